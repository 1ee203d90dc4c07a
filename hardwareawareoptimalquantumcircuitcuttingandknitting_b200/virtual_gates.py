"""Virtual (cut) gates: quasi-probability decompositions and their knit rules.

Host-side mirror of the reference operator API in
``third_party/qvm/qvm/virtual_gates.py``:

* ``VirtualBinaryGate`` (``:17-55``) with ``num_instantiations``,
  ``_instantiations()``, ``instantiate(id)`` and ``knit(results, clbit_idx)``;
* the families ``VirtualMove`` (``:58-124``, wire cut, 8 instantiations),
  ``VirtualCZ`` (``:153-194``), ``VirtualCX`` (``:197-206``), ``VirtualCY``
  (``:209-220``), ``VirtualRZZ`` (``:226-286``), ``VirtualCPhase`` (``:294-310``);
* ``VirtualGateEndpoint`` (``:127-150``), ``WireCut`` (``:9-14``) and the registry
  ``VIRTUAL_GATE_TYPES`` (``:313-319``).

Here the decompositions are *data* (``_TABLES``): each instantiation is a list of
``(gate name, params, qubit)`` / ``("measure", (), qubit)`` entries, checked
entry-for-entry against tables dumped from the reference classes
(``tests/golden/instantiation_tables.json``).  The device never sees gate names:
``compiler.py`` turns every (slot, variant) into ``pre-matrix, measure?,
post-matrix`` records, and the knit rules become the coefficient pairs returned
by ``knit_coefficients()`` which the knit kernels consume.

``VirtualCPhase`` reproduces the reference's angles as written (``rz(-theta/4)``
around an RZZ decomposition at ``-theta/2``) including the fact that it is *not*
numerically a CP(theta) decomposition.  Parity means
"same as the reference", not "fixed".
"""
from __future__ import annotations

import abc
from math import cos, pi, sin
from typing import Sequence

from .circuit import Barrier, Gate, QuantumCircuit, QuantumRegister

__all__ = [
    "WireCut", "VirtualBinaryGate", "VirtualMove", "VirtualGateEndpoint", "VirtualCZ",
    "VirtualCX", "VirtualCY", "VirtualRZZ", "VirtualCPhase", "VIRTUAL_GATE_TYPES",
    "RZZ_ACCURACY",
]

RZZ_ACCURACY = 0.00001      # virtual_gates.py:223

_M = "measure"
# entries: (name, params, qubit).  Order inside one instantiation = order in which the
# reference appends to its QuantumCircuit(2, 1); only the per-qubit order matters.
_CZ_TABLE = (
    (("sdg", (), 0), ("sdg", (), 1)),
    (("s", (), 0), ("s", (), 1)),
    ((_M, (), 0),),
    ((_M, (), 0), ("z", (), 1)),
    ((_M, (), 1),),
    (("z", (), 0), (_M, (), 1)),
)
_MOVE_TABLE = (
    (),
    (("x", (), 1),),
    (("h", (), 0), (_M, (), 0), ("h", (), 1)),
    (("h", (), 0), (_M, (), 0), ("x", (), 1), ("h", (), 1)),
    (("sdg", (), 0), ("h", (), 0), (_M, (), 0), ("h", (), 1), ("s", (), 1)),
    (("sdg", (), 0), ("h", (), 0), (_M, (), 0), ("x", (), 1), ("h", (), 1), ("s", (), 1)),
    ((_M, (), 0),),
    ((_M, (), 0), ("x", (), 1)),
)


def _wrap(table, before, after):
    """Sandwich every instantiation between two fixed op lists (compose order)."""
    return tuple(tuple(before) + tuple(inst) + tuple(after) for inst in table)


def _rzz_table(theta: float):
    """``VirtualRZZ._instantiations`` (``virtual_gates.py:230-260``); ``m = -theta``."""
    m = -theta
    inst0 = ()
    inst1 = (("z", (), 0), ("z", (), 1))
    if abs(cos(m / 2)) < RZZ_ACCURACY:
        return (inst1,)
    if abs(sin(m / 2)) < RZZ_ACCURACY:
        return (inst0,)
    return (
        inst0,
        inst1,
        (("rz", (-pi / 2,), 0), (_M, (), 1)),
        ((_M, (), 0), ("rz", (-pi / 2,), 1)),
        (("rz", (pi / 2,), 0), (_M, (), 1)),
        ((_M, (), 0), ("rz", (pi / 2,), 1)),
    )


def _table_to_circuits(table) -> list[QuantumCircuit]:
    out = []
    for inst in table:
        qc = QuantumCircuit(2, 1)
        for name, params, q in inst:
            if name == _M:
                qc.measure(q, 0)
            else:
                qc.append(Gate(name, 1, params), [q])
        out.append(qc)
    return out


class WireCut(Barrier):
    """Marker left on a wire where it is cut (``virtual_gates.py:9-14``)."""

    def __init__(self, num_qubits: int = 1, label: str | None = None) -> None:
        super().__init__(num_qubits, label)
        self.name = "wire_cut"


class VirtualBinaryGate(Barrier, abc.ABC):
    """Two-qubit gate replaced by a sum over local instantiations."""

    def __init__(self, original_gate: Gate, label: str = "") -> None:
        self._original_gate = original_gate
        super().__init__(original_gate.num_qubits, label if label else f"v_{original_gate.name}")
        self.name = f"v_{original_gate.name}"
        # NB: shares the list with the original gate, as the reference does
        # (virtual_gates.py:22) - VirtualCPhase relies on it.
        self.params = original_gate.params
        for inst in self._instantiations():
            self._check_instantiation(inst)

    @property
    def original_gate(self) -> Gate:
        return self._original_gate

    @property
    def num_instantiations(self) -> int:
        return len(self._table())

    @abc.abstractmethod
    def _table(self) -> tuple:
        """Instantiations as data: tuple of tuples of ``(name, params, qubit)``."""

    @abc.abstractmethod
    def knit_coefficients(self) -> list[tuple[float, float]]:
        """``[(a_i, b_i)]`` such that the knitted distribution is
        ``sum_i a_i * r_i[bit=0] + b_i * r_i[bit=1]``."""

    def _instantiations(self) -> list[QuantumCircuit]:
        return _table_to_circuits(self._table())

    def instantiate(self, inst_id: int) -> QuantumCircuit:
        return self._instantiations()[inst_id]

    def knit(self, results: Sequence, clbit_idx: int):
        """Combine the ``num_instantiations`` distributions of one chunk.

        ``results`` are device-resident ``QuasiDistr`` objects.  With pruning
        disabled (``accuracy == 0``) this is one fused kernel over the coefficient
        pairs; with pruning it replays the reference's operation order
        (split, then the signed chain, pruning after every step) on the device.
        """
        from .quasi_distr import knit_level
        if len(results) != self.num_instantiations:
            raise ValueError(f"{self.name}: expected {self.num_instantiations} results, got {len(results)}")
        return knit_level(self, list(results), clbit_idx)

    @staticmethod
    def _check_instantiation(inst: QuantumCircuit) -> None:
        assert len(inst.qubits) == 2
        assert len(inst.clbits) == 1
        for instr in inst.data:
            assert len(instr.qubits) == 1
            assert len(instr.clbits) <= 1

    #: how the reference evaluates knit(): "chain" = 0.5*((r00-r01)+(r10-r11)+...-...)
    #: left to right; "rzz" = the cos/sin form (virtual_gates.py:262-286)
    knit_form = "chain"


class VirtualMove(VirtualBinaryGate):
    """Wire cut as a virtual SWAP onto a fresh qubit (``virtual_gates.py:58-124``)."""

    def __init__(self, originalGate: Gate) -> None:
        super().__init__(originalGate, label=f"VirtualMove {originalGate.label}")

    def _table(self):
        return _MOVE_TABLE

    def knit_coefficients(self):
        signs = (+1, +1, +1, -1, +1, -1, +1, -1)
        return [(0.5 * s, -0.5 * s) for s in signs]


class VirtualGateEndpoint(Barrier):
    """One end of a virtual gate inside a fragment (``virtual_gates.py:127-150``)."""

    def __init__(self, virtual_gate: VirtualBinaryGate, vgate_idx: int, qubit_idx: int) -> None:
        self._virtual_gate = virtual_gate
        self.vgate_idx = vgate_idx
        self.qubit_idx = qubit_idx
        super().__init__(1, label=f"v_{virtual_gate.name}_{vgate_idx}_{qubit_idx}")
        self.name = "vgate_endpoint"

    @property
    def virtual_gate(self) -> VirtualBinaryGate:
        return self._virtual_gate

    def instantiate(self, inst_id: int) -> QuantumCircuit:
        """One-qubit, <=1-clbit circuit holding this end's share of instantiation
        ``inst_id`` (the reference returns it as an ``Instruction``)."""
        assert 0 <= inst_id < self._virtual_gate.num_instantiations
        return self._circuit_on_index(self._virtual_gate.instantiate(inst_id), self.qubit_idx)

    @staticmethod
    def _circuit_on_index(circuit: QuantumCircuit, index: int) -> QuantumCircuit:
        new_circuit = QuantumCircuit(QuantumRegister(1, "q"), *circuit.cregs)
        qubit = circuit.qubits[index]
        for instr in circuit.data:
            if len(instr.qubits) == 1 and instr.qubits[0] == qubit:
                new_circuit.append(instr.operation, (new_circuit.qubits[0],), instr.clbits)
        return new_circuit


class VirtualCZ(VirtualBinaryGate):
    def _table(self):
        return _CZ_TABLE

    def knit_coefficients(self):
        signs = (+1, +1, +1, -1, +1, -1)
        return [(0.5 * s, -0.5 * s) for s in signs]


_CX_TABLE = _wrap(_CZ_TABLE, (("h", (), 1),), (("h", (), 1),))
_CY_TABLE = _wrap(_CX_TABLE, (("rz", (-pi / 2,), 1),), (("rz", (pi / 2,), 1),))


class VirtualCX(VirtualCZ):
    def _table(self):
        return _CX_TABLE


class VirtualCY(VirtualCX):
    def _table(self):
        return _CY_TABLE


class VirtualRZZ(VirtualBinaryGate):
    knit_form = "rzz"

    def __init__(self, original_gate: Gate, label: str = "") -> None:
        super().__init__(original_gate, label)

    def _table(self):
        return _rzz_table(self.params[0])

    def knit_coefficients(self):
        m = -self.params[0]
        c, s = cos(m / 2), sin(m / 2)
        if abs(c) < RZZ_ACCURACY:
            return [(s ** 2, 0.0)]
        if abs(s) < RZZ_ACCURACY:
            return [(c ** 2, 0.0)]
        cs = c * s
        return [(c ** 2, 0.0), (s ** 2, 0.0), (cs, -cs), (cs, -cs), (-cs, cs), (-cs, cs)]


class VirtualCPhase(VirtualRZZ):
    def __init__(self, original_gate: Gate, label: str = "") -> None:
        super().__init__(original_gate, label)
        self.params[0] = -self.params[0] / 2        # in place, aliasing the gate (virtual_gates.py:297)

    def _table(self):
        lam = self.params[0]
        return _wrap(super()._table(), (("rz", (lam / 2,), 0),), (("rz", (lam / 2,), 1),))


VIRTUAL_GATE_TYPES: dict[str, type[VirtualBinaryGate]] = {
    "cx": VirtualCX,
    "cy": VirtualCY,
    "cz": VirtualCZ,
    "rzz": VirtualRZZ,
    "cp": VirtualCPhase,
}
