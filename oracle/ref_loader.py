"""Load the REAL reference knit code under a qiskit stub (TEST INFRASTRUCTURE).

Build-container only: ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may call this.  It is used by
``tests/golden/make_golden.py`` to produce the committed fixtures and by the
``not gpu`` tests *when the reference tree is present* to re-check the oracle
live.

``third_party/qvm/qvm/quasi_distr.py`` imports only ``typing`` and loads as is.
``third_party/qvm/qvm/virtual_gates.py`` needs ``qiskit.circuit.{Barrier, Gate,
QuantumCircuit, Instruction, QuantumRegister}``; the stub below records the
handful of calls those classes make (``x h z s sdg rz measure compose``) so
that every ``_instantiations()`` table and every ``knit()`` runs unmodified
(SURVEY.md C.2).  No reference source is copied: the modules are imported from
where they lie.
"""
import importlib
import os
import sys
import types

REFERENCE_QVM = "/root/reference/third_party/qvm"


def available():
    return os.path.isfile(os.path.join(REFERENCE_QVM, "qvm", "virtual_gates.py"))


class _Op:
    def __init__(self, name, qubits, clbits=(), params=()):
        self.name, self.qubits, self.clbits, self.params = name, tuple(qubits), tuple(clbits), list(params)


class _Entry:
    """One element of ``QuantumCircuit.data``: (operation, qubits, clbits)."""

    def __init__(self, op):
        self.operation, self.qubits, self.clbits = op, op.qubits, op.clbits


class _StubCircuit:
    def __init__(self, nq=0, nc=0):
        self.qubits = list(range(nq))
        self.clbits = list(range(nc))
        self.cregs = []
        self.data = []

    def _add(self, name, q, clbits=(), params=()):
        self.data.append(_Entry(_Op(name, (q,), clbits, params)))

    def x(self, q): self._add("x", q)
    def h(self, q): self._add("h", q)
    def z(self, q): self._add("z", q)
    def s(self, q): self._add("s", q)
    def sdg(self, q): self._add("sdg", q)
    def rz(self, theta, q): self._add("rz", q, (), (theta,))
    def measure(self, q, c): self._add("measure", q, (c,))

    def compose(self, other, inplace=False):
        new = _StubCircuit(len(self.qubits), len(self.clbits))
        new.data = list(self.data) + list(other.data)
        return new

    def append(self, *a, **k):  # only used by _define(), never on the knit path
        pass


class _StubBarrier:
    def __init__(self, num_qubits=1, label=None):
        self.num_qubits, self.label = num_qubits, label


class StubGate:
    def __init__(self, name, num_qubits, params, label=None):
        self.name, self.num_qubits, self.params, self.label = name, num_qubits, list(params), label


_loaded = None


def load():
    """-> (virtual_gates module, quasi_distr module) of the reference, unmodified."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present (expected on the GPU box)")
    saved = {k: sys.modules.get(k) for k in ("qiskit", "qiskit.circuit", "qvm", "qvm.quasi_distr",
                                             "qvm.virtual_gates")}
    qk, qc = types.ModuleType("qiskit"), types.ModuleType("qiskit.circuit")
    for name, obj in dict(Barrier=_StubBarrier, Gate=StubGate, QuantumCircuit=_StubCircuit,
                          Instruction=object, QuantumRegister=object).items():
        setattr(qc, name, obj)
    qk.circuit = qc
    sys.modules["qiskit"], sys.modules["qiskit.circuit"] = qk, qc
    sys.path.insert(0, REFERENCE_QVM)
    try:
        for k in ("qvm", "qvm.quasi_distr", "qvm.virtual_gates"):
            sys.modules.pop(k, None)
        vg = importlib.import_module("qvm.virtual_gates")
        qd = importlib.import_module("qvm.quasi_distr")
    finally:
        sys.path.remove(REFERENCE_QVM)
        for k in ("qiskit", "qiskit.circuit"):
            if saved[k] is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = saved[k]
    _loaded = (vg, qd)
    return _loaded


def make_vgate(vg, kind, theta=None):
    """Instantiate the reference class for ``kind`` the way the cutter does
    (``src/HwAwareCutter/Cutter.py:589`` gate cuts, ``:629`` wire cuts)."""
    if kind == "move":
        return vg.VirtualMove(StubGate("swap", 2, [], label="wc"))
    if kind in ("rzz", "cp"):
        return vg.VIRTUAL_GATE_TYPES[kind](StubGate(kind, 2, [theta]), f"{kind} cut")
    return vg.VIRTUAL_GATE_TYPES[kind](StubGate(kind, 2, []), f"{kind} cut")


def dump_table(gate):
    """[[(name, qubit, has_clbit, params), ...] per instantiation] straight from the reference."""
    out = []
    for inst in gate._instantiations():
        out.append([[e.operation.name, int(e.qubits[0]), len(e.clbits), [float(p) for p in e.operation.params]]
                    for e in inst.data])
    return out
