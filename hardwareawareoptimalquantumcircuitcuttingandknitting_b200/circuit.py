"""Qiskit-free circuit IR for the fragment-execution + knit hot path.

The reference builds everything on ``qiskit.circuit.QuantumCircuit``
(``third_party/qvm/qvm/virtual_circuit.py:4-9``).  Qiskit is not part of this
image, and the hot path only ever needs: registers (a fragment *is* a quantum
register, ``virtual_circuit.py:8``), an ordered instruction list, one- and
two-qubit unitaries given as matrices, measurements into classical bits, and
barriers (the virtual gates subclass ``Barrier``, ``virtual_gates.py:17``).
This module provides exactly that, with Qiskit's conventions:

* little-endian: qubit 0 is the least-significant bit of a basis-state index,
* a result key is an integer whose bit ``i`` is classical bit ``i``
  (``quasi_distr.py:13-20`` parses Qiskit's MSB-first strings into such ints),
* gate matrices follow Qiskit's standard-gate definitions (u, rx, ry, rz, p, r,
  cx with the *control first*, cp, rzz ...).

Nothing here touches the GPU; the compiler (``compiler.py``) flattens these
objects into the POD arrays the C ABI takes.
"""
from __future__ import annotations

import cmath
import math
from dataclasses import dataclass, field
from typing import Iterable, Iterator, Sequence

import numpy as np

__all__ = [
    "QuantumRegister", "ClassicalRegister", "Qubit", "Clbit", "Operation", "Gate",
    "Measure", "Barrier", "CircuitInstruction", "QuantumCircuit", "gate_matrix",
    "GATE_NUM_QUBITS",
]


# --------------------------------------------------------------------------- bits / registers
# Bits are created once, by their register, and compared by identity (eq=False keeps object.__hash__: the
# host compiler hashes qubits tens of thousands of times per circuit).
@dataclass(frozen=True, eq=False)
class Qubit:
    register: "QuantumRegister"
    index: int

    def __repr__(self) -> str:  # pragma: no cover - cosmetic
        return f"{self.register.name}[{self.index}]"


@dataclass(frozen=True, eq=False)
class Clbit:
    register: "ClassicalRegister"
    index: int

    def __repr__(self) -> str:  # pragma: no cover - cosmetic
        return f"{self.register.name}[{self.index}]"


class _Register:
    _bit_type: type = Qubit
    _counter = 0

    def __init__(self, size: int, name: str | None = None) -> None:
        if size < 0:
            raise ValueError("register size must be >= 0")
        if name is None:
            type(self)._counter += 1
            name = f"{'q' if self._bit_type is Qubit else 'c'}{type(self)._counter}"
        self.size = int(size)
        self.name = name
        self._bits = tuple(self._bit_type(self, i) for i in range(self.size))

    def __len__(self) -> int:
        return self.size

    def __iter__(self) -> Iterator:
        return iter(self._bits)

    def __getitem__(self, i):
        return self._bits[i]

    # identity semantics: two registers with the same name are still two registers
    def __hash__(self) -> int:
        return id(self)

    def __eq__(self, other) -> bool:
        return self is other

    def __repr__(self) -> str:  # pragma: no cover - cosmetic
        return f"{type(self).__name__}({self.size}, '{self.name}')"


class QuantumRegister(_Register):
    """A fragment is a ``QuantumRegister`` (alias ``Fragment`` in the reference)."""
    _bit_type = Qubit


class ClassicalRegister(_Register):
    _bit_type = Clbit


# --------------------------------------------------------------------------- operations
class Operation:
    """Base of everything that can sit in a circuit."""
    name: str = "op"
    num_qubits: int = 1
    num_clbits: int = 0
    params: list

    def __repr__(self) -> str:  # pragma: no cover - cosmetic
        return f"{type(self).__name__}({self.name}, {getattr(self, 'params', [])})"


class Gate(Operation):
    """A named unitary on 1 or 2 qubits.  ``matrix`` overrides the name lookup."""

    def __init__(self, name: str, num_qubits: int, params: Sequence[float] = (),
                 label: str | None = None, matrix: np.ndarray | None = None) -> None:
        self.name = name
        self.num_qubits = int(num_qubits)
        self.params = [float(p) for p in params]
        self.label = label
        self._matrix = None if matrix is None else np.asarray(matrix, dtype=np.complex128)

    def to_matrix(self) -> np.ndarray:
        if self._matrix is not None:
            return self._matrix
        return gate_matrix(self.name, self.params)


class Measure(Operation):
    name = "measure"
    num_qubits = 1
    num_clbits = 1

    def __init__(self) -> None:
        self.params = []


class Barrier(Operation):
    """Directive without effect on the state (same role as qiskit's ``Barrier``)."""
    name = "barrier"

    def __init__(self, num_qubits: int = 1, label: str | None = None) -> None:
        self.num_qubits = int(num_qubits)
        self.label = label
        self.params = []


@dataclass
class CircuitInstruction:
    operation: Operation
    qubits: tuple = ()
    clbits: tuple = ()

    def __iter__(self):  # allows ``op, qubits, clbits = instr``
        return iter((self.operation, self.qubits, self.clbits))


# --------------------------------------------------------------------------- gate matrices
_SQ2 = 1.0 / math.sqrt(2.0)


def _u(theta: float, phi: float, lam: float) -> np.ndarray:
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -cmath.exp(1j * lam) * s],
                     [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]], dtype=np.complex128)


def gate_matrix(name: str, params: Sequence[float] = ()) -> np.ndarray:
    """Qiskit standard-gate matrices (little-endian for two-qubit gates: the
    first qubit argument is the less-significant index bit)."""
    p = list(params)
    if name in ("id", "i"):
        return np.eye(2, dtype=np.complex128)
    if name == "x":
        return np.array([[0, 1], [1, 0]], dtype=np.complex128)
    if name == "y":
        return np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
    if name == "z":
        return np.array([[1, 0], [0, -1]], dtype=np.complex128)
    if name == "h":
        return np.array([[_SQ2, _SQ2], [_SQ2, -_SQ2]], dtype=np.complex128)
    if name == "s":
        return np.array([[1, 0], [0, 1j]], dtype=np.complex128)
    if name == "sdg":
        return np.array([[1, 0], [0, -1j]], dtype=np.complex128)
    if name == "t":
        return np.array([[1, 0], [0, cmath.exp(1j * math.pi / 4)]], dtype=np.complex128)
    if name == "tdg":
        return np.array([[1, 0], [0, cmath.exp(-1j * math.pi / 4)]], dtype=np.complex128)
    if name == "sx":
        return 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=np.complex128)
    if name == "rx":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
    if name == "ry":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if name == "rz":
        return np.array([[cmath.exp(-0.5j * p[0]), 0], [0, cmath.exp(0.5j * p[0])]], dtype=np.complex128)
    if name in ("p", "u1"):
        return np.array([[1, 0], [0, cmath.exp(1j * p[0])]], dtype=np.complex128)
    if name == "r":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -1j * cmath.exp(-1j * p[1]) * s],
                         [-1j * cmath.exp(1j * p[1]) * s, c]], dtype=np.complex128)
    if name in ("u", "u3"):
        return _u(p[0], p[1], p[2])
    if name == "u2":
        return _u(math.pi / 2, p[0], p[1])
    # ---- two-qubit gates; index = q0 + 2*q1 with (q0, q1) the argument order
    if name == "cx":      # control = first arg (bit 0), target = second (bit 1)
        m = np.zeros((4, 4), dtype=np.complex128)
        m[0, 0] = m[2, 2] = 1      # control 0: identity
        m[3, 1] = m[1, 3] = 1      # control 1: flip bit 1
        return m
    if name == "cy":
        m = np.zeros((4, 4), dtype=np.complex128)
        m[0, 0] = m[2, 2] = 1
        m[3, 1] = 1j
        m[1, 3] = -1j
        return m
    if name == "cz":
        return np.diag([1, 1, 1, -1]).astype(np.complex128)
    if name == "cp":
        return np.diag([1, 1, 1, cmath.exp(1j * p[0])]).astype(np.complex128)
    if name == "rzz":
        a, b = cmath.exp(-0.5j * p[0]), cmath.exp(0.5j * p[0])
        return np.diag([a, b, b, a]).astype(np.complex128)
    if name == "swap":
        m = np.zeros((4, 4), dtype=np.complex128)
        m[0, 0] = m[3, 3] = m[1, 2] = m[2, 1] = 1
        return m
    raise ValueError(f"unknown gate '{name}'")


GATE_NUM_QUBITS = {
    **{g: 1 for g in ("id", "i", "x", "y", "z", "h", "s", "sdg", "t", "tdg", "sx", "rx", "ry", "rz",
                      "p", "u1", "r", "u", "u3", "u2")},
    **{g: 2 for g in ("cx", "cy", "cz", "cp", "rzz", "swap")},
}


# --------------------------------------------------------------------------- circuit
class QuantumCircuit:
    """Ordered instruction list over quantum and classical registers.

    Constructor mirrors the two qiskit forms the reference uses:
    ``QuantumCircuit(nq, nc)`` (``virtual_gates.py:65``) and
    ``QuantumCircuit(*qregs, *cregs)`` (``virtual_circuit.py:99,119``).
    """

    def __init__(self, *regs, name: str | None = None) -> None:
        self.qregs: list[QuantumRegister] = []
        self.cregs: list[ClassicalRegister] = []
        self.data: list[CircuitInstruction] = []
        self.name = name or "circuit"
        if regs and all(isinstance(r, int) for r in regs):
            if len(regs) > 2:
                raise ValueError("QuantumCircuit(nq[, nc]) takes at most two ints")
            if regs[0]:
                self.add_register(QuantumRegister(regs[0], "q"))
            if len(regs) == 2 and regs[1]:
                self.add_register(ClassicalRegister(regs[1], "c"))
        else:
            for r in regs:
                self.add_register(r)

    # ---- registers
    def add_register(self, reg) -> None:
        if isinstance(reg, QuantumRegister):
            if reg not in self.qregs:
                self.qregs.append(reg)
        elif isinstance(reg, ClassicalRegister):
            if reg not in self.cregs:
                self.cregs.append(reg)
        else:
            raise TypeError(f"not a register: {reg!r}")

    @property
    def qubits(self) -> list[Qubit]:
        return [q for r in self.qregs for q in r]

    @property
    def clbits(self) -> list[Clbit]:
        return [c for r in self.cregs for c in r]

    @property
    def num_qubits(self) -> int:
        return sum(len(r) for r in self.qregs)

    @property
    def num_clbits(self) -> int:
        return sum(len(r) for r in self.cregs)

    def qubit_index(self, q: Qubit) -> int:
        off = 0
        for r in self.qregs:
            if q.register is r:
                return off + q.index
            off += len(r)
        raise ValueError(f"qubit {q} not in circuit")

    def clbit_index(self, c: Clbit) -> int:
        off = 0
        for r in self.cregs:
            if c.register is r:
                return off + c.index
            off += len(r)
        raise ValueError(f"clbit {c} not in circuit")

    # ---- construction
    def _q(self, q) -> Qubit:
        return q if isinstance(q, Qubit) else self.qubits[q]

    def _c(self, c) -> Clbit:
        return c if isinstance(c, Clbit) else self.clbits[c]

    def append(self, op: Operation, qubits: Iterable = (), clbits: Iterable = ()) -> None:
        qs = tuple(self._q(q) for q in qubits)
        cs = tuple(self._c(c) for c in clbits)
        self.data.append(CircuitInstruction(op, qs, cs))

    def _g(self, name, qubits, params=()):
        self.append(Gate(name, len(qubits), params), qubits)

    def x(self, q): self._g("x", [q])
    def y(self, q): self._g("y", [q])
    def z(self, q): self._g("z", [q])
    def h(self, q): self._g("h", [q])
    def s(self, q): self._g("s", [q])
    def sdg(self, q): self._g("sdg", [q])
    def t(self, q): self._g("t", [q])
    def tdg(self, q): self._g("tdg", [q])
    def sx(self, q): self._g("sx", [q])
    def id(self, q): self._g("id", [q])
    def rx(self, theta, q): self._g("rx", [q], [theta])
    def ry(self, theta, q): self._g("ry", [q], [theta])
    def rz(self, theta, q): self._g("rz", [q], [theta])
    def p(self, lam, q): self._g("p", [q], [lam])
    def r(self, theta, phi, q): self._g("r", [q], [theta, phi])
    def u(self, theta, phi, lam, q): self._g("u", [q], [theta, phi, lam])
    def cx(self, c, t): self._g("cx", [c, t])
    def cy(self, c, t): self._g("cy", [c, t])
    def cz(self, a, b): self._g("cz", [a, b])
    def cp(self, lam, a, b): self._g("cp", [a, b], [lam])
    def rzz(self, theta, a, b): self._g("rzz", [a, b], [theta])
    def swap(self, a, b): self._g("swap", [a, b])

    def unitary(self, matrix, qubits, label: str = "unitary") -> None:
        m = np.asarray(matrix, dtype=np.complex128)
        nq = len(qubits)
        if m.shape != (1 << nq, 1 << nq):
            raise ValueError("matrix shape does not match qubit count")
        self.append(Gate(label, nq, (), matrix=m), qubits)

    def barrier(self, *qubits) -> None:
        qs = list(qubits) if qubits else self.qubits
        self.append(Barrier(len(qs)), qs)

    def measure(self, q, c) -> None:
        self.append(Measure(), [q], [c])

    def measure_all(self) -> None:
        """Qiskit semantics: add a fresh creg ``meas`` with one bit per qubit,
        a barrier, then ``measure(qubit i -> meas[i])``
        (used by every generator in ``benchmarks/helper_functions.py:132-203``)."""
        creg = ClassicalRegister(self.num_qubits, "meas")
        self.add_register(creg)
        self.barrier()
        for i, q in enumerate(self.qubits):
            self.append(Measure(), [q], [creg[i]])

    def copy(self) -> "QuantumCircuit":
        new = QuantumCircuit(*self.qregs, *self.cregs, name=self.name)
        new.data = [CircuitInstruction(i.operation, i.qubits, i.clbits) for i in self.data]
        return new

    def __iter__(self) -> Iterator[CircuitInstruction]:
        return iter(self.data)

    def __len__(self) -> int:
        return len(self.data)

    def count_ops(self) -> dict[str, int]:
        out: dict[str, int] = {}
        for ins in self.data:
            out[ins.operation.name] = out.get(ins.operation.name, 0) + 1
        return out

    def decompose_two_qubit(self) -> "QuantumCircuit":
        """One level of the Qiskit definitions the cutter relies on
        (``Cutter.py:84`` calls ``inputCirc.decompose()``): after it the only
        two-qubit gate left is ``cx``.  One-qubit gates keep
        their names - their matrices are what matters to the simulator."""
        new = QuantumCircuit(*self.qregs, *self.cregs, name=self.name)
        for ins in self.data:
            op, qs, cs = ins.operation, ins.qubits, ins.clbits
            if isinstance(op, Gate) and op._matrix is None and op.num_qubits == 2:
                a, b = qs
                if op.name == "cz":
                    new.h(b); new.cx(a, b); new.h(b)
                    continue
                if op.name == "cp":
                    lam = op.params[0]
                    new.p(lam / 2, a); new.cx(a, b); new.p(-lam / 2, b); new.cx(a, b); new.p(lam / 2, b)
                    continue
                if op.name == "swap":
                    new.cx(a, b); new.cx(b, a); new.cx(a, b)
                    continue
                if op.name == "cy":
                    new.sdg(b); new.cx(a, b); new.s(b)
                    continue
                if op.name == "rzz":
                    new.cx(a, b); new.rz(op.params[0], b); new.cx(a, b)
                    continue
            new.data.append(CircuitInstruction(op, qs, cs))
        return new
