"""Adapters between Qiskit objects and this package's Qiskit-free IR.

The reference hands qvm a ``qiskit.QuantumCircuit`` whose cut gates are instances
of ``qvm.virtual_gates.Virtual*`` (``src/HwAwareCutter/Cutter.py:575-643``).
``circuit_from_qiskit`` walks such a circuit by duck typing only - ``qregs``,
``cregs``, ``data`` entries with ``operation`` (``name``, ``params``, ``num_qubits``,
optional ``to_matrix()``), ``qubits``, ``clbits`` - so it needs neither qiskit nor the
reference to be importable.  Virtual gates are recognised by class name
(``VirtualMove``, ``VirtualCX`` ...) and rebuilt from their ``original_gate`` /
``_original_gate`` (name + params) with this package's classes, which carry the same
instantiation tables (``tests/golden/instantiation_tables.json``).

Qiskit is not installed in the build image: the adapter is exercised against
duck-typed fakes and against the reference's own ``Virtual*`` classes loaded under a
stub (``tests/test_adapter.py``); it has NOT been run against a real qiskit install
(SURVEY.md 8f-2).
"""
from __future__ import annotations

import numpy as np

from .circuit import (Barrier, ClassicalRegister, Gate, GATE_NUM_QUBITS, Measure, QuantumCircuit,
                      QuantumRegister)
from .virtual_gates import VIRTUAL_GATE_TYPES, VirtualMove, WireCut

__all__ = ["circuit_from_qiskit", "VIRTUAL_CLASS_KINDS"]

VIRTUAL_CLASS_KINDS = {"VirtualMove": "move", "VirtualCX": "cx", "VirtualCZ": "cz", "VirtualCY": "cy",
                       "VirtualRZZ": "rzz", "VirtualCPhase": "cp"}
_RENAME = {"u1": "p", "cnot": "cx", "cphase": "cp", "cu1": "cp", "i": "id"}


def _original_params(op, kind: str) -> list[float]:
    """Parameters of the gate the virtual gate replaces.  The reference's VirtualCPhase has
    already overwritten params[0] with -theta/2 in place (virtual_gates.py:297): undo it."""
    orig = getattr(op, "original_gate", None) or getattr(op, "_original_gate", None)
    params = [float(p) for p in (getattr(orig, "params", None) or getattr(op, "_params", None) or [])]
    if kind == "cp" and params:
        own = getattr(op, "_params", None) or getattr(op, "params", None)
        if own is not None and float(own[0]) == params[0]:      # aliased list: already -theta/2
            params = [-2.0 * params[0]] + params[1:]
    return params


def circuit_from_qiskit(qc) -> QuantumCircuit:
    qmap, cmap = {}, {}
    qregs, cregs = [], []
    for reg in qc.qregs:
        new = QuantumRegister(len(reg), getattr(reg, "name", None))
        qregs.append(new)
        for i, bit in enumerate(reg):
            qmap[bit] = new[i]
    for reg in qc.cregs:
        new = ClassicalRegister(len(reg), getattr(reg, "name", None))
        cregs.append(new)
        for i, bit in enumerate(reg):
            cmap[bit] = new[i]
    out = QuantumCircuit(*qregs, *cregs, name=getattr(qc, "name", None))
    for ins in qc.data:
        op = getattr(ins, "operation", None)
        if op is None:                      # old tuple form (operation, qubits, clbits)
            op, qubits, clbits = ins
        else:
            qubits, clbits = ins.qubits, ins.clbits
        qs = [qmap[q] for q in qubits]
        cs = [cmap[c] for c in clbits]
        cls = type(op).__name__
        name = _RENAME.get(op.name, op.name)
        if cls in VIRTUAL_CLASS_KINDS:
            kind = VIRTUAL_CLASS_KINDS[cls]
            if kind == "move":
                orig = getattr(op, "original_gate", None) or getattr(op, "_original_gate", None)
                out.append(VirtualMove(Gate("swap", 2, (), label=getattr(orig, "label", None))), qs)
            else:
                out.append(VIRTUAL_GATE_TYPES[kind](Gate(kind, 2, _original_params(op, kind)),
                                                    getattr(op, "label", "") or ""), qs)
        elif cls == "WireCut":
            out.append(WireCut(1, getattr(op, "label", None)), qs)
        elif name == "barrier" or cls == "Barrier":
            out.append(Barrier(len(qs), getattr(op, "label", None)), qs)
        elif name == "measure":
            out.append(Measure(), qs, cs)
        elif name in GATE_NUM_QUBITS:
            out.append(Gate(name, len(qs), [float(p) for p in getattr(op, "params", [])]), qs)
        elif hasattr(op, "to_matrix") and len(qs) <= 2:
            # qiskit's to_matrix() is little-endian in the gate's own qubit order: same as ours
            out.append(Gate(name, len(qs), (), matrix=np.asarray(op.to_matrix(), dtype=np.complex128)), qs)
        else:
            raise NotImplementedError(f"cannot convert operation '{op.name}' ({cls}) on {len(qs)} qubits; "
                                      "decompose the circuit to one- and two-qubit gates first")
    return out
