"""B200Backend: the fragment-execution plug-in.

The reference binds every fragment to a Qiskit backend and only relies on the
duck type ``backend.run(list_of_circuits, shots=int) -> job`` with
``job.result().get_counts() -> dict | list[dict]`` whose keys are MSB-first
bitstrings with a space between classical registers, the last-added register
leftmost (``third_party/qvm/qvm/run.py:42,48-56``, ``virtual_circuit.py:82-95``,
``quasi_distr.py:13-20``).  ``B200Backend`` honours that duck type - each circuit
is compiled as a label-free program and its *exact* outcome distribution is
returned as float "counts" that sum to ``shots`` (``from_counts`` divides by the
total, ``quasi_distr.py:14``) - so it can also stand in for ``AerSimulator()`` in
``getCircResultFromBackend`` (``src/HwAwareCutter/Utilities.py:39-69``).

``run_virtual_circuit`` recognises this class and bypasses circuit objects,
strings and dicts entirely (``VirtualCircuit.simulate_fragments``).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .circuit import QuantumCircuit, QuantumRegister
from .compiler import FragmentExecutor, FragmentProgram
from .quasi_distr import default_device

__all__ = ["B200Backend", "B200Job", "B200Result"]


class B200Result:
    def __init__(self, counts: list[dict]) -> None:
        self._counts = counts

    def get_counts(self):
        if any(len(c) == 0 for c in self._counts):
            raise ValueError("No counts for experiment (circuit has no measurement)")
        return self._counts[0] if len(self._counts) == 1 else self._counts


class B200Job:
    def __init__(self, result: B200Result) -> None:
        self._result = result

    def result(self) -> B200Result:
        return self._result


class B200Backend:
    name = "b200_statevector"

    def __init__(self, device=None) -> None:
        self._device = device

    def _format_key(self, key: int, cregs) -> str:
        parts, off = [], 0
        for reg in cregs:
            bits = "".join("1" if (key >> (off + i)) & 1 else "0" for i in reversed(range(len(reg))))
            parts.append(bits)
            off += len(reg)
        return " ".join(reversed(parts))

    def exact_distribution(self, circuit: QuantumCircuit) -> dict[int, float]:
        """Exact outcome distribution of one measurement-terminated circuit (key bit i = clbit i)."""
        device = self._device if self._device is not None else default_device()
        handle = _lib.get_handle(getattr(device, "index", None) or 0)
        whole = QuantumRegister(circuit.num_qubits, "all")
        remap = {q: whole[i] for i, q in enumerate(circuit.qubits)}
        flat = QuantumCircuit(whole, *circuit.cregs)
        for ins in circuit.data:
            flat.append(ins.operation, [remap[q] for q in ins.qubits], ins.clbits)
        prog = FragmentProgram(flat, whole, circuit.num_clbits)
        if not prog.measures_anything:
            return {}
        row = FragmentExecutor(prog, device).run(handle)[0].cpu().numpy()
        clbits = prog.out_clbits                         # every written clbit, ascending (mid-circuit ones too)
        out = {}
        for i in np.nonzero(row)[0]:
            key = 0
            for j, c in enumerate(clbits):
                key |= ((int(i) >> j) & 1) << c
            out[key] = float(row[i])
        return out

    def run(self, circuits, shots: int = 1024, **_ignored) -> B200Job:
        if isinstance(circuits, QuantumCircuit) or not isinstance(circuits, (list, tuple)):
            circuits = [circuits]
        counts = []
        for circ in circuits:
            if not isinstance(circ, QuantumCircuit):     # a qiskit circuit (duck typed): convert
                from .adapters import circuit_from_qiskit
                circ = circuit_from_qiskit(circ)
            dist = self.exact_distribution(circ)
            counts.append({self._format_key(k, circ.cregs): v * shots for k, v in dist.items()})
        return B200Job(B200Result(counts))
