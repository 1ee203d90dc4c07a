"""Exact branching statevector semantics (TEST INFRASTRUCTURE - see oracle/__init__.py).

Replaces ``backend.run(instantiations, shots=...)`` + ``get_counts()``
(``third_party/qvm/qvm/run.py:42,48,54-55``; the arithmetic lives in qiskit-aer
0.13.0, third-party, not vendored) by the *exact* outcome distribution of each
circuit, per SURVEY.md A.2:

* start in |0...0>, apply the ops in order;
* a measurement whose qubit is used again later splits every live branch into
  its two projected, un-normalised states and records the outcome in the clbit;
* terminal measurements are read off |amp|^2 at the end;
* ``p(key)`` = squared norm of everything that ends with classical value
  ``key`` (bit i of key = clbit i); unwritten clbits are 0.

Plain numpy, one op at a time, no fusion - slow on purpose, easy to audit.
Accepts any circuit object exposing ``qubits``, ``clbits`` and ``data`` whose
entries have ``operation`` (``name``, ``params``, optional ``_matrix``),
``qubits`` and ``clbits``.
"""
import numpy as np

from . import gates


def _apply_1q(psi, n, q, u):
    v = psi.reshape(1 << (n - 1 - q), 2, 1 << q)
    return np.einsum("ab,xbl->xal", u, v).reshape(-1)


def _apply_2q(psi, n, q0, q1, u):
    """``u`` is little-endian in (q0, q1): row/col index = b(q0) + 2 b(q1)."""
    t = psi.reshape((2,) * n)                  # axis k <-> qubit n-1-k
    a0, a1 = n - 1 - q0, n - 1 - q1
    u4 = u.reshape(2, 2, 2, 2)                 # [o1, o0, i1, i0]
    t = np.tensordot(u4, t, axes=([2, 3], [a1, a0]))   # -> [o1, o0, rest...]
    t = np.moveaxis(t, [0, 1], [a1, a0])
    return np.ascontiguousarray(t).reshape(-1)


def lower(circuit):
    """Flatten a circuit into ('u1', q, U) / ('u2', q0, q1, U) / ('m', q, c) tuples."""
    qidx = {q: i for i, q in enumerate(circuit.qubits)}
    cidx = {c: i for i, c in enumerate(circuit.clbits)}
    ops = []
    for ins in circuit.data:
        op = ins.operation
        name = op.name
        if name in ("barrier", "wire_cut") or type(op).__name__ in ("Barrier", "WireCut"):
            continue
        qs = [qidx[q] for q in ins.qubits]
        if name == "measure":
            ops.append(("m", qs[0], cidx[ins.clbits[0]]))
            continue
        u = getattr(op, "_matrix", None)
        if u is None:
            u = gates.matrix(name, op.params)
        if len(qs) == 1:
            ops.append(("u1", qs[0], np.asarray(u, dtype=complex)))
        elif len(qs) == 2:
            ops.append(("u2", qs[0], qs[1], np.asarray(u, dtype=complex)))
        else:
            raise ValueError(f"oracle: {name} on {len(qs)} qubits")
    return ops


def exact_distribution(circuit):
    """dict[int, float]: exact probability of every classical outcome (zeros omitted)."""
    n = len(circuit.qubits)
    ops = lower(circuit)
    last_use = {}
    for i, op in enumerate(ops):
        for q in ((op[1],) if op[0] in ("u1", "m") else (op[1], op[2])):
            last_use[q] = i
    psi0 = np.zeros(1 << n, dtype=complex)
    psi0[0] = 1.0
    branches = [(psi0, 0)]
    deferred = []                                   # (qubit, clbit) terminal measurements
    for i, op in enumerate(ops):
        if op[0] == "u1":
            branches = [(_apply_1q(p, n, op[1], op[2]), c) for p, c in branches]
        elif op[0] == "u2":
            branches = [(_apply_2q(p, n, op[1], op[2], op[3]), c) for p, c in branches]
        else:
            q, cb = op[1], op[2]
            if last_use[q] == i:
                deferred.append((q, cb))
                continue
            new = []
            sel = ((np.arange(1 << n) >> q) & 1).astype(bool)
            for p, c in branches:
                p0 = np.where(sel, 0, p)
                p1 = np.where(sel, p, 0)
                if np.any(p0):
                    new.append((p0, c & ~(1 << cb)))
                if np.any(p1):
                    new.append((p1, (c & ~(1 << cb)) | (1 << cb)))
            branches = new
    out = {}
    idx = np.arange(1 << n, dtype=np.int64)
    dkey = np.zeros(1 << n, dtype=np.int64)
    dmask = 0
    for q, cb in deferred:
        dkey |= ((idx >> q) & 1) << cb
        dmask |= 1 << cb
    for p, c in branches:
        prob = (p.real * p.real + p.imag * p.imag)
        keys = dkey | (c & ~dmask)
        uniq, inv = np.unique(keys, return_inverse=True)
        sums = np.bincount(inv, weights=prob)
        for k, v in zip(uniq.tolist(), sums.tolist()):
            if v != 0.0:
                out[k] = out.get(k, 0.0) + v
    return out


def dense(dist, nbits):
    v = np.zeros(1 << nbits)
    for k, p in dist.items():
        v[k] = p
    return v
