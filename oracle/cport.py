"""ctypes access to the plain-C oracle ``oracle/c/qck_oracle.c`` (TEST INFRASTRUCTURE).

Used as the timed CPU baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``:
kind "port", all host cores through OpenMP) and, in tests, as a second
independent implementation next to the numpy oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import gates

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "libqck_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_DIR, "qck_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        res = subprocess.run(["make", "-C", _DIR] + (["-B"] if force else []), capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building the C oracle failed:\n" + res.stdout + res.stderr)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_num_threads.restype = C.c_int
        L.oracle_knit_outer.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_uint64,
                                        C.c_uint64, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_knit_contract.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_int64), C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
        L.oracle_apply_1q.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_apply_cx.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_apply_cz.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_apply_2q.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oracle_probabilities.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n=None):
    """Use ``n`` OpenMP threads (default: every host core) whatever OMP_NUM_THREADS says - torchrun sets it
    to 1 for its workers, which made the multi-GPU reference arm run on one core."""
    import os
    lib().oracle_set_num_threads(int(n or os.cpu_count() or 1))
    return num_threads()


def knit_outer(tables, masks, y_begin, y_end, want_output=True, out=None):
    L = lib()
    tabs = [np.ascontiguousarray(t, dtype=np.float64) for t in tables]
    ptrs = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    cm = (C.c_uint64 * len(tabs))(*masks)
    if out is not None:
        assert out.dtype == np.float64 and out.flags.c_contiguous and len(out) == y_end - y_begin
    elif want_output:
        out = np.empty(y_end - y_begin)
    s, m = C.c_double(), C.c_double()
    L.oracle_knit_outer(len(tabs), ptrs, cm, y_begin, y_end, out.ctypes.data if out is not None else None,
                        C.byref(s), C.byref(m))
    return out, s.value, m.value


def knit_contract(tables, masks, n_out_bits, weights, rows):
    L = lib()
    tabs = [np.ascontiguousarray(t, dtype=np.float64) for t in tables]
    ptrs = (C.c_void_p * len(tabs))(*[t.ctypes.data for t in tabs])
    cm = (C.c_uint64 * len(tabs))(*masks)
    strides = (C.c_int64 * len(tabs))(*[t.shape[1] for t in tabs])
    w = np.ascontiguousarray(weights, dtype=np.float64)
    r = np.ascontiguousarray(rows, dtype=np.int32)
    out = np.empty(1 << n_out_bits)
    L.oracle_knit_contract(len(tabs), ptrs, cm, strides, n_out_bits, len(w), w.ctypes.data, r.ctypes.data,
                           out.ctypes.data)
    return out


def simulate_probabilities(circuit):
    """|amp|^2 of a measurement-free view of ``circuit`` (measurements and barriers skipped):
    index bit q = qubit q.  Only for circuits whose measurements are all terminal."""
    L = lib()
    qidx = {q: i for i, q in enumerate(circuit.qubits)}
    n = len(circuit.qubits)
    state = np.zeros(2 << n)
    state[0] = 1.0
    for ins in circuit.data:
        op = ins.operation
        name = op.name
        if name in ("barrier", "measure", "wire_cut") or type(op).__name__ == "Barrier":
            continue
        qs = [qidx[q] for q in ins.qubits]
        u = getattr(op, "_matrix", None)
        if len(qs) == 1:
            m = np.ascontiguousarray(gates.matrix(name, op.params) if u is None else u, dtype=complex)
            L.oracle_apply_1q(state.ctypes.data, n, qs[0], m.ctypes.data)
        elif u is None and name == "cx":
            L.oracle_apply_cx(state.ctypes.data, n, qs[0], qs[1])
        elif u is None and name == "cz":
            L.oracle_apply_cz(state.ctypes.data, n, qs[0], qs[1])
        else:
            m = np.ascontiguousarray(gates.matrix(name, op.params) if u is None else u, dtype=complex)
            L.oracle_apply_2q(state.ctypes.data, n, qs[0], qs[1], m.ctypes.data)
    prob = np.empty(1 << n)
    L.oracle_probabilities(state.ctypes.data, n, prob.ctypes.data)
    return prob
