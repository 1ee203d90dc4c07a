"""Hellinger fidelity on the device.

``src/HwAwareCutter/Utilities.py:224`` scores a cut run with qiskit's
``hellinger_fidelity(uncut, cut)`` (qiskit-terra 0.25.2.1
``quantum_info/analysis/distance.py``, not vendored): both inputs are
normalised by their own totals, ``H^2 = 1/2 sum_x (sqrt p_x - sqrt q_x)^2`` over
the union of keys and ``F = (1 - H^2)^2``.  With ``S_p = sum p``, ``S_q = sum q``
and ``BC = sum sqrt(p q)`` this is ``F = (BC / sqrt(S_p S_q))^2`` - three sums one
streaming kernel produces (``qck_hellinger``).  For factorised distributions
(no virtual gates) ``hellinger_fidelity_factored`` evaluates the same three sums
with ``qck_knit_outer`` over square-rooted tables without materialising either
side.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from .quasi_distr import QuasiDistr, default_device

__all__ = ["hellinger_fidelity", "hellinger_fidelity_factored"]


def _dense(x, num_bits, device):
    import torch
    if isinstance(x, QuasiDistr):
        return x.values
    if isinstance(x, torch.Tensor):
        return x
    return QuasiDistr(dict(x), num_bits=num_bits, device=device, accuracy=0.0).values


def hellinger_fidelity(p, q, num_bits: int | None = None, device=None) -> float:
    """``p``, ``q``: dicts, ``QuasiDistr`` or dense float64 CUDA tensors of equal length.
    ``F({}, {}) = 1`` as in qiskit (both normalisations are skipped for empty inputs)."""
    import torch
    if isinstance(p, dict) and isinstance(q, dict):
        if not p and not q:
            return 1.0
        if num_bits is None:
            num_bits = max(max((int(k).bit_length() for k in p), default=0),
                           max((int(k).bit_length() for k in q), default=0))
    device = default_device() if device is None else device
    dp, dq = _dense(p, num_bits, device), _dense(q, num_bits, device)
    if dp.numel() != dq.numel():
        raise ValueError("distributions have different widths")
    handle = _lib.get_handle(dp.device.index or 0)
    res = torch.empty(3, dtype=torch.float64, device=dp.device)
    handle.check(handle.lib.qck_hellinger(handle.ptr, dp.data_ptr(), dq.data_ptr(), dp.numel(), res.data_ptr(),
                                          torch.cuda.current_stream(dp.device).cuda_stream))
    sp, sq, bc = res.cpu().tolist()
    if sp == 0.0 or sq == 0.0:
        return 1.0 if sp == sq else 0.0
    return (bc / (sp * sq) ** 0.5) ** 2


def hellinger_fidelity_factored(tables_p, masks_p, tables_q, masks_q, n_bits: int, device=None,
                                y_range: tuple[int, int] | None = None) -> tuple[float, float, float]:
    """-> (S_p, S_q, BC) over ``y_range`` for p = prod_f tables_p[f][pext(y, masks_p[f])] and
    likewise q.  All tables must be non-negative (true probabilities)."""
    import torch
    device = default_device() if device is None else device
    handle = _lib.get_handle(getattr(device, "index", None) or 0)
    stream = torch.cuda.current_stream(device).cuda_stream
    y0, y1 = y_range if y_range is not None else (0, 1 << n_bits)

    def outer_sum(tables, masks):
        ptrs = (C.c_void_p * len(tables))(*[t.data_ptr() for t in tables])
        cm = (C.c_uint64 * len(tables))(*masks)
        st = torch.zeros(4, dtype=torch.float64, device=device)
        handle.check(handle.lib.qck_knit_outer(handle.ptr, len(tables), ptrs, cm, n_bits, y0, y1, None,
                                               st.data_ptr(), stream))
        return st

    def sqrt_of(t):
        out = torch.empty_like(t)
        handle.check(handle.lib.qck_qd_sqrt(handle.ptr, t.data_ptr(), out.data_ptr(), t.numel(), stream))
        return out

    sp = outer_sum(tables_p, masks_p)
    sq = outer_sum(tables_q, masks_q)
    roots = [sqrt_of(t) for t in tables_p] + [sqrt_of(t) for t in tables_q]
    bc = outer_sum(roots, list(masks_p) + list(masks_q))
    return float(sp[0].item()), float(sq[0].item()), float(bc[0].item())
