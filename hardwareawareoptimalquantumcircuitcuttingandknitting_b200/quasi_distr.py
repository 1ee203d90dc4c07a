"""Device-resident quasi-probability distributions.

Mirror of ``third_party/qvm/qvm/quasi_distr.py`` (``QuasiDistr``, ``:6-86``): the
reference keeps a sparse ``dict[int, float]`` and prunes ``|v| <= ACCURACY`` at
every construction (``:3,7-10``).  Here a distribution over ``num_bits`` classical
bits is a dense float64 vector of ``2**num_bits`` entries in HBM; every operation
is one streaming kernel of ``csrc/quasi.cu`` / ``csrc/reduce.cu`` reached through
the C ABI, with the same pruning rule applied after every operation.

``ACCURACY`` is read at call time like the reference's module global;
``0.0`` (the default here) is the exact mode, ``1e-5`` reproduces the
reference's pruning.  The dict-like read API (``items``,
``keys``, ``get``, ``[]``, ``len``) copies the non-zero entries to the host on
first use, so it is only meant for small distributions.
"""
from __future__ import annotations

import ctypes as C
from typing import Union

import numpy as np

from . import _lib

ACCURACY = 0.0
REFERENCE_ACCURACY = 1e-5      # quasi_distr.py:3

__all__ = ["QuasiDistr", "ACCURACY", "REFERENCE_ACCURACY", "knit_level", "default_device"]


def default_device():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: this package has no CPU fallback for simulation or knitting")
    return torch.device("cuda", torch.cuda.current_device())


def _stream(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


class QuasiDistr:
    def __init__(self, data, num_bits: int | None = None, device=None, accuracy: float | None = None,
                 _pruned: bool = False) -> None:
        import torch
        self.accuracy = ACCURACY if accuracy is None else float(accuracy)
        self._dict = None
        if isinstance(data, torch.Tensor):
            if data.dtype != torch.float64 or data.dim() != 1 or not data.is_cuda:
                raise TypeError("QuasiDistr needs a 1-D float64 CUDA tensor")
            n = data.numel()
            if n & (n - 1) or n == 0:
                raise ValueError("length must be a power of two")
            self.num_bits = n.bit_length() - 1
            if num_bits is not None and num_bits != self.num_bits:
                raise ValueError("num_bits does not match the tensor length")
            self.values = data
            self.device = data.device
            if not _pruned:
                self._prune()
        else:
            data = dict(data)
            if num_bits is None:
                num_bits = max((int(k).bit_length() for k in data), default=0)
            self.num_bits = int(num_bits)
            self.device = torch.device(device) if device is not None else default_device()
            host = np.zeros(1 << self.num_bits, dtype=np.float64)
            for k, v in data.items():
                if abs(v) > self.accuracy:
                    host[int(k)] = v
            self.values = torch.from_numpy(host).to(self.device)

    # ---------------------------------------------------------------- helpers
    @property
    def _h(self) -> "_lib.Handle":
        return _lib.get_handle(self.device.index or 0)

    def _prune(self) -> None:
        if self.accuracy > 0.0:
            h = self._h
            h.check(h.lib.qck_qd_prune(h.ptr, self.values.data_ptr(), self.values.numel(), self.accuracy,
                                       _stream(self.device)))

    def _new(self, values) -> "QuasiDistr":
        return QuasiDistr(values, accuracy=self.accuracy, _pruned=True)

    def _same_shape(self, other: "QuasiDistr") -> None:
        if other.num_bits != self.num_bits:
            raise ValueError("QuasiDistr widths differ")

    # ---------------------------------------------------------------- construction
    @staticmethod
    def from_counts(counts: dict, num_bits: int | None = None, device=None,
                    accuracy: float | None = None) -> "QuasiDistr":
        """``quasi_distr.py:13-20``: keys are MSB-first bitstrings (spaces between registers
        ignored), values are normalised by their total; float 'counts' are accepted."""
        shots = sum(counts.values())
        data = {int("".join(key.split()), 2): value / shots for key, value in counts.items()}
        if num_bits is None and counts:
            num_bits = max(len("".join(k.split())) for k in counts)
        return QuasiDistr(data, num_bits=num_bits, device=device, accuracy=accuracy)

    # ---------------------------------------------------------------- dict-like reads (host copy)
    def to_dict(self) -> dict[int, float]:
        if self._dict is None:
            host = self.values.cpu().numpy()
            nz = np.nonzero(host)[0]
            self._dict = {int(k): float(host[k]) for k in nz}
        return self._dict

    def items(self):
        return self.to_dict().items()

    def keys(self):
        return self.to_dict().keys()

    def values_list(self):
        return list(self.to_dict().values())

    def get(self, key, default=None):
        return self.to_dict().get(key, default)

    def __getitem__(self, key):
        return self.to_dict()[key]

    def __contains__(self, key):
        return key in self.to_dict()

    def __len__(self):
        return len(self.to_dict())

    def __iter__(self):
        return iter(self.to_dict())

    def __eq__(self, other):
        if isinstance(other, QuasiDistr):
            return self.to_dict() == other.to_dict()
        if isinstance(other, dict):
            return self.to_dict() == other
        return NotImplemented

    def __repr__(self) -> str:  # pragma: no cover - cosmetic
        return f"QuasiDistr(num_bits={self.num_bits}, accuracy={self.accuracy})"

    def support_mask(self) -> int:
        """OR of the keys with a non-zero value, computed on the device (one small read-back instead of a
        full device -> host copy and a Python loop per merge operand)."""
        import torch
        if self._dict is not None:
            m = 0
            for k in self._dict:
                m |= k
            return m
        if self.num_bits == 0:
            return 0
        idx = torch.nonzero(self.values).reshape(-1, 1)
        bits = ((idx >> torch.arange(self.num_bits, device=self.device)) & 1).any(dim=0).cpu().tolist()
        return sum(1 << b for b, on in enumerate(bits) if on)

    # ---------------------------------------------------------------- algebra (quasi_distr.py:45-86)
    def nearest_probability_distribution(self) -> dict[int, float]:
        """``quasi_distr.py:28-43``; returns the plain dict the reference returns."""
        out = self.nearest_probability_distribution_dense()
        host = out.cpu().numpy()
        return {int(k): float(host[k]) for k in np.nonzero(host)[0]}

    def nearest_probability_distribution_dense(self):
        h = self._h
        out = self.values.clone()
        beta, num = C.c_double(), C.c_double()
        h.check(h.lib.qck_npd(h.ptr, out.data_ptr(), out.numel(), max(self.accuracy, 0.0), C.byref(beta),
                              C.byref(num), _stream(self.device)))
        return out

    def split(self, bit_index: int) -> tuple["QuasiDistr", "QuasiDistr"]:
        """Halves with ``bit_index`` = 0 / 1, the bit removed from the key.  The knit only ever
        splits on the current top bit (``virtual_circuit.py:60-67``), which keeps all lower
        key bits in place; other bits are not supported."""
        import torch
        if bit_index != self.num_bits - 1:
            raise NotImplementedError("split is only supported on the most-significant key bit")
        h = self._h
        half = self.values.numel() // 2
        lo = torch.empty(half, dtype=torch.float64, device=self.device)
        hi = torch.empty(half, dtype=torch.float64, device=self.device)
        h.check(h.lib.qck_qd_split(h.ptr, self.values.data_ptr(), self.values.numel(), bit_index,
                                   lo.data_ptr(), hi.data_ptr(), self.accuracy, _stream(self.device)))
        return self._new(lo), self._new(hi)

    def merge(self, other: "QuasiDistr") -> "QuasiDistr":
        """XOR-key outer product (``quasi_distr.py:55-60``); the supports must be disjoint,
        which is what the reference silently relies on."""
        import torch
        self._same_shape(other)
        ma, mb = self.support_mask(), other.support_mask()
        if ma & mb:
            raise ValueError("merge: key supports overlap")
        h = self._h
        out = torch.empty_like(self.values)
        h.check(h.lib.qck_qd_merge(h.ptr, self.values.data_ptr(), ma, other.values.data_ptr(), mb,
                                   out.data_ptr(), out.numel(), self.accuracy, _stream(self.device)))
        return self._new(out)

    def _axpby(self, a: float, b: float, other: "QuasiDistr | None") -> "QuasiDistr":
        import torch
        h = self._h
        out = torch.empty_like(self.values)
        h.check(h.lib.qck_qd_axpby(h.ptr, a, self.values.data_ptr(), b,
                                   other.values.data_ptr() if other is not None else None,
                                   out.data_ptr(), out.numel(), self.accuracy, _stream(self.device)))
        return self._new(out)

    def __add__(self, other: "QuasiDistr") -> "QuasiDistr":
        self._same_shape(other)
        return self._axpby(1.0, 1.0, other)

    def __sub__(self, other: "QuasiDistr") -> "QuasiDistr":
        self._same_shape(other)
        return self._axpby(1.0, -1.0, other)

    def __mul__(self, other: Union[int, float, "QuasiDistr"]) -> "QuasiDistr":
        if isinstance(other, QuasiDistr):
            return self.merge(other)
        elif isinstance(other, float) or isinstance(other, int):
            return self._axpby(float(other), 0.0, None)
        raise TypeError(f"Cannot multiply QuasiDistr by {type(other)}")

    def __rmul__(self, other):
        return self.__mul__(other)


def knit_level(vgate, results: list[QuasiDistr], clbit_idx: int) -> QuasiDistr:
    """``Virtual*.knit`` on the device (``virtual_gates.py:105-124,179-194,262-286``).

    Exact mode: one fused kernel over the coefficient pairs.  With pruning: the reference's
    sequence of split / + / - / * (left to right, pruning after each) replayed with the
    elementwise kernels, because there the order is part of the result."""
    import torch
    from math import cos, sin
    from .virtual_gates import RZZ_ACCURACY
    first = results[0]
    acc = first.accuracy
    if acc <= 0.0:
        h = first._h
        n = first.values.numel()
        for r in results:
            first._same_shape(r)
        if clbit_idx != first.num_bits - 1:
            raise NotImplementedError("knit is only supported on the most-significant key bit")
        coefs = vgate.knit_coefficients()
        ptrs = (C.c_void_p * len(results))(*[r.values.data_ptr() for r in results])
        c0 = (C.c_double * len(results))(*[c[0] for c in coefs])
        c1 = (C.c_double * len(results))(*[c[1] for c in coefs])
        out = torch.empty(n // 2, dtype=torch.float64, device=first.device)
        h.check(h.lib.qck_qd_knit_level(h.ptr, len(results), ptrs, n, clbit_idx, c0, c1, out.data_ptr(),
                                        _stream(first.device)))
        return QuasiDistr(out, accuracy=acc, _pruned=True)
    if vgate.knit_form == "chain":
        signs = [1 if c[0] > 0 else -1 for c in vgate.knit_coefficients()]
        total = None
        for r, sg in zip(results, signs):
            r0, r1 = r.split(clbit_idx)
            diff = r0 - r1
            total = diff if total is None else (total + diff if sg > 0 else total - diff)
        return 0.5 * total
    m_theta = -vgate.params[0]
    c, s = cos(m_theta / 2), sin(m_theta / 2)
    if abs(c) < RZZ_ACCURACY:
        r, _ = results[0].split(clbit_idx)
        return r * s ** 2
    if abs(s) < RZZ_ACCURACY:
        r, _ = results[0].split(clbit_idx)
        return r * c ** 2
    r0, _ = results[0].split(clbit_idx)
    r1, _ = results[1].split(clbit_idx)
    r23 = results[2] + results[3]
    r45 = results[4] + results[5]
    r230, r231 = r23.split(clbit_idx)
    r450, r451 = r45.split(clbit_idx)
    return (r0 * c ** 2) + (r1 * s ** 2) + (r230 - r231 - r450 + r451) * c * s
