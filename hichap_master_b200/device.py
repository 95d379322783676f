"""Device-memory plumbing for the C ABI: torch is used only to own HBM buffers, pick the
stream and move bytes between host and device.  No torch op computes anything on the path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi

ROW_ALIGN = 128  # int32 elements -> 512-byte rows: every 128-column chunk a warp streams is full


_inited = set()


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _abi.HcError("hichap_master_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _inited:
        with torch.cuda.device(idx):
            _abi.check(_abi.lib().hc_init(), "hc_init")
        _inited.add(idx)
    return dev


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def as_int32_counts(arr) -> np.ndarray:
    """Contact counts as a contiguous int32 array.  The tiles are int32 (cooler stores counts as int32
    too, matrixBuilding.py:196); the reference's own matrices are int64 / float arrays holding integers.
    Anything that is not an integer in int32 range is rejected instead of being truncated or wrapped."""
    a = np.asarray(arr)
    if a.dtype == np.int32:
        return np.ascontiguousarray(a)
    if a.dtype == np.bool_ or np.issubdtype(a.dtype, np.integer):
        if a.size and (a.min() < -2**31 or a.max() > 2**31 - 1):
            raise OverflowError("contact counts exceed the int32 range of the dense tiles")
        return np.ascontiguousarray(a, dtype=np.int32)
    if np.issubdtype(a.dtype, np.floating):
        if a.size and not np.all(np.isfinite(a)):
            raise TypeError("contact matrix holds NaN/inf; integer counts expected")
        if a.size and (a.min() < -2**31 or a.max() > 2**31 - 1):
            raise OverflowError("contact counts exceed the int32 range of the dense tiles")
        out = a.astype(np.int32)
        if a.size and not np.array_equal(out, a):
            raise TypeError("contact matrix holds non-integral values; integer counts expected "
                            "(corrected float matrices do not go through the int32 tiles)")
        return np.ascontiguousarray(out)
    raise TypeError("unsupported matrix dtype %s" % a.dtype)


class DenseBatch:
    """Several zero-initialised int32 row-major matrices in ONE device buffer (the "dense
    tiles" of the binning kernel), plus the small device tables the kernels index them by."""

    def __init__(self, sizes, device=None):
        dev = require_cuda(device)
        self.sizes = [int(n) for n in sizes]
        self.lds = [max(ROW_ALIGN, round_up(n, ROW_ALIGN)) for n in self.sizes]
        offs, cur = [], 0
        for n, ld in zip(self.sizes, self.lds):
            offs.append(cur)
            cur += n * ld
        self.offsets = offs
        self.numel = max(cur, ROW_ALIGN)
        self.buf = torch.zeros(self.numel, dtype=torch.int32, device=dev)
        self.mat_off = torch.tensor(offs, dtype=torch.int64, device=dev)
        self.mat_n = torch.tensor(self.sizes, dtype=torch.int32, device=dev)
        self.mat_ld = torch.tensor(self.lds, dtype=torch.int32, device=dev)
        bin_off = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.h_bin_off = bin_off
        self.bin_off = torch.from_numpy(bin_off).to(dev)
        self.h_mat_n = (C.c_int32 * len(self.sizes))(*self.sizes)
        self.nbins = int(bin_off[-1])
        self.device = dev

    def __len__(self):
        return len(self.sizes)

    def view(self, i) -> torch.Tensor:
        """(n, ld) strided view of matrix i (columns >= n are zero padding)."""
        n, ld, o = self.sizes[i], self.lds[i], self.offsets[i]
        return self.buf[o:o + n * ld].view(n, ld)

    def mat_ptr(self, i) -> C.c_void_p:
        return C.c_void_p(self.buf.data_ptr() + 4 * self.offsets[i])

    def to_numpy(self, i, dtype=np.int64) -> np.ndarray:
        n = self.sizes[i]
        return self.view(i)[:, :n].cpu().numpy().astype(dtype, copy=False)

    def load(self, i, arr: np.ndarray):
        """Upload a host matrix (any integer dtype) into slot i."""
        n = self.sizes[i]
        assert arr.shape == (n, n), (arr.shape, n)
        self.view(i)[:, :n].copy_(torch.from_numpy(as_int32_counts(arr)))

    @classmethod
    def from_numpy(cls, mats, device=None):
        b = cls([m.shape[0] for m in mats], device)
        for i, m in enumerate(mats):
            b.load(i, m)
        return b


class PairColumns:
    """Columnar valid pairs resident in HBM: chromosome index / fragment mid-point per mate
    and the optional allelic mark (0 Both, 1 R1, 2 R2, 3 other)."""

    def __init__(self, c1, p1, c2, p2, mark=None, device=None):
        dev = require_cuda(device)

        def up(x, dt):
            if isinstance(x, torch.Tensor):
                return x.to(device=dev, dtype=dt).contiguous()
            return torch.from_numpy(np.ascontiguousarray(x, dtype={torch.int32: np.int32, torch.uint8: np.uint8}[dt])).to(dev)

        self.c1, self.p1, self.c2, self.p2 = (up(x, torch.int32) for x in (c1, p1, c2, p2))
        self.mark = None if mark is None else up(mark, torch.uint8)
        self.n = int(self.c1.numel())
        assert self.p1.numel() == self.n and self.c2.numel() == self.n and self.p2.numel() == self.n
        self.device = dev
