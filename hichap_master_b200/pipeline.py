"""The matrix stage as one call: valid pairs (host or HBM) -> per-chromosome contact matrices
-> cis-only ICE weights (-> upper-triangular records on the host).  This is the path
``TraditionalMatrixConstruction`` drives in the reference (matrixBuilding.py:617-717) for the
local resolutions: TraditionalMatrixBuilding (:640) then `cooler balance --cis-only` (:713).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi, kernels
from ._abi import check, lib
from .device import DenseBatch, PairColumns, ptr, require_cuda, stream_ptr


class HostPairs:
    """Columnar pairs in pinned host memory (what a parser hands to the stage)."""

    def __init__(self, c1, p1, c2, p2):
        def pin(x):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32))
            t = t.to("cpu", torch.int32).contiguous()
            return t if t.is_pinned() else t.pin_memory()
        self.c1, self.p1, self.c2, self.p2 = (pin(x) for x in (c1, p1, c2, p2))
        self.n = int(self.c1.numel())
        self.nbytes = 16 * self.n


class LocalStage:
    """Reusable buffers for repeated runs of the local-resolution stage on one GPU."""

    def __init__(self, sizes, max_pairs: int, device=None):
        self.dev = require_cuda(device)
        self.batch = DenseBatch(sizes, self.dev)
        self.max_pairs = int(max_pairs)
        self.cols = [torch.empty(max(self.max_pairs, 4), dtype=torch.int32, device=self.dev) for _ in range(4)]
        self.weights_host = torch.empty(self.batch.nbins, dtype=torch.float64).pin_memory()
        self.pool = kernels.PinnedPool()
        self.banded = True      # banded binning (hc_bin_pairs_local_banded)
        self.bin_work = None

    def upload(self, hp: HostPairs) -> PairColumns:
        assert hp.n <= self.max_pairs
        for d, s in zip(self.cols, (hp.c1, hp.p1, hp.c2, hp.p2)):
            d[:hp.n].copy_(s, non_blocking=True)
        return PairColumns(*(d[:hp.n] for d in self.cols), device=self.dev)

    def run(self, pairs: PairColumns, res: int, records=False, weights_to_host=True, **ice_kw):
        """zero tiles -> bin -> [extract upper-triangular records] -> filters -> ICE."""
        b = self.batch
        b.buf.zero_()
        if self.banded:
            self.bin_work = kernels.bin_pairs_local_banded(pairs, res, b, check_bounds=False, work=self.bin_work)
        else:
            kernels.bin_pairs_local(pairs, res, b, check_bounds=False)
        recs = None
        d2h = 0
        if records:
            recs, nbytes = kernels.dense_batch_triu_records(b, self.pool)
            d2h += nbytes
        params = kernels.ice_params(**ice_kw)
        bias = kernels.ice_dense_filters(b, params)
        results, info = kernels.ice_dense_iterate(b, bias, params)
        if weights_to_host:
            self.weights_host.copy_(bias, non_blocking=False)
            d2h += 8 * b.nbins
        return dict(bias=bias, results=results, info=info, records=recs, d2h_bytes=d2h)
