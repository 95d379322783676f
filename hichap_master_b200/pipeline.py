"""The matrix stage as one call: valid pairs (host or HBM) -> per-chromosome contact matrices
-> cis-only ICE weights (-> upper-triangular records on the host).  This is the path
``TraditionalMatrixConstruction`` drives in the reference (matrixBuilding.py:617-717) for the
local resolutions: TraditionalMatrixBuilding (:640) then `cooler balance --cis-only` (:713).
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels
from ._abi import check, lib
from .device import DenseBatch, PairColumns, ptr, require_cuda, stream_ptr


class HostPairs:
    """Columnar pairs in pinned host memory (what a parser hands to the stage): chromosome index
    as uint8 (255 = filtered; fewer bytes over PCIe) and fragment mid-point as int32."""

    def __init__(self, c1, p1, c2, p2):
        def pin(x, dt):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            t = t.to("cpu")
            if dt == torch.uint8:
                if int(t.max().item() if t.numel() else 0) >= 255:
                    raise ValueError("more than 254 chromosomes: use int32 PairColumns")
                t = torch.where(t < 0, torch.full_like(t, 255), t).to(torch.uint8)
            else:
                t = t.to(torch.int32)
            t = t.contiguous()
            return t if t.is_pinned() else t.pin_memory()
        self.c1, self.c2 = pin(c1, torch.uint8), pin(c2, torch.uint8)
        self.p1, self.p2 = pin(p1, torch.int32), pin(p2, torch.int32)
        self.n = int(self.p1.numel())
        self.nbytes = 10 * self.n


class LocalStage:
    """Reusable buffers for repeated runs of the local-resolution stage on one GPU."""

    def __init__(self, sizes, max_pairs: int, device=None):
        self.dev = require_cuda(device)
        self.batch = DenseBatch(sizes, self.dev)
        self.max_pairs = int(max_pairs)
        n = max(self.max_pairs, 4)
        self.cols = [torch.empty(n, dtype=torch.int32, device=self.dev) for _ in range(4)]
        self.chrom8 = [torch.empty(n, dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self.weights_host = torch.empty(self.batch.nbins, dtype=torch.float64).pin_memory()
        self.pool = kernels.PinnedPool()
        self.banded = True      # banded binning (hc_bin_pairs_local_banded)
        self.bin_work = None
        self.side = torch.cuda.Stream(device=self.dev)
        self.copy = torch.cuda.Stream(device=self.dev)
        # out-of-range pairs (position beyond the chromosome end): counted on the device during binning,
        # read once per run where the results are synchronised anyway (the reference raises IndexError)
        self.oob = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.oob_host = torch.zeros(1, dtype=torch.int64).pin_memory()

    def upload(self, hp: HostPairs) -> PairColumns:
        assert hp.n <= self.max_pairs
        n = hp.n
        self.cols[1][:n].copy_(hp.p1, non_blocking=True)
        self.cols[3][:n].copy_(hp.p2, non_blocking=True)
        for k, (src, dst) in enumerate(((hp.c1, self.cols[0]), (hp.c2, self.cols[2]))):
            self.chrom8[k][:n].copy_(src, non_blocking=True)
            check(lib().hc_widen_u8_i32(ptr(self.chrom8[k]), ptr(dst), n, stream_ptr()), "hc_widen_u8_i32")
        return PairColumns(*(d[:n] for d in self.cols), device=self.dev)

    def run(self, pairs: PairColumns, res: int, records=False, weights_to_host=True, **ice_kw):
        """zero tiles -> bin -> [upper-triangular records, copied out on a side stream while ICE
        runs] -> filters -> ICE."""
        b = self.batch
        b.buf.zero_()
        self.oob.zero_()
        if self.banded:
            self.bin_work = kernels.bin_pairs_local_banded(pairs, res, b, check_bounds=False, work=self.bin_work,
                                                           oob=self.oob)
        else:
            kernels.bin_pairs_local(pairs, res, b, check_bounds=False, oob=self.oob)
        return self._after_binning(records, weights_to_host, **ice_kw)

    def run_from_host(self, hp: HostPairs, res: int, records=False, weights_to_host=True, chunk_pairs=1 << 24,
                      **ice_kw):
        """``run`` fed straight from pinned host columns: the pairs cross PCIe in chunks on a copy
        stream and every chunk is binned (uint8 chromosome columns read as they are) while the next
        ones are in flight, so that only the last chunk's binning is exposed after the transfer."""
        assert hp.n <= self.max_pairs
        b, n = self.batch, hp.n
        main = torch.cuda.current_stream()
        b.buf.zero_()
        self.oob.zero_()
        bb = kernels.BandedBinning(b, res, work=self.bin_work, oob=self.oob)
        ready = torch.cuda.Event()
        ready.record(main)                 # the device columns may still be read by work queued earlier on `main`
        self.copy.wait_event(ready)
        chunk = max(16, int(chunk_pairs) // 16 * 16)
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            with torch.cuda.stream(self.copy):
                self.chrom8[0][lo:hi].copy_(hp.c1[lo:hi], non_blocking=True)
                self.cols[1][lo:hi].copy_(hp.p1[lo:hi], non_blocking=True)
                self.chrom8[1][lo:hi].copy_(hp.c2[lo:hi], non_blocking=True)
                self.cols[3][lo:hi].copy_(hp.p2[lo:hi], non_blocking=True)
                landed = torch.cuda.Event()
                landed.record(self.copy)
            main.wait_event(landed)
            bb.accumulate(self.chrom8[0][lo:hi], self.cols[1][lo:hi], self.chrom8[1][lo:hi], self.cols[3][lo:hi])
        self.bin_work = bb.finish(check_bounds=False)
        return self._after_binning(records, weights_to_host, **ice_kw)

    def _after_binning(self, records, weights_to_host, **ice_kw):
        b = self.batch
        recs, d2h = None, 0
        main = torch.cuda.current_stream()
        if records:
            binned = torch.cuda.Event()
            binned.record(main)
            self.side.wait_event(binned)
            with torch.cuda.stream(self.side):
                recs, nbytes = kernels.dense_batch_triu_records(b, self.pool, sync=False)
            d2h += nbytes
        self.oob_host.copy_(self.oob, non_blocking=True)       # 8 bytes; complete before ice_dense_iterate returns
        params = kernels.ice_params(**ice_kw)
        bias = kernels.ice_dense_filters(b, params)
        results, info = kernels.ice_dense_iterate(b, bias, params)
        n_oob = int(self.oob_host.item())
        if n_oob:
            raise IndexError("%d pair(s) fall outside the intra-chromosomal matrices (position beyond the "
                             "chromosome length in the genomeSize file)" % n_oob)
        if weights_to_host:
            self.weights_host.copy_(bias, non_blocking=False)
            d2h += 8 * b.nbins
        if records:
            self.side.synchronize()
        return dict(bias=bias, results=results, info=info, records=recs, d2h_bytes=d2h, oob_pairs=n_oob)
