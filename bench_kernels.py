#!/usr/bin/env python
"""Per-kernel measurements for the rows of SURVEY.md section 8(d) that bench.py's headline step
does not cover: the sort path (keys -> radix sort -> reduce-by-key -> symmetric CSR), ICE on the
CSR, and the two-step allelic correction.  One JSON line per kernel with algorithmic bytes,
CUDA-event time and the fraction of the measured HBM peak.  Single GPU.

    python bench_kernels.py [--pairs 200000000] [--res 10000] [--twostep-n 6232]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=200_000_000)
    ap.add_argument("--res", type=int, default=10000)
    ap.add_argument("--twostep-n", type=int, default=6232)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()

    import torch
    from hichap_master_b200 import _abi, kernels, matrixBuilding as mb, synth
    from hichap_master_b200.device import DenseBatch, PairColumns

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0

    def ev(fn, reps=args.reps, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps, out

    def emit(kernel, ms, alg_bytes, **extra):
        gbs = alg_bytes / (ms * 1e6)
        print(json.dumps(dict(kernel=kernel, ms=ms, algorithmic_bytes=alg_bytes, GBps=gbs, frac_of_measured_peak=gbs / peak,
                              peak_GBps=peak, **extra)), flush=True)

    # ---- (a') sort path on a genome-wide matrix --------------------------------------------------
    genome = {c: l for c, l in synth.HG19.items() if c not in ("Y", "M")}
    order = [str(i) for i in range(1, 23)] + ["X"]
    c1, p1, c2, p2 = synth.genome_pairs_torch(genome, order, args.pairs, 4, dev, trans_frac=0.25)
    pairs = PairColumns(c1, p1, c2, p2, device=dev)
    P = pairs.n
    bins, total = mb._bins_from_genome(genome, args.res, [(c, c) for c in order])
    start = mb._start_table(bins, order, dev)
    chrom_bins = torch.tensor([genome[c] // args.res + 1 for c in order], dtype=torch.int32, device=dev)
    col_bits = kernels.key_col_bits(total)
    passes = (2 * col_bits + 7) // 8

    keys = torch.empty(2 * P, dtype=torch.int64, device=dev)
    n_valid = torch.zeros(1, dtype=torch.int64, device=dev)
    oob = torch.zeros(1, dtype=torch.int64, device=dev)

    def make_keys():
        _abi.check(_abi.lib().hc_pairs_to_keys(kernels.ptr(pairs.c1), kernels.ptr(pairs.p1), kernels.ptr(pairs.c2),
                                               kernels.ptr(pairs.p2), P, args.res, kernels.ptr(start),
                                               kernels.ptr(chrom_bins), len(order), 0, col_bits, kernels.ptr(keys),
                                               kernels.ptr(n_valid), kernels.ptr(oob), kernels.stream_ptr()), "keys")
    ms, _ = ev(make_keys)
    emit("pairs_to_keys_kernel", ms, 16.0 * P + 16.0 * P, pairs=P, note="16 B read + 2 keys x 8 B written per pair")

    src = keys.clone()
    tmp = torch.empty_like(keys)
    work = torch.empty(int(_abi.lib().hc_sort_work_bytes(2 * P)), dtype=torch.uint8, device=dev)
    import ctypes as C
    flag = C.c_int32(0)

    def do_sort():
        src.copy_(keys)
        _abi.check(_abi.lib().hc_sort_keys_u64(kernels.ptr(src), kernels.ptr(tmp), 2 * P, 0, 2 * col_bits, kernels.ptr(work),
                                               C.byref(flag), kernels.stream_ptr()), "sort")
    t_copy, _ = ev(lambda: src.copy_(keys))
    ms, _ = ev(do_sort)
    ms -= t_copy
    nk = 2 * P
    emit("radix sort (histogram + %d onesweep passes)" % passes, ms, (8.0 + 16.0 * passes) * nk, keys=nk, key_bits=2 * col_bits,
         Gkeys_per_s=nk / (ms * 1e6))

    ms, csr = ev(lambda: kernels.pairs_to_csr(pairs, args.res, start, chrom_bins, total, False), reps=2)
    Z_sym = csr.nnz
    Z = (Z_sym + total) // 2
    emit("sort path total (keys + sort + reduce-by-key -> CSR)", ms, (40.0 + 16.0 * passes) * P + 12.0 * Z, pairs=P, nnz_upper=Z,
         nnz_stored=Z_sym, bins=total, note="SURVEY 8(d): (40+16*passes)*P + 12*Z; this build sorts 2 keys per pair")

    # ---- (b) ICE on the CSR ------------------------------------------------------------------
    def ice():
        return mb.ice_balance_sparse(csr, bins, cis_only=False)
    ms, (w, st) = ev(ice, reps=1, warm=1)
    per_iter = st["loop_ms"] / max(st["iters"], 1)
    emit("ice_csr iteration (stream + 3 stat kernels)", per_iter, 8.0 * Z + 24.0 * total, iters=st["iters"], converged=st["converged"],
         time_to_convergence_ms=ms, loop_ms=st["loop_ms"], streamed_bytes=8.0 * Z_sym,
         GBps_streamed=8.0 * Z_sym / (per_iter * 1e6), note="algorithmic 8*Z+24*n; symmetric CSR streams 8 B x 2Z")

    del csr, keys, src, tmp, pairs
    torch.cuda.empty_cache()

    # ---- (c) two-step correction on one chromosome-sized triple ---------------------------------
    n = args.twostep_n
    L = n * 40000 - 1
    a, b = synth.genome_pairs_torch({"1": L}, ["1"], 40_000_000, 7, dev)[1::2]
    z = torch.zeros_like(a)
    g = torch.Generator(device=dev); g.manual_seed(3)
    cls = torch.rand(a.numel(), generator=g, device=dev)
    mark = (torch.rand(a.numel(), generator=g, device=dev) * 3).to(torch.uint8)
    T = DenseBatch([n], dev); H = DenseBatch([n, n], dev)
    kernels.bin_pairs_local(PairColumns(z, a, z, b, device=dev), 40000, T)
    from hichap_master_b200.construction import _sub_batch
    for hap, (lo, hi) in enumerate(((0.0, 0.07), (0.07, 0.14))):
        sel = (cls >= lo) & (cls < hi)
        pc = PairColumns(z[sel], a[sel], z[sel], b[sel], mark[sel], device=dev)
        kernels.bin_pairs_local(pc, 40000, _sub_batch(H, hap, 1), _abi.HC_BIN_SYM_BOTH)
        kernels.bin_pairs_local(pc, 40000, _sub_batch(H, hap, 1), _abi.HC_BIN_ONESIDED)
    ms, _ = ev(lambda: mb.two_step_device(T, 0, H, 0, H, 1))
    emit("two-step correction (M and P of one chromosome)", ms, 52.0 * n * n, n=n,
         note="SURVEY 8(d): 52*N^2 (TM once, MM/PM 4 reads + fp64 write each)")
    tp = H.buf.data_ptr()
    ms, _ = ev(lambda: kernels.rowstats(tp, H.lds[0], n, n, dev))
    emit("rowstats_kernel", ms, 4.0 * n * n, n=n)


if __name__ == "__main__":
    main()
