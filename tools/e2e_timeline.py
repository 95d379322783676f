#!/usr/bin/env python
"""CUDA-event timeline of the end-to-end C2 step (`LocalStage.run_from_host`): when the H2D copies,
the binning, the record extraction + D2H and the ICE loop start and end relative to the step start.
Single GPU; prints one JSON line.  python tools/e2e_timeline.py [--pairs N]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=400_000_000)
    ap.add_argument("--res", type=int, default=40000)
    args = ap.parse_args()
    import torch
    from hichap_master_b200 import kernels, synth
    from hichap_master_b200.pipeline import HostPairs, LocalStage
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    genome = {c: l for c, l in synth.HG19.items() if c not in ("Y", "M")}
    order = [str(i) for i in range(1, 23)] + ["X"]
    c1, p1, c2, p2 = synth.genome_pairs_torch(genome, order, args.pairs, 4, dev, trans_frac=0.0)
    host = HostPairs(c1.cpu(), p1.cpu(), c2.cpu(), p2.cpu())
    del c1, p1, c2, p2
    sizes = [genome[c] // args.res + 1 for c in order]
    st = LocalStage(sizes, args.pairs, dev)
    for _ in range(2):
        st.run_from_host(host, args.res, records=True)
    torch.cuda.synchronize()

    marks = {}
    def mark(name, stream):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        marks[name] = e

    # the same sequence as run_from_host / _after_binning, with event marks
    main_s = torch.cuda.current_stream()
    b, n, res = st.batch, host.n, args.res
    mark("t0", main_s)
    b.buf.zero_()
    bb = kernels.BandedBinning(b, res, work=st.bin_work)
    ready = torch.cuda.Event(); ready.record(main_s); st.copy.wait_event(ready)
    chunk = 1 << 24
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        with torch.cuda.stream(st.copy):
            st.chrom8[0][lo:hi].copy_(host.c1[lo:hi], non_blocking=True)
            st.cols[1][lo:hi].copy_(host.p1[lo:hi], non_blocking=True)
            st.chrom8[1][lo:hi].copy_(host.c2[lo:hi], non_blocking=True)
            st.cols[3][lo:hi].copy_(host.p2[lo:hi], non_blocking=True)
            landed = torch.cuda.Event(); landed.record(st.copy)
        main_s.wait_event(landed)
        bb.accumulate(st.chrom8[0][lo:hi], st.cols[1][lo:hi], st.chrom8[1][lo:hi], st.cols[3][lo:hi])
    mark("h2d_done", st.copy)
    mark("last_chunk_binned", main_s)
    bb.finish(check_bounds=False)
    mark("binned", main_s)
    binned = torch.cuda.Event(); binned.record(main_s); st.side.wait_event(binned)
    with torch.cuda.stream(st.side):
        recs, nbytes = kernels.dense_batch_triu_records(b, st.pool, sync=False)
        mark("records_d2h_done", st.side)
    params = kernels.ice_params()
    bias = kernels.ice_dense_filters(b, params)
    mark("filters_done", main_s)
    results, info = kernels.ice_dense_iterate(b, bias, params)
    mark("ice_done", main_s)
    st.weights_host.copy_(bias, non_blocking=False)
    st.side.synchronize()
    torch.cuda.synchronize()
    t0 = marks.pop("t0")
    out = {k: round(t0.elapsed_time(e), 3) for k, e in marks.items()}
    out.update(pairs=n, h2d_bytes=host.nbytes, d2h_record_bytes=nbytes, ice_loop_ms=info.loop_ms,
               h2d_GBps=round(host.nbytes / out["h2d_done"] / 1e6, 1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
