#!/usr/bin/env python
"""bench.py -- the matrix-stage hot path on BASELINE.json's config C2:
hg19 chr1-22,X intra-chromosomal matrices at 40 kb from synthetic cis valid pairs, binned on
the GPU and ICE-balanced (`cooler balance --ignore-diags 1 --cis-only` semantics) to convergence.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl ours|reference]

One "step" = zero the dense tiles, bin all pairs, run the bin filters and iterate every
chromosome to convergence.  `value` = device-resident step time (ms, lower is better);
`e2e` = the same through the host-facing call: pinned host columns -> HBM, the step, the
upper-triangular records and the weight vector back to the host.  N > 1: chromosomes are
LPT-sharded over ranks (no collective on the data path); time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES = 40000
METRIC = "matrix_stage_time_to_ice_convergence"
UNIT = "ms"


def c2_genome():
    from hichap_master_b200 import synth
    g = {c: l for c, l in synth.HG19.items() if c not in ("Y", "M")}
    order = [str(i) for i in range(1, 23)] + ["X"]
    return g, order


def pair_shares(genome, order, total):
    lens = np.array([genome[c] for c in order], dtype=np.float64)
    share = np.floor(total * lens / lens.sum()).astype(np.int64)
    share[0] += total - share.sum()
    return share


def workload_name(pairs):
    return ("C2: hg19 chr1-22,X intra-chromosomal 40 kb matrices, %d synthetic cis valid pairs, "
            "binning + cis-only ICE (ignore_diags=1, cooler defaults) to convergence" % pairs)


# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hichap_master_b200 import _abi, kernels, shard, synth
    from hichap_master_b200.device import PairColumns
    from hichap_master_b200.pipeline import HostPairs, LocalStage

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    genome, order = c2_genome()
    sizes_all = [genome[c] // RES + 1 for c in order]
    shares = pair_shares(genome, order, args.pairs)
    mine = shard.chromosome_shards(sizes_all, world)[rank]
    sizes = [sizes_all[i] for i in mine]

    # ---- synthetic inputs, generated on the device, one seeded stream per chromosome ---------
    cs, p1s, p2s = [], [], []
    for li, gi in enumerate(mine):
        c = order[gi]
        _, a, _, b = synth.genome_pairs_torch({c: genome[c]}, [c], int(shares[gi]), 2000 + gi, dev)
        cs.append(torch.full((a.numel(),), li, dtype=torch.int32, device=dev)); p1s.append(a); p2s.append(b)
    c1 = torch.cat(cs); p1 = torch.cat(p1s); p2 = torch.cat(p2s)
    del cs, p1s, p2s
    g = torch.Generator(device=dev); g.manual_seed(99 + rank)
    perm = torch.randperm(c1.numel(), generator=g, device=dev)
    c1, p1, p2 = c1[perm].contiguous(), p1[perm].contiguous(), p2[perm].contiguous()
    del perm
    n_local = int(c1.numel())
    pairs = PairColumns(c1, p1, c1, p2, device=dev)
    stage = LocalStage(sizes, n_local, dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """device time of `steps` calls, max over ranks (ms per step) + last result"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item()), out

    os.environ.setdefault("HC_ICE_TIME_KERNEL", "1")    # per-launch events around the ICE stream kernel (roofline)
    step = lambda: stage.run(pairs, RES, records=False, weights_to_host=False)
    for _ in range(args.warmup):
        step()
    launches0 = _abi.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, out = timed(step, args.steps)
    launches = (_abi.launch_count() - launches0) // args.steps
    # roofline of the dominant kernel (the fused ICE iteration), from the same timed steps
    iters = out["results"]["iters"].astype(np.int64)
    info = out["info"]
    packed = bool(info.packed)
    # algorithmic bytes of the stream kernel: the matrix once per iteration -- 4 B per cell as int32 tiles, or
    # (packed encoding, default) 1 B per cell + 8 B per overflow cell (col, extra count)
    cell_iters = float(sum(int(it) * n * n for it, n in zip(iters, sizes)))
    mean_iters = cell_iters / max(float(sum(n * n for n in sizes)), 1.0)
    ice_bytes = (1.0 * cell_iters + 8.0 * float(info.overflow_cells) * mean_iters) if packed else 4.0 * cell_iters
    ice_kernel = "ice_q8_mma_kernel" if packed else "ice_dense_stream_kernel"
    loop_ms = float(info.loop_ms)
    n_iter_launches = int(max(iters))   # stream-kernel launches that had work (graph replays run in chunks of 8)

    # ---- end to end through the host-facing call ---------------------------------------------
    if args.skip_e2e:
        if rank == 0:
            sampler.stop()
            print(json.dumps({"tuning_only": True, "ms_per_step": ms_step, "ice_loop_ms": loop_ms,
                              "ice_GBps": ice_bytes / (loop_ms * 1e6), "ice_launches": n_iter_launches, "ice_kernel": ice_kernel,
                              "stream_full_ms": float(info.stream_full_ms), "stream_full_launches": int(info.stream_full_launches),
                              "pack_ms": float(info.pack_ms), "overflow_cells": int(info.overflow_cells), "iters_max": int(max(iters)),
                              "variant": os.environ.get("HC_ICE_VARIANT"), "item_kb": os.environ.get("HC_ICE_ITEM_KB")}))
        if world > 1:
            dist.destroy_process_group()
        return
    host = HostPairs(c1.cpu(), p1.cpu(), c1.cpu(), p2.cpu())
    if os.environ.get("HC_E2E_CHUNKED", "1") == "1":     # chunked H2D overlapped with binning (default)
        e2e_step = lambda: stage.run_from_host(host, RES, records=True, weights_to_host=True)
    else:                                                   # A/B: whole-column upload, then the step
        e2e_step = lambda: stage.run(stage.upload(host), RES, records=True, weights_to_host=True)
    for _ in range(max(1, min(args.warmup, 2))):
        o2 = e2e_step()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e, o2 = timed(e2e_step, e2e_steps)
    clocks = sampler.stop() if rank == 0 else None
    h2d = torch.tensor([float(host.nbytes)], dtype=torch.float64, device=dev)
    d2h = torch.tensor([float(o2["d2h_bytes"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(h2d); dist.all_reduce(d2h)

    # ---- per-kernel breakdown (rank 0, outside the timed regions) ----------------------------
    breakdown = None
    if rank == 0:
        def ev_time(fn, reps=3):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        t_zero = ev_time(lambda: stage.batch.buf.zero_())
        t_bin_direct = ev_time(lambda: (stage.batch.buf.zero_(), kernels.bin_pairs_local(pairs, RES, stage.batch, check_bounds=False))) - t_zero
        t_bin = ev_time(lambda: (stage.batch.buf.zero_(), kernels.bin_pairs_local_banded(
            pairs, RES, stage.batch, check_bounds=False, work=stage.bin_work))) - t_zero
        params = kernels.ice_params()
        t_filt = ev_time(lambda: kernels.ice_dense_filters(stage.batch, params))
        sq = float(sum(n * n for n in sizes))
        breakdown = {
            "zero_tiles_ms": t_zero, "binning_ms": t_bin, "binning_direct_atomics_ms": t_bin_direct, "ice_filters_ms": t_filt, "ice_loop_ms": loop_ms,
            "binning_GBps": 16.0 * n_local / (t_bin * 1e6), "binning_Gpairs_per_s": n_local / (t_bin * 1e6),
            "ice_filters_GBps": 4.0 * sq / (t_filt * 1e6),
            "ice_iters": [int(i) for i in iters], "ice_iter_launches": n_iter_launches,
        }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None          # dram__bytes_read.sum + dram__bytes_write.sum of one full launch, from the committed ncu capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = float(tj["traffic_bytes_per_launch"]) if world == 1 and str(tj.get("kernel", "ice_dense_stream_kernel")).startswith(ice_kernel) else None
        traffic_note = tj.get("note", "")
    except Exception:
        traffic, traffic_note = None, ""
    loop_achieved = ice_bytes / (loop_ms * 1e6) if loop_ms > 0 else 0.0
    # the stream kernel on its own: the first launch of every graph replay is bracketed by CUDA events inside the graph
    # (HC_ICE_TIME_KERNEL=1) and those with all chromosomes still active are averaged; falls back to the loop-level figure
    full_bytes = (1.0 if packed else 4.0) * float(sum(n * n for n in sizes)) + (8.0 * float(info.overflow_cells) if packed else 0.0)
    kernel_ms = float(info.stream_full_ms) if int(info.stream_full_launches) > 0 else 0.0
    achieved = full_bytes / (kernel_ms * 1e6) if kernel_ms > 0 else loop_achieved
    line = {
        "metric": METRIC, "value": ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None,
        "dtype": "int32 counts (ICE streams them as u8 + overflow list, s32 tensor-core plane sums) / f64 weights" if packed else "int32 counts / f64 weights",
        "data": "synthetic",
        "config": {"workload": workload_name(args.pairs), "pairs": args.pairs, "bins": int(sum(sizes_all)),
                   "resolution": RES, "chromosomes": len(order),
                   "parallelism": "chromosomes LPT-sharded by N^2 over %d GPU(s), no collective" % world,
                   "host_format": "pinned columns: chromosome uint8 + mid-point int32 per mate (10 B/pair)",
                   "l2": "inputs larger than L2 (%.1f GB pair columns, %.2f GB int32 tiles on rank 0)"
                         % (16e-9 * n_local, 4e-9 * stage.batch.numel)},
        "throughput_Mpairs_per_s": args.pairs / (ms_step * 1e3),
        "e2e": {"value": ms_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d.item()),
                "d2h_bytes_per_step": int(d2h.item()), "steps": e2e_steps,
                "Mpairs_per_s": args.pairs / (ms_e2e * 1e3)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": ice_kernel, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "frac_of_nominal_8TBps": achieved / 8000.0,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650",
                     "algorithmic_bytes_per_launch": full_bytes if kernel_ms > 0 else ice_bytes / max(n_iter_launches, 1),
                     "launches": int(info.stream_full_launches) if kernel_ms > 0 else n_iter_launches,
                     "avg_launch_ms": kernel_ms if kernel_ms > 0 else loop_ms / max(n_iter_launches, 1),
                     "timing": ("CUDA events around the first stream-kernel launch of every graph replay of the last timed step in "
                                "which every chromosome was still active (event-record nodes inside the replayed graph)") if kernel_ms > 0
                               else "CUDA events around the whole iteration loop (stream + update kernels, launch gaps, polls)",
                     "loop": {"achieved": loop_achieved, "frac": loop_achieved / peak, "ms": loop_ms, "launches": n_iter_launches,
                              "note": "algorithmic bytes of all iterations / time of the whole loop incl. update kernels and gaps"},
                     "encoding": ("uint8 cells + overflow list, built once per call in %.3f ms (%d overflow cells)"
                                  % (float(info.pack_ms), int(info.overflow_cells))) if packed else "int32 tiles",
                     "traffic": traffic, "traffic_note": traffic_note},
        "breakdown": breakdown, "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.pairs, threads=1)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------
SAMPLE_CHROMS = ["21", "22"]


def _cpu_one_chrom(job):
    """oracle port on one chromosome: bin (NumPy restatement) + cis-only ICE restatement"""
    from oracle import cooler_ice, hichap_oracle as ho
    c, length, p1, p2 = job
    n = length // RES + 1
    z = np.zeros(p1.size, np.int32)
    t0 = time.perf_counter()
    M = ho.bin_local_dense(z, p1, z, p2, [n], RES)[0]
    rec = ho.dense_to_triu_records(M)
    t1 = time.perf_counter()
    w, st = cooler_ice.balance(rec["bin1"], rec["bin2"], rec["IF"].astype(np.int32), n, [0, n], cis_only=True)
    t2 = time.perf_counter()
    return c, t1 - t0, t2 - t1, int(st["iters"][0]), int(rec.size)


def _cpu_sample(pairs_total):
    from hichap_master_b200 import synth
    genome, order = c2_genome()
    shares = pair_shares(genome, order, pairs_total)
    jobs = []
    for c in SAMPLE_CHROMS:
        i = order.index(c)
        a, b = synth.cis_pairs(c, genome[c], int(shares[i]), 2000 + i)
        jobs.append((c, genome[c], a, b))
    n_sample = sum(j[2].size for j in jobs)
    return jobs, n_sample


def cpu_baseline(pairs_total, threads=1, sample=None):
    jobs, n_sample = sample if sample is not None else _cpu_sample(pairs_total)
    t0 = time.perf_counter()
    if threads > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(threads, len(jobs))) as pool:
            res = pool.map(_cpu_one_chrom, jobs)
        used = min(threads, len(jobs))
    else:
        res = [_cpu_one_chrom(j) for j in jobs]
        used = 1
    wall = time.perf_counter() - t0
    scale = pairs_total / float(n_sample)
    return {"value": wall * 1e3 * scale, "unit": UNIT, "cores": used, "kind": "port",
            "sample": ("oracle port (NumPy restatement of matrixBuilding.py:595-603 + cooler balance --cis-only) on "
                       "chromosomes %s of the same workload = %d of %d pairs, %.1f s measured; value is that time "
                       "x %.1f (pair ratio) -- extrapolated. The reference's own interpreted binning loop is "
                       "2.64 us/pair/resolution (SURVEY.md probe), slower than this vectorised port"
                       % ("+".join(SAMPLE_CHROMS), n_sample, pairs_total, wall, scale)),
            "measured_s": wall, "scale": scale,
            "per_chrom": [{"chrom": c, "bin_s": tb, "ice_s": ti, "iters": it, "nnz": nz} for c, tb, ti, it, nz in res],
            "host_cores": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the Python-2 reference cannot be
    installed) on the host cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    times = []
    cb = None
    sample = _cpu_sample(args.pairs)
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(args.pairs, threads=threads, sample=sample)
        if i >= args.warmup:
            times.append(cb["value"])
    v = float(np.mean(times))
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "int64 counts / f64 weights", "data": "synthetic",
            "config": {"workload": workload_name(args.pairs), "pairs": args.pairs, "resolution": RES},
            "cpu_baseline": cb,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_c4(args):
    """--config C4 (BASELINE.json configs[3]): hg19 genome-wide 10 kb matrix, pairs spread over the
    ranks as a parser would deliver them, distributed sort/exchange into row-block CSR shards, ICE
    with one NCCL allreduce of the marginal vector per iteration.  Prints one JSON line (not the
    driver's headline config)."""
    import torch
    import torch.distributed as dist
    from hichap_master_b200 import distributed as hd, kernels, matrixBuilding as mb, synth
    from hichap_master_b200.device import PairColumns

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    res = 10000 if args.config == "C4" else 5000
    genome, order = c2_genome()
    bins, total = mb._bins_from_genome(genome, res, [(c, c) for c in order])
    start = mb._start_table(bins, order, dev)
    chrom_bins = torch.tensor([genome[c] // res + 1 for c in order], dtype=torch.int32, device=dev)
    n_local = args.pairs // world
    c1, p1, c2, p2 = synth.genome_pairs_torch(genome, order, n_local, 4000 + rank, dev, trans_frac=0.25)
    pairs = PairColumns(c1, p1, c2, p2, device=dev)
    torch.cuda.synchronize()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    comm = hd.nccl_comm_from_process_group(dev) if world > 1 else None
    allreduce = (lambda t: dist.all_reduce(t)) if world > 1 else None
    out = {}
    for rep in range(2):                       # rep 0 warms NCCL, allocator and caches
        sync_all()
        t0 = time.perf_counter()
        if world > 1:
            csr, cuts = hd.build_row_block_csr(pairs, res, start, chrom_bins, total)
        else:
            csr, cuts = kernels.pairs_to_csr(pairs, res, start, chrom_bins, total, False), [0, total]
        sync_all()
        t1 = time.perf_counter()
        w, st = mb.ice_balance_sparse(csr, bins, cis_only=False, comm=comm, allreduce=allreduce)
        sync_all()
        t2 = time.perf_counter()
        nnz = torch.tensor([float(csr.nnz)], dtype=torch.float64, device=dev)
        mx = nnz.clone()
        if world > 1:
            dist.all_reduce(nnz); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        tm = torch.tensor([t1 - t0, t2 - t1, st["loop_ms"] / 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        out = dict(build_s=float(tm[0]), ice_s=float(tm[1]), loop_s=float(tm[2]), iters=st["iters"],
                   converged=st["converged"], nnz_stored=float(nnz), nnz_max_rank=float(mx))
        del csr
    if rank == 0:
        Z = (out["nnz_stored"] + total) / 2
        per_iter = out["loop_s"] / max(out["iters"], 1)
        print(json.dumps({
            "config": {"workload": "%s: hg19 genome-wide %d kb (%d bins), %d synthetic pairs (75%% cis / 25%% trans), "
                                   "row-block sharded CSR over %d GPU(s), one NCCL allreduce per ICE iteration"
                                   % (args.config, res // 1000, total, args.pairs, world)},
            "n_gpus": world, "binning_to_csr_s": out["build_s"], "ice_time_to_convergence_s": out["ice_s"],
            "ice_loop_s": out["loop_s"], "ice_iters": out["iters"], "converged": out["converged"],
            "nnz_upper": Z, "nnz_stored_total": out["nnz_stored"], "nnz_imbalance": out["nnz_max_rank"] * world / out["nnz_stored"],
            "ice_iter_ms": per_iter * 1e3, "ice_algorithmic_GBps_aggregate": (8 * Z + 24 * total) / per_iter / 1e9,
            "ice_streamed_GBps_per_gpu": 8 * out["nnz_max_rank"] / per_iter / 1e9}))
    if comm:
        kernels.nccl_comm_destroy(comm)
    if world > 1:
        dist.destroy_process_group()


def run_c3(args):
    """--config C3 (BASELINE.json configs[2]): allelic maternal/paternal matrices at 40 kb with the
    two-step (ICE of the traditional matrices + SNP-density / coverage) correction.  Single GPU;
    prints one JSON line with the stage breakdown."""
    import torch
    from hichap_master_b200 import _abi, kernels, matrixBuilding as mb, synth
    from hichap_master_b200.construction import _sub_batch
    from hichap_master_b200.device import DenseBatch, PairColumns

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    genome, order = c2_genome()
    nchrom = len(order)
    pairs_n = args.pairs if args.pairs != 400_000_000 else 100_000_000
    c1, p1, c2, p2 = synth.genome_pairs_torch(genome, order, pairs_n, 3, dev, trans_frac=0.0)
    g = torch.Generator(device=dev); g.manual_seed(33)
    cls = torch.rand(pairs_n, generator=g, device=dev)          # Bi 86 %, M_M 6 %, P_P 6 %, M_P 1 %, P_M 1 %
    mark = torch.multinomial(torch.tensor([0.30, 0.35, 0.35], device=dev), pairs_n, replacement=True, generator=g).to(torch.uint8)
    allp = PairColumns(c1, p1, c2, p2, device=dev)
    hap = []
    for lo, hi in ((0.86, 0.92), (0.92, 0.98)):
        sel = (cls >= lo) & (cls < hi)
        hap.append(PairColumns(c1[sel], p1[sel], c2[sel], p2[sel], mark[sel], device=dev))
    sizes = [genome[c] // RES + 1 for c in order]
    T = DenseBatch(sizes, dev)
    H = DenseBatch(sizes + sizes, dev)

    def ev(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b), out

    def bin_all():
        T.buf.zero_(); H.buf.zero_()
        kernels.bin_pairs_local_banded(allp, RES, T, check_bounds=False)
        for k in range(2):
            v = _sub_batch(H, k * nchrom, nchrom)
            kernels.bin_pairs_local(hap[k], RES, v, _abi.HC_BIN_SYM_BOTH, check_bounds=False)
            kernels.bin_pairs_local(hap[k], RES, v, _abi.HC_BIN_ONESIDED, check_bounds=False)

    def correct_all():
        return kernels.twostep_batch(T, H)

    res_t, outs = {}, None
    for rep in range(4):                         # first pass warms the allocator; report the last
        outs = None                              # release the 4.7 GB of corrected matrices before re-allocating
        res_t["binning_ms"], _ = ev(bin_all)
        res_t["ice_traditional_ms"], (w, st) = ev(lambda: kernels.ice_balance_dense(T, None, ignore_diags=1))
        res_t["two_step_ms"], outs = ev(correct_all)
    sq = float(sum(n * n for n in sizes))
    print(json.dumps({
        "config": {"workload": "C3: hg19 chr1-22,X allelic matrices at 40 kb, %d synthetic pairs (86%% bi-allelic, 6%% M_M, 6%% P_P; "
                               "Both/R1/R2 = 30/35/35%%), ICE of the traditional matrices + two-step correction of M and P" % pairs_n},
        "n_gpus": 1, **res_t, "total_ms": sum(res_t.values()), "ice_iters_max": int(max(st["iters"])),
        "two_step_algorithmic_GBps": 52.0 * sq / (res_t["two_step_ms"] * 1e6),
        "gap_rows_total": int(sum(g.size for g in outs[1]))}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=400_000_000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="kernel tuning runs only: no e2e leg (line is not a bench result)")
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5"], help="C2 = the driver's headline workload")
    args = ap.parse_args()
    if args.config == "C3":
        return run_c3(args)
    if args.config in ("C4", "C5"):
        if args.pairs == 400_000_000:
            args.pairs = 1_000_000_000 if args.config == "C4" else 2_000_000_000
        return run_c4(args)
    if args.impl == "reference":
        args.steps = min(args.steps, 3)
        args.warmup = min(args.warmup, 1)
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
