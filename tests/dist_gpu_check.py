"""Run under torchrun (one rank per GPU): the row-block sharded genome-wide path -- distributed
sort/exchange into per-rank CSR row blocks, ICE with ONE in-stream NCCL allreduce of the marginal
vector per iteration -- must reproduce the CPU oracle and the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import CHROMS, SMALL_GENOME  # noqa: E402
from hichap_master_b200 import distributed as hd, kernels, matrixBuilding as mb, synth  # noqa: E402
from hichap_master_b200.device import PairColumns  # noqa: E402
from oracle import cooler_ice, hichap_oracle as ho  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    genome = {c: l for c, l in SMALL_GENOME.items() if ho.chrom_passes(c, CHROMS)}
    order = ho.sort_chromosomes(genome)
    res = 20000
    c1, p1, c2, p2 = synth.genome_pairs(genome, order, 1_500_000, 17, trans_frac=0.25)
    mine = np.arange(c1.size) % world == rank                  # what a parser would hand this rank
    table, total = ho.chro_bins(genome, res)
    start = torch.tensor([table[c][0] for c in order], dtype=torch.int64, device=dev)
    chrom_bins = torch.tensor([genome[c] // res + 1 for c in order], dtype=torch.int32, device=dev)
    pairs = PairColumns(c1[mine], p1[mine], c2[mine], p2[mine], device=dev)
    csr, cuts = hd.build_row_block_csr(pairs, res, start, chrom_bins, total)
    comm = hd.nccl_comm_from_process_group(dev)
    w, st = mb.ice_balance_sparse(csr, table, cis_only=False, comm=comm, allreduce=lambda t: dist.all_reduce(t))
    # every rank holds the full, identical weight vector
    wt = torch.from_numpy(w.copy()).to(dev)
    lo, hi = wt.clone(), wt.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = torch.equal(torch.nan_to_num(lo), torch.nan_to_num(hi))
    # rows are complete and nnz is balanced
    nnz = torch.tensor([csr.nnz], dtype=torch.int64, device=dev)
    allnnz = [torch.zeros_like(nnz) for _ in range(world)]
    dist.all_gather(allnnz, nnz)
    ok = True
    msg = []
    if rank == 0:
        sgl = np.array([table[c][0] for c in order], np.int64)
        b1 = p1.astype(np.int64) // res + sgl[c1]; b2 = p2.astype(np.int64) // res + sgl[c2]
        a, b = np.minimum(b1, b2), np.maximum(b1, b2)
        key, cnt = np.unique(a * total + b, return_counts=True)
        off = mb.chrom_offsets_from_bins(table)
        ref, rst = cooler_ice.balance(key // total, key % total, cnt, total, off, cis_only=False)
        good = ~np.isnan(ref)
        err = float(np.max(np.abs(w[good] - ref[good]) / np.abs(ref[good])))
        nan_same = np.array_equal(np.isnan(w), np.isnan(ref))
        ok = (nan_same and err < 1e-6 and st["iters"] == rst["iters"] and bool(same))
        if not ok:
            bad = np.nonzero(np.isnan(w) != np.isnan(ref))[0]
            print("DIST_CHECK detail: nan_same=%s err_ok=%s iters %r vs %r same=%r mismatching bins %s w=%s ref=%s"
                  % (nan_same, err < 1e-6, st["iters"], rst["iters"], same, bad[:10], w[bad[:10]], ref[bad[:10]]))
        sizes = [int(x.item()) for x in allnnz]
        bal = max(sizes) / (sum(sizes) / world)
        msg = ["world=%d iters=%d (oracle %d) max_rel_err=%.2e identical_across_ranks=%s nnz/rank=%s imbalance=%.3f "
               "launches=%d loop_ms=%.2f cuts=%s" % (world, st["iters"], rst["iters"], err, same, sizes, bal,
                                                     st["launches"], st["loop_ms"], cuts)]
        # the cuts balance the distinct keys each rank holds before the exchange (a proxy for the stored entries);
        # rows are the sharding granularity: within 5 % of the mean, or at most two (dense) rows above it
        rows_nnz = np.bincount(np.concatenate([key // total, key % total]), minlength=total)
        ok = ok and (bal <= 1.05 or max(sizes) <= sum(sizes) / world + 2 * rows_nnz.max())
        print("DIST_CHECK", "OK" if ok else "FAIL", *msg)
    kernels.nccl_comm_destroy(comm)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
