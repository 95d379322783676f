"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: chromosome LPT sharding, row-block
boundaries, the key exchange plan, and the invariant the sharded ICE relies on -- per-rank
marginals of complete row blocks, summed by one allreduce, equal the single-process marginals
and drive the oracle to the same weights."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from hichap_master_b200 import shard
from hichap_master_b200.distributed import exchange_plan, row_cuts_from_counts
from oracle import cooler_ice


def test_lpt_assignment_balances_hg19():
    from hichap_master_b200 import synth
    sizes = [synth.HG19[c] // 40000 + 1 for c in [str(i) for i in range(1, 23)] + ["X"]]
    for world in (1, 2, 4, 8):
        parts = shard.chromosome_shards(sizes, world)
        assert sorted(i for p in parts for i in p) == list(range(23))
        loads = [sum(sizes[i] ** 2 for i in p) for p in parts]
        assert max(loads) <= 1.12 * sum(loads) / world or world == 8 and max(loads) <= 1.2 * sum(loads) / world
    assert shard.chromosome_shards(sizes, 8) == shard.chromosome_shards(sizes, 8)   # deterministic


def test_row_block_splits_balance_nnz():
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 50, size=1000)
    counts[:100] += 400
    for world in (1, 2, 3, 8):
        cuts = row_cuts_from_counts(counts, world)
        assert cuts[0] == 0 and cuts[-1] == 1000 and all(a <= b for a, b in zip(cuts, cuts[1:]))
        loads = [counts[a:b].sum() for a, b in zip(cuts, cuts[1:])]
        assert max(loads) <= counts.sum() / world + counts.max()
    assert row_cuts_from_counts(np.zeros(10, int), 4)[-1] == 10


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = load_golden("ice_restated.npz")
        off = g["chrom_offsets"]; n = int(off[-1])
        b1, b2, cnt = g["gw_bin1"], g["gw_bin2"], g["gw_count"]
        # symmetric (row, col, count) entries, the layout of the CSR path; each rank starts from
        # an arbitrary half of them (what a parser would hand it), sorted by row like the keys
        offd = b1 != b2
        r = np.concatenate([b1, b2[offd]]); c = np.concatenate([b2, b1[offd]]); v = np.concatenate([cnt, cnt[offd]])
        mine = np.arange(r.size) % world == rank
        r, c, v = r[mine], c[mine], v[mine]
        o = np.lexsort((c, r)); r, c, v = r[o], c[o], v[o]
        hist = np.bincount(r, minlength=n)
        total = torch.from_numpy(hist.copy()); dist.all_reduce(total)
        cuts = row_cuts_from_counts(total.numpy(), world)
        send = exchange_plan(hist, cuts)
        assert sum(send) == r.size
        recv = torch.empty(world, dtype=torch.int64)
        dist.all_to_all_single(recv, torch.tensor(send, dtype=torch.int64))
        recv = [int(x) for x in recv]
        payload = torch.from_numpy(np.stack([r, c, v], 1).astype(np.int64).ravel())
        inbox = torch.empty(3 * sum(recv), dtype=torch.int64)
        dist.all_to_all_single(inbox, payload, output_split_sizes=[3 * x for x in recv],
                               input_split_sizes=[3 * x for x in send])
        rr, cc, vv = inbox.numpy().reshape(-1, 3).T
        assert rr.size == 0 or (rr.min() >= cuts[rank] and rr.max() < cuts[rank + 1])   # complete rows only
        # sharded ICE: every rank iterates the same bias; marginals of the local rows + allreduce
        data = vv.astype(float); data[np.abs(rr - cc) < 1] = 0
        def marg_of(bias):
            m = torch.from_numpy(np.bincount(rr, weights=bias[rr] * bias[cc] * data, minlength=n))
            dist.all_reduce(m)
            return m.numpy()
        nnz = torch.from_numpy(np.bincount(rr, weights=(data != 0).astype(float), minlength=n)); dist.all_reduce(nnz)
        bias0, _ = cooler_ice.initial_bias(b1, b2, cnt, n, off, False, 1)
        assert np.array_equal(nnz.numpy() < 10, cooler_ice.marginalize(b1, b2, (np.where(np.abs(b1 - b2) < 1, 0, cnt) != 0).astype(float), n) < 10)
        bias = bias0.copy()
        for it in range(200):
            m = marg_of(bias); nz = m[m != 0]
            mm = m / nz.mean(); mm[mm == 0] = 1; bias /= mm
            if nz.var() < 1e-5:
                break
        scale = nz.mean(); bias[bias == 0] = np.nan; bias /= np.sqrt(scale)
        ok = ~np.isnan(g["weight_gw"])
        assert np.array_equal(np.isnan(bias), ~ok)
        assert np.max(np.abs(bias[ok] - g["weight_gw"][ok]) / g["weight_gw"][ok]) < 1e-9
        assert it + 1 == int(g["iters_gw"])
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_row_block_sharded_ice_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _entries_worker(rank, world, port, out):
    """Host logic of build_row_block_csr's exchange on gloo: reduced upper / lower entry lists per rank ->
    exchange_entry_lists -> (numpy stand-in for the sort + add-counts reduce) == the owner's rows of the full matrix."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hichap_master_b200 import kernels
        from hichap_master_b200.distributed import exchange_entry_lists
        g = load_golden("ice_restated.npz")
        n = int(g["chrom_offsets"][-1])
        b1, b2, cnt = g["gw_bin1"].astype(np.int64), g["gw_bin2"].astype(np.int64), g["gw_count"].astype(np.int64)
        cb, vb = kernels.key_col_bits(n), kernels.entry_cnt_bits(n)
        assert 2 * cb + vb <= 63
        # every rank holds a share of each cell's pairs (cells are split unevenly, some only on one rank)
        rng = np.random.default_rng(5)
        share = rng.integers(0, cnt + 1) if world == 2 else cnt
        mine = share if rank == 0 else cnt - share
        keep = mine > 0
        r, c, v = b1[keep], b2[keep], mine[keep]
        up = np.sort(((r << cb | c) << vb) | v)
        offd = r != c
        lo = np.sort((((c[offd] << cb) | r[offd]) << vb) | v[offd])
        inbox, cuts = exchange_entry_lists(torch.from_numpy(up), torch.from_numpy(lo), n)
        e = np.sort(inbox.numpy())
        cell, val = e >> vb, e & ((1 << vb) - 1)
        ucell, inv = np.unique(cell, return_inverse=True)
        uval = np.bincount(inv, weights=val).astype(np.int64) if cell.size else np.zeros(0, np.int64)
        # expected: rows [cuts[rank], cuts[rank+1]) of the symmetric matrix
        R = np.concatenate([b1, b2[b1 != b2]]); Cc = np.concatenate([b2, b1[b1 != b2]]); V = np.concatenate([cnt, cnt[b1 != b2]])
        sel = (R >= cuts[rank]) & (R < cuts[rank + 1])
        o = np.lexsort((Cc[sel], R[sel]))
        assert np.array_equal(ucell, ((R[sel] << cb) | Cc[sel])[o])
        assert np.array_equal(uval, V[sel][o])
        assert cuts[0] == 0 and cuts[-1] == n
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_entry_list_exchange_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entries_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
