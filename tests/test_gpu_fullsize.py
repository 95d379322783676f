"""Full-size (BASELINE.json C2 shape: hg19 chr1-22,X @ 40 kb, 400 M pairs) GPU checks through
size-independent properties -- the oracle cannot run at this size in test time:
count conservation, symmetry, no writes into the row padding, flat balanced marginals after ICE,
idempotence of a second balancing pass, and sortedness + checksum for the radix sort."""
import numpy as np
import pytest

from hichap_master_b200 import synth

pytestmark = pytest.mark.gpu

RES = 40000
PAIRS = 400_000_000


@pytest.fixture(scope="module")
def c2(cuda_device):
    import torch
    from hichap_master_b200 import kernels
    from hichap_master_b200.device import DenseBatch, PairColumns
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a large-memory GPU")
    genome = {c: l for c, l in synth.HG19.items() if c not in ("Y", "M")}
    order = [str(i) for i in range(1, 23)] + ["X"]
    c1, p1, c2_, p2 = synth.genome_pairs_torch(genome, order, PAIRS, 2, cuda_device, trans_frac=0.0)
    pairs = PairColumns(c1, p1, c2_, p2, device=cuda_device)
    sizes = [genome[c] // RES + 1 for c in order]
    batch = DenseBatch(sizes, cuda_device)
    kernels.bin_pairs_local_banded(pairs, RES, batch)
    return dict(pairs=pairs, batch=batch, sizes=sizes, order=order, genome=genome)


def test_fullsize_binning_conserves_counts_and_is_symmetric(c2, cuda_device):
    import torch
    b = c2["batch"]
    total_upper = 0
    for i, n in enumerate(c2["sizes"]):
        v = b.view(i)
        M = v[:, :n]
        assert int(v[:, n:].abs().sum().item()) == 0                      # row padding untouched
        if i in (0, 20, 22):                                              # chr1, chr21, chrX: full symmetry check
            assert torch.equal(M, M.t())
        total_upper += int(torch.triu(M).sum(dtype=torch.int64).item())
    assert total_upper == PAIRS                                           # every pair counted exactly once
    # per-chromosome pair counts match the input
    per = torch.bincount(c2["pairs"].c1.long(), minlength=len(c2["sizes"])).cpu().numpy()
    for i, n in enumerate(c2["sizes"]):
        if i in (0, 7, 22):
            assert int(torch.triu(b.view(i)[:, :n]).sum(dtype=torch.int64).item()) == int(per[i])


def test_fullsize_banded_equals_direct_binning(c2, cuda_device):
    import torch
    from hichap_master_b200 import kernels
    from hichap_master_b200.device import DenseBatch
    direct = DenseBatch(c2["sizes"], cuda_device)
    kernels.bin_pairs_local(c2["pairs"], RES, direct)
    assert torch.equal(direct.buf, c2["batch"].buf)                       # bit-exact, all 1.2 GB


def test_fullsize_ice_balances_every_chromosome(c2, cuda_device):
    import torch
    from hichap_master_b200 import kernels
    b = c2["batch"]
    w, st = kernels.ice_balance_dense(b, None, ignore_diags=1)
    assert all(st["converged_per_chrom"]) and max(st["iters"]) < 200
    off = b.h_bin_off
    wt = torch.from_numpy(np.nan_to_num(w)).to(cuda_device)
    for i in (0, 11, 20, 22):
        n = c2["sizes"][i]
        wi = wt[off[i]:off[i + 1]]
        M = b.view(i)[:, :n].to(torch.float64)
        M.fill_diagonal_(0.0)                                             # --ignore-diags 1
        marg = (M * wi[None, :]).sum(1) * wi
        kept = ~torch.isnan(torch.from_numpy(w[off[i]:off[i + 1]]).to(cuda_device))
        m = marg[kept]
        assert abs(float(m.mean()) - 1.0) < 1e-3 and float(m.var(unbiased=False)) < 1e-4
        # scale/var bookkeeping is consistent with the definition
        assert st["scale"][i] > 0
    # determinism: a second run gives bit-identical weights (no atomics in the iteration)
    w2, st2 = kernels.ice_balance_dense(b, None, ignore_diags=1)
    assert st2["iters"] == st["iters"]
    assert np.array_equal(np.nan_to_num(w2), np.nan_to_num(w)) and np.array_equal(np.isnan(w2), np.isnan(w))


def test_fullsize_counts_and_weights_equal_the_oracle_on_chr1_chr21_chr22(c2, cuda_device):
    """Full-size parity proper: the pairs of chr1 (largest, slowest to converge), chr21 and chr22 of the 400 M-pair
    workload go through the CPU oracle (NumPy binning + cooler-balance restatement); the GPU tiles must be
    bit-exact and the packed-encoding ICE weights (overflow cells, ~80 iterations) within 1e-6 with the same
    NaN mask and iteration count."""
    import torch
    from hichap_master_b200 import kernels
    from oracle import cooler_ice, hichap_oracle as ho
    b = c2["batch"]
    w, st = kernels.ice_balance_dense(b, None, ignore_diags=1)
    off = b.h_bin_off
    pairs = c2["pairs"]
    for i in (21, 20, 0):
        sel = pairs.c1 == i
        p1, p2 = pairs.p1[sel].cpu().numpy(), pairs.p2[sel].cpu().numpy()
        n = c2["sizes"][i]
        z = np.zeros(p1.size, np.int32)
        M = ho.bin_local_dense(z, p1, z, p2, [n], RES)[0]
        assert np.array_equal(b.to_numpy(i), M), "chromosome %s: counts differ" % c2["order"][i]
        ref, rst = cooler_ice.balance_dense(M, cis_only=True)
        wi = w[off[i]:off[i + 1]]
        assert np.array_equal(np.isnan(wi), np.isnan(ref)), c2["order"][i]
        ok = ~np.isnan(ref)
        err = float(np.max(np.abs(wi[ok] - ref[ok]) / np.abs(ref[ok])))
        print("chr%s: %d pairs, %d bins, iters %d (oracle %d), max rel err %.2e, max count %d"
              % (c2["order"][i], p1.size, n, st["iters"][i], rst["iters"][0], err, int(M.max())))
        assert st["iters"][i] == rst["iters"][0]
        assert err < 1e-6
        np.testing.assert_allclose(st["scale"][i], rst["scale"][0], rtol=1e-6)


def test_c4_shaped_csr_ice_equals_the_oracle(cuda_device):
    """BASELINE.json configs[3] shape at a size the oracle finishes in a minute: genome-wide 10 kb matrix of
    chr1-3 (69 050 bins), 50 M pairs with 25 % trans, sort path -> symmetric CSR -> genome-wide ICE; upper-triangular
    records bit-exact against NumPy, weights against the cooler-balance restatement."""
    import torch
    from hichap_master_b200 import kernels, matrixBuilding as mb
    from hichap_master_b200.device import PairColumns
    from oracle import cooler_ice
    genome = {c: synth.HG19[c] for c in ("1", "2", "3")}
    order = ["1", "2", "3"]
    res = 10000
    c1, p1, c2_, p2 = synth.genome_pairs_torch(genome, order, 50_000_000, 44, cuda_device, trans_frac=0.25)
    bins, csr = mb.bin_traditional_sparse(PairColumns(c1, p1, c2_, p2, device=cuda_device), genome, res)
    total = csr.nbins
    w, st = mb.ice_balance_sparse(csr, bins)
    b1, b2, v = (t.cpu().numpy() for t in kernels.csr_upper_records(csr))
    start = np.array([bins[c][0] for c in order], np.int64)
    h1 = p1.cpu().numpy().astype(np.int64) // res + start[c1.cpu().numpy()]
    h2 = p2.cpu().numpy().astype(np.int64) // res + start[c2_.cpu().numpy()]
    key, cnt = np.unique(np.minimum(h1, h2) * total + np.maximum(h1, h2), return_counts=True)
    assert np.array_equal(b1.astype(np.int64) * total + b2, key) and np.array_equal(v, cnt)     # counts bit-exact
    off = mb.chrom_offsets_from_bins(bins)
    ref, rst = cooler_ice.balance(key // total, key % total, cnt, total, off, cis_only=False)
    assert np.array_equal(np.isnan(w), np.isnan(ref))
    ok = ~np.isnan(ref)
    err = float(np.max(np.abs(w[ok] - ref[ok]) / np.abs(ref[ok])))
    print("C4-shaped: %d bins, %d upper-triangle pixels, iters %d (oracle %d), max rel err %.2e" % (total, key.size, st["iters"], rst["iters"], err))
    assert st["iters"] == rst["iters"] and st["converged"] == rst["converged"]
    assert err < 1e-6


def test_fullsize_radix_sort_sortedness_and_checksum(cuda_device):
    import torch
    from hichap_master_b200 import kernels
    g = torch.Generator(device=cuda_device); g.manual_seed(5)
    n = 200_000_000
    keys = torch.randint(0, 1 << 40, (n,), generator=g, device=cuda_device, dtype=torch.int64)
    csum, cxor = int(keys.sum().item()), int(torch.bitwise_xor(keys[: n // 2], keys[n // 2:]).sum().item())
    out, _ = kernels.sort_keys_u64(keys.clone(), 40)
    assert bool((out[1:] >= out[:-1]).all())                              # sorted
    assert int(out.sum().item()) == csum                                  # same multiset (checksum)
    assert int(out[0].item()) == int(keys.min().item()) and int(out[-1].item()) == int(keys.max().item())
    # spot check against torch's sort on a slice of distinct high bits
    sel = keys[keys < (1 << 30)]
    exp = torch.sort(sel).values
    assert torch.equal(out[: exp.numel()], exp)
    del cxor
