"""GPU parity tests of the sort path: hand-written radix sort, reduce-by-key into the symmetric
CSR, the upper-triangular records and ICE on the CSR, against NumPy / the oracle."""
import numpy as np
import pytest

from conftest import CHROMS, SMALL_GENOME, load_golden
from hichap_master_b200 import synth
from oracle import cooler_ice
from oracle import hichap_oracle as ho

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.fixture(scope="module")
def K(cuda_device):
    from hichap_master_b200 import kernels
    return kernels


@pytest.mark.parametrize("n,bits", [(0, 40), (1, 40), (5, 16), (4095, 40), (4096, 40), (4097, 34), (100_000, 40),
                                    (3_000_001, 40), (1_000_000, 64), (50_000, 8)])
def test_radix_sort_matches_numpy(K, cuda_device, n, bits):
    import torch
    rng = np.random.default_rng(n + bits)
    hi = (1 << bits) - 1 if bits < 64 else (1 << 63) - 1
    keys = rng.integers(0, hi, size=n, dtype=np.int64, endpoint=True)
    if n > 10:
        keys[::7] = keys[3]                       # duplicates
    t = torch.from_numpy(keys.copy()).to(cuda_device)
    out, _ = K.sort_keys_u64(t, bits if bits < 64 else 63)
    assert np.array_equal(out.cpu().numpy(), np.sort(keys))


def test_radix_sort_padding_keys_sort_last(K, cuda_device):
    import torch
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 1 << 34, size=20_000, dtype=np.int64)
    keys[rng.random(keys.size) < 0.3] = -1        # ~0 as uint64
    out, _ = K.sort_keys_u64(torch.from_numpy(keys.copy()).to(cuda_device), 34)
    exp = np.sort(keys.view(np.uint64)).view(np.int64)
    assert np.array_equal(out.cpu().numpy(), exp)


def small_case(seed, n_pairs, trans):
    genome = {c: l for c, l in SMALL_GENOME.items() if ho.chrom_passes(c, CHROMS)}
    order = ho.sort_chromosomes(genome)
    c1, p1, c2, p2 = synth.genome_pairs(genome, order, n_pairs, seed, trans_frac=trans)
    return genome, order, c1, p1, c2, p2


@pytest.mark.parametrize("res,cis_only,npairs", [(40000, False, 60_000), (40000, True, 60_000), (500000, False, 5_000),
                                                 (10000, False, 7), (40000, False, 0)])
def test_pairs_to_csr_matches_oracle(K, cuda_device, res, cis_only, npairs):
    from hichap_master_b200 import matrixBuilding as mb
    from hichap_master_b200.device import PairColumns
    genome, order, c1, p1, c2, p2 = small_case(5, max(npairs, 1), 0.2)
    c1, p1, c2, p2 = c1[:npairs], p1[:npairs], c2[:npairs], p2[:npairs]
    bins, csr = mb.bin_traditional_sparse(PairColumns(c1, p1, c2, p2), genome, res, cis_only=cis_only)
    table, total = ho.chro_bins(genome, res)
    assert bins == table and csr.nbins == total
    start = np.array([table[c][0] for c in order], np.int64)
    keep = (c1 == c2) if cis_only else np.ones(c1.size, bool)
    M = ho.bin_whole_dense(c1[keep], p1[keep], c2[keep], p2[keep], start, start, total, res)
    # full symmetric CSR
    rp = csr.row_ptr.cpu().numpy(); col = csr.col.cpu().numpy(); cnt = csr.cnt.cpu().numpy()
    x, y = np.nonzero(M)
    assert rp[-1] == x.size == col.size
    assert np.array_equal(np.repeat(np.arange(total), np.diff(rp)), x)
    assert np.array_equal(col, y) and np.array_equal(cnt, M[x, y])
    # the reference's per-block dictionary
    got = mb.WholeCsrToSparseDict(bins, csr)
    exp = ho.whole_matrix_to_sparse_dict(table, M)
    assert set(got) == set(exp)
    for k in exp:
        for f in ("bin1", "bin2", "IF"):
            assert np.array_equal(got[k][f], exp[k][f]), (k, f)


def csr_from_pixels(K, dev, b1, b2, cnt, n):
    """symmetric CSR tensors from upper-triangular pixels (host construction, test helper)"""
    import torch
    off = b1 != b2
    r = np.concatenate([b1, b2[off]]); c = np.concatenate([b2, b1[off]]); v = np.concatenate([cnt, cnt[off]])
    o = np.lexsort((c, r))
    r, c, v = r[o], c[o], v[o]
    rp = np.zeros(n + 1, np.int64)
    np.add.at(rp, r + 1, 1)
    rp = np.cumsum(rp)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    return K.SymCsr(t(rp, np.int64), t(c, np.int32), t(v, np.int32), n)


def check_weights(w, ref, what):
    assert np.array_equal(np.isnan(w), np.isnan(ref)), what + ": NaN mask differs"
    ok = ~np.isnan(ref)
    err = np.max(np.abs(w[ok] - ref[ok]) / np.abs(ref[ok])) if ok.any() else 0.0
    print("%s: max rel err %.3e over %d bins" % (what, err, int(ok.sum())))
    assert err < RTOL, what


@pytest.mark.parametrize("window", ["0", "1"])
def test_ice_csr_restated_golden(K, cuda_device, monkeypatch, window):
    """window=1: the opt-in stream kernel that stages a bias window with a TMA bulk copy (HC_CSR_WINDOW)."""
    monkeypatch.setenv("HC_CSR_WINDOW", window)
    g = load_golden("ice_restated.npz")
    off = g["chrom_offsets"]
    n = int(off[-1])
    csr = csr_from_pixels(K, cuda_device, g["gw_bin1"], g["gw_bin2"], g["gw_count"], n)
    bias, st = K.ice_balance_csr(csr, np.array([0, n]), chrom_off=off)
    check_weights(bias.cpu().numpy(), g["weight_gw"], "csr genome-wide")
    assert st["iters"] == int(g["iters_gw"]) and st["converged"]
    csr = csr_from_pixels(K, cuda_device, g["bin1"], g["bin2"], g["count"], n)
    bias, st = K.ice_balance_csr(csr, off, chrom_off=off)
    check_weights(bias.cpu().numpy(), g["weight_cis"], "csr cis-only")
    assert st["iters"] == [int(i) for i in g["iters_cis"]]
    np.testing.assert_allclose(st["scale"], g["scale_cis"], rtol=RTOL)


@pytest.mark.parametrize("kw", [dict(), dict(ignore_diags=0), dict(ignore_diags=3), dict(mad_max=0, min_nnz=0),
                                dict(max_iters=4)])
def test_ice_csr_end_to_end_vs_oracle(K, cuda_device, kw):
    """pairs -> sort path -> CSR -> ICE, against the pixel-based oracle."""
    from hichap_master_b200 import matrixBuilding as mb
    from hichap_master_b200.device import PairColumns
    genome, order, c1, p1, c2, p2 = small_case(9, 300_000, 0.25)
    res = 40000
    bins, csr = mb.bin_traditional_sparse(PairColumns(c1, p1, c2, p2), genome, res)
    w, st = mb.ice_balance_sparse(csr, bins, cis_only=False, **kw)
    table, total = ho.chro_bins(genome, res)
    start = np.array([table[c][0] for c in order], np.int64)
    b1 = p1.astype(np.int64) // res + start[c1]; b2 = p2.astype(np.int64) // res + start[c2]
    lo, hi = np.minimum(b1, b2), np.maximum(b1, b2)
    key, cnt = np.unique(lo * total + hi, return_counts=True)
    off = mb.chrom_offsets_from_bins(bins)
    ref, rst = cooler_ice.balance(key // total, key % total, cnt, total, off, cis_only=False, **kw)
    check_weights(w, ref, "e2e %r" % (kw,))
    assert st["iters"] == rst["iters"] and st["converged"] == rst["converged"]


def test_power_of_two_bins_last_diagonal_cell_survives_padding(K, cuda_device):
    """nbins == 2^col_bits (256 bins): the key of cell (255, 255) is all ones in the low 2*col_bits bits and would tie
    with the padding key of dropped pairs; the sort must look at one more bit (kernels.key_sort_bits)."""
    import torch
    from hichap_master_b200 import kernels
    from hichap_master_b200.device import PairColumns
    res, nb = 1000, 256
    assert kernels.key_sort_bits(nb) == 17 and kernels.key_sort_bits(255) == 16 and kernels.key_sort_bits(257) == 18
    # dropped pairs (filtered chromosome -1) come FIRST in the input: a stable sort on 16 bits would leave their
    # padding keys in front of the (255, 255) key
    c1 = np.array([-1, -1, -1, 0, 0, 0, 0], np.int32)
    p1 = np.array([5, 5, 5, 255_500, 255_100, 1_000, 254_000], np.int32)
    p2 = np.array([5, 5, 5, 255_900, 255_200, 2_000, 255_999], np.int32)
    start = torch.zeros(1, dtype=torch.int64, device=cuda_device)
    chrom_bins = torch.tensor([nb], dtype=torch.int32, device=cuda_device)
    csr = kernels.pairs_to_csr(PairColumns(c1, p1, c1, p2, device=cuda_device), res, start, chrom_bins, nb, False)
    b1, b2, v = (t.cpu().numpy() for t in kernels.csr_upper_records(csr))
    got = {(int(a), int(b)): int(c) for a, b, c in zip(b1, b2, v)}
    assert got == {(1, 2): 1, (254, 255): 1, (255, 255): 2}
