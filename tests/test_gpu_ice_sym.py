"""GPU tests of the symmetric packed ICE path (hc_ice_sym.cu: upper-triangular 256 x 256 uint8 blocks, one persistent
dataflow kernel for the whole loop) beyond the variants in test_gpu_parity.py: determinism, equality with the other two
encodings on a chr21-sized problem, degenerate chromosomes."""
import numpy as np
import pytest

from hichap_master_b200 import synth
from oracle import cooler_ice
from oracle import hichap_oracle as ho

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.fixture(scope="module")
def mb(cuda_device):
    from hichap_master_b200 import matrixBuilding
    return matrixBuilding


def test_sym_path_chr21_and_chr22_sized_vs_oracle_and_other_encodings(mb, monkeypatch):
    mats = []
    for c, seed in (("21", 1), ("22", 2)):
        L = synth.HG19[c]
        p1, p2 = synth.cis_pairs(c, L, 1_500_000, seed=seed)
        n = L // 40000 + 1
        z = np.zeros(p1.size, np.int32)
        mats.append(ho.bin_local_dense(z, p1, z, p2, [n], 40000)[0])
    out = {}
    for mode in ("2", "1", "0"):
        monkeypatch.setenv("HC_ICE_PACKED", mode)
        out[mode] = mb.ice_balance_dense(mats)
    off = np.concatenate([[0], np.cumsum([m.shape[0] for m in mats])])
    b1, b2, cnt = [], [], []
    for m, lo in zip(mats, off[:-1]):
        x, y = np.nonzero(np.triu(m)); b1.append(x + lo); b2.append(y + lo); cnt.append(m[x, y])
    ref, rst = cooler_ice.balance(np.concatenate(b1), np.concatenate(b2), np.concatenate(cnt), int(off[-1]), off, cis_only=True)
    ok = ~np.isnan(ref)
    for mode, (w, st) in out.items():
        assert np.array_equal(np.isnan(w), np.isnan(ref)), mode
        err = float(np.max(np.abs(w[ok] - ref[ok]) / np.abs(ref[ok])))
        print("HC_ICE_PACKED=%s: iters %r (oracle %r), max rel err %.2e" % (mode, st["iters"], rst["iters"], err))
        assert st["iters"] == rst["iters"] and err < RTOL, mode
        np.testing.assert_allclose(st["scale"], rst["scale"], rtol=RTOL)
    # determinism: every sum of the dataflow kernel has a fixed order, whatever CTA ends up doing it
    monkeypatch.setenv("HC_ICE_PACKED", "2")
    w2, st2 = mb.ice_balance_dense(mats)
    assert np.array_equal(np.nan_to_num(w2), np.nan_to_num(out["2"][0])) and st2["iters"] == out["2"][1]["iters"]


def test_sym_path_degenerate_chromosomes(mb, monkeypatch):
    monkeypatch.setenv("HC_ICE_PACKED", "2")
    rng = np.random.default_rng(4)
    A = rng.poisson(5.0, size=(300, 300)); A = np.triu(A) + np.triu(A, 1).T
    Z = np.zeros((33, 33), np.int64)                       # nothing to balance: NaN weights, converged after one pass
    one = np.array([[7]])
    w, st = mb.ice_balance_dense([A, Z, one])
    x, y = np.nonzero(np.triu(A))
    ref, rst = cooler_ice.balance(x, y, A[x, y], 334, [0, 300, 333, 334], cis_only=True)
    assert np.array_equal(np.isnan(w), np.isnan(ref)) and np.isnan(w[300:]).all()
    ok = ~np.isnan(ref)
    assert np.max(np.abs(w[ok] - ref[ok]) / np.abs(ref[ok])) < RTOL
    assert st["iters"] == rst["iters"]
