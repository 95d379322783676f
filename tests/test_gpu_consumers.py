"""GPU tests of the downstream-consumer drop-ins (hichap_master_b200/structureFind.py) against
outputs of the reference's own StructureFind methods (tests/golden/consumers.npz) and the oracle on a larger case."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import hichap_oracle as ho

pytestmark = pytest.mark.gpu
RTOL = 1e-9        # sums of non-negative float64 terms in a different (fixed) order than NumPy's


@pytest.fixture(scope="module")
def sf(cuda_device):
    from hichap_master_b200 import matrixBuilding  # noqa: F401  (import order: matrixBuilding first)
    from hichap_master_b200 import structureFind
    return structureFind


def test_consumers_golden(sf):
    g = load_golden("consumers.npz")
    cM = sf.balanced_matrix(g["M"].astype(np.int64), g["weight"])
    np.testing.assert_allclose(cM, g["cM"], rtol=1e-15)
    assert np.array_equal(cM == 0, g["cM"] == 0)
    db, G, NG = sf.Distance_Decay(g["cM"], None)
    np.testing.assert_allclose(db, g["dd_auto"], rtol=RTOL)
    assert np.array_equal(G, g["dd_auto_G"]) and np.array_equal(NG, g["dd_auto_NG"])
    db2, G2, NG2 = sf.Distance_Decay(g["cM"], g["dd_given_G"])
    np.testing.assert_allclose(db2, g["dd_given"], rtol=RTOL)
    assert np.array_equal(NG2, g["dd_given_NG"])
    np.testing.assert_allclose(sf.Observed_Expected(g["cM"], g["dd_auto"].copy()), g["OE"], rtol=RTOL)
    for t in ("ttest", "chitest"):
        np.testing.assert_allclose(sf.Get_DI(g["cM"], g["gap"], g["window"], t), g["DI_" + t], rtol=RTOL, atol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(sf.bias_handle(g["weight"].reshape(-1, 1)), g["bias_handled"])
    b = sf.peak_biases(np.array([2.0, 0.0, np.nan, 0.5]))
    assert b[0] == 0.5 and b[1] == 0.0 and np.isnan(b[2]) and b[3] == 2.0


def test_consumers_larger_vs_oracle(sf):
    rng = np.random.default_rng(12)
    n = 700
    bias = np.exp(rng.normal(0, 0.3, n))
    d = np.abs(np.subtract.outer(np.arange(n), np.arange(n))) + 1.0
    M = rng.poisson(25.0 * bias[:, None] * bias[None, :] / d)
    M = np.triu(M) + np.triu(M, 1).T
    M[100:110, :] = 0; M[:, 100:110] = 0
    w = 1.0 / np.sqrt(np.maximum(M.sum(1), 1.0)); w[100:110] = np.nan
    cM = sf.balanced_matrix(M, w)
    np.testing.assert_allclose(cM, ho.balanced_matrix(M, w), rtol=1e-15)
    db, G, NG = sf.Distance_Decay(cM)
    rb, rG, rNG = ho.distance_decay(cM)
    np.testing.assert_allclose(db, rb, rtol=RTOL)
    assert np.array_equal(G, rG) and np.array_equal(NG, rNG)
    np.testing.assert_allclose(sf.Observed_Expected(cM, db.copy()), ho.observed_expected(cM, rb), rtol=RTOL)
    window = rng.integers(2, 30, n)
    for t in ("ttest", "chitest"):
        np.testing.assert_allclose(sf.Get_DI(cM, G, window, t), ho.get_di(cM, G, window, t), rtol=1e-8, atol=1e-12)


def test_gap_npz_round_trip(sf, tmp_path):
    """the gap NPZ the stage writes (matrixBuilding.py:1616-1617) is read back the way CallPeaks does (:1988-1992)"""
    gaps = {"40000": {"M1": np.array([3, 4, 9]), "P1": np.array([])}}
    path = str(tmp_path / "S_Imputated_Gap.npz")
    np.savez(path, **gaps)
    got = sf.load_gap(path, 40000, ["M1", "P1"])
    assert np.array_equal(got["M1"], [3, 4, 9]) and got["P1"].size == 0
