"""Build-container-only tests: the oracle restatement against the LIVE reference executed
through oracle/ref_shim.py on fresh random inputs (skipped where /root/reference is absent)."""
import io

import numpy as np
import pytest

from conftest import SMALL_GENOME
from hichap_master_b200 import synth
from oracle import hichap_oracle as ho
from oracle import ref_shim

pytestmark = pytest.mark.reference


@pytest.mark.parametrize("seed,chroms", [(101, ["#", "X"]), (102, []), (103, ["2", "X"])])
def test_traditional_building_equals_reference(seed, chroms, small_genome_file):
    mod = ref_shim.load()
    genome = {c: l for c, l in SMALL_GENOME.items() if c != "M"}
    names = list(genome)
    c1, p1, c2, p2 = synth.genome_pairs(genome, names, 3000, seed, trans_frac=0.2)
    text = "".join(synth.valid23_lines(names, c1, p1, c2, p2))
    rw, rl = mod.TraditionalMatrixBuilding(io.StringIO(text), small_genome_file, [1000000], [100000], chroms)
    ow, ol = ho.traditional_matrix_building(text.splitlines(True), small_genome_file, [1000000], [100000], chroms)
    for ref, mine in ((rw[1000000], ow[1000000]), (rl[100000], ol[100000])):
        assert set(ref) == set(mine)
        for k in ref:
            for f in ("bin1", "bin2", "IF"):
                assert np.array_equal(ref[k][f], mine[k][f]), (k, f)


def test_bin_tables_equal_reference(small_genome_file):
    mod = ref_shim.load()
    for chroms in (["#", "X"], [], ["#"]):
        genome = ho.load_genome(small_genome_file, chroms)
        assert genome == mod.Load_Genome(small_genome_file, chroms)
        assert ho.sort_chromosomes(genome) == mod.Sort_Chromosomes(genome)
        for res in (40000, 500000, 4_800_000):
            assert ho.chro_bins(genome, res) == mod.Get_Chro_Bins(small_genome_file, res, chroms)
            assert ho.chro_bins_haplotypes(genome, res) == mod.Get_Chro_Bins_Haplotypes(small_genome_file, res, chroms)


@pytest.mark.parametrize("seed", [7, 8, 9])
def test_two_step_equals_reference(seed):
    mod = ref_shim.load()
    rng = np.random.default_rng(seed)
    n = 48
    dens = np.clip(rng.gamma(2.0, 0.3, size=n), 0, 1)
    if seed == 9:
        dens[:] = 1.0                                  # gap-free -> sum rule
    lam = 4.0 * dens[:, None] * dens[None, :]
    mm, pm = rng.poisson(lam), rng.poisson(0.7 * lam)
    tm = rng.poisson(6 * lam); tm = tm + tm.T + mm + pm
    r = mod.TwoStepCorrection(tm, mm, pm)
    o = ho.two_step_correction(tm, mm, pm)
    np.testing.assert_allclose(o[0], r[0], rtol=1e-12)
    np.testing.assert_allclose(o[1], r[1], rtol=1e-12)
    assert np.array_equal(o[2], r[2]) and np.array_equal(o[3], r[3])
    for fn_o, fn_r, arg in ((ho.gap_defined, mod.Gap_defined, mm), (ho.gap_defined_lowres, mod.Gap_definedLowRes, tm)):
        assert np.array_equal(fn_o(arg), fn_r(arg))
    S = mm / np.linspace(0.5, 1.5, n)[:, None]
    for gap in (np.array([]), np.array([1, 5, 6, 30])):
        np.testing.assert_array_equal(ho.trans2symmetry(S, gap), mod.Trans2symmetry(S, gap))
    np.testing.assert_array_equal(ho.correct_vc(S, 2 / 3), mod.Correct_VC(S, 2 / 3))
