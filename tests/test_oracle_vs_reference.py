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


def test_neighborhood_index_equals_reference():
    mod = ref_shim.load()
    for L in list(range(0, 26)) + [40, 100]:
        ri, rj = mod.GetNeighborhoodIndex(L)
        oi, oj = ho.neighborhood_index(L)
        assert list(oi) == list(ri) and list(oj) == list(rj), L


def _live_imputation(tmp_path, seed, whole_res, params, drop_mm_onesided=False, n=20000):
    """Run the reference's HaplotypeMatrixBuilding and the oracle's imputation on the same fresh beds."""
    import os
    from conftest import CHROMS
    from oracle import make_golden as mg
    from test_oracle_golden import _allelic_columns
    names, c1, p1, c2, p2, cls, mark = mg.imputation_inputs(seed=seed, n=n)
    if drop_mm_onesided:
        mark = np.where(cls == 1, 0, mark).astype(np.uint8)       # every M_M line is 'Both'
    gs = synth.write_genome_size(str(tmp_path / "genomeSize"), SMALL_GENOME)
    bed_dir = mg.write_allelic_beds(str(tmp_path), names, c1, p1, c2, p2, cls, mark)
    out_dir = str(tmp_path / "out")
    os.makedirs(out_dir)
    genome = ho.load_genome(gs, CHROMS)
    g = dict(names=np.array(names), c1=c1, p1=p1, c2=c2, p2=p2, cls=cls, mark=mark)
    order, k1, q1, k2, q2, cls, mark, keep = _allelic_columns(g, genome)
    files = {tag: tuple(a[cls == k] for a in (k1, q1, k2, q2, mark)) for tag, k in (("M_M", 1), ("P_P", 2))}
    starts = {}
    for res in whole_res:
        hb, _ = ho.chro_bins_haplotypes(genome, res)
        starts[res] = (np.array([hb["M" + c][0] for c in order]), np.array([hb["P" + c][0] for c in order]))

    def reference():
        return mg.run_reference_haplotype(bed_dir, gs, whole_res, [1000000], CHROMS, out_dir, imputation=params)[1]

    def oracle(ds):
        un = {res: ds["UnImputated_Whole"][res]["Matrix"].astype(np.int64) for res in whole_res}
        imp = {res: un[res].copy() for res in whole_res}
        for res in whole_res:
            for tag, own in (("M_M", 0), ("P_P", 1)):
                ho.bin_whole_onesided(*files[tag], starts[res][own], res, imp[res])
        return ho.impute_inter_chromosomal(un, imp, whole_res, starts, files, *params)
    return reference, oracle


@pytest.mark.parametrize("seed,whole_res,params", [(201, [500000], (1500000, 2, 0.7)),
                                                    (202, [1000000, 250000], (2000000, 1, 0.5)),
                                                    (203, [250000], (750000, 2, 0.95))])
def test_imputation_equals_reference_live(tmp_path, seed, whole_res, params):
    reference, oracle = _live_imputation(tmp_path, seed, whole_res, params)
    ds = reference()
    imp = oracle(ds)
    for res in whole_res:
        assert np.array_equal(imp[res], ds["Imputated_Whole"][res]["Matrix"]), res


def test_imputation_error_behaviour_equals_reference(tmp_path):
    """The reference dies with NameError when its P_P loop needs `M_M_sub` and the M_M loop never set it,
    and with IndexError when the stale window belongs to a coarser resolution."""
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    reference, oracle = _live_imputation(tmp_path / "a", 204, [500000], (1500000, 2, 0.7), drop_mm_onesided=True)
    with pytest.raises(NameError):
        reference()
    # the oracle raises the same way on the same inputs (un-imputed matrices do not matter for the error)
    zeros = {"UnImputated_Whole": {500000: {"Matrix": np.zeros((82, 82), np.int64)}}}
    with pytest.raises(NameError):
        oracle(zeros)
    reference, oracle = _live_imputation(tmp_path / "b", 205, [250000, 500000], (1500000, 2, 0.7))
    with pytest.raises(IndexError):
        reference()
    zeros = {"UnImputated_Whole": {250000: {"Matrix": np.zeros((160, 160), np.int64)},
                                   500000: {"Matrix": np.zeros((82, 82), np.int64)}}}
    with pytest.raises(IndexError):
        oracle(zeros)


@pytest.mark.parametrize("seed", [31, 32])
def test_genome_wide_and_intra_correction_equal_reference_live(seed, small_genome_file):
    """GenomeWideMatrixCorrection (:857-901) and IntraChromMatrixCorrection (:1026-1041) on fresh random matrices."""
    from conftest import CHROMS
    mod = ref_shim.load()
    rng = np.random.default_rng(seed)
    res = 500000
    bins, total = mod.Get_Chro_Bins(small_genome_file, res, CHROMS)
    hbins, htotal = mod.Get_Chro_Bins_Haplotypes(small_genome_file, res, CHROMS)
    dens_t = np.clip(rng.gamma(2.0, 0.4, size=total), 0.05, None)
    T = rng.poisson(40.0 * dens_t[:, None] * dens_t[None, :]); T = np.triu(T) + np.triu(T, 1).T
    dens_h = np.clip(rng.gamma(2.0, 0.3, size=htotal), 0.02, None)
    H = rng.poisson(6.0 * dens_h[:, None] * dens_h[None, :])          # imputed matrices are not symmetric
    H[5, :] = 0; H[:, 5] = 0
    ref = mod.GenomeWideMatrixCorrection(bins, hbins, T, H)
    mine = ho.genome_wide_matrix_correction(bins, hbins, T, H)
    np.testing.assert_allclose(mine, ref, rtol=1e-12, atol=0)
    tra, hap = {}, {}
    for c, (lo, hi) in bins.items():
        tra[c] = T[lo:hi + 1, lo:hi + 1]
        for h in "MP":
            a, b = hbins[h + c]
            hap[h + c] = H[a:b + 1, a:b + 1]
    rn, rg = mod.IntraChromMatrixCorrection(tra, hap)
    on, og = ho.intra_chrom_matrix_correction(tra, hap)
    assert set(rn) == set(on) and set(rg) == set(og)
    for k in rn:
        np.testing.assert_allclose(on[k], rn[k], rtol=1e-12, atol=0)
        assert np.array_equal(np.asarray(og[k], int), np.asarray(rg[k], int))
