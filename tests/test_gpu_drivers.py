"""GPU tests of the stage drivers (same signatures as the reference): HaplotypeMatrixBuilding run
end to end on the five allelic beds against the golden vectors produced by the reference's own
HaplotypeMatrixBuilding (matrixBuilding.py:1044-1638), and TraditionalMatrixConstruction with two
replicates against the oracle."""
import os

import numpy as np
import pytest

from conftest import CHROMS, SMALL_GENOME, SORTED_SMALL, load_golden, unflatten
from hichap_master_b200 import synth
from oracle import cooler_ice
from oracle import hichap_oracle as ho

pytestmark = pytest.mark.gpu
RTOL = 1e-6
CLASS_FILES = ["Bi_Allelic", "M_M", "P_P", "M_P", "P_M"]


def write_allelic_beds(bed_dir, g, scale=1):
    os.makedirs(bed_dir, exist_ok=True)
    names = [str(x) for x in g["names"]]
    for k, tag in enumerate(CLASS_FILES):
        sel = g["cls"] == k
        mk = g["mark"][sel] if tag in ("M_M", "P_P") else None
        with open(os.path.join(bed_dir, "S_Valid_%s.bed" % tag), "w") as fh:
            fh.writelines(synth.allelic_lines(names, g["c1"][sel], g["p1"][sel], g["c2"][sel], g["p2"][sel], mk))
    return bed_dir


def test_haplotype_matrix_building_golden(cuda_device, tmp_path, small_genome_file):
    from hichap_master_b200 import matrixBuilding as mb
    g = load_golden("allelic_small.npz")
    bed_dir = write_allelic_beds(str(tmp_path / "beds"), g)
    out_dir = str(tmp_path / "out"); os.makedirs(out_dir)
    prefix, ds = mb.HaplotypeMatrixBuilding(out_dir, bed_dir, small_genome_file, [500000], [80000],
                                            10000000, 2, 0.9, CHROMS)
    assert prefix == str(g["prefix"]) == "S_"
    for key in ("Tradition_Whole", "UnImputated_Whole", "Imputated_Whole"):
        ref = unflatten(g, key)["500000"]
        assert np.array_equal(ds[key][500000]["Matrix"], ref["Matrix"]), key
        assert {k: tuple(v) for k, v in ds[key][500000]["Bins"].items()} == {k: tuple(int(x) for x in v) for k, v in ref["Bins"].items()}
    for key in ("Tradition_Local", "UnImputated_Local", "Imputated_Local"):
        ref = unflatten(g, key)["80000"]
        assert set(ds[key][80000]) == set(ref)
        for c in ref:
            assert np.array_equal(ds[key][80000][c], ref[c]), (key, c)
    # gap NPZ: same key structure the reference writes and StructureFind.py:1988-1992 reads
    gap = np.load(os.path.join(out_dir, "S_Imputated_Gap.npz"), allow_pickle=True)
    assert gap.files == ["80000"]
    got, ref = gap["80000"].item(), unflatten(g, "Gap")["80000"]
    assert set(got) == set(ref)
    for k in ref:
        assert np.array_equal(np.asarray(got[k], np.int64), ref[k]), k
    # corrected matrices as handed to the cool writer (float records)
    st = np.load(os.path.join(out_dir, "S_Imputated_Haplotype_Multi.npz"), allow_pickle=True)
    refb = unflatten(g, "Balanced_Local")["80000"]
    for k in refb:
        rec = st["80000|%s" % k]
        assert np.array_equal(rec["bin1"], refb[k]["bin1"]) and np.array_equal(rec["bin2"], refb[k]["bin2"])
        np.testing.assert_allclose(rec["IF"], refb[k]["IF"], rtol=RTOL)
    # genome-wide corrected matrix: intra block of the first chromosome = upper triangle of the golden
    gw = g["GenomeWide|500000"]
    hb = {k: tuple(int(x) for x in v) for k, v in unflatten(g, "Imputated_Whole")["500000"]["Bins"].items()}
    lo, hi = hb["M1"]
    exp = ho.dense_to_triu_records(gw[lo:hi + 1, lo:hi + 1])
    rec = st["500000|M1"]
    assert np.array_equal(rec["bin1"], exp["bin1"]) and np.array_equal(rec["bin2"], exp["bin2"])
    np.testing.assert_allclose(rec["IF"], exp["IF"], rtol=RTOL)
    # traditional store carries the ICE weights
    tr = np.load(os.path.join(out_dir, "S_Traditional_Multi.npz"), allow_pickle=True)
    assert tr["weight|80000"].shape == (sum(SMALL_GENOME[c] // 80000 + 1 for c in SORTED_SMALL),)
    assert os.path.isfile(os.path.join(out_dir, "Hap_genomeSize"))
    # a missing bed is the reference's Exception (matrixBuilding.py:1075)
    os.remove(os.path.join(bed_dir, "S_Valid_P_M.bed"))
    with pytest.raises(Exception, match="Missing file P_M.bed"):
        mb.HaplotypeMatrixBuilding(out_dir, bed_dir, small_genome_file, [500000], [80000], chroms=CHROMS)


def test_haplotype_construction_merges_replicates(cuda_device, tmp_path, small_genome_file):
    from hichap_master_b200 import matrixBuilding as mb
    g = load_golden("allelic_small.npz")
    reps = [write_allelic_beds(str(tmp_path / ("rep%d" % i)), g) for i in (1, 2)]
    out = str(tmp_path / "ws"); os.makedirs(out)
    mb.HaplotypeMatrixConstruction(out, reps, small_genome_file, [500000], [80000], chroms=CHROMS)
    one = np.load(os.path.join(out, "Cooler", "S_UnImputated_Haplotype_Multi.npz"), allow_pickle=True)
    two = np.load(os.path.join(out, "Cooler", "Merged_UnImputated_Haplotype_Multi.npz"), allow_pickle=True)
    for k in ("80000|M1", "80000|PX", "500000|M1_P2"):
        assert np.array_equal(two[k]["bin1"], one[k]["bin1"]) and np.array_equal(two[k]["IF"], 2 * one[k]["IF"])
    # doubled counts: the two-step correction of the merged data equals the oracle on 2x matrices
    tra = {c: 2 * v for c, v in unflatten(g, "Tradition_Local")["80000"].items()}
    hap = {c: 2 * v for c, v in unflatten(g, "Imputated_Local")["80000"].items()}
    nor, gaps = ho.intra_chrom_matrix_correction(tra, hap)
    st = np.load(os.path.join(out, "Cooler", "Merged_Imputated_Haplotype_Multi.npz"), allow_pickle=True)
    for k in ("M1", "P10"):
        exp = ho.dense_to_triu_records(nor[k])
        np.testing.assert_allclose(st["80000|" + k]["IF"], exp["IF"], rtol=RTOL)
    gap = np.load(os.path.join(out, "Cooler", "Merged_Imputated_Gap.npz"), allow_pickle=True)["80000"].item()
    for k in gaps:
        assert np.array_equal(np.asarray(gap[k], np.int64), gaps[k])


def test_traditional_construction_two_replicates(cuda_device, tmp_path, small_genome_file):
    from hichap_master_b200 import matrixBuilding as mb
    genome = {c: l for c, l in SMALL_GENOME.items() if c != "M"}
    names = list(genome)
    reps, cols = [], []
    for i, seed in enumerate((51, 52)):
        d = tmp_path / ("rep%d" % i); d.mkdir()
        c1, p1, c2, p2 = synth.genome_pairs(genome, names, 150_000, seed, trans_frac=0.2)
        # two part files per replicate, like the chunked *_Valid.bed outputs of the filtering stage
        half = c1.size // 2
        for j, sl in enumerate((slice(0, half), slice(half, None))):
            with open(d / ("R%d_part%d_Valid.bed" % (i, j)), "w") as fh:
                fh.writelines(synth.valid23_lines(names, c1[sl], p1[sl], c2[sl], p2[sl]))
        reps.append(str(d)); cols.append((c1, p1, c2, p2))
    out = str(tmp_path / "ws"); os.makedirs(out)
    files = mb.TraditionalMatrixConstruction(out, reps, small_genome_file, [500000], [40000], CHROMS, balance=True)
    assert [os.path.basename(f) for f in files] == ["R0_part0_Multi.npz", "R1_part0_Multi.npz", "Merged_Multi.npz"] or len(files) == 3
    merged = np.load(os.path.join(out, "Cooler", "Merged_Multi.npz"), allow_pickle=True)
    g = ho.load_genome(small_genome_file, CHROMS)
    order = ho.sort_chromosomes(g)
    lines = []
    for c1, p1, c2, p2 in cols:
        lines.extend(synth.valid23_lines(names, c1, p1, c2, p2))
    whole, local = ho.traditional_matrix_building(lines, small_genome_file, [500000], [40000], CHROMS)
    for k, rec in local[40000].items():
        got = merged["40000|%s" % k]
        assert all(np.array_equal(got[f], rec[f]) for f in ("bin1", "bin2", "IF")), k
    for k, rec in whole[500000].items():
        got = merged["500000|%s" % k]
        assert all(np.array_equal(got[f], rec[f]) for f in ("bin1", "bin2", "IF")), k
    # weights of the merged store == cooler restatement (cis-only for the local resolution)
    sizes = [g[c] // 40000 + 1 for c in order]
    off = np.concatenate([[0], np.cumsum(sizes)])
    b1 = np.concatenate([local[40000][c]["bin1"] + off[i] for i, c in enumerate(order)])
    b2 = np.concatenate([local[40000][c]["bin2"] + off[i] for i, c in enumerate(order)])
    cnt = np.concatenate([local[40000][c]["IF"] for c in order]).astype(np.int64)
    ref, _ = cooler_ice.balance(b1, b2, cnt, int(off[-1]), off, cis_only=True)
    w = merged["weight|40000"]
    assert np.array_equal(np.isnan(w), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.max(np.abs(w[ok] - ref[ok]) / ref[ok]) < RTOL


def _write_replicates(tmp_path, n_pairs=150_000, seeds=(51, 52)):
    genome = {c: l for c, l in SMALL_GENOME.items() if c != "M"}
    names = list(genome)
    reps = []
    for i, seed in enumerate(seeds):
        d = tmp_path / ("rep%d" % i); d.mkdir()
        c1, p1, c2, p2 = synth.genome_pairs(genome, names, n_pairs, seed, trans_frac=0.2)
        with open(d / ("R%d_Valid.bed" % i), "w") as fh:
            fh.writelines(synth.valid23_lines(names, c1, p1, c2, p2))
        reps.append(str(d))
    return reps


def _stores_equal(a, b, what):
    keys = sorted(k for k in a if "|" in k and not k.startswith(("weight", "bins")))
    assert keys == sorted(k for k in b if "|" in k and not k.startswith(("weight", "bins"))), what
    for k in keys:
        assert all(np.array_equal(a[k][f], b[k][f]) for f in ("bin1", "bin2", "IF")), (what, k)
    for k in (k for k in a if k.startswith("weight|")):
        wa, wb = a[k], b[k]
        assert np.array_equal(np.isnan(wa), np.isnan(wb)), (what, k)
        ok = ~np.isnan(wa)
        assert np.max(np.abs(wa[ok] - wb[ok]) / np.abs(wa[ok])) < RTOL, (what, k)


def test_traditional_construction_sort_path_equals_dense_path(cuda_device, tmp_path, small_genome_file, monkeypatch):
    """The drop-in picks the sort path (symmetric CSR, CSR ICE) by itself when the dense tiles exceed the budget
    (genome-wide 10 kb would be 369 GB): with the budget forced to zero the stores must equal the dense-path
    stores -- records bit-exact, weights within tolerance -- replicate merge included."""
    from hichap_master_b200 import matrixBuilding as mb
    from hichap_master_b200.construction import MatrixStore
    reps = _write_replicates(tmp_path)
    dense = str(tmp_path / "dense"); os.makedirs(dense)
    mb.TraditionalMatrixConstruction(dense, reps, small_genome_file, [500000], [40000], CHROMS)
    monkeypatch.setenv("HC_DENSE_BUDGET_GB", "0")
    sparse = str(tmp_path / "sparse"); os.makedirs(sparse)
    mb.TraditionalMatrixConstruction(sparse, reps, small_genome_file, [500000], [40000], CHROMS)
    for f in ("R0_Multi.npz", "R1_Multi.npz", "Merged_Multi.npz"):
        a, b = MatrixStore.load(os.path.join(dense, "Cooler", f)), MatrixStore.load(os.path.join(sparse, "Cooler", f))
        _stores_equal(a, b, f)
    # and the callable itself
    text = open(os.path.join(reps[0], "R0_Valid.bed")).read()
    import io
    w1, l1 = mb.TraditionalMatrixBuilding(io.StringIO(text), small_genome_file, [500000], [40000], CHROMS)
    monkeypatch.delenv("HC_DENSE_BUDGET_GB")
    w0, l0 = mb.TraditionalMatrixBuilding(io.StringIO(text), small_genome_file, [500000], [40000], CHROMS)
    for lib0, lib1 in ((w0[500000], w1[500000]), (l0[40000], l1[40000])):
        assert set(lib0) == set(lib1)
        for k in lib0:
            assert all(np.array_equal(lib0[k][f], lib1[k][f]) for f in ("bin1", "bin2", "IF")), k


def test_dense_entry_points_refuse_impossible_allocations(cuda_device, small_genome_file, tmp_path):
    """TraditionalMatrixInAllelic returns dense matrices by contract: a resolution whose tiles cannot fit raises
    MemoryError instead of dying inside the allocator."""
    from hichap_master_b200 import matrixBuilding as mb
    import io
    big = tmp_path / "genomeSize_big"
    big.write_text("chr1\t249250621\nchr2\t243199373\n")
    with pytest.raises(MemoryError):
        mb.TraditionalMatrixInAllelic(io.StringIO("chr1\t100\tchr1\t5000\n"), str(big), [500], [], CHROMS)
