"""TEST INFRASTRUCTURE ONLY -- generates the golden vectors under ``tests/golden/`` by running
the UNMODIFIED reference (``/root/reference/HiCHap/matrixBuilding.py`` through
``oracle/ref_shim.py``) on small seeded inputs.  Run in the build container only:

    python -m oracle.make_golden

Fixtures (all < 1 MB):
  traditional_small.npz  TraditionalMatrixBuilding (matrixBuilding.py:528-613) on 23-column text
  allelic_small.npz      HaplotypeMatrixBuilding (:1044-1638) end to end with the `cat` pipes,
                         the cool writer and the `cooler balance` subprocess stubbed out:
                         raw traditional / un-imputed / imputed matrices, two-step corrected
                         matrices and gap lists, GenomeWideMatrixCorrection output
  twostep_cases.npz      TwoStepCorrection (:984-1023) on a gap-free and a gappy triple
  consumers.npz          StructureFind.py: Distance_Decay, the O/E loop of Get_PCA, Get_DI, bias_handle
  building_blocks.npz    Coverage_M, Gap_defined(+LowRes), Non_Gap_Defined, Trans2symmetry(+LowRes), Correct_VC
  ice_restated.npz       NOT from the reference (its ICE is the un-vendored `cooler`): outputs of
                         oracle/cooler_ice.py, kept as a regression anchor ("parity unpinned")
"""
from __future__ import annotations

import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import cooler_ice, ref_shim  # noqa: E402
from hichap_master_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

SMALL_GENOME = {"1": 6_010_000, "2": 4_800_000, "10": 3_333_333, "X": 5_000_001, "Y": 2_000_000, "M": 16571}
CHROMS = ["#", "X"]


def small_pairs(n, seed, trans_frac=0.15):
    """Pairs over the small genome, including chromosomes the filter drops (Y, M)."""
    order = list(SMALL_GENOME.keys())            # file order, NOT the sorted order
    big = {c: l for c, l in SMALL_GENOME.items() if c != "M"}
    names = [c for c in order if c != "M"]
    c1, p1, c2, p2 = synth.genome_pairs(big, names, n, seed, trans_frac)
    return names, c1, p1, c2, p2


class _FakePopen:
    """`cat files` -> text stream; anything else (the `cooler balance` shell string) -> no-op."""

    def __init__(self, cmd, **kw):
        if isinstance(cmd, (list, tuple)) and cmd and cmd[0] == "cat":
            self.stdout = io.StringIO("".join(open(f).read() for f in cmd[1:]))
        else:
            self.stdout = io.StringIO("")

    def communicate(self):
        return "", ""


def run_reference_haplotype(bed_dir, genome_size, whole_res, local_res, chroms, out_dir,
                            imputation=(10000000, 2, 0.9)):
    """Execute the reference's HaplotypeMatrixBuilding with I/O side effects stubbed."""
    mod = ref_shim.load()
    saved = (mod.subprocess, mod.NPZ2Cooler)
    calls = []
    mod.subprocess = types.SimpleNamespace(Popen=_FakePopen, PIPE=-1, call=lambda *a, **k: 0)
    mod.NPZ2Cooler = lambda **kw: calls.append(kw)
    try:
        prefix, datasets = mod.HaplotypeMatrixBuilding(out_dir, bed_dir, genome_size, whole_res, local_res,
                                                       imputation[0], imputation[1], imputation[2], chroms)
    finally:
        mod.subprocess, mod.NPZ2Cooler = saved
    return prefix, datasets, calls


def flatten(prefix, obj, out):
    """nested dict of arrays -> flat npz keys joined by '|'."""
    if isinstance(obj, dict):
        for k, v in obj.items():
            flatten("%s|%s" % (prefix, k), v, out)
    elif isinstance(obj, tuple) and len(obj) == 2 and all(isinstance(x, (int, np.integer)) for x in obj):
        out[prefix] = np.asarray(obj, dtype=np.int64)
    else:
        out[prefix] = np.asarray(obj)


def make_traditional():
    mod = ref_shim.load()
    names, c1, p1, c2, p2 = small_pairs(20000, seed=11)
    with tempfile.TemporaryDirectory() as td:
        gs = synth.write_genome_size(os.path.join(td, "genomeSize"), SMALL_GENOME)
        text = "".join(synth.valid23_lines(names, c1, p1, c2, p2))
        whole, local = mod.TraditionalMatrixBuilding(io.StringIO(text), gs, [500000], [40000], CHROMS)
    out = dict(names=np.array(names), c1=c1, p1=p1, c2=c2, p2=p2,
               whole_res=np.array([500000]), local_res=np.array([40000]))
    flatten("whole", whole, out)
    flatten("local", local, out)
    np.savez_compressed(os.path.join(GOLDEN, "traditional_small.npz"), **out)
    return out


def allelic_inputs(seed=21, n=24000):
    """Five allelic beds: class split and Both/R1/R2 marks as in SURVEY.md section 8d (C3)."""
    rng = np.random.default_rng(seed)
    names, c1, p1, c2, p2 = small_pairs(n, seed=seed, trans_frac=0.1)
    cls = rng.choice(5, size=n, p=[0.50, 0.22, 0.22, 0.03, 0.03])   # Bi, M_M, P_P, M_P, P_M
    mark = rng.choice(3, size=n, p=[0.3, 0.35, 0.35]).astype(np.uint8)
    return names, c1, p1, c2, p2, cls, mark


CLASS_FILES = ["Bi_Allelic", "M_M", "P_P", "M_P", "P_M"]


def write_allelic_beds(td, names, c1, p1, c2, p2, cls, mark):
    bed_dir = os.path.join(td, "beds")
    os.makedirs(bed_dir, exist_ok=True)
    for k, tag in enumerate(CLASS_FILES):
        sel = cls == k
        # M_M / P_P carry the Both/R1/R2 mark (filtering.py:913-958); the others are 4-column here
        mk = mark[sel] if tag in ("M_M", "P_P") else None
        with open(os.path.join(bed_dir, "S_Valid_%s.bed" % tag), "w") as fh:
            fh.writelines(synth.allelic_lines(names, c1[sel], p1[sel], c2[sel], p2[sel], mk))
    return bed_dir


def imputation_inputs(seed=33, n=60000):
    """Allelic beds with many inter-chromosomal contacts, so that the neighbourhood vote of the
    imputation (matrixBuilding.py:1302-1378, :1416-1492) fires in every branch."""
    rng = np.random.default_rng(seed)
    names, c1, p1, c2, p2 = small_pairs(n, seed=seed, trans_frac=0.45)
    cls = rng.choice(5, size=n, p=[0.20, 0.32, 0.32, 0.08, 0.08])
    mark = rng.choice(3, size=n, p=[0.4, 0.3, 0.3]).astype(np.uint8)
    return names, c1, p1, c2, p2, cls, mark


IMPUTATION_CASES = [      # (wholeRes in call order, Imputation_region, Imputation_min, Imputation_ratio)
    ([500000, 250000], 1500000, 2, 0.6),
    ([500000], 1000000, 1, 0.9),
    ([250000], 2500000, 3, 0.55),
]


def make_imputation():
    """imputation_small.npz: genome-wide un-imputed / imputed haplotype matrices of the reference's
    HaplotypeMatrixBuilding for several (wholeRes, region, min, ratio) settings on one set of beds."""
    names, c1, p1, c2, p2, cls, mark = imputation_inputs()
    out = dict(names=np.array(names), c1=c1, p1=p1, c2=c2, p2=p2, cls=cls, mark=mark, ncases=np.array(len(IMPUTATION_CASES)))
    with tempfile.TemporaryDirectory() as td:
        gs = synth.write_genome_size(os.path.join(td, "genomeSize"), SMALL_GENOME)
        bed_dir = write_allelic_beds(td, names, c1, p1, c2, p2, cls, mark)
        for k, (whole_res, region, imin, ratio) in enumerate(IMPUTATION_CASES):
            out_dir = os.path.join(td, "out%d" % k)
            os.makedirs(out_dir)
            _, ds, _ = run_reference_haplotype(bed_dir, gs, whole_res, [1000000], CHROMS, out_dir,
                                               imputation=(region, imin, ratio))
            out["case%d|params" % k] = np.array([region, imin, ratio], dtype=np.float64)
            out["case%d|whole_res" % k] = np.array(whole_res)
            for res in whole_res:
                un, imp = ds["UnImputated_Whole"][res]["Matrix"], ds["Imputated_Whole"][res]["Matrix"]
                out["case%d|un|%d" % (k, res)] = un.astype(np.int32)
                out["case%d|imp|%d" % (k, res)] = imp.astype(np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "imputation_small.npz"), **out)
    return out


def make_allelic():
    names, c1, p1, c2, p2, cls, mark = allelic_inputs()
    whole_res, local_res = [500000], [80000]
    with tempfile.TemporaryDirectory() as td:
        gs = synth.write_genome_size(os.path.join(td, "genomeSize"), SMALL_GENOME)
        bed_dir = write_allelic_beds(td, names, c1, p1, c2, p2, cls, mark)
        out_dir = os.path.join(td, "out")
        os.makedirs(out_dir)
        prefix, ds, calls = run_reference_haplotype(bed_dir, gs, whole_res, local_res, CHROMS, out_dir)
        gap = np.load(os.path.join(out_dir, prefix + "Imputated_Gap.npz"), allow_pickle=True)
        gap_local = {k: gap[k].item() for k in gap.files}
    # the float ('Imputated') cool datasets are the last two NPZ2Cooler calls (matrixBuilding.py:1621-1633)
    balanced_whole, balanced_local = calls[-2]["datasets"], calls[-1]["datasets"]
    mod = ref_shim.load()
    nor = {}
    for res in local_res:
        nor_lib, _ = mod.IntraChromMatrixCorrection(ds["Tradition_Local"][res], ds["Imputated_Local"][res])
        nor[res] = nor_lib
    out = dict(names=np.array(names), c1=c1, p1=p1, c2=c2, p2=p2, cls=cls, mark=mark,
               whole_res=np.array(whole_res), local_res=np.array(local_res), prefix=np.array(prefix))
    for key in ("Tradition_Whole", "Tradition_Local", "UnImputated_Whole", "UnImputated_Local",
                "Imputated_Whole", "Imputated_Local"):
        flatten(key, ds[key], out)
    flatten("Gap", gap_local, out)
    flatten("Nor_Local", nor, out)
    flatten("Balanced_Local", balanced_local, out)
    for res in whole_res:
        gw = mod.GenomeWideMatrixCorrection(ds["Tradition_Whole"][res]["Bins"], ds["Imputated_Whole"][res]["Bins"],
                                            ds["Tradition_Whole"][res]["Matrix"], ds["Imputated_Whole"][res]["Matrix"])
        out["GenomeWide|%d" % res] = gw
    del balanced_whole
    np.savez_compressed(os.path.join(GOLDEN, "allelic_small.npz"), **out)
    return out


def make_twostep_cases():
    mod = ref_shim.load()
    rng = np.random.default_rng(31)
    out = {}
    # (1) gap-free: every row well covered -> threshold capped at 0.2 -> no gap rows -> SUM rule
    n = 70
    tm = rng.poisson(6.0, size=(n, n)); tm = tm + tm.T
    mm = rng.poisson(1.5, size=(n, n))
    pm = rng.poisson(1.2, size=(n, n))
    # (2) gappy and asymmetric (imputed matrices are not symmetric), sparse rows present
    n2 = 90
    dens = np.clip(rng.gamma(2.0, 0.25, size=n2), 0.0, 1.0)
    dens[10:14] = 0.0
    dens[40] = 0.02
    lam = 3.0 * dens[:, None] * dens[None, :]
    mm2 = rng.poisson(lam); pm2 = rng.poisson(0.8 * lam)
    tm2 = rng.poisson(8.0 * lam); tm2 = tm2 + tm2.T + mm2 + pm2
    for tag, (a, b, c) in {"nogap": (tm, mm, pm), "gappy": (tm2, mm2, pm2)}.items():
        a, b, c = (np.asarray(x, dtype=np.int64) for x in (a, b, c))
        nm, npm, gm, gp = mod.TwoStepCorrection(a, b, c)
        out.update({tag + "|TM": a, tag + "|MM": b, tag + "|PM": c, tag + "|Nor_MM": nm,
                    tag + "|Nor_PM": npm, tag + "|Gap_M": np.asarray(gm), tag + "|Gap_P": np.asarray(gp)})
    assert out["nogap|Gap_M"].size == 0 and out["nogap|Gap_P"].size == 0
    assert out["gappy|Gap_M"].size > 0
    np.savez_compressed(os.path.join(GOLDEN, "twostep_cases.npz"), **out)
    return out


def make_building_blocks():
    """The reference's own helper functions on the inputs of twostep_cases.npz (plus a rectangular
    Correct_VC case): Coverage_M :904, Gap_defined :915, Gap_definedLowRes :742, Non_Gap_Defined :932,
    Trans2symmetry :945, Trans2symmetryLowRes :770, Correct_VC :780."""
    mod = ref_shim.load()
    g = np.load(os.path.join(GOLDEN, "twostep_cases.npz"), allow_pickle=True)
    rng = np.random.default_rng(77)
    out = {}
    for tag in ("nogap", "gappy"):
        mm = g[tag + "|MM"]
        gap = np.asarray(mod.Gap_defined(mm))
        S = mm / rng.uniform(0.3, 1.0, size=mm.shape[0])[:, None]          # what TwoStepCorrection feeds (:1007)
        sym = mod.Trans2symmetry(S, gap)
        out.update({tag + "|M": mm, tag + "|Coverage": mod.Coverage_M(mm), tag + "|Gap": gap,
                    tag + "|GapLowRes": np.asarray(mod.Gap_definedLowRes(mm)),
                    tag + "|NonGap": np.asarray(mod.Non_Gap_Defined(mm.shape[0], gap)),
                    tag + "|S": S, tag + "|Sym": sym, tag + "|SymLowRes": mod.Trans2symmetryLowRes(S),
                    tag + "|VC": mod.Correct_VC(sym, 2.0 / 3)})
    # the max rule needs two gap rows with asymmetric entries between them
    S = rng.gamma(1.0, 2.0, size=(40, 40)); S[rng.random((40, 40)) < 0.5] = 0.0
    gap = np.array([3, 4, 17, 39])
    out.update({"forced|S": S, "forced|Gap": gap, "forced|Sym": mod.Trans2symmetry(S, gap)})
    X = rng.poisson(2.0, size=(37, 53)).astype(float); X[5, :] = 0; X[:, 11] = 0
    out.update({"rect|X": X, "rect|VC": mod.Correct_VC(X, 0.5)})
    np.savez_compressed(os.path.join(GOLDEN, "building_blocks.npz"), **out)
    return out


class _GapArray(np.ndarray):
    """ndarray whose ``== None`` is a plain False, as it was for the NumPy the 2019 reference ran on (the reference tests
    ``if G_array == None`` at StructureFind.py:215; NumPy 2 would compare elementwise and the ``if`` would raise)."""

    def __eq__(self, other):
        if other is None:
            return False
        return np.ndarray.__eq__(self, other)

    __hash__ = None


def make_consumers():
    """Outputs of the reference's StructureFind methods that first consume the matrix stage's results:
    Distance_Decay (:201-272) with its own gap rule and with a given gap list, the O/E loop of Get_PCA (:318-326),
    Get_DI (:804-840) in both flavours, bias_handle (:1948-1952)."""
    sf = ref_shim.load_structure()
    obj = object.__new__(sf.StructureFind)
    rng = np.random.default_rng(99)
    n = 120
    bias = np.exp(rng.normal(0, 0.3, n))
    d = np.abs(np.subtract.outer(np.arange(n), np.arange(n))) + 1.0
    M = rng.poisson(40.0 * bias[:, None] * bias[None, :] / d).astype(float)
    M = np.triu(M) + np.triu(M, 1).T
    M[30:34, :] = 0; M[:, 30:34] = 0; M[77, :] = 0; M[:, 77] = 0
    w = 1.0 / np.sqrt(np.maximum(M.sum(1), 1.0)); w[[30, 31, 32, 33, 77]] = np.nan
    cM = np.nan_to_num(M * w[:, None] * w[None, :])            # what cooler hands CallPeaks (:2005-2007)
    out = {"M": M, "weight": w, "cM": cM}
    db, G, NG = obj.Distance_Decay(cM, None)
    out.update(dd_auto=db.copy(), dd_auto_G=np.asarray(G), dd_auto_NG=np.asarray(NG))
    given = np.array([3, 30, 31, 32, 33, 77, 119]).view(_GapArray)
    db2, G2, NG2 = obj.Distance_Decay(cM, given)
    out.update(dd_given=db2.copy(), dd_given_G=np.asarray(G2), dd_given_NG=np.asarray(NG2))
    # the O/E loop of Get_PCA, verbatim semantics (:318-326); the PCA that follows is outside the scope
    decline = db.copy()
    decline[decline == 0] = decline[np.nonzero(decline)].min()
    OE = np.zeros(cM.shape)
    for i in range(n):
        for j in range(n):
            if cM[i][j] != 0:
                OE[i][j] = cM[i][j] / decline[abs(i - j)]
    out["OE"] = OE
    window = rng.integers(1, 9, n)
    gap = np.array([30, 31, 32, 33, 77])
    for t in ("ttest", "chitest"):
        obj.test_type = t
        with np.errstate(invalid="ignore", divide="ignore"):
            out["DI_" + t] = obj.Get_DI(cM, gap, window)
    out.update(window=window, gap=gap)
    b = w.copy().reshape(-1, 1)
    out["bias_handled"] = obj.bias_handle(b.copy())
    np.savez_compressed(os.path.join(GOLDEN, "consumers.npz"), **out)
    return out


def make_ice_restated(trad):
    """Regression anchor for the cooler restatement (NOT a reference output)."""
    out = {}
    # cis-only on the four local matrices of the traditional fixture, denser input for a real run
    names, c1, p1, c2, p2 = small_pairs(400000, seed=41, trans_frac=0.0)
    order = ["1", "2", "10", "X"]
    remap = {names.index(c): i for i, c in enumerate(order)}
    keep = np.isin(c1, list(remap.keys())) & (c1 == c2)
    cid = np.vectorize(remap.get)(c1[keep]).astype(np.int32)
    res = 40000
    sizes = [SMALL_GENOME[c] // res + 1 for c in order]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    b1 = p1[keep] // res + offs[cid]
    b2 = p2[keep] // res + offs[cid]
    lo, hi = np.minimum(b1, b2), np.maximum(b1, b2)
    key, cnt = np.unique(lo.astype(np.int64) * offs[-1] + hi, return_counts=True)
    bin1, bin2 = key // offs[-1], key % offs[-1]
    w, st = cooler_ice.balance(bin1, bin2, cnt, int(offs[-1]), offs, cis_only=True, ignore_diags=1)
    out.update(bin1=bin1, bin2=bin2, count=cnt.astype(np.int32), chrom_offsets=offs, weight_cis=w,
               scale_cis=st["scale"], iters_cis=np.array(st["iters"]), var_cis=np.array(st["var"]))
    # genome-wide: add trans contacts so the matrix is one connected component
    names, c1, p1, c2, p2 = small_pairs(400000, seed=42, trans_frac=0.25)
    keep = np.isin(c1, list(remap.keys())) & np.isin(c2, list(remap.keys()))
    ca = np.vectorize(remap.get)(c1[keep]); cb = np.vectorize(remap.get)(c2[keep])
    b1 = p1[keep] // res + offs[ca]
    b2 = p2[keep] // res + offs[cb]
    lo, hi = np.minimum(b1, b2), np.maximum(b1, b2)
    key, cnt = np.unique(lo.astype(np.int64) * offs[-1] + hi, return_counts=True)
    bin1, bin2 = key // offs[-1], key % offs[-1]
    w2, st2 = cooler_ice.balance(bin1, bin2, cnt, int(offs[-1]), offs, cis_only=False, ignore_diags=1)
    out.update(gw_bin1=bin1, gw_bin2=bin2, gw_count=cnt.astype(np.int32), weight_gw=w2, scale_gw=np.array(st2["scale"]), iters_gw=np.array(st2["iters"]),
               var_gw=np.array(st2["var"]))
    np.savez_compressed(os.path.join(GOLDEN, "ice_restated.npz"), **out)
    return out


def main():
    if not ref_shim.available():
        raise SystemExit("reference not found at %s" % ref_shim.REFERENCE_ROOT)
    os.makedirs(GOLDEN, exist_ok=True)
    trad = make_traditional()
    make_allelic()
    make_imputation()
    make_twostep_cases()
    make_building_blocks()
    make_consumers()
    make_ice_restated(trad)
    for f in sorted(os.listdir(GOLDEN)):
        print("%-28s %8d bytes" % (f, os.path.getsize(os.path.join(GOLDEN, f))))


if __name__ == "__main__":
    main()
