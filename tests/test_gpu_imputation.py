"""GPU parity of the inter-chromosomal imputation (`hc_impute_inter` behind
`matrixBuilding.impute_inter_chromosomal` and the `HaplotypeMatrixBuilding` driver) against matrices the
reference's own HaplotypeMatrixBuilding produced (tests/golden/imputation_small.npz) and against the
oracle on larger seeded inputs.  Bit-exact (integer counts)."""
import os

import numpy as np
import pytest

from conftest import CHROMS, SMALL_GENOME, load_golden
from hichap_master_b200 import synth
from oracle import hichap_oracle as ho
from test_gpu_drivers import write_allelic_beds

pytestmark = pytest.mark.gpu


def _device_case(mb, cuda_device, genome, order, whole_res, files, un_np):
    """un / imp one-matrix batches (imp = un + cis one-sided) and start tables on the device."""
    import torch
    from hichap_master_b200 import _abi, kernels
    from hichap_master_b200.device import DenseBatch, PairColumns
    cols = {tag: PairColumns(*files[tag], device=cuda_device) for tag in files}
    un, imp, starts = {}, {}, {}
    for res in whole_res:
        hb, htot = ho.chro_bins_haplotypes(genome, res)
        sm = torch.tensor([hb["M" + c][0] for c in order], dtype=torch.int64, device=cuda_device)
        sp = torch.tensor([hb["P" + c][0] for c in order], dtype=torch.int64, device=cuda_device)
        U = DenseBatch([htot], cuda_device)
        n, ld = U.sizes[0], U.lds[0]
        U.buf[:n * ld].view(n, ld)[:, :n] = torch.from_numpy(un_np[res].astype(np.int32)).to(cuda_device)
        I = DenseBatch([htot], cuda_device)
        I.buf.copy_(U.buf)
        kernels.bin_pairs_whole(cols["M_M"], res, sm, sm, I, _abi.HC_BIN_ONESIDED)
        kernels.bin_pairs_whole(cols["P_P"], res, sp, sp, I, _abi.HC_BIN_ONESIDED)
        un[res], imp[res], starts[res] = U, I, (sm, sp)
    return cols, un, imp, starts


def test_imputation_matches_reference_golden(cuda_device, small_genome_file):
    from hichap_master_b200 import matrixBuilding as mb
    from test_oracle_golden import _imputation_case
    g = load_golden("imputation_small.npz")
    genome = ho.load_genome(small_genome_file, CHROMS)
    for k in range(int(g["ncases"])):
        order, files, _, whole_res, region, imin, ratio = _imputation_case(g, k, genome)
        # the product drops filtered lines at parse time; relative file order is what matters
        files = {t: tuple(a[(c[0] >= 0) & (c[2] >= 0)] for a in c) for t, c in files.items()}
        un_np = {res: g["case%d|un|%d" % (k, res)] for res in whole_res}
        cols, un, imp, starts = _device_case(mb, cuda_device, genome, order, whole_res, files, un_np)
        mb.impute_inter_chromosomal(un, imp, cols["M_M"], cols["P_P"], starts, whole_res, region, imin, ratio)
        for res in whole_res:
            assert np.array_equal(un[res].to_numpy(0), un_np[res])                      # read only
            assert np.array_equal(imp[res].to_numpy(0), g["case%d|imp|%d" % (k, res)]), (k, res)


def test_haplotype_building_driver_with_imputation_golden(cuda_device, tmp_path, small_genome_file):
    """The driver with the reference's signature, -region/-min/-ratio as in case 0 of the golden."""
    from hichap_master_b200 import matrixBuilding as mb
    g = load_golden("imputation_small.npz")
    bed_dir = write_allelic_beds(str(tmp_path / "beds"), g)
    out_dir = str(tmp_path / "out"); os.makedirs(out_dir)
    region, imin, ratio = g["case0|params"]
    whole_res = [int(r) for r in g["case0|whole_res"]]
    _, ds = mb.HaplotypeMatrixBuilding(out_dir, bed_dir, small_genome_file, whole_res, [1000000], int(region),
                                       int(imin), float(ratio), CHROMS)
    for res in whole_res:
        assert np.array_equal(ds["UnImputated_Whole"][res]["Matrix"], g["case0|un|%d" % res])
        assert np.array_equal(ds["Imputated_Whole"][res]["Matrix"], g["case0|imp|%d" % res])


@pytest.mark.parametrize("seed,whole_res,params", [(301, [200000], (1000000, 2, 0.6)),
                                                    (302, [400000, 100000], (800000, 1, 0.75)),
                                                    (303, [100000], (3000000, 5, 0.5))])
def test_imputation_vs_oracle_seeded(cuda_device, seed, whole_res, params):
    """Larger genome, many one-sided inter-chromosomal lines, all four branches, big neighbourhood discs."""
    from hichap_master_b200 import matrixBuilding as mb
    genome = {"1": 24_000_000, "2": 19_500_000, "3": 17_000_000, "X": 15_000_001}
    order = ho.sort_chromosomes(genome)
    rng = np.random.default_rng(seed)
    n = 400_000
    c1, p1, c2, p2 = synth.genome_pairs(genome, order, n, seed, trans_frac=0.5)
    cls = rng.choice(5, size=n, p=[0.10, 0.35, 0.35, 0.10, 0.10])
    mark = rng.choice(4, size=n, p=[0.5, 0.25, 0.2, 0.05]).astype(np.uint8)       # 3 = other text: treated as R2
    files = {tag: tuple(a[cls == k] for a in (c1, p1, c2, p2, mark)) for tag, k in (("M_M", 1), ("P_P", 2))}
    un_np, starts_np = {}, {}
    for res in whole_res:
        hb, htot = ho.chro_bins_haplotypes(genome, res)
        sm = np.array([hb["M" + c][0] for c in order]); sp = np.array([hb["P" + c][0] for c in order])
        H = np.zeros((htot, htot), np.int64)
        for k, (s1, s2, both) in {1: (sm, sm, True), 2: (sp, sp, True), 3: (sm, sp, False), 4: (sp, sm, False)}.items():
            sel = (cls == k) & ((mark == 0) | (not both))
            ho.bin_whole_dense(c1[sel], p1[sel], c2[sel], p2[sel], s1, s2, htot, res, out=H)
        un_np[res], starts_np[res] = H, (sm, sp)
    exp = {res: un_np[res].copy() for res in whole_res}
    for res in whole_res:
        for tag, own in (("M_M", 0), ("P_P", 1)):
            ho.bin_whole_onesided(*files[tag], starts_np[res][own], res, exp[res])
    cis_only = {res: exp[res].copy() for res in whole_res}
    ho.impute_inter_chromosomal(un_np, exp, whole_res, starts_np, files, *params)
    cols, un, imp, starts = _device_case(mb, cuda_device, genome, order, whole_res, files, un_np)
    mb.impute_inter_chromosomal(un, imp, cols["M_M"], cols["P_P"], starts, whole_res, *params)
    for res in whole_res:
        assert np.array_equal(imp[res].to_numpy(0), exp[res]), res
        assert (exp[res] - cis_only[res]).sum() > 100


def test_imputation_error_behaviour(cuda_device):
    """NameError when the P_P R1 branch needs the stale M_M window and no M_M line produced one;
    IndexError when the stale window comes from a coarser resolution (both as the reference)."""
    from hichap_master_b200 import matrixBuilding as mb
    genome = {"1": 24_000_000, "2": 19_500_000}
    order = ["1", "2"]
    z = lambda *v: np.array(v, np.int32)
    # one P_P R1 inter-chromosomal line in the middle of both chromosomes, no M_M lines at all
    pp = (z(0), z(12_000_000), z(1), z(9_000_000), np.array([1], np.uint8))
    mm0 = tuple(a[:0] for a in pp)
    res = 500000
    hb, htot = ho.chro_bins_haplotypes(genome, res)
    un_np = {res: np.zeros((htot, htot), np.int64)}
    cols, un, imp, starts = _device_case(mb, cuda_device, genome, order, [res], {"M_M": mm0, "P_P": pp}, un_np)
    with pytest.raises(NameError):
        mb.impute_inter_chromosomal(un, imp, cols["M_M"], cols["P_P"], starts, [res], 1500000, 2, 0.9)
    # ... but a P_P file without such a line is fine, as in the reference
    pp2 = (z(0), z(12_000_000), z(1), z(9_000_000), np.array([0], np.uint8))
    cols, un, imp, starts = _device_case(mb, cuda_device, genome, order, [res], {"M_M": mm0, "P_P": pp2}, un_np)
    mb.impute_inter_chromosomal(un, imp, cols["M_M"], cols["P_P"], starts, [res], 1500000, 2, 0.9)
    assert int(imp[res].buf.sum().item()) == 0
    # stale window from the coarser, later resolution: IndexError at the finer one
    mm = (z(0), z(12_000_000), z(1), z(9_000_000), np.array([2], np.uint8))
    wr = [250000, 500000]
    un_np = {r: np.zeros((ho.chro_bins_haplotypes(genome, r)[1],) * 2, np.int64) for r in wr}
    cols, un, imp, starts = _device_case(mb, cuda_device, genome, order, wr, {"M_M": mm, "P_P": pp}, un_np)
    with pytest.raises(IndexError):
        mb.impute_inter_chromosomal(un, imp, cols["M_M"], cols["P_P"], starts, wr, 1500000, 2, 0.9)
    exp = {r: un_np[r].copy() for r in wr}
    with pytest.raises(IndexError):
        ho.impute_inter_chromosomal(un_np, exp, wr, {r: tuple(t.cpu().numpy() for t in starts[r]) for r in wr},
                                    {"M_M": mm, "P_P": pp}, 1500000, 2, 0.9)
