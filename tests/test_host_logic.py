"""CPU tests of the host-side mirror: genome/bin tables and the text -> columnar parser,
against the oracle restatement (itself pinned to the reference)."""
import io

import numpy as np
import pytest

from conftest import CHROMS, SMALL_GENOME
from hichap_master_b200 import pairs, synth
from oracle import hichap_oracle as ho


def test_genome_tables_match_oracle(small_genome_file):
    from hichap_master_b200 import matrixBuilding as mb
    for chroms in (["#", "X"], [], ["#"], ["X", "Y"]):
        genome = mb.Load_Genome(small_genome_file, chroms)
        assert genome == ho.load_genome(small_genome_file, chroms)
        assert mb.Sort_Chromosomes(genome) == ho.sort_chromosomes(genome)
        for res in (40000, 500000, 4_800_000):
            assert mb.Get_Chro_Bins(small_genome_file, res, chroms) == ho.chro_bins(genome, res)
            assert mb.Get_Chro_Bins_Haplotypes(small_genome_file, res, chroms) == ho.chro_bins_haplotypes(genome, res)
        hg = mb.Load_HaplotypeGenome(small_genome_file, chroms)
        assert set(hg) == {t + c for c in genome for t in "MP"}


def test_lstrip_chr_is_a_character_set():
    from hichap_master_b200 import matrixBuilding as mb
    # 'chrchr1' -> '1', 'rch2' -> '2' (str.lstrip strips the set {c,h,r}; matrixBuilding.py:359)
    assert mb.Sort_Chromosomes(["chrchr1", "rch2", "chrX"]) == ["1", "2", "X"]


@pytest.mark.parametrize("as_bytes", [False, True])
def test_valid23_parser_matches_oracle(small_genome_file, as_bytes):
    genome = ho.load_genome(small_genome_file, CHROMS)
    order = ho.sort_chromosomes(genome)
    names = [c for c in SMALL_GENOME if c != "M"]
    big = {c: SMALL_GENOME[c] for c in names}
    c1, p1, c2, p2 = synth.genome_pairs(big, names, 2000, 5, trans_frac=0.3)
    text = "".join(synth.valid23_lines(names, c1, p1, c2, p2))
    exp = ho.parse_pairs(text.splitlines(True), genome, CHROMS, "valid23")
    stream = io.BytesIO(text.encode()) if as_bytes else io.StringIO(text)
    got = pairs.read_pairs(stream, order, CHROMS, "valid23")
    keep = (got[0] >= 0) & (got[2] >= 0)          # the product keeps dropped rows as chrom -1
    for a, b in zip(got[:4], exp[:4]):
        assert np.array_equal(a[keep], b)
    assert got[4] is None


def test_allelic_parser_ragged_columns(small_genome_file):
    genome = ho.load_genome(small_genome_file, CHROMS)
    order = ho.sort_chromosomes(genome)
    text = ("chr1\t100\tchr1\t90000\tBoth\n"
            "chr2\t5\tchrX\t7\n"                      # 4-column Bi_Allelic style line
            "chr10\t1\tchr10\t2\tR1\n"
            "chrY\t1\tchr1\t2\tR2\n"                  # dropped by the filter
            "chrX\t11\tchrX\t12\tR2\n")
    exp = ho.parse_pairs(text.splitlines(True), genome, CHROMS, "allelic")
    got = pairs.read_pairs(io.StringIO(text), order, CHROMS, "allelic")
    keep = (got[0] >= 0) & (got[2] >= 0)
    for a, b in zip(got, exp):
        assert np.array_equal(a[keep], b)
    assert list(got[4]) == [0, 3, 1, 2, 2]
    # iterable of lines (what `for line in bed_IO` accepts) and empty input
    got2 = pairs.read_pairs(text.splitlines(True), order, CHROMS, "allelic")
    assert all(np.array_equal(a, b) for a, b in zip(got, got2))
    empty = pairs.read_pairs(io.StringIO(""), order, CHROMS, "allelic")
    assert empty[0].size == 0 and empty[4].size == 0


def test_unknown_chromosome_raises_keyerror(small_genome_file):
    genome = ho.load_genome(small_genome_file, CHROMS)
    order = ho.sort_chromosomes(genome)
    with pytest.raises(KeyError):
        pairs.read_pairs(io.StringIO("chr7\t1\tchr1\t2\tBoth\n"), order, CHROMS, "allelic")


def test_cli_flags_match_reference_matrix_subcommand():
    """scripts/hichap:392-427: same flags and defaults for `matrix`."""
    from hichap_master_b200.__main__ import getargs
    a = getargs(["matrix", "-b", "r1", "r2", "-o", "out", "-N", "-gs", "gs", "-wR", "1000000", "-C"])
    assert a.bedPath == ["r1", "r2"] and a.NonAllelic and a.wholeRes == [1000000] and a.chroms == []
    d = getargs(["matrix", "-b", "r1", "-o", "out", "-gs", "gs"])
    assert d.localRes == [500000, 40000] and d.wholeRes is None and d.chroms == ["#", "X"]
    assert (d.ImputationRatio, d.ImputationMin, d.ImputationRegion, d.NonAllelic) == (0.9, 2, 10000000, False)


# ---- native multithreaded ingest (hc_ingest_parse): CPU-only, no GPU needed -------------------
@pytest.mark.parametrize("layout", ["valid23", "allelic"])
@pytest.mark.parametrize("chroms", [["#", "X"], [], ["2", "X"]])
def test_native_parser_matches_oracle(tmp_path, small_genome_file, layout, chroms):
    genome = ho.load_genome(small_genome_file, chroms)
    order = ho.sort_chromosomes(genome)
    names = [c for c in SMALL_GENOME if c != "M"]
    big = {c: SMALL_GENOME[c] for c in names}
    files, lines_all = [], []
    for k, seed in enumerate((5, 6, 7)):
        c1, p1, c2, p2 = synth.genome_pairs(big, names, 4000 + k, seed, trans_frac=0.3)
        if layout == "valid23":
            lines = list(synth.valid23_lines(names, c1, p1, c2, p2))
        else:
            mark = np.random.default_rng(seed).integers(0, 3, size=c1.size)
            lines = list(synth.allelic_lines(names, c1, p1, c2, p2, mark if k < 2 else None))   # last file: 4 columns
        if k == 1:
            lines[10] = lines[10].replace("\t", "  \t ", 3)      # runs of mixed whitespace
            lines.insert(20, "\n")                               # blank line
            lines[-1] = lines[-1].rstrip("\n")                   # no trailing newline at EOF
        f = tmp_path / ("part%d.bed" % k)
        f.write_text("".join(lines))
        files.append(str(f)); lines_all.extend(l if l.endswith("\n") else l + "\n" for l in lines)
    exp = ho.parse_pairs(lines_all, genome, chroms, layout)
    for nthreads in (1, 3):
        got = pairs.read_pair_files(files, order, chroms, layout, nthreads=nthreads)
        for a, b in zip(got[:4], exp[:4]):
            assert a.dtype == np.int32 and np.array_equal(a, b)
        if layout == "allelic":
            assert np.array_equal(got[4], exp[4])
        else:
            assert got[4] is None


def test_native_parser_errors_and_edge_cases(tmp_path, small_genome_file):
    genome = ho.load_genome(small_genome_file, CHROMS)
    order = ho.sort_chromosomes(genome)
    empty = tmp_path / "empty.bed"; empty.write_text("")
    got = pairs.read_pair_files([str(empty)], order, CHROMS, "allelic")
    assert got[0].size == 0 and got[4].size == 0
    assert pairs.read_pair_files([], order, CHROMS, "valid23")[0].size == 0
    bad = tmp_path / "bad.bed"; bad.write_text("chr7\t1\tchr1\t2\tBoth\n")      # passes '#' filter, not in genome
    with pytest.raises(KeyError):
        pairs.read_pair_files([str(bad)], order, CHROMS, "allelic")
    ok = tmp_path / "ok.bed"; ok.write_text("chr7\t1\tchrY\t2\tBoth\nchrchr1\t5\trch2\t9\tR1\n")  # filtered mate wins; set-strip
    got = pairs.read_pair_files([str(ok)], order, CHROMS, "allelic")
    assert list(got[0]) == [0] and list(got[2]) == [1] and list(got[4]) == [1]
    junk = tmp_path / "junk.bed"; junk.write_text("chr1\tx\tchr1\t2\tBoth\n")
    with pytest.raises(ValueError):
        pairs.read_pair_files([str(junk)], order, CHROMS, "allelic")
    with pytest.raises(IOError):
        pairs.read_pair_files([str(tmp_path / "missing.bed")], order, CHROMS, "allelic")


def test_native_parser_whitespace_and_last_token(tmp_path):
    """str.split() semantics: runs of blanks / tabs, leading blanks, trailing blanks, blank lines; the allelic mark
    is the LAST token of the line (matrixBuilding.py:1270 `line[-1]`), which for a 4-column line is the position."""
    from hichap_master_b200 import pairs
    order = ["1", "2", "X"]
    text = ("chr1\t100\tchr2\t200\tBoth\nchr1 300   chrX\t400\nchr2\t5\tchr2\t6\tR1  \n\n"
            "  chrX\t7\tchr1\t8\tfoo bar\tR2\nchrM\t1\tchr1\t2\tBoth\n")
    path = tmp_path / "a.bed"
    path.write_text(text)
    c1, p1, c2, p2, mark = pairs.read_pair_files([str(path)], order, ["#", "X"], "allelic")
    assert c1.tolist() == [0, 0, 1, 2] and p1.tolist() == [100, 300, 5, 7]
    assert c2.tolist() == [1, 2, 1, 0] and p2.tolist() == [200, 400, 6, 8]
    assert mark.tolist() == [0, 3, 1, 2]
    # same lines through the stream parser (pandas) used for file-like inputs
    import io
    d1, q1, d2, q2, mk = pairs.read_pairs(io.StringIO("chr1\t100\tchr2\t200\tBoth\nchr2\t5\tchr2\t6\tR1\n"), order, ["#", "X"], "allelic")
    assert d1.tolist() == [0, 1] and mk.tolist() == [0, 1]


def test_cool_export_tables(tmp_path, small_genome_file):
    """npz store -> (bins, pixels) as cooler.create_cooler takes them: cooler's bin table (ceil(len / res) bins,
    last one clipped), pixels sorted by (bin1_id, bin2_id) with global ids, weights re-laid on those bins."""
    from conftest import CHROMS
    from hichap_master_b200 import cool_export
    from hichap_master_b200.construction import MatrixStore
    genome = ho.load_genome(small_genome_file, CHROMS)
    order = ho.sort_chromosomes(genome)
    rng = np.random.default_rng(3)
    res = 500000
    # local (intra-chromosomal) store with weights over HiCHap's len // res + 1 bins per chromosome
    store = MatrixStore(str(tmp_path / "local.npz"))
    dense = {}
    for c in order:
        n = genome[c] // res + 1
        M = rng.poisson(2.0, size=(n, n)); M = np.triu(M) + np.triu(M, 1).T
        if genome[c] % res == 0:
            M[-1, :] = 0; M[:, -1] = 0            # HiCHap's extra bin past the chromosome end is never hit
        dense[c] = M
    store.add(res, {c: ho.dense_to_triu_records(dense[c]) for c in order})
    w = rng.random(sum(genome[c] // res + 1 for c in order))
    store.set_weight(res, w, dict(tol=1e-5, converged=True, scale=np.array([1.0, 2.0])))
    path = store.save()
    bins, pixels, attrs = cool_export.store_tables(np.load(path, allow_pickle=True), res, genome)
    nb = {c: -(-genome[c] // res) for c in order}
    assert len(bins) == sum(nb.values()) and list(bins["chrom"].unique()) == order
    assert (bins["end"] - bins["start"]).max() == res and int(bins["end"].iloc[-1]) == genome[order[-1]]
    assert attrs["converged"] is True and attrs["tol"] == 1e-5
    assert pixels["count"].dtype == np.int32
    key = pixels["bin1_id"].to_numpy() * len(bins) + pixels["bin2_id"].to_numpy()
    assert np.all(np.diff(key) > 0) and np.all(pixels["bin1_id"] <= pixels["bin2_id"])
    off, hoff = 0, 0
    for c in order:
        sel = (pixels["bin1_id"] >= off) & (pixels["bin1_id"] < off + nb[c])
        R = np.zeros((nb[c], nb[c]), np.int64)
        R[pixels["bin1_id"][sel] - off, pixels["bin2_id"][sel] - off] = pixels["count"][sel]
        assert np.array_equal(R, np.triu(dense[c])[:nb[c], :nb[c]])
        assert np.array_equal(bins["weight"].to_numpy()[off:off + nb[c]], w[hoff:hoff + nb[c]])
        off += nb[c]; hoff += genome[c] // res + 1
    # genome-wide store: blocks keyed 'c' and 'c1_c2', its own bin table
    table, total = ho.chro_bins(genome, res)
    W = rng.poisson(0.7, size=(total, total)); W = np.triu(W) + np.triu(W, 1).T
    gw = MatrixStore(str(tmp_path / "whole.npz"))
    gw.add(res, ho.whole_matrix_to_sparse_dict(table, W), table)
    bins2, pix2, _ = cool_export.store_tables(np.load(gw.save(), allow_pickle=True), res, genome)
    assert len(pix2) == int((np.triu(W) != 0).sum())
    # the same ids as the dense matrix when no chromosome length is a multiple of the resolution
    if all(genome[c] % res for c in genome):
        R = np.zeros_like(W)
        R[pix2["bin1_id"], pix2["bin2_id"]] = pix2["count"]
        assert np.array_equal(R, np.triu(W))
    with pytest.raises(RuntimeError):
        cool_export.write_cool(path, small_genome_file, str(tmp_path / "x.mcool"))       # cooler is not installed here


def test_entry_word_layout_invariants():
    """Sort-path entries: (row << (cb + vb)) | (col << vb) | count must leave bit 63 clear (torch int64 / searchsorted on the
    raw words), keep the padding key ~0 strictly above every real entry on the sorted bits, and the count field must hold
    what the int32 CSR can hold (or less, with the overflow fallback)."""
    from hichap_master_b200 import kernels
    for nbins in (1, 2, 3, 255, 256, 257, 75_918, 303_641, 524_288, 607_271, 3_000_000, (1 << 24) - 1, 1 << 24):
        cb, vb = kernels.key_col_bits(nbins), kernels.entry_cnt_bits(nbins)
        assert (1 << cb) >= nbins and (cb == 1 or (1 << (cb - 1)) < nbins)
        assert 1 <= vb <= 31 and 2 * cb + vb <= 63
        top = (((nbins - 1) << cb | (nbins - 1)) << vb) | ((1 << vb) - 1)       # the largest real entry
        assert 0 < top < (1 << 63)
        # bits the sorts look at: two bin fields (+ the padding bit when nbins is a power of two)
        nbits = kernels.key_sort_bits(nbins)
        assert nbits == 2 * cb + (1 if nbins == (1 << cb) else 0) and vb + nbits <= 64
        pad_key = ((1 << 64) - 1) >> vb & ((1 << nbits) - 1)
        assert pad_key > (top >> vb)                                             # padding sorts strictly last
        # swapping row and col keeps the word in range and is an involution
        r, c, v = nbins - 1, 0, 5
        e = ((r << cb | c) << vb) | v
        sw = ((c << cb | r) << vb) | v
        assert ((sw >> vb) & ((1 << cb) - 1)) == r and (sw >> (cb + vb)) == c and (e & ((1 << vb) - 1)) == (sw & ((1 << vb) - 1))


def test_entry_count_bits_test_hook(monkeypatch):
    from hichap_master_b200 import kernels
    monkeypatch.setenv("HC_ENTRY_CNT_BITS", "6")
    assert kernels.entry_cnt_bits(300) == 6
    monkeypatch.setenv("HC_ENTRY_CNT_BITS", "40")
    assert kernels.entry_cnt_bits(300) == 31          # never beyond what the layout leaves
