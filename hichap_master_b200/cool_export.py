"""Bridge from the ``.npz`` matrix stores to the reference's on-disk contract, the multi-resolution
``.cool`` file (``NPZ2Cooler``, matrixBuilding.py:100-303; ``cooler balance`` puts the weights in
``bins/weight``).  ``cooler`` / ``h5py`` are not part of this image, so the stage itself writes ``.npz``
stores (``construction.MatrixStore``); this module turns a store into exactly the tables
``cooler.create_cooler`` takes -- the bin table of ``cooler.binnify`` with a ``weight`` column, and the
upper-triangular pixel table sorted by (bin1_id, bin2_id) -- and writes the file when ``cooler`` is
importable.  Host-side glue only (NumPy / pandas), no device work.

    python -m hichap_master_b200.cool_export Traditional_Multi.npz genomeSize out.mcool [-C '#' X]
"""
from __future__ import annotations

import argparse
import ast

import numpy as np
import pandas as pd

from .matrixBuilding import Load_Genome, Sort_Chromosomes


def binnify(genome: dict, order, res: int) -> pd.DataFrame:
    """``cooler.binnify``: fixed-width bins per chromosome, the last one clipped to the chromosome end
    (ceil(len / res) bins -- one fewer than HiCHap's ``len // res + 1`` when ``res`` divides ``len``)."""
    chrom, start, end = [], [], []
    for c in order:
        n = -(-genome[c] // res)
        s = np.arange(n, dtype=np.int64) * res
        chrom.extend([c] * n)
        start.append(s)
        end.append(np.minimum(s + res, genome[c]))
    return pd.DataFrame({"chrom": pd.Categorical(chrom, categories=list(order), ordered=True),
                         "start": np.concatenate(start) if start else np.zeros(0, np.int64),
                         "end": np.concatenate(end) if end else np.zeros(0, np.int64)})


def store_resolutions(store) -> list:
    return sorted({int(k.split("|")[0]) for k in store.files if k.split("|")[0].isdigit()})


def store_tables(store, res: int, genome: dict):
    """(bins, pixels) of one resolution of a store.  Keys ``res|c`` hold the upper-triangular records of
    chromosome c with chromosome-local bin indices; ``res|c1_c2`` the records of an inter-chromosomal block
    (matrixBuilding.py:457-524).  ``weight|res`` is laid out over HiCHap's bins (len // res + 1 per
    chromosome, chromosomes in the order of the store's own bin table when it has one, else sorted)."""
    keys = [k for k in store.files if k.startswith("%d|" % res)]

    def base(c):                                             # 'M1' / 'P1' haplotype labels share chromosome 1's length
        return c if c in genome else c[1:]

    if "bins|%d" % res in store.files:                       # genome-wide store: its own chromosome order / offsets
        table = store["bins|%d" % res]
        order = [str(c) for c in table["chrom"]]
        hichap_off = {str(c): int(s) for c, s in zip(table["chrom"], table["start"])}
    else:
        names = [k.split("|", 1)[1] for k in keys if "_" not in k.split("|", 1)[1]]
        plain = Sort_Chromosomes(sorted({base(c) for c in names}))
        order = sorted(names, key=lambda c: ("" if c in genome else c[0], plain.index(base(c))))
        hichap_off, run = {}, 0
        for c in order:
            hichap_off[c] = run
            run += genome[base(c)] // res + 1
    sizes = {c: genome[base(c)] for c in order}
    bins = binnify(sizes, order, res)
    cool_n = {c: -(-sizes[c] // res) for c in order}
    cool_off, run = {}, 0
    for c in order:
        cool_off[c] = run
        run += cool_n[c]
    b1, b2, cnt = [], [], []
    for k in keys:
        name = k.split("|", 1)[1]
        ca, cb = (name, name) if "_" not in name else name.split("_")
        if ca not in cool_off or cb not in cool_off:
            continue
        rec = store[k]
        if rec.size == 0:
            continue
        if int(rec["bin1"].max()) >= cool_n[ca] or int(rec["bin2"].max()) >= cool_n[cb]:
            raise ValueError("records of %s reach beyond the chromosome end" % name)
        b1.append(rec["bin1"].astype(np.int64) + cool_off[ca])
        b2.append(rec["bin2"].astype(np.int64) + cool_off[cb])
        cnt.append(rec["IF"])
    if b1:
        b1, b2, cnt = np.concatenate(b1), np.concatenate(b2), np.concatenate(cnt)
        o = np.lexsort((b2, b1))
        b1, b2, cnt = b1[o], b2[o], cnt[o]
    else:
        b1 = b2 = np.zeros(0, np.int64)
        cnt = np.zeros(0, np.float64)
    integral = cnt.size == 0 or bool(np.all(cnt == np.rint(cnt)))
    pixels = pd.DataFrame({"bin1_id": b1, "bin2_id": b2, "count": cnt.astype(np.int32) if integral else cnt})
    attrs = None
    if "weight|%d" % res in store.files:
        w = np.asarray(store["weight|%d" % res], np.float64)
        bins["weight"] = np.concatenate([w[hichap_off[c]:hichap_off[c] + cool_n[c]] for c in order]) if order else w[:0]
        if "weight_attrs|%d" % res in store.files:
            try:
                attrs = ast.literal_eval(str(store["weight_attrs|%d" % res]))
            except (ValueError, SyntaxError):
                attrs = None
    return bins, pixels, attrs


def write_cool(store_path: str, genomeSize: str, out_path: str, chroms=("#", "X")):
    """One ``out_path::res`` cooler per resolution of the store (the layout NPZ2Cooler writes)."""
    try:
        import cooler
    except ImportError as e:                                   # not in this image
        raise RuntimeError("writing .cool files needs the `cooler` package (and h5py)") from e
    genome = Load_Genome(genomeSize, list(chroms))
    store = np.load(store_path, allow_pickle=True)
    for i, res in enumerate(store_resolutions(store)):
        bins, pixels, attrs = store_tables(store, res, genome)
        uri = "%s::%d" % (out_path, res)
        cooler.create_cooler(uri, bins, pixels, ordered=True, mode="w" if i == 0 else "a",
                             dtypes={"count": pixels["count"].dtype})
        if attrs is not None:
            import h5py
            with h5py.File(out_path, "r+") as h5:
                h5["%d/bins/weight" % res].attrs.update({k: v for k, v in attrs.items() if np.isscalar(v)})
    return out_path


def main(argv=None):
    ap = argparse.ArgumentParser(description="npz matrix store -> multi-resolution .cool (needs cooler)")
    ap.add_argument("store")
    ap.add_argument("genomeSize")
    ap.add_argument("out")
    ap.add_argument("-C", "--chroms", nargs="*", default=["#", "X"])
    a = ap.parse_args(argv)
    print(write_cool(a.store, a.genomeSize, a.out, a.chroms))


if __name__ == "__main__":
    main()
