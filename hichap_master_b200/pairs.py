"""Valid-pair text -> columnar arrays (the step right before the binning kernel).

Mirrors the per-line logic of HiCHap/matrixBuilding.py:573-580 (23-column ``*_Valid.bed``:
chromosomes in columns 1 and 8, fragment mid-points in columns 6 and 13, format documented at
filtering.py:16-47) and :822-829 / :1131-1141 (4/5-column allelic beds: c1,p1,c2,p2[,mark]).
Parsing is host work; the kernels consume the columnar result.
"""
from __future__ import annotations

import io
import os

import numpy as np
import pandas as pd

MARK_ID = {"Both": 0, "R1": 1, "R2": 2}
MARK_OTHER = 3


def chrom_passes(c: str, chroms) -> bool:
    """matrixBuilding.py:360 / :577."""
    return (not chroms) or (c.isdigit() and ("#" in chroms)) or (c in chroms)


def _as_text_stream(bed_io):
    if isinstance(bed_io, (str, bytes)) and not isinstance(bed_io, io.IOBase):
        raise TypeError("bed_IO must be an iterable / file-like stream of lines, as in the reference")
    if hasattr(bed_io, "read"):
        return bed_io
    lines = list(bed_io)
    if lines and isinstance(lines[0], bytes):
        return io.BytesIO(b"".join(l if l.endswith(b"\n") else l + b"\n" for l in lines))
    return io.StringIO("".join(l if l.endswith("\n") else l + "\n" for l in lines))


def _chrom_ids(names: pd.Series, order, chroms) -> np.ndarray:
    """name column -> index into the sorted chromosome table; -1 = dropped by the filter.
    A name that passes the filter but is missing from the genomeSize table raises KeyError,
    as the reference's dictionary lookup does (matrixBuilding.py:584)."""
    cid = {c: i for i, c in enumerate(order)}
    uniq = names.unique()
    table = {}
    for raw in uniq:
        c = str(raw).lstrip("chr")
        if not chrom_passes(c, chroms):
            table[raw] = -1
        else:
            if c not in cid:
                raise KeyError(c)
            table[raw] = cid[c]
    return names.map(table).to_numpy(np.int32)


def read_pairs(bed_io, order, chroms, layout="valid23"):
    """Returns (c1, p1, c2, p2, mark) NumPy columns; mark is None for the 23-column layout."""
    stream = _as_text_stream(bed_io)
    empty = (np.zeros(0, np.int32),) * 4
    try:
        if layout == "valid23":
            df = pd.read_csv(stream, sep=r"\s+", header=None, usecols=[1, 6, 8, 13], dtype=str,
                             engine="c", na_filter=False)
            df.columns = ["c1", "p1", "c2", "p2"]
            last = None
        else:
            df = pd.read_csv(stream, sep=r"\s+", header=None, names=[0, 1, 2, 3, 4], dtype=str,
                             engine="c", na_filter=False)
            last = df[4].where(df[4] != "", df[3])   # line[-1] of a 4-column line is column 3
            df = df[[0, 1, 2, 3]]
            df.columns = ["c1", "p1", "c2", "p2"]
    except pd.errors.EmptyDataError:
        return empty + ((None if layout == "valid23" else np.zeros(0, np.uint8)),)
    c1 = _chrom_ids(df["c1"], order, chroms)
    c2 = _chrom_ids(df["c2"], order, chroms)
    p1 = df["p1"].astype(np.int64).to_numpy()
    p2 = df["p2"].astype(np.int64).to_numpy()
    if (p1.max(initial=0) > np.iinfo(np.int32).max) or (p2.max(initial=0) > np.iinfo(np.int32).max):
        raise OverflowError("fragment mid-point does not fit int32")
    mark = None
    if last is not None:
        mark = last.map(MARK_ID).fillna(MARK_OTHER).to_numpy(np.uint8)
    return c1, p1.astype(np.int32), c2, p2.astype(np.int32), mark


def read_pair_files(paths, order, chroms, layout="valid23", nthreads=0):
    """Native multithreaded parser (``hc_ingest_parse``) for bed FILES: same columns as
    ``read_pairs`` except that lines dropped by the chromosome filter are omitted instead of
    being kept with chromosome -1.  Mirrors `cat files | per-line loop` of the reference."""
    import ctypes as C
    from . import _abi
    lib = _abi.lib()
    paths = [os.fsencode(p) for p in paths]
    arr = lambda xs: (C.c_char_p * max(len(xs), 1))(*xs)
    names = [c.encode() for c in order]
    filt = [c.encode() for c in chroms]
    n = C.c_int64(0)
    rc = lib.hc_ingest_parse(arr(paths), len(paths), 0 if layout == "valid23" else 1, arr(names), len(names),
                             arr(filt), len(filt), int(nthreads), C.byref(n))
    if rc != 0:
        msg = lib.hc_last_error().decode("utf-8", "replace")
        if msg.startswith("KeyError: "):
            raise KeyError(msg[len("KeyError: "):])
        if msg.startswith("IOError: "):
            raise IOError(msg[len("IOError: "):])
        raise ValueError(msg)
    n = int(n.value)
    cols = [np.empty(n, np.int32) for _ in range(4)]
    mark = np.empty(n, np.uint8) if layout != "valid23" else None
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    _abi.check(lib.hc_ingest_fetch(ptr(cols[0]), ptr(cols[1]), ptr(cols[2]), ptr(cols[3]),
                                   ptr(mark) if mark is not None else None), "hc_ingest_fetch")
    return cols[0], cols[1], cols[2], cols[3], mark
