"""Partitioning of the matrix stage across GPUs.

Intra-chromosomal configs shard by chromosome with no communication: cis-only balancing runs
an independent loop per chromosome and TwoStepCorrection is per chromosome
(matrixBuilding.py:1031-1039), so chromosomes are assigned to ranks by longest-processing-time
greedy on their cost (dense: N^2; sparse: nnz).  A genome-wide matrix is split into contiguous
row blocks balanced by nnz; each ICE iteration then needs one allreduce of the marginal vector.
"""
from __future__ import annotations

import numpy as np


def lpt_assign(costs, nranks: int):
    """Longest-processing-time greedy.  Returns a list (per rank) of item indices, each list in
    ascending index order; deterministic (ties -> lower index first, lower rank first)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * nranks
    out = [[] for _ in range(nranks)]
    for i in order:
        r = min(range(nranks), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    return [sorted(x) for x in out]


def chromosome_shards(sizes, nranks: int, nnz=None):
    """Chromosome -> rank assignment for the dense (cost N^2) or sparse (cost nnz) path."""
    cost = np.asarray(sizes, dtype=np.float64) ** 2 if nnz is None else np.asarray(nnz, dtype=np.float64)
    return lpt_assign(cost, nranks)


def row_block_splits(row_ptr, nranks: int):
    """Contiguous row blocks of an upper-triangular CSR with ~equal nnz: returns nranks+1 row
    boundaries.  Rows near the top of each chromosome are denser in upper-triangular form, so
    the split is on cumulative nnz, not on row count."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    nrows, total = row_ptr.size - 1, int(row_ptr[-1])
    cuts = [0]
    for r in range(1, nranks):
        cuts.append(int(np.searchsorted(row_ptr, total * r / nranks, side="left")))
    cuts.append(nrows)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts
