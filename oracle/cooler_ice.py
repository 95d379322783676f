"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the ICE balancing that
HiCHap delegates to the third-party command ``cooler balance``.

PARITY UNPINNED: the arithmetic is NOT in /root/reference.  HiCHap shells out to
``cooler balance --ignore-diags 1 [--cis-only] --force <file>::<res>``
(matrixBuilding.py:708, :713, :1537, :1542, :1761, :1766).  ``cooler`` is an
un-vendored, un-pinned dependency (README.md:27; no install_requires in
setup.py:23-38) and is not installed in the build container, and the reference
ships no tests / golden vectors for this boundary.  This file restates the
published algorithm of ``cooler.balance.balance_cooler`` (cooler 0.8.x, the
Python-2.7-compatible series contemporaneous with matrixBuilding.py's May-2019
header) with the CLI defaults that HiCHap's command lines leave in force:

    mad_max=5  min_nnz=10  min_count=0  tol=1e-5  max_iters=200
    rescale_marginals=True  ignore_diags=1 (passed)  cis_only (passed for local res)
    non-convergence: weights are still stored (policy ``store_final``)

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

Input convention (what a .cool file holds): upper-triangular pixels
``bin1 <= bin2`` with integer ``count``; ``chrom_offsets`` has nchrom+1 entries.
"""
from __future__ import annotations

import numpy as np


def marginalize(bin1, bin2, data, n):
    """cooler ``_marginalize``: each pixel contributes to both of its bins."""
    # cooler reduces the per-chunk marginals onto np.zeros(n_bins): the result is float64
    return (np.zeros(n) + np.bincount(bin1, weights=data, minlength=n)
            + np.bincount(bin2, weights=data, minlength=n))


def _mad(x):
    return np.median(np.abs(x - np.median(x)))


def _masked_counts(bin1, bin2, count, chrom_offsets, cis_only, ignore_diags):
    """cooler ``_zero_trans`` (cis_only) and ``_zero_diags`` applied to a copy."""
    data = np.array(count, dtype=float)
    if cis_only:
        chrom_of = np.searchsorted(np.asarray(chrom_offsets)[1:], np.arange(chrom_offsets[-1]),
                                   side="right")
        data[chrom_of[bin1] != chrom_of[bin2]] = 0
    if ignore_diags:
        data[np.abs(bin1 - bin2) < ignore_diags] = 0
    return data


def initial_bias(bin1, bin2, count, n, chrom_offsets, cis_only=False, ignore_diags=1,
                 mad_max=5, min_nnz=10, min_count=0):
    """The pre-iteration bin filters of ``balance_cooler``: min_nnz on binarised
    marginals, min_count, then MAD-max on per-chromosome-median-normalised
    log-marginals.  Returns (bias0, masked float counts)."""
    data0 = _masked_counts(bin1, bin2, count, chrom_offsets, cis_only, ignore_diags)
    bias = np.ones(n, float)
    if min_nnz > 0:
        nnz_marg = marginalize(bin1, bin2, (data0 != 0).astype(float), n)
        bias[nnz_marg < min_nnz] = 0
    marg = marginalize(bin1, bin2, data0, n)
    if min_count:
        bias[marg < min_count] = 0
    if mad_max > 0:
        with np.errstate(invalid="ignore", divide="ignore"), \
                __import__("warnings").catch_warnings():
            __import__("warnings").simplefilter("ignore")
            for lo, hi in zip(chrom_offsets[:-1], chrom_offsets[1:]):
                c_marg = marg[lo:hi]
                marg[lo:hi] /= np.median(c_marg[c_marg > 0])
            log_nz = np.log(marg[marg > 0])
            med = np.median(log_nz)
            cutoff = np.exp(med - mad_max * _mad(log_nz))
            bias[marg < cutoff] = 0
    return bias, data0


def _iterate(bias, lo, hi, bin1, bin2, data0, n, tol, max_iters):
    """One independent balancing loop on bias[lo:hi] (the whole vector when
    genome-wide).  Returns (scale, var, iters, converged)."""
    var, iters, nz = 0.0, 0, np.array([])
    for _ in range(max_iters):
        iters += 1
        marg = marginalize(bin1, bin2, bias[bin1] * bias[bin2] * data0, n)[lo:hi]
        nz = marg[marg != 0]
        if nz.size == 0:
            bias[lo:hi] = np.nan
            return np.nan, 0.0, iters, True
        marg = marg / nz.mean()
        marg[marg == 0] = 1
        bias[lo:hi] /= marg
        var = nz.var()
        if var < tol:
            break
    return nz.mean(), var, iters, bool(var < tol)


def balance(bin1, bin2, count, n, chrom_offsets, cis_only=False, ignore_diags=1,
            mad_max=5, min_nnz=10, min_count=0, tol=1e-5, max_iters=200,
            rescale_marginals=True):
    """Restatement of ``cooler.balance.balance_cooler``.

    Returns (weight[n] float64 with NaN for filtered bins, stats) where stats has
    ``scale`` (array per chromosome when cis_only), ``var``, ``converged``,
    ``iters`` (list per chromosome when cis_only)."""
    bin1 = np.asarray(bin1, np.int64)
    bin2 = np.asarray(bin2, np.int64)
    chrom_offsets = np.asarray(chrom_offsets, np.int64)
    bias, data0 = initial_bias(bin1, bin2, count, n, chrom_offsets, cis_only, ignore_diags,
                               mad_max, min_nnz, min_count)
    if cis_only:
        nchrom = len(chrom_offsets) - 1
        scales = np.ones(nchrom)
        iters_all, conv_all, var = [], [], 0.0
        for c in range(nchrom):
            lo, hi = int(chrom_offsets[c]), int(chrom_offsets[c + 1])
            # pixels are sorted by bin1, and cis pixels of c have lo <= bin1 < hi; trans
            # pixels are already zero in data0, so restricting to this span is exact
            plo, phi = np.searchsorted(bin1, [lo, hi], side="left")
            sl = slice(plo, phi)
            scale, var, it, conv = _iterate(bias, lo, hi, bin1[sl], bin2[sl], data0[sl], n,
                                            tol, max_iters)
            b = bias[lo:hi]
            if not np.isnan(scale):
                b[b == 0] = np.nan
            scales[c] = scale
            if rescale_marginals:
                bias[lo:hi] /= np.sqrt(scale)
            iters_all.append(it)
            conv_all.append(conv)
        stats = dict(scale=scales, var=var, converged=bool(var < tol), iters=iters_all,
                     converged_per_chrom=conv_all)
    else:
        scale, var, it, conv = _iterate(bias, 0, n, bin1, bin2, data0, n, tol, max_iters)
        if not np.isnan(scale):
            bias[bias == 0] = np.nan
        if rescale_marginals:
            bias /= np.sqrt(scale)
        stats = dict(scale=scale, var=var, converged=conv, iters=it)
    stats.update(tol=tol, min_nnz=min_nnz, min_count=min_count, mad_max=mad_max,
                 cis_only=cis_only, ignore_diags=ignore_diags, divisive_weights=False)
    return bias, stats


def balance_dense(M, **kw):
    """Convenience for one symmetric dense intra-chromosomal matrix."""
    x, y = np.nonzero(np.triu(M))
    return balance(x, y, M[x, y], M.shape[0], [0, M.shape[0]], **kw)
