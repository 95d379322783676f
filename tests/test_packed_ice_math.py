"""CPU model of the arithmetic of the packed ICE kernel (hc_ice.cu: ice_write_digits / ice_q8_mma_kernel):
the bias as 64-bit fixed point relative to the chromosome's largest bias, split into 8 byte planes; exact
integer plane sums; recombination in fp64.  Checks the numerical claims DESIGN.md makes for it against exact
rational arithmetic -- no GPU, no library call (the kernels themselves are checked against the oracle in
tests/test_gpu_parity.py)."""
import math
from fractions import Fraction

import numpy as np
import pytest


def bias_planes(b):
    """(E, planes[8][n]) as ice_write_digits builds them: F = floor(b * 2^(64 - E)), plane 0 = top byte."""
    mx = float(np.max(b))
    E = math.frexp(mx)[1] if mx > 0 else 0                  # mx < 2^E  (frexp: mx = m * 2^E, 0.5 <= m < 1)
    F = [int(math.ldexp(float(x), 64 - E)) if x > 0 else 0 for x in b]      # ldexp and int() are exact
    assert all(0 <= f < (1 << 64) for f in F)
    planes = np.array([[(f >> (8 * (7 - p))) & 255 for f in F] for p in range(8)], dtype=np.int64)
    return E, planes, F


def kernel_row_sum(v, E, planes):
    """One row: int32-style plane sums, then the kernel's fp64 recombination: lane t4 holds planes 2*t4 and
    2*t4+1 -> c_even * w_even + c_odd * w_odd, scaled by 2^(E-64), then two xor-shuffle additions."""
    C = planes @ v.astype(np.int64)                          # exact (|C| < 2^31 for <= 32768 columns)
    assert np.all(C < (1 << 31))
    lane = []
    for t4 in range(4):
        we, wo = math.ldexp(1.0, 8 * (7 - 2 * t4)), math.ldexp(1.0, 8 * (6 - 2 * t4))
        lane.append((float(C[2 * t4]) * we + float(C[2 * t4 + 1]) * wo) * math.ldexp(1.0, E - 64))
    a = [lane[0] + lane[1], lane[1] + lane[0], lane[2] + lane[3], lane[3] + lane[2]]     # xor 1
    return a[0] + a[2]                                                                      # xor 2


def exact_row_sum(v, b):
    return sum(Fraction(int(x)) * Fraction(float(y)) for x, y in zip(v, b))


@pytest.mark.parametrize("spread,tol", [(0.4, 4e-16), (3.0, 4e-16), (8.0, 1e-13)])
def test_plane_sums_match_exact_arithmetic(spread, tol):
    """Biases within 2^11 of the largest one are represented exactly, so the row sum is the exactly rounded
    dot product up to the three fp64 additions of the recombination; wider ranges lose the low bits of the SMALL
    biases only (absolute error below 2^(E-64) per unit of count)."""
    rng = np.random.default_rng(int(spread * 10))
    n = 4096
    b = np.exp(rng.normal(0, spread, n))
    b[rng.integers(0, n, 50)] = 0.0                              # masked bins
    E, planes, F = bias_planes(b)
    if spread <= 3.0 and b[b > 0].min() * 2048 >= b.max():
        assert all(Fraction(f, 1 << (64 - E)) == Fraction(float(x)) for f, x in zip(F, b))      # exact representation
    for _ in range(20):
        v = rng.poisson(1.5, n).clip(0, 255).astype(np.uint8)
        v[rng.integers(0, n, 30)] = 255
        got = kernel_row_sum(v, E, planes)
        ref = exact_row_sum(v, b)
        assert ref > 0
        assert abs(Fraction(got) - ref) / ref <= tol
        # truncation bound: every bias is cut by less than 2^(E-64)
        assert abs(Fraction(got) - ref) <= Fraction(int(v.astype(np.int64).sum()) + 4, 1 << (64 - E)) + ref * Fraction(1, 1 << 51)


def test_zero_rows_and_zero_bias_are_exact_zeros():
    """cooler drops bins with marg == 0 from mean / variance: an all-zero row must give exactly 0.0."""
    b = np.array([0.0, 1.25, 3.5, 0.0, 1e-3])
    E, planes, _ = bias_planes(b)
    assert kernel_row_sum(np.zeros(5, np.uint8), E, planes) == 0.0
    assert kernel_row_sum(np.array([7, 0, 0, 9, 0], np.uint8), E, planes) == 0.0       # counts only against masked bins
    assert kernel_row_sum(np.array([0, 2, 0, 0, 0], np.uint8), E, planes) == 2.5


def test_int32_plane_sums_cannot_overflow_within_a_segment():
    """A K segment is at most 1024 k-tiles = 32768 columns: 255 * 255 * 32768 < 2^31."""
    assert 255 * 255 * 32768 < 2 ** 31 <= 255 * 255 * (32768 + 512)      # and one more 512-column chunk would not fit
