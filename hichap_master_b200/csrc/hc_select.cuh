// Single-CTA order statistics (radix select on order-preserving 64-bit keys) and the two
// NumPy reductions built on them that the reference path relies on:
//   np.percentile(..., q) with the default 'linear' method  (matrixBuilding.py:920, :1004, :884)
//   np.median                                                (cooler balance MAD-max filter)
// Exactness matters: the gap-row lists and the filter masks are compared bit-for-bit with the
// oracle, so the interpolation below reproduces NumPy's floating-point expression order and
// uses explicitly un-fused arithmetic.
#pragma once
#include "hc_common.cuh"

struct HcSelectSmem {
    unsigned int hist[256];
    unsigned long long bcast[2];
    double red[32];
    long long redll[32];
};

// k-th smallest (0-based) key among {keyf(i) : i in [0,n), validf(i)}.  Requires 0 <= k < count.
template <class KeyF, class ValidF>
__device__ unsigned long long block_select_kth(KeyF keyf, ValidF validf, long long n, long long k,
                                               HcSelectSmem* sm) {
    // The keys of one call are close to each other (log-marginals, coverages: same sign and exponent), so their leading
    // bytes coincide -- and a radix pass over a byte that is the same for every key is the slowest kind: all threads
    // increment ONE shared-memory counter (0.5 ms of the 0.8 ms of the C2 filters went there).  One min / max reduction
    // finds the common leading bytes; the passes start below them.
    unsigned long long kmin = ~0ull, kmax = 0ull;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        if (validf(i)) {
            const unsigned long long key = keyf(i);
            kmin = key < kmin ? key : kmin;
            kmax = key > kmax ? key : kmax;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
    }
    __syncthreads();
    {
        unsigned long long* slots = reinterpret_cast<unsigned long long*>(sm->redll);      // 32 x 8 bytes: min; max goes to sm->red
        unsigned long long* slots2 = reinterpret_cast<unsigned long long*>(sm->red);
        if ((threadIdx.x & 31) == 0) { slots[threadIdx.x >> 5] = kmin; slots2[threadIdx.x >> 5] = kmax; }
        __syncthreads();
        kmin = ~0ull; kmax = 0ull;
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) {
            kmin = slots[w] < kmin ? slots[w] : kmin;
            kmax = slots2[w] > kmax ? slots2[w] : kmax;
        }
        __syncthreads();
    }
    int top = 56;                                   // highest byte in which two keys differ
    while (top > 0 && ((kmin ^ kmax) >> top) == 0ull) top -= 8;
    unsigned long long mask = top == 56 ? 0ull : (~0ull << (top + 8));
    unsigned long long prefix = kmin & mask;
    for (int shift = top; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) sm->hist[i] = 0;
        __syncthreads();
        // One CTA streams the keys from L2, so the pass is bound by load latency: four independent elements per trip.
        // (Merging equal digits inside a warp with match.any before the atomic was measured and is slower: +10 ms on the
        // 303 k bins of C4 -- the convergence barrier serialises the loads.)
        for (long long base = threadIdx.x; base < n; base += 4ll * blockDim.x) {
            bool ok[4];
            unsigned long long key[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long i = base + (long long)u * blockDim.x;
                ok[u] = i < n && validf(i);
                key[u] = ok[u] ? keyf(i) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (ok[u] && (key[u] & mask) == prefix) atomicAdd(&sm->hist[(key[u] >> shift) & 255ull], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {                    // warp 0: lane l owns digits 8l .. 8l+7
            const int lane = threadIdx.x;
            unsigned int h[8];
            long long s = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) { h[e] = sm->hist[8 * lane + e]; s += h[e]; }
            long long incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            long long run = incl - s;
            if ((k >= run && k < incl) || (lane == 31 && k >= incl)) {   // the second case cannot happen for k < count
                unsigned long long digit = 8ull * lane + 7ull;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (k < run + h[e]) { digit = 8ull * lane + e; break; }
                    if (e < 7) run += h[e];
                }
                sm->bcast[0] = digit;
                sm->bcast[1] = (unsigned long long)(k - run);
            }
        }
        __syncthreads();
        prefix |= sm->bcast[0] << shift;
        mask |= 0xffull << shift;
        k = (long long)sm->bcast[1];
        __syncthreads();
    }
    return prefix;
}

// (k-th, (k+1)-th) smallest values of the valid elements; second == first when k+1 >= count.
template <class ValF, class ValidF>
__device__ void block_select_pair(ValF valf, ValidF validf, long long n, long long count, long long k,
                                  HcSelectSmem* sm, double* lo, double* hi) {
    auto keyf = [&](long long i) { return f64_key(valf(i)); };
    unsigned long long klo = block_select_kth(keyf, validf, n, k, sm);
    // count elements <= klo, and the smallest key > klo
    long long le = 0;
    unsigned long long nxt = ~0ull;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        if (validf(i)) {
            unsigned long long key = keyf(i);
            if (key <= klo) ++le; else if (key < nxt) nxt = key;
        }
    }
    le = block_sum_ll(le, sm->redll);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, nxt, o);
        nxt = other < nxt ? other : nxt;
    }
    __syncthreads();
    unsigned long long* slots = reinterpret_cast<unsigned long long*>(sm->red);
    if ((threadIdx.x & 31) == 0) slots[threadIdx.x >> 5] = nxt;
    __syncthreads();
    nxt = ~0ull;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) nxt = slots[w] < nxt ? slots[w] : nxt;
    __syncthreads();
    *lo = key_f64(klo);
    if (k + 1 >= count) *hi = *lo;
    else *hi = (le >= k + 2) ? *lo : key_f64(nxt);
}

// np.percentile(valid values, 100*q) with method='linear'.  count == 0 -> NaN.
template <class ValF, class ValidF>
__device__ double block_percentile(ValF valf, ValidF validf, long long n, long long count, double q,
                                   HcSelectSmem* sm) {
    if (count <= 0) return __longlong_as_double(0x7ff8000000000000ll);
    // numpy: virtual_index = n*q + (alpha + q*(1 - alpha - beta)) - 1 with alpha = beta = 1
    double vi = __dadd_rn(__dadd_rn(__dmul_rn((double)count, q), __dadd_rn(1.0, __dmul_rn(q, -1.0))), -1.0);
    long long prev;
    double lo, hi, t;
    if (vi >= (double)(count - 1)) {
        prev = count - 1;
        block_select_pair(valf, validf, n, count, prev, sm, &lo, &hi);
        hi = lo;
        t = __dadd_rn(vi, -(double)prev);  // numpy computes gamma before clipping the indexes
        t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    } else if (vi < 0.0) {
        block_select_pair(valf, validf, n, count, 0, sm, &lo, &hi);
        hi = lo;
        t = 0.0;
    } else {
        prev = (long long)floor(vi);
        block_select_pair(valf, validf, n, count, prev, sm, &lo, &hi);
        t = __dadd_rn(vi, -(double)prev);
    }
    // numpy _lerp: a + (b-a)*t, replaced by b - (b-a)*(1-t) where t >= 0.5
    double d = __dadd_rn(hi, -lo);
    if (t >= 0.5) return __dadd_rn(hi, -__dmul_rn(d, __dadd_rn(1.0, -t)));
    return __dadd_rn(lo, __dmul_rn(d, t));
}

// np.median(valid values).  count == 0 -> NaN.
template <class ValF, class ValidF>
__device__ double block_median(ValF valf, ValidF validf, long long n, long long count, HcSelectSmem* sm) {
    if (count <= 0) return __longlong_as_double(0x7ff8000000000000ll);
    double lo, hi;
    if (count & 1) {
        block_select_pair(valf, validf, n, count, count / 2, sm, &lo, &hi);
        return lo;
    }
    block_select_pair(valf, validf, n, count, count / 2 - 1, sm, &lo, &hi);
    return __dmul_rn(__dadd_rn(lo, hi), 0.5);
}
