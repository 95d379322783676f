// (b) ICE balancing of a batch of dense symmetric tiles: the SYMMETRIC packed path (default of hc_ice_dense_balance).
// Same algorithm as hc_ice.cu (cooler balance restated in oracle/cooler_ice.py; HiCHap call sites
// matrixBuilding.py:708, :713, :1537, :1542, :1761, :1766).
//
// Two ideas on top of the packed uint8 + tensor-core kernel of hc_ice.cu:
//
// 1. Only the UPPER triangle is stored and streamed.  A chromosome is cut into 256 x 256 blocks; block (I, J), I <= J,
//    is 64 KB of uint8 cells min(w * count, 255) (cells below the diagonal of a diagonal block are zero, the diagonal
//    holds half of its weight).  One pass over a block yields BOTH the partial row sums of its rows (block x bias of
//    the columns) and the partial column sums of its columns (block^T x bias of the rows) -- by symmetry the latter
//    are the row sums of the mirrored block that is never stored.  The block is staged in shared memory by ONE bulk
//    asynchronous copy (cp.async.bulk + mbarrier, double buffered); the plain fragment comes out with ldmatrix, the
//    transposed one with ldmatrix.trans (b16 granularity) + 4 PRMT, and both products are exact integer
//    mma.sync.m16n8k32.u8.u8.s32 against the byte planes of the 64-bit fixed-point bias (tools/sym_fragment_model.py
//    checks the register-level data movement on the CPU).  Bytes per iteration: N^2 / 2 instead of N^2 (hc_ice.cu
//    packed) or 4 N^2 (int32 tiles).
//
// 2. No launch and no grid-wide synchronisation per iteration.  Cis-only balancing iterates every chromosome on its
//    own, so ONE persistent kernel runs the whole loop as a dataflow: CTAs draw blocks of any chromosome that has
//    some; the CTA that completes the last block touching block index b reduces the partial sums of those 256 bins
//    (phase A); the CTA that completes the last phase A of a chromosome does the O(n) update -- mean / variance,
//    bias update, convergence test, new byte planes (phase B) -- and re-opens the chromosome's ticket counter for
//    the next iteration.  While one chromosome is in its (latency-bound) update the others keep the HBM pipe busy.
//    Every sum has a fixed order, so results are deterministic.
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <vector>
#include "hc_common.cuh"

namespace {

constexpr int BLK = 256;                     // block side (bins)
constexpr int BLK_BYTES = BLK * BLK;         // 64 KB
constexpr int SYM_THREADS = 512;             // 16 warps: 4 (row groups of 64) x 4 (column groups of 64)
constexpr int MAX_GROUPS = 4;                // phase B: groups of 4 bins per thread -> n <= 512 * 4 * 4 = 8192
constexpr int MAX_BINS = SYM_THREADS * 4 * MAX_GROUPS;
constexpr int MAX_BLOCKS = MAX_BINS / BLK;   // 32

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// weighted count of cell (r, j) under cooler's _zero_diags + _marginalize on upper-triangular pixels (see hc_ice.cu)
__device__ __forceinline__ long long wcount(int v, int j, int r, int kd) {
    const int d = j - r;
    if (d == 0) return kd == 0 ? 2ll * v : 0ll;
    return (d < kd && d > -kd) ? 0ll : (long long)v;
}
// the byte stored for cell (r, j), j >= r, and what is left for the overflow list of row r (both orientations of an
// off-diagonal cell carry the same extra; the diagonal stores half of its weight because U + U^T counts it twice)
__device__ __forceinline__ int stored_byte(int v, int j, int r, int kd, long long* extra) {
    const long long w = wcount(v, j, r, kd);
    if (j == r) {
        const long long half = w >> 1;                       // w is 0 or 2 v
        const long long b = half > 255 ? 255 : half;
        *extra = w - 2 * b;
        return (int)b;
    }
    const long long b = w > 255 ? 255 : (w < 0 ? 0 : w);
    *extra = w - b;
    return (int)b;
}

struct SymTables {              // per chromosome, device
    const int32_t* n;           // bins
    const int32_t* nblk;        // blocks per side
    const int64_t* pad_off;     // offset in the padded vectors (nblk * 256 entries per chromosome)
    const int64_t* part_off;    // offset of the chromosome's partial planes: part[part_off + plane * npad + idx]
    const int32_t* blk_off;     // offset of the chromosome's per-block bookkeeping
    const int32_t* item_first;  // [nprob + 1] prefix of items
    const int32_t* prio;        // chromosomes, largest first (order of the ticket scan)
    const int32_t* rank_of;     // inverse of prio
    const int64_t* bin_off;     // concatenated (unpadded) bins (the caller's table)
};

// ---------------------------------------------------------------------------------------
// pack: int32 tiles -> upper-triangular 256 x 256 blocks of uint8 in ldmatrix order
// ---------------------------------------------------------------------------------------
// A block is [strip 0..15][k-tile 0..7] tiles of 512 B; a tile is four 8 x 8 b16 matrices (h, k) in the order
// (0,0) (1,0) (0,1) (1,1); row i of matrix (h, k) = the 16 cells (row 8h + i, columns 16k .. 16k + 15) of the tile.
// One warp per tile, lane l writes the 16 bytes at tile + 16 l.
__global__ void __launch_bounds__(256)
sym_pack_kernel(const int32_t* __restrict__ mats, const int64_t* __restrict__ mat_off, const int32_t* __restrict__ mat_ld,
                SymTables T, const int4* __restrict__ items, int nitems, int kd, uint8_t* __restrict__ q8,
                unsigned long long* __restrict__ ovf_cnt, int32_t* __restrict__ ovf_lo, int32_t* __restrict__ ovf_hi) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (long long)nitems * 128) return;
    const int it = (int)(w >> 7), tile = (int)(w & 127);
    const int4 d = items[it];
    const int p = d.z, I = d.w & 0xffff, J = d.w >> 16;
    const long long off = ((long long)(uint32_t)d.x) | ((long long)d.y << 32);
    const int n = T.n[p];
    const int64_t ld = mat_ld[p];
    const int32_t* M = mats + mat_off[p];
    const int s = tile >> 3, t = tile & 7;
    const int m = lane >> 3, i = lane & 7, h = m & 1, k = m >> 1;
    const int r = I * BLK + s * 16 + 8 * h + i;
    const int c0 = J * BLK + t * 32 + 16 * k;
    uint32_t word[4] = {0u, 0u, 0u, 0u};
    if (r < n) {
        const int64_t grow = T.bin_off[p] + r;
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
            const int cb = c0 + 4 * v4;
            if (cb < n) {                    // ld is a multiple of 128 >= n: a 4-column group starting below n is inside the row
                const int4 a = ld_stream_v4(M + (int64_t)r * ld + cb);
                const int x[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = cb + e;
                    if (c < n && c >= r) {
                        long long extra;
                        const int b = stored_byte(x[e], c, r, kd, &extra);
                        word[v4] |= (uint32_t)b << (8 * e);
                        if (extra > 0) {     // rare: listed for row r, and for row c when off the diagonal
                            atomicAdd(ovf_cnt + grow, 1ull); atomicMin(ovf_lo + grow, c); atomicMax(ovf_hi + grow, c);
                            if (c != r) {
                                const int64_t gc = T.bin_off[p] + c;
                                atomicAdd(ovf_cnt + gc, 1ull); atomicMin(ovf_lo + gc, r); atomicMax(ovf_hi + gc, r);
                            }
                        }
                    }
                }
            }
        }
    }
    *reinterpret_cast<uint4*>(q8 + off + (long long)tile * 512 + 16 * lane) = make_uint4(word[0], word[1], word[2], word[3]);
}

// exclusive scan of the per-row overflow counts (single CTA; bins are O(1e5)), total -> v[n]
__global__ void __launch_bounds__(1024) sym_scan_kernel(int64_t* __restrict__ v, long long n) {
    __shared__ long long sh[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long b0 = 0; b0 < n; b0 += 1024) {
        const long long i = b0 + threadIdx.x;
        const long long x = i < n ? v[i] : 0;
        sh[threadIdx.x] = x;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const long long t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) v[i] = carry + sh[threadIdx.x] - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) v[n] = carry;
}

// overflow cells of every row, in column order (one warp per global row), from the span the pack kernel found
__global__ void __launch_bounds__(256)
sym_ovf_fill_kernel(const int32_t* __restrict__ mats, const int64_t* __restrict__ mat_off, const int32_t* __restrict__ mat_ld,
                    SymTables T, int nprob, int kd, const int64_t* __restrict__ ovf_ptr, const int32_t* __restrict__ ovf_lo,
                    const int32_t* __restrict__ ovf_hi, int32_t* __restrict__ ovf_col, int32_t* __restrict__ ovf_val) {
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= T.bin_off[nprob]) return;
    int64_t out = ovf_ptr[g];
    if (ovf_ptr[g + 1] == out) return;
    int p = 0;
    while (p + 1 < nprob && T.bin_off[p + 1] <= g) ++p;
    const int r = (int)(g - T.bin_off[p]);
    const int32_t* row = mats + mat_off[p] + (int64_t)r * mat_ld[p];
    const int hi = ovf_hi[g];
    for (int j0 = ovf_lo[g]; j0 <= hi; j0 += 32) {
        const int j = j0 + lane;
        long long extra = 0;
        if (j <= hi) (void)stored_byte(row[j], max(j, r), min(j, r), kd, &extra);      // cell (min, max): same count by symmetry
        const unsigned mk = __ballot_sync(0xffffffffu, extra > 0);
        if (extra > 0) {
            const int64_t at = out + __popc(mk & ((1u << lane) - 1u));
            ovf_col[at] = j;
            ovf_val[at] = (int32_t)extra;
        }
        out += __popc(mk);
    }
}

// ---------------------------------------------------------------------------------------
// the persistent dataflow kernel
// ---------------------------------------------------------------------------------------
struct SymArgs {
    SymTables T; int nprob;
    const uint8_t* q8; const int4* items;       // item: {tile offset lo, hi, chromosome, I | J << 16}
    double* bias; double* marg;                 // padded layout
    uint8_t* dig1; uint8_t* dig2;               // byte planes: B fragments of the columns (per k-tile) / of the rows (per strip pair)
    int32_t* dig_exp;
    double* part;                               // per chromosome: nblk planes of row partials, then nblk planes of column partials
    int32_t* tick; int32_t* blkdone; int32_t* adone; int32_t* iters; int32_t* n_active;
    unsigned long long* avail;                  // bit r: the chromosome of priority rank r has tickets left this iteration
    int32_t* abort_flag;                        // set by a CTA that waited implausibly long for work (dataflow bug guard)
    long long spin_limit;                       // clock64 ticks a CTA may wait for work before it raises abort_flag
    long long* stats;                           // HC_SYM_STATS=1: per CTA {blocks, cycles waiting for work, cycles in phases, cycles
                                                // waiting for a bulk copy, total cycles, waits for work, failed tickets, phases run}
    double* blk_sum; long long* blk_cnt;        // per block: sum / count of the non-zero marginals (phase A -> phase B)
    const int64_t* ovf_ptr; const int32_t* ovf_col; const int32_t* ovf_val;
    hc_ice_result* results; double tol; int max_iters;
};

__device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], const uint2& b) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint2 ldcg_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_volatile_i32(const int32_t* p) {
    int r;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ double block_max_sym(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmax(t, red[w]);
    return t;
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long r;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}

// thread 0: start the bulk copies of a block (64 KB of tiles) and of the bias byte planes of its columns (dig1, 2 KB)
// and rows (dig2, 2 KB) into a buffer; all three complete on the buffer's mbarrier.  The planes were written by the
// phase B that re-opened the chromosome's tickets, i.e. before this block's ticket could be drawn.
constexpr int DIG_BYTES = 8 * 256;                          // 8 k-tiles (or strip pairs) x 256 B
constexpr int BUF_BYTES = BLK_BYTES + 2 * DIG_BYTES;
__device__ __forceinline__ void sym_issue_copy(const SymArgs& A, int item, unsigned char* buf, unsigned long long* bar) {
    const int4 d = __ldg(A.items + item);
    const long long off = ((long long)(uint32_t)d.x) | ((long long)d.y << 32);
    const int p = d.z, I = d.w & 0xffff, J = d.w >> 16;
    const int64_t lo = A.T.pad_off[p];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // earlier generic-proxy reads of buf precede the async writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"((uint32_t)BUF_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(buf)), "l"(A.q8 + off), "r"((uint32_t)BLK_BYTES), "r"(smem_u32(bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(buf + BLK_BYTES)), "l"(A.dig1 + 8 * lo + (int64_t)(8 * J) * 256), "r"((uint32_t)DIG_BYTES), "r"(smem_u32(bar)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(buf + BLK_BYTES + DIG_BYTES)), "l"(A.dig2 + 8 * lo + (int64_t)(8 * I) * 256), "r"((uint32_t)DIG_BYTES), "r"(smem_u32(bar)) : "memory");
}

// thread 0: turn a drawn ticket into an item (or -1).  The drawer of a chromosome's LAST ticket clears its bit in the
// availability mask -- that happens before the block's completion, hence before the phase B that sets the bit again.
__device__ __forceinline__ int sym_ticket_item(const SymArgs& A, int r, int t) {
    const int p = A.T.prio[r];
    const int n = A.T.item_first[p + 1] - A.T.item_first[p];
    if (t >= n) return -1;
    if (t == n - 1) atomicAnd(A.avail, ~(1ull << r));
    return A.T.item_first[p] + t;
}

// byte planes of 4 consecutive bins j4 .. j4+3 (fixed point F = round(b * 2^(64 - E)), plane 0 most significant):
// dig1 = B fragments of product 1 (k-tile of 32 columns: lane 4 * plane + (column % 16) / 4, register column / 16);
// dig2 = B fragments of product 2 (strip pair of 32 rows: k-slot 4q + j <-> row {2q, 2q+1, 8+2q, 9+2q}[j] of a strip,
// register = which strip of the pair)
__device__ __forceinline__ void sym_store_digits(const double (&bv)[4], int j4, double pow2, uint8_t* __restrict__ dig1,
                                                 uint8_t* __restrict__ dig2) {
    unsigned long long F[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) F[e] = bv[e] > 0.0 ? __double2ull_rn(bv[e] * pow2) : 0ull;
    uint32_t* d1 = reinterpret_cast<uint32_t*>(dig1);
    uint16_t* d2 = reinterpret_cast<uint16_t*>(dig2);
    const int t = j4 >> 5, cc = j4 & 31, sub = (cc & 15) >> 2, reg = cc >> 4;
    const int sp = j4 >> 5, rr = j4 & 15, strip = (j4 >> 4) & 1;      // rows rr .. rr+3 of strip `strip` of pair sp
    const int q0 = (rr & 7) >> 1, hi = rr >> 3;                      // rows rr, rr+1 -> slot pair of q0; rr+2, rr+3 -> q0 + 1
#pragma unroll
    for (int pl = 0; pl < 8; ++pl) {
        const int sh = 8 * (7 - pl);
        const uint32_t b0 = (uint32_t)((F[0] >> sh) & 255ull), b1 = (uint32_t)((F[1] >> sh) & 255ull);
        const uint32_t b2 = (uint32_t)((F[2] >> sh) & 255ull), b3 = (uint32_t)((F[3] >> sh) & 255ull);
        d1[t * 64 + (pl * 4 + sub) * 2 + reg] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        // word (lane 4 pl + q, register strip) holds rows {2q, 2q+1} in bytes 0-1 and {8+2q, 9+2q} in bytes 2-3
        d2[(sp * 64 + (pl * 4 + q0) * 2 + strip) * 2 + hi] = (uint16_t)(b0 | (b1 << 8));
        d2[(sp * 64 + (pl * 4 + q0 + 1) * 2 + strip) * 2 + hi] = (uint16_t)(b2 | (b3 << 8));
    }
}

__device__ __forceinline__ int digits_exp(double mx, double* pow2) {
    const int E = (mx > 1.0e-300 && mx < 1.0e300) ? (int)((__double_as_longlong(mx) >> 52) & 0x7ff) - 1023 + 1 : 0;
    *pow2 = __longlong_as_double((long long)(1023 + 64 - E) << 52);
    return E;
}

// phase A: marginals of the 256 bins of block index b of chromosome p (all partial sums are in) + the block's share
// of sum / count over the non-zero marginals
__device__ void sym_phase_a(const SymArgs& A, int p, int b, double* red, long long* redll, double* xch) {
    const int nblk = A.T.nblk[p], n = A.T.n[p];
    const int64_t lo = A.T.pad_off[p], npad = (int64_t)nblk * BLK;
    const double* part = A.part + A.T.part_off[p];
    const int tid = threadIdx.x, idx = b * BLK + (tid & 255);
    double t = 0.0;
    // the partial planes are summed four at a time (independent loads in flight), in a fixed order
    auto plane_sum = [&](int first, int last) {       // planes first .. last-1 at idx
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        int k = first;
        for (; k + 3 < last; k += 4) {
            t0 += __ldcg(part + (int64_t)k * npad + idx); t1 += __ldcg(part + (int64_t)(k + 1) * npad + idx);
            t2 += __ldcg(part + (int64_t)(k + 2) * npad + idx); t3 += __ldcg(part + (int64_t)(k + 3) * npad + idx);
        }
        for (; k < last; ++k) t0 += __ldcg(part + (int64_t)k * npad + idx);
        return (t0 + t1) + (t2 + t3);
    };
    if (tid < 256) {                                  // row partials of blocks (b, J), J = b .. nblk-1
        t = plane_sum(b, nblk);
        const int64_t g = A.T.bin_off[p] + idx;       // + the row's overflow cells
        if (idx < n) {
            const double* bw = A.bias + lo;
            for (int64_t e = A.ovf_ptr[g]; e < A.ovf_ptr[g + 1]; ++e) t = fma((double)A.ovf_val[e], __ldcg(bw + A.ovf_col[e]), t);
        }
    } else {                                          // column partials of blocks (I, b), I = 0 .. b
        t = plane_sum(nblk, nblk + b + 1);
        xch[tid & 255] = t;
    }
    __syncthreads();
    double m = 0.0;
    if (tid < 256) {
        m = idx < n ? __ldcg(A.bias + lo + idx) * (t + xch[tid]) : 0.0;
        A.marg[lo + idx] = m;
    }
    double s = m != 0.0 ? m : 0.0;
    long long c = m != 0.0 ? 1 : 0;
    s = block_sum(s, red);
    c = block_sum_ll(c, redll);
    if (tid == 0) { A.blk_sum[A.T.blk_off[p] + b] = s; A.blk_cnt[A.T.blk_off[p] + b] = c; }
}

// phase B: the O(n) update of chromosome p.  Returns nothing; re-opens the ticket counter unless the chromosome is done.
__device__ void sym_phase_b(const SymArgs& A, int p, double* red, long long* redll) {
    const int nblk = A.T.nblk[p], n = A.T.n[p];
    const int64_t lo = A.T.pad_off[p];
    const int npad = nblk * BLK;
    const int k = __ldcg(A.iters + p) + 1;
    double s = 0.0;
    long long c = 0;
    for (int b = 0; b < nblk; ++b) { s += __ldcg(A.blk_sum + A.T.blk_off[p] + b); c += __ldcg(A.blk_cnt + A.T.blk_off[p] + b); }
    double* bw = A.bias + lo;
    double* mg = A.marg + lo;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    bool finished;
    if (c == 0) {        // nothing left to balance: cooler sets bias = NaN, scale = NaN, var = 0
        for (int j = threadIdx.x; j < npad; j += blockDim.x) bw[j] = nan;
        if (threadIdx.x == 0) { hc_ice_result r; r.scale = nan; r.var = 0.0; r.iters = k; r.converged = 1; A.results[p] = r; }
        finished = true;
    } else {
        const double mean = s / (double)c;
        double m[4 * MAX_GROUPS], bb[4 * MAX_GROUPS];
        double v = 0.0, mx = 0.0;
#pragma unroll
        for (int gq = 0; gq < MAX_GROUPS; ++gq) {
            const int j4 = 4 * ((int)threadIdx.x + gq * SYM_THREADS);
#pragma unroll
            for (int e = 0; e < 4; ++e) { m[4 * gq + e] = 0.0; bb[4 * gq + e] = 0.0; }
            if (j4 < npad) {
                const double2 x = __ldcg(reinterpret_cast<const double2*>(mg + j4)), y = __ldcg(reinterpret_cast<const double2*>(mg + j4 + 2));
                const double2 u = __ldcg(reinterpret_cast<const double2*>(bw + j4)), w = __ldcg(reinterpret_cast<const double2*>(bw + j4 + 2));
                m[4 * gq] = x.x; m[4 * gq + 1] = x.y; m[4 * gq + 2] = y.x; m[4 * gq + 3] = y.y;
                bb[4 * gq] = u.x; bb[4 * gq + 1] = u.y; bb[4 * gq + 2] = w.x; bb[4 * gq + 3] = w.y;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double mm = m[4 * gq + e];
                if (mm != 0.0) { const double d = mm - mean; v += d * d; }
                double q = mm / mean;
                if (q == 0.0) q = 1.0;
                bb[4 * gq + e] = bb[4 * gq + e] / q;
                mx = fmax(mx, bb[4 * gq + e]);
            }
        }
        const double var = block_sum(v, red) / (double)c;
        mx = block_max_sym(mx, red);
        double pow2;
        const int E = digits_exp(mx, &pow2);
#pragma unroll
        for (int gq = 0; gq < MAX_GROUPS; ++gq) {
            const int j4 = 4 * ((int)threadIdx.x + gq * SYM_THREADS);
            if (j4 < npad) {
                double bv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) bv[e] = bb[4 * gq + e];
                *reinterpret_cast<double2*>(bw + j4) = make_double2(bv[0], bv[1]);
                *reinterpret_cast<double2*>(bw + j4 + 2) = make_double2(bv[2], bv[3]);
                sym_store_digits(bv, j4, pow2, A.dig1 + 8 * lo, A.dig2 + 8 * lo);
            }
        }
        finished = var < A.tol || k >= A.max_iters;
        if (threadIdx.x == 0) {
            A.dig_exp[p] = E;
            hc_ice_result r; r.scale = mean; r.var = var; r.iters = k; r.converged = var < A.tol;
            A.results[p] = r;
        }
    }
    // bookkeeping for the next iteration, then (last) re-open the tickets
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) A.blkdone[A.T.blk_off[p] + b] = 0;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        A.iters[p] = k;
        A.adone[p] = 0;
        __threadfence();
        if (finished) atomicSub(A.n_active, 1);
        else { atomicExch(A.tick + p, 0); __threadfence(); atomicOr(A.avail, 1ull << A.T.rank_of[p]); }
    }
}

// CTA-wide: block (I, J) of chromosome p has been counted; fi / fj say whether that completed the partial planes of
// block index I / J.  Runs the phases this CTA has thereby become responsible for.
__device__ void sym_phases(const SymArgs& A, int p, int I, int J, int fi, int fj, double* red, long long* redll, double* xch,
                           int* sh_last) {
    __threadfence();
    if (fi) sym_phase_a(A, p, I, red, redll, xch);
    if (fj) { __syncthreads(); sym_phase_a(A, p, J, red, redll, xch); }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *sh_last = (atomicAdd(A.adone + p, fi + fj) + fi + fj == A.T.nblk[p]);
    __syncthreads();
    if (*sh_last) { __threadfence(); sym_phase_b(A, p, red, redll); }
    __syncthreads();
}

constexpr int NBUF = 3;

// The persistent kernel.  One CTA per SM; every CTA runs the loop
//     [B0] -> phases the previous blocks made this CTA responsible for (rare) -> block of slot s: wait for its bulk
//     copy, both integer products, plane sums -> fp64 -> [B1] -> partial sums to global -> [B2] -> thread 0: fence,
//     completion counters
// with everything that needs a round trip to L2 issued one block ahead of its use by thread 0: the ticket atomic
// and the availability mask are requested in block k and looked at in block k + 1 (the drawn block then gets its bulk
// copies, two blocks before it is processed); the completion counters bumped after block k are looked at after block
// k + 1.  So the critical path of a block is the shared-memory work plus three CTA barriers.
__global__ void __launch_bounds__(SYM_THREADS, 1)
sym_ice_kernel(SymArgs A) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* red_r = reinterpret_cast<double*>(smem_raw + NBUF * BUF_BYTES);   // [4][256]
    double* red_c = red_r + 4 * BLK;                                           // [4][256]
    __shared__ __align__(8) unsigned long long bar[NBUF];
    __shared__ double red[32];
    __shared__ long long redll[32];
    __shared__ int sh_item[NBUF];       // block held by each buffer (-1 = none)
    __shared__ int sh_flag[8];          // resolved completion of an earlier block: valid, p, I, J, fi, fj; [6] = last
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wi = warp >> 2, wj = warp & 3;
    const int g = lane >> 2, q = lane & 3;
    const double w_even = __longlong_as_double((long long)(1023 + 8 * (7 - 2 * q)) << 52);
    const double w_odd = __longlong_as_double((long long)(1023 + 8 * (6 - 2 * q)) << 52);
    uint32_t phasebits = 0u;            // expected parity of each buffer's mbarrier
    // thread 0 only: the outstanding ticket, the last availability mask seen, the completion counters not yet looked at
    int tk_n = 0, tk_r = 0, tk_t0 = 0, tk_t1 = 0;           // up to two tickets (of the same chromosome) in flight
    unsigned long long mask_seen = 0ull;
    int pend_valid = 0, pend_p = 0, pend_I = 0, pend_J = 0, pend_oldI = 0, pend_oldJ = 0;
    long long st_items = 0, st_wait = 0, st_phase = 0, st_copy = 0, st_nwait = 0, st_fail = 0, st_nphase = 0;
    const long long st_t0 = clock64();
    if (tid == 0) {
        for (int x = 0; x < NBUF; ++x) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[x])), "r"(1) : "memory");
            sh_item[x] = -1;
        }
        sh_flag[0] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto resolve_pending = [&]() {      // thread 0: did the block before last complete a block index?
        sh_flag[0] = 0;
        if (pend_valid) {
            const int need = A.T.nblk[pend_p] + 1;
            const int fi = pend_I == pend_J ? (pend_oldI + 2 == need) : (pend_oldI + 1 == need);
            const int fj = pend_I == pend_J ? 0 : (pend_oldJ + 1 == need);
            if (fi || fj) { sh_flag[0] = 1; sh_flag[1] = pend_p; sh_flag[2] = pend_I; sh_flag[3] = pend_J; sh_flag[4] = fi; sh_flag[5] = fj; }
            pend_valid = 0;
        }
    };
    for (int it = 0;; ++it) {
        const int s = it % NBUF;
        __syncthreads();                                                          // [B0]
        if (sh_flag[0]) {
            const long long t = clock64();
            sym_phases(A, sh_flag[1], sh_flag[2], sh_flag[3], sh_flag[4], sh_flag[5], red, redll, red_r, &sh_flag[6]);
            st_phase += clock64() - t; ++st_nphase;
        }
        int cur = sh_item[s];
        if (cur < 0) {
            // Nothing in hand (then no buffer holds a block: they fill in order).  Settle what is outstanding, then wait for
            // work or for the end.
            __syncthreads();
            if (tid == 0) resolve_pending();
            __syncthreads();
            if (sh_flag[0]) {
                const long long t = clock64();
                sym_phases(A, sh_flag[1], sh_flag[2], sh_flag[3], sh_flag[4], sh_flag[5], red, redll, red_r, &sh_flag[6]);
                st_phase += clock64() - t; ++st_nphase;
                if (tid == 0) sh_flag[0] = 0;
            }
            if (tid == 0) {
                int got = -1;
                const long long t_start = clock64();
                unsigned long long skip = 0ull;     // chromosomes whose bit is still set but whose tickets just ran out
                while (got < 0) {
                    const unsigned long long m = ld_volatile_u64(A.avail) & ~skip;
                    mask_seen = m;
                    if (m != 0ull) {
                        const int r = __ffsll((long long)m) - 1;
                        got = sym_ticket_item(A, r, atomicAdd(A.tick + A.T.prio[r], 1));
                        if (got < 0) { skip |= 1ull << r; ++st_fail; }
                        continue;
                    }
                    skip = 0ull;
                    if (ld_volatile_i32(A.n_active) <= 0 || ld_volatile_i32(A.abort_flag) != 0) break;
                    if (clock64() - t_start > A.spin_limit) { atomicExch(A.abort_flag, 1); break; }
                    __nanosleep(100);
                }
                sh_item[s] = got;
                if (got >= 0) sym_issue_copy(A, got, smem_raw + s * BUF_BYTES, &bar[s]);
                st_wait += clock64() - t_start; ++st_nwait;
            }
            __syncthreads();
            cur = sh_item[s];
            if (cur < 0) break;
        }
        if (tid == 0) {
            // request a ticket for every buffer that will be empty after this block (looked at after the block) and refresh
            // the mask
            const int empty = min(2, 1 + (sh_item[(s + 1) % NBUF] < 0) + (sh_item[(s + 2) % NBUF] < 0));
            if (mask_seen != 0ull) {
                tk_r = __ffsll((long long)mask_seen) - 1;
                if (empty >= 1) tk_t0 = atomicAdd(A.tick + A.T.prio[tk_r], 1);
                if (empty >= 2) tk_t1 = atomicAdd(A.tick + A.T.prio[tk_r], 1);
                tk_n = empty;
            }
            mask_seen = ld_volatile_u64(A.avail);
        }
        const int4 d = __ldg(A.items + cur);
        const int p = d.z, I = d.w & 0xffff, J = d.w >> 16;
        const int nblk = A.T.nblk[p];
        const int64_t npad = (int64_t)nblk * BLK;
        const int dexp = __ldcg(A.dig_exp + p);                // used after the products: latency hidden
        ++st_items;
        const long long st_c0 = clock64();
        {                               // wait for the block (and its bias planes) to land
            const uint32_t par = (phasebits >> s) & 1u;
            uint32_t done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(par) : "memory");
            }
            phasebits ^= 1u << s;
        }
        st_copy += clock64() - st_c0;
        const unsigned char* bufp = smem_raw + s * BUF_BYTES;
        const uint32_t base = smem_u32(bufp);
        uint2 bf1[2], bf2[2];           // bias planes of this warp's columns (product 1) and rows (product 2)
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) bf1[kt] = *reinterpret_cast<const uint2*>(bufp + BLK_BYTES + (2 * wj + kt) * 256 + 8 * lane);
#pragma unroll
        for (int sp = 0; sp < 2; ++sp) bf2[sp] = *reinterpret_cast<const uint2*>(bufp + BLK_BYTES + DIG_BYTES + (2 * wi + sp) * 256 + 8 * lane);
        int c1[4][4], c2[2][2][4];
#pragma unroll
        for (int x = 0; x < 4; ++x) { c1[x][0] = 0; c1[x][1] = 0; c1[x][2] = 0; c1[x][3] = 0; }
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
            for (int h = 0; h < 2; ++h) { c2[kt][h][0] = 0; c2[kt][h][1] = 0; c2[kt][h][2] = 0; c2[kt][h][3] = 0; }
        const int mj = lane >> 3, mi = lane & 7;         // ldmatrix.trans: lane supplies row mi of matrix mj
#pragma unroll
        for (int sp = 0; sp < 2; ++sp) {
#pragma unroll
            for (int kt = 0; kt < 2; ++kt) {
                const uint32_t tileA = base + (uint32_t)(((4 * wi + 2 * sp) * 8 + 2 * wj + kt) * 512);
                const uint32_t tileB = tileA + 8 * 512;                 // the strip below, same k-tile
                uint32_t fa[4], fb[4];
                ldsm_x4(fa, tileA + 16 * lane);
                ldsm_x4(fb, tileB + 16 * lane);
                mma_u8(c1[2 * sp], fa, bf1[kt]);
                mma_u8(c1[2 * sp + 1], fb, bf1[kt]);
#pragma unroll
                for (int h = 0; h < 2; ++h) {           // 16-column half h: matrices A(0,h) A(1,h) B(0,h) B(1,h)
                    uint32_t r[4], f[4];
                    ldsm_x4_trans(r, (mj < 2 ? tileA : tileB) + (uint32_t)(128 * ((mj & 1) + 2 * h) + 16 * mi));
                    f[0] = __byte_perm(r[0], r[1], 0x6420);     // column 2g,   rows {2q, 2q+1, 8+2q, 9+2q} of strip a
                    f[1] = __byte_perm(r[0], r[1], 0x7531);     // column 2g+1
                    f[2] = __byte_perm(r[2], r[3], 0x6420);     // ... of strip b
                    f[3] = __byte_perm(r[2], r[3], 0x7531);
                    mma_u8(c2[kt][h], f, bf2[sp]);
                }
            }
        }
        // plane sums -> fp64, reduced over the 4 lanes of a group; cross-warp sums through shared memory
        const double scale = __longlong_as_double((long long)(1023 + dexp - 64) << 52);   // 2^(E - 64)
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            double va = ((double)c1[x][0] * w_even + (double)c1[x][1] * w_odd) * scale;
            double vb = ((double)c1[x][2] * w_even + (double)c1[x][3] * w_odd) * scale;
            va += __shfl_xor_sync(0xffffffffu, va, 1); va += __shfl_xor_sync(0xffffffffu, va, 2);
            vb += __shfl_xor_sync(0xffffffffu, vb, 1); vb += __shfl_xor_sync(0xffffffffu, vb, 2);
            if (q == 0) {
                const int r0 = (4 * wi + x) * 16 + g;
                red_r[wj * BLK + r0] = va;
                red_r[wj * BLK + r0 + 8] = vb;
            }
        }
#pragma unroll
        for (int kt = 0; kt < 2; ++kt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double va = ((double)c2[kt][h][0] * w_even + (double)c2[kt][h][1] * w_odd) * scale;   // column 2g
                double vb = ((double)c2[kt][h][2] * w_even + (double)c2[kt][h][3] * w_odd) * scale;   // column 2g+1
                va += __shfl_xor_sync(0xffffffffu, va, 1); va += __shfl_xor_sync(0xffffffffu, va, 2);
                vb += __shfl_xor_sync(0xffffffffu, vb, 1); vb += __shfl_xor_sync(0xffffffffu, vb, 2);
                if (q == 0) {
                    const int cc = (2 * wj + kt) * 32 + 16 * h + 2 * g;
                    red_c[wi * BLK + cc] = va;
                    red_c[wi * BLK + cc + 1] = vb;
                }
            }
        __syncthreads();                                                          // [B1]
        double* part = A.part + A.T.part_off[p];
        if (tid < 256) {
            const double t = ((red_r[tid] + red_r[BLK + tid]) + red_r[2 * BLK + tid]) + red_r[3 * BLK + tid];
            part[(int64_t)J * npad + I * BLK + tid] = t;                      // rows of block I, partial over the columns of block J
        } else {
            const int cdx = tid - 256;
            const double t = ((red_c[cdx] + red_c[BLK + cdx]) + red_c[2 * BLK + cdx]) + red_c[3 * BLK + cdx];
            part[(int64_t)(nblk + I) * npad + J * BLK + cdx] = t;             // columns of block J, partial over the rows of block I
        }
        __syncthreads();                                                          // [B2]
        // thread 0: make the CTA's partial sums visible, look at the counters the PREVIOUS block bumped (no wait: they were
        // bumped a block ago), bump this block's.  Block (I, J) feeds the bins of index I (row partials) and J (column
        // partials); an index has all its nblk + 1 partial planes when its counter reaches nblk + 1.
        if (tid == 0) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            resolve_pending();
            pend_p = p; pend_I = I; pend_J = J; pend_valid = 1;
            int* bd = A.blkdone + A.T.blk_off[p];
            if (I == J) pend_oldI = atomicAdd(bd + I, 2);
            else { pend_oldI = atomicAdd(bd + I, 1); pend_oldJ = atomicAdd(bd + J, 1); }
            sh_item[s] = -1;
            // the tickets requested at the top of this block: their blocks go into the earliest empty buffers in processing
            // order (s+1, s+2, then s again) and their bulk copies start now, one to three blocks before they are needed
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                if (k < tk_n) {
                    const int got = sym_ticket_item(A, tk_r, k == 0 ? tk_t0 : tk_t1);
                    if (got >= 0) {
                        const int x = sh_item[(s + 1) % NBUF] < 0 ? (s + 1) % NBUF : (sh_item[(s + 2) % NBUF] < 0 ? (s + 2) % NBUF : s);
                        sh_item[x] = got;
                        sym_issue_copy(A, got, smem_raw + x * BUF_BYTES, &bar[x]);
                    } else { mask_seen &= ~(1ull << tk_r); ++st_fail; }      // its tickets ran out: try the next chromosome
                }
            }
            tk_n = 0;
        }
    }
    if (tid == 0 && A.stats != nullptr) {
        long long* o = A.stats + 8 * (long long)blockIdx.x;
        o[0] = st_items; o[1] = st_wait; o[2] = st_phase; o[3] = st_copy; o[4] = clock64() - st_t0; o[5] = st_nwait; o[6] = st_fail; o[7] = st_nphase;
    }
}

// user bias (concatenated bins) -> padded layout + both digit layouts + exponent (one CTA per chromosome)
__global__ void __launch_bounds__(SYM_THREADS)
sym_init_kernel(SymArgs A, const double* __restrict__ bias_in) {
    __shared__ double red[32];
    const int p = blockIdx.x;
    const int n = A.T.n[p], npad = A.T.nblk[p] * BLK;
    const int64_t lo = A.T.pad_off[p], b0 = A.T.bin_off[p];
    double mx = 0.0;
    for (int j = threadIdx.x; j < npad; j += blockDim.x) {
        const double v = j < n ? bias_in[b0 + j] : 0.0;
        A.bias[lo + j] = v;
        mx = fmax(mx, v);
    }
    mx = block_max_sym(mx, red);
    double pow2;
    const int E = digits_exp(mx, &pow2);
    if (threadIdx.x == 0) A.dig_exp[p] = E;
    __syncthreads();
    for (int j4 = 4 * threadIdx.x; j4 < npad; j4 += 4 * blockDim.x) {
        double bv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) bv[e] = A.bias[lo + j4 + e];
        sym_store_digits(bv, j4, pow2, A.dig1 + 8 * lo, A.dig2 + 8 * lo);
    }
}

// final weights: bias == 0 -> NaN; divide by sqrt(scale) when rescaling (cooler balance_cooler tail)
__global__ void __launch_bounds__(256)
sym_finalize_kernel(SymTables T, int nprob, const hc_ice_result* __restrict__ results, int rescale,
                    const double* __restrict__ padded, double* __restrict__ bias) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= T.bin_off[nprob]) return;
    int p = 0;
    while (p + 1 < nprob && T.bin_off[p + 1] <= g) ++p;
    const hc_ice_result r = results[p];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double b = padded[T.pad_off[p] + (g - T.bin_off[p])];
    if (isnan(r.scale)) b = nan;
    else {
        if (b == 0.0) b = nan;
        if (rescale) b = b / sqrt(r.scale);
    }
    bias[g] = b;
}

struct Scratch {
    cudaStream_t s;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t st) : s(st) {}
    cudaError_t alloc(void** p, size_t bytes) {
        const cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, s);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, s); }
};

}  // namespace

// Returns HC_OK / an error, or +1 when this path does not apply (a chromosome with more than 8192 bins, or nothing to
// balance): the caller then uses the full-matrix packed kernel of hc_ice.cu.
int hc_ice_dense_balance_sym(const int32_t* mats, const int64_t* mat_off, const int32_t* mat_n, const int32_t* mat_ld,
                             const int64_t* bin_off, int32_t nprob, const int32_t* h_mat_n, const hc_ice_params* P,
                             double* bias, hc_ice_result* results, hc_ice_run_info* h_info, cudaStream_t s) {
    (void)mat_n;
    if (nprob > 64) return 1;           // the availability mask is one 64-bit word
    for (int p = 0; p < nprob; ++p) if (h_mat_n[p] > MAX_BINS) return 1;
    int64_t nbins = 0;
    for (int p = 0; p < nprob; ++p) nbins += h_mat_n[p];
    std::vector<int32_t> h_ld(nprob);
    HC_CUDA(hc_read_small(h_ld.data(), mat_ld, sizeof(int32_t) * nprob, s));
    // every table is indexed by the caller's chromosome index; only the ticket scan walks `prio` (largest first: the
    // largest chromosome is the critical path of the dataflow)
    std::vector<int32_t> h_nblk(nprob), h_blkoff(nprob, 0), h_first(nprob + 1, 0), h_prio(nprob);
    std::vector<int64_t> h_pad(nprob, 0), h_part(nprob, 0);
    for (int p = 0; p < nprob; ++p) h_prio[p] = p;
    std::stable_sort(h_prio.begin(), h_prio.end(), [&](int a, int b) { return h_mat_n[a] > h_mat_n[b]; });
    std::vector<int4> h_items;
    int64_t npad = 0, part_n = 0, qbytes = 0;
    int32_t nblk_tot = 0;
    for (int p = 0; p < nprob; ++p) {
        const int n = h_mat_n[p], nb = (n + BLK - 1) / BLK;
        if (h_ld[p] < n || (h_ld[p] & 127) != 0) {
            hc_set_error("hc_ice_dense_balance: ld must be >= n and a multiple of 128 elements (matrix %d: n=%d ld=%d)", p, n, h_ld[p]);
            return HC_ERR_ARG;
        }
        h_nblk[p] = nb; h_blkoff[p] = nblk_tot; h_pad[p] = npad; h_part[p] = part_n;
        nblk_tot += nb; npad += (int64_t)nb * BLK; part_n += 2ll * nb * nb * BLK;
        h_first[p] = (int32_t)h_items.size();
        for (int I = 0; I < nb; ++I)
            for (int J = I; J < nb; ++J) {
                int4 d;
                d.x = (int)(uint32_t)(qbytes & 0xffffffffll); d.y = (int)(qbytes >> 32); d.z = p; d.w = I | (J << 16);
                h_items.push_back(d);
                qbytes += BLK_BYTES;
            }
    }
    h_first[nprob] = (int32_t)h_items.size();
    const int nitems = (int)h_items.size();
    if (nitems == 0) return 1;      // nothing to balance: the generic path produces the defined (NaN) result
    if (h_info) { h_info->launches = 0; h_info->loop_ms = 0.f; h_info->packed = 3; h_info->pack_ms = 0.f; h_info->overflow_cells = 0; h_info->stream_full_ms = 0.f; h_info->stream_full_launches = 0; }

    Scratch scratch(s);
    int32_t* d_i32 = nullptr;      // n | nblk | blk_off | item_first[nprob + 1] | prio | rank_of
    int64_t* d_i64 = nullptr;      // pad_off | part_off
    HC_CUDA(scratch.alloc((void**)&d_i32, sizeof(int32_t) * (6 * (size_t)nprob + 1)));
    HC_CUDA(scratch.alloc((void**)&d_i64, sizeof(int64_t) * 2 * (size_t)nprob));
    std::vector<int32_t> h_i32;
    h_i32.insert(h_i32.end(), h_mat_n, h_mat_n + nprob);
    h_i32.insert(h_i32.end(), h_nblk.begin(), h_nblk.end());
    h_i32.insert(h_i32.end(), h_blkoff.begin(), h_blkoff.end());
    h_i32.insert(h_i32.end(), h_first.begin(), h_first.end());
    h_i32.insert(h_i32.end(), h_prio.begin(), h_prio.end());
    {
        std::vector<int32_t> h_rank(nprob);
        for (int r = 0; r < nprob; ++r) h_rank[h_prio[r]] = r;
        h_i32.insert(h_i32.end(), h_rank.begin(), h_rank.end());
    }
    std::vector<int64_t> h_i64;
    h_i64.insert(h_i64.end(), h_pad.begin(), h_pad.end());
    h_i64.insert(h_i64.end(), h_part.begin(), h_part.end());
    HC_CUDA(cudaMemcpyAsync(d_i32, h_i32.data(), sizeof(int32_t) * h_i32.size(), cudaMemcpyHostToDevice, s));
    HC_CUDA(cudaMemcpyAsync(d_i64, h_i64.data(), sizeof(int64_t) * h_i64.size(), cudaMemcpyHostToDevice, s));
    int4* d_items = nullptr;
    HC_CUDA(scratch.alloc((void**)&d_items, sizeof(int4) * (size_t)nitems));
    HC_CUDA(cudaMemcpyAsync(d_items, h_items.data(), sizeof(int4) * (size_t)nitems, cudaMemcpyHostToDevice, s));
    SymTables T;
    T.n = d_i32; T.nblk = d_i32 + nprob; T.blk_off = d_i32 + 2 * nprob; T.item_first = d_i32 + 3 * nprob;
    T.prio = d_i32 + 4 * nprob + 1; T.rank_of = d_i32 + 5 * nprob + 1;
    T.pad_off = d_i64; T.part_off = d_i64 + nprob; T.bin_off = bin_off;

    cudaEvent_t evp0 = nullptr, evp1 = nullptr, ev0 = nullptr, ev1 = nullptr;
    struct EvGuard { cudaEvent_t* e[4]; ~EvGuard() { for (auto p : e) if (*p) cudaEventDestroy(*p); } } evg{{&evp0, &evp1, &ev0, &ev1}};
    if (h_info) { cudaEventCreate(&evp0); cudaEventCreate(&evp1); cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(evp0, s); }

    // ---- encoding -------------------------------------------------------------------------------------------
    uint8_t* d_q8 = nullptr;
    HC_CUDA(scratch.alloc((void**)&d_q8, (size_t)qbytes));
    int64_t* d_ovf_ptr = nullptr;      // [nbins + 1] | lo[nbins] hi[nbins] (int32)
    HC_CUDA(scratch.alloc((void**)&d_ovf_ptr, sizeof(int64_t) * ((size_t)nbins + 1) + sizeof(int32_t) * 2 * (size_t)nbins));
    int32_t* d_lo = reinterpret_cast<int32_t*>(d_ovf_ptr + nbins + 1);
    int32_t* d_hi = d_lo + nbins;
    HC_CUDA(cudaMemsetAsync(d_ovf_ptr, 0, sizeof(int64_t) * ((size_t)nbins + 1), s));
    HC_CUDA(cudaMemsetAsync(d_lo, 0x7f, sizeof(int32_t) * (size_t)nbins, s));
    HC_CUDA(cudaMemsetAsync(d_hi, 0xff, sizeof(int32_t) * (size_t)nbins, s));
    sym_pack_kernel<<<(unsigned)(((long long)nitems * 128 * 32 + 255) / 256), 256, 0, s>>>(
        mats, mat_off, mat_ld, T, d_items, nitems, P->ignore_diags, d_q8, reinterpret_cast<unsigned long long*>(d_ovf_ptr), d_lo, d_hi);
    HC_LAUNCH_CHECK();
    sym_scan_kernel<<<1, 1024, 0, s>>>(d_ovf_ptr, nbins);
    HC_LAUNCH_CHECK();
    long long novf = 0;
    HC_CUDA(hc_read_small(&novf, d_ovf_ptr + nbins, sizeof(long long), s));     // also: the host tables above are uploaded
    int32_t* d_ovf = nullptr;
    HC_CUDA(scratch.alloc((void**)&d_ovf, 2 * (size_t)std::max(novf, 1ll) * sizeof(int32_t)));
    if (novf > 0) {
        sym_ovf_fill_kernel<<<(unsigned)((nbins * 32 + 255) / 256), 256, 0, s>>>(mats, mat_off, mat_ld, T, nprob, P->ignore_diags,
                                                                                 d_ovf_ptr, d_lo, d_hi, d_ovf, d_ovf + novf);
        HC_LAUNCH_CHECK();
    }

    // ---- vectors, partial planes, bookkeeping ------------------------------------------------------------------
    double* d_vec = nullptr;           // bias[npad] | marg[npad]
    HC_CUDA(scratch.alloc((void**)&d_vec, sizeof(double) * 2 * (size_t)npad));
    HC_CUDA(cudaMemsetAsync(d_vec, 0, sizeof(double) * 2 * (size_t)npad, s));
    uint8_t* d_dig = nullptr;          // dig1[8 npad] | dig2[8 npad]
    HC_CUDA(scratch.alloc((void**)&d_dig, 16 * (size_t)npad));
    HC_CUDA(cudaMemsetAsync(d_dig, 0, 16 * (size_t)npad, s));
    double* d_part = nullptr;
    HC_CUDA(scratch.alloc((void**)&d_part, sizeof(double) * (size_t)part_n));
    HC_CUDA(cudaMemsetAsync(d_part, 0, sizeof(double) * (size_t)part_n, s));
    // int32 book: tick[nprob] | adone[nprob] | iters[nprob] | dig_exp[nprob] | n_active | abort | blkdone[nblk_tot]
    int32_t* d_book = nullptr;
    const size_t book_n = 4 * (size_t)nprob + 2 + (size_t)nblk_tot;
    HC_CUDA(scratch.alloc((void**)&d_book, sizeof(int32_t) * book_n));
    std::vector<int32_t> h_book(book_n, 0);
    std::vector<hc_ice_result> h_res(nprob);
    int active = 0;
    for (int p = 0; p < nprob; ++p) {
        h_res[p].scale = NAN; h_res[p].var = 0.0; h_res[p].iters = 0; h_res[p].converged = 1;     // stands for empty chromosomes
        if (h_mat_n[p] > 0) ++active;
    }
    h_book[4 * (size_t)nprob] = active;
    HC_CUDA(cudaMemcpyAsync(d_book, h_book.data(), sizeof(int32_t) * book_n, cudaMemcpyHostToDevice, s));
    HC_CUDA(cudaMemcpyAsync(results, h_res.data(), sizeof(hc_ice_result) * (size_t)nprob, cudaMemcpyHostToDevice, s));
    double* d_blk = nullptr;           // blk_sum[nblk_tot] | blk_cnt[nblk_tot] (int64) | availability mask (uint64)
    HC_CUDA(scratch.alloc((void**)&d_blk, 16 * (size_t)nblk_tot + 8));
    unsigned long long h_avail = 0ull;              // every chromosome with blocks starts with its tickets open
    for (int r = 0; r < nprob; ++r) if (h_first[h_prio[r] + 1] > h_first[h_prio[r]]) h_avail |= 1ull << r;
    HC_CUDA(cudaMemcpyAsync(reinterpret_cast<unsigned char*>(d_blk) + 16 * (size_t)nblk_tot, &h_avail, 8, cudaMemcpyHostToDevice, s));
    HC_CUDA(cudaStreamSynchronize(s));      // host staging vectors go out of use

    SymArgs A;
    A.T = T; A.nprob = nprob; A.q8 = d_q8; A.items = d_items;
    A.bias = d_vec; A.marg = d_vec + npad; A.dig1 = d_dig; A.dig2 = d_dig + 8 * (size_t)npad;
    A.tick = d_book; A.adone = d_book + nprob; A.iters = d_book + 2 * nprob; A.dig_exp = d_book + 3 * nprob;
    A.n_active = d_book + 4 * nprob; A.abort_flag = d_book + 4 * nprob + 1; A.blkdone = d_book + 4 * nprob + 2;
    A.spin_limit = 4000000000ll;        // ~2 s at 1.9 GHz: no chromosome's update takes anywhere near that
    A.stats = nullptr;
    const bool want_stats = getenv("HC_SYM_STATS") != nullptr && atoi(getenv("HC_SYM_STATS")) != 0;
    const int grid = std::min(hc_num_sms(), nitems);
    if (want_stats) {
        HC_CUDA(scratch.alloc((void**)&A.stats, sizeof(long long) * 8 * (size_t)grid));
        HC_CUDA(cudaMemsetAsync(A.stats, 0, sizeof(long long) * 8 * (size_t)grid, s));
    }
    A.part = d_part; A.blk_sum = d_blk; A.blk_cnt = reinterpret_cast<long long*>(d_blk + nblk_tot);
    A.avail = reinterpret_cast<unsigned long long*>(d_blk + 2 * (size_t)nblk_tot);
    A.ovf_ptr = d_ovf_ptr; A.ovf_col = d_ovf; A.ovf_val = d_ovf + novf;
    A.results = results; A.tol = P->tol; A.max_iters = P->max_iters;
    sym_init_kernel<<<nprob, SYM_THREADS, 0, s>>>(A, bias);
    HC_LAUNCH_CHECK();
    if (h_info && evp1) cudaEventRecord(evp1, s);

    const size_t smem = NBUF * (size_t)BUF_BYTES + 2 * 4 * BLK * sizeof(double);
    HC_CUDA(cudaFuncSetAttribute(sym_ice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (h_info && ev0) cudaEventRecord(ev0, s);
    sym_ice_kernel<<<grid, SYM_THREADS, smem, s>>>(A);
    HC_LAUNCH_CHECK();
    if (h_info && ev1) cudaEventRecord(ev1, s);
    sym_finalize_kernel<<<(unsigned)((nbins + 255) / 256), 256, 0, s>>>(T, nprob, results, P->rescale_marginals, d_vec, bias);
    HC_LAUNCH_CHECK();
    int32_t h_abort = 0;
    const cudaError_t e = hc_read_small(&h_abort, A.abort_flag, sizeof(int32_t), s);
    if (e != cudaSuccess) { hc_set_error("hc_ice_dense_balance (symmetric): %s", cudaGetErrorString(e)); return HC_ERR_CUDA; }
    if (h_abort) { hc_set_error("hc_ice_dense_balance (symmetric): the dataflow kernel stalled (a CTA waited > 2 s for work)"); return HC_ERR_CUDA; }
    if (want_stats) {
        std::vector<long long> h_st(8 * (size_t)grid);
        HC_CUDA(cudaMemcpyAsync(h_st.data(), A.stats, sizeof(long long) * h_st.size(), cudaMemcpyDeviceToHost, s));
        HC_CUDA(cudaStreamSynchronize(s));
        double sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int b = 0; b < grid; ++b) for (int k = 0; k < 8; ++k) sum[k] += (double)h_st[8 * (size_t)b + k];
        fprintf(stderr, "[sym_ice_kernel] %d CTAs; per CTA: %.0f blocks, %.0f waits for work, %.0f failed tickets, %.0f phase runs; "
                "cycles: total %.0f = waiting for work %.1f %% + phases %.1f %% + waiting for a bulk copy %.1f %% + rest %.1f %%\n",
                grid, sum[0] / grid, sum[5] / grid, sum[6] / grid, sum[7] / grid, sum[4] / grid, 100 * sum[1] / sum[4], 100 * sum[2] / sum[4],
                100 * sum[3] / sum[4], 100 * (sum[4] - sum[1] - sum[2] - sum[3]) / sum[4]);
    }
    if (h_info) {
        h_info->launches = 5 + (novf > 0 ? 1 : 0);
        h_info->packed = 3;
        h_info->overflow_cells = novf;
        if (ev0) cudaEventElapsedTime(&h_info->loop_ms, ev0, ev1);
        if (evp0) cudaEventElapsedTime(&h_info->pack_ms, evp0, evp1);
        h_info->stream_full_ms = h_info->loop_ms;      // one kernel runs the whole loop
        h_info->stream_full_launches = 1;
    }
    return HC_OK;
}
