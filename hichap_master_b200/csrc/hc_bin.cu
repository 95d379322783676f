// (a) Valid-pair binning into dense int32 tiles + dense->sparse marshalling.
// Replaces the interpreted per-line loops of HiCHap/matrixBuilding.py:568-603 (traditional),
// :817-852 (traditional-in-allelic), :1131-1161 / :1169-1199 (M_M / P_P 'Both'),
// :1207-1243 (M_P / P_M), :1284-1301 / :1398-1415 (one-sided R1/R2), and the np.triu/np.nonzero
// marshalling at :457-524.
//
// Roofline: HBM-bound streaming read of the columnar pairs (16 B/pair, +1 B mark) with 128-bit
// loads; the += 1 updates are 32-bit RED atomics resolved in L2 (not HBM traffic).
#include "hc_common.cuh"

namespace {

constexpr int BIN_THREADS = 256;
constexpr int PAIRS_PER_THREAD = 4;  // one int4 per column per thread

struct PairCols {
    const int32_t* c1; const int32_t* p1; const int32_t* c2; const int32_t* p2; const uint8_t* mark;
};

__device__ __forceinline__ bool mode_accepts(int mode, int mk) {
    // HC_BIN_SYM_ALL: everything; SYM_BOTH: mark==Both only; ONESIDED: mark!=Both only
    return mode == HC_BIN_SYM_ALL || (mode == HC_BIN_SYM_BOTH ? mk == 0 : mk != 0);
}

template <bool WHOLE>
__device__ __forceinline__ void apply_pair(int c1, int p1, int c2, int p2, int mk, uint32_t res, int mode,
                                           int32_t* __restrict__ mats, const int64_t* __restrict__ t0,
                                           const int64_t* __restrict__ t1, const int32_t* __restrict__ mat_n,
                                           const int32_t* __restrict__ mat_ld, int nchrom, int64_t whole_ld,
                                           int32_t whole_n, unsigned long long* oob) {
    if (c1 < 0 || c2 < 0 || c1 >= nchrom || c2 >= nchrom) return;   // filtered chromosome
    if (!mode_accepts(mode, mk)) return;
    if ((!WHOLE || mode == HC_BIN_ONESIDED) && c1 != c2) return;     // cis only
    if (p1 < 0 || p2 < 0) { if (oob) atomicAdd(oob, 1ull); return; }
    int64_t b1 = (uint32_t)p1 / res, b2 = (uint32_t)p2 / res;
    int32_t* M;
    int64_t ld, n;
    if (WHOLE) {
        b1 += t0[c1]; b2 += t1[c2];
        M = mats; ld = whole_ld; n = whole_n;
    } else {
        M = mats + t0[c1]; ld = mat_ld[c1]; n = mat_n[c1];
    }
    if (b1 >= n || b2 >= n) { if (oob) atomicAdd(oob, 1ull); return; }
    if (mode == HC_BIN_ONESIDED) {
        // R1: row = first mate; anything else (R2): row = second mate  (matrixBuilding.py:1298-1301)
        if (mk == 1) atomicAdd(&M[b1 * ld + b2], 1); else atomicAdd(&M[b2 * ld + b1], 1);
    } else {
        atomicAdd(&M[b1 * ld + b2], 1);
        if (b1 != b2) atomicAdd(&M[b2 * ld + b1], 1);
    }
}

template <bool WHOLE>
__global__ void __launch_bounds__(BIN_THREADS)
bin_pairs_kernel(PairCols in, int64_t npairs, uint32_t res, int mode, int32_t* __restrict__ mats,
                 const int64_t* __restrict__ t0, const int64_t* __restrict__ t1,
                 const int32_t* __restrict__ mat_n, const int32_t* __restrict__ mat_ld, int nchrom,
                 int64_t whole_ld, int32_t whole_n, unsigned long long* oob) {
    const int64_t nvec = npairs / PAIRS_PER_THREAD;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const int4 a = ld_stream_v4(in.c1 + 4 * v), b = ld_stream_v4(in.p1 + 4 * v);
        const int4 c = ld_stream_v4(in.c2 + 4 * v), d = ld_stream_v4(in.p2 + 4 * v);
        uint32_t mk4 = 0;
        if (in.mark) mk4 = *reinterpret_cast<const uint32_t*>(in.mark + 4 * v);
        apply_pair<WHOLE>(a.x, b.x, c.x, d.x, mk4 & 255, res, mode, mats, t0, t1, mat_n, mat_ld, nchrom, whole_ld, whole_n, oob);
        apply_pair<WHOLE>(a.y, b.y, c.y, d.y, (mk4 >> 8) & 255, res, mode, mats, t0, t1, mat_n, mat_ld, nchrom, whole_ld, whole_n, oob);
        apply_pair<WHOLE>(a.z, b.z, c.z, d.z, (mk4 >> 16) & 255, res, mode, mats, t0, t1, mat_n, mat_ld, nchrom, whole_ld, whole_n, oob);
        apply_pair<WHOLE>(a.w, b.w, c.w, d.w, (mk4 >> 24) & 255, res, mode, mats, t0, t1, mat_n, mat_ld, nchrom, whole_ld, whole_n, oob);
    }
    // ragged tail (< 4 pairs)
    if (blockIdx.x == 0 && threadIdx.x < (int)(npairs - nvec * PAIRS_PER_THREAD)) {
        const int64_t i = nvec * PAIRS_PER_THREAD + threadIdx.x;
        apply_pair<WHOLE>(in.c1[i], in.p1[i], in.c2[i], in.p2[i], in.mark ? in.mark[i] : 0, res, mode, mats, t0, t1,
                          mat_n, mat_ld, nchrom, whole_ld, whole_n, oob);
    }
}

int bin_grid(int64_t npairs) {
    int64_t nvec = (npairs + PAIRS_PER_THREAD - 1) / PAIRS_PER_THREAD;
    int64_t blocks = (nvec + BIN_THREADS - 1) / BIN_THREADS;
    int64_t cap = (int64_t)hc_num_sms() * 8;  // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : (int)blocks;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- dense -> sparse -------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
row_nonzero_count_kernel(const T* __restrict__ M, int64_t ld, int nrows, int ncols, int triu,
                         int64_t* __restrict__ row_cnt) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    const T* row = M + (int64_t)warp * ld;
    int cnt = 0;
    for (int j = (triu ? warp : 0) + lane; j < ncols; j += 32) cnt += (row[j] != T(0));
    cnt = warp_sum_i(cnt);
    if (lane == 0) row_cnt[warp] = cnt;
}

// single-CTA exclusive scan of row_cnt[0..n) in place, total to row_cnt[n]
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(int64_t* __restrict__ v, int n) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        long long x = i < n ? v[i] : 0, incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        long long woff = 0;
        for (int w = 0; w < wid; ++w) woff += warp_tot[w];
        const long long carry = carry_s;
        if (i < n) v[i] = carry + woff + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) v[n] = carry_s;
}

template <typename T>
__global__ void __launch_bounds__(256)
row_nonzero_extract_kernel(const T* __restrict__ M, int64_t ld, int nrows, int ncols, int triu,
                           const int64_t* __restrict__ row_ptr, int32_t* __restrict__ bin1,
                           int32_t* __restrict__ bin2, T* __restrict__ val) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    const T* row = M + (int64_t)warp * ld;
    int64_t out = row_ptr[warp];
    const int j0 = triu ? warp : 0;
    for (int base = j0; base < ncols; base += 32) {   // ascending columns: row-major order
        const int j = base + lane;
        T x = j < ncols ? row[j] : T(0);
        const unsigned m = __ballot_sync(0xffffffffu, x != T(0));
        if (x != T(0)) {
            const int64_t o = out + __popc(m & ((1u << lane) - 1u));
            bin1[o] = warp; bin2[o] = j; val[o] = x;
        }
        out += __popc(m);
    }
}

// ---- whole batch at once: upper-triangular records in the reference's 24-byte layout ------
struct Rec24 { long long bin1; long long bin2; double IF; };   // matrixBuilding.py:460-461 S_dtype

__device__ __forceinline__ int batch_problem(const int64_t* __restrict__ bin_off, int nprob, int64_t g) {
    int lo = 0, hi = nprob - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (bin_off[mid] <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
batch_triu_count_kernel(const int32_t* __restrict__ mats, const int64_t* __restrict__ mat_off,
                        const int32_t* __restrict__ mat_n, const int32_t* __restrict__ mat_ld,
                        const int64_t* __restrict__ bin_off, int nprob, int64_t* __restrict__ row_cnt) {
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per (global) row
    const int lane = threadIdx.x & 31;
    if (g >= bin_off[nprob]) return;
    const int p = batch_problem(bin_off, nprob, g);
    const int r = (int)(g - bin_off[p]), n = mat_n[p];
    const int32_t* row = mats + mat_off[p] + (int64_t)r * mat_ld[p];
    int cnt = 0;
    for (int j = r + lane; j < n; j += 32) cnt += (row[j] != 0);
    cnt = warp_sum_i(cnt);
    if (lane == 0) row_cnt[g] = cnt;
}

__global__ void __launch_bounds__(256)
batch_triu_records_kernel(const int32_t* __restrict__ mats, const int64_t* __restrict__ mat_off,
                          const int32_t* __restrict__ mat_n, const int32_t* __restrict__ mat_ld,
                          const int64_t* __restrict__ bin_off, int nprob, const int64_t* __restrict__ row_ptr,
                          Rec24* __restrict__ rec) {
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= bin_off[nprob]) return;
    const int p = batch_problem(bin_off, nprob, g);
    const int r = (int)(g - bin_off[p]), n = mat_n[p];
    const int32_t* row = mats + mat_off[p] + (int64_t)r * mat_ld[p];
    int64_t out = row_ptr[g];
    for (int base = r; base < n; base += 32) {
        const int j = base + lane;
        const int x = j < n ? row[j] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, x != 0);
        if (x != 0) {
            Rec24 v; v.bin1 = r; v.bin2 = j; v.IF = (double)x;
            rec[out + __popc(m & ((1u << lane) - 1u))] = v;
        }
        out += __popc(m);
    }
}

// multi-CTA exclusive scan is not needed: bins are O(1e5); one CTA with a 64-bit carry
}  // namespace

extern "C" int hc_dense_batch_triu_count(const int32_t* mats, const int64_t* mat_off, const int32_t* mat_n,
                                         const int32_t* mat_ld, const int64_t* bin_off, int32_t nprob,
                                         int64_t nbins, int64_t* row_ptr, void* stream) {
    HC_REQUIRE(nprob > 0 && nbins >= 0, "nprob>0, nbins>=0");
    cudaStream_t s = (cudaStream_t)stream;
    if (nbins > 0) {
        const int64_t blocks = (nbins * 32 + 255) / 256;
        batch_triu_count_kernel<<<(unsigned)blocks, 256, 0, s>>>(mats, mat_off, mat_n, mat_ld, bin_off, nprob, row_ptr);
        HC_LAUNCH_CHECK();
    }
    HC_REQUIRE(nbins < (1ll << 31), "too many bins for the single-CTA scan");
    exclusive_scan_kernel<<<1, 1024, 0, s>>>(row_ptr, (int)nbins);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

extern "C" int hc_dense_batch_triu_records(const int32_t* mats, const int64_t* mat_off, const int32_t* mat_n,
                                           const int32_t* mat_ld, const int64_t* bin_off, int32_t nprob,
                                           int64_t nbins, const int64_t* row_ptr, void* records, void* stream) {
    HC_REQUIRE(nprob > 0 && nbins >= 0, "nprob>0, nbins>=0");
    if (nbins == 0) return HC_OK;
    const int64_t blocks = (nbins * 32 + 255) / 256;
    batch_triu_records_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        mats, mat_off, mat_n, mat_ld, bin_off, nprob, row_ptr, reinterpret_cast<Rec24*>(records));
    HC_LAUNCH_CHECK();
    return HC_OK;
}

extern "C" int hc_bin_pairs_local(const int32_t* c1, const int32_t* p1, const int32_t* c2, const int32_t* p2,
                                  const uint8_t* mark, int64_t npairs, int32_t res, int32_t mode, int32_t* mats,
                                  const int64_t* mat_off, const int32_t* mat_n, const int32_t* mat_ld,
                                  int32_t nchrom, unsigned long long* oob, void* stream) {
    HC_REQUIRE(npairs >= 0 && res > 0 && nchrom > 0, "npairs>=0, res>0, nchrom>0");
    HC_REQUIRE(mode >= HC_BIN_SYM_ALL && mode <= HC_BIN_ONESIDED, "mode");
    HC_REQUIRE(mode == HC_BIN_SYM_ALL || mark != nullptr, "mark column required for this mode");
    if (npairs == 0) return HC_OK;
    HC_REQUIRE(aligned16(c1) && aligned16(p1) && aligned16(c2) && aligned16(p2), "pair columns must be 16-byte aligned");
    HC_REQUIRE(mark == nullptr || (reinterpret_cast<uintptr_t>(mark) & 3u) == 0, "mark must be 4-byte aligned");
    PairCols in{c1, p1, c2, p2, mark};
    bin_pairs_kernel<false><<<bin_grid(npairs), BIN_THREADS, 0, (cudaStream_t)stream>>>(
        in, npairs, (uint32_t)res, mode, mats, mat_off, nullptr, mat_n, mat_ld, nchrom, 0, 0, oob);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

extern "C" int hc_bin_pairs_whole(const int32_t* c1, const int32_t* p1, const int32_t* c2, const int32_t* p2,
                                  const uint8_t* mark, int64_t npairs, int32_t res, int32_t mode,
                                  const int64_t* start1, const int64_t* start2, int32_t nchrom, int32_t* M,
                                  int32_t total, int64_t ld, unsigned long long* oob, void* stream) {
    HC_REQUIRE(npairs >= 0 && res > 0 && nchrom > 0 && total > 0 && ld >= total, "sizes");
    HC_REQUIRE(mode >= HC_BIN_SYM_ALL && mode <= HC_BIN_ONESIDED, "mode");
    HC_REQUIRE(mode == HC_BIN_SYM_ALL || mark != nullptr, "mark column required for this mode");
    if (npairs == 0) return HC_OK;
    HC_REQUIRE(aligned16(c1) && aligned16(p1) && aligned16(c2) && aligned16(p2), "pair columns must be 16-byte aligned");
    HC_REQUIRE(mark == nullptr || (reinterpret_cast<uintptr_t>(mark) & 3u) == 0, "mark must be 4-byte aligned");
    PairCols in{c1, p1, c2, p2, mark};
    bin_pairs_kernel<true><<<bin_grid(npairs), BIN_THREADS, 0, (cudaStream_t)stream>>>(
        in, npairs, (uint32_t)res, mode, M, start1, start2, nullptr, nullptr, nchrom, ld, total, oob);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

extern "C" int hc_dense_nonzero_count(const void* M, int64_t ld, int32_t nrows, int32_t ncols, int32_t triu,
                                      int32_t is_f64, int64_t* row_ptr, void* stream) {
    HC_REQUIRE(nrows >= 0 && ncols >= 0 && ld >= ncols, "shape");
    cudaStream_t s = (cudaStream_t)stream;
    if (nrows > 0) {
        const int blocks = (nrows * 32 + 255) / 256;
        if (is_f64) row_nonzero_count_kernel<double><<<blocks, 256, 0, s>>>((const double*)M, ld, nrows, ncols, triu, row_ptr);
        else row_nonzero_count_kernel<int32_t><<<blocks, 256, 0, s>>>((const int32_t*)M, ld, nrows, ncols, triu, row_ptr);
        HC_LAUNCH_CHECK();
    }
    exclusive_scan_kernel<<<1, 1024, 0, s>>>(row_ptr, nrows);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

extern "C" int hc_dense_nonzero_extract(const void* M, int64_t ld, int32_t nrows, int32_t ncols, int32_t triu,
                                        int32_t is_f64, const int64_t* row_ptr, int32_t* bin1, int32_t* bin2,
                                        void* val, void* stream) {
    HC_REQUIRE(nrows >= 0 && ncols >= 0 && ld >= ncols, "shape");
    if (nrows == 0) return HC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = (nrows * 32 + 255) / 256;
    if (is_f64) row_nonzero_extract_kernel<double><<<blocks, 256, 0, s>>>((const double*)M, ld, nrows, ncols, triu, row_ptr, bin1, bin2, (double*)val);
    else row_nonzero_extract_kernel<int32_t><<<blocks, 256, 0, s>>>((const int32_t*)M, ld, nrows, ncols, triu, row_ptr, bin1, bin2, (int32_t*)val);
    HC_LAUNCH_CHECK();
    return HC_OK;
}
