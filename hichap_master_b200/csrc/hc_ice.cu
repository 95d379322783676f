// (b) ICE balancing on dense symmetric int32 tiles  ==  `cooler balance --ignore-diags K [--cis-only]`
// (HiCHap call sites: matrixBuilding.py:708, :713, :1537, :1542, :1761, :1766; algorithm restated
// in oracle/cooler_ice.py from cooler.balance.balance_cooler).
//
// One launch == one ICE iteration for EVERY problem (chromosome) of the batch:
//   prologue (per CTA, redundantly, from L2): mean / variance of the previous marginals over
//       the non-zero bins, bias update  b <- b / (marg / mean), convergence test  var < tol,
//       new bias segment staged in shared memory;
//   body: streaming pass over the CTA's rows of the int32 matrix with 128-bit loads,
//       fp64 dot product against the staged bias, warp-shuffle reduction, one store per row.
// Nothing goes back to the host inside the loop except a done-counter poll.
//
// Roofline: HBM-bound, 4*N^2 algorithmic bytes per iteration per problem.
#include <vector>
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include "hc_common.cuh"
#include "hc_select.cuh"

namespace {

__device__ __forceinline__ double band_weight(int j, int r, int kd) {
    // pixel weight under cooler's _zero_diags + _marginalize on upper-triangular pixels:
    // |i-j| < ignore_diags -> dropped; a kept diagonal pixel (ignore_diags == 0) counts twice
    const int d = j - r;
    if (d == 0) return kd == 0 ? 2.0 : 0.0;
    return (d < kd && d > -kd) ? 0.0 : 1.0;
}

// ---------------------------------------------------------------------------------------
// filter marginals: per bin, count and sum of kept pixels
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ice_dense_marginals_kernel(const int32_t* __restrict__ mats, const int64_t* __restrict__ mat_off,
                           const int32_t* __restrict__ mat_n, const int32_t* __restrict__ mat_ld,
                           const int64_t* __restrict__ bin_off, int nprob, int kd,
                           double* __restrict__ nnz_marg, double* __restrict__ marg) {
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // global bin == one warp
    const int lane = threadIdx.x & 31;
    if (g >= bin_off[nprob]) return;
    int p = 0;
    while (p + 1 < nprob && bin_off[p + 1] <= g) ++p;
    const int r = (int)(g - bin_off[p]), n = mat_n[p];
    const int64_t ld = mat_ld[p];
    const int32_t* row = mats + mat_off[p] + (int64_t)r * ld;
    const int nvec = (int)(ld >> 2);
    long long s = 0;
    int c = 0;
    for (int v = lane; v < nvec; v += 32) {
        const int4 a = ld_stream_v4(row + 4 * v);
        const int j = 4 * v;
        const int x[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (j + e < n) {
                const double w = band_weight(j + e, r, kd);
                s += (long long)(w * x[e]);
                c += (x[e] != 0) ? (int)w : 0;
            }
        }
    }
    s = warp_sum_ll(s);
    c = warp_sum_i(c);
    if (lane == 0) { marg[g] = (double)s; nnz_marg[g] = (double)c; }
}

// ---------------------------------------------------------------------------------------
// bin filters
// ---------------------------------------------------------------------------------------
// grid = nchrom.  bias <- 1, then min_nnz / min_count masks; marg[lo:hi] /= median(marg[lo:hi][>0])
__global__ void __launch_bounds__(1024)
ice_filter_chrom_kernel(const double* __restrict__ nnz_marg, double* __restrict__ marg,
                        const int64_t* __restrict__ chrom_off, int min_nnz, int min_count, int do_mad,
                        double* __restrict__ bias) {
    __shared__ HcSelectSmem sm;
    const int64_t lo = chrom_off[blockIdx.x], hi = chrom_off[blockIdx.x + 1];
    const long long n = hi - lo;
    double* m = marg + lo;
    long long cnt = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        double b = 1.0;
        if (min_nnz > 0 && nnz_marg[lo + i] < (double)min_nnz) b = 0.0;
        if (min_count != 0 && m[i] < (double)min_count) b = 0.0;
        bias[lo + i] = b;
        cnt += (m[i] > 0.0);
    }
    if (!do_mad) return;
    cnt = block_sum_ll(cnt, sm.redll);
    const double med = block_median([&](long long i) { return m[i]; }, [&](long long i) { return m[i] > 0.0; },
                                    n, cnt, &sm);
    __syncthreads();
    for (long long i = threadIdx.x; i < n; i += blockDim.x) m[i] = m[i] / med;
}

// single CTA.  logs of the positive normalised marginals -> median, MAD -> cutoff -> mask
__global__ void __launch_bounds__(1024)
ice_filter_mad_kernel(const double* __restrict__ marg, int64_t nbins, double mad_max,
                      double* __restrict__ bias, double* __restrict__ work) {
    __shared__ HcSelectSmem sm;
    double* lg = work;           // log(marg) where marg > 0
    double* dev = work + nbins;  // |lg - median|
    long long cnt = 0;
    for (long long i = threadIdx.x; i < nbins; i += blockDim.x) {
        const double v = marg[i];
        const bool ok = v > 0.0;  // false for NaN
        lg[i] = ok ? log(v) : 0.0;
        cnt += ok;
    }
    cnt = block_sum_ll(cnt, sm.redll);
    auto valid = [&](long long i) { return marg[i] > 0.0; };
    const double med = block_median([&](long long i) { return lg[i]; }, valid, nbins, cnt, &sm);
    __syncthreads();
    for (long long i = threadIdx.x; i < nbins; i += blockDim.x) dev[i] = fabs(lg[i] - med);
    __syncthreads();
    const double mad = block_median([&](long long i) { return dev[i]; }, valid, nbins, cnt, &sm);
    const double cutoff = exp(med - mad_max * mad);
    for (long long i = threadIdx.x; i < nbins; i += blockDim.x)
        if (marg[i] < cutoff) bias[i] = 0.0;
}

// ---------------------------------------------------------------------------------------
// the fused iteration
// ---------------------------------------------------------------------------------------
struct IceDenseArgs {
    const int32_t* mats; const int64_t* mat_off; const int32_t* mat_n; const int32_t* mat_ld;
    const int64_t* pad_off;     // start of each problem in the padded (ld-strided, 128 B aligned) vectors
    const int32_t* item_prob; const int32_t* item_row0; const int32_t* item_nrows;   // work items (~equal bytes)
    const unsigned int* nitems; unsigned int* queue;   // global work queue: warps draw items until it runs dry
    int32_t* iter;              // iteration counter, advanced on the device (the loop is replayed as a CUDA graph)
    double* bias;               // padded layout; updated in place by the update kernel
    double* marg;               // padded layout; fresh marginals of this iteration
    hc_ice_result* results; int32_t* done; int32_t* n_done;
    double tol; int kd; int max_iters; int nprob;
    unsigned int queue_start;   // number of warps of the stream kernel: items below it are pre-assigned
};

// one column chunk (128 columns, 4 per lane) of RG rows: acc[q] += sum_j w * A[rq][j] * b[j]
template <int RG>
__device__ __forceinline__ void consume_chunk(const int4 (&a)[RG], const double* __restrict__ b, int j, int jc,
                                              int rg, int kd, int kspan, double (&acc)[RG]) {
    const double2 b01 = *reinterpret_cast<const double2*>(b + j);        // through L1: reused by every row
    const double2 b23 = *reinterpret_cast<const double2*>(b + j + 2);
    const bool band = jc <= rg + RG - 1 + kspan && jc + 127 >= rg - kspan;   // warp-uniform
#pragma unroll
    for (int q = 0; q < RG; ++q) {
        double x0 = (double)a[q].x, x1 = (double)a[q].y, x2 = (double)a[q].z, x3 = (double)a[q].w;
        if (band) {
            const int r = rg + q;
            x0 *= band_weight(j, r, kd); x1 *= band_weight(j + 1, r, kd);
            x2 *= band_weight(j + 2, r, kd); x3 *= band_weight(j + 3, r, kd);
        }
        acc[q] = fma(x0, b01.x, acc[q]); acc[q] = fma(x1, b01.y, acc[q]);
        acc[q] = fma(x2, b23.x, acc[q]); acc[q] = fma(x3, b23.y, acc[q]);
    }
}

// Streaming half of an iteration: marg[r] = b[r] * sum_j w(r,j) A[r][j] b[j] for every row of every
// unconverged chromosome.  Persistent warps draw row-group items (~48 KB each) from a global queue,
// so the launch stays balanced however the chromosomes differ in size; the next item is drawn
// before the current one is processed, hiding the atomic's latency behind the row loads.
template <int RG, int U, int MINB>
__global__ void __launch_bounds__(256, MINB)
ice_dense_stream_kernel(IceDenseArgs A) {
    const int lane = threadIdx.x & 31;
    const int kd = A.kd, kspan = kd > 0 ? kd - 1 : 0;
    const unsigned nitems = *A.nitems;
    // the first item of a warp is its global index (4736 simultaneous draws on one address would serialise
    // the start of every launch); the queue counter is re-armed to the warp count by the update kernel
    unsigned item = 0;
    if (lane == 0) {
        item = (blockIdx.x * 256u + threadIdx.x) >> 5;
        if (item == 0) atomicAdd(A.iter, 1);       // exactly one warp per launch owns index 0: it opens iteration k
    }
    item = __shfl_sync(0xffffffffu, item, 0);
    while (item < nitems) {
        unsigned next = 0;
        if (lane == 0) next = atomicAdd(A.queue, 1u);
        const int p = A.item_prob[item];
        if (!A.done[p]) {
            const int ld = A.mat_ld[p];            // multiple of 128: every 128-column chunk is full
            const int nchunk = ld >> 7;
            const int64_t lo = A.pad_off[p];
            const int32_t* mat = A.mats + A.mat_off[p];
            const double* __restrict__ b = A.bias + lo;
            double* mout = A.marg + lo;
            const int r0 = A.item_row0[item], r1 = r0 + A.item_nrows[item];
            for (int rg = r0; rg < r1; rg += RG) {
                const int nr = min(RG, r1 - rg);
                const int32_t* base = mat + (int64_t)rg * ld + 4 * lane;
                int roff[RG];                      // ragged last group: re-read the last valid row, discard below
#pragma unroll
                for (int q = 0; q < RG; ++q) roff[q] = min(q, nr - 1) * ld;
                double acc[RG];
#pragma unroll
                for (int q = 0; q < RG; ++q) acc[q] = 0.0;
                int c0 = 0;
                for (; c0 + U <= nchunk; c0 += U) {   // RG x U independent 128-bit streaming loads in flight per lane
                    int4 a[U][RG];
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int q = 0; q < RG; ++q) a[u][q] = ld_stream_v4(base + roff[q] + (c0 + u) * 128);
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        consume_chunk<RG>(a[u], b, (c0 + u) * 128 + 4 * lane, (c0 + u) * 128, rg, kd, kspan, acc);
                }
                for (; c0 < nchunk; ++c0) {
                    int4 a[RG];
#pragma unroll
                    for (int q = 0; q < RG; ++q) a[q] = ld_stream_v4(base + roff[q] + c0 * 128);
                    consume_chunk<RG>(a, b, c0 * 128 + 4 * lane, c0 * 128, rg, kd, kspan, acc);
                }
#pragma unroll
                for (int q = 0; q < RG; ++q) {
                    const double sacc = warp_sum(acc[q]);
                    if (lane == 0 && q < nr) mout[rg + q] = b[rg + q] * sacc;
                }
            }
        }
        item = __shfl_sync(0xffffffffu, next, 0);
    }
}

// Vector half of an iteration (grid = one CTA per chromosome): mean / variance of the fresh
// marginals over the non-zero bins, bias update b /= marg/mean (in place), convergence test,
// scale / iteration bookkeeping -- all on the device.
__global__ void __launch_bounds__(1024)
ice_dense_update_kernel(IceDenseArgs A) {
    __shared__ double red[32];
    __shared__ long long redll[32];
    const int p = blockIdx.x;
    const int k = *A.iter;                                // opened by this iteration's stream kernel
    if (p == 0 && threadIdx.x == 0) *A.queue = A.queue_start;   // re-arm the work queue for the next iteration
    if (A.done[p] || k > A.max_iters) return;
    const int n = A.mat_n[p];
    if (n == 0) return;
    const int64_t lo = A.pad_off[p];
    const double* m_in = A.marg + lo;
    double* bw = A.bias + lo;
    double s = 0.0;
    long long c = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double m = m_in[j];
        if (m != 0.0) { s += m; ++c; }
    }
    s = block_sum(s, red);
    c = block_sum_ll(c, redll);
    if (c == 0) {   // nothing left to balance: cooler sets bias = NaN, scale = NaN, var = 0
        if (threadIdx.x == 0) {
            hc_ice_result r; r.scale = __longlong_as_double(0x7ff8000000000000ll); r.var = 0.0;
            r.iters = k; r.converged = 1;
            A.results[p] = r; A.done[p] = 1; atomicAdd(A.n_done, 1);
        }
        return;
    }
    const double mean = s / (double)c;
    double v = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const double m = m_in[j];
        if (m != 0.0) { const double d = m - mean; v += d * d; }
    }
    const double var = block_sum(v, red) / (double)c;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double m = m_in[j] / mean;
        if (m == 0.0) m = 1.0;
        bw[j] = bw[j] / m;
    }
    if (threadIdx.x == 0) {
        hc_ice_result r; r.scale = mean; r.var = var; r.iters = k; r.converged = var < A.tol;
        A.results[p] = r;
        if (var < A.tol || k >= A.max_iters) { A.done[p] = 1; atomicAdd(A.n_done, 1); }
    }
}

// user bias (concatenated bins) <-> padded internal layout
__global__ void __launch_bounds__(256)
ice_pad_bias_kernel(const int64_t* __restrict__ bin_off, const int64_t* __restrict__ pad_off, int nprob,
                    const double* __restrict__ bias, double* __restrict__ padded) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= bin_off[nprob]) return;
    int p = 0;
    while (p + 1 < nprob && bin_off[p + 1] <= g) ++p;
    padded[pad_off[p] + (g - bin_off[p])] = bias[g];
}

// final weights: bias==0 -> NaN; divide by sqrt(scale) when rescaling (cooler balance_cooler tail)
__global__ void __launch_bounds__(256)
ice_finalize_kernel(const int64_t* __restrict__ bin_off, const int64_t* __restrict__ pad_off, int nprob,
                    const hc_ice_result* __restrict__ results, int rescale, const double* __restrict__ padded,
                    double* __restrict__ bias) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= bin_off[nprob]) return;
    int p = 0;
    while (p + 1 < nprob && bin_off[p + 1] <= g) ++p;
    const hc_ice_result r = results[p];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double b = padded[pad_off[p] + (g - bin_off[p])];
    if (isnan(r.scale)) b = nan;
    else {
        if (b == 0.0) b = nan;
        if (rescale) b = b / sqrt(r.scale);
    }
    bias[g] = b;
}

typedef void (*StreamKernel)(IceDenseArgs);
struct StreamVariant { StreamKernel fn; int rg, u, minb; };
// tuned on B200 (profiles/): RG rows share the bias loads; RG*U 128-bit loads in flight per lane
const StreamVariant kVariants[] = {
    {ice_dense_stream_kernel<4, 2, 4>, 4, 2, 4},   // default: 5.5 TB/s on C2 (profiles/r1c_ice_variants_v3.log)
    {ice_dense_stream_kernel<4, 2, 3>, 4, 2, 3},
    {ice_dense_stream_kernel<2, 2, 4>, 2, 2, 4},
    {ice_dense_stream_kernel<2, 4, 3>, 2, 4, 3},
    {ice_dense_stream_kernel<4, 4, 2>, 4, 4, 2},
    {ice_dense_stream_kernel<1, 4, 4>, 1, 4, 4},
};

}  // namespace

extern "C" int hc_ice_dense_marginals(const int32_t* mats, const int64_t* mat_off, const int32_t* mat_n,
                                      const int32_t* mat_ld, const int64_t* bin_off, int32_t nprob,
                                      int32_t ignore_diags, double* nnz_marg, double* marg, void* stream) {
    HC_REQUIRE(nprob > 0 && ignore_diags >= 0, "nprob>0, ignore_diags>=0");
    int64_t total = 0;
    HC_CUDA(hc_read_small(&total, bin_off + nprob, sizeof(int64_t), (cudaStream_t)stream));
    if (total == 0) return HC_OK;
    const int64_t blocks = (total * 32 + 255) / 256;
    ice_dense_marginals_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        mats, mat_off, mat_n, mat_ld, bin_off, nprob, ignore_diags, nnz_marg, marg);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

extern "C" int hc_ice_filter_bins(const double* nnz_marg, double* marg, int64_t nbins, const int64_t* chrom_off,
                                  int32_t nchrom, const hc_ice_params* P, double* bias, double* work, void* stream) {
    HC_REQUIRE(nbins >= 0 && nchrom > 0 && P != nullptr, "nbins>=0, nchrom>0, params");
    if (nbins == 0) return HC_OK;
    const int do_mad = P->mad_max > 0.0;
    ice_filter_chrom_kernel<<<nchrom, 1024, 0, (cudaStream_t)stream>>>(nnz_marg, marg, chrom_off, P->min_nnz,
                                                                      P->min_count, do_mad, bias);
    HC_LAUNCH_CHECK();
    if (do_mad) {
        ice_filter_mad_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(marg, nbins, P->mad_max, bias, work);
        HC_LAUNCH_CHECK();
    }
    return HC_OK;
}

extern "C" int hc_ice_dense_balance(const int32_t* mats, const int64_t* mat_off, const int32_t* mat_n,
                                    const int32_t* mat_ld, const int64_t* bin_off, int32_t nprob,
                                    const int32_t* h_mat_n, const hc_ice_params* P, double* bias, double* work,
                                    hc_ice_result* results, hc_ice_run_info* h_info, void* stream) {
    HC_REQUIRE(nprob > 0 && h_mat_n != nullptr && P != nullptr, "nprob>0, h_mat_n, params");
    HC_REQUIRE(P->max_iters >= 1 && P->ignore_diags >= 0, "max_iters>=1, ignore_diags>=0");
    int64_t nbins = 0;
    for (int p = 0; p < nprob; ++p) {
        HC_REQUIRE(h_mat_n[p] >= 0, "matrix side");
        nbins += h_mat_n[p];
    }
    if (h_info) { h_info->launches = 0; h_info->loop_ms = 0.f; }
    if (nbins == 0) return HC_OK;
    // The iteration loop is replayed as a CUDA graph, which cannot be captured on the legacy default
    // stream: run on a private stream ordered after the caller's stream (the call synchronises
    // before returning, so the caller's later work is ordered after it).
    static thread_local cudaStream_t private_stream[64] = {nullptr};
    int dev = 0;
    HC_CUDA(cudaGetDevice(&dev));
    HC_REQUIRE(dev >= 0 && dev < 64, "device index");
    if (!private_stream[dev]) HC_CUDA(cudaStreamCreateWithFlags(&private_stream[dev], cudaStreamNonBlocking));
    cudaStream_t s = private_stream[dev];
    {
        cudaEvent_t e_in;
        HC_CUDA(cudaEventCreateWithFlags(&e_in, cudaEventDisableTiming));
        HC_CUDA(cudaEventRecord(e_in, (cudaStream_t)stream));
        HC_CUDA(cudaStreamWaitEvent(s, e_in, 0));
        cudaEventDestroy(e_in);
    }

    // variant: 4 rows per warp step is the fastest per byte, but a warp item is at least RG rows; when the
    // batch is small (one large chromosome per GPU in the 8-way sharded run) fewer rows per item keep
    // every resident warp busy.  HC_ICE_VARIANT overrides (tuning).
    int vi = -1;
    if (const char* e = getenv("HC_ICE_VARIANT")) vi = atoi(e);
    if (vi < 0 || vi >= (int)(sizeof(kVariants) / sizeof(kVariants[0]))) {
        int64_t rows = 0;
        for (int p = 0; p < nprob; ++p) rows += h_mat_n[p];
        const int64_t warps = (int64_t)hc_num_sms() * 4 * 8;
        vi = rows / 4 >= warps + warps / 4 ? 0 : (rows / 2 >= warps + warps / 4 ? 2 : 5);   // <4,2,4> | <2,2,4> | <1,4,4>
    }
    const StreamVariant V = kVariants[vi];
    int item_kb = 32;   // bytes of matrix per work item: small enough that the last item is a short tail
    if (const char* e = getenv("HC_ICE_ITEM_KB")) item_kb = std::max(1, atoi(e));

    // padded internal vectors: problem p owns [pad_off[p], pad_off[p] + ld_p), 128-byte aligned, so the
    // kernel can read the bias of any 4-column group with two aligned 16-byte loads and never
    // needs a column bound check (matrix padding columns are zero, padded bias entries are zero)
    std::vector<int32_t> h_ld(nprob);
    HC_CUDA(hc_read_small(h_ld.data(), mat_ld, sizeof(int32_t) * nprob, s));
    std::vector<int64_t> h_pad(nprob + 1, 0);
    for (int p = 0; p < nprob; ++p) {
        if (h_ld[p] < h_mat_n[p] || (h_ld[p] & 127) != 0) {
            hc_set_error("hc_ice_dense_balance: ld must be >= n and a multiple of 128 elements (matrix %d: n=%d ld=%d)",
                         p, h_mat_n[p], h_ld[p]);
            return HC_ERR_ARG;
        }
        h_pad[p + 1] = h_pad[p] + h_ld[p];
    }
    const int64_t npad = h_pad[nprob];
    (void)work;   // scratch is allocated stream-ordered below; `work` is kept for ABI stability

    // ---- work items: RG-aligned row groups of ~item_kb KB, largest chromosomes first -----------
    std::vector<int32_t> h_done(nprob, 0);
    std::vector<int32_t> item_prob, item_row0, item_nrows;
    auto build_items = [&]() {
        item_prob.clear(); item_row0.clear(); item_nrows.clear();
        std::vector<int> order(nprob);
        for (int p = 0; p < nprob; ++p) order[p] = p;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h_mat_n[a] > h_mat_n[b]; });
        for (int p : order) {
            const int n = h_mat_n[p];
            if (n == 0 || h_done[p]) continue;
            int rows = (int)((int64_t)item_kb * 1024 / ((int64_t)h_ld[p] * 4));
            rows = std::max(V.rg, rows / V.rg * V.rg);
            for (int r0 = 0; r0 < n; r0 += rows) {
                item_prob.push_back(p); item_row0.push_back(r0); item_nrows.push_back(std::min(rows, n - r0));
            }
        }
    };
    build_items();
    const size_t max_items = item_prob.size();

    double* d_vec = nullptr;     // [npad] bias | [npad] marg | [nprob+1] pad_off (as int64)
    HC_CUDA(cudaMallocAsync((void**)&d_vec, (2 * (size_t)npad + nprob + 1) * sizeof(double), s));
    double* biasp = d_vec;
    double* marg = d_vec + npad;
    int64_t* d_pad = reinterpret_cast<int64_t*>(d_vec + 2 * npad);
    HC_CUDA(cudaMemsetAsync(d_vec, 0, 2 * (size_t)npad * sizeof(double), s));
    HC_CUDA(cudaMemcpyAsync(d_pad, h_pad.data(), (nprob + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    ice_pad_bias_kernel<<<(unsigned)((nbins + 255) / 256), 256, 0, s>>>(bin_off, d_pad, nprob, bias, biasp);
    HC_LAUNCH_CHECK();
    int32_t* d_tab = nullptr;    // 3 item tables | done[nprob] | n_done | queue | iter | nitems
    const size_t tab_ints = 3 * max_items + (size_t)nprob + 4;
    HC_CUDA(cudaMallocAsync((void**)&d_tab, tab_ints * sizeof(int32_t), s));
    auto upload_items = [&]() -> cudaError_t {
        const size_t n = item_prob.size();
        const int32_t n32 = (int32_t)n;
        cudaError_t e = cudaMemcpyAsync(d_tab + 3 * max_items + nprob + 3, &n32, sizeof(int32_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab, item_prob.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab + max_items, item_row0.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab + 2 * max_items, item_nrows.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);   // the host vectors may be rebuilt afterwards
        return e;
    };
    HC_CUDA(cudaMemsetAsync(d_tab + 3 * max_items, 0, ((size_t)nprob + 3) * sizeof(int32_t), s));
    HC_CUDA(upload_items());
    const int grid = hc_num_sms() * V.minb;    // persistent: every resident slot filled exactly once
    const unsigned int queue_start = (unsigned)grid * 8u;
    HC_CUDA(cudaMemcpyAsync(d_tab + 3 * max_items + nprob + 1, &queue_start, sizeof(unsigned int), cudaMemcpyHostToDevice, s));

    IceDenseArgs A;
    A.mats = mats; A.mat_off = mat_off; A.mat_n = mat_n; A.mat_ld = mat_ld; A.pad_off = d_pad;
    A.item_prob = d_tab; A.item_row0 = d_tab + max_items; A.item_nrows = d_tab + 2 * max_items;
    A.done = d_tab + 3 * max_items;
    A.n_done = A.done + nprob;
    A.queue = reinterpret_cast<unsigned int*>(A.n_done + 1);
    A.iter = A.n_done + 2;
    A.nitems = reinterpret_cast<const unsigned int*>(A.n_done + 3);
    A.bias = biasp; A.marg = marg; A.results = results;
    A.tol = P->tol; A.kd = P->ignore_diags; A.max_iters = P->max_iters; A.nprob = nprob; A.queue_start = queue_start;

    int nonempty = 0;
    for (int p = 0; p < nprob; ++p) nonempty += h_mat_n[p] > 0;
    if (nonempty != nprob) {   // empty problems never get work: give them a defined result
        std::vector<hc_ice_result> h_res(nprob);
        for (int p = 0; p < nprob; ++p) { h_res[p].scale = NAN; h_res[p].var = 0.0; h_res[p].iters = 0; h_res[p].converged = 1; }
        HC_CUDA(cudaMemcpyAsync(results, h_res.data(), sizeof(hc_ice_result) * nprob, cudaMemcpyHostToDevice, s));
        HC_CUDA(cudaStreamSynchronize(s));
    }
    // the bias vector is the only data that should live in L1: no shared-memory carve-out
    cudaFuncSetAttribute(V.fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    const int poll = P->poll_every > 0 ? P->poll_every : 8;
    int launches = 0, h_ndone = 0, seen_done = 0;
    int rc = HC_OK;
    // `poll` iterations (stream + update kernel each) captured once and replayed: the iteration index,
    // the work-list length and every convergence decision live in device memory, so the graph is static
    cudaGraphExec_t gexec = nullptr;
    bool use_graph = true;
    if (const char* e = getenv("HC_ICE_GRAPH")) use_graph = atoi(e) != 0;
    if (use_graph) {
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            for (int i = 0; i < poll; ++i) {
                V.fn<<<grid, 256, 0, s>>>(A);
                ice_dense_update_kernel<<<nprob, 1024, 0, s>>>(A);
            }
            e = cudaStreamEndCapture(s, &graph);
        }
        if (e == cudaSuccess) e = cudaGraphInstantiate(&gexec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) { gexec = nullptr; (void)cudaGetLastError(); }   // plain launches below
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // device time of the iteration loop, for the roofline
    if (h_info) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, s); }
    for (int k0 = 0; k0 < P->max_iters; k0 += poll) {
        cudaError_t e = cudaSuccess;
        if (gexec) e = cudaGraphLaunch(gexec, s);
        else for (int i = 0; i < poll; ++i) {
            V.fn<<<grid, 256, 0, s>>>(A);
            ice_dense_update_kernel<<<nprob, 1024, 0, s>>>(A);
        }
        hc_count_launch(2 * poll);
        launches += 2 * poll;
        if (e == cudaSuccess) e = hc_read_small(&h_ndone, A.n_done, sizeof(int32_t), s);   // not a memcpy: see hc_read_small
        if (e == cudaSuccess && h_ndone >= nonempty) break;
        if (e == cudaSuccess && h_ndone != seen_done) {
            // drop the converged chromosomes from the work list (their items would only be skipped)
            seen_done = h_ndone;
            e = hc_read_small(h_done.data(), A.done, sizeof(int32_t) * nprob, s);
            if (e == cudaSuccess) { build_items(); e = upload_items(); }
        }
        if (e != cudaSuccess) { hc_set_error("hc_ice_dense_balance: %s", cudaGetErrorString(e)); rc = HC_ERR_CUDA; break; }
    }
    if (h_info && ev0) cudaEventRecord(ev1, s);
    if (gexec) cudaGraphExecDestroy(gexec);
    if (rc == HC_OK) {
        const int64_t blocks = (nbins + 255) / 256;
        ice_finalize_kernel<<<(unsigned)blocks, 256, 0, s>>>(bin_off, d_pad, nprob, results, P->rescale_marginals, biasp, bias);
        hc_count_launch();
        ++launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { hc_set_error("hc_ice_dense_balance: %s", cudaGetErrorString(e)); rc = HC_ERR_CUDA; }
    }
    cudaFreeAsync(d_tab, s);
    cudaFreeAsync(d_vec, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (rc == HC_OK && e != cudaSuccess) { hc_set_error("hc_ice_dense_balance: %s", cudaGetErrorString(e)); rc = HC_ERR_CUDA; }
    if (h_info) {
        h_info->launches = launches;
        if (ev0 && e == cudaSuccess) cudaEventElapsedTime(&h_info->loop_ms, ev0, ev1);
        if (ev0) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); }
    }
    return rc;
}
