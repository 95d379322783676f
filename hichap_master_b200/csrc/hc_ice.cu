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
    const int64_t* bin_off;
    const int32_t* cta_prob; const int32_t* cta_row0; const int32_t* cta_row1;
    double* bias[2];   // ping-pong: launch k reads b_{k-2} from bias[k&1], writes b_{k-1} to bias[(k+1)&1]
    double* marg[2];   // launch k reads marg_{k-1} from marg[(k-1)&1], writes marg_k to marg[k&1]
    hc_ice_result* results; int32_t* done; int32_t* n_done;
    double tol; int kd; int max_iters;
};

__global__ void __launch_bounds__(1024)
ice_dense_iter_kernel(IceDenseArgs A, int k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sb = reinterpret_cast<double*>(smem_raw);  // staged bias, ld doubles (padding = 0)
    __shared__ double red[32];
    __shared__ long long redll[32];

    const int p = A.cta_prob[blockIdx.x];
    // the leader CTA of this same launch may set done[p] while we start: read it once, CTA-uniformly
    __shared__ int done_s;
    if (threadIdx.x == 0) done_s = *reinterpret_cast<volatile int32_t*>(A.done + p);
    __syncthreads();
    if (done_s) return;
    const int n = A.mat_n[p];
    const int64_t ld = A.mat_ld[p], lo = A.bin_off[p];
    const int row0 = A.cta_row0[blockIdx.x], row1 = A.cta_row1[blockIdx.x];
    const bool leader = row0 == 0;

    // ---- prologue: reduction over previous marginals, bias update, convergence test --------
    if (k == 1) {
        const double* b0 = A.bias[0] + lo;
        for (int j = threadIdx.x; j < (int)ld; j += blockDim.x) sb[j] = j < n ? b0[j] : 0.0;
    } else {
        const double* mprev = A.marg[(k - 1) & 1] + lo;
        const double* bprev = A.bias[k & 1] + lo;
        double s = 0.0;
        long long c = 0;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const double m = mprev[j];
            if (m != 0.0) { s += m; ++c; }
        }
        s = block_sum(s, red);
        c = block_sum_ll(c, redll);
        if (c == 0) {  // nothing left to balance: cooler sets bias = NaN, scale = NaN, var = 0
            if (leader && threadIdx.x == 0) {
                hc_ice_result r; r.scale = __longlong_as_double(0x7ff8000000000000ll); r.var = 0.0;
                r.iters = k - 1; r.converged = 1;
                A.results[p] = r; A.done[p] = 1; atomicAdd(A.n_done, 1);
            }
            return;
        }
        const double mean = s / (double)c;
        double v = 0.0;
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const double m = mprev[j];
            if (m != 0.0) { const double d = m - mean; v += d * d; }
        }
        const double var = block_sum(v, red) / (double)c;
        double* bout = A.bias[(k + 1) & 1] + lo;
        for (int j = threadIdx.x; j < (int)ld; j += blockDim.x) {
            double b = 0.0;
            if (j < n) {
                double m = mprev[j] / mean;
                if (m == 0.0) m = 1.0;
                b = bprev[j] / m;
                if (leader) bout[j] = b;
            }
            sb[j] = b;
        }
        if (var < A.tol || k - 1 >= A.max_iters) {
            if (leader && threadIdx.x == 0) {
                hc_ice_result r; r.scale = mean; r.var = var; r.iters = k - 1; r.converged = var < A.tol;
                A.results[p] = r; A.done[p] = 1; atomicAdd(A.n_done, 1);
            }
            return;
        }
    }
    __syncthreads();

    // ---- body: marg_k[r] = b[r] * sum_j w(r,j) A[r][j] b[j] ---------------------------------
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nvec = (int)(ld >> 2);
    const int nchunk = (nvec + 31) >> 5;
    const int32_t* mat = A.mats + A.mat_off[p];
    double* mout = A.marg[k & 1] + lo;
    const double2* sb2 = reinterpret_cast<const double2*>(sb);
    const int kd = A.kd, kspan = kd > 0 ? kd - 1 : 0;
    constexpr int U = 4;
    for (int r = row0 + wid; r < row1; r += nw) {
        const int32_t* row = mat + (int64_t)r * ld;
        double acc0 = 0.0, acc1 = 0.0;
        for (int c0 = 0; c0 < nchunk; c0 += U) {
            int4 a[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int v = (c0 + u) * 32 + lane;
                a[u] = v < nvec ? ld_stream_v4(row + 4 * v) : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int v = (c0 + u) * 32 + lane;
                if (v < nvec) {
                    const double2 b01 = sb2[2 * v], b23 = sb2[2 * v + 1];
                    double x0 = (double)a[u].x, x1 = (double)a[u].y, x2 = (double)a[u].z, x3 = (double)a[u].w;
                    const int jc = (c0 + u) * 128;  // warp-uniform: does this 128-column chunk touch the band?
                    if (jc <= r + kspan && jc + 127 >= r - kspan) {
                        const int j = 4 * v;
                        x0 *= band_weight(j, r, kd); x1 *= band_weight(j + 1, r, kd);
                        x2 *= band_weight(j + 2, r, kd); x3 *= band_weight(j + 3, r, kd);
                    }
                    acc0 = fma(x0, b01.x, acc0); acc1 = fma(x1, b01.y, acc1);
                    acc0 = fma(x2, b23.x, acc0); acc1 = fma(x3, b23.y, acc1);
                }
            }
        }
        const double acc = warp_sum(acc0 + acc1);
        if (lane == 0) mout[r] = sb[r] * acc;
    }
}

// final weights: bias==0 -> NaN; divide by sqrt(scale) when rescaling (cooler balance_cooler tail)
__global__ void __launch_bounds__(256)
ice_finalize_kernel(const int64_t* __restrict__ bin_off, int nprob, const hc_ice_result* __restrict__ results,
                    const double* b0, const double* b1, int rescale, double* out) {  // out may alias b0
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= bin_off[nprob]) return;
    int p = 0;
    while (p + 1 < nprob && bin_off[p + 1] <= g) ++p;
    const hc_ice_result r = results[p];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double b = (r.iters & 1) ? b1[g] : b0[g];
    if (isnan(r.scale)) b = nan;
    else {
        if (b == 0.0) b = nan;
        if (rescale) b = b / sqrt(r.scale);
    }
    out[g] = b;
}

}  // namespace

extern "C" int hc_ice_dense_marginals(const int32_t* mats, const int64_t* mat_off, const int32_t* mat_n,
                                      const int32_t* mat_ld, const int64_t* bin_off, int32_t nprob,
                                      int32_t ignore_diags, double* nnz_marg, double* marg, void* stream) {
    HC_REQUIRE(nprob > 0 && ignore_diags >= 0, "nprob>0, ignore_diags>=0");
    int64_t total = 0;
    HC_CUDA(cudaMemcpyAsync(&total, bin_off + nprob, sizeof(int64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    HC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
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
    cudaStream_t s = (cudaStream_t)stream;
    int64_t nbins = 0;
    int max_n = 0;
    double sum_sq = 0.0;
    for (int p = 0; p < nprob; ++p) {
        HC_REQUIRE(h_mat_n[p] >= 0, "matrix side");
        nbins += h_mat_n[p];
        max_n = std::max(max_n, h_mat_n[p]);
        sum_sq += (double)h_mat_n[p] * (double)h_mat_n[p];
    }
    if (h_info) { h_info->launches = 0; h_info->loop_ms = 0.f; }
    if (nbins == 0) return HC_OK;

    // ---- launch shape: the staged bias segment decides how many CTAs fit per SM -------------
    std::vector<int32_t> h_ld(nprob);
    HC_CUDA(cudaMemcpyAsync(h_ld.data(), mat_ld, sizeof(int32_t) * nprob, cudaMemcpyDeviceToHost, s));
    HC_CUDA(cudaStreamSynchronize(s));
    int ld_max = 0;
    for (int p = 0; p < nprob; ++p) {
        HC_REQUIRE(h_ld[p] >= h_mat_n[p] && (h_ld[p] & 3) == 0, "ld must be >= n and a multiple of 4");
        ld_max = std::max(ld_max, h_ld[p]);
    }
    const size_t smem = (size_t)ld_max * sizeof(double);
    int threads, ctas_per_sm;
    if (smem <= 54 * 1024) { threads = 256; ctas_per_sm = 4; }
    else if (smem <= 110 * 1024) { threads = 512; ctas_per_sm = 2; }
    else if (smem <= 220 * 1024) { threads = 1024; ctas_per_sm = 1; }
    else {
        hc_set_error("hc_ice_dense_balance: matrix side %d needs %zu B of shared memory for the staged bias; "
                     "use the CSR path for matrices this large", max_n, smem);
        return HC_ERR_UNSUPPORTED;
    }
    HC_CUDA(cudaFuncSetAttribute(ice_dense_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    // ---- static work plan: CTAs per problem proportional to n^2, contiguous row ranges -----
    const int G = hc_num_sms() * ctas_per_sm;
    std::vector<int32_t> cta_prob, cta_row0, cta_row1;
    for (int p = 0; p < nprob; ++p) {
        const int n = h_mat_n[p];
        if (n == 0) continue;
        int nc = (int)llround((double)G * ((double)n * n) / sum_sq);
        nc = std::max(1, std::min(nc, n));
        for (int c = 0; c < nc; ++c) {
            cta_prob.push_back(p);
            cta_row0.push_back((int32_t)((int64_t)n * c / nc));
            cta_row1.push_back((int32_t)((int64_t)n * (c + 1) / nc));
        }
    }
    const int ncta = (int)cta_prob.size();

    // scratch carved from `work` (3*nbins doubles) + a small device block for tables/flags
    double* bias1 = work;
    double* marg0 = work + nbins;
    double* marg1 = work + 2 * nbins;
    int32_t* d_tab = nullptr;
    const size_t tab_ints = (size_t)3 * ncta + nprob + 1;
    HC_CUDA(cudaMallocAsync((void**)&d_tab, tab_ints * sizeof(int32_t), s));
    HC_CUDA(cudaMemcpyAsync(d_tab, cta_prob.data(), ncta * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    HC_CUDA(cudaMemcpyAsync(d_tab + ncta, cta_row0.data(), ncta * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    HC_CUDA(cudaMemcpyAsync(d_tab + 2 * ncta, cta_row1.data(), ncta * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    HC_CUDA(cudaMemsetAsync(d_tab + 3 * ncta, 0, (nprob + 1) * sizeof(int32_t), s));

    IceDenseArgs A;
    A.mats = mats; A.mat_off = mat_off; A.mat_n = mat_n; A.mat_ld = mat_ld; A.bin_off = bin_off;
    A.cta_prob = d_tab; A.cta_row0 = d_tab + ncta; A.cta_row1 = d_tab + 2 * ncta;
    A.bias[0] = bias; A.bias[1] = bias1; A.marg[0] = marg0; A.marg[1] = marg1;
    A.results = results; A.done = d_tab + 3 * ncta; A.n_done = d_tab + 3 * ncta + nprob;
    A.tol = P->tol; A.kd = P->ignore_diags; A.max_iters = P->max_iters;

    int nonempty = 0;
    for (int p = 0; p < nprob; ++p) nonempty += h_mat_n[p] > 0;
    // empty problems never get a CTA: give them a defined result
    if (nonempty != nprob) {
        std::vector<hc_ice_result> h_res(nprob);
        for (int p = 0; p < nprob; ++p) { h_res[p].scale = NAN; h_res[p].var = 0.0; h_res[p].iters = 0; h_res[p].converged = 1; }
        HC_CUDA(cudaMemcpyAsync(results, h_res.data(), sizeof(hc_ice_result) * nprob, cudaMemcpyHostToDevice, s));
        HC_CUDA(cudaStreamSynchronize(s));
    }

    const int poll = P->poll_every > 0 ? P->poll_every : 8;
    int launches = 0, h_done = 0;
    int rc = HC_OK;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // device time of the iteration loop, for the roofline
    if (h_info) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, s); }
    for (int k = 1; k <= P->max_iters + 1; ++k) {
        ice_dense_iter_kernel<<<ncta, threads, smem, s>>>(A, k);
        hc_count_launch();
        ++launches;
        if (k % poll == 0 || k == P->max_iters + 1) {
            cudaError_t e = cudaMemcpyAsync(&h_done, A.n_done, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { hc_set_error("hc_ice_dense_balance: %s", cudaGetErrorString(e)); rc = HC_ERR_CUDA; break; }
            if (h_done >= nonempty) break;
        }
    }
    if (h_info && ev0) cudaEventRecord(ev1, s);
    if (rc == HC_OK) {
        const int64_t blocks = (nbins + 255) / 256;
        ice_finalize_kernel<<<(unsigned)blocks, 256, 0, s>>>(bin_off, nprob, results, bias, bias1,
                                                             P->rescale_marginals, bias);
        hc_count_launch();
        ++launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { hc_set_error("hc_ice_dense_balance: %s", cudaGetErrorString(e)); rc = HC_ERR_CUDA; }
    }
    cudaFreeAsync(d_tab, s);
    cudaError_t e = cudaStreamSynchronize(s);  // host tables above must outlive the copies
    if (rc == HC_OK && e != cudaSuccess) { hc_set_error("hc_ice_dense_balance: %s", cudaGetErrorString(e)); rc = HC_ERR_CUDA; }
    if (h_info) {
        h_info->launches = launches;
        if (ev0 && e == cudaSuccess) cudaEventElapsedTime(&h_info->loop_ms, ev0, ev1);
        if (ev0) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); }
    }
    return rc;
}
