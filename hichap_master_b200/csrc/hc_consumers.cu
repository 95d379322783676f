// the first downstream consumers of the matrix stage's outputs -- the weight vector and
// the (corrected) matrices -- in HiCHap/StructureFind.py:
//   balanced matrix       cooler.matrix(balance=True).fetch(chrom) + np.nan_to_num   (StructureFind.py:2005-2007)
//   Distance_Decay        mean contact per genomic distance with gap columns removed (:201-272)
//   observed / expected   the O/E matrix Get_PCA builds from that curve               (:321-329)
//   Get_DI                directionality index, t-test or chi-square flavour          (:804-840)
// Streaming kernels over dense tiles; every reduction has a fixed order (deterministic).
#include <math.h>
#include "hc_common.cuh"

namespace {

// out[i][j] = nan_to_num(M[i][j] * w[i] * w[j]); NaN weights (filtered bins) give NaN products -> 0
__global__ void __launch_bounds__(256)
balance_apply_kernel(const int32_t* __restrict__ M, int64_t ld, int n, const double* __restrict__ w,
                     double* __restrict__ out, int64_t ld_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    const double wi = w[i], v = (double)M[(int64_t)i * ld + j] * wi * w[j];      // (count * w_i) * w_j, cooler's order
    double r = v;
    if (isnan(v)) r = 0.0;
    else if (isinf(v)) r = v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;   // np.nan_to_num
    out[(int64_t)i * ld_out + j] = r;
}

// non-zero count per COLUMN (Distance_Decay's gap rule looks at columns, StructureFind.py:218)
__global__ void __launch_bounds__(256)
colnnz_f64_kernel(const double* __restrict__ M, int64_t ld, int n, int32_t* __restrict__ colnnz) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int c = 0;
    for (int i = 0; i < n; ++i) c += (M[(int64_t)i * ld + j] != 0.0);
    colnnz[j] = c;
}

// one warp per distance d: sum of M[i][j] over |i - j| == d with column j not a gap (both triangles), fixed order
__global__ void __launch_bounds__(256)
distance_sum_kernel(const double* __restrict__ M, int64_t ld, int n, const uint8_t* __restrict__ gap,
                    double* __restrict__ dsum) {
    const int d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (d >= n) return;
    double s = 0.0;
    for (int i = lane; i + d < n; i += 32) {
        const int j = i + d;
        if (!gap[j]) s += M[(int64_t)i * ld + j];            // entry (i, j): column j
        if (d > 0 && !gap[i]) s += M[(int64_t)j * ld + i];   // entry (j, i): column i
    }
    s = warp_sum(s);
    if (lane == 0) dsum[d] = s;
}

// OE[i][j] = M[i][j] / decline[|i - j|] where M != 0 (decline has no zeros: the caller replaced them, :318-319)
__global__ void __launch_bounds__(256)
oe_kernel(const double* __restrict__ M, int64_t ld, int n, const double* __restrict__ decline, double* __restrict__ out,
          int64_t ld_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    const double v = M[(int64_t)i * ld + j];
    out[(int64_t)i * ld_out + j] = v != 0.0 ? v / decline[i > j ? i - j : j - i] : 0.0;
}

// one thread per bin (the windows are a few tens of bins)
__global__ void __launch_bounds__(128)
di_kernel(const double* __restrict__ M, int64_t ld, int n, const uint8_t* __restrict__ gap,
          const int32_t* __restrict__ window_bin, int chitest, double* __restrict__ di) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int w = window_bin[j];
    double bias = 0.0;
    if (!gap[j] && w >= 1 && j >= w && j <= n - w - 1) {
        double up_sum = 0.0, down_sum = 0.0;
        for (int t = 1; t <= w; ++t) { up_sum += M[(int64_t)(j - t) * ld + j]; down_sum += M[(int64_t)(j + t) * ld + j]; }
        if (!chitest) {
            const double um = up_sum / w, dm = down_sum / w;
            const double den = (double)w * (double)(w - 1);
            double uq = 0.0, dq = 0.0;
            for (int t = 1; t <= w; ++t) {
                const double a = M[(int64_t)(j - t) * ld + j] - um, b = M[(int64_t)(j + t) * ld + j] - dm;
                uq += a * a / den; dq += b * b / den;
            }
            const double s = sqrt(uq + dq);
            if (s != 0.0) bias = (dm - um) / s;               // NaN when w == 1 (0/0), as NumPy gives the reference
        } else {
            const double e = (up_sum + down_sum) / 2.0;
            if (up_sum != down_sum && e != 0.0) {
                const double dd = down_sum - up_sum;
                bias = dd / fabs(dd) * ((up_sum - e) * (up_sum - e) / e + (down_sum - e) * (down_sum - e) / e);
            }
        }
    }
    di[j] = bias;
}

}  // namespace

extern "C" int hc_balance_apply_i32(const int32_t* M, int64_t ld, int32_t n, const double* weight, double* out,
                                    int64_t ld_out, void* stream) {
    HC_REQUIRE(n > 0 && ld >= n && ld_out >= n && M && weight && out, "arguments");
    balance_apply_kernel<<<dim3((n + 255) / 256, n), 256, 0, (cudaStream_t)stream>>>(M, ld, n, weight, out, ld_out);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

extern "C" int hc_colnnz_f64(const double* M, int64_t ld, int32_t n, int32_t* colnnz, void* stream) {
    HC_REQUIRE(n > 0 && ld >= n && M && colnnz, "arguments");
    colnnz_f64_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(M, ld, n, colnnz);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

extern "C" int hc_distance_sums_f64(const double* M, int64_t ld, int32_t n, const uint8_t* gapflag, double* dsum,
                                    void* stream) {
    HC_REQUIRE(n > 0 && ld >= n && M && gapflag && dsum, "arguments");
    distance_sum_kernel<<<(unsigned)(((int64_t)n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(M, ld, n, gapflag, dsum);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

extern "C" int hc_observed_expected_f64(const double* M, int64_t ld, int32_t n, const double* decline, double* out,
                                        int64_t ld_out, void* stream) {
    HC_REQUIRE(n > 0 && ld >= n && ld_out >= n && M && decline && out, "arguments");
    oe_kernel<<<dim3((n + 255) / 256, n), 256, 0, (cudaStream_t)stream>>>(M, ld, n, decline, out, ld_out);
    HC_LAUNCH_CHECK();
    return HC_OK;
}

extern "C" int hc_directionality_index_f64(const double* M, int64_t ld, int32_t n, const uint8_t* gapflag,
                                           const int32_t* window_bin, int32_t chitest, double* di, void* stream) {
    HC_REQUIRE(n > 0 && ld >= n && M && gapflag && window_bin && di, "arguments");
    di_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(M, ld, n, gapflag, window_bin, chitest, di);
    HC_LAUNCH_CHECK();
    return HC_OK;
}
