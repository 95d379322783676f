// Inter-chromosomal imputation of one-sided allelic contacts on the genome-wide haplotype matrix.
// Replaces the per-line interpreted loops of HiCHap/matrixBuilding.py:1302-1378 (M_M file) and
// :1416-1492 (P_P file): a read whose allele is known on one mate only and whose mates lie on
// different chromosomes is credited to the maternal or the paternal copy of the other chromosome
// by a vote over a small disc of the UN-imputed matrix (GetNeighborhoodIndex :721-732 -- note its
// centre is (L+1, L+1), one off the window centre) around the two candidate cells.
//
// The un-imputed matrix is read-only during the whole pass (the reference votes on
// UnImputated_Whole_Lib and credits Imputated_Whole_Lib), so every line is independent: one warp per
// qualifying line, lanes over the disc points, shuffle reduction, one atomic per credited line.
//
// Bug-for-bug with the reference (the parity target; see oracle/hichap_oracle.py
// impute_inter_chromosomal, pinned on reference outputs in tests/golden/imputation_small.npz):
//   * R2 lines add the start of chromosome c1 to mate 2's bin and of c2 to mate 1's (:1343-1345);
//   * the P_P R1 branch votes with the stale `M_M_sub` window left by the M_M loop (:1448), i.e. a
//     constant -- the caller passes its disc sum (stale_sum) or says that the reference would have
//     raised there (stale_state 1: NameError, 2: IndexError) and the kernel reports whether such a
//     line exists (*stale_needed);
//   * P_P R2 credits [hap_bin1][bin2], M_M R2 credits [bin2][hap_bin1] (:1376 vs :1488).
#include "hc_common.cuh"

namespace {

struct ImputeArgs {
    const int32_t *c1, *p1, *c2, *p2;
    const uint8_t* mark;
    long long npairs;
    FastDiv res;
    const int64_t *start_m, *start_p;
    int nchrom, own_is_p;
    const int32_t* un;
    int32_t* imp;
    long long total, ld;
    int s;                              // half width of the window (Imputation_region // res)
    const int32_t *nb_i, *nb_j;         // disc points as window indices in [0, 2s]
    int npts;
    long long imin;
    double ratio;
    int stale_state;
    long long stale_sum;
    long long* last_qualifying;         // max index of a line that reached the window cut (M_M file: defines the stale window)
    int32_t* stale_needed;              // set when a P_P R1 line reached the vote
};

__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the disc of un[(r - s + i_k)][(c - s + j_k)], lanes striding over the points
__device__ __forceinline__ long long disc_sum(const ImputeArgs& a, long long r, long long c, int lane) {
    long long acc = 0;
    const int32_t* base = a.un + (r - a.s) * a.ld + (c - a.s);
    for (int k = lane; k < a.npts; k += 32) acc += __ldg(base + (long long)a.nb_i[k] * a.ld + a.nb_j[k]);
    return warp_sum_i64(acc);
}

__global__ void __launch_bounds__(256) impute_inter_kernel(ImputeArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long my_last = -1;
    for (long long base = warp * 32; base < a.npairs; base += nwarps * 32) {
        const long long i = base + lane;
        bool q = false;
        bool r1 = false;
        long long known = 0, m_o = 0, p_o = 0;
        if (i < a.npairs) {
            const int mk = a.mark[i];
            const int ca = a.c1[i], cb = a.c2[i];
            const int pa = a.p1[i], pb = a.p2[i];
            if (mk != 0 && ca >= 0 && cb >= 0 && ca < a.nchrom && cb < a.nchrom && ca != cb && pa >= 0 && pb >= 0) {
                const long long b1 = fast_div((uint32_t)pa, a.res), b2 = fast_div((uint32_t)pb, a.res);
                const int64_t* own = a.own_is_p ? a.start_p : a.start_m;
                r1 = mk == 1;
                const long long bk = r1 ? b1 : b2, bo = r1 ? b2 : b1;     // known-allele mate / the other mate
                known = bk + own[ca];
                m_o = bo + a.start_m[cb];
                p_o = bo + a.start_p[cb];
                const long long s = a.s, n = a.total;
                q = !(known < s || m_o < s || p_o < s) && !(known + s + 1 > n || m_o + s + 1 > n || p_o + s + 1 > n);
            }
        }
        if (q) my_last = i;   // i grows along the loop
        unsigned todo = __ballot_sync(0xffffffffu, q);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const long long k = __shfl_sync(0xffffffffu, known, src);
            const long long m = __shfl_sync(0xffffffffu, m_o, src);
            const long long p = __shfl_sync(0xffffffffu, p_o, src);
            const bool one = __shfl_sync(0xffffffffu, (int)r1, src) != 0;
            long long A, B;
            long long t0r, t0c, t1r, t1c;      // cell credited when A wins / when B wins
            if (!a.own_is_p) {
                if (one) { A = disc_sum(a, k, m, lane); B = disc_sum(a, k, p, lane); }
                else { A = disc_sum(a, m, k, lane); B = disc_sum(a, p, k, lane); }
                t0r = k; t0c = m; t1r = k; t1c = p;
            } else if (one) {
                if (a.stale_state != 0) {
                    if (lane == 0) *a.stale_needed = 1;
                    continue;
                }
                A = a.stale_sum;
                B = disc_sum(a, k, p, lane);
                t0r = k; t0c = m; t1r = k; t1c = p;
            } else {
                A = disc_sum(a, p, k, lane);
                B = disc_sum(a, m, k, lane);
                t0r = p; t0c = k; t1r = m; t1c = k;
            }
            if (lane == 0) {
                const double tot = (double)(A + B);
                const double ra = (double)A / tot, rb = (double)B / tot;     // 0/0 = NaN: both tests fail, as in NumPy
                if (A >= a.imin && ra > a.ratio) atomicAdd(a.imp + t0r * a.ld + t0c, 1);
                else if (B >= a.imin && rb > a.ratio) atomicAdd(a.imp + t1r * a.ld + t1c, 1);
            }
        }
    }
    if (a.last_qualifying) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_last = max(my_last, __shfl_xor_sync(0xffffffffu, my_last, o));
        if (lane == 0 && my_last >= 0) atomicMax(a.last_qualifying, my_last);
    }
}

}  // namespace

extern "C" int hc_impute_inter(const int32_t* c1, const int32_t* p1, const int32_t* c2, const int32_t* p2,
                               const uint8_t* mark, int64_t npairs, int32_t res, const int64_t* start_m,
                               const int64_t* start_p, int32_t nchrom, int32_t own_is_p, const int32_t* un,
                               int32_t* imp, int32_t total, int64_t ld, int32_t half_width, const int32_t* nb_i,
                               const int32_t* nb_j, int32_t npts, int64_t imputation_min, double imputation_ratio,
                               int32_t stale_state, int64_t stale_sum, long long* last_qualifying,
                               int32_t* stale_needed, void* stream) {
    HC_REQUIRE(npairs >= 0 && res > 0 && nchrom > 0 && total > 0 && ld >= total, "sizes");
    HC_REQUIRE(half_width >= 0 && npts >= 0 && (npts == 0 || (nb_i != nullptr && nb_j != nullptr)), "neighbourhood");
    HC_REQUIRE(stale_state >= 0 && stale_state <= 2, "stale_state");
    HC_REQUIRE(!own_is_p || stale_needed != nullptr, "stale_needed is required for the P_P file");
    if (npairs == 0) return HC_OK;
    HC_REQUIRE(mark != nullptr, "mark column required");
    ImputeArgs a;
    a.c1 = c1; a.p1 = p1; a.c2 = c2; a.p2 = p2; a.mark = mark; a.npairs = npairs;
    a.res = make_fast_div((uint32_t)res);
    a.start_m = start_m; a.start_p = start_p; a.nchrom = nchrom; a.own_is_p = own_is_p != 0;
    a.un = un; a.imp = imp; a.total = total; a.ld = ld; a.s = half_width; a.nb_i = nb_i; a.nb_j = nb_j; a.npts = npts;
    a.imin = imputation_min; a.ratio = imputation_ratio; a.stale_state = stale_state; a.stale_sum = stale_sum;
    a.last_qualifying = last_qualifying; a.stale_needed = stale_needed;
    long long blocks = (npairs + 255) / 256;
    const long long cap = (long long)hc_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    impute_inter_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    HC_LAUNCH_CHECK();
    return HC_OK;
}
