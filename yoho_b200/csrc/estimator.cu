// YOHO-C / YOHO-O transformation estimators (tests/estimator.py:28-141,279-340) in FP64 on the device.
//
// COMPILED WITH -fmad=false: every FP64 operation below rounds exactly as written (fma() only where it is
// spelled out), so the selected hypothesis and the inlier mask are reproducible bit for bit against the
// C restatement in oracle/estimator_oracle.c.  Arithmetic specification: DESIGN.md "Estimator arithmetic".
//
//   hypothesis  (E2, Threepps2Tran :55-63)  Kabsch from three matches, R = V U^T with NO reflection fix.
//       np.linalg.svd's sign for the null-space pair of the rank-2 cross-covariance is LAPACK noise; here the
//       SVD is a one-sided Jacobi iteration and the sign is s = sign(det H) (or the caller's override).
//   score       (E3, overlap_cal :66-70)    #{ m : ||k0_m - (R k1_m + t)||^2 < d^2 }, strict.
//   selection   (E4/E5 :132,:333)           first hypothesis with the strictly largest count (> 0).
#include "common.cuh"

namespace {

struct V3 { double x, y, z; };
__device__ __forceinline__ double dotp(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 crossp(const V3& a, const V3& b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Jacobi rotation of columns (p,q) of W and V (column-major storage: w[col]).
__device__ __forceinline__ void jacobi_pair(V3& wp, V3& wq, V3& vp, V3& vq) {
    const double alpha = dotp(wp, wp), beta = dotp(wq, wq), gamma = dotp(wp, wq);
    if (gamma == 0.0) return;
    const double zeta = (beta - alpha) / (2.0 * gamma);
    double t = 1.0 / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    if (zeta < 0.0) t = -t;
    const double c = 1.0 / sqrt(1.0 + t * t);
    const double s = c * t;
    V3 a = wp, b = wq;
    wp = V3{c * a.x - s * b.x, c * a.y - s * b.y, c * a.z - s * b.z};
    wq = V3{s * a.x + c * b.x, s * a.y + c * b.y, s * a.z + c * b.z};
    a = vp; b = vq;
    vp = V3{c * a.x - s * b.x, c * a.y - s * b.y, c * a.z - s * b.z};
    vq = V3{s * a.x + c * b.x, s * a.y + c * b.y, s * a.z + c * b.z};
}

__device__ void kabsch3(const double* __restrict__ k0, const double* __restrict__ k1, int i0, int i1, int i2,
                        int sign_override, double T[12]) {
    // centroids: ((p0 + p1) + p2) / 3
    double c0[3], c1[3], a[3][3], b[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        c0[d] = ((k0[3 * i0 + d] + k0[3 * i1 + d]) + k0[3 * i2 + d]) / 3.0;
        c1[d] = ((k1[3 * i0 + d] + k1[3 * i1 + d]) + k1[3 * i2 + d]) / 3.0;
    }
    const int id[3] = {i0, i1, i2};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            a[i][d] = k1[3 * id[i] + d] - c1[d];
            b[i][d] = k0[3 * id[i] + d] - c0[d];
        }
    double H[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) H[r][c] = fma(a[2][r], b[2][c], fma(a[1][r], b[1][c], a[0][r] * b[0][c]));

    double s;
    if (sign_override > 0) s = 1.0;
    else if (sign_override < 0) s = -1.0;
    else {
        const double m0 = H[1][1] * H[2][2] - H[1][2] * H[2][1];
        const double m1 = H[1][0] * H[2][2] - H[1][2] * H[2][0];
        const double m2 = H[1][0] * H[2][1] - H[1][1] * H[2][0];
        const double det = (H[0][0] * m0 - H[0][1] * m1) + H[0][2] * m2;
        s = det < 0.0 ? -1.0 : 1.0;
    }

    V3 w[3] = {{H[0][0], H[1][0], H[2][0]}, {H[0][1], H[1][1], H[2][1]}, {H[0][2], H[1][2], H[2][2]}};
    V3 v[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 8; ++sweep) {
        jacobi_pair(w[0], w[1], v[0], v[1]);
        jacobi_pair(w[0], w[2], v[0], v[2]);
        jacobi_pair(w[1], w[2], v[1], v[2]);
    }
    const double n0 = dotp(w[0], w[0]), n1 = dotp(w[1], w[1]), n2 = dotp(w[2], w[2]);
    const double nn[3] = {n0, n1, n2};
    int i1st = 0;
    if (nn[1] > nn[i1st]) i1st = 1;
    if (nn[2] > nn[i1st]) i1st = 2;
    int i2nd = -1;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        if (j == i1st) continue;
        if (i2nd < 0 || nn[j] > nn[i2nd]) i2nd = j;
    }
    double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (nn[i1st] != 0.0) {
        const double s1 = sqrt(nn[i1st]);
        const V3 wa = w[i1st], wb = w[i2nd];
        const V3 u1{wa.x / s1, wa.y / s1, wa.z / s1};
        const V3 v1 = v[i1st], v2 = v[i2nd];
        V3 u2;
        if (nn[i2nd] <= 1e-28 * nn[i1st]) {
            int j = 0;
            const double au[3] = {fabs(u1.x), fabs(u1.y), fabs(u1.z)};
            if (au[1] < au[j]) j = 1;
            if (au[2] < au[j]) j = 2;
            const double uj = j == 0 ? u1.x : (j == 1 ? u1.y : u1.z);
            u2 = V3{(j == 0 ? 1.0 : 0.0) - uj * u1.x, (j == 1 ? 1.0 : 0.0) - uj * u1.y, (j == 2 ? 1.0 : 0.0) - uj * u1.z};
        } else {
            const double s2 = sqrt(nn[i2nd]);
            u2 = V3{wb.x / s2, wb.y / s2, wb.z / s2};
            const double pr = dotp(u1, u2);
            u2 = V3{u2.x - pr * u1.x, u2.y - pr * u1.y, u2.z - pr * u1.z};
        }
        const double l2 = sqrt(dotp(u2, u2));
        u2 = V3{u2.x / l2, u2.y / l2, u2.z / l2};
        const V3 u3 = crossp(u1, u2), v3 = crossp(v1, v2);
        const double V1[3] = {v1.x, v1.y, v1.z}, V2[3] = {v2.x, v2.y, v2.z}, V3_[3] = {v3.x, v3.y, v3.z};
        const double U1[3] = {u1.x, u1.y, u1.z}, U2[3] = {u2.x, u2.y, u2.z}, U3[3] = {u3.x, u3.y, u3.z};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) R[r][c] = (V1[r] * U1[c] + V2[r] * U2[c]) + s * (V3_[r] * U3[c]);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        T[4 * r + 0] = R[r][0]; T[4 * r + 1] = R[r][1]; T[4 * r + 2] = R[r][2];
        T[4 * r + 3] = c0[r] - fma(R[r][2], c1[2], fma(R[r][1], c1[1], R[r][0] * c1[0]));
    }
}

// The inlier test on loaded coordinates: p = k1 point (x, y, z), q = k0 point (a, b, c).
__device__ __forceinline__ bool inlier_xyz(double x, double y, double z, double a, double b, double c, const double T[12], double thr2) {
    const double px = fma(T[2], z, fma(T[1], y, T[0] * x)) + T[3];
    const double py = fma(T[6], z, fma(T[5], y, T[4] * x)) + T[7];
    const double pz = fma(T[10], z, fma(T[9], y, T[8] * x)) + T[11];
    const double dx = a - px, dy = b - py, dz = c - pz;
    return fma(dz, dz, fma(dy, dy, dx * dx)) < thr2;
}
__device__ __forceinline__ bool is_inlier(const double* __restrict__ k0, const double* __restrict__ k1, int m,
                                          const double T[12], double thr2) {
    return inlier_xyz(k1[3 * m], k1[3 * m + 1], k1[3 * m + 2], k0[3 * m], k0[3 * m + 1], k0[3 * m + 2], T, thr2);
}

// Transform of hypothesis h.  YOHO-O: the given transform (through `order`).  YOHO-C: Kabsch of the triplet with the
// sign rule / the caller's sign; signs[h] == 2 takes the caller's transform `fixed[h]` instead (reference replay of the
// rank-deficient triplets, whose LAPACK completion is arbitrary: yoho_b200/estimator.py).
__device__ __forceinline__ void hypothesis_transform(const double* __restrict__ k0, const double* __restrict__ k1,
                                                     const int32_t* __restrict__ hyp, const int8_t* __restrict__ signs,
                                                     const double* __restrict__ fixed, const double* __restrict__ trans,
                                                     const int32_t* __restrict__ order, int h, double T[12]) {
    const double* src = nullptr;
    if (trans) src = trans + 12 * (size_t)(order ? order[h] : h);
    else if (signs && fixed && signs[h] == 2) src = fixed + 12 * (size_t)h;
    if (src) {
#pragma unroll
        for (int i = 0; i < 12; ++i) T[i] = src[i];
    } else {
        kabsch3(k0, k1, hyp[3 * h], hyp[3 * h + 1], hyp[3 * h + 2], signs ? (int)signs[h] : 0, T);
    }
}

// Four hypotheses per CTA of eight warps.  Warp w < 4 forms the transform of hypothesis 4*blockIdx.x + w (Kabsch from the triplet,
// or the given transform: YOHO-O); then every thread keeps all four transforms in registers and walks the matches with a stride
// of 256, testing each loaded match against the four: one set of (stride-24-byte) coordinate loads serves four hypotheses, and
// the kernel — one partial wave, bound by the latency of those loads — has eleven dependent rounds of them per thread at 2800
// matches.  The transforms go to `T_all` for select_kernel.
constexpr int SC_H = 4;
constexpr int SC_T = 256;
__global__ void __launch_bounds__(SC_T, 2) score_kernel(const double* __restrict__ k0, const double* __restrict__ k1, int M,
                                                         const int32_t* __restrict__ hyp, const int8_t* __restrict__ signs,
                                                         const double* __restrict__ fixed,
                                                         const double* __restrict__ trans, const int32_t* __restrict__ order,
                                                         int n_hyp, double thr2, double* __restrict__ T_all,
                                                         int32_t* __restrict__ counts) {
    __shared__ double Ts[SC_H][12];
    __shared__ int cs[SC_T / 32][SC_H];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h0 = blockIdx.x * SC_H;
    if (w < SC_H) {
        double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        const int h = h0 + w;
        if (h < n_hyp) hypothesis_transform(k0, k1, hyp, signs, fixed, trans, order, h, T);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                Ts[w][i] = T[i];
                if (h < n_hyp) T_all[12 * (size_t)h + i] = T[i];
            }
        }
    }
    __syncthreads();
    double T[SC_H][12];
#pragma unroll
    for (int q = 0; q < SC_H; ++q)
#pragma unroll
        for (int i = 0; i < 12; ++i) T[q][i] = Ts[q][i];
    int n[SC_H];
#pragma unroll
    for (int q = 0; q < SC_H; ++q) n[q] = 0;
#pragma unroll 2
    for (int m = threadIdx.x; m < M; m += SC_T) {
        const double x = k1[3 * m], y = k1[3 * m + 1], z = k1[3 * m + 2];
        const double a = k0[3 * m], b = k0[3 * m + 1], c = k0[3 * m + 2];
#pragma unroll
        for (int q = 0; q < SC_H; ++q) n[q] += inlier_xyz(x, y, z, a, b, c, T[q], thr2) ? 1 : 0;
    }
#pragma unroll
    for (int q = 0; q < SC_H; ++q) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) n[q] += __shfl_xor_sync(0xffffffffu, n[q], o);
        if (lane == 0) cs[w][q] = n[q];
    }
    __syncthreads();
    if (threadIdx.x < SC_H && h0 + (int)threadIdx.x < n_hyp) {
        int tot = 0;
#pragma unroll
        for (int v = 0; v < SC_T / 32; ++v) tot += cs[v][threadIdx.x];
        counts[h0 + threadIdx.x] = tot;
    }
}

// First strictly-best hypothesis, its transform and inlier mask.  Single CTA.
__global__ void __launch_bounds__(1024) select_kernel(const double* __restrict__ k0, const double* __restrict__ k1, int M,
                                                     const int32_t* __restrict__ hyp, const int8_t* __restrict__ signs,
                                                     const double* __restrict__ fixed,
                                                     const double* __restrict__ trans, const int32_t* __restrict__ order,
                                                     int n_hyp, double thr2, const int32_t* __restrict__ counts,
                                                     const double* __restrict__ T_all, double* __restrict__ T_out, int32_t* __restrict__ best_iter,
                                                     int32_t* __restrict__ n_inl, uint8_t* __restrict__ mask) {
    __shared__ unsigned long long red[32];
    __shared__ double Ts[12];
    __shared__ int bi_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    // key: larger count first, then smaller index  ->  maximise (count << 32) | (0xffffffff - index)
    unsigned long long best = 0;
    for (int h = t; h < n_hyp; h += 1024) {
        const unsigned long long k = ((unsigned long long)(unsigned)counts[h] << 32) | (unsigned)(0xffffffffu - (unsigned)h);
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, best, o);
        best = u > best ? u : best;
    }
    if (lane == 0) red[w] = best;
    __syncthreads();
    if (t == 0) {
        unsigned long long b = 0;
        for (int i = 0; i < 32; ++i) b = red[i] > b ? red[i] : b;
        const int cnt = (int)(b >> 32);
        const int bi = cnt > 0 ? (int)(0xffffffffu - (unsigned)(b & 0xffffffffu)) : -1;
        bi_s = bi;
        *best_iter = bi;
        *n_inl = cnt > 0 ? cnt : 0;
        double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        if (bi >= 0) {             // the transform score_kernel formed for this hypothesis (bit for bit what it scored)
            for (int i = 0; i < 12; ++i) T[i] = T_all[12 * (size_t)bi + i];
        }
        for (int i = 0; i < 12; ++i) { Ts[i] = T[i]; T_out[i] = T[i]; }
    }
    __syncthreads();
    if (mask) {
        double T[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) T[i] = Ts[i];
        const bool any = bi_s >= 0;
        for (int m = t; m < M; m += 1024) mask[m] = (any && is_inlier(k0, k1, m, T, thr2)) ? 1 : 0;
    }
}

// ---- Philox4x32-10 ------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * ctr.x;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * ctr.z;
        ctr = make_uint4((unsigned)(p1 >> 32) ^ ctr.y ^ key.x, (unsigned)p1, (unsigned)(p0 >> 32) ^ ctr.w ^ key.y, (unsigned)p0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

// E1 + draws.  Single CTA.  members are kept in ascending match order per bin, as the reference's lists.
constexpr int DRAW_T = 512;       // 512 threads: 128 registers per thread (1024 threads spilled 1 KB per thread at the 64-register cap)
__global__ void __launch_bounds__(DRAW_T, 1) c_draw_kernel(const int64_t* __restrict__ dr_index, int M, int iters,
                                                     unsigned long long seed, int32_t* __restrict__ members_ws,
                                                     int32_t* __restrict__ hyp, int32_t* __restrict__ status) {
    __shared__ int cnt[YG];
    __shared__ int off[YG + 1];
    __shared__ double cdf[YG];
    __shared__ int ok_s, bad_s;
    __shared__ uint8_t bins[16384];            // rotation bin of every match (M <= 16384 staged; larger M reads global)
    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    if (t < YG) cnt[t] = 0;
    if (t == 0) bad_s = 0;
    __syncthreads();
    // histogram: lanes of a warp that hold the same bin elect one leader (match_any), so a dominant bin (half of the matches of a
    // registered pair share the planted rotation) costs one shared-memory atomic per warp instead of 32 serialised ones
    for (int m0 = 0; m0 < M; m0 += DRAW_T) {
        const int m = m0 + t;
        int b = 64;                                                      // lanes past the end form their own group
        if (m < M) {
            // a rotation index outside [0,60) (a stale DR_index file) must not corrupt shared memory: clamp, and report status 2
            const long long raw = dr_index[m];
            b = raw < 0 ? 0 : (raw >= YG ? YG - 1 : (int)raw);
            if (raw != (long long)b) bad_s = 1;
            if (m < 16384) bins[m] = (uint8_t)b;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        if (b < YG && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&cnt[b], __popc(peers));
    }
    __syncthreads();
    // DR_statictic weights (tests/estimator.py:41-51): the per-bin products in parallel, the two running sums serially in bin
    // order on one thread (the same additions in the same order as the reference's loops), the divisions in parallel again
    __shared__ double w_s[YG];
    __shared__ double tot_s, last_s;
    if (t < YG) {
        double p = 0.0;
        if (cnt[t] >= 2) {
            const double num = (double)cnt[t] / 100.0;
            p = num * (num - 0.01) * (num - 0.02);
        }
        w_s[t] = p;
    }
    __syncthreads();
    if (t == 0) {
        double tot = 0.0;
        int o = 0;
        for (int i = 0; i < YG; ++i) { off[i] = o; o += cnt[i]; tot += w_s[i]; }
        off[YG] = o;
        tot_s = tot;
        ok_s = !(tot < 1e-4) && !bad_s;
        *status = bad_s ? 2 : (ok_s ? 0 : 1);
    }
    __syncthreads();
    if (ok_s) {
        // np.random.choice: cdf = cumsum(p / sum); cdf /= cdf[-1]
        if (t < YG) w_s[t] = w_s[t] / tot_s;
        __syncthreads();
        if (t == 0) {
            double c = 0.0;
            for (int i = 0; i < YG; ++i) { c += w_s[i]; cdf[i] = c; }
            last_s = c;
        }
        __syncthreads();
        if (t < YG) cdf[t] = cdf[t] / last_s;
    }
    // stable bucket fill (members of a bin in ascending match order, like the reference's append loop): per chunk of DRAW_T matches
    // a match's position = bin start + matches of the bin in earlier chunks + in earlier warps of this chunk + in earlier lanes of
    // its warp (match_any).  Replaces a per-bin scan of all matches (60 x M/32 dependent ballot steps).
    {
        __shared__ int wcnt[DRAW_T / 32][64];
        __shared__ int run[64];
        if (t < 64) run[t] = t < YG ? off[t] : 0;
        for (int m0 = 0; m0 < M; m0 += DRAW_T) {
            for (int i = t; i < (DRAW_T / 32) * 64; i += DRAW_T) (&wcnt[0][0])[i] = 0;
            __syncthreads();
            const int m = m0 + t;
            const int b = m < M ? (m < 16384 ? (int)bins[m] : (int)min(max(dr_index[m], (int64_t)0), (int64_t)(YG - 1))) : 64;
            const unsigned peers = __match_any_sync(0xffffffffu, b);
            const int before = __popc(peers & ((1u << lane) - 1u));
            if (b < YG && before == 0) wcnt[warp][b] = __popc(peers);
            __syncthreads();
            if (b < YG) {
                int pre = run[b];
                for (int w = 0; w < warp; ++w) pre += wcnt[w][b];
                members_ws[pre + before] = m;
            }
            __syncthreads();
            if (t < YG) {
                int tot = 0;
#pragma unroll
                for (int w = 0; w < DRAW_T / 32; ++w) tot += wcnt[w][t];
                run[t] += tot;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (!ok_s) {
        for (int i = t; i < 3 * iters; i += DRAW_T) hyp[i] = 0;
        return;
    }
    for (int it = t; it < iters; it += DRAW_T) {
        const uint4 r = philox4x32(make_uint4((unsigned)it, 0u, 0x59484f43u, 0u),
                                   make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
        // 53-bit uniform in [0,1) like random_sample(): (a >> 5, b >> 6)
        const double u = ((double)(r.x >> 5) * 67108864.0 + (double)(r.y >> 6)) / 9007199254740992.0;
        // searchsorted(cdf, u, side='right') over the non-decreasing cdf, clamped to the last bin: number of entries of
        // cdf[0..58] that are <= u (binary search; the linear scan it replaces cost up to 59 FP64 compares per draw)
        int lo_b = 0, hi_b = YG - 1;
        while (lo_b < hi_b) {
            const int mid = (lo_b + hi_b) >> 1;
            if (u < cdf[mid]) hi_b = mid; else lo_b = mid + 1;
        }
        int bin = lo_b;
        while (cnt[bin] == 0 && bin > 0) --bin;            // unreachable guard
        const uint4 r2 = philox4x32(make_uint4((unsigned)it, 1u, 0x59484f43u, 0u),
                                    make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
        const unsigned c = (unsigned)cnt[bin];
        const int base = off[bin];
        hyp[3 * it + 0] = members_ws[base + (int)(((unsigned long long)r2.x * c) >> 32)];
        hyp[3 * it + 1] = members_ws[base + (int)(((unsigned long long)r2.y * c) >> 32)];
        hyp[3 * it + 2] = members_ws[base + (int)(((unsigned long long)r2.z * c) >> 32)];
    }
}

// Random evaluation order: rank of a Philox key (ties by index).  Keys are generated once, then ranked with
// O(M^2) compares on cached keys (M is a few thousand).
__global__ void o_keys_kernel(int M, unsigned long long seed, unsigned long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const uint4 r = philox4x32(make_uint4((unsigned)i, 2u, 0x59484f4fu, 0u), make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    keys[i] = ((unsigned long long)r.x << 32) | r.y;
}
// 32 items per CTA; warp w compares its lane's key with every eighth key (slice w) staged in shared memory, the eight partial
// ranks are added through shared memory.  M compares per item spread over 8 threads and M/32 CTAs.
__global__ void __launch_bounds__(256) o_rank_kernel(int M, const unsigned long long* __restrict__ keys, int32_t* __restrict__ order) {
    __shared__ unsigned long long ks[2048];
    __shared__ int part[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    const unsigned long long ki = i < M ? keys[i] : 0ull;
    int rank = 0;
    for (int base = 0; base < M; base += 2048) {
        for (int j = threadIdx.x; j < 2048; j += 256) ks[j] = base + j < M ? keys[base + j] : ~0ull;
        __syncthreads();
        const int lim = M - base < 2048 ? M - base : 2048;
        for (int j = w; j < lim; j += 8) {
            const unsigned long long kj = ks[j];
            rank += (kj < ki || (kj == ki && base + j < i)) ? 1 : 0;
        }
        __syncthreads();
    }
    part[w][lane] = rank;
    __syncthreads();
    if (w == 0 && i < M) {
        int r = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) r += part[q][lane];
        order[r] = i;
    }
}

}  // namespace

extern "C" int yoho_c_draw(yoho_ctx* ctx, const int64_t* dr_index, int M, int iters, uint64_t seed, int32_t* hyp,
                           int32_t* status, void* stream) {
    YARG(ctx && dr_index && hyp && status && M >= 0 && iters >= 0);
    YCHECK(cudaSetDevice(ctx->device));
    if (int rc = yoho_ws_reserve(ctx, (size_t)(M + 1) * 4)) return rc;
    c_draw_kernel<<<1, DRAW_T, 0, (cudaStream_t)stream>>>(dr_index, M, iters, seed, (int32_t*)ctx->ws, hyp, status);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

static int score_and_select(yoho_ctx* ctx, const double* k0, const double* k1, int M, const int32_t* hyp,
                            const int8_t* signs, const double* fixed, const double* trans, const int32_t* order, int n_hyp,
                            double dist,
                            double* T, int32_t* best_iter, int32_t* n_inl, uint8_t* mask, int32_t* counts, cudaStream_t st) {
    // workspace: the transforms of all hypotheses (96 B each), then the counts when the caller does not want them
    if (int rc = yoho_ws_reserve(ctx, (size_t)(n_hyp + 1) * (96 + 4))) return rc;
    double* T_all = (double*)ctx->ws;
    int32_t* cnt = counts ? counts : (int32_t*)(T_all + 12 * (size_t)(n_hyp + 1));
    const double thr2 = dist * dist;
    if (n_hyp > 0) {
        score_kernel<<<(n_hyp + SC_H - 1) / SC_H, SC_T, 0, st>>>(k0, k1, M, hyp, signs, fixed, trans, order, n_hyp, thr2, T_all, cnt);
        ctx->launches++;
    }
    select_kernel<<<1, 1024, 0, st>>>(k0, k1, M, hyp, signs, fixed, trans, order, n_hyp, thr2, cnt, T_all, T, best_iter, n_inl, mask);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

extern "C" int yoho_c_ransac(yoho_ctx* ctx, const double* k0, const double* k1, int M, const int32_t* hyp,
                             const int8_t* signs, const double* fixed, int iters, double inlier_dist, double* T,
                             int32_t* best_iter, int32_t* n_inl, uint8_t* mask, int32_t* counts, void* stream) {
    YARG(ctx && k0 && k1 && T && best_iter && n_inl && M >= 0 && iters >= 0 && (iters == 0 || hyp));
    YCHECK(cudaSetDevice(ctx->device));
    return score_and_select(ctx, k0, k1, M, hyp, signs, fixed, nullptr, nullptr, iters, inlier_dist, T, best_iter, n_inl, mask,
                            counts, (cudaStream_t)stream);
}

extern "C" int yoho_o_order(yoho_ctx* ctx, int M, uint64_t seed, int32_t* order, void* stream) {
    YARG(ctx && order && M >= 0);
    if (M == 0) return YOHO_OK;
    YCHECK(cudaSetDevice(ctx->device));
    if (int rc = yoho_ws_reserve(ctx, (size_t)M * 8)) return rc;
    unsigned long long* keys = (unsigned long long*)ctx->ws;
    o_keys_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(M, seed, keys);
    o_rank_kernel<<<(M + 31) / 32, 256, 0, (cudaStream_t)stream>>>(M, keys, order);
    ctx->launches += 2;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

extern "C" int yoho_o_score(yoho_ctx* ctx, const double* k0, const double* k1, int M, const double* trans,
                            const int32_t* order, int H, double inlier_dist, double* T, int32_t* best_iter,
                            int32_t* n_inl, uint8_t* mask, int32_t* counts, void* stream) {
    YARG(ctx && k0 && k1 && T && best_iter && n_inl && M >= 0 && H >= 0 && (H == 0 || trans));
    YCHECK(cudaSetDevice(ctx->device));
    return score_and_select(ctx, k0, k1, M, nullptr, nullptr, nullptr, trans, order, H, inlier_dist, T, best_iter, n_inl, mask,
                            counts, (cudaStream_t)stream);
}
