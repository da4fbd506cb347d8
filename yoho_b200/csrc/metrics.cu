// Evaluation metrics on the device: the per-pair feature-match count of tests/evaluator.py:49-71 and the
// Redwood registration error / RRE / RTE of utils/RR_cal.py:13-65,273-301.  FP64, compiled with -fmad=false so every operation
// rounds as written (the thresholded decisions `dist < threshold`, `p <= err2` only differ from the reference on exact ties of
// differently-associated sums).  Tiny kernels: one block per pair for the match count, one thread per pair for the errors.
#include <cmath>
#include "common.cuh"

namespace {

constexpr int FMR_THREADS = 256;

// tests/evaluator.py:57-66.  keys0/keys1: matched keypoints of all pairs concatenated, pair p owns rows offsets[p]..offsets[p+1]-1;
// gt: [n][4][4] (a 3x4 ground truth is passed with the row 0 0 0 1: the homogeneous divide is then by exactly 1).
__global__ void __launch_bounds__(FMR_THREADS) fmr_count_kernel(const double* __restrict__ keys0, const double* __restrict__ keys1,
                                                                 const long long* __restrict__ offsets, const double* __restrict__ gt,
                                                                 double thr, int* __restrict__ counts) {
    const int pair = blockIdx.x;
    __shared__ double T[16];
    __shared__ int wsum[FMR_THREADS / 32];
    if (threadIdx.x < 16) T[threadIdx.x] = gt[(size_t)pair * 16 + threadIdx.x];
    __syncthreads();
    const long long beg = offsets[pair], end = offsets[pair + 1];
    int cnt = 0;
    for (long long m = beg + threadIdx.x; m < end; m += FMR_THREADS) {
        const double x = keys1[3 * m], y = keys1[3 * m + 1], z = keys1[3 * m + 2];
        const double hx = x * T[0] + y * T[1] + z * T[2] + T[3];
        const double hy = x * T[4] + y * T[5] + z * T[6] + T[7];
        const double hz = x * T[8] + y * T[9] + z * T[10] + T[11];
        const double hw = x * T[12] + y * T[13] + z * T[14] + T[15];
        const double dx = keys0[3 * m] - hx / hw, dy = keys0[3 * m + 1] - hy / hw, dz = keys0[3 * m + 2] - hz / hw;
        cnt += sqrt(dx * dx + dy * dy + dz * dz) < thr ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < FMR_THREADS / 32; ++w) s += wsum[w];
        counts[pair] = s;
    }
}

// 4x4 inverse by Gauss-Jordan elimination with partial pivoting (np.linalg.inv, RR_cal.py:273,289, is LU with partial pivoting:
// same pivots, differently ordered roundings).  Returns false for a singular matrix.
__device__ bool inv4(const double* a, double (&inv)[4][4]) {
    double m[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { m[i][j] = a[4 * i + j]; m[i][4 + j] = i == j ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(m[r][c]) > fabs(m[piv][c])) piv = r;
        if (m[piv][c] == 0.0) return false;
        if (piv != c) for (int j = 0; j < 8; ++j) { const double t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t; }
        const double d = m[c][c];
        for (int j = 0; j < 8; ++j) m[c][j] /= d;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double f = m[r][c];
            if (f != 0.0) for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) inv[i][j] = m[i][4 + j];
    return true;
}

// Largest eigenvector of a symmetric 4x4 matrix by cyclic Jacobi rotations (nibabel's mat2quat takes it from numpy's eigh).
__device__ void largest_eigvec4(double (&A)[4][4], double (&vec)[4]) {
    double V[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 16; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
        if (off < 1e-300) break;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {            // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) {            // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; ++i) if (A[i][i] > A[best][best]) best = i;
    for (int k = 0; k < 4; ++k) vec[k] = V[k][best];
}

// est, gt: [n][4][4]; info: [n][6][6].  p = Redwood error (RR_cal.py:48-65 of inv(gt) @ est), rre in degrees (RR_cal.py:13-33,
// including its float32 pi), rte (RR_cal.py:35-46).
__global__ void reg_err_kernel(const double* __restrict__ est, const double* __restrict__ gt, const double* __restrict__ info, int n,
                               double* __restrict__ p_out, double* __restrict__ rre, double* __restrict__ rte) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* E = est + (size_t)i * 16;
    const double* G = gt + (size_t)i * 16;
    if (p_out) {
        double gi[4][4];
        double p = nan("");
        if (inv4(G, gi)) {
            double tr[3][4];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 4; ++c) tr[r][c] = gi[r][0] * E[c] + gi[r][1] * E[4 + c] + gi[r][2] * E[8 + c] + gi[r][3] * E[12 + c];
            // nibabel.quaternions.mat2quat: Qxx,Qyx,Qzx,Qxy,... = M.flat
            const double Qxx = tr[0][0], Qyx = tr[0][1], Qzx = tr[0][2], Qxy = tr[1][0], Qyy = tr[1][1], Qzy = tr[1][2],
                         Qxz = tr[2][0], Qyz = tr[2][1], Qzz = tr[2][2];
            double K[4][4];
            K[0][0] = (Qxx - Qyy - Qzz) / 3.0; K[1][1] = (Qyy - Qxx - Qzz) / 3.0; K[2][2] = (Qzz - Qxx - Qyy) / 3.0; K[3][3] = (Qxx + Qyy + Qzz) / 3.0;
            K[1][0] = K[0][1] = (Qyx + Qxy) / 3.0; K[2][0] = K[0][2] = (Qzx + Qxz) / 3.0; K[2][1] = K[1][2] = (Qzy + Qyz) / 3.0;
            K[3][0] = K[0][3] = (Qyz - Qzy) / 3.0; K[3][1] = K[1][3] = (Qzx - Qxz) / 3.0; K[3][2] = K[2][3] = (Qxy - Qyx) / 3.0;
            double v[4];
            largest_eigvec4(K, v);
            double q[4] = {v[3], v[0], v[1], v[2]};
            if (q[0] < 0.0) for (int k = 0; k < 4; ++k) q[k] = -q[k];
            const double er[6] = {tr[0][3], tr[1][3], tr[2][3], q[1], q[2], q[3]};
            const double* I = info + (size_t)i * 36;
            double acc = 0.0;
            for (int c = 0; c < 6; ++c) {
                double vc = 0.0;
                for (int r = 0; r < 6; ++r) vc += er[r] * I[6 * r + c];
                acc += vc * er[c];
            }
            p = acc / I[0];
        }
        p_out[i] = p;
    }
    if (rre) {
        double tr = 0.0;                                   // trace(R_gt^T R_est) = sum_ij R_gt[i][j] R_est[i][j], diagonal by diagonal
        for (int d = 0; d < 3; ++d) tr += G[d] * E[d] + G[4 + d] * E[4 + d] + G[8 + d] * E[8 + d];
        double e = (tr - 1.0) / 2.0;
        e = fmin(fmax(e, -1.0), 1.0);
        rre[i] = 180.0 * acos(e) / (double)3.14159274101257324f;
    }
    if (rte) {
        const double dx = G[3] - E[3], dy = G[7] - E[7], dz = G[11] - E[11];
        rte[i] = sqrt(dx * dx + dy * dy + dz * dz);
    }
}

}  // namespace

extern "C" int yoho_fmr_batch(yoho_ctx* ctx, const double* keys0, const double* keys1, const int64_t* offsets, const double* gt, int n_pairs,
                              double threshold, int32_t* counts, void* stream) {
    YARG(ctx && n_pairs >= 0 && (n_pairs == 0 || (offsets && gt && counts)));
    if (n_pairs == 0) return YOHO_OK;
    YCHECK(cudaSetDevice(ctx->device));
    fmr_count_kernel<<<n_pairs, FMR_THREADS, 0, (cudaStream_t)stream>>>(keys0, keys1, (const long long*)offsets, gt, threshold, counts);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

extern "C" int yoho_registration_errors(yoho_ctx* ctx, const double* est, const double* gt, const double* info, int n, double* p, double* rre_deg,
                                        double* rte, void* stream) {
    YARG(ctx && n >= 0 && (n == 0 || (est && gt)) && (!p || info));
    if (n == 0) return YOHO_OK;
    YCHECK(cudaSetDevice(ctx->device));
    reg_err_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(est, gt, info, n, p, rre_deg, rte);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
