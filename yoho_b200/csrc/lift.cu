// Group-feature lift tail (YOHO_testset.py:153-166): for each of the 60 group rotations, rotate the
// keypoints (Keys @ R_g^T, float64), find each rotated keypoint's nearest point in that rotation's down-sampled cloud
// (knn_module.KNN(1) on a float64 source against float32 targets -> float64 distances, utils/knn_search.py:17-24), and
// take that point's 32-d FCGF feature: out[k, :, g] = feat_g[nn_g(k), :]  ->  the [K,32,60] input of PartI.
// The FCGF backbone that produces feat_g stays out of scope.
#include "common.cuh"

namespace {

constexpr int LK = 64;        // keypoints per CTA
constexpr int LCH = 1024;     // cloud points staged per pass

// grid (ceil(K/64), 60).  256 threads: kp = t & 63, part = t >> 6 scans points [256*part, 256*part+256) of every pass in
// ascending order with strict '<', parts are merged lexicographically on (d2, index): the first minimal index wins, as
// torch.min does.  The comparison is on d2 = (dx^2+dy^2)+dz^2; sqrt(d2+1e-7) is monotone, so the argmin is the same unless
// two distinct d2 round to one double after the sqrt.
__global__ void __launch_bounds__(256) lift_kernel(const double* __restrict__ kps, int K, const double* __restrict__ rot,
                                                  const float* __restrict__ pts, const float* __restrict__ feats,
                                                  const int* __restrict__ offsets, float* __restrict__ out,
                                                  int64_t* __restrict__ nn_out) {
    __shared__ float sp[LCH * 3];
    __shared__ double bd[4][LK];
    __shared__ int bi[4][LK];
    __shared__ int nn[LK];
    const int g = blockIdx.y;
    const int t = threadIdx.x;
    const int kp = t & (LK - 1), part = t >> 6;
    const int k = blockIdx.x * LK + kp;
    const int off = offsets[g], n = offsets[g + 1] - off;
    const double* R = rot + g * 9;
    double qx = 0, qy = 0, qz = 0;
    if (k < K) {
        const double x = kps[3 * k], y = kps[3 * k + 1], z = kps[3 * k + 2];
        qx = (x * R[0] + y * R[1]) + z * R[2];      // Keys @ R_g^T
        qy = (x * R[3] + y * R[4]) + z * R[5];
        qz = (x * R[6] + y * R[7]) + z * R[8];
    }
    double best = INFINITY;
    int besti = 0;
    for (int base = 0; base < n; base += LCH) {
        const int m = n - base < LCH ? n - base : LCH;
        __syncthreads();
        for (int i = t; i < m * 3; i += 256) sp[i] = pts[(size_t)(off + base) * 3 + i];
        __syncthreads();
        const int lo = part * 256, hi = lo + 256 < m ? lo + 256 : m;
        for (int i = lo; i < hi; ++i) {
            const double dx = qx - (double)sp[3 * i], dy = qy - (double)sp[3 * i + 1], dz = qz - (double)sp[3 * i + 2];
            const double d2 = (dx * dx + dy * dy) + dz * dz;
            if (d2 < best) { best = d2; besti = base + i; }
        }
    }
    bd[part][kp] = best;
    bi[part][kp] = besti;
    __syncthreads();
    if (t < LK) {
        double b = bd[0][t];
        int ix = bi[0][t];
#pragma unroll
        for (int p = 1; p < 4; ++p)
            if (bd[p][t] < b || (bd[p][t] == b && bi[p][t] < ix)) { b = bd[p][t]; ix = bi[p][t]; }
        nn[t] = ix;
        const int kk = blockIdx.x * LK + t;
        if (nn_out && kk < K) nn_out[(size_t)g * K + kk] = ix;
    }
    __syncthreads();
    for (int i = t; i < LK * YF; i += 256) {
        const int q = i >> 5, c = i & 31;
        const int kk = blockIdx.x * LK + q;
        if (kk < K && n > 0) out[((size_t)kk * YF + c) * YG + g] = feats[((size_t)off + nn[q]) * YF + c];
    }
}

}  // namespace

extern "C" int yoho_lift_group_features(yoho_ctx* ctx, const double* kps, int K, const float* pts, const float* feats,
                                        const int32_t* offsets, float* out, int64_t* nn_out, void* stream) {
    YARG(ctx && kps && pts && feats && offsets && out && K >= 0);
    if (K == 0) return YOHO_OK;
    YCHECK(cudaSetDevice(ctx->device));
    dim3 grid((K + LK - 1) / LK, YG);
    lift_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kps, K, ctx->d_rot, pts, feats, offsets, out, nn_out);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
