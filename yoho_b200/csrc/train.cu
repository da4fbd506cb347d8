// Training-time twins of the hot kernels, FP32, reference tensor layouts ([B,C,60], group axis innermost):
//
//   * group convolution forward / backward-data / backward-weight — Comb_Conv's `data[:,:,Nei].reshape(B,C,60,13)` +
//     Conv2d(C,O,(1,13)) (utils/network.py:12-21,46-52,80-84) and its two gradients, without materialising the 13x gathered
//     tensor.  Weights are read in the reference's own layout W[O,C,1,13] (they change every optimiser step: no packing).
//   * rotation-correlation backward — gradient of cor[b,a] = sum_{f,g} des1[b,f,P[a][g]] des2[b,f,g], the score of
//     Batch_hard_Rindex_loss.eqvloss (train/loss_val.py:27-31) and of PartI_train.Des2DR (utils/network.py:115-118); the forward
//     is yoho_rot_argmax's `cor_out`.
//
// Deterministic: every output element is produced by one thread with a fixed summation order (no atomics).
#include "common.cuh"

namespace {

// out[b][co][g] = bias[co] + sum_{ci,k} W(k,ci,co) * in[b][ci][idx[g][k]]
// W(k,ci,co) = w[co*s_co + ci*s_ci + k]: forward (ci = c, co = o): s_co = C*13, s_ci = 13;
// backward-data (ci = o, co = c, idx = inverse tables): s_co = 13, s_ci = C*13.
constexpr int TC_CO = 64;      // output channels per CTA
constexpr int TC_CI = 8;       // input channels per shared-memory chunk
__global__ void __launch_bounds__(256) gconv_nchw_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ bias, const int* __restrict__ idx,
                                                        float* __restrict__ out, int Cin, int Cout, long s_co, long s_ci) {
    __shared__ float ins[TC_CI][YG];
    __shared__ float ws[YT][TC_CI][TC_CO];
    __shared__ unsigned char nb[YG][YT];
    const int b = blockIdx.x, co0 = blockIdx.y * TC_CO, t = threadIdx.x;
    const int co = t % TC_CO, gq = t / TC_CO;                 // 4 groups of 15 group elements
    for (int i = t; i < YG * YT; i += 256) nb[i / YT][i % YT] = (unsigned char)idx[i];
    float acc[15];
    const float bv = (bias && co0 + co < Cout) ? bias[co0 + co] : 0.f;
#pragma unroll
    for (int i = 0; i < 15; ++i) acc[i] = bv;
    const float* inb = in + (size_t)b * Cin * YG;
    for (int c0 = 0; c0 < Cin; c0 += TC_CI) {
        __syncthreads();
        for (int i = t; i < TC_CI * YG; i += 256) {
            const int c = i / YG, g = i % YG;
            ins[c][g] = (c0 + c < Cin) ? inb[(size_t)(c0 + c) * YG + g] : 0.f;
        }
        for (int i = t; i < YT * TC_CI * TC_CO; i += 256) {
            const int o = i % TC_CO, c = (i / TC_CO) % TC_CI, k = i / (TC_CO * TC_CI);
            ws[k][c][o] = (co0 + o < Cout && c0 + c < Cin) ? w[(size_t)(co0 + o) * s_co + (size_t)(c0 + c) * s_ci + k] : 0.f;
        }
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < TC_CI; ++c) {
#pragma unroll
            for (int k = 0; k < YT; ++k) {
                const float wv = ws[k][c][co];
#pragma unroll
                for (int i = 0; i < 15; ++i) acc[i] = fmaf(wv, ins[c][nb[gq * 15 + i][k]], acc[i]);
            }
        }
    }
    if (co0 + co < Cout) {
        float* ob = out + ((size_t)b * Cout + co0 + co) * YG + gq * 15;
#pragma unroll
        for (int i = 0; i < 15; ++i) ob[i] = acc[i];
    }
}

// dW[o][c][k] = sum_b sum_g dy[b][o][g] * x[b][c][N[g][k]];  CTA = 32 o x 32 c, thread = 4 (o,c) pairs x 13 taps.
// db[o] = sum_{b,g} dy[b][o][g] (CTAs with blockIdx.y == 0).
__global__ void __launch_bounds__(256) gconv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         const int* __restrict__ nei, float* __restrict__ dw,
                                                         float* __restrict__ db, int B, int C, int O) {
    __shared__ float xs[32][YG + 1];
    __shared__ float ds[32][YG + 1];
    __shared__ unsigned char nb[YG][YT];
    const int o0 = blockIdx.x * 32, c0 = blockIdx.y * 32, t = threadIdx.x;
    for (int i = t; i < YG * YT; i += 256) nb[i / YT][i % YT] = (unsigned char)nei[i];
    const int c = t % 32, oq = t / 32;                        // o = oq + 8*p, p = 0..3
    float acc[4][YT];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int k = 0; k < YT; ++k) acc[p][k] = 0.f;
    float bsum = 0.f;
    for (int b = 0; b < B; ++b) {
        __syncthreads();
        for (int i = t; i < 32 * YG; i += 256) {
            const int r = i / YG, g = i % YG;
            xs[r][g] = (c0 + r < C) ? x[((size_t)b * C + c0 + r) * YG + g] : 0.f;
            ds[r][g] = (o0 + r < O) ? dy[((size_t)b * O + o0 + r) * YG + g] : 0.f;
        }
        __syncthreads();
#pragma unroll 2
        for (int g = 0; g < YG; ++g) {
            float xv[YT];
#pragma unroll
            for (int k = 0; k < YT; ++k) xv[k] = xs[c][nb[g][k]];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float d = ds[oq + 8 * p][g];
#pragma unroll
                for (int k = 0; k < YT; ++k) acc[p][k] = fmaf(d, xv[k], acc[p][k]);
            }
        }
        if (db && blockIdx.y == 0 && t < 32) {
            float s = 0.f;
            for (int g = 0; g < YG; ++g) s += ds[t][g];
            bsum += s;
        }
    }
    if (c0 + c < C) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int o = o0 + oq + 8 * p;
            if (o < O) {
                float* dst = dw + ((size_t)o * C + c0 + c) * YT;
#pragma unroll
                for (int k = 0; k < YT; ++k) dst[k] = acc[p][k];
            }
        }
    }
    if (db && blockIdx.y == 0 && t < 32 && o0 + t < O) db[o0 + t] = bsum;
}

// One CTA per row b.  g1[f][j] = sum_a dC[a] d2[f][Pinv[a][j]],  g2[f][g] = sum_a dC[a] d1[f][P[a][g]].
__global__ void __launch_bounds__(256) rot_cor_backward_kernel(const float* __restrict__ des1, const float* __restrict__ des2,
                                                              const float* __restrict__ gcor, const uint8_t* __restrict__ perm,
                                                              float* __restrict__ g1, float* __restrict__ g2, int F) {
    __shared__ float d1[YF][YG + 1], d2[YF][YG + 1];
    __shared__ float dc[YG];
    __shared__ uint8_t P[YG][YG], Pinv[YG][YG];
    const int b = blockIdx.x, t = threadIdx.x;
    for (int i = t; i < YG * YG; i += 256) {
        const int a = i / YG, g = i % YG;
        const uint8_t j = perm[i];                            // P[a][g]
        P[a][g] = j;
        Pinv[a][j] = (uint8_t)g;
    }
    for (int i = t; i < F * YG; i += 256) {
        d1[i / YG][i % YG] = des1[(size_t)b * F * YG + i];
        d2[i / YG][i % YG] = des2[(size_t)b * F * YG + i];
    }
    if (t < YG) dc[t] = gcor[(size_t)b * YG + t];
    __syncthreads();
    for (int i = t; i < F * YG; i += 256) {
        const int f = i / YG, j = i % YG;
        float s1 = 0.f, s2 = 0.f;
        for (int a = 0; a < YG; ++a) {
            s1 = fmaf(dc[a], d2[f][Pinv[a][j]], s1);
            s2 = fmaf(dc[a], d1[f][P[a][j]], s2);
        }
        if (g1) g1[(size_t)b * F * YG + i] = s1;
        if (g2) g2[(size_t)b * F * YG + i] = s2;
    }
}

}  // namespace

// inverse neighbour tables: for every tap k, g -> N[g][k] is a permutation of the group (N[g][k] = idx(R_{h_k} R_g)), so
// sum_{g : N[g][k] = j} dy[g] = dy[Ninv[j][k]].
static int ensure_inverse_tables(yoho_ctx* ctx) {
    if (ctx->d_idx_full_inv) return YOHO_OK;
    std::vector<int> n(YG * YT), inv(YG * YT, -1);
    YCHECK(cudaMemcpy(n.data(), ctx->d_idx_full, sizeof(int) * YG * YT, cudaMemcpyDeviceToHost));
    for (int g = 0; g < YG; ++g)
        for (int k = 0; k < YT; ++k) {
            if (inv[n[g * YT + k] * YT + k] != -1) {
                yoho_set_error("neighbour table column %d is not a permutation of the group", k);
                return YOHO_ERR_ARG;
            }
            inv[n[g * YT + k] * YT + k] = g;
        }
    YCHECK(cudaMalloc((void**)&ctx->d_idx_full_inv, sizeof(int) * YG * YT));
    YCHECK(cudaMemcpy(ctx->d_idx_full_inv, inv.data(), sizeof(int) * YG * YT, cudaMemcpyHostToDevice));
    return YOHO_OK;
}

extern "C" int yoho_gconv_train_forward(yoho_ctx* ctx, const float* x, const float* weight, const float* bias, int B, int C, int O,
                                        float* y, void* stream) {
    YARG(ctx && x && weight && y && B >= 0 && C > 0 && O > 0);
    if (B == 0) return YOHO_OK;
    YCHECK(cudaSetDevice(ctx->device));
    dim3 grid(B, (O + TC_CO - 1) / TC_CO);
    gconv_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, weight, bias, ctx->d_idx_full, y, C, O, (long)C * YT, (long)YT);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

extern "C" int yoho_gconv_train_backward(yoho_ctx* ctx, const float* x, const float* weight, const float* dy, int B, int C, int O,
                                         float* dx, float* dweight, float* dbias, void* stream) {
    YARG(ctx && weight && dy && B >= 0 && C > 0 && O > 0 && (x || !dweight));
    YCHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (dx && B > 0) {
        if (int rc = ensure_inverse_tables(ctx)) return rc;
        dim3 grid(B, (C + TC_CO - 1) / TC_CO);
        gconv_nchw_kernel<<<grid, 256, 0, st>>>(dy, weight, nullptr, ctx->d_idx_full_inv, dx, O, C, (long)YT, (long)C * YT);
        ctx->launches++;
    }
    if (dweight) {
        dim3 grid((O + 31) / 32, (C + 31) / 32);
        gconv_wgrad_kernel<<<grid, 256, 0, st>>>(x, dy, ctx->d_idx_full, dweight, dbias, B, C, O);
        ctx->launches++;
    }
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

extern "C" int yoho_rot_correlation_backward(yoho_ctx* ctx, const float* des1, const float* des2, const float* grad_cor, int M,
                                             float* grad_des1, float* grad_des2, void* stream) {
    YARG(ctx && des1 && des2 && grad_cor && M >= 0 && (grad_des1 || grad_des2));
    if (M == 0) return YOHO_OK;
    YCHECK(cudaSetDevice(ctx->device));
    rot_cor_backward_kernel<<<M, 256, 0, (cudaStream_t)stream>>>(des1, des2, grad_cor, ctx->d_perm, grad_des1, grad_des2, YF);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
