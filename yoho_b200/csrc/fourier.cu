// Group-Fourier transforms of the PartI activations (see yoho_b200/fourier.py and DESIGN.md §2.2).
//
// One kernel, two small dense products per (keypoint, 128-channel block) with the channel axis contiguous:
//     mid[m][c] = sum_k M1[k][m] * in[k][c]              (k, m in 0..59: group elements or Fourier coefficients)
//     mid      += bias[c] (+ resid[m][c]);  mid = relu(mid*scale[c] + shift[c])          (optional, group domain)
//     out[m][c] = sum_k M2[k][m] * mid[k][c]              (optional second product)
// used as   forward:            M1 = F^T                                    a1 (group)   -> X^ (bf16 hi/lo)
//           inverse+act+forward: M1 = F, bias/BN/ReLU, M2 = F^T             Y^ (Fourier) -> X^ of the next layer
//           inverse+act:        M1 = F, bias + shortcut, BN/ReLU            Y^           -> a3 (group, bf16 hi/lo)
// F is orthogonal, so the inverse transform is F^T.  FP32 FMA in a fixed order (k ascending): deterministic.
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int XC = 128;      // channels per CTA
constexpr int XM = 64;       // padded row count of the transform matrices

struct XfArgs {
    const float* in;         // [B][60][C]
    const float* m1;         // [60][64]: m1[k][m]
    const float* m2;         // nullable
    const float* bias;       // nullable [C]
    const float* resid;      // nullable [B][60][C]
    const float* scale;      // nullable [C] (with shift): BN + ReLU
    const float* shift;
    unsigned short* out_hi;  // nullable [B][60][C]
    unsigned short* out_lo;
    float* out_f32;          // nullable
    int B, C;
};

__device__ __forceinline__ void small_product(const float* __restrict__ ms, const float* __restrict__ xs, int m0, int c0,
                                              float (&acc)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < YG; ++k) {
        const float4 ma = *reinterpret_cast<const float4*>(ms + k * XM + m0);
        const float4 mb = *reinterpret_cast<const float4*>(ms + k * XM + m0 + 4);
        const float4 xv = *reinterpret_cast<const float4*>(xs + k * XC + c0);
        const float mm[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
        const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(mm[i], xx[j], acc[i][j]);
    }
}

__global__ void __launch_bounds__(256, 2) group_transform_kernel(const XfArgs p) {
    extern __shared__ __align__(16) float sm[];
    float* in_s = sm;                       // [60][128]
    float* m1_s = in_s + YG * XC;           // [60][64]
    float* m2_s = m1_s + YG * XM;           // [60][64]
    float* mid_s = m2_s + YG * XM;          // [64][128]
    const int b = blockIdx.x, cb = blockIdx.y * XC, t = threadIdx.x;
    const float* src = p.in + (size_t)b * YG * p.C + cb;
    for (int i = t; i < YG * XC / 4; i += 256) {
        const int k = i / (XC / 4), c4 = i % (XC / 4);
        *reinterpret_cast<float4*>(in_s + k * XC + c4 * 4) = *reinterpret_cast<const float4*>(src + (size_t)k * p.C + c4 * 4);
    }
    for (int i = t; i < YG * XM / 4; i += 256) {
        reinterpret_cast<float4*>(m1_s)[i] = reinterpret_cast<const float4*>(p.m1)[i];
        if (p.m2) reinterpret_cast<float4*>(m2_s)[i] = reinterpret_cast<const float4*>(p.m2)[i];
    }
    __syncthreads();
    const int m0 = (t >> 5) * 8, c0 = (t & 31) * 4;
    float acc[8][4];
    small_product(m1_s, in_s, m0, c0, acc);
    float bs[4] = {0, 0, 0, 0}, sc[4] = {1, 1, 1, 1}, sh[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (p.bias) bs[j] = p.bias[cb + c0 + j];
        if (p.scale) { sc[j] = p.scale[cb + c0 + j]; sh[j] = p.shift[cb + c0 + j]; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + i;
        if (m >= YG) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bs[j];
        if (p.resid) {
            const float4 r = *reinterpret_cast<const float4*>(p.resid + ((size_t)b * YG + m) * p.C + cb + c0);
            v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        if (p.scale) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = v[j];
        if (p.m2) *reinterpret_cast<float4*>(mid_s + m * XC + c0) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (p.m2) {
        __syncthreads();
        small_product(m2_s, mid_s, m0, c0, acc);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + i;
        if (m >= YG) continue;
        const size_t o = ((size_t)b * YG + m) * p.C + cb + c0;
        if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + o) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (p.out_hi) {
            unsigned short h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __nv_bfloat16 hh = __float2bfloat16_rn(acc[i][j]);
                h[j] = __bfloat16_as_ushort(hh);
                l[j] = __bfloat16_as_ushort(__float2bfloat16_rn(acc[i][j] - __bfloat162float(hh)));
            }
            *reinterpret_cast<uint2*>(p.out_hi + o) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
            *reinterpret_cast<uint2*>(p.out_lo + o) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
        }
    }
}

constexpr size_t XF_SMEM = (size_t)(YG * XC + 2 * YG * XM + XM * XC) * sizeof(float);

}  // namespace

int group_transform(yoho_ctx* ctx, const float* in, int B, int C, const float* m1, const float* m2, const float* bias,
                    const float* resid, const float* scale, const float* shift, void* out_hi, void* out_lo, float* out_f32,
                    cudaStream_t st) {
    YARG(C % XC == 0 && B > 0 && in && m1);
    YCHECK(cudaFuncSetAttribute(group_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XF_SMEM));
    XfArgs p{in, m1, m2, bias, resid, scale, shift, (unsigned short*)out_hi, (unsigned short*)out_lo, out_f32, B, C};
    group_transform_kernel<<<dim3(B, C / XC), 256, XF_SMEM, st>>>(p);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
