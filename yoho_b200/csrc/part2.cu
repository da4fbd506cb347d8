// PartII residual-rotation network and its post-processing
// (tests/extractor.py:125-138,185-201; utils/network.py:218-278; utils/r_eval.py:94-110).
//
// Only group element 0 of the last group convolution is consumed (utils/network.py:272-276: the mean is dead
// code and the head output is sliced at [:,:,0,0]), so the layers are evaluated on the receptive field of
// g=0 only: Conv_init at the 45 two-hop elements, comb_layer_in at the 13 one-hop elements, comb_layer_out and
// the 1x1 head at g=0.  This is exact (same sums, same order) and 5.6x cheaper.
#include <cuda_bf16.h>
#include "common.cuh"

int gconv_forward_grouped(yoho_ctx* ctx, const GLayer* const* Ls, const GConvArgs* as, int n, cudaStream_t st);

int gconv_split_bf16(yoho_ctx* ctx, const float* x, void* hi, void* lo, size_t n, cudaStream_t st);   // gconv_tc.cu

namespace {

// z0a[m][g][0:128] = relu(BN_init(concat_c(P_r FCGF_B, FCGF_A, P_r YOHO_B, YOHO_A)))   one CTA per match.
// batch_create swaps fragment 0<->1 (tests/extractor.py:132-137): "eqv0" tensors come from fragment id1 (B)
// and are permuted along g by P[pre_idx] (utils/network.py:266-268).
__global__ void __launch_bounds__(256) part2_assemble_kernel(const float* __restrict__ fcgf0, const float* __restrict__ fcgf1,
                                                            const float* __restrict__ yoho0, const float* __restrict__ yoho1,
                                                            const int64_t* __restrict__ pairs, const int64_t* __restrict__ pre_idx,
                                                            const uint8_t* __restrict__ perm, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, float* __restrict__ z0a,
                                                            unsigned short* __restrict__ z_hi, unsigned short* __restrict__ z_lo, int M) {
    __shared__ float s[4][YF][YG + 1];
    __shared__ uint8_t pr[64];
    const int m = blockIdx.x, t = threadIdx.x;
    const int64_t ra = pairs ? pairs[2 * (size_t)m] : m;
    const int64_t rb = pairs ? pairs[2 * (size_t)m + 1] : m;
    const int r = (int)pre_idx[m];
    if (t < YG) pr[t] = perm[r * YG + t];
    const float* src[4] = {fcgf1 + (size_t)rb * YF * YG, fcgf0 + (size_t)ra * YF * YG,
                           yoho1 + (size_t)rb * YF * YG, yoho0 + (size_t)ra * YF * YG};
#pragma unroll
    for (int q = 0; q < 4; ++q)
        for (int i = t; i < YF * YG; i += 256) s[q][i / YG][i % YG] = src[q][i];
    __syncthreads();
    const size_t o = (size_t)m * YG * 128;
    for (int i = t; i < YG * 128; i += 256) {
        const int g = i >> 7, c = i & 127;
        const int q = c >> 5, cc = c & 31;
        const int gs = (q == 0 || q == 2) ? pr[g] : g;
        const float v = fmaxf(fmaf(s[q][cc][gs], scale[c], shift[c]), 0.f);
        if (z_hi) {   // tensor-core path: bf16 hi/lo split
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            z_hi[o + i] = __bfloat16_as_ushort(h);
            z_lo[o + i] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
        } else {
            z0a[o + i] = v;
        }
    }
}

// Last 1x1 conv (128 -> 4), quaternion normalisation, R(q) in float32 as the reference evaluates it on numpy
// float32 scalars, R = R(q) @ Rgroup_f32[idx] and t = k0 - R k1 in float64.  One warp per match.
__global__ void __launch_bounds__(128) part2_head_kernel(const float* __restrict__ h2, int pitch, const float* __restrict__ w3,
                                                        const float* __restrict__ b3, const int64_t* __restrict__ pairs,
                                                        const int64_t* __restrict__ pre_idx, const float* __restrict__ rot32,
                                                        const double* __restrict__ kps0, const double* __restrict__ kps1,
                                                        float* __restrict__ quat, double* __restrict__ trans, int M) {
    const int m = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    float h[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < 128; c += 32) {
        const float v = h2[(size_t)m * pitch + c];
#pragma unroll
        for (int o = 0; o < 4; ++o) h[o] = fmaf(v, w3[c * 4 + o], h[o]);   // w3 packed [128][4]
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) h[o] += __shfl_xor_sync(0xffffffffu, h[o], s);
        h[o] += b3[o];
    }
    if (lane != 0) return;
    const float n = sqrtf(((h[0] * h[0] + h[1] * h[1]) + h[2] * h[2]) + h[3] * h[3]);
    const float w = h[0] / n, x = h[1] / n, y = h[2] / n, z = h[3] / n;
    quat[4 * (size_t)m + 0] = w; quat[4 * (size_t)m + 1] = x; quat[4 * (size_t)m + 2] = y; quat[4 * (size_t)m + 3] = z;
    if (!trans) return;
    // matrix_from_quaternion on float32 scalars: every product/sum rounds to float32, left to right.
    const float two = 2.f;
#define M3(a, b, c) __fmul_rn(__fmul_rn(a, b), c)
    float q[9];
    q[0] = __fsub_rn(__fsub_rn(1.f, M3(two, y, y)), M3(two, z, z));
    q[1] = __fsub_rn(M3(two, x, y), M3(two, z, w));
    q[2] = __fadd_rn(M3(two, x, z), M3(two, y, w));
    q[3] = __fadd_rn(M3(two, x, y), M3(two, z, w));
    q[4] = __fsub_rn(__fsub_rn(1.f, M3(two, x, x)), M3(two, z, z));
    q[5] = __fsub_rn(M3(two, y, z), M3(two, x, w));
    q[6] = __fsub_rn(M3(two, x, z), M3(two, y, w));
    q[7] = __fadd_rn(M3(two, y, z), M3(two, x, w));
    q[8] = __fsub_rn(__fsub_rn(1.f, M3(two, x, x)), M3(two, y, y));
#undef M3
    const float* rg = rot32 + (size_t)pre_idx[m] * 9;
    double R[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            R[3 * i + j] = ((double)q[3 * i] * (double)rg[j] + (double)q[3 * i + 1] * (double)rg[3 + j]) +
                           (double)q[3 * i + 2] * (double)rg[6 + j];
    const int64_t ra = pairs ? pairs[2 * (size_t)m] : m;
    const int64_t rb = pairs ? pairs[2 * (size_t)m + 1] : m;
    const double* k0 = kps0 + 3 * (size_t)ra;
    const double* k1 = kps1 + 3 * (size_t)rb;
    double* T = trans + 12 * (size_t)m;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        T[4 * i + 0] = R[3 * i]; T[4 * i + 1] = R[3 * i + 1]; T[4 * i + 2] = R[3 * i + 2];
        T[4 * i + 3] = k0[i] - ((k1[0] * R[3 * i] + k1[1] * R[3 * i + 1]) + k1[2] * R[3 * i + 2]);
    }
}

__global__ void gather_kps_kernel(const double* __restrict__ kps0, const double* __restrict__ kps1,
                                  const int64_t* __restrict__ pairs, int M, double* __restrict__ o0, double* __restrict__ o1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * M) {
        const int m = i / 3, d = i % 3;
        o0[i] = kps0[3 * pairs[2 * (size_t)m] + d];
        o1[i] = kps1[3 * pairs[2 * (size_t)m + 1] + d];
    }
}

// z3[m][c] = bias[c] + z1[m][g = 0][c] + the five tap-split partial sums of the last group convolution, added in split order.
__global__ void part2_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, const float* __restrict__ z1,
                                    int zero_pos, float* __restrict__ z3, unsigned short* __restrict__ z3_hi,
                                    unsigned short* __restrict__ z3_lo, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // float4 index over [n][64]
    if (i >= n * 64) return;
    const int m = i >> 6, c4 = i & 63;
    const float4* p = reinterpret_cast<const float4*>(part) + (size_t)m * 5 * 64 + c4;
    float4 acc = p[0];
#pragma unroll
    for (int q = 1; q < 5; ++q) {
        const float4 v = p[q * 64];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const float4 b = reinterpret_cast<const float4*>(bias)[c4];
    const float4 r = reinterpret_cast<const float4*>(z1 + ((size_t)m * 45 + zero_pos) * 256)[c4];
    acc.x = (acc.x + b.x) + r.x; acc.y = (acc.y + b.y) + r.y; acc.z = (acc.z + b.z) + r.z; acc.w = (acc.w + b.w) + r.w;
    reinterpret_cast<float4*>(z3)[i] = acc;
    if (z3_hi) {                                                  // bf16 hi/lo image for the tensor-core head
        const float f[4] = {acc.x, acc.y, acc.z, acc.w};
        unsigned short h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat16 hh = __float2bfloat16_rn(f[k]);
            h[k] = __bfloat16_as_ushort(hh);
            l[k] = __bfloat16_as_ushort(__float2bfloat16_rn(f[k] - __bfloat162float(hh)));
        }
        reinterpret_cast<uint2*>(z3_hi)[i] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        reinterpret_cast<uint2*>(z3_lo)[i] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
}

constexpr int P2_CHUNK = 4096;

}  // namespace

extern "C" int yoho_part2_forward(yoho_ctx* ctx, const float* fcgf0, const float* fcgf1, const float* yoho0,
                                  const float* yoho1, const int64_t* pairs, const int64_t* pre_idx, int M,
                                  const double* kps0, const double* kps1, float* quat, double* trans, void* stream) {
    YARG(ctx && fcgf0 && fcgf1 && yoho0 && yoho1 && pre_idx && quat && M >= 0);
    YARG(!trans || (kps0 && kps1));
    if (!ctx->has_p2) {
        yoho_set_error("No model exists: yoho_part2_load has not been called");
        return YOHO_ERR_NOWEIGHTS;
    }
    if (M == 0) return YOHO_OK;
    cudaStream_t st = (cudaStream_t)stream;
    YCHECK(cudaSetDevice(ctx->device));
    const size_t per = sizeof(float) * ((size_t)YG * 128 + 45 * 256 * 2 + 13 * 512 + 256 + 512 + 256 + 5 * 256 + 256);
    const int chunk = M < P2_CHUNK ? M : P2_CHUNK;
    if (int rc = yoho_ws_reserve(ctx, per * (size_t)chunk)) return rc;
    for (int s = 0; s < M; s += chunk) {
        const int n = (M - s) < chunk ? (M - s) : chunk;
        float* z0a = (float*)ctx->ws;
        float* z1 = z0a + (size_t)n * YG * 128;
        float* a1 = z1 + (size_t)n * 45 * 256;
        float* a2 = a1 + (size_t)n * 45 * 256;
        float* z3 = a2 + (size_t)n * 13 * 512;
        float* h1 = z3 + (size_t)n * 256;
        float* h2 = h1 + (size_t)n * 512;
        float* zpart = h2 + (size_t)n * 256;         // [n][5][256] tap-split partial sums of the last group convolution
        unsigned short* z3_hi = (unsigned short*)(zpart + (size_t)n * 5 * 256);      // bf16 hi|lo image of z3 (tensor-core head)
        unsigned short* z3_lo = z3_hi + (size_t)n * 256;
        unsigned short* h1_hi = (unsigned short*)h1;
        unsigned short* h1_lo = h1_hi + (size_t)n * 512;
        const int64_t* pr = pairs ? pairs + 2 * (size_t)s : nullptr;
        // without a match list the inputs are already per-match rows: advance them with the chunk
        const size_t adv = pairs ? 0 : (size_t)s * YF * YG;
        // tensor-core path (>= 128 rows in the smallest layer): the activation regions hold bf16 hi|lo halves
        const bool tc = ctx->gconv_impl >= 1 && ctx->p2_init.w_hi && ctx->p2_a.w_hi && ctx->p2_b.w_hi && n >= 128;
        unsigned short* z0_hi = (unsigned short*)z0a;
        unsigned short* z0_lo = z0_hi + (size_t)n * YG * 128;
        unsigned short* a1_hi = (unsigned short*)a1;
        unsigned short* a1_lo = a1_hi + (size_t)n * 45 * 256;
        unsigned short* a2_hi = (unsigned short*)a2;
        unsigned short* a2_lo = a2_hi + (size_t)n * 13 * 512;
        part2_assemble_kernel<<<n, 256, 0, st>>>(fcgf0 + adv, fcgf1 + adv, yoho0 + adv, yoho1 + adv, pr, pre_idx + s,
                                                 ctx->d_perm, ctx->p2_bn_init.scale, ctx->p2_bn_init.shift, z0a,
                                                 tc ? z0_hi : nullptr, tc ? z0_lo : nullptr, n);
        ctx->launches++;
        GConvArgs a{};
        a.B = n;
        // Conv_init at the 45 two-hop elements: raw z1 (shortcut) + a1 = relu(BN_a(z1))
        a.idx = ctx->d_idx_p2_init; a.Jin = YG; a.Jout = 45;
        a.out_raw = z1; a.scale = ctx->p2_bn_a.scale; a.shift = ctx->p2_bn_a.shift;
        if (tc) { a.act_hi = z0_hi; a.act_lo = z0_lo; a.out_hi = a1_hi; a.out_lo = a1_lo; } else { a.act = z0a; a.out_act = a1; }
        if (int rc = gconv_forward(ctx, ctx->p2_init, a, st)) return rc;
        // comb_layer_in at the 13 one-hop elements
        a.idx = ctx->d_idx_p2_a; a.Jin = 45; a.Jout = 13;
        a.out_raw = nullptr; a.scale = ctx->p2_bn_b.scale; a.shift = ctx->p2_bn_b.shift;
        if (tc) { a.act_hi = a1_hi; a.act_lo = a1_lo; a.out_hi = a2_hi; a.out_lo = a2_lo; } else { a.act = a1; a.out_act = a2; }
        if (int rc = gconv_forward(ctx, ctx->p2_a, a, st)) return rc;
        // comb_layer_out at g=0 + shortcut z1[:, g=0]
        a.idx = ctx->d_idx_p2_b; a.Jin = 13; a.Jout = 1;
        a.resid = z1; a.Jres = 45; a.resid_off = ctx->hop2_zero_pos; a.resid_per_j = 0;
        a.out_raw = z3; a.out_act = nullptr; a.out_hi = a.out_lo = nullptr; a.scale = a.shift = nullptr;
        if (tc) { a.act_hi = a2_hi; a.act_lo = a2_lo; } else { a.act = a2; }
        if (tc && ctx->p2_b_split[4].w_hi && !(ctx->tc_flags & 1024)) {
            GConvArgs fs[5];
            const GLayer* Ls[5];
            int t0 = 0;
            for (int q = 0; q < 5; ++q) {
                GConvArgs f{};
                f.B = n; f.Jin = 13; f.Jout = 1; f.idx = ctx->d_idx_p2_b + t0;
                f.act_hi = a2_hi; f.act_lo = a2_lo;
                f.out_raw = zpart; f.omap = ctx->d_idx_ident + q; f.ogroup = 256; f.out_J = 5;
                fs[q] = f; Ls[q] = &ctx->p2_b_split[q];
                t0 += ctx->p2_b_split[q].taps;
            }
            if (int rc = gconv_forward_grouped(ctx, Ls, fs, 5, st)) return rc;
            part2_reduce_kernel<<<(n * 64 + 255) / 256, 256, 0, st>>>(zpart, ctx->p2_b.bias, z1, ctx->hop2_zero_pos, z3, z3_hi, z3_lo, n);
            ctx->launches++;
        } else {
            if (int rc = gconv_forward(ctx, ctx->p2_b, a, st)) return rc;
            if (tc) { if (int rc = gconv_split_bf16(ctx, z3, z3_hi, z3_lo, (size_t)n * 256, st)) return rc; }
        }
        a.act_hi = a.act_lo = nullptr;
        // head: 256 -> 512 -> 128 with BN+ReLU, as 1-tap layers.  Tensor-core path: two dense GEMMs (the second zero-padded to one
        // 256-column tile, 128 valid), h2 rows 256 floats apart
        a.resid = nullptr; a.idx = ctx->d_idx_one; a.Jin = 1; a.Jout = 1;
        const bool tc_head = tc && ctx->p2_fc1.w_hi && ctx->p2_fc2_pad.w_hi;
        if (tc_head) {
            a.act = nullptr; a.act_hi = z3_hi; a.act_lo = z3_lo; a.out_raw = nullptr; a.out_act = nullptr; a.out_hi = h1_hi; a.out_lo = h1_lo;
            a.scale = ctx->p2_bn1.scale; a.shift = ctx->p2_bn1.shift;
            if (int rc = gconv_forward(ctx, ctx->p2_fc1, a, st)) return rc;
            a.act_hi = h1_hi; a.act_lo = h1_lo; a.out_hi = a.out_lo = nullptr; a.out_act = h2; a.n_valid = 128;
            a.scale = ctx->p2_bn2_pad.scale; a.shift = ctx->p2_bn2_pad.shift;
            if (int rc = gconv_forward(ctx, ctx->p2_fc2_pad, a, st)) return rc;
            a.n_valid = 0; a.act_hi = a.act_lo = nullptr;
        } else {
        a.act = z3; a.out_raw = nullptr; a.out_act = h1; a.scale = ctx->p2_bn1.scale; a.shift = ctx->p2_bn1.shift;
        if (int rc = gconv_forward(ctx, ctx->p2_fc1, a, st)) return rc;
        a.act = h1; a.out_act = h2; a.scale = ctx->p2_bn2.scale; a.shift = ctx->p2_bn2.shift;
        if (int rc = gconv_forward(ctx, ctx->p2_fc2, a, st)) return rc;
        }
        const double* k0 = kps0;
        const double* k1 = kps1;
        if (!pairs && trans) { k0 = kps0 + 3 * (size_t)s; k1 = kps1 + 3 * (size_t)s; }
        part2_head_kernel<<<(n + 3) / 4, 128, 0, st>>>(h2, tc_head ? 256 : 128, ctx->p2_fc3.w, ctx->p2_fc3.bias, pr, pre_idx + s, ctx->d_rot32,
                                                       k0, k1, quat + 4 * (size_t)s, trans ? trans + 12 * (size_t)s : nullptr, n);
        ctx->launches++;
    }
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

extern "C" int yoho_gather_kps(yoho_ctx* ctx, const double* kps0, const double* kps1, const int64_t* pairs, int M,
                               double* out0, double* out1, void* stream) {
    YARG(ctx && kps0 && kps1 && pairs && out0 && out1 && M >= 0);
    if (M == 0) return YOHO_OK;
    YCHECK(cudaSetDevice(ctx->device));
    gather_kps_kernel<<<(3 * M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(kps0, kps1, pairs, M, out0, out1);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
