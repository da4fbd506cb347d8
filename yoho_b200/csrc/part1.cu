// PartI descriptor network (utils/network.py:67-105,140-147) around the gather-GEMM layers.
//
//   x0[B,32,60] --transpose--> [B,60,32]
//   y1 = GC_in(x0)                       raw kept (shortcut), a1 = relu(BN_a(y1))           (:87-88, :55)
//   a2 = relu(BN_b(GC_a(a1)))                                                                  (:55-57)
//   a3 = relu(BN_o(GC_b(a2) + y1))                                                             (:57-65,:91)
//   y4 = GC_out(a3)                                                                            (:92)
//   e = y4 + x0 ; inv = normalise(mean_g e) ; eqv = e / max(||e||_c, 1e-4) ; desc = mean_g eqv (:98-103, tests/matcher.py:35)
#include <cuda_bf16.h>
#include "common.cuh"
#include "tcgen05.cuh"      // mbarrier / 1-D TMA bulk-copy helpers

// fourier_tc.cu / gconv_tc.cu / abi.cu
int group_transform_tc(yoho_ctx* ctx, const void* in_hi, const void* in_lo, int B, int C, const void* m1_hi, const void* m1_lo, const void* m2_hi,
                       const void* m2_lo, const float* bias, const float* resid, const float* scale, const float* shift,
                       void* out_hi, void* out_lo, cudaStream_t st, const void* in2_hi, const void* in2_lo);
int gconv_forward_grouped(yoho_ctx* ctx, const GLayer* const* Ls, const GConvArgs* as, int n, cudaStream_t st);
int group_finalize_tc(yoho_ctx* ctx, const void* y4_hi, const void* y4_lo, int B, const void* minv_hi, const void* minv_lo, const float* bias4,
                      const float* x, float* eqv, float* inv, float* desc, cudaStream_t st);
void yoho_prof_begin(yoho_ctx* ctx, int cls, double flops, cudaStream_t st);
void yoho_prof_end(yoho_ctx* ctx, cudaStream_t st);

namespace {

// [B][32][60] -> [B][60][32]; one CTA per keypoint.
// With hi/lo given the transposed tile is written as a bf16 split instead (tensor-core path).
__global__ void __launch_bounds__(128) transpose_in_kernel(const float* __restrict__ x, float* __restrict__ xt,
                                                          unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, int B) {
    __shared__ float s[YF][YG + 1];
    const int b = blockIdx.x;
    const float* src = x + (size_t)b * YF * YG;
    for (int i = threadIdx.x; i < YF * YG; i += blockDim.x) s[i / YG][i % YG] = src[i];
    __syncthreads();
    const size_t o = (size_t)b * YF * YG;
    for (int i = threadIdx.x; i < YF * YG; i += blockDim.x) {
        const float v = s[i % YF][i / YF];
        if (hi) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            hi[o + i] = __bfloat16_as_ushort(h);
            lo[o + i] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
        } else {
            xt[o + i] = v;
        }
    }
}

// numpy's float32 pairwise summation of 60 contiguous values followed by /60
// (np.mean(feats, axis=-1), tests/matcher.py:35): 8 strided partial sums, tree-combined, 4-element tail.
__device__ __forceinline__ float numpy_mean60(const float* a, int stride) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j * stride];
    for (int i = 8; i < 56; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[(i + j) * stride]);
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (int i = 56; i < 60; ++i) res = __fadd_rn(res, a[i * stride]);
    return __fdiv_rn(res, 60.0f);
}

// One CTA (64 threads) per keypoint.  x [B][32][60].  Layer-4 result either as y4 [B][60][32] (bias included) or, on
// the tensor-core path, as Z [B][60][512] = a3 . W_cat with the 13-tap gather still to do:
//     y4[g][c] = bias[c] + sum_k Z[N[g][k]][k*32 + c]      (taps in ascending order)
__global__ void __launch_bounds__(64) part1_finalize_kernel(const float* __restrict__ y4, const float* __restrict__ z,
                                                           const int* __restrict__ nei, const float* __restrict__ bias4,
                                                           const float* __restrict__ x,
                                                           float* __restrict__ eqv, float* __restrict__ inv,
                                                           float* __restrict__ desc, int B) {
    __shared__ float e[YF][YG + 1];
    __shared__ float invn[YG];
    __shared__ float red[YF];
    const int b = blockIdx.x;
    const int t = threadIdx.x;
    const float* xs = x + (size_t)b * YF * YG;
    if (z) {
        const float* zs = z + (size_t)b * YG * 512;
        for (int i = t; i < YF * YG; i += 64) {
            const int g = i / YF, c = i % YF;
            float acc = bias4[c];
#pragma unroll
            for (int k = 0; k < YT; ++k) acc += zs[(size_t)nei[g * YT + k] * 512 + k * 32 + c];
            e[c][g] = acc;
        }
    } else {
        const float* ys = y4 + (size_t)b * YG * YF;
        // e[c][g] = y4[g][c] + x[c][g]
        for (int i = t; i < YF * YG; i += 64) {
            const int g = i / YF, c = i % YF;
            e[c][g] = ys[i];
        }
    }
    __syncthreads();
    for (int i = t; i < YF * YG; i += 64) {
        const int c = i / YG, g = i % YG;
        e[c][g] = e[c][g] + xs[i];
    }
    __syncthreads();
    if (t < YG) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < YF; ++c) ss = fmaf(e[c][t], e[c][t], ss);
        invn[t] = fmaxf(sqrtf(ss), 1e-4f);   // torch.clamp_min(torch.norm(eqv, dim=1), 1e-4)
    }
    // invariant pooling uses the UN-normalised e (utils/network.py:99 precedes :102)
    float m = 0.f;
    if (t < YF) {
        float s = 0.f;
        for (int g = 0; g < YG; ++g) s += e[t][g];
        m = s / 60.0f;
        red[t] = m * m;
    }
    __syncthreads();
    if (t < YF && inv) {
        float ss = 0.f;
        for (int c = 0; c < YF; ++c) ss += red[c];
        inv[(size_t)b * YF + t] = m / fmaxf(sqrtf(ss), 1e-4f);
    }
    float* out = eqv + (size_t)b * YF * YG;
    for (int i = t; i < YF * YG; i += 64) {
        const int c = i / YG, g = i % YG;
        const float v = e[c][g] / invn[g];
        e[c][g] = v;
        out[i] = v;
    }
    __syncthreads();
    if (t < YF && desc) desc[(size_t)b * YF + t] = numpy_mean60(&e[t][0], 1);
}

// Input side of the all-Fourier PartI: X0[m][c] = sum_g F[m][g] x[c][g], written as the bf16 hi/lo pair [B][60][32] that the
// layer-1 per-irrep GEMMs consume.  CTAs of 128 threads stride over the keypoints with F^T resident in shared memory; thread
// (c = t % 32, quarter = t / 32) accumulates the 16 coefficient rows m = 16*quarter .. +15 of its channel in registers (one
// conflict-free scalar read of x and four broadcast 16-byte reads of F^T per 16 FMAs).
__global__ void __launch_bounds__(128) fourier_in_kernel(const float* __restrict__ x, const float* __restrict__ F,
                                                        unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, int B) {
    __shared__ float xs[YF][YG + 1];
    __shared__ __align__(16) float Ft[YG][64];           // Ft[g][m] = F[m][g], m padded to 64 with zeros
    const int t = threadIdx.x;
    for (int i = t; i < YG * 64; i += 128) {
        const int g = i >> 6, m = i & 63;
        Ft[g][m] = m < YG ? __ldg(F + m * YG + g) : 0.f;
    }
    const int c = t & 31, m0 = (t >> 5) * 16;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const float* src = x + (size_t)b * YF * YG;
        __syncthreads();                                 // previous keypoint's reads of xs are done (and Ft is complete)
        for (int i = t; i < YF * YG; i += 128) xs[i / YG][i % YG] = src[i];
        __syncthreads();
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll 4
        for (int g = 0; g < YG; ++g) {
            const float xv = xs[c][g];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                const float4 f = *reinterpret_cast<const float4*>(&Ft[g][m0 + 4 * i4]);
                acc[4 * i4 + 0] = fmaf(f.x, xv, acc[4 * i4 + 0]); acc[4 * i4 + 1] = fmaf(f.y, xv, acc[4 * i4 + 1]);
                acc[4 * i4 + 2] = fmaf(f.z, xv, acc[4 * i4 + 2]); acc[4 * i4 + 3] = fmaf(f.w, xv, acc[4 * i4 + 3]);
            }
        }
        const size_t o = (size_t)b * YG * YF + c;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (m0 + i < YG) {
                const __nv_bfloat16 h = __float2bfloat16_rn(acc[i]);
                hi[o + (m0 + i) * YF] = __bfloat16_as_ushort(h);
                lo[o + (m0 + i) * YF] = __bfloat16_as_ushort(__float2bfloat16_rn(acc[i] - __bfloat162float(h)));
            }
        }
    }
}

// Output side of the all-Fourier PartI: the layer-4 result arrives as Fourier coefficients Y4 [B][60 m][32 c] (no bias);
//     e[c][g] = bias4[c] + sum_m F[m][g] Y4[m][c] + x[c][g]
// followed by the same tail as part1_finalize_kernel (both L2 norms, invariant pool, numpy's pairwise mean_g).
// CTAs of 128 threads stride over the keypoints with F resident in shared memory; thread (g = t % 64 < 60, half = t / 64)
// accumulates 16 channels of its group element in registers.
__global__ void __launch_bounds__(128) part1_finalize_fourier_kernel(const float* __restrict__ y4f, const float* __restrict__ F,
                                                                    const float* __restrict__ bias4, const float* __restrict__ x,
                                                                    float* __restrict__ eqv, float* __restrict__ inv,
                                                                    float* __restrict__ desc, int B) {
    __shared__ float Fs[YG][64];                         // Fs[m][g], g padded to 64 (threads g >= 60 read zeros)
    __shared__ __align__(16) float ys[YG][YF];
    __shared__ float e[YF][YG + 1];
    __shared__ float invn[YG];
    __shared__ float red[YF];
    const int t = threadIdx.x;
    for (int i = t; i < YG * 64; i += 128) {
        const int m = i >> 6, g = i & 63;
        Fs[m][g] = g < YG ? __ldg(F + m * YG + g) : 0.f;
    }
    const int g = t & 63, c0 = (t >> 6) * 16;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const float* xs = x + (size_t)b * YF * YG;
        const float* yb = y4f + (size_t)b * YG * YF;
        __syncthreads();                                 // previous keypoint done with ys / e / red (and Fs is complete)
        for (int i = t; i < YG * YF / 4; i += 128) reinterpret_cast<float4*>(&ys[0][0])[i] = reinterpret_cast<const float4*>(yb)[i];
        __syncthreads();
        {
            float acc[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[c] = __ldg(bias4 + c0 + c);
#pragma unroll 4
            for (int m = 0; m < YG; ++m) {
                const float f = Fs[m][g];
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 y = *reinterpret_cast<const float4*>(&ys[m][c0 + c4 * 4]);
                    acc[c4 * 4 + 0] = fmaf(f, y.x, acc[c4 * 4 + 0]); acc[c4 * 4 + 1] = fmaf(f, y.y, acc[c4 * 4 + 1]);
                    acc[c4 * 4 + 2] = fmaf(f, y.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(f, y.w, acc[c4 * 4 + 3]);
                }
            }
            if (g < YG) {
#pragma unroll
                for (int c = 0; c < 16; ++c) e[c0 + c][g] = acc[c];
            }
        }
        __syncthreads();
        for (int i = t; i < YF * YG; i += 128) {
            const int c = i / YG, gg = i % YG;
            e[c][gg] = e[c][gg] + xs[i];
        }
        __syncthreads();
        if (t < YG) {
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < YF; ++c) ss = fmaf(e[c][t], e[c][t], ss);
            invn[t] = fmaxf(sqrtf(ss), 1e-4f);   // torch.clamp_min(torch.norm(eqv, dim=1), 1e-4)
        }
        // invariant pooling uses the UN-normalised e (utils/network.py:99 precedes :102)
        float m = 0.f;
        if (t >= 64 && t < 64 + YF) {
            float sum = 0.f;
            for (int gg = 0; gg < YG; ++gg) sum += e[t - 64][gg];
            m = sum / 60.0f;
            red[t - 64] = m * m;
        }
        __syncthreads();
        if (t >= 64 && t < 64 + YF && inv) {
            float ss = 0.f;
            for (int c = 0; c < YF; ++c) ss += red[c];
            inv[(size_t)b * YF + t - 64] = m / fmaxf(sqrtf(ss), 1e-4f);
        }
        float* out = eqv + (size_t)b * YF * YG;
        for (int i = t; i < YF * YG; i += 128) {
            const int c = i / YG, gg = i % YG;
            const float v = e[c][gg] / invn[gg];
            e[c][gg] = v;
            out[i] = v;
        }
        __syncthreads();
        if (t < YF && desc) desc[(size_t)b * YF + t] = numpy_mean60(&e[t][0], 1);
    }
}

// ---- warp-MMA input / output side (default) -------------------------------------------------------------------------------------
// Both sides are 64 x 64 x 32 products per keypoint against the CONSTANT transform matrix, far too small for a tcgen05 tile but
// shared-memory-wavefront bound as SIMT code (61 / 84 us per 5000 keypoints against a 12 us HBM floor).  Here the matrix lives in
// registers as mma.sync A fragments (bf16 hi and lo, one 16-row tile per warp) for the whole kernel, the per-keypoint operand is
// read from global memory straight into B fragments (FP32 -> bf16 hi/lo split in registers: no shared-memory staging at all),
// and D = A_hi B_hi + A_lo B_hi + A_hi B_lo accumulates in FP32 (the same three-product scheme as the tensor-core GEMMs).
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {           // low half = a
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    hi = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
    lo = pack_bf16x2(a - __bfloat162float(ha), b - __bfloat162float(hb));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A fragments (hi, lo) of the 16-row tile `rt` of the 64 x 64 (zero-padded) matrix M[r][k] = transposed ? F[k][r] : F[r][k].
__device__ __forceinline__ void load_matrix_frags(const float* __restrict__ F, bool transposed, int rt, int lane,
                                                  uint32_t (&ah)[4][4], uint32_t (&al)[4][4]) {
    const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = 16 * rt + gid + ((i & 1) ? 8 : 0);
            const int k = 16 * t + 2 * tig + ((i & 2) ? 8 : 0);
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int kk = k + e;
                v[e] = (r < YG && kk < YG) ? __ldg(F + (transposed ? kk * YG + r : r * YG + kk)) : 0.f;
            }
            split2(v[0], v[1], ah[t][i], al[t][i]);
        }
}

__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {     // packed conversions (cvt.rn.bf16x2.f32)
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// X0[b][m][c] (bf16 hi / lo) = sum_g F[m][g] x[b][c][g].  CTA = 4 warps = the four 16-row tiles of m; CTAs stride over keypoints.
//  * the 7 680-byte [32][60] tile of a keypoint arrives by ONE 1-D TMA bulk copy into a ring of IN_ST shared-memory stages, issued
//    IN_ST keypoints ahead by one thread (mbarrier complete_tx): the register-fed first version (B fragments straight from global
//    memory) was latency bound at 52 us;
//  * the CTA converts the tile ONCE to packed bf16 hi / lo pairs (g even | g odd = exactly a B-fragment register) in a padded
//    shared-memory image, and the four warps read their fragments from it with conflict-free 32-bit loads: with every warp
//    splitting the whole tile itself (second version, 36 us) the kernel was bound by the FP32 -> bf16 conversion pipe, 4x redundantly.
constexpr int IN_ST = 4;
constexpr int CV_PITCH = 36;                  // 32 (g-pair) columns + 4: bank = 4 c + pair -> the 8 x 4 lanes of a fragment load hit 32 banks
__global__ void __launch_bounds__(128) fourier_in_mma_kernel(const float* __restrict__ x, const float* __restrict__ F,
                                                            unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, int B) {
    __shared__ __align__(128) float xs[IN_ST][YF * YG];
    __shared__ uint32_t cvh[YF * CV_PITCH], cvl[YF * CV_PITCH];
    __shared__ __align__(8) unsigned long long full[IN_ST];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const int n_mine = blockIdx.x < B ? (B - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < IN_ST; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
        for (int s = 0; s < IN_ST && s < n_mine; ++s) {
            mbar_expect_tx(&full[s], YF * YG * 4);
            bulk_g2s(&xs[s][0], x + (size_t)(blockIdx.x + (size_t)s * gridDim.x) * YF * YG, YF * YG * 4, &full[s]);
        }
    }
    for (int i = threadIdx.x; i < YF * CV_PITCH; i += 128) { cvh[i] = 0u; cvl[i] = 0u; }     // pair columns 30, 31 (g = 60..63) stay zero
    uint32_t ah[4][4], al[4][4];
    load_matrix_frags(F, false, w, lane, ah, al);
    __syncthreads();                                               // barrier initialisation + zeroed padding visible
    for (int it = 0; it < n_mine; ++it) {
        const int b = blockIdx.x + it * gridDim.x, s = it % IN_ST;
        mbar_wait(&full[s], (it / IN_ST) & 1);
        // cooperative split of the tile: 960 (channel, g-pair) items, 7.5 per thread
        for (int i = threadIdx.x; i < YF * (YG / 2); i += 128) {
            const int c = i / (YG / 2), pr = i - c * (YG / 2);
            const float2 v = *reinterpret_cast<const float2*>(&xs[s][c * YG + 2 * pr]);
            split_pack(v.x, v.y, cvh[c * CV_PITCH + pr], cvl[c * CV_PITCH + pr]);
        }
        __syncthreads();                                           // image complete; stage s is free again
        if (threadIdx.x == 0 && it + IN_ST < n_mine) {
            mbar_expect_tx(&full[s], YF * YG * 4);
            bulk_g2s(&xs[s][0], x + (size_t)(b + (size_t)IN_ST * gridDim.x) * YF * YG, YF * YG * 4, &full[s]);
        }
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint32_t bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {                          // column n = gid of n-tile j is channel 8j + gid
                const int o = (8 * j + gid) * CV_PITCH + 8 * t + tig;
                bh0[j] = cvh[o]; bl0[j] = cvl[o]; bh1[j] = cvh[o + 4]; bl1[j] = cvl[o + 4];
            }
            // the n-tile is the INNER loop: consecutive mma.sync instructions belong to four independent accumulator chains
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_bf16(acc[j], ah[t], bh0[j], bh1[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_bf16(acc[j], al[t], bh0[j], bh1[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_bf16(acc[j], ah[t], bl0[j], bl1[j]);
        }
        __syncthreads();                                           // every warp has read the image: the next keypoint may overwrite it
        // thread: rows m = 16w + gid (+8), channels 8j + 2 tig + {0,1}: one bf16 pair (4 bytes) per n-tile, hi and lo
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = 16 * w + gid + 8 * h;
            if (m < YG) {
                const size_t o = ((size_t)b * YG + m) * YF + 2 * tig;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t ph, pl;
                    split_pack(acc[j][2 * h], acc[j][2 * h + 1], ph, pl);
                    *reinterpret_cast<uint32_t*>(hi + o + 8 * j) = ph;
                    *reinterpret_cast<uint32_t*>(lo + o + 8 * j) = pl;
                }
            }
        }
    }
}

// Output side: e[c][g] = bias4[c] + sum_m F[m][g] Y4[b][m][c] + x[b][c][g], then the tail of PartI_network.forward
// (utils/network.py:98-103) and the matcher's numpy mean (tests/matcher.py:35).  Warp w owns the group elements 16w .. 16w+15.
// Both tiles of a keypoint (Y4 coefficients and x) are prefetched by two TMA bulk copies into a two-stage ring; the Y4 tile is
// converted once per CTA to packed bf16 hi / lo pairs along m (= B-fragment registers) in a 40-column image (conflict-free reads).
constexpr int FIN_ST = 2;
constexpr int FV_PITCH = 40;
__global__ void __launch_bounds__(128) part1_finalize_mma_kernel(const float* __restrict__ y4f, const float* __restrict__ F,
                                                                const float* __restrict__ bias4, const float* __restrict__ x,
                                                                float* __restrict__ eqv, float* __restrict__ inv,
                                                                float* __restrict__ desc, int B) {
    __shared__ __align__(128) float ys[FIN_ST][YG * YF];
    __shared__ __align__(128) float xs[FIN_ST][YF * YG];
    __shared__ uint32_t cvh[(YG / 2) * FV_PITCH], cvl[(YG / 2) * FV_PITCH];   // [m-pair][channel]; pairs 30, 31 (m = 60..63) read as zero
    __shared__ float es[YF][YG + 1];
    __shared__ float part[4][YF];
    __shared__ __align__(8) unsigned long long full[FIN_ST];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, gid = lane >> 2, tig = lane & 3;
    const int n_mine = blockIdx.x < B ? (B - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    auto prefetch = [&](int it) {                                  // one thread
        const int s = it % FIN_ST;
        const size_t b = blockIdx.x + (size_t)it * gridDim.x;
        mbar_expect_tx(&full[s], 2 * YF * YG * 4);
        bulk_g2s(&xs[s][0], x + b * YF * YG, YF * YG * 4, &full[s]);
        bulk_g2s(&ys[s][0], y4f + b * YG * YF, YG * YF * 4, &full[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < FIN_ST; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
        for (int it = 0; it < FIN_ST && it < n_mine; ++it) prefetch(it);
    }
    uint32_t ah[4][4], al[4][4];
    load_matrix_frags(F, true, w, lane, ah, al);                  // A[g][m] = F[m][g]
    float bias[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { bias[j][0] = __ldg(bias4 + 8 * j + 2 * tig); bias[j][1] = __ldg(bias4 + 8 * j + 2 * tig + 1); }
    __syncthreads();
    for (int it = 0; it < n_mine; ++it) {
        const size_t b = blockIdx.x + (size_t)it * gridDim.x;
        const int s = it % FIN_ST;
        mbar_wait(&full[s], (it / FIN_ST) & 1);
        for (int i = threadIdx.x; i < (YG / 2) * YF; i += 128) {   // 960 (m-pair, channel) items
            const int pr = i / YF, c = i - pr * YF;
            split_pack(ys[s][(2 * pr) * YF + c], ys[s][(2 * pr + 1) * YF + c], cvh[pr * FV_PITCH + c], cvl[pr * FV_PITCH + c]);
        }
        __syncthreads();                                           // image complete (and the previous keypoint's es / part consumed)
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint32_t bh0[4], bl0[4], bh1[4], bl1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {                          // column n = gid of n-tile j is channel 8j + gid; rows m = 16t + 2 tig (+8)
                const int o = (8 * t + tig) * FV_PITCH + 8 * j + gid;
                const bool ok1 = 8 * t + tig + 4 < YG / 2;
                bh0[j] = cvh[o]; bl0[j] = cvl[o];
                bh1[j] = ok1 ? cvh[o + 4 * FV_PITCH] : 0u; bl1[j] = ok1 ? cvl[o + 4 * FV_PITCH] : 0u;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_bf16(acc[j], ah[t], bh0[j], bh1[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_bf16(acc[j], al[t], bh0[j], bh1[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_bf16(acc[j], ah[t], bl0[j], bl1[j]);
        }
        // thread: group elements g = 16w + gid (+8), channels c = 8j + 2 tig + {0,1}
        const float* xb = &xs[s][0];
        float e[2][8], ss[2] = {0.f, 0.f}, cs[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int g = 16 * w + gid + 8 * h;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int c = 8 * j + 2 * tig + i;
                    const float v = g < YG ? (acc[j][2 * h + i] + bias[j][i]) + xb[c * YG + g] : 0.f;
                    e[h][2 * j + i] = v;
                    ss[h] = fmaf(v, v, ss[h]);
                }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) cs[k] = e[0][k] + e[1][k];    // invariant pooling uses the UN-normalised e (utils/network.py:99 precedes :102)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            ss[h] += __shfl_xor_sync(0xffffffffu, ss[h], 1);
            ss[h] += __shfl_xor_sync(0xffffffffu, ss[h], 2);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 4);
            cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 8);
            cs[k] += __shfl_xor_sync(0xffffffffu, cs[k], 16);
        }
        if (gid == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) part[w][8 * (k >> 1) + 2 * tig + (k & 1)] = cs[k];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int g = 16 * w + gid + 8 * h;
            if (g < YG) {
                const float nrm = fmaxf(sqrtf(ss[h]), 1e-4f);      // torch.clamp_min(torch.norm(eqv, dim=1), 1e-4)
#pragma unroll
                for (int k = 0; k < 8; ++k) es[8 * (k >> 1) + 2 * tig + (k & 1)][g] = e[h][k] / nrm;
            }
        }
        __syncthreads();                                           // es / part complete; everyone is done with stage s and the image
        if (threadIdx.x == 0 && it + FIN_ST < n_mine) prefetch(it + FIN_ST);
        float* out = eqv + b * YF * YG;
        for (int i = threadIdx.x; i < YF * YG; i += 128) out[i] = es[i / YG][i % YG];
        if (w == 2) {
            if (desc) desc[b * YF + lane] = numpy_mean60(&es[lane][0], 1);
        } else if (w == 1 && inv) {
            const float m = (((part[0][lane] + part[1][lane]) + part[2][lane]) + part[3][lane]) / 60.0f;
            float q = m * m;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            inv[b * YF + lane] = m / fmaxf(sqrtf(q), 1e-4f);
        }
    }
}

__global__ void __launch_bounds__(64) group_mean_kernel(const float* __restrict__ eqv, float* __restrict__ desc, int K) {
    __shared__ float e[YF][YG + 1];
    const int b = blockIdx.x;
    const float* src = eqv + (size_t)b * YF * YG;
    for (int i = threadIdx.x; i < YF * YG; i += 64) e[i / YG][i % YG] = src[i];
    __syncthreads();
    if (threadIdx.x < YF) desc[(size_t)b * YF + threadIdx.x] = numpy_mean60(&e[threadIdx.x][0], 1);
}

constexpr int P1_CHUNK = 8192;  // keypoints per pass (3.6 GB of intermediates); a 5000-keypoint fragment is one pass, which keeps
                                // the persistent tensor-core kernels at 16-32 full waves of tiles instead of 6-7

}  // namespace

extern "C" int yoho_part1_forward(yoho_ctx* ctx, const float* x, int B, float* eqv, float* inv, float* desc_mean,
                                  void* stream) {
    YARG(ctx && x && eqv && B >= 0);
    if (!ctx->has_p1) {
        yoho_set_error("No model exists: yoho_part1_load has not been called");
        return YOHO_ERR_NOWEIGHTS;
    }
    cudaStream_t st = (cudaStream_t)stream;
    YCHECK(cudaSetDevice(ctx->device));
    const bool tc_on = ctx->gconv_impl >= 1 && ctx->p1_in.w_hi && ctx->p1_a.w_hi && ctx->p1_b.w_hi && ctx->p1_out.w_hi;
    // per keypoint: xt 32, y1 256, a1 256 (fp32, or bf16 hi+lo = same bytes), a2 512 (same), a3 256, y4 32 floats x 60
    const bool fourier_on = tc_on && ctx->gconv_impl == 3 && ctx->has_p1f;
    // group-Fourier path: X0 32, Y1 256, X1 256, Y2 512, X2 512, Y3 256, X3 256 (bf16 hi|lo pairs = fp32 bytes), Y4 32
    const size_t direct_f = 32 + 256 + 256 + 512 + 256 + 512, fourier_f = 32 + 256 + 256 + 512 + 512 + 256 + 256 + 32;
    const size_t per_kp = (size_t)YG * (fourier_on && fourier_f > direct_f ? fourier_f : direct_f) * sizeof(float);
    const int n_chunks = (B + P1_CHUNK - 1) / P1_CHUNK;
    const int chunk = n_chunks > 0 ? (B + n_chunks - 1) / n_chunks : 0;      // balanced passes
    if (int rc = yoho_ws_reserve(ctx, per_kp * (size_t)chunk)) return rc;
    for (int s = 0; s < B; s += chunk) {
        const int n = (B - s) < chunk ? (B - s) : chunk;
        const bool tc = tc_on && n * YG >= 128;   // a tensor-core tile is 128 rows
        float* xt = (float*)ctx->ws;
        float* y1 = xt + (size_t)n * YG * 32;
        float* a1 = y1 + (size_t)n * YG * 256;
        float* a2 = a1 + (size_t)n * YG * 256;
        float* a3 = a2 + (size_t)n * YG * 512;
        float* y4 = a3 + (size_t)n * YG * 256;
        // tensor-core path: the same regions hold bf16 hi|lo halves instead of fp32
        unsigned short* xt_hi = (unsigned short*)xt;
        unsigned short* xt_lo = xt_hi + (size_t)n * YG * 32;
        unsigned short* a1_hi = (unsigned short*)a1;
        unsigned short* a1_lo = a1_hi + (size_t)n * YG * 256;
        unsigned short* a2_hi = (unsigned short*)a2;
        unsigned short* a2_lo = a2_hi + (size_t)n * YG * 512;
        unsigned short* a3_hi = (unsigned short*)a3;
        unsigned short* a3_lo = a3_hi + (size_t)n * YG * 256;
        const float* xs = x + (size_t)s * YF * YG;
        // ---- the whole stack in the group-Fourier domain: per-irrep GEMMs for all four layers, transforms only
        // at the three BatchNorm/ReLU points, shortcut added as Fourier coefficients.  Needs the tcgen05 transform kernel.
        if (fourier_on && ctx->has_p1f_io && n >= 128 && (ctx->tc_flags & 256) && !(ctx->tc_flags & 512)) {
            const size_t R = (size_t)n * YG;
            unsigned short* w16 = (unsigned short*)ctx->ws;
            unsigned short* X0h = w16;            unsigned short* X0l = X0h + R * 32;
            unsigned short* Y1h = X0l + R * 32;   unsigned short* Y1l = Y1h + R * 256;
            unsigned short* X1h = Y1l + R * 256;  unsigned short* X1l = X1h + R * 256;
            unsigned short* Y2h = X1l + R * 256;  unsigned short* Y2l = Y2h + R * 512;
            unsigned short* X2h = Y2l + R * 512;  unsigned short* X2l = X2h + R * 512;
            unsigned short* Y3h = X2l + R * 512;  unsigned short* Y3l = Y3h + R * 256;
            unsigned short* X3h = Y3l + R * 256;  unsigned short* X3l = X3h + R * 256;
            float* Y4 = (float*)(X3l + R * 256);
            const bool simt_io = (ctx->tc_flags & 8192) != 0;                    // test twin: FP32 SIMT input / output side
            const int sgrid = n < 8 * ctx->num_sms ? n : 8 * ctx->num_sms;      // keypoint-striding CTAs, transform matrix resident on chip
            // the warp-MMA kernels are persistent: exactly one wave of resident CTAs (the TMA ring of a CTA prefetches ITS next keypoints)
            static int occ_in = 0, occ_fin = 0;
            if (!occ_in) {
                YCHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_in, fourier_in_mma_kernel, 128, 0));
                YCHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_fin, part1_finalize_mma_kernel, 128, 0));
                if (occ_in < 1) occ_in = 1;
                if (occ_fin < 1) occ_fin = 1;
            }
            const int grid_in = n < occ_in * ctx->num_sms ? n : occ_in * ctx->num_sms;
            const int grid_fin = n < occ_fin * ctx->num_sms ? n : occ_fin * ctx->num_sms;
            if (simt_io) fourier_in_kernel<<<sgrid, 128, 0, st>>>(xs, ctx->d_F, X0h, X0l, n);
            else fourier_in_mma_kernel<<<grid_in, 128, 0, st>>>(xs, ctx->d_F, X0h, X0l, n);
            ctx->launches++;
            GConvArgs f{};
            f.B = n; f.Jin = YG; f.out_J = YG;
            GConvArgs fs[8];
            const GLayer* Ls[8];
            auto layer = [&](const GLayer* Lr, const unsigned short* ih, const unsigned short* il, unsigned short* oh, unsigned short* ol,
                             float* oraw, int ogroup, bool out4) -> int {
                for (int q = 0; q < ctx->nf; ++q) {
                    const int r = ctx->nf - 1 - q;                           // largest irrep first: short tiles make up the tail
                    f.act_hi = ih; f.act_lo = il; f.idx = ctx->d_fidx[r]; f.Jout = ctx->fd[r];
                    f.omap = out4 ? ctx->d_fomap_out[r] : ctx->d_fomap[r]; f.ogroup = ogroup;
                    f.out_raw = oraw; f.out_hi = oh; f.out_lo = ol;
                    f.n_valid = out4 ? ctx->fd[r] * 32 : 0;
                    fs[q] = f; Ls[q] = &Lr[r];
                }
                return gconv_forward_grouped(ctx, Ls, fs, ctx->nf, st);
            };
            // layer 1 (32 -> 256): Y1 = Fourier coefficients of y1 (without bias: it joins in the group domain below)
            if (int rc = layer(ctx->p1f_in, X0h, X0l, Y1h, Y1l, nullptr, 256, false)) return rc;
            // X1 = F relu(BN_a(F^T Y1 + b1))
            yoho_prof_begin(ctx, 8, 2.0 * n * 256 * 7200.0, st);
            if (int rc = group_transform_tc(ctx, Y1h, Y1l, n, 256, ctx->d_inv_hi, ctx->d_inv_lo, ctx->d_fwd_hi, ctx->d_fwd_lo, ctx->p1_in.bias, nullptr,
                                            ctx->p1_bn_a.scale, ctx->p1_bn_a.shift, X1h, X1l, st, nullptr, nullptr)) return rc;
            yoho_prof_end(ctx, st);
            if (int rc = layer(ctx->p1f_a, X1h, X1l, Y2h, Y2l, nullptr, 512, false)) return rc;
            yoho_prof_begin(ctx, 8, 2.0 * n * 512 * 7200.0, st);
            if (int rc = group_transform_tc(ctx, Y2h, Y2l, n, 512, ctx->d_inv_hi, ctx->d_inv_lo, ctx->d_fwd_hi, ctx->d_fwd_lo, ctx->p1_a.bias, nullptr,
                                            ctx->p1_bn_b.scale, ctx->p1_bn_b.shift, X2h, X2l, st, nullptr, nullptr)) return rc;
            yoho_prof_end(ctx, st);
            if (int rc = layer(ctx->p1f_b, X2h, X2l, Y3h, Y3l, nullptr, 256, false)) return rc;
            // X3 = F relu(BN_o(F^T (Y3 + Y1) + b3 + b1)): the identity shortcut y1 = F^T Y1 + b1 rides along as coefficients
            yoho_prof_begin(ctx, 8, 2.0 * n * 256 * 7200.0, st);
            if (int rc = group_transform_tc(ctx, Y3h, Y3l, n, 256, ctx->d_inv_hi, ctx->d_inv_lo, ctx->d_fwd_hi, ctx->d_fwd_lo, ctx->d_p1_bias31, nullptr,
                                            ctx->p1_bn_out.scale, ctx->p1_bn_out.shift, X3h, X3l, st, Y1h, Y1l)) return rc;
            yoho_prof_end(ctx, st);
            // layer 4 (256 -> 32): Fourier coefficients of y4 [n][60][32], then the output side (inverse transform, residual, norms,
            // pools).  Tuning flag 2048: coefficients as a bf16 hi/lo pair and the output side on tensor cores (fourier_tc.cu).
            if (ctx->tc_flags & 2048) {
                unsigned short* Y4h = (unsigned short*)Y4;
                unsigned short* Y4l = Y4h + R * 32;
                if (int rc = layer(ctx->p1f_out, X3h, X3l, Y4h, Y4l, nullptr, 32, true)) return rc;
                if (int rc = group_finalize_tc(ctx, Y4h, Y4l, n, ctx->d_inv_hi, ctx->d_inv_lo, ctx->p1_out.bias, xs, eqv + (size_t)s * YF * YG,
                                               inv ? inv + (size_t)s * YF : nullptr, desc_mean ? desc_mean + (size_t)s * YF : nullptr, st)) return rc;
            } else {
            if (int rc = layer(ctx->p1f_out, X3h, X3l, nullptr, nullptr, Y4, 32, true)) return rc;
            const int fgrid = n < 6 * ctx->num_sms ? n : 6 * ctx->num_sms;
            if (simt_io)
                part1_finalize_fourier_kernel<<<fgrid, 128, 0, st>>>(Y4, ctx->d_F, ctx->p1_out.bias, xs, eqv + (size_t)s * YF * YG,
                                                                      inv ? inv + (size_t)s * YF : nullptr,
                                                                      desc_mean ? desc_mean + (size_t)s * YF : nullptr, n);
            else
                part1_finalize_mma_kernel<<<grid_fin, 128, 0, st>>>(Y4, ctx->d_F, ctx->p1_out.bias, xs, eqv + (size_t)s * YF * YG,
                                                                 inv ? inv + (size_t)s * YF : nullptr,
                                                                 desc_mean ? desc_mean + (size_t)s * YF : nullptr, n);
            ctx->launches++;
            }
            continue;
        }
        transpose_in_kernel<<<n, 128, 0, st>>>(xs, xt, tc ? xt_hi : nullptr, tc ? xt_lo : nullptr, n);
        ctx->launches++;
        GConvArgs a{};
        a.idx = ctx->d_idx_full; a.B = n; a.Jin = YG; a.Jout = YG;
        // layer 1: raw y1 (shortcut) + a1 = relu(BN_a(y1))
        a.resid = nullptr; a.out_raw = y1;
        a.scale = ctx->p1_bn_a.scale; a.shift = ctx->p1_bn_a.shift;
        if (tc) { a.act_hi = xt_hi; a.act_lo = xt_lo; a.out_hi = a1_hi; a.out_lo = a1_lo; } else { a.act = xt; a.out_act = a1; }
        if (int rc = gconv_forward(ctx, ctx->p1_in, a, st)) return rc;
        a.out_act = nullptr;
        // layer 2: a2 = relu(BN_b(GC_a(a1)))
        a.out_raw = nullptr;
        a.scale = ctx->p1_bn_b.scale; a.shift = ctx->p1_bn_b.shift;
        if (tc) { a.act_hi = a1_hi; a.act_lo = a1_lo; a.out_hi = a2_hi; a.out_lo = a2_lo; } else { a.act = a1; a.out_act = a2; }
        if (int rc = gconv_forward(ctx, ctx->p1_a, a, st)) return rc;
        // layer 3: a3 = relu(BN_o(GC_b(a2) + y1))
        a.resid = y1; a.Jres = YG; a.resid_off = 0; a.resid_per_j = 1;
        a.scale = ctx->p1_bn_out.scale; a.shift = ctx->p1_bn_out.shift;
        if (tc) { a.act_hi = a2_hi; a.act_lo = a2_lo; a.out_hi = a3_hi; a.out_lo = a3_lo; } else { a.act = a2; a.out_act = a3; }
        if (int rc = gconv_forward(ctx, ctx->p1_b, a, st)) return rc;
        // layer 4: y4 = GC_out(a3).  Tensor-core path: one dense GEMM Z = a3 . W_cat [256 x 13*32] (a3 is read once, not
        // 13 times); the 13-tap gather-add happens on the 100 KB Z tile of each keypoint inside the finalize kernel.
        a.resid = nullptr; a.out_raw = y4; a.out_act = nullptr; a.out_hi = a.out_lo = nullptr; a.scale = a.shift = nullptr;
        const bool dense4 = tc && ctx->p1_out_cat.w_hi;
        if (dense4) {
            a.act_hi = a3_hi; a.act_lo = a3_lo; a.idx = ctx->d_idx_ident; a.n_valid = YT * 32;
            if (int rc = gconv_forward(ctx, ctx->p1_out_cat, a, st)) return rc;
        } else {
            if (tc) { a.act_hi = a3_hi; a.act_lo = a3_lo; } else { a.act = a3; }
            if (int rc = gconv_forward(ctx, ctx->p1_out, a, st)) return rc;
        }
        part1_finalize_kernel<<<n, 64, 0, st>>>(dense4 ? nullptr : y4, dense4 ? y4 : nullptr, ctx->d_idx_full, ctx->p1_out.bias, xs,
                                                eqv + (size_t)s * YF * YG, inv ? inv + (size_t)s * YF : nullptr,
                                                desc_mean ? desc_mean + (size_t)s * YF : nullptr, n);
        ctx->launches++;
    }
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

extern "C" int yoho_group_mean(yoho_ctx* ctx, const float* eqv, int K, float* desc, void* stream) {
    YARG(ctx && eqv && desc && K >= 0);
    if (K == 0) return YOHO_OK;
    YCHECK(cudaSetDevice(ctx->device));
    group_mean_kernel<<<K, 64, 0, (cudaStream_t)stream>>>(eqv, desc, K);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
