// Mutual 1-NN on tensor cores with exact FP32 verification (tests/matcher.py:37-48, utils/knn_search.py:17-66).
//
// The reference's answer is an ARGMIN over FP32 values dist = sqrt(sum_f (a_f - b_f)^2 + 1e-7) with first-occurrence ties; the
// SIMT kernel (match.cu, nn_tile_kernel) evaluates all Ka x Kb of them in that arithmetic: 0.8 G subtract+FMA pairs, issue bound
// at ~98 us for 5000 x 5000.  Here the K x K work moves to tcgen05:
//
//   1. nn_prep_kernel   descriptors -> bf16 hi / lo images in the UMMA K-major SWIZZLE_64B layout (a row = 32 channels = 64 B),
//                       squared norms, the largest norm.
//   2. nn_tc_kernel     per 128-row tile: dot = A_hi B_hi + A_lo B_hi + A_hi B_lo (M128 x N256 x K32, six tcgen05.mma per column
//                       tile, FP32 in TMEM); the epilogue thread that owns a row (one TMEM lane) turns every dot product into the
//                       approximate squared distance |a|^2 + |b|^2 - 2 dot and keeps its FOUR smallest.  The approximation error
//                       is bounded (3-product bf16 split: 2^-15 |a||b|, plus FP32 rounding of the norms) by
//                       tol = 1e-4 (|a|^2 + max|b|^2), so the true argmin — and every column that ties with it after the
//                       reference's rounding — lies within 2 tol of the smallest approximate value.  Those (normally one or two)
//                       candidates are re-evaluated in the reference's exact arithmetic (one subtraction and one FMA per channel
//                       in ascending channel order, IEEE sqrt) and merged as 64-bit (distance bits, index) keys with atomicMin,
//                       exactly the keys the SIMT kernel produces.  Rows and columns swap roles in a second set of CTAs
//                       (blockIdx.y), which gives the column minima without any cross-lane reduction.
//   3. nn_fix_kernel    rows whose four candidates ALL fell inside the window (many near-duplicates) are flagged and re-scanned
//                       exhaustively in the exact arithmetic, one warp per row; normally no row is flagged.
//
// The result is bit-identical to nn_tile_kernel's (tests/test_gpu_parity.py, test_gpu_fullsize.py run both).
#include <cuda_bf16.h>
#include <math_constants.h>
#include "common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int NT_M = 128;            // rows per CTA (UMMA M)
constexpr int NT_N = 256;            // columns per MMA tile (UMMA N)
constexpr int ROW_B = 64;            // bytes per image row (32 bf16)
constexpr int TOPK = 4;
constexpr float TOL_REL = 1e-4f;

// K-major operand tile, 64-byte rows, SWIZZLE_64B, 8-row groups 512 B apart (same image as gconv_tc.cu's operands).
__device__ __forceinline__ uint64_t desc_sw64(const void* smem_tile) {
    const uint64_t addr = (uint64_t)(smem_u32(smem_tile) >> 4) & 0x3FFFull;
    return addr | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}

// One thread per row of the padded image (rows >= K are zero, their norm +inf so that they never become a candidate).
__global__ void nn_prep_kernel(const float* __restrict__ d, int K, int Kpad, uint8_t* __restrict__ img_hi, uint8_t* __restrict__ img_lo,
                               float* __restrict__ norm, unsigned int* __restrict__ nmax_bits) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Kpad) return;
    float v[YF];
    float n2 = 0.f;
#pragma unroll
    for (int q = 0; q < YF / 4; ++q) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < K) x = reinterpret_cast<const float4*>(d + (size_t)r * YF)[q];
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        n2 = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, n2))));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {                              // 16-byte chunk j = channels 8j .. 8j+7
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
            const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
            h[e] = *reinterpret_cast<const uint32_t*>(&hh);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(a - __uint_as_float(h[e] << 16), b - __uint_as_float(h[e] & 0xffff0000u));
            l[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const size_t off = (size_t)r * ROW_B + (size_t)((j ^ ((r >> 1) & 3)) << 4);
        *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(img_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    norm[r] = r < K ? n2 : CUDART_INF_F;
    if (r < K) atomicMax(nmax_bits, __float_as_uint(n2));     // non-negative floats order like their bit patterns
}

struct NnSide {
    const float* d;              // [K][32] FP32 descriptors (exact re-evaluation)
    const uint8_t* hi;           // bf16 images [Kpad][64 B]
    const uint8_t* lo;
    const float* norm;           // [Kpad]
    int K, Kpad;
};

struct NnArgs {
    NnSide side[2];              // blockIdx.y = 0: rows = side 0, columns = side 1; blockIdx.y = 1: swapped
    unsigned long long* best[2]; // best[dir][row]
    uint8_t* flag[2];            // overflow flags per row
    const unsigned int* nmax_bits[2];   // largest squared norm of side 0 / side 1
    int nsplit;                  // column tiles of a row tile are split over this many CTAs
};

struct __align__(8) NnBars {
    unsigned long long a_full, b_full[2], b_empty[2], t_full, t_empty;
    uint32_t tmem_base;
};

__device__ __forceinline__ unsigned long long exact_key(const float (&a)[YF], const float* __restrict__ brow, int col) {
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < YF / 4; ++q) {
        const float4 b = reinterpret_cast<const float4*>(brow)[q];
        float dlt = __fsub_rn(a[4 * q], b.x);     acc = __fmaf_rn(dlt, dlt, acc);
        dlt = __fsub_rn(a[4 * q + 1], b.y);       acc = __fmaf_rn(dlt, dlt, acc);
        dlt = __fsub_rn(a[4 * q + 2], b.z);       acc = __fmaf_rn(dlt, dlt, acc);
        dlt = __fsub_rn(a[4 * q + 3], b.w);       acc = __fmaf_rn(dlt, dlt, acc);
    }
    const float dist = __fsqrt_rn(__fadd_rn(acc, 1e-7f));
    return ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)col;
}

// 192 threads: warps 0-3 = epilogue (warp w owns TMEM lanes 32w..32w+31 = rows), warp 4 = TMA producer, warp 5 = MMA issuer + TMEM.
constexpr int NN_SMEM = 1024 + 2 * NT_M * ROW_B + 4 * NT_N * ROW_B + 2 * NT_N * 4 + 128;     // 84 096 B: two CTAs per SM
__global__ void __launch_bounds__(192) nn_tc_kernel(const NnArgs p) {
    extern __shared__ __align__(1024) uint8_t nn_smem_raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)nn_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t (*a_s)[NT_M * ROW_B] = reinterpret_cast<uint8_t (*)[NT_M * ROW_B]>(sm);                          // [2]: hi, lo
    uint8_t (*b_s)[2][NT_N * ROW_B] = reinterpret_cast<uint8_t (*)[2][NT_N * ROW_B]>(sm + 2 * NT_M * ROW_B);   // [stage][hi, lo]
    float (*nb_s)[NT_N] = reinterpret_cast<float (*)[NT_N]>(sm + 2 * NT_M * ROW_B + 4 * NT_N * ROW_B);
    NnBars& bars = *reinterpret_cast<NnBars*>(sm + 2 * NT_M * ROW_B + 4 * NT_N * ROW_B + 2 * NT_N * 4);
    const int dir = blockIdx.y;
    const NnSide R = p.side[dir], C = p.side[dir ^ 1];
    const int m_tile = blockIdx.x / p.nsplit, part = blockIdx.x - m_tile * p.nsplit;
    if (m_tile * NT_M >= R.K) return;                                      // uniform per CTA
    const int n_tiles = C.Kpad / NT_N;
    const int t_lo = (int)((long long)n_tiles * part / p.nsplit), t_hi = (int)((long long)n_tiles * (part + 1) / p.nsplit);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT_N >> 3) << 17) | ((uint32_t)(NT_M >> 4) << 24);

    if (threadIdx.x == 0) {
        mbar_init(&bars.a_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&bars.b_full[s], 1); mbar_init(&bars.b_empty[s], 1); }
        mbar_init(&bars.t_full, 1);
        mbar_init(&bars.t_empty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&bars.tmem_base)), "r"(NT_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars.tmem_base;

    if (warp == 4) {
        if (lane == 0 && t_hi > t_lo) {
            mbar_expect_tx(&bars.a_full, 2 * NT_M * ROW_B);
            bulk_g2s(a_s[0], R.hi + (size_t)m_tile * NT_M * ROW_B, NT_M * ROW_B, &bars.a_full);
            bulk_g2s(a_s[1], R.lo + (size_t)m_tile * NT_M * ROW_B, NT_M * ROW_B, &bars.a_full);
            for (int t = t_lo; t < t_hi; ++t) {
                const int it = t - t_lo, s = it & 1;
                mbar_wait(&bars.b_empty[s], ((it >> 1) & 1) ^ 1);
                mbar_expect_tx(&bars.b_full[s], 2 * NT_N * ROW_B);
                bulk_g2s(b_s[s][0], C.hi + (size_t)t * NT_N * ROW_B, NT_N * ROW_B, &bars.b_full[s]);
                bulk_g2s(b_s[s][1], C.lo + (size_t)t * NT_N * ROW_B, NT_N * ROW_B, &bars.b_full[s]);
            }
        }
    } else if (warp == 5) {
        if (lane == 0 && t_hi > t_lo) {
            mbar_wait(&bars.a_full, 0);
            const uint64_t a_hi = desc_sw64(a_s[0]), a_lo = desc_sw64(a_s[1]);
            for (int t = t_lo; t < t_hi; ++t) {
                const int it = t - t_lo, s = it & 1;
                mbar_wait(&bars.t_empty, (it & 1) ^ 1);                    // the epilogue has drained the accumulator
                mbar_wait(&bars.b_full[s], (it >> 1) & 1);
                tc_fence_after();
                const uint64_t b_hi = desc_sw64(b_s[s][0]), b_lo = desc_sw64(b_s[s][1]);
#pragma unroll
                for (uint32_t ks = 0; ks < 2; ++ks) {                      // K = 32 = two 16-element steps, 32 bytes apart
                    const uint64_t adv = (uint64_t)(ks * 2);
                    tc_mma(tmem, a_hi + adv, b_hi + adv, IDESC, ks ? 1u : 0u);
                    tc_mma(tmem, a_lo + adv, b_hi + adv, IDESC, 1u);
                    tc_mma(tmem, a_hi + adv, b_lo + adv, IDESC, 1u);
                }
                tc_commit(&bars.b_empty[s]);
                tc_commit(&bars.t_full);
            }
        }
    } else {
        // ---- epilogue: one thread per row ----
        const int row = m_tile * NT_M + warp * 32 + lane;
        const bool ok = row < R.K;
        float a[YF];
#pragma unroll
        for (int q = 0; q < YF / 4; ++q) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) x = reinterpret_cast<const float4*>(R.d + (size_t)row * YF)[q];
            a[4 * q] = x.x; a[4 * q + 1] = x.y; a[4 * q + 2] = x.z; a[4 * q + 3] = x.w;
        }
        const float na = ok ? R.norm[row] : 0.f;
        float v0 = CUDART_INF_F, v1 = CUDART_INF_F, v2 = CUDART_INF_F, v3 = CUDART_INF_F;
        int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
        for (int t = t_lo; t < t_hi; ++t) {
            const int it = t - t_lo;
            float* nb = nb_s[it & 1];
            nb[threadIdx.x] = C.norm[t * NT_N + threadIdx.x];
            nb[threadIdx.x + 128] = C.norm[t * NT_N + threadIdx.x + 128];
            asm volatile("bar.sync 1, 128;\n" ::: "memory");               // the four epilogue warps only
            mbar_wait(&bars.t_full, it & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int cc = 0; cc < NT_N / 32; ++cc) {
                uint32_t r[32];
                tmem_ld32(taddr + cc * 32, r);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float d = fmaf(-2.f, __uint_as_float(r[i]), na + nb[cc * 32 + i]);
                    if (d < v3) {
                        const int c = t * NT_N + cc * 32 + i;
                        if (d < v2) {
                            v3 = v2; i3 = i2;
                            if (d < v1) {
                                v2 = v1; i2 = i1;
                                if (d < v0) { v1 = v0; i1 = i0; v0 = d; i0 = c; } else { v1 = d; i1 = c; }
                            } else { v2 = d; i2 = c; }
                        } else { v3 = d; i3 = c; }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars.t_empty);
        }
        if (ok && i0 >= 0) {
            const float nbmax = __uint_as_float(*p.nmax_bits[dir ^ 1]);
            const float thr = v0 + 2.f * TOL_REL * (na + nbmax) + 1e-30f;
            unsigned long long best = exact_key(a, C.d + (size_t)i0 * YF, i0);
            if (i1 >= 0 && v1 <= thr) { const unsigned long long k = exact_key(a, C.d + (size_t)i1 * YF, i1); best = k < best ? k : best; }
            if (i2 >= 0 && v2 <= thr) { const unsigned long long k = exact_key(a, C.d + (size_t)i2 * YF, i2); best = k < best ? k : best; }
            if (i3 >= 0 && v3 <= thr) {                                    // the window may hold more than four columns: exhaustive re-scan
                const unsigned long long k = exact_key(a, C.d + (size_t)i3 * YF, i3); best = k < best ? k : best;
                p.flag[dir][row] = 1;
            }
            atomicMin(&p.best[dir][row], best);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(NT_N) : "memory");
    }
}

// Exhaustive exact re-scan of the flagged rows, one warp per row.
__global__ void __launch_bounds__(256) nn_fix_kernel(const NnArgs p) {
    const int dir = blockIdx.y;
    const NnSide R = p.side[dir], C = p.side[dir ^ 1];
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= R.K || !p.flag[dir][row]) return;
    float a[YF];
#pragma unroll
    for (int q = 0; q < YF / 4; ++q) {
        const float4 x = reinterpret_cast<const float4*>(R.d + (size_t)row * YF)[q];
        a[4 * q] = x.x; a[4 * q + 1] = x.y; a[4 * q + 2] = x.z; a[4 * q + 3] = x.w;
    }
    unsigned long long best = ~0ull;
    for (int c = lane; c < C.K; c += 32) {
        const unsigned long long k = exact_key(a, C.d + (size_t)c * YF, c);
        best = k < best ? k : best;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, best, o);
        best = u < best ? u : best;
    }
    if (lane == 0) atomicMin(&p.best[dir][row], best);
}

}  // namespace

size_t nn_tc_ws_bytes(int Ka, int Kb) {
    const size_t pa = (size_t)((Ka + NT_N - 1) / NT_N) * NT_N, pb = (size_t)((Kb + NT_N - 1) / NT_N) * NT_N;
    return (pa + pb) * (2 * ROW_B + 4 + 1) + 1024;
}

// rowbest / colbest must already hold ~0 (fill_u64_kernel).  `ws` = nn_tc_ws_bytes(Ka, Kb) bytes, 1024-byte aligned.
int nn_pass_tc(yoho_ctx* ctx, const float* dA, int Ka, const float* dB, int Kb, unsigned long long* rowbest,
               unsigned long long* colbest, void* ws, cudaStream_t st) {
    const int pa = ((Ka + NT_N - 1) / NT_N) * NT_N, pb = ((Kb + NT_N - 1) / NT_N) * NT_N;
    uint8_t* w = (uint8_t*)ws;
    uint8_t* a_hi = w;                       w += (size_t)pa * ROW_B;
    uint8_t* a_lo = w;                       w += (size_t)pa * ROW_B;
    uint8_t* b_hi = w;                       w += (size_t)pb * ROW_B;
    uint8_t* b_lo = w;                       w += (size_t)pb * ROW_B;
    float* na = (float*)w;                   w += (size_t)pa * 4;
    float* nb = (float*)w;                   w += (size_t)pb * 4;
    unsigned int* nmax = (unsigned int*)w;   w += 256;
    uint8_t* fa = w;                         w += pa;
    uint8_t* fb = w;
    YCHECK(cudaMemsetAsync(nmax, 0, 256 + (size_t)pa + pb, st));          // nmax[0..1] and both flag arrays
    nn_prep_kernel<<<(pa + 127) / 128, 128, 0, st>>>(dA, Ka, pa, a_hi, a_lo, na, nmax);
    nn_prep_kernel<<<(pb + 127) / 128, 128, 0, st>>>(dB, Kb, pb, b_hi, b_lo, nb, nmax + 1);
    NnArgs p;
    p.side[0] = NnSide{dA, a_hi, a_lo, na, Ka, pa};
    p.side[1] = NnSide{dB, b_hi, b_lo, nb, Kb, pb};
    p.best[0] = rowbest; p.best[1] = colbest;
    p.flag[0] = fa; p.flag[1] = fb;
    p.nmax_bits[0] = nmax; p.nmax_bits[1] = nmax + 1;
    const int mt = (((Ka > Kb ? Ka : Kb) + NT_M - 1) / NT_M);
    int nsplit = (2 * ctx->num_sms + 2 * mt - 1) / (2 * mt);               // ~one wave of two resident CTAs per SM over both directions
    const int min_tiles = (pa < pb ? pa : pb) / NT_N;
    if (nsplit > min_tiles) nsplit = min_tiles;
    if (nsplit < 1) nsplit = 1;
    p.nsplit = nsplit;
    dim3 grid(mt * nsplit, 2);
    static bool attr_set = false;
    if (!attr_set) {
        YCHECK(cudaFuncSetAttribute(nn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NN_SMEM));
        attr_set = true;
    }
    nn_tc_kernel<<<grid, 192, NN_SMEM, st>>>(p);
    dim3 gfix(((Ka > Kb ? Ka : Kb) + 7) / 8, 2);
    nn_fix_kernel<<<gfix, 256, 0, st>>>(p);
    ctx->launches += 4;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
