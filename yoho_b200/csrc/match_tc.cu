// Mutual 1-NN on tensor cores with exact FP32 verification (tests/matcher.py:37-48, utils/knn_search.py:17-66).
//
// The reference's answer is an ARGMIN over FP32 values dist = sqrt(sum_f (a_f - b_f)^2 + 1e-7) with first-occurrence ties; the
// SIMT kernel (match.cu, nn_tile_kernel) evaluates all Ka x Kb of them in that arithmetic: 0.8 G subtract+FMA pairs, issue bound
// at ~98 us for 5000 x 5000.  Here the K x K work moves to tcgen05:
//
//   1. nn_sum / nn_prep descriptors, centred on the common mean of both sets (distances are unchanged, norms shrink to the spread),
//                       -> bf16 hi / lo images in the UMMA K-major SWIZZLE_64B layout (a row = 32 channels = 64 B), squared norms,
//                       the largest norm.
//   2. nn_tc_kernel     per 128-row tile: dot = A_hi B_hi + A_lo B_hi + A_hi B_lo (M128 x N256 x K32, six tcgen05.mma per column
//                       tile, FP32 in TMEM); the epilogue thread that owns a row (one TMEM lane) turns every dot product into the
//                       approximate squared distance |a|^2 + |b|^2 - 2 dot and keeps the FOUR smallest minima over chunks of 32
//                       consecutive columns (two instructions per element).  The approximation error is bounded (3-product bf16
//                       split: 2^-15 |a||b|, plus FP32 rounding of the norms) by tol = 1e-4 (|a|^2 + max|b|^2), so the true
//                       argmin — and every column that ties with it after the reference's rounding — lies in a chunk whose
//                       minimum is within 2 tol of the smallest one; the kernel writes those (normally one) chunk ids per row.
//   2b. nn_verify_kernel every column of the candidate chunks is evaluated in the reference's exact arithmetic (one subtraction and
//                       one FMA per channel in ascending channel order, IEEE sqrt), one warp per candidate list, and merged as
//                       64-bit (distance bits, index) keys with atomicMin — exactly the keys the SIMT kernel produces.  Rows and columns swap roles in a second set of CTAs
//                       (blockIdx.y), which gives the column minima without any cross-lane reduction.
//                       A row with a FOURTH list entry inside the window (many near-duplicates) may have in-window chunks that were
//                       not kept: its warp scans all columns exactly instead (and sets the row's flag); normally no row does.
//
// The result is bit-identical to nn_tile_kernel's (tests/test_gpu_parity.py, test_gpu_fullsize.py run both).
#include <cuda_bf16.h>
#include <math_constants.h>
#include "common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int NT_M = 128;            // rows per CTA (UMMA M)
constexpr int NT_N = 128;            // columns per MMA tile (UMMA N); image rows are padded to a multiple of 256
constexpr int ROW_B = 64;            // bytes per image row (32 bf16)
constexpr int TOPK = 4;
constexpr float TOL_REL = 1e-4f;

// K-major operand tile, 64-byte rows, SWIZZLE_64B, 8-row groups 512 B apart (same image as gconv_tc.cu's operands).
__device__ __forceinline__ uint64_t desc_sw64(const void* smem_tile) {
    const uint64_t addr = (uint64_t)(smem_u32(smem_tile) >> 4) & 0x3FFFull;
    return addr | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}

// Channel sums of a descriptor set (for the common centre), one CTA of 256 threads per 64 rows, atomicAdd of 32 partial sums.
constexpr int SUM_ROWS = 64;
__global__ void __launch_bounds__(256) nn_sum_kernel(const float* __restrict__ d0, int K0, const float* __restrict__ d1, int K1,
                                                    float* __restrict__ sum) {
    __shared__ float red[8][YF];
    const float* d = blockIdx.y ? d1 : d0;
    const int K = blockIdx.y ? K1 : K0;
    const int f = threadIdx.x & 31, w = threadIdx.x >> 5;
    float s = 0.f;
    const int r0 = blockIdx.x * SUM_ROWS, r1 = min(K, r0 + SUM_ROWS);
#pragma unroll
    for (int r = r0 + w; r < r1; r += 8) s += d[(size_t)r * YF + f];
    red[w][f] = s;
    __syncthreads();
    if (w == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i][f];
        atomicAdd(&sum[f], t);
    }
}

// One thread per row of the padded image (rows >= K are zero, their norm +inf so that they never become a candidate).
// Rows are CENTRED on the common mean of both sets before the split: |a - b| does not change, but the norms — and with them the
// cancellation error of |a|^2 + |b|^2 - 2 a.b — shrink to the spread of the descriptors (PartI descriptors cluster tightly:
// un-centred, almost every row of a real pair kept more than four columns inside its verification window).
struct PrepSide {
    const float* d; int K, Kpad;
    uint8_t* img_hi; uint8_t* img_lo; float* norm; unsigned int* nmax_bits;
    unsigned long long* best; uint8_t* flag;
};
__global__ void nn_prep_kernel(const PrepSide s0, const PrepSide s1, const float* __restrict__ sum, float inv_count) {
    const PrepSide& S = blockIdx.y ? s1 : s0;
    const float* __restrict__ d = S.d;
    const int K = S.K, Kpad = S.Kpad;
    uint8_t* __restrict__ img_hi = S.img_hi;
    uint8_t* __restrict__ img_lo = S.img_lo;
    float* __restrict__ norm = S.norm;
    unsigned int* __restrict__ nmax_bits = S.nmax_bits;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= Kpad) return;
    if (r < K) { S.best[r] = ~0ull; S.flag[r] = 0; }
    float v[YF];
    float n2 = 0.f;
#pragma unroll
    for (int q = 0; q < YF / 4; ++q) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < K) {
            x = reinterpret_cast<const float4*>(d + (size_t)r * YF)[q];
            const float4 m = reinterpret_cast<const float4*>(sum)[q];
            x.x -= m.x * inv_count; x.y -= m.y * inv_count; x.z -= m.z * inv_count; x.w -= m.w * inv_count;
        }
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
    int4* cand[2];               // [row][nsplit][2 halves]: the four smallest chunks (32 columns each) of a column share, -1 = none
    float4* cval[2];             //                            and their minima  |b|^2 - 2 a.b
    const unsigned int* nmax_bits[2];   // largest squared norm of side 0 / side 1
    int nsplit;                  // column tiles of a row tile are split over this many CTAs
};

struct __align__(8) NnBars {
    unsigned long long a_full, b_full[2], b_empty[2], t_full[2], t_empty[2];
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

// 320 threads: warps 0-7 = epilogue (warp w owns TMEM lanes 32 (w % 4) .. +31 = rows, and half w / 4 of every tile's columns),
// warp 8 = TMA producer, warp 9 = MMA issuer + TMEM.  Column tiles are 128 wide with TWO TMEM accumulators, so the MMAs of tile
// t+1 run while the epilogue scans tile t; two CTAs per SM.
constexpr int NN_THREADS = 320;
constexpr int NN_SMEM = 1024 + 2 * NT_M * ROW_B + 4 * NT_N * ROW_B + 2 * NT_N * 4 + 128;
__global__ void __launch_bounds__(NN_THREADS) nn_tc_kernel(const NnArgs p) {
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
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars.b_full[s], 1); mbar_init(&bars.b_empty[s], 1);
            mbar_init(&bars.t_full[s], 1); mbar_init(&bars.t_empty[s], 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&bars.tmem_base)), "r"(2 * NT_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars.tmem_base;

    if (warp == 8) {
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
    } else if (warp == 9) {
        if (lane == 0 && t_hi > t_lo) {
            mbar_wait(&bars.a_full, 0);
            const uint64_t a_hi = desc_sw64(a_s[0]), a_lo = desc_sw64(a_s[1]);
            for (int t = t_lo; t < t_hi; ++t) {
                const int it = t - t_lo, s = it & 1;
                mbar_wait(&bars.t_empty[s], ((it >> 1) & 1) ^ 1);          // the epilogue has drained accumulator s
                mbar_wait(&bars.b_full[s], (it >> 1) & 1);
                tc_fence_after();
                const uint64_t b_hi = desc_sw64(b_s[s][0]), b_lo = desc_sw64(b_s[s][1]);
                const uint32_t d_tmem = tmem + s * NT_N;
#pragma unroll
                for (uint32_t ks = 0; ks < 2; ++ks) {                      // K = 32 = two 16-element steps, 32 bytes apart
                    const uint64_t adv = (uint64_t)(ks * 2);
                    tc_mma(d_tmem, a_hi + adv, b_hi + adv, IDESC, ks ? 1u : 0u);
                    tc_mma(d_tmem, a_lo + adv, b_hi + adv, IDESC, 1u);
                    tc_mma(d_tmem, a_hi + adv, b_lo + adv, IDESC, 1u);
                }
                tc_commit(&bars.b_empty[s]);
                tc_commit(&bars.t_full[s]);
            }
        }
    } else {
        // ---- epilogue: two threads per row (one per column half of every tile) ----
        const int q = warp & 3, half = warp >> 2;
        const int row = m_tile * NT_M + q * 32 + lane;
        const bool ok = row < R.K;
        // The four smallest CHUNK minima (a chunk = 32 consecutive columns) of key = |b|^2 - 2 a.b (|a|^2 is constant along a row):
        // two instructions per element, no per-element index bookkeeping; the exact arithmetic then runs over every column of
        // the chunks whose minimum lies inside the verification window (normally one chunk, 32 columns).
        float v0 = CUDART_INF_F, v1 = CUDART_INF_F, v2 = CUDART_INF_F, v3 = CUDART_INF_F;
        int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
        constexpr int CPT = NT_N / 32;                                     // chunks per tile
        for (int t = t_lo; t < t_hi; ++t) {
            const int it = t - t_lo, s = it & 1;
            float* nb = nb_s[s];
            if (threadIdx.x < NT_N) nb[threadIdx.x] = C.norm[t * NT_N + threadIdx.x];
            asm volatile("bar.sync 1, 256;\n" ::: "memory");               // the eight epilogue warps only
            mbar_wait(&bars.t_full[s], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + s * NT_N + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int cc = half * (CPT / 2); cc < (half + 1) * (CPT / 2); ++cc) {
                uint32_t r[32];
                tmem_ld32(taddr + cc * 32, r);
                float m0 = CUDART_INF_F, m1 = CUDART_INF_F, m2 = CUDART_INF_F, m3 = CUDART_INF_F;      // four independent chains
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 n4 = *reinterpret_cast<const float4*>(nb + cc * 32 + 4 * j);      // broadcast read
                    m0 = fminf(m0, fmaf(-2.f, __uint_as_float(r[4 * j]), n4.x));
                    m1 = fminf(m1, fmaf(-2.f, __uint_as_float(r[4 * j + 1]), n4.y));
                    m2 = fminf(m2, fmaf(-2.f, __uint_as_float(r[4 * j + 2]), n4.z));
                    m3 = fminf(m3, fmaf(-2.f, __uint_as_float(r[4 * j + 3]), n4.w));
                }
                const float cmin = fminf(fminf(m0, m1), fminf(m2, m3));
                if (cmin < v3) {
                    const int c = t * CPT + cc;
                    if (cmin < v2) {
                        v3 = v2; i3 = i2;
                        if (cmin < v1) {
                            v2 = v1; i2 = i1;
                            if (cmin < v0) { v1 = v0; i1 = i0; v0 = cmin; i0 = c; } else { v1 = cmin; i1 = c; }
                        } else { v2 = cmin; i2 = c; }
                    } else { v3 = cmin; i3 = c; }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars.t_empty[s]);
        }
        // The four smallest chunk minima of this thread's share of the columns go to nn_verify_kernel, which merges the lists of a
        // row and runs the exact arithmetic (inside this kernel the 32-column evaluations were latency bound and cost more than
        // the search itself).
        if (ok) {
            const size_t o = ((size_t)row * p.nsplit + part) * 2 + half;
            p.cand[dir][o] = make_int4(i0, i1, i2, i3);
            p.cval[dir][o] = make_float4(v0, v1, v2, v3);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(2 * NT_N) : "memory");
    }
}

// Exact evaluation of the candidate chunks, one warp per row.  The lists of the row (one per split part and column half, four
// entries each, <= 32 entries: one per lane) are merged: a column can beat the best one only if its approximate key is within
// 2 tol of the smallest key of the ROW, tol = TOL_REL (|a|^2 + max |b|^2).  Chunks a list did not keep have a minimum >= that
// list's fourth entry, so the merged set is complete unless some fourth entry lies inside the window (-> flag, exhaustive
// re-scan).  For every chunk in the window lane l evaluates column 32 c + l (the 32 descriptor rows of a chunk are one contiguous
// 4 KB block) in the reference's arithmetic.
__global__ void __launch_bounds__(256) nn_verify_kernel(const NnArgs p) {
    const int dir = blockIdx.y;
    const NnSide R = p.side[dir], C = p.side[dir ^ 1];
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= R.K) return;
    const int n_ent = p.nsplit * 2 * 4;                                    // <= 32 (nsplit <= 4)
    float v = CUDART_INF_F;
    int id = -1;
    if (lane < n_ent) {
        v = reinterpret_cast<const float*>(p.cval[dir] + (size_t)row * p.nsplit * 2)[lane];
        id = reinterpret_cast<const int*>(p.cand[dir] + (size_t)row * p.nsplit * 2)[lane];
    }
    float vmin = id >= 0 ? v : CUDART_INF_F;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    const float thr = vmin + 2.f * TOL_REL * (R.norm[row] + __uint_as_float(*p.nmax_bits[dir ^ 1])) + 1e-30f;
    const bool in_win = id >= 0 && v <= thr;
    unsigned todo = __ballot_sync(0xffffffffu, in_win);
    const bool overflow = __ballot_sync(0xffffffffu, in_win && (lane & 3) == 3) != 0;
    float a[YF];
#pragma unroll
    for (int j = 0; j < YF / 4; ++j) {
        const float4 x = reinterpret_cast<const float4*>(R.d + (size_t)row * YF)[j];
        a[4 * j] = x.x; a[4 * j + 1] = x.y; a[4 * j + 2] = x.z; a[4 * j + 3] = x.w;
    }
    unsigned long long best = ~0ull;
    if (overflow) {                                                        // warp-uniform: exhaustive exact scan of this row (rare)
        if (lane == 0) p.flag[dir][row] = 1;
        for (int col = lane; col < C.K; col += 32) {
            const unsigned long long key = exact_key(a, C.d + (size_t)col * YF, col);
            best = key < best ? key : best;
        }
        todo = 0;
    }
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int col = __shfl_sync(0xffffffffu, id, src) * 32 + lane;
        if (col < C.K) {
            const unsigned long long key = exact_key(a, C.d + (size_t)col * YF, col);
            best = key < best ? key : best;
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xffffffffu, best, o);
        best = u < best ? u : best;
    }
    if (lane == 0) atomicMin(&p.best[dir][row], best);
}

}  // namespace

constexpr int NN_PAD = 256;
constexpr int NN_MAX_SPLIT = 4;       // 4 parts x 2 halves x 4 entries = 32 candidates per row = one per lane of the verifying warp
size_t nn_tc_ws_bytes(int Ka, int Kb) {
    const size_t pa = (size_t)((Ka + NN_PAD - 1) / NN_PAD) * NN_PAD, pb = (size_t)((Kb + NN_PAD - 1) / NN_PAD) * NN_PAD;
    return (pa + pb) * (2 * ROW_B + 4 + 1 + (size_t)NN_MAX_SPLIT * 2 * (sizeof(int4) + sizeof(float4))) + 4096;
}

// rowbest / colbest are initialised here (nn_prep_kernel).  `ws` = nn_tc_ws_bytes(Ka, Kb) bytes, 1024-byte aligned.
int nn_pass_tc(yoho_ctx* ctx, const float* dA, int Ka, const float* dB, int Kb, unsigned long long* rowbest,
               unsigned long long* colbest, void* ws, cudaStream_t st) {
    const int pa = ((Ka + NN_PAD - 1) / NN_PAD) * NN_PAD, pb = ((Kb + NN_PAD - 1) / NN_PAD) * NN_PAD;
    uint8_t* w = (uint8_t*)ws;
    uint8_t* a_hi = w;                       w += (size_t)pa * ROW_B;
    uint8_t* a_lo = w;                       w += (size_t)pa * ROW_B;
    uint8_t* b_hi = w;                       w += (size_t)pb * ROW_B;
    uint8_t* b_lo = w;                       w += (size_t)pb * ROW_B;
    float* na = (float*)w;                   w += (size_t)pa * 4;
    float* nb = (float*)w;                   w += (size_t)pb * 4;
    unsigned int* nmax = (unsigned int*)w;   w += 256;                     // [0..1] largest centred norms, [32..63] channel sums
    uint8_t* fa = w;                         w += pa;
    uint8_t* fb = w;                         w += pb;
    w = (uint8_t*)(((uintptr_t)w + 255) & ~(uintptr_t)255);
    int4* ca = (int4*)w;                     w += (size_t)pa * NN_MAX_SPLIT * 2 * sizeof(int4);
    int4* cb = (int4*)w;                     w += (size_t)pb * NN_MAX_SPLIT * 2 * sizeof(int4);
    float4* va = (float4*)w;                 w += (size_t)pa * NN_MAX_SPLIT * 2 * sizeof(float4);
    float4* vb = (float4*)w;
    float* csum = (float*)(nmax + 32);
    YCHECK(cudaMemsetAsync(nmax, 0, 256, st));                            // nmax and the channel sums
    const int kmax = Ka > Kb ? Ka : Kb, pmax = pa > pb ? pa : pb;
    nn_sum_kernel<<<dim3((kmax + SUM_ROWS - 1) / SUM_ROWS, 2), 256, 0, st>>>(dA, Ka, dB, Kb, csum);
    const float inv_count = 1.0f / (float)((double)Ka + (double)Kb);
    const PrepSide s0{dA, Ka, pa, a_hi, a_lo, na, nmax, rowbest, fa}, s1{dB, Kb, pb, b_hi, b_lo, nb, nmax + 1, colbest, fb};
    nn_prep_kernel<<<dim3((pmax + 127) / 128, 2), 128, 0, st>>>(s0, s1, csum, inv_count);
    NnArgs p;
    p.side[0] = NnSide{dA, a_hi, a_lo, na, Ka, pa};
    p.side[1] = NnSide{dB, b_hi, b_lo, nb, Kb, pb};
    p.best[0] = rowbest; p.best[1] = colbest;
    p.flag[0] = fa; p.flag[1] = fb;
    p.cand[0] = ca; p.cand[1] = cb;
    p.cval[0] = va; p.cval[1] = vb;
    p.nmax_bits[0] = nmax; p.nmax_bits[1] = nmax + 1;
    const int mt = (((Ka > Kb ? Ka : Kb) + NT_M - 1) / NT_M);
    int nsplit = (2 * ctx->num_sms) / (2 * mt);                            // at most ONE wave of two resident CTAs per SM over both directions
    const int min_tiles = (pa < pb ? pa : pb) / NT_N;
    if (nsplit > min_tiles) nsplit = min_tiles;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > NN_MAX_SPLIT) nsplit = NN_MAX_SPLIT;
    p.nsplit = nsplit;
    dim3 grid(mt * nsplit, 2);
    // per-device attribute; cheap enough to set on every launch (one process may drive several devices)
    YCHECK(cudaFuncSetAttribute(nn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NN_SMEM));
    nn_tc_kernel<<<grid, NN_THREADS, NN_SMEM, st>>>(p);
    dim3 gver(((Ka > Kb ? Ka : Kb) + 7) / 8, 2);
    nn_verify_kernel<<<gver, 256, 0, st>>>(p);
    ctx->launches += 4;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
