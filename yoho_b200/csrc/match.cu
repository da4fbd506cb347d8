// Descriptor matching (tests/matcher.py:19-49, utils/knn_search.py:17-66,138-154) and the 60-way
// rotation-correlation argmax (tests/extractor.py:74-78).
#include "common.cuh"

size_t nn_tc_ws_bytes(int Ka, int Kb);                                                        // match_tc.cu
int nn_pass_tc(yoho_ctx* ctx, const float* dA, int Ka, const float* dB, int Kb, unsigned long long* rowbest,
               unsigned long long* colbest, void* ws, cudaStream_t st);

namespace {

// ---------------------------------------------------------------------------------------------------
// Mutual 1-NN.  One pass over the Ka x Kb distance matrix in 64x64 tiles; each CTA reduces its tile to
// 64 row minima and 64 column minima and merges them into global 64-bit keys with atomicMin:
//     key = (float_bits(sqrt(d2 + 1e-7)) << 32) | index
// Distances are positive, so the unsigned order of the key is the lexicographic (distance, index) order:
// equal float distances resolve to the LOWEST index, which is torch.min's first-occurrence rule
// (utils/knn_search.py:41).  d2 is the reference's direct form sum_f (a_f - b_f)^2 accumulated in FP32 with
// one FMA per channel in ascending channel order (the reference's summation order is library-defined).
// ---------------------------------------------------------------------------------------------------
constexpr int MT = 64;

__global__ void __launch_bounds__(256) nn_tile_kernel(const float* __restrict__ dA, int Ka, const float* __restrict__ dB,
                                                     int Kb, unsigned long long* __restrict__ rowbest,
                                                     unsigned long long* __restrict__ colbest) {
    __shared__ float As[YF][MT + 4];
    __shared__ float Bs[YF][MT + 4];
    __shared__ unsigned long long colred[16][MT];
    const int t = threadIdx.x;
    const int a0 = blockIdx.y * MT, b0 = blockIdx.x * MT;
    for (int i = t; i < MT * YF; i += 256) {
        const int r = i / YF, f = i % YF;
        As[f][r] = (a0 + r < Ka) ? dA[(size_t)(a0 + r) * YF + f] : 0.f;
        Bs[f][r] = (b0 + r < Kb) ? dB[(size_t)(b0 + r) * YF + f] : 0.f;
    }
    __syncthreads();
    const int ta = t >> 4, tb = t & 15;   // 16 x 16 threads, 4 x 4 distances each
    // Packed FP32 (sub.f32x2 / fma.rn.f32x2, sm_100): the two lanes of a packed register are two NEIGHBOURING COLUMNS of the
    // tile, so every distance is still one subtraction and one FMA per channel in ascending channel order (bit-identical to
    // the scalar loop) at half the instruction count — this kernel is issue bound.
    unsigned long long acc2[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc2[i][0] = 0ull; acc2[i][1] = 0ull; }
#pragma unroll 8
    for (int f = 0; f < YF; ++f) {
        const float4 av = *reinterpret_cast<const float4*>(&As[f][ta * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[f][tb * 4]);
        const float a[4] = {av.x, av.y, av.z, av.w};
        unsigned long long b01, b23;
        asm("mov.b64 %0, {%1, %2};" : "=l"(b01) : "f"(bv.x), "f"(bv.y));
        asm("mov.b64 %0, {%1, %2};" : "=l"(b23) : "f"(bv.z), "f"(bv.w));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned long long aa, d01, d23;
            asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a[i]));
            asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(aa), "l"(b01));
            asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(aa), "l"(b23));
            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc2[i][0]) : "l"(d01));
            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc2[i][1]) : "l"(d23));
        }
    }
    float d2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        asm("mov.b64 {%0, %1}, %2;" : "=f"(d2[i][0]), "=f"(d2[i][1]) : "l"(acc2[i][0]));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(d2[i][2]), "=f"(d2[i][3]) : "l"(acc2[i][1]));
    }
    unsigned long long rkey[4], ckey[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { rkey[i] = ~0ull; ckey[i] = ~0ull; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ia = a0 + ta * 4 + i, ib = b0 + tb * 4 + j;
            if (ia < Ka && ib < Kb) {
                const float d = __fsqrt_rn(__fadd_rn(d2[i][j], 1e-7f));
                const unsigned long long hi = (unsigned long long)__float_as_uint(d) << 32;
                const unsigned long long kr = hi | (unsigned)ib;
                const unsigned long long kc = hi | (unsigned)ia;
                rkey[i] = kr < rkey[i] ? kr : rkey[i];
                ckey[j] = kc < ckey[j] ? kc : ckey[j];
            }
        }
    // row minima: reduce over the 16 tb lanes (consecutive lanes of a half-warp)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unsigned long long v = rkey[i];
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
            v = u < v ? u : v;
        }
        if (tb == 0 && a0 + ta * 4 + i < Ka) atomicMin(&rowbest[a0 + ta * 4 + i], v);
    }
    // column minima: reduce over the 16 ta groups through shared memory
#pragma unroll
    for (int j = 0; j < 4; ++j) colred[ta][tb * 4 + j] = ckey[j];
    __syncthreads();
    if (t < MT) {
        unsigned long long v = colred[0][t];
#pragma unroll
        for (int r = 1; r < 16; ++r) v = colred[r][t] < v ? colred[r][t] : v;
        if (b0 + t < Kb) atomicMin(&colbest[b0 + t], v);
    }
}

__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// Mutual filter + ordered compaction (tests/matcher.py:41-48).  Single CTA, 1024 threads.
__global__ void __launch_bounds__(1024) mutual_compact_kernel(const unsigned long long* __restrict__ rowbest, int Ka,
                                                             const unsigned long long* __restrict__ colbest, int Kb,
                                                             int64_t* __restrict__ pairs, int32_t* __restrict__ n_pairs,
                                                             int32_t* __restrict__ nnA, int32_t* __restrict__ nnB) {
    __shared__ int warp_tot[32];
    __shared__ int base_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) base_s = 0;
    __syncthreads();
    if (nnB)
        for (int i = t; i < Kb; i += 1024) nnB[i] = (int32_t)(colbest[i] & 0xffffffffu);
    for (int start = 0; start < Ka; start += 1024) {
        const int a = start + t;
        int flag = 0, j = 0;
        if (a < Ka) {
            j = (int)(rowbest[a] & 0xffffffffu);
            if (nnA) nnA[a] = j;
            flag = ((int)(colbest[j] & 0xffffffffu) == a);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        const int pre = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_tot[w] = __popc(bal);
        __syncthreads();
        int woff = 0;
        for (int i = 0; i < w; ++i) woff += warp_tot[i];
        const int base = base_s;
        if (flag) {
            const int o = base + woff + pre;
            pairs[2 * (size_t)o] = a;
            pairs[2 * (size_t)o + 1] = j;
        }
        __syncthreads();
        if (t == 0) {
            int tot = 0;
            for (int i = 0; i < 32; ++i) tot += warp_tot[i];
            base_s = base + tot;
        }
        __syncthreads();
    }
    if (t == 0) *n_pairs = base_s;
}

__global__ void nn1_out_kernel(const unsigned long long* __restrict__ best, int m, float* dist, int64_t* idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const unsigned long long k = best[i];
        if (dist) dist[i] = __uint_as_float((unsigned)(k >> 32));
        if (idx) idx[i] = (int64_t)(k & 0xffffffffu);
    }
}

__global__ void pad32_kernel(const float* __restrict__ src, int n, int F, float* __restrict__ dst) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)n * YF) {
        const int r = (int)(i / YF), f = (int)(i % YF);
        dst[i] = f < F ? src[(size_t)r * F + f] : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------
// Rotation-correlation argmax.  Persistent 64-thread CTAs (four per SM) walk the matches; both [32x60] tiles of a match (7680 B
// each, contiguous in HBM) are staged in shared memory by two 1-D TMA bulk copies that signal one mbarrier, TWO matches ahead of
// the one being computed (two stages).  60 threads form the 60x60 channel-contracted product of the two tiles in 6x10 register
// blocks, then each sums, for its rotation, the 60 permuted entries (four quarters of the group axis, added in a fixed order);
// the argmax takes the lowest index among equal values (torch.argmax).
// ---------------------------------------------------------------------------------------------------
constexpr int TILE_FLOATS = YF * YG;           // 1920
constexpr int TILE_BYTES = TILE_FLOATS * 4;    // 7680
constexpr int ROT_ST = 2;
constexpr int ROT_T = 64;

struct __align__(128) RotSmem {
    float s1[ROT_ST][TILE_FLOATS];
    float s2[ROT_ST][TILE_FLOATS];
    float S[YG][YG + 1];                       // S[g][g'] = sum_f des2[f][g] * des1[f][g']
    float part[4][64];
    uint8_t pt[YG * YG];                       // pt[g*60 + a] = P[a][g]
    unsigned long long bar[ROT_ST];
};

__global__ void __launch_bounds__(ROT_T) rot_argmax_kernel(const float* __restrict__ des1, const int64_t* __restrict__ rows1,
                                                          const float* __restrict__ des2, const int64_t* __restrict__ rows2,
                                                          int row_stride, int M, const uint8_t* __restrict__ perm_t,
                                                          int64_t* __restrict__ idx_out, float* __restrict__ cor_out) {
    extern __shared__ __align__(128) uint8_t rot_smem_raw[];
    RotSmem& sm = *reinterpret_cast<RotSmem*>(rot_smem_raw);
    const int t = threadIdx.x;
    auto fetch = [&](int m, int st) {          // thread 0: both tiles of match m -> stage st
        const int64_t r1 = rows1 ? rows1[(size_t)m * row_stride] : m;
        const int64_t r2 = rows2 ? rows2[(size_t)m * row_stride] : m;
        const uint32_t bar_a = smem_u32(&sm.bar[st]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_a), "r"(2 * TILE_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                         smem_u32(sm.s1[st])),
                     "l"(des1 + (size_t)r1 * TILE_FLOATS), "r"(TILE_BYTES), "r"(bar_a)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                         smem_u32(sm.s2[st])),
                     "l"(des2 + (size_t)r2 * TILE_FLOATS), "r"(TILE_BYTES), "r"(bar_a)
                     : "memory");
    };
    if (t == 0) {
        for (int st = 0; st < ROT_ST; ++st) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&sm.bar[st])));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
        for (int st = 0; st < ROT_ST; ++st) {
            const int m = blockIdx.x + st * gridDim.x;
            if (m < M) fetch(m, st);
        }
    }
    for (int i = t; i < YG * YG / 4; i += ROT_T) reinterpret_cast<uint32_t*>(sm.pt)[i] = __ldg(reinterpret_cast<const uint32_t*>(perm_t) + i);
    __syncthreads();
    const int gi = (t / 6) * 6, gj = (t % 6) * 10;       // this thread's 6 x 10 block of S (t < 60)
    int it = 0;
    for (int m = blockIdx.x; m < M; m += gridDim.x, ++it) {
        const int st = it % ROT_ST;
        {   // wait for both tiles of this match
            const uint32_t bar_a = smem_u32(&sm.bar[st]), par = (uint32_t)(it / ROT_ST) & 1u;
            uint32_t ok = 0;
            while (!ok) {
                asm volatile(
                    "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                    : "=r"(ok)
                    : "r"(bar_a), "r"(par)
                    : "memory");
            }
        }
        const float* s1 = sm.s1[st];
        const float* s2 = sm.s2[st];
        // Step 1: the 60x60 channel-contracted product S[g][g'] = sum_f des2[f][g] des1[f][g'] in 6x10 register blocks (60 threads):
        // eight 8-byte shared reads per 60 FMAs (a 4x4 blocking — two 16-byte reads per 16 FMAs — was bound by shared-memory
        // wavefronts at 38 us per 2800 matches; the sum over f runs in the same order, so S is unchanged bit for bit).
        if (t < YG) {
            float acc[6][10];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 10; ++j) acc[i][j] = 0.f;
#pragma unroll 2
            for (int f = 0; f < YF; ++f) {
                float uu[6], vv[10];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float2 u = *reinterpret_cast<const float2*>(s2 + f * YG + gi + 2 * i);
                    uu[2 * i] = u.x; uu[2 * i + 1] = u.y;
                }
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const float2 v = *reinterpret_cast<const float2*>(s1 + f * YG + gj + 2 * j);
                    vv[2 * j] = v.x; vv[2 * j + 1] = v.y;
                }
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 10; ++j) acc[i][j] = fmaf(uu[i], vv[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 10; ++j) sm.S[gi + i][gj + j] = acc[i][j];
        }
        __syncthreads();                       // S complete; the tiles of this stage are consumed
        if (t == 0) {
            const int mn = m + ROT_ST * gridDim.x;
            if (mn < M) fetch(mn, st);
        }
        // Step 2: cor[a] = sum_g S[g][P[a][g]] — 60 gathered adds per rotation instead of 1920 gathered multiply-adds (the
        // permutation acts on the group axis only, so it commutes with the channel sum); four quarter sums added in a fixed order.
        float c = -INFINITY;
        if (t < YG) {
            float q4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float acc = 0.f;
#pragma unroll
                for (int g = q * 15; g < q * 15 + 15; ++g) acc += sm.S[g][sm.pt[g * YG + t]];
                q4[q] = acc;
            }
            c = ((q4[0] + q4[1]) + q4[2]) + q4[3];
            if (cor_out) cor_out[(size_t)m * YG + t] = c;
        }
        // argmax with lowest-index tie-break: inside each warp, then warp 1's (rotations 32..59) against warp 0's
        float best = c;
        int bi = t;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (t == 32) { sm.part[0][0] = best; sm.part[0][1] = __int_as_float(bi); }
        __syncthreads();                       // also: every read of S is done before the next match's step 1 writes it
        if (t == 0) {
            const float ob = sm.part[0][0];
            const int oi = __float_as_int(sm.part[0][1]);
            if (ob > best) bi = oi;            // equal values keep the lower index (warp 0's)
            idx_out[m] = bi;
        }
    }
}

// Tensor-core search with exact verification (match_tc.cu) when both sides have at least one 256-row tile; the FP32 SIMT kernel
// otherwise (and as the test twin: tuning flag 16384).  Both produce the same keys.
bool nn_use_tc(const yoho_ctx* ctx, int Ka, int Kb) {
    return ctx->gconv_impl >= 1 && !(ctx->tc_flags & 16384) && Ka >= 256 && Kb >= 256;
}

int nn_pass(yoho_ctx* ctx, const float* dA, int Ka, const float* dB, int Kb, unsigned long long* rowbest,
            unsigned long long* colbest, void* tc_ws, cudaStream_t st) {
    if (tc_ws && nn_use_tc(ctx, Ka, Kb)) return nn_pass_tc(ctx, dA, Ka, dB, Kb, rowbest, colbest, tc_ws, st);
    const size_t n = (size_t)Ka + Kb;   // rowbest and colbest are contiguous
    fill_u64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rowbest, n, ~0ull);
    ctx->launches++;
    dim3 grid((Kb + MT - 1) / MT, (Ka + MT - 1) / MT);
    nn_tile_kernel<<<grid, 256, 0, st>>>(dA, Ka, dB, Kb, rowbest, colbest);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" int yoho_mutual_nn(yoho_ctx* ctx, const float* dA, int Ka, const float* dB, int Kb, int64_t* pairs,
                              int32_t* n_pairs, int32_t* nnA, int32_t* nnB, void* stream) {
    YARG(ctx && dA && dB && pairs && n_pairs && Ka > 0 && Kb > 0);
    cudaStream_t st = (cudaStream_t)stream;
    YCHECK(cudaSetDevice(ctx->device));
    const size_t keys = align_up(((size_t)Ka + Kb) * 8, 1024);
    if (int rc = yoho_ws_reserve(ctx, keys + nn_tc_ws_bytes(Ka, Kb))) return rc;
    unsigned long long* rowbest = (unsigned long long*)ctx->ws;
    unsigned long long* colbest = rowbest + Ka;
    if (int rc = nn_pass(ctx, dA, Ka, dB, Kb, rowbest, colbest, (char*)ctx->ws + keys, st)) return rc;
    mutual_compact_kernel<<<1, 1024, 0, st>>>(rowbest, Ka, colbest, Kb, pairs, n_pairs, nnA, nnB);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

extern "C" int yoho_nn1(yoho_ctx* ctx, const float* source, int m, const float* target, int n, int F, float* dist,
                        int64_t* idx, void* stream) {
    YARG(ctx && source && target && m > 0 && n > 0 && F >= 1 && F <= YF);
    cudaStream_t st = (cudaStream_t)stream;
    YCHECK(cudaSetDevice(ctx->device));
    const size_t keys = align_up(((size_t)m + n) * 8, 1024);
    const size_t padb = F == YF ? 0 : align_up(((size_t)m + n) * YF * 4, 1024);
    if (int rc = yoho_ws_reserve(ctx, keys + padb + nn_tc_ws_bytes(m, n))) return rc;
    unsigned long long* rowbest = (unsigned long long*)ctx->ws;
    unsigned long long* colbest = rowbest + m;
    const float* s = source;
    const float* tg = target;
    if (F != YF) {   // zero channels add exactly 0 to every distance
        float* ps = (float*)((char*)ctx->ws + keys);
        float* pt = ps + (size_t)m * YF;
        pad32_kernel<<<(unsigned)(((size_t)m * YF + 255) / 256), 256, 0, st>>>(source, m, F, ps);
        pad32_kernel<<<(unsigned)(((size_t)n * YF + 255) / 256), 256, 0, st>>>(target, n, F, pt);
        ctx->launches += 2;
        s = ps; tg = pt;
    }
    if (int rc = nn_pass(ctx, s, m, tg, n, rowbest, colbest, (char*)ctx->ws + keys + padb, st)) return rc;
    nn1_out_kernel<<<(m + 255) / 256, 256, 0, st>>>(rowbest, m, dist, idx);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

extern "C" int yoho_rot_argmax(yoho_ctx* ctx, const float* des1, const int64_t* rows1, const float* des2,
                               const int64_t* rows2, int row_stride, int M, int64_t* idx, float* cor_out, void* stream) {
    YARG(ctx && des1 && des2 && idx && M >= 0 && row_stride >= 1);
    if (M == 0) return YOHO_OK;
    YCHECK(cudaSetDevice(ctx->device));
    // per-device attribute; cheap enough to set on every launch (one process may drive several devices)
    YCHECK(cudaFuncSetAttribute(rot_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RotSmem)));
    const int grid = M < 4 * ctx->num_sms ? M : 4 * ctx->num_sms;
    rot_argmax_kernel<<<grid, ROT_T, sizeof(RotSmem), (cudaStream_t)stream>>>(des1, rows1, des2, rows2, row_stride, M, ctx->d_perm_t, idx, cor_out);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
