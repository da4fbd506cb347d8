// Group-Fourier transforms on tcgen05 (same contract as fourier_mma.cu / fourier.cu):
//     mid[m][c] = sum_k M1[k][m] in[k][c] ;  pointwise (bias / shortcut / BN+ReLU) ;  out[m][c] = sum_k M2[k][m] mid[k][c]
// computed TRANSPOSED so that the 128 channels of a tile are the UMMA M axis (= TMEM lanes, one channel per epilogue thread):
//     D1[c][m] = sum_k X[k][c] M1t[m][k]        A = X, 128 channels x 64 k, MN-major (channels contiguous: the [k][c] rows of the
//                                               activation tensor land in shared memory unchanged), B = M1t [64 m][64 k] K-major
//     D2[c][m'] = sum_m mid[c][m] M2t[m'][m]     A = the pointwise result, written by the epilogue as a K-major image, B = M2t
// Each product is 3 bf16 MMAs (hi*hi, lo*hi, hi*lo), M128 x N64 x K16, four K steps.  The transform is HBM-bound (60x60 per
// channel): a persistent CTA per SM streams (keypoint, 128-channel) tiles through a 3-stage cp.async ring; warps 0-3 load,
// warp 12 issues the MMAs, two sets of four epilogue warps (4-7, 8-11) alternate tiles so that the pointwise stage, the
// hi/lo packing and the stores of one tile overlap the loads and MMAs of the next.
#include <cuda_bf16.h>
#include "common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int XT_THREADS = 416;
constexpr int XCH = 128;                     // channels per tile (UMMA M)
constexpr int D_TILE = 64 * XCH * 2;         // bytes of one [64 k][128 c] bf16 operand image (hi or lo): 2 x 8 atoms of 1 KB
constexpr int M_IMG = 64 * 64 * 2;           // bytes of one [64][64] bf16 matrix image
constexpr int A2_TILE = XCH * 64 * 2;        // bytes of one [128 c][64 k] K-major image (hi or lo)
constexpr int NST = 3;

struct XtArgs {
    const unsigned short* in_hi;      // [B][60][C] bf16 hi/lo split of the input
    const unsigned short* in_lo;
    const unsigned short* m1_hi;      // [64 m][64 k] bf16: M1^T, zero padded
    const unsigned short* m1_lo;
    const unsigned short* m2_hi;      // nullable
    const unsigned short* m2_lo;
    const float* bias;
    const float* resid;               // [B][60][C] fp32, nullable
    const float* scale;
    const float* shift;
    unsigned short* out_hi;           // [B][60][C]
    unsigned short* out_lo;
    int B, C, tiles;
    int desc_swap;                    // debugging aid: swap the LBO / SBO fields of the MN-major descriptor
};

struct __align__(8) XtBars {
    unsigned long long full[NST], empty[NST];
    unsigned long long acc1_full[2], acc1_empty[2];
    unsigned long long mid_full[2];
    unsigned long long acc2_full[2];
    uint32_t tmem_base;
};

// D=F32, A=B=BF16, N=64, M=128; bit 15: A is MN-major (cute::UMMA::InstrDescriptor a_major)
constexpr uint32_t IDESC_MN = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t IDESC_K = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ uint32_t pack2(float a, float b, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const uint32_t hu = *reinterpret_cast<const uint32_t*>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hu << 16), b - __uint_as_float(hu & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l);
    return hu;
}

template <bool TWO>
__global__ void __launch_bounds__(XT_THREADS, 1) group_transform_tc_kernel(const XtArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* mats = smem;                                   // m1h, m1l, (m2h, m2l)
    uint8_t* stages = mats + (TWO ? 4 : 2) * M_IMG;         // NST x {hi, lo} data images
    uint8_t* mids = stages + NST * 2 * D_TILE;              // TWO: 2 x {hi, lo} K-major images
    XtBars* bars = (XtBars*)(mids + (TWO ? 2 * 2 * A2_TILE : 0));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // transform matrices -> K-major SWIZZLE_128B images (row n at n*128 B, 16-byte chunk j at j ^ (n & 7))
    for (int i = threadIdx.x; i < 64 * 8; i += XT_THREADS) {
        const int r = i >> 3, j = i & 7;
        const uint32_t o = r * 128 + ((j ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(mats + o) = reinterpret_cast<const uint4*>(p.m1_hi)[i];
        *reinterpret_cast<uint4*>(mats + M_IMG + o) = reinterpret_cast<const uint4*>(p.m1_lo)[i];
        if (TWO) {
            *reinterpret_cast<uint4*>(mats + 2 * M_IMG + o) = reinterpret_cast<const uint4*>(p.m2_hi)[i];
            *reinterpret_cast<uint4*>(mats + 3 * M_IMG + o) = reinterpret_cast<const uint4*>(p.m2_lo)[i];
        }
    }
    // k rows 60..63 of every data image stay zero: rows 4..7 of the k-group-7 atom of both channel blocks
    for (int i = threadIdx.x; i < NST * 2 * 2 * 32; i += XT_THREADS) {
        const int img = i / 64, rem = i % 64, nb = rem / 32, q = rem % 32;          // 32 x 16 B = rows 4..7 of one atom
        *reinterpret_cast<uint4*>(stages + img * D_TILE + (nb * 8 + 7) * 1024 + 512 + q * 16) = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&bars->full[s], 128); mbar_init(&bars->empty[s], 1); }
        for (int e = 0; e < 2; ++e) {
            mbar_init(&bars->acc1_full[e], 1); mbar_init(&bars->acc1_empty[e], 128);
            mbar_init(&bars->mid_full[e], 128); mbar_init(&bars->acc2_full[e], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&bars->tmem_base)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // matrix images / zero rows -> async proxy (UMMA)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int cblocks = p.C / XCH;

    if (warp < 4) {
        // ================= producers: [60 k][128 c] hi/lo rows -> MN-major swizzled atoms =================
        uint32_t stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int b = tile / cblocks, cb = (tile - b * cblocks) * XCH;
            const size_t base = (size_t)b * YG * p.C + cb;
            uint8_t* st = stages + stage * 2 * D_TILE;
            mbar_wait(&bars->empty[stage], phase ^ 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int item = threadIdx.x + 128 * i;               // 60 rows x 16 chunks
                if (item < YG * 16) {
                    const int k = item >> 4, j16 = item & 15, nb = j16 >> 3, j = j16 & 7;
                    const uint32_t o = (nb * 8 + (k >> 3)) * 1024 + (k & 7) * 128 + ((j ^ (k & 7)) << 4);
                    const size_t g = base + (size_t)k * p.C + j16 * 8;
                    cp_async16(st + o, p.in_hi + g, true);
                    cp_async16(st + D_TILE + o, p.in_lo + g, true);
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&bars->full[stage])) : "memory");
            if (++stage == NST) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 12) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t lbo = p.desc_swap ? 1024u : 8192u, sbo = p.desc_swap ? 8192u : 1024u;
            const uint64_t b1h = umma_desc(mats), b1l = umma_desc(mats + M_IMG);
            const uint64_t b2h = umma_desc(mats + 2 * M_IMG), b2l = umma_desc(mats + 3 * M_IMG);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.tiles + (TWO ? gridDim.x : 0); tile += gridDim.x, ++it) {
                if (tile < p.tiles) {
                    // product 1 of tile `it`
                    const int e = it & 1;
                    const uint32_t par = (uint32_t)(it >> 1) & 1u;
                    mbar_wait(&bars->acc1_empty[e], par ^ 1);
                    mbar_wait(&bars->full[stage], phase);
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // cp.async data -> UMMA (async proxy)
                    tc_fence_after();
                    const uint8_t* st = stages + stage * 2 * D_TILE;
                    const uint64_t ah = umma_desc_mn(st, lbo, sbo), al = umma_desc_mn(st + D_TILE, lbo, sbo);
                    const uint32_t d = tmem_base + e * 64;
#pragma unroll
                    for (uint32_t ks = 0; ks < 4; ++ks) {
                        const uint64_t aadv = (uint64_t)(ks * 128);            // two 8-row K groups = 2048 B, in 16-byte units
                        const uint64_t badv = (uint64_t)(ks * 2);              // 32 bytes
                        tc_mma(d, ah + aadv, b1h + badv, IDESC_MN, ks ? 1u : 0u);
                        tc_mma(d, al + aadv, b1h + badv, IDESC_MN, 1u);
                        tc_mma(d, ah + aadv, b1l + badv, IDESC_MN, 1u);
                    }
                    tc_commit(&bars->empty[stage]);
                    tc_commit(&bars->acc1_full[e]);
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
                if (TWO && it > 0) {
                    // product 2 of tile `it - 1` (its pointwise stage ran while product 1 of tile `it` was issued)
                    const int e = (it - 1) & 1;
                    const uint32_t par = (uint32_t)((it - 1) >> 1) & 1u;
                    mbar_wait(&bars->mid_full[e], par);
                    tc_fence_after();
                    const uint8_t* md = mids + e * 2 * A2_TILE;
                    const uint64_t ah = umma_desc(md), al = umma_desc(md + A2_TILE);
                    const uint32_t d = tmem_base + 128 + e * 64;
#pragma unroll
                    for (uint32_t ks = 0; ks < 4; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);
                        tc_mma(d, ah + adv, b2h + adv, IDESC_K, ks ? 1u : 0u);
                        tc_mma(d, al + adv, b2h + adv, IDESC_K, 1u);
                        tc_mma(d, ah + adv, b2l + adv, IDESC_K, 1u);
                    }
                    tc_commit(&bars->acc2_full[e]);
                }
            }
        }
    } else {
        // ================= epilogue sets: warps 4-7 take even local tiles, warps 8-11 odd ones =================
        const int e = (warp - 4) >> 2;
        const int q = warp & 3;                              // TMEM lane quadrant this warp may read
        const int cl = q * 32 + lane;                        // channel of this thread inside the tile
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != e) continue;
            const uint32_t par = (uint32_t)(it >> 1) & 1u;
            const int b = tile / cblocks, cb = (tile - b * cblocks) * XCH;
            const int c = cb + cl;
            const float bias = p.bias ? __ldg(p.bias + c) : 0.f;
            const float sc = p.scale ? __ldg(p.scale + c) : 1.f, sh = p.scale ? __ldg(p.shift + c) : 0.f;
            const size_t row0 = (size_t)b * YG * p.C + c;    // element offset of (b, m = 0, c)
            float rcur[16];
            if (p.resid) {
#pragma unroll
                for (int i = 0; i < 16; ++i) rcur[i] = __ldg(p.resid + row0 + (size_t)i * p.C);
            }
            mbar_wait(&bars->acc1_full[e], par);
            tc_fence_after();
            uint8_t* md = mids + e * 2 * A2_TILE;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t v[16];
                tmem_ld16(tmem_base + e * 64 + ch * 16 + lane_off, v);
                float rnext[16];
                if (p.resid && ch < 3) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int m = (ch + 1) * 16 + i;
                        rnext[i] = m < YG ? __ldg(p.resid + row0 + (size_t)m * p.C) : 0.f;
                    }
                }
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x = __uint_as_float(v[i]) + bias;
                    if (p.resid) x += rcur[i];
                    if (p.scale) x = fmaxf(fmaf(x, sc, sh), 0.f);
                    f[i] = (ch * 16 + i) < YG ? x : 0.f;
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) hi[i] = pack2(f[2 * i], f[2 * i + 1], lo[i]);
                if (TWO) {
                    // row `cl` of the K-major image: 16 values = chunks 2ch, 2ch+1 of the 128-byte row
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t o = cl * 128 + (((2 * ch + h) ^ (cl & 7)) << 4);
                        *reinterpret_cast<uint4*>(md + o) = make_uint4(hi[4 * h], hi[4 * h + 1], hi[4 * h + 2], hi[4 * h + 3]);
                        *reinterpret_cast<uint4*>(md + A2_TILE + o) = make_uint4(lo[4 * h], lo[4 * h + 1], lo[4 * h + 2], lo[4 * h + 3]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int m = ch * 16 + i;
                        if (m < YG) {
                            const size_t o = row0 + (size_t)m * p.C;
                            p.out_hi[o] = (unsigned short)(i & 1 ? hi[i >> 1] >> 16 : hi[i >> 1] & 0xffffu);
                            p.out_lo[o] = (unsigned short)(i & 1 ? lo[i >> 1] >> 16 : lo[i >> 1] & 0xffffu);
                        }
                    }
                }
                if (p.resid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) rcur[i] = rnext[i];
                }
            }
            tc_fence_before();
            mbar_arrive(&bars->acc1_empty[e]);
            if (TWO) {
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the K-major image (generic stores) -> UMMA
                mbar_arrive(&bars->mid_full[e]);
                mbar_wait(&bars->acc2_full[e], par);
                tc_fence_after();
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t v[16];
                    tmem_ld16(tmem_base + 128 + e * 64 + ch * 16 + lane_off, v);
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) hi[i] = pack2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]), lo[i]);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int m = ch * 16 + i;
                        if (m < YG) {
                            const size_t o = row0 + (size_t)m * p.C;
                            p.out_hi[o] = (unsigned short)(i & 1 ? hi[i >> 1] >> 16 : hi[i >> 1] & 0xffffu);
                            p.out_lo[o] = (unsigned short)(i & 1 ? lo[i >> 1] >> 16 : lo[i >> 1] & 0xffffu);
                        }
                    }
                }
                tc_fence_before();       // orders these TMEM reads before the next mid_full arrive (which lets product 2 overwrite)
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

template <bool TWO>
constexpr size_t xt_smem() { return 1024 + (size_t)(TWO ? 4 : 2) * M_IMG + (size_t)NST * 2 * D_TILE + (TWO ? 2 * 2 * A2_TILE : 0) + sizeof(XtBars) + 64; }

}  // namespace

int group_transform_tc(yoho_ctx* ctx, const void* in_hi, const void* in_lo, int B, int C, const void* m1_hi, const void* m1_lo, const void* m2_hi,
                       const void* m2_lo, const float* bias, const float* resid, const float* scale, const float* shift,
                       void* out_hi, void* out_lo, cudaStream_t st) {
    YARG(C % XCH == 0 && B > 0 && in_hi && in_lo && m1_hi && m1_lo && out_hi && out_lo);
    XtArgs p{(const unsigned short*)in_hi, (const unsigned short*)in_lo, (const unsigned short*)m1_hi, (const unsigned short*)m1_lo,
             (const unsigned short*)m2_hi, (const unsigned short*)m2_lo, bias, resid, scale, shift, (unsigned short*)out_hi, (unsigned short*)out_lo,
             B, C, B * (C / XCH), (ctx->tc_flags & 512) ? 1 : 0};
    const int grid = p.tiles < ctx->num_sms ? p.tiles : ctx->num_sms;
    if (m2_hi) {
        YCHECK(cudaFuncSetAttribute(group_transform_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xt_smem<true>()));
        group_transform_tc_kernel<true><<<grid, XT_THREADS, xt_smem<true>(), st>>>(p);
    } else {
        YCHECK(cudaFuncSetAttribute(group_transform_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xt_smem<false>()));
        group_transform_tc_kernel<false><<<grid, XT_THREADS, xt_smem<false>(), st>>>(p);
    }
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
