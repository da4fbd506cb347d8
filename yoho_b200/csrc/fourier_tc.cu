// Group-Fourier transforms on tcgen05 (same contract as fourier_mma.cu / fourier.cu):
//     mid[m][c] = sum_k M1[k][m] in[k][c] ;  pointwise (bias / shortcut / BN+ReLU) ;  out[m][c] = sum_k M2[k][m] mid[k][c]
// computed TRANSPOSED so that the 128 channels of a tile are the UMMA M axis (= TMEM lanes, one channel per epilogue thread):
//     D1[c][m] = sum_k X[k][c] M1t[m][k]        A = X, 128 channels x 64 k, MN-major (channels contiguous: the [k][c] rows of the
//                                               activation tensor land in shared memory unchanged), B = M1t [64 m][64 k] K-major
//     D2[c][m'] = sum_m mid[c][m] M2t[m'][m]     A = the pointwise result, written by the epilogue as a K-major image, B = M2t
// Each product is 3 bf16 MMAs (hi*hi, lo*hi, hi*lo), M128 x N64 x K16, four K steps.  The transform is HBM-bound (60x60 per
// channel): a persistent CTA per SM streams (keypoint, 128-channel) tiles through a 3-stage cp.async ring; warps 0-3 load,
// warp 12 issues the MMAs, warps 4-11 are the epilogue.  One-product kernels: two sets of four warps (4-7, 8-11) alternate
// tiles.  Two-product kernel: all eight warps share every tile (half of the coefficient columns each) and are software
// pipelined — pointwise stage of tile i, then the stores of tile i-1 — so nobody waits for the round trip through product 2.
#include <cuda_bf16.h>
#include "common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int XT_THREADS = 448;               // 4 producer warps, 8 epilogue warps, product-1 issuer, product-2 issuer
constexpr int XCH = 128;                     // channels per tile (UMMA M)
constexpr int D_TILE = 64 * XCH * 2;         // bytes of one [64 k][128 c] bf16 operand image (hi or lo): 2 x 8 atoms of 1 KB
constexpr int M_IMG = 64 * 64 * 2;           // bytes of one [64][64] bf16 matrix image
constexpr int A2_TILE = XCH * 64 * 2;        // bytes of one [128 c][64 k] K-major image (hi or lo)
constexpr int NST_MAX = 4;
constexpr int NACC1_MAX = 4;

struct XtArgs {
    const unsigned short* in_hi;      // [B][60][C] bf16 hi/lo split of the input
    const unsigned short* in_lo;
    const unsigned short* in2_hi;     // FS: a second input of the same shape, ADDED to the first before product 1 (a shortcut kept
    const unsigned short* in2_lo;     //     in the Fourier domain: the transform is linear, so its images simply join the accumulation)
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
    int reserved;
};

struct __align__(8) XtBars {
    unsigned long long full[NST_MAX], empty[NST_MAX];
    unsigned long long acc1_full[NACC1_MAX], acc1_empty[NACC1_MAX];
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

// RES: a [60][128] FP32 shortcut tile rides in the stage ring next to the operand images (the epilogue reads it from shared
// memory, one conflict-free word per lane, instead of 60 dependent global loads per thread).  C is a template parameter so
// that every store address is base + immediate.
template <bool RES, bool FS>
struct StageBytes { static constexpr int value = (FS ? 4 : 2) * D_TILE + (RES ? YG * XCH * 4 : 0); };

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

// one value -> bf16 hi / lo halves, stored to element `o` of the two output tensors
__device__ __forceinline__ void store_split(unsigned short* __restrict__ oh, unsigned short* __restrict__ ol, size_t o, float x) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    oh[o] = __bfloat16_as_ushort(h);
    ol[o] = __bfloat16_as_ushort(l);
}

// Epilogue of the two-product kernel.  Warp (q = warp % 4, H = (warp - 4) / 4) owns channels q*32.. (its TMEM lane quadrant)
// and coefficient columns 32H..32H+31 (H = 1: 28 valid ones).  Software pipelined: iteration `it` runs the pointwise stage of
// tile `it` (accumulator 1 -> K-major image for product 2) and THEN the stores of tile `it - 1` (accumulator 2), so the round
// trip through the product-2 issuer is hidden behind the previous tile's stores.  H and C are compile-time: every store is
// base register + immediate, no per-element predicates.  Bias and the folded BN (scale, shift) are mandatory here.
template <int H, int C, int NACC1>
__device__ __forceinline__ void two_epilogue(const XtArgs& p, XtBars* bars, uint8_t* mids, uint32_t tmem_base, uint32_t acc2_col, int q, int lane) {
    constexpr int NV = H == 0 ? 32 : YG - 32;            // valid coefficient columns in this half
    constexpr int cblocks = C / XCH;
    const int cl = q * 32 + lane;
    const uint32_t lane_off = ((uint32_t)(q * 32) << 16) + (uint32_t)(32 * H);
    unsigned short* oh_prev = nullptr;
    unsigned short* ol_prev = nullptr;
    int it = 0;
    auto store_prev = [&](int pit) {
        const int e = pit & 1;
        uint32_t v[32];
        mbar_wait(&bars->acc2_full[e], (uint32_t)(pit >> 1) & 1u);
        tc_fence_after();
        tmem_ld32_nowait(tmem_base + acc2_col + e * 64 + lane_off, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        tc_fence_before();       // orders these TMEM reads before this thread's next mid_full arrive (-> product 2 may overwrite)
        unsigned short* oh = oh_prev + (size_t)(32 * H) * C;
        unsigned short* ol = ol_prev + (size_t)(32 * H) * C;
#pragma unroll
        for (int m = 0; m < NV; ++m) store_split(oh, ol, (size_t)m * C, __uint_as_float(v[m]));
    };
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        const int e = it & 1;
        const int b = tile / cblocks, cb = (tile - b * cblocks) * XCH;
        const int c = cb + cl;
        const float bias = __ldg(p.bias + c), sc = __ldg(p.scale + c), sh = __ldg(p.shift + c);
        const int a = it % NACC1;
        mbar_wait(&bars->acc1_full[a], (uint32_t)(it / NACC1) & 1u);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32_nowait(tmem_base + a * 64 + lane_off, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        tc_fence_before();
        mbar_arrive(&bars->acc1_empty[a]);               // this thread's part of accumulator 1 is in registers
#pragma unroll
        for (int m = 0; m < 32; ++m)
            v[m] = m < NV ? __float_as_uint(fmaxf(fmaf(__uint_as_float(v[m]) + bias, sc, sh), 0.f)) : 0u;
        // row `cl` of the K-major image: this thread's 32 values = chunks 4H..4H+3 of the 128-byte row (hi and lo)
        uint8_t* md = mids + e * 2 * A2_TILE + cl * 128;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) hi[i] = pack2(__uint_as_float(v[8 * jj + 2 * i]), __uint_as_float(v[8 * jj + 2 * i + 1]), lo[i]);
            const uint32_t o = ((uint32_t)(4 * H + jj) ^ (uint32_t)(cl & 7)) << 4;
            *reinterpret_cast<uint4*>(md + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(md + A2_TILE + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the K-major image (generic stores) -> UMMA
        mbar_arrive(&bars->mid_full[e]);
        if (it > 0) store_prev(it - 1);
        oh_prev = p.out_hi + (size_t)b * YG * C + c;     // (b, m = 0, c); row m is at + m * C (immediate offsets)
        ol_prev = p.out_lo + (size_t)b * YG * C + c;
    }
    if (it > 0) store_prev(it - 1);
}

template <bool TWO, bool RES, bool FS, int C>
__global__ void __launch_bounds__(XT_THREADS, 1) group_transform_tc_kernel(const XtArgs p) {
    static_assert(!(TWO && RES), "the two-product kernel carries no FP32 shortcut tile (its shortcut is a Fourier-domain input: FS)");
    static_assert(!FS || TWO, "FS is a variant of the two-product kernel");
    constexpr int STAGE = StageBytes<RES, FS>::value;
    constexpr int IMGS = FS ? 4 : 2;                     // operand images per stage
    // The two-product kernel is a longer pipeline (pointwise stage + second product between load and store): product 1 runs up
    // to four tiles ahead of the epilogue into four TMEM accumulators, so a stage is released as soon as its tile has LANDED
    // and been multiplied — the load ring (4 stages) stays in flight independently of the epilogue's progress.
    constexpr int NST = FS ? 2 : TWO ? 4 : 3;            // FS: two 64 KB stages (same bytes in flight as four 32 KB ones)
    constexpr int NACC1 = TWO ? 4 : 2;
    constexpr uint32_t ACC2_COL = NACC1 * 64;
    constexpr uint32_t TMEM_COLS = TWO ? 512 : 128;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* mats = smem;                                   // m1h, m1l, (m2h, m2l)
    uint8_t* stages = mats + (TWO ? 4 : 2) * M_IMG;         // NST x {hi image, lo image, (shortcut tile)}
    uint8_t* mids = stages + NST * STAGE;                   // TWO: 2 x {hi, lo} K-major images
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
    for (int i = threadIdx.x; i < NST * IMGS * 2 * 32; i += XT_THREADS) {
        const int img = i / 64, rem = i % 64, nb = rem / 32, q = rem % 32;          // 32 x 16 B = rows 4..7 of one atom
        *reinterpret_cast<uint4*>(stages + (img / IMGS) * STAGE + (img % IMGS) * D_TILE + (nb * 8 + 7) * 1024 + 512 + q * 16) = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&bars->full[s], 128);
            mbar_init(&bars->empty[s], RES ? 1 + 128 : 1);    // tcgen05.commit (+ the epilogue threads that read the shortcut tile)
        }
        for (int e = 0; e < NACC1; ++e) { mbar_init(&bars->acc1_full[e], 1); mbar_init(&bars->acc1_empty[e], TWO ? 256 : 128); }
        for (int e = 0; e < 2; ++e) { mbar_init(&bars->mid_full[e], 256); mbar_init(&bars->acc2_full[e], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&bars->tmem_base)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // matrix images / zero rows -> async proxy (UMMA)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    constexpr int cblocks = C / XCH;

    if (warp < 4) {
        // ================= producers: [60 k][128 c] hi/lo rows -> MN-major swizzled atoms (+ the shortcut rows) =================
        uint32_t stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int b = tile / cblocks, cb = (tile - b * cblocks) * XCH;
            const size_t base = (size_t)b * YG * C + cb;
            uint8_t* st = stages + stage * STAGE;
            mbar_wait(&bars->empty[stage], phase ^ 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int item = threadIdx.x + 128 * i;               // 60 rows x 16 chunks
                if (item < YG * 16) {
                    const int k = item >> 4, j16 = item & 15, nb = j16 >> 3, j = j16 & 7;
                    const uint32_t o = (nb * 8 + (k >> 3)) * 1024 + (k & 7) * 128 + ((j ^ (k & 7)) << 4);
                    const size_t g = base + (size_t)k * C + j16 * 8;
                    cp_async16(st + o, p.in_hi + g, true);
                    cp_async16(st + D_TILE + o, p.in_lo + g, true);
                    if (FS) {
                        cp_async16(st + 2 * D_TILE + o, p.in2_hi + g, true);
                        cp_async16(st + 3 * D_TILE + o, p.in2_lo + g, true);
                    }
                }
            }
            if (RES) {
#pragma unroll
                for (int i = 0; i < 15; ++i) {
                    const int item = threadIdx.x + 128 * i;           // 60 rows x 32 chunks of 4 floats
                    const int k = item >> 5, j = item & 31;
                    cp_async16(st + 2 * D_TILE + k * (XCH * 4) + j * 16, p.resid + base + (size_t)k * C + j * 4, true);
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&bars->full[stage])) : "memory");
            if (++stage == NST) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 12) {
        // ================= product-1 issuer =================
        if (lane == 0) {
            const uint64_t b1h = umma_desc(mats), b1l = umma_desc(mats + M_IMG);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
                const int a = it % NACC1;
                mbar_wait(&bars->acc1_empty[a], ((uint32_t)(it / NACC1) & 1u) ^ 1u);
                mbar_wait(&bars->full[stage], phase);
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // cp.async data -> UMMA (async proxy)
                tc_fence_after();
                const uint8_t* st = stages + stage * STAGE;
                // MN-major operand: 64-channel blocks 8192 B apart (LBO), 8-row k groups 1024 B apart (SBO)
                const uint64_t ah = umma_desc_mn(st, 8192u, 1024u), al = umma_desc_mn(st + D_TILE, 8192u, 1024u);
                const uint64_t a2h = umma_desc_mn(st + 2 * D_TILE, 8192u, 1024u), a2l = umma_desc_mn(st + 3 * D_TILE, 8192u, 1024u);
                const uint32_t d = tmem_base + a * 64;
#pragma unroll
                for (uint32_t ks = 0; ks < 4; ++ks) {
                    const uint64_t aadv = (uint64_t)(ks * 128);            // two 8-row K groups = 2048 B, in 16-byte units
                    const uint64_t badv = (uint64_t)(ks * 2);              // 32 bytes
                    tc_mma(d, ah + aadv, b1h + badv, IDESC_MN, ks ? 1u : 0u);
                    tc_mma(d, al + aadv, b1h + badv, IDESC_MN, 1u);
                    tc_mma(d, ah + aadv, b1l + badv, IDESC_MN, 1u);
                    if (FS) {
                        tc_mma(d, a2h + aadv, b1h + badv, IDESC_MN, 1u);
                        tc_mma(d, a2l + aadv, b1h + badv, IDESC_MN, 1u);
                        tc_mma(d, a2h + aadv, b1l + badv, IDESC_MN, 1u);
                    }
                }
                tc_commit(&bars->empty[stage]);
                tc_commit(&bars->acc1_full[a]);
                if (++stage == NST) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 13) {
        // ================= product-2 issuer (its own thread: never queued behind a product 1 that waits for a load) =================
        if (TWO && lane == 0) {
            const uint64_t b2h = umma_desc(mats + 2 * M_IMG), b2l = umma_desc(mats + 3 * M_IMG);
            int it = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
                const int e = it & 1;
                mbar_wait(&bars->mid_full[e], (uint32_t)(it >> 1) & 1u);
                tc_fence_after();
                const uint8_t* md = mids + e * 2 * A2_TILE;
                const uint64_t ah = umma_desc(md), al = umma_desc(md + A2_TILE);
                const uint32_t d = tmem_base + ACC2_COL + e * 64;
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
    } else if constexpr (TWO) {
        // ================= epilogue, two products: all eight warps work on EVERY tile (two_epilogue below) =================
        if (warp < 8) two_epilogue<0, C, NACC1>(p, bars, mids, tmem_base, ACC2_COL, warp & 3, lane);
        else two_epilogue<1, C, NACC1>(p, bars, mids, tmem_base, ACC2_COL, warp & 3, lane);
    } else {
        // ================= epilogue sets, one product: warps 4-7 take even local tiles, warps 8-11 odd ones =================
        const int e = (warp - 4) >> 2;
        const int q = warp & 3;                              // TMEM lane quadrant this warp may read
        const int cl = q * 32 + lane;                        // channel of this thread inside the tile
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const bool act = p.scale != nullptr;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != e) continue;
            const uint32_t par = (uint32_t)(it >> 1) & 1u;
            const int b = tile / cblocks, cb = (tile - b * cblocks) * XCH;
            const int c = cb + cl;
            const float bias = p.bias ? __ldg(p.bias + c) : 0.f;
            const float sc = act ? __ldg(p.scale + c) : 1.f, sh = act ? __ldg(p.shift + c) : 0.f;
            unsigned short* oh = p.out_hi + (size_t)b * YG * C + c;      // (b, m = 0, c); row m is at + m * C (immediate offsets)
            unsigned short* ol = p.out_lo + (size_t)b * YG * C + c;
            const int stage = it % NST;
            const float* rs = (const float*)(stages + stage * STAGE + 2 * D_TILE) + cl;
            mbar_wait(&bars->acc1_full[e], par);
            tc_fence_after();
            uint32_t v[64];
            tmem_ld32_nowait(tmem_base + e * 64 + lane_off, v);
            tmem_ld32_nowait(tmem_base + e * 64 + 32 + lane_off, v + 32);
            if (RES) mbar_wait(&bars->full[stage], (uint32_t)(it / NST) & 1u);   // the shortcut tile has landed (same barrier the MMA waited on)
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            tc_fence_before();
            mbar_arrive(&bars->acc1_empty[e]);               // accumulator 1 is in registers
#pragma unroll
            for (int m = 0; m < YG; ++m) {
                float x = __uint_as_float(v[m]) + bias;
                if (RES) x += rs[m * XCH];
                if (act) x = fmaxf(fmaf(x, sc, sh), 0.f);
                store_split(oh, ol, (size_t)m * C, x);
            }
            if (RES) mbar_arrive(&bars->empty[stage]);       // shortcut tile consumed: the stage may be refilled
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Output side of the all-Fourier PartI on the same machinery (tuning flag 2048; the FP32 SIMT twin is part1_finalize_fourier_kernel):
//     e[c][g] = bias4[c] + sum_m F[m][g] Y4[m][c] + x[c][g] ;  eqv = e / max(||e[:, g]||, 1e-4) ;  inv, desc as in part1.cu
// Y4 arrives as a bf16 hi/lo pair [B][60][32].  A tile is FOUR keypoints x 32 channels = the 128 lanes of the UMMA M axis, so an
// epilogue warp (TMEM lane quadrant q) holds one keypoint with one channel per lane and the 60 group elements as 60 columns:
// the channel norms are warp-shuffle sums, the pools over the group axis are per-thread sums over registers, and every thread
// reads / writes the 60 contiguous floats of its (keypoint, channel) row of x / eqv.  One product (inverse transform), same
// producer / issuer / alternating epilogue sets as the one-product transform kernel.
struct FinArgs {
    const unsigned short* in_hi;      // [B][60][32] bf16 hi/lo of the layer-4 Fourier coefficients
    const unsigned short* in_lo;
    const unsigned short* m1_hi;      // inverse transform, [64 g][64 m] bf16 hi/lo
    const unsigned short* m1_lo;
    const float* bias4;               // [32]
    const float* x;                   // [B][32][60] PartI input (the outer residual)
    float* eqv;                       // [B][32][60]
    float* inv;                       // [B][32] or nullptr
    float* desc;                      // [B][32] or nullptr
    int B, tiles;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// STAGED (tuning flag 4096 on top of 2048, NOT yet run on hardware): the x / eqv rows of a keypoint (7680 contiguous bytes) pass
// through a per-warp shared-memory tile, so the global accesses are lane-contiguous 16-byte pieces instead of 32 rows x 16 bytes
// per warp instruction (the measured bound of the unstaged variant: 68 us against a 20 us HBM floor).
template <bool STAGED>
__global__ void __launch_bounds__(XT_THREADS, 1) group_finalize_tc_kernel(const FinArgs p) {
    constexpr int NST = 3;
    constexpr int STAGE = 2 * D_TILE;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* mats = smem;
    uint8_t* stages = mats + 2 * M_IMG;
    uint8_t* xtiles = stages + NST * STAGE;                  // STAGED: 8 x 7680 B, one [32][60] FP32 tile per epilogue warp
    XtBars* bars = (XtBars*)(xtiles + (STAGED ? 8 * YF * YG * 4 : 0));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 8; i += XT_THREADS) {
        const int r = i >> 3, j = i & 7;
        const uint32_t o = r * 128 + ((j ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(mats + o) = reinterpret_cast<const uint4*>(p.m1_hi)[i];
        *reinterpret_cast<uint4*>(mats + M_IMG + o) = reinterpret_cast<const uint4*>(p.m1_lo)[i];
    }
    for (int i = threadIdx.x; i < NST * 2 * 2 * 32; i += XT_THREADS) {          // k rows 60..63 stay zero
        const int img = i / 64, rem = i % 64, nb = rem / 32, q = rem % 32;
        *reinterpret_cast<uint4*>(stages + (img / 2) * STAGE + (img % 2) * D_TILE + (nb * 8 + 7) * 1024 + 512 + q * 16) = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&bars->full[s], 128); mbar_init(&bars->empty[s], 1); }
        for (int e = 0; e < 2; ++e) { mbar_init(&bars->acc1_full[e], 1); mbar_init(&bars->acc1_empty[e], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&bars->tmem_base)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < 4) {
        // producers: tile = keypoints 4*tile .. 4*tile+3; 16-byte chunk j16 of coefficient row k holds channels 8*(j16 % 4).. of
        // keypoint j16 / 4 (missing keypoints of the last tile are zero filled)
        uint32_t stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            uint8_t* st = stages + stage * STAGE;
            mbar_wait(&bars->empty[stage], phase ^ 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int item = threadIdx.x + 128 * i;               // 60 rows x 16 chunks
                if (item < YG * 16) {
                    const int k = item >> 4, j16 = item & 15, nb = j16 >> 3, j = j16 & 7;
                    const uint32_t o = (nb * 8 + (k >> 3)) * 1024 + (k & 7) * 128 + ((j ^ (k & 7)) << 4);
                    const int b = tile * 4 + (j16 >> 2);
                    const bool ok = b < p.B;
                    const size_t g = ((size_t)(ok ? b : 0) * YG + k) * YF + (j16 & 3) * 8;
                    cp_async16(st + o, p.in_hi + g, ok);
                    cp_async16(st + D_TILE + o, p.in_lo + g, ok);
                }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&bars->full[stage])) : "memory");
            if (++stage == NST) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 12) {
        if (lane == 0) {
            const uint64_t b1h = umma_desc(mats), b1l = umma_desc(mats + M_IMG);
            uint32_t stage = 0, phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
                const int a = it & 1;
                mbar_wait(&bars->acc1_empty[a], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                mbar_wait(&bars->full[stage], phase);
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                tc_fence_after();
                const uint8_t* st = stages + stage * STAGE;
                const uint64_t ah = umma_desc_mn(st, 8192u, 1024u), al = umma_desc_mn(st + D_TILE, 8192u, 1024u);
                const uint32_t d = tmem_base + a * 64;
#pragma unroll
                for (uint32_t ks = 0; ks < 4; ++ks) {
                    const uint64_t aadv = (uint64_t)(ks * 128), badv = (uint64_t)(ks * 2);
                    tc_mma(d, ah + aadv, b1h + badv, IDESC_MN, ks ? 1u : 0u);
                    tc_mma(d, al + aadv, b1h + badv, IDESC_MN, 1u);
                    tc_mma(d, ah + aadv, b1l + badv, IDESC_MN, 1u);
                }
                tc_commit(&bars->empty[stage]);
                tc_commit(&bars->acc1_full[a]);
                if (++stage == NST) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < 12) {
        const int e = (warp - 4) >> 2;                       // epilogue set: even / odd local tiles
        const int q = warp & 3;                              // TMEM lane quadrant = keypoint of the tile; lane = channel
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float b4 = __ldg(p.bias4 + lane);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != e) continue;
            const int b = tile * 4 + q;
            const bool ok = b < p.B;                         // warp-uniform
            float* sx = reinterpret_cast<float*>(xtiles) + (warp - 4) * (YF * YG);
            if (STAGED && ok) {                              // this keypoint's x tile, lane-contiguous, while the product is still running
                const float4* gx = reinterpret_cast<const float4*>(p.x + (size_t)b * YF * YG);
#pragma unroll
                for (int t4 = 0; t4 < YF * YG / 4 / 32; ++t4) reinterpret_cast<float4*>(sx)[t4 * 32 + lane] = gx[t4 * 32 + lane];
                __syncwarp();
            }
            mbar_wait(&bars->acc1_full[e], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            uint32_t v[64];
            tmem_ld32_nowait(tmem_base + e * 64 + lane_off, v);
            tmem_ld32_nowait(tmem_base + e * 64 + 32 + lane_off, v + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            tc_fence_before();
            mbar_arrive(&bars->acc1_empty[e]);
            if (!ok) continue;
            const size_t row = ((size_t)b * YF + lane) * YG;              // this thread's 60 contiguous floats of x / eqv
            float ev[YG];
#pragma unroll
            for (int g4 = 0; g4 < YG / 4; ++g4) {
                const float4 xv = STAGED ? *reinterpret_cast<const float4*>(sx + lane * YG + 4 * g4)
                                         : *reinterpret_cast<const float4*>(p.x + row + 4 * g4);
                ev[4 * g4 + 0] = (__uint_as_float(v[4 * g4 + 0]) + b4) + xv.x;
                ev[4 * g4 + 1] = (__uint_as_float(v[4 * g4 + 1]) + b4) + xv.y;
                ev[4 * g4 + 2] = (__uint_as_float(v[4 * g4 + 2]) + b4) + xv.z;
                ev[4 * g4 + 3] = (__uint_as_float(v[4 * g4 + 3]) + b4) + xv.w;
            }
            // invariant pooling uses the UN-normalised e (utils/network.py:99 precedes :102)
            if (p.inv) {
                float sum = 0.f;
#pragma unroll
                for (int g = 0; g < YG; ++g) sum += ev[g];
                const float mean = sum / 60.0f;
                const float ss = warp_sum(mean * mean);
                p.inv[(size_t)b * YF + lane] = mean / fmaxf(sqrtf(ss), 1e-4f);
            }
#pragma unroll
            for (int g = 0; g < YG; ++g) {
                const float ss = warp_sum(ev[g] * ev[g]);                  // sum over the 32 channels of this keypoint
                ev[g] = ev[g] / fmaxf(sqrtf(ss), 1e-4f);                    // torch.clamp_min(torch.norm(eqv, dim=1), 1e-4)
            }
            if (STAGED) {
                __syncwarp();                                // every lane has read its x row
#pragma unroll
                for (int g4 = 0; g4 < YG / 4; ++g4)
                    *reinterpret_cast<float4*>(sx + lane * YG + 4 * g4) = make_float4(ev[4 * g4], ev[4 * g4 + 1], ev[4 * g4 + 2], ev[4 * g4 + 3]);
                __syncwarp();
                float4* ge = reinterpret_cast<float4*>(p.eqv + (size_t)b * YF * YG);
#pragma unroll
                for (int t4 = 0; t4 < YF * YG / 4 / 32; ++t4) ge[t4 * 32 + lane] = reinterpret_cast<const float4*>(sx)[t4 * 32 + lane];
                __syncwarp();                                // the tile is free for this warp's next keypoint
            } else {
#pragma unroll
                for (int g4 = 0; g4 < YG / 4; ++g4)
                    *reinterpret_cast<float4*>(p.eqv + row + 4 * g4) = make_float4(ev[4 * g4], ev[4 * g4 + 1], ev[4 * g4 + 2], ev[4 * g4 + 3]);
            }
            if (p.desc) {
                // numpy's float32 pairwise mean of 60 contiguous values (np.mean(feats, axis=-1), tests/matcher.py:35)
                float r8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) r8[j] = ev[j];
#pragma unroll
                for (int i = 8; i < 56; i += 8)
#pragma unroll
                    for (int j = 0; j < 8; ++j) r8[j] = __fadd_rn(r8[j], ev[i + j]);
                float res = __fadd_rn(__fadd_rn(__fadd_rn(r8[0], r8[1]), __fadd_rn(r8[2], r8[3])),
                                      __fadd_rn(__fadd_rn(r8[4], r8[5]), __fadd_rn(r8[6], r8[7])));
#pragma unroll
                for (int i = 56; i < 60; ++i) res = __fadd_rn(res, ev[i]);
                p.desc[(size_t)b * YF + lane] = __fdiv_rn(res, 60.0f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(128) : "memory");
    }
}

template <bool TWO, bool RES, bool FS>
constexpr size_t xt_smem() { return 1024 + (size_t)(TWO ? 4 : 2) * M_IMG + (size_t)(FS ? 2 : TWO ? 4 : 3) * StageBytes<RES, FS>::value + (TWO ? 2 * 2 * A2_TILE : 0) + sizeof(XtBars) + 64; }
static_assert(xt_smem<true, false, true>() <= 232448 && xt_smem<true, false, false>() <= 232448 && xt_smem<false, true, false>() <= 232448, "227 KB per CTA");

template <bool TWO, bool RES, bool FS, int C>
int xt_launch(yoho_ctx* ctx, const XtArgs& p, cudaStream_t st) {
    YCHECK(cudaFuncSetAttribute(group_transform_tc_kernel<TWO, RES, FS, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xt_smem<TWO, RES, FS>()));
    const int grid = p.tiles < ctx->num_sms ? p.tiles : ctx->num_sms;
    group_transform_tc_kernel<TWO, RES, FS, C><<<grid, XT_THREADS, xt_smem<TWO, RES, FS>(), st>>>(p);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

}  // namespace

// The shapes PartI uses are instantiated: forward only (C=256), inverse -> activation -> forward (C=256, 512), the latter with a
// Fourier-domain shortcut input (C=256), inverse + FP32 shortcut -> activation (C=256).  The caller falls back to the warp-MMA
// kernel for any other shape.
bool group_transform_tc_supported(int C, bool two, bool res) { return (C == 256 && !(two && res)) || (C == 512 && two && !res); }

int group_transform_tc(yoho_ctx* ctx, const void* in_hi, const void* in_lo, int B, int C, const void* m1_hi, const void* m1_lo, const void* m2_hi,
                       const void* m2_lo, const float* bias, const float* resid, const float* scale, const float* shift,
                       void* out_hi, void* out_lo, cudaStream_t st, const void* in2_hi, const void* in2_lo) {
    YARG(B > 0 && in_hi && in_lo && m1_hi && m1_lo && out_hi && out_lo && group_transform_tc_supported(C, m2_hi != nullptr, resid != nullptr));
    YARG((in2_hi == nullptr) == (in2_lo == nullptr) && (!in2_hi || (m2_hi && C == 256)));
    XtArgs p{(const unsigned short*)in_hi, (const unsigned short*)in_lo, (const unsigned short*)in2_hi, (const unsigned short*)in2_lo,
             (const unsigned short*)m1_hi, (const unsigned short*)m1_lo,
             (const unsigned short*)m2_hi, (const unsigned short*)m2_lo, bias, resid, scale, shift, (unsigned short*)out_hi, (unsigned short*)out_lo,
             B, C, B * (C / XCH), 0};
    if (m2_hi) {
        YARG(bias && scale && shift);
        if (in2_hi) return xt_launch<true, false, true, 256>(ctx, p, st);
        return C == 512 ? xt_launch<true, false, false, 512>(ctx, p, st) : xt_launch<true, false, false, 256>(ctx, p, st);
    }
    if (resid) return xt_launch<false, true, false, 256>(ctx, p, st);
    return xt_launch<false, false, false, 256>(ctx, p, st);
}

// PartI output side on tensor cores (tuning flag 2048; + 4096: shared-memory staged row accesses): see group_finalize_tc_kernel.
int group_finalize_tc(yoho_ctx* ctx, const void* y4_hi, const void* y4_lo, int B, const void* minv_hi, const void* minv_lo, const float* bias4,
                      const float* x, float* eqv, float* inv, float* desc, cudaStream_t st) {
    YARG(B > 0 && y4_hi && y4_lo && minv_hi && minv_lo && bias4 && x && eqv);
    FinArgs p{(const unsigned short*)y4_hi, (const unsigned short*)y4_lo, (const unsigned short*)minv_hi, (const unsigned short*)minv_lo,
              bias4, x, eqv, inv, desc, B, (B + 3) / 4};
    const bool staged = (ctx->tc_flags & 4096) != 0;
    const size_t smem = 1024 + 2 * M_IMG + 3 * 2 * D_TILE + (staged ? 8 * YF * YG * 4 : 0) + sizeof(XtBars) + 64;
    const int grid = p.tiles < ctx->num_sms ? p.tiles : ctx->num_sms;
    if (staged) {
        YCHECK(cudaFuncSetAttribute(group_finalize_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        group_finalize_tc_kernel<true><<<grid, XT_THREADS, smem, st>>>(p);
    } else {
        YCHECK(cudaFuncSetAttribute(group_finalize_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        group_finalize_tc_kernel<false><<<grid, XT_THREADS, smem, st>>>(p);
    }
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
