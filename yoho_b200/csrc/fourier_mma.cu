// Group-Fourier transforms on the warp-level tensor-core path (mma.sync m16n8k16, bf16 x3 split, FP32 accumulate).
// Same contract as group_transform_kernel in fourier.cu (which stays as the FP32 SIMT reference of this kernel):
//     mid[m][c] = sum_k M1[k][m] in[k][c] ;  pointwise (bias / shortcut / BN+ReLU) ;  out[m][c] = sum_k M2[k][m] mid[k][c]
// One tile = one keypoint x 128 channels; 16 warps = 4 (16 output rows each) x 4 (32 channels each); persistent CTAs.
// The transform is memory-bound (60x60 per channel), so the legacy warp MMA is enough here; the big GEMMs use tcgen05.
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int MP = 72;                // padded row length (bf16) of the 64x64 transform matrices: 144 B, conflict-free ldmatrix

struct XmArgs {
    const unsigned short* in_hi;      // [B][60][C] bf16 hi/lo split of the input
    const unsigned short* in_lo;
    const __nv_bfloat16* m1_hi;       // [64 m][64 k] bf16: M1^T (row = output index m, col = input index k), zero padded
    const __nv_bfloat16* m1_lo;
    const __nv_bfloat16* m2_hi;       // nullable
    const __nv_bfloat16* m2_lo;
    const float* bias;
    const float* resid;
    const float* scale;
    const float* shift;
    unsigned short* out_hi;           // [B][60][C]
    unsigned short* out_lo;
    int B, C;
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_hi_lo(float a, float b, uint32_t& lo) {
    // two values -> packed bf16x2 hi and lo parts (first value in the low half) with the paired conversion instruction
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const uint32_t hu = *reinterpret_cast<const uint32_t*>(&h);
    const float ha = __uint_as_float(hu << 16), hb = __uint_as_float(hu & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
    lo = *reinterpret_cast<const uint32_t*>(&l);
    return hu;
}

// Persistent: one CTA per SM loops over (keypoint, 128-channel) tiles.  The FP32 input tile (and the shortcut tile, if any) of
// the NEXT tile streams into a staging buffer with cp.async while the current tile is converted, multiplied and written
// back, so global loads are always in flight (the kernel is memory-bound: ~2 x 30 KB in, 30 KB out per tile).
constexpr int XT = 512;                       // threads: 16 warps = 4 (16 rows) x 4 (XC/4 channels)

// [60][XC] fp32 tile -> shared (row stride XC floats)
template <int XC>
__device__ __forceinline__ void stage_tile(float* dst, const float* src, int C, int t) {
    for (int i = t; i < YG * (XC / 4); i += XT) {
        const int k = i / (XC / 4), c4 = i % (XC / 4);
        cp_async16(dst + k * XC + c4 * 4, src + (size_t)k * C + c4 * 4, true);
    }
}
// [60][XC] bf16 tile -> shared operand tile (row stride XP = XC + 8)
template <int XC>
__device__ __forceinline__ void stage_bf16(__nv_bfloat16* dst, const unsigned short* src, int C, int t) {
    constexpr int XP = XC + 8;
    for (int i = t; i < YG * (XC / 8); i += XT) {
        const int k = i / (XC / 8), q = i % (XC / 8);
        cp_async16(dst + k * XP + q * 8, src + (size_t)k * C + q * 8, true);
    }
}

// acc[nt][4] (NT n-tiles of 8 channels) = M^T (rows m0..m0+15) x data (64 k x channels n_base..n_base+8*NT-1), 3 split products
template <int NT, int XP>
__device__ __forceinline__ void warp_product32(const __nv_bfloat16* mh, const __nv_bfloat16* ml, const __nv_bfloat16* xh,
                                               const __nv_bfloat16* xl, int m0, int n_base, int lane, float (&acc)[NT][4]) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
    const int arow = m0 + (lane & 15), acol = (lane >> 4) * 8;
    const int brow = (lane & 7) + ((lane >> 3) & 1) * 8, bcol = (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t ah[4], al[4];
        ldsm_x4(ah, mh + arow * MP + ks * 16 + acol);
        ldsm_x4(al, ml + arow * MP + ks * 16 + acol);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
            uint32_t bh[4], bl[4];
            const int off = (ks * 16 + brow) * XP + n_base + np * 16 + bcol;
            ldsm_x4_t(bh, xh + off);
            ldsm_x4_t(bl, xl + off);
            mma16816(acc[2 * np], ah, bh[0], bh[1]);
            mma16816(acc[2 * np], al, bh[0], bh[1]);
            mma16816(acc[2 * np], ah, bl[0], bl[1]);
            mma16816(acc[2 * np + 1], ah, bh[2], bh[3]);
            mma16816(acc[2 * np + 1], al, bh[2], bh[3]);
            mma16816(acc[2 * np + 1], ah, bl[2], bl[3]);
        }
    }
}

template <int XC>
__global__ void __launch_bounds__(XT, XC == 128 ? 1 : 2) group_transform_mma_kernel(const XmArgs p) {
    constexpr int XP = XC + 8;                 // padded row length (bf16) of the data tiles: conflict-free ldmatrix
    constexpr int NT = XC / 32;                // 8-channel n-tiles per warp
    constexpr int STG = YG * XC;               // floats of one staged shortcut tile
    extern __shared__ __align__(16) uint8_t smraw[];
    __nv_bfloat16* m1h = (__nv_bfloat16*)smraw;          // [64][MP]
    __nv_bfloat16* m1l = m1h + 64 * MP;
    __nv_bfloat16* m2h = m1l + 64 * MP;
    __nv_bfloat16* m2l = m2h + 64 * MP;
    __nv_bfloat16* xbuf = m2l + 64 * MP;                 // 2 x {hi,lo} x [64][XP]: input tiles (double-buffered), reused as output staging
    __nv_bfloat16* yh = xbuf + 4 * 64 * XP;              // [64][XP]   intermediate tile (two-stage transforms)
    __nv_bfloat16* yl = yh + 64 * XP;
    float* rstage = (float*)(yl + 64 * XP);              // [60][128] fp32 shortcut staging
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int cblocks = p.C / XC;
    const int tiles = p.B * cblocks;

    for (int i = t; i < 64 * 8; i += XT) {
        const int r = i >> 3, q = i & 7;
        *reinterpret_cast<uint4*>(m1h + r * MP + q * 8) = reinterpret_cast<const uint4*>(p.m1_hi)[i];
        *reinterpret_cast<uint4*>(m1l + r * MP + q * 8) = reinterpret_cast<const uint4*>(p.m1_lo)[i];
        if (p.m2_hi) {
            *reinterpret_cast<uint4*>(m2h + r * MP + q * 8) = reinterpret_cast<const uint4*>(p.m2_hi)[i];
            *reinterpret_cast<uint4*>(m2l + r * MP + q * 8) = reinterpret_cast<const uint4*>(p.m2_lo)[i];
        }
    }
    for (int i = t; i < 4 * XP; i += XT) {               // rows 60..63 of the data tiles stay zero
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) xbuf[s4 * 64 * XP + YG * XP + i] = __float2bfloat16_rn(0.f);
        yh[YG * XP + i] = __float2bfloat16_rn(0.f); yl[YG * XP + i] = __float2bfloat16_rn(0.f);
    }
    int tile = blockIdx.x;
    if (tile < tiles) {
        const int b = tile / cblocks, cb = (tile - b * cblocks) * XC;
        const size_t o = (size_t)b * YG * p.C + cb;
        stage_bf16<XC>(xbuf, p.in_hi + o, p.C, t);
        stage_bf16<XC>(xbuf + 64 * XP, p.in_lo + o, p.C, t);
    }
    cp_async_commit();
    const int m0 = (warp & 3) * 16, n_base = (warp >> 2) * (XC / 4);
    const int r0 = m0 + (lane >> 2), cq = 2 * (lane & 3);
    int it = 0;
    for (; tile < tiles; tile += gridDim.x, ++it) {
        const int b = tile / cblocks, cb = (tile - b * cblocks) * XC;
        __nv_bfloat16* xh = xbuf + (it & 1) * 2 * 64 * XP;
        __nv_bfloat16* xl = xh + 64 * XP;
        // group A: this tile's shortcut; group B: the next tile's input
        if (p.resid) stage_tile<XC>(rstage, p.resid + (size_t)b * YG * p.C + cb, p.C, t);
        cp_async_commit();
        const int nxt = tile + gridDim.x;
        if (nxt < tiles) {
            const int nb = nxt / cblocks, ncb = (nxt - nb * cblocks) * XC;
            const size_t o = (size_t)nb * YG * p.C + ncb;
            __nv_bfloat16* nx = xbuf + ((it + 1) & 1) * 2 * 64 * XP;
            stage_bf16<XC>(nx, p.in_hi + o, p.C, t);
            stage_bf16<XC>(nx + 64 * XP, p.in_lo + o, p.C, t);
        }
        cp_async_commit();
        cp_async_wait<2>();                                // this tile's input has landed (two younger groups may be in flight)
        __syncthreads();
        float acc[NT][4];
        warp_product32<NT, XP>(m1h, m1l, xh, xl, m0, n_base, lane, acc);
        if (p.resid) { cp_async_wait<1>(); }               // shortcut tile landed (the next input may still be in flight)
        __syncthreads();                                   // ... and every warp is done reading xh/xl
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int cl = n_base + nt * 8 + cq, c = cb + cl;
            float b0 = 0.f, b1 = 0.f, s0 = 1.f, s1 = 1.f, h0 = 0.f, h1 = 0.f;
            if (p.bias) { b0 = p.bias[c]; b1 = p.bias[c + 1]; }
            if (p.scale) { s0 = p.scale[c]; s1 = p.scale[c + 1]; h0 = p.shift[c]; h1 = p.shift[c + 1]; }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int m = r0 + 8 * hf;
                float v0 = acc[nt][2 * hf] + b0, v1 = acc[nt][2 * hf + 1] + b1;
                if (p.resid && m < YG) {
                    const float2 rr = *reinterpret_cast<const float2*>(rstage + m * XC + cl);
                    v0 += rr.x; v1 += rr.y;
                }
                if (p.scale) { v0 = fmaxf(fmaf(v0, s0, h0), 0.f); v1 = fmaxf(fmaf(v1, s1, h1), 0.f); }
                if (m >= YG) { v0 = 0.f; v1 = 0.f; }
                acc[nt][2 * hf] = v0; acc[nt][2 * hf + 1] = v1;
            }
        }
        if (p.m2_hi) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t lo;
                    const uint32_t hi = pack_hi_lo(acc[nt][2 * hf], acc[nt][2 * hf + 1], lo);
                    const int o = (r0 + 8 * hf) * XP + n_base + nt * 8 + cq;
                    *reinterpret_cast<uint32_t*>(yh + o) = hi;
                    *reinterpret_cast<uint32_t*>(yl + o) = lo;
                }
            __syncthreads();
            warp_product32<NT, XP>(m2h, m2l, yh, yl, m0, n_base, lane, acc);
        }
        // result -> staging tile (xh/xl are free: product 1 finished before the barrier above)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int m = r0 + 8 * hf;
                if (m >= YG) continue;
                uint32_t lo;
                const uint32_t hi = pack_hi_lo(acc[nt][2 * hf], acc[nt][2 * hf + 1], lo);
                const int o = m * XP + n_base + nt * 8 + cq;
                *reinterpret_cast<uint32_t*>(xh + o) = hi;
                *reinterpret_cast<uint32_t*>(xl + o) = lo;
            }
        __syncthreads();
        for (int i = t; i < YG * (XC / 8); i += XT) {      // coalesced write of the [60][128] bf16 tiles
            const int m = i / (XC / 8), q = i % (XC / 8);
            const size_t o = ((size_t)b * YG + m) * p.C + cb + q * 8;
            *reinterpret_cast<uint4*>(p.out_hi + o) = *reinterpret_cast<const uint4*>(xh + m * XP + q * 8);
            *reinterpret_cast<uint4*>(p.out_lo + o) = *reinterpret_cast<const uint4*>(xl + m * XP + q * 8);
        }
        __syncthreads();                                   // before this x buffer is refilled (two tiles ahead) and rstage is reused
    }
    cp_async_wait<0>();
}


// ---- warp-autonomous variant (default) ------------------------------------------------------------------------------------
// Each warp owns a (keypoint, 32-channel) tile end to end: it streams its own [60][32] hi/lo columns with cp.async (double
// buffered), holds all 64 output rows in registers (4 m-tiles x 4 n-tiles), hands the intermediate tile to its second product
// through its private shared-memory buffer and writes the result back itself.  No block-wide barrier in the main loop; the
// transform matrices are the only shared state.  Per k-step a warp issues 12 ldmatrix for 48 mma (the block-tiled kernel
// above: 16 for 24), which is what the shared-memory pipe needed.  Accumulation order per output element is the same as in
// the block-tiled kernel, so the two produce identical bits.
constexpr int WC = 32;                       // channels per warp tile
constexpr int WP = WC + 8;                   // padded row length (bf16): 80 B, conflict-free ldmatrix / stmatrix
constexpr int WBUF = 64 * WP;                // elements of one [64][WP] operand tile

__device__ __forceinline__ void stsm_x4(void* p, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};\n" :: "r"(smem_u32(p)), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

__device__ __forceinline__ void warp_stage(__nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo, const unsigned short* hi, const unsigned short* lo, int C, int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int i = lane + 32 * j;
        if (i < YG * 4) {
            const int k = i >> 2, q = i & 3;
            cp_async16(dst_hi + k * WP + q * 8, hi + (size_t)k * C + q * 8, true);
            cp_async16(dst_lo + k * WP + q * 8, lo + (size_t)k * C + q * 8, true);
        }
    }
}

// acc[mt][nt][4] = M^T (64 rows) x data (64 k x 32 channels), 3 split products, same order as warp_product32
__device__ __forceinline__ void warp_product_full(const __nv_bfloat16* mh, const __nv_bfloat16* ml, const __nv_bfloat16* xh,
                                                  const __nv_bfloat16* xl, int lane, float (&acc)[4][4][4]) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    const int arow = lane & 15, acol = (lane >> 4) * 8;
    const int brow = (lane & 7) + ((lane >> 3) & 1) * 8, bcol = (lane >> 4) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t bh[2][4], bl[2][4];
#pragma unroll
        for (int np = 0; np < 2; ++np) {
            const int off = (ks * 16 + brow) * WP + np * 16 + bcol;
            ldsm_x4_t(bh[np], xh + off);
            ldsm_x4_t(bl[np], xl + off);
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            uint32_t ah[4], al[4];
            ldsm_x4(ah, mh + (mt * 16 + arow) * MP + ks * 16 + acol);
            ldsm_x4(al, ml + (mt * 16 + arow) * MP + ks * 16 + acol);
            // the three products of one accumulator are issued four MMAs apart (same per-accumulator order: hh, lh, hl), so
            // consecutive tensor instructions are independent
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                mma16816(acc[mt][2 * np], ah, bh[np][0], bh[np][1]);
                mma16816(acc[mt][2 * np + 1], ah, bh[np][2], bh[np][3]);
            }
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                mma16816(acc[mt][2 * np], al, bh[np][0], bh[np][1]);
                mma16816(acc[mt][2 * np + 1], al, bh[np][2], bh[np][3]);
            }
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                mma16816(acc[mt][2 * np], ah, bl[np][0], bl[np][1]);
                mma16816(acc[mt][2 * np + 1], ah, bl[np][2], bl[np][3]);
            }
        }
    }
}

// accumulators -> bf16 hi/lo [64][WP] tiles (rows >= 60 are written as zeros by the caller's masking)
__device__ __forceinline__ void warp_store_tiles(__nv_bfloat16* th, __nv_bfloat16* tl, int lane, const float (&acc)[4][4][4]) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) h[nt] = pack_hi_lo(acc[mt][nt][2 * hf], acc[mt][nt][2 * hf + 1], l[nt]);
            const int o = (mt * 16 + hf * 8 + (lane & 7)) * WP + (lane >> 3) * 8;   // lane -> row (lane&7) of matrix (lane>>3) = n-tile
            stsm_x4(th + o, h[0], h[1], h[2], h[3]);
            stsm_x4(tl + o, l[0], l[1], l[2], l[3]);
        }
}

// VW warps per CTA (one CTA per SM), NBUF input buffers per warp: <8,2> prefetches the next tile while computing (248 registers),
// <16,1> trades the prefetch for twice the warps (128 registers) and lets the other warps of the scheduler hide the load.
template <int VW, int NBUF>
__global__ void __launch_bounds__(VW * 32, 1) group_transform_warp_kernel(const XmArgs p) {
    extern __shared__ __align__(16) uint8_t smraw[];
    __nv_bfloat16* m1h = (__nv_bfloat16*)smraw;          // [64][MP]
    __nv_bfloat16* m1l = m1h + 64 * MP;
    __nv_bfloat16* m2h = m1l + 64 * MP;
    __nv_bfloat16* m2l = m2h + 64 * MP;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    __nv_bfloat16* wbuf = m2l + 64 * MP + (size_t)warp * NBUF * 2 * WBUF;   // this warp's NBUF x {hi, lo} x [64][WP]
    for (int i = t; i < 64 * 8; i += VW * 32) {
        const int r = i >> 3, q = i & 7;
        *reinterpret_cast<uint4*>(m1h + r * MP + q * 8) = reinterpret_cast<const uint4*>(p.m1_hi)[i];
        *reinterpret_cast<uint4*>(m1l + r * MP + q * 8) = reinterpret_cast<const uint4*>(p.m1_lo)[i];
        if (p.m2_hi) {
            *reinterpret_cast<uint4*>(m2h + r * MP + q * 8) = reinterpret_cast<const uint4*>(p.m2_hi)[i];
            *reinterpret_cast<uint4*>(m2l + r * MP + q * 8) = reinterpret_cast<const uint4*>(p.m2_lo)[i];
        }
    }
    for (int i = lane; i < NBUF * 2 * 4 * WP; i += 32)   // rows 60..63 of the operand tiles stay zero
        wbuf[(i / (4 * WP)) * WBUF + YG * WP + i % (4 * WP)] = __float2bfloat16_rn(0.f);
    __syncthreads();
    const int cblocks = p.C / WC;
    const long long tiles = (long long)p.B * cblocks;
    const long long stride = (long long)gridDim.x * VW;
    long long tile = (long long)blockIdx.x * VW + warp;
    if (tile < tiles) {
        const int b = (int)(tile / cblocks), cb = (int)(tile - (long long)b * cblocks) * WC;
        const size_t o = (size_t)b * YG * p.C + cb;
        warp_stage(wbuf, wbuf + WBUF, p.in_hi + o, p.in_lo + o, p.C, lane);
    }
    cp_async_commit();
    const int r0 = lane >> 2, cq = 2 * (lane & 3);
    for (int it = 0; tile < tiles; tile += stride, ++it) {
        const int b = (int)(tile / cblocks), cb = (int)(tile - (long long)b * cblocks) * WC;
        __nv_bfloat16* xh = wbuf + (NBUF == 2 ? (it & 1) : 0) * 2 * WBUF;
        __nv_bfloat16* xl = xh + WBUF;
        const long long nxt = tile + stride;
        if (NBUF == 2) {
            if (nxt < tiles) {
                const int nb = (int)(nxt / cblocks), ncb = (int)(nxt - (long long)nb * cblocks) * WC;
                const size_t o = (size_t)nb * YG * p.C + ncb;
                __nv_bfloat16* nx = wbuf + ((it + 1) & 1) * 2 * WBUF;
                warp_stage(nx, nx + WBUF, p.in_hi + o, p.in_lo + o, p.C, lane);
            }
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        float acc[4][4][4];
        warp_product_full(m1h, m1l, xh, xl, lane, acc);
        // pointwise stage on the accumulators: row m = mt*16 + r0 + 8*hf, channel cb + nt*8 + cq + {0,1}
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int c = cb + nt * 8 + cq;
            float b0 = 0.f, b1 = 0.f, s0 = 1.f, s1 = 1.f, h0 = 0.f, h1 = 0.f;
            if (p.bias) { b0 = __ldg(p.bias + c); b1 = __ldg(p.bias + c + 1); }
            if (p.scale) { s0 = __ldg(p.scale + c); s1 = __ldg(p.scale + c + 1); h0 = __ldg(p.shift + c); h1 = __ldg(p.shift + c + 1); }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int m = mt * 16 + r0 + 8 * hf;
                    float v0 = acc[mt][nt][2 * hf] + b0, v1 = acc[mt][nt][2 * hf + 1] + b1;
                    if (p.resid && m < YG) {
                        const float2 rr = __ldg(reinterpret_cast<const float2*>(p.resid + ((size_t)b * YG + m) * p.C + c));
                        v0 += rr.x; v1 += rr.y;
                    }
                    if (p.scale) { v0 = fmaxf(fmaf(v0, s0, h0), 0.f); v1 = fmaxf(fmaf(v1, s1, h1), 0.f); }
                    if (m >= YG) { v0 = 0.f; v1 = 0.f; }
                    acc[mt][nt][2 * hf] = v0; acc[mt][nt][2 * hf + 1] = v1;
                }
        }
        __syncwarp();                                      // every lane is done reading xh/xl
        if (p.m2_hi) {
            warp_store_tiles(xh, xl, lane, acc);           // intermediate tile (rows 60..63 = 0) in place of the input
            __syncwarp();
            warp_product_full(m2h, m2l, xh, xl, lane, acc);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
                    if (48 + r0 + 8 * hf >= YG) { acc[3][nt][2 * hf] = 0.f; acc[3][nt][2 * hf + 1] = 0.f; }
            __syncwarp();
        }
        warp_store_tiles(xh, xl, lane, acc);               // result staging (rows 60..63 stay zero)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {                      // 8 rows x 64 B per instruction, full sectors
            const int i = lane + 32 * j;
            if (i < YG * 4) {
                const int m = i >> 2, q = i & 3;
                const size_t o = ((size_t)b * YG + m) * p.C + cb + q * 8;
                *reinterpret_cast<uint4*>(p.out_hi + o) = *reinterpret_cast<const uint4*>(xh + m * WP + q * 8);
                *reinterpret_cast<uint4*>(p.out_lo + o) = *reinterpret_cast<const uint4*>(xl + m * WP + q * 8);
            }
        }
        __syncwarp();                                      // before this buffer is refilled
        if (NBUF == 1) {
            if (nxt < tiles) {
                const int nb = (int)(nxt / cblocks), ncb = (int)(nxt - (long long)nb * cblocks) * WC;
                const size_t o = (size_t)nb * YG * p.C + ncb;
                warp_stage(xh, xl, p.in_hi + o, p.in_lo + o, p.C, lane);
            }
            cp_async_commit();
        }
    }
    cp_async_wait<0>();
}

template <int VW, int NBUF>
constexpr size_t xw_smem() { return (size_t)(4 * 64 * MP + VW * NBUF * 2 * WBUF) * sizeof(__nv_bfloat16); }

template <int VW, int NBUF>
int launch_warp_kernel(yoho_ctx* ctx, const XmArgs& p, cudaStream_t st) {
    YCHECK(cudaFuncSetAttribute(group_transform_warp_kernel<VW, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xw_smem<VW, NBUF>()));
    const long long tiles = (long long)p.B * (p.C / WC);
    const long long ctas = (tiles + VW - 1) / VW;
    group_transform_warp_kernel<VW, NBUF><<<(int)(ctas < ctx->num_sms ? ctas : ctx->num_sms), VW * 32, xw_smem<VW, NBUF>(), st>>>(p);
    return YOHO_OK;
}

template <int XC>
constexpr size_t xm_smem() { return (size_t)(4 * 64 * MP + 6 * 64 * (XC + 8)) * sizeof(__nv_bfloat16) + (size_t)YG * XC * sizeof(float); }

}  // namespace

int group_transform_tc(yoho_ctx* ctx, const void* in_hi, const void* in_lo, int B, int C, const void* m1_hi, const void* m1_lo, const void* m2_hi,
                       const void* m2_lo, const float* bias, const float* resid, const float* scale, const float* shift,
                       void* out_hi, void* out_lo, cudaStream_t st, const void* in2_hi, const void* in2_lo);
bool group_transform_tc_supported(int C, bool two, bool res);

int group_transform_mma(yoho_ctx* ctx, const void* in_hi, const void* in_lo, int B, int C, const void* m1_hi, const void* m1_lo, const void* m2_hi,
                        const void* m2_lo, const float* bias, const float* resid, const float* scale, const float* shift,
                        void* out_hi, void* out_lo, cudaStream_t st) {
    YARG(C % 128 == 0 && B > 0 && in_hi && in_lo && m1_hi && m1_lo && out_hi && out_lo);
    if ((ctx->tc_flags & 256) && group_transform_tc_supported(C, m2_hi != nullptr, resid != nullptr))      // tcgen05 variant (fourier_tc.cu)
        return group_transform_tc(ctx, in_hi, in_lo, B, C, m1_hi, m1_lo, m2_hi, m2_lo, bias, resid, scale, shift, out_hi, out_lo, st, nullptr, nullptr);
    XmArgs p{(const unsigned short*)in_hi, (const unsigned short*)in_lo, (const __nv_bfloat16*)m1_hi, (const __nv_bfloat16*)m1_lo, (const __nv_bfloat16*)m2_hi, (const __nv_bfloat16*)m2_lo,
             bias, resid, scale, shift, (unsigned short*)out_hi, (unsigned short*)out_lo, B, C};
    if ((ctx->tc_flags & (8 | 32)) == 0) {   // default: warp-autonomous 32-channel tiles, one CTA per SM
        if (int rc = (ctx->tc_flags & 64) ? launch_warp_kernel<16, 1>(ctx, p, st) : (ctx->tc_flags & 128) ? launch_warp_kernel<12, 1>(ctx, p, st)
                                                                                  : launch_warp_kernel<8, 2>(ctx, p, st)) return rc;
    } else if (ctx->tc_flags & 8) {      // 128-channel block tiles, one CTA per SM
        YCHECK(cudaFuncSetAttribute(group_transform_mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xm_smem<128>()));
        const int tiles = B * (C / 128);
        group_transform_mma_kernel<128><<<tiles < ctx->num_sms ? tiles : ctx->num_sms, XT, xm_smem<128>(), st>>>(p);
    } else {                      // flag 32: 64-channel block tiles, two CTAs per SM
        YCHECK(cudaFuncSetAttribute(group_transform_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xm_smem<64>()));
        const int tiles = B * (C / 64);
        const int cap = 2 * ctx->num_sms;
        group_transform_mma_kernel<64><<<tiles < cap ? tiles : cap, XT, xm_smem<64>(), st>>>(p);
    }
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}
