// tcgen05 group convolution: the gather-GEMM of gconv_simt.cu on 5th-generation tensor cores.
//
//   D[(b,j), o] = sum_k sum_c act[b, idx[j][k], c] * W_k[c][o]          FP32-accurate via a 2-term BF16 split
//       act = a_hi + a_lo,  W = w_hi + w_lo   (bf16 each)      D = a_hi*w_hi + a_lo*w_hi + a_hi*w_lo  (+O(2^-17))
//
// The 1e-4 descriptor bar rules out single-pass TF32/BF16 (measured 6e-4 / 4e-3 on the shipped checkpoint,
// SURVEY.md §7 hard part 1); three BF16 products per MAC measure 8e-6, at one third of the dense
// BF16 peak by construction.
//
// Kernel (persistent, one CTA per SM, 448 threads, warp-specialised):
//   warps 0-3   A producers: four lanes per tile row gather its 32-channel slice (hi and lo) for the current tap
//               straight from L2 with 16-byte cp.async into the 64B-swizzled K-major layout UMMA expects — the
//               group gather idx[j][k] is pure address arithmetic here.
//   warp  12    W producer: the weights are pre-packed on the host as ready-made swizzled 16 KB tiles, so one
//               cp.async.bulk (TMA, 1-D) per tile lands them in shared memory and signals the stage mbarrier.
//   warp  13    MMA issuer: one thread issues 6 tcgen05.mma (M128 x N256 x K16, kind::f16) per 32-wide K block (four
//               48 KB stages), accumulating all taps x Cin channels of a tile in TMEM; tcgen05.commit releases the stage.
//   warps 4-11  epilogue (two warps per TMEM lane quadrant, splitting the columns): tcgen05.ld the 128x256 FP32 accumulator (double-buffered in the 512 TMEM columns, so the
//               next tile's MMAs overlap), add bias / residual, apply the NEXT layer's folded BN + ReLU, and write
//               the activation as a bf16 hi/lo pair (the next layer's A operand) and/or FP32.
#include <vector>
#include <cuda_bf16.h>
#include "common.cuh"
#include "tcgen05.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                       // bf16 elements per K block = one 64-byte swizzle row (SWIZZLE_64B)
constexpr int A_TILE = BM * BK * 2;          // 8 KB (hi or lo)
constexpr int THREADS = 448;               // 4 A-producer warps, 8 epilogue warps, W producer, MMA issuer
constexpr int EPI_THREADS = 256;
constexpr int EPI_ROW = 64;                 // bytes per staged row (unpadded, XOR-swizzled: see coalesced_store)
constexpr int EPI_WBUF = 32 * EPI_ROW;      // per-epilogue-warp staging buffer (2 KB)
constexpr int MAX_STAGES = 8;

// Tile width BN = 256 for the wide layers (4 stages of 48 KB, two 256-column accumulators = all of TMEM) and
// BN = 32 for the 256 -> 32 output layer (8 stages of 20 KB; that layer is bound by streaming A from L2).
// K blocks are 32 wide (64-byte rows, SWIZZLE_64B): with 64-wide blocks only TWO 96 KB stages fit, and the refill of a stage
// (~1.6 us: commit -> producers -> L2) could not hide behind the 0.8 us of MMAs of the other one — the tensor pipe idled a
// third of the time (ncu: 66-68 % active, the issuing thread spinning on the `full` barrier).  Four half-size stages keep
// three refills in flight for the same shared memory.
// SPLIT: the hi*hi products accumulate in one TMEM accumulator and the two small cross products (a_lo*w_hi, a_hi*w_lo)
// in a second one, summed in FP32 (round-to-nearest) by the epilogue.  The tensor core truncates when it adds into
// the accumulator, so the error grows with the number of accumulation steps at full magnitude: SPLIT cuts that chain
// from 3*K/16 to K/16 steps.  With BN = 256 it uses all 512 TMEM columns for one tile (no epilogue overlap).
template <int BN, bool SPLIT>
struct Cfg {
    static constexpr int STAGES = BN == 256 ? 4 : 8;
    static constexpr int W_TILE = BN * BK * 2;                        // bytes (hi or lo)
    static constexpr int STAGE_BYTES = 2 * A_TILE + 2 * W_TILE;
    static constexpr int ACC_COLS = (SPLIT ? 2 : 1) * BN;             // TMEM columns of one tile's accumulator(s)
    static constexpr int NACC = (512 / ACC_COLS) >= 2 ? 2 : 1;        // accumulator buffers in flight
    static constexpr int TMEM_COLS = NACC * ACC_COLS < 32 ? 32 : NACC * ACC_COLS;   // power of two >= 32
    // kind::f16 instruction descriptor: D=F32, A=B=BF16, both K-major, N=BN, M=128 (cute::UMMA::InstrDescriptor)
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    static constexpr size_t SMEM_BYTES = 1024 /* alignment slack */ + (size_t)STAGES * STAGE_BYTES + 3 * 512 * sizeof(float) +
                                         64 * 13 * sizeof(int) + 256 + 8 * EPI_WBUF;
};

// One GEMM of a launch.  A launch is a list of up to MAX_GROUPS GEMMs sharing the activation tensor and the output
// tensor (the per-irrep GEMMs of a group-Fourier layer run as ONE persistent launch, so there is one tail, not five).
constexpr int MAX_GROUPS = 5;
struct TcGroup {
    const uint8_t* w_hi;         // [n_tiles][nkb] swizzled tiles
    const uint8_t* w_lo;
    const int* idx;              // [Jout][taps] input-row table
    const int* omap;             // [Jout][Cout/ogroup] output-row table (nullptr: plain [rows][Cout] output)
    int Jout, taps, Cout, nkb, m_total, n_tiles, tile_begin, idx_off, omap_off;
    int n_valid;                 // columns >= n_valid are not written
    int n_mma;                   // UMMA N of this GEMM (multiple of 16, <= BN): rows n_mma..BN-1 of its W tiles are zero padding and are
                                 // neither fetched nor multiplied (PartI layer 4 in the group-Fourier domain: d*32 of 256 columns)
};

struct TcArgs {
    const __nv_bfloat16* a_hi;   // [B][Jin][Cin]
    const __nv_bfloat16* a_lo;
    const float* bias;
    int Jin, Cin;
    int ngroups, total_tiles;
    TcGroup grp[MAX_GROUPS];
    const float* resid;
    int Jres, resid_off, resid_per_j;
    float* out_raw;
    float* out_act;              // FP32 activation (optional)
    __nv_bfloat16* out_hi;       // bf16 split activation (optional)
    __nv_bfloat16* out_lo;
    const float* scale;
    const float* shift;
    int flags;                   // bit0: non-blocking producer completion (bit1, once a lane-map choice, is ignored)
    int cluster;                 // 1: launched as 2-CTA clusters that share every weight tile (each CTA fetches half of it and multicasts
                                 // it to both); `tile` then counts PAIR tiles: the two CTAs take row tiles 2 m and 2 m + 1 of one column tile
    // remapped output rows (group-Fourier layers): columns are groups of `ogroup`; group i of GEMM row (b,j) is written to
    // row b*out_J + omap[j*n_groups + i] of an [.., ogroup]-wide output.
    int ogroup, out_J;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// 1-D bulk copy global -> the same shared-memory offset of every CTA in `mask`, signalling the mbarrier at the same offset in each
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, unsigned long long* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tc_commit_mc(unsigned long long* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

__device__ __forceinline__ int find_group(const TcArgs& p, int tile) {
    int g = 0;
    while (g + 1 < p.ngroups && tile >= p.grp[g + 1].tile_begin) ++g;
    return g;
}

// K-major operand tile, 64-byte rows, SWIZZLE_64B (16-byte chunk c of row r stored at c ^ ((r >> 1) & 3)), 8-row groups 512 B
// apart (cute::UMMA::SmemDescriptor, layout type 4).
__device__ __forceinline__ uint64_t umma_desc_sw64(const void* smem_tile) {
    const uint64_t addr = (uint64_t)(smem_u32(smem_tile) >> 4) & 0x3FFFull;
    return addr | (1ull << 16) /* LBO (unused for swizzled K-major) */ | (32ull << 32) /* SBO = 512 B */ |
           (1ull << 46) /* descriptor version: Blackwell */ | (4ull << 61) /* SWIZZLE_64B */;
}

struct __align__(8) Barriers {
    unsigned long long full[MAX_STAGES];
    unsigned long long empty[MAX_STAGES];
    unsigned long long tmem_full[2];
    unsigned long long tmem_empty[2];
    uint32_t tmem_base;
};

// Epilogue store of a [32 rows x 64 bytes] block held one row per lane (four uint4 registers per lane): staged through a
// per-warp 2 KB shared-memory buffer so that each warp-wide store instruction writes 64 contiguous bytes of 8 rows instead of
// 16 bytes of 32 different rows (the L1TEX tag stage serialises on cache lines touched per instruction).  The buffer is
// unpadded and XOR-swizzled — 16-byte piece i of row r lives at r*64 + ((i ^ (r >> 1)) & 3)*16 — which makes both the row-per-
// lane writes and the 4-lanes-per-row reads conflict-free (the former 80-byte pitch cost two wavefronts per read quarter).
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
// `rowoff[it]` = BYTE offset (from `base`) of row it*8 + lane/4 at this block's first column — the caller shuffles the
// per-lane row offsets once per 32-column chunk and reuses them for every tensor it writes (FP32, hi, lo).
template <int NB>
__device__ __forceinline__ void coalesced_store(uint32_t wb, const uint4* regs, uint8_t* base, const uint32_t* rowoff, int lane, uint32_t okmask) {
    static_assert(NB == 64, "the swizzle below is written for four 16-byte pieces per row");
    constexpr int Q = NB / 16;                       // 16-byte pieces per row
    __syncwarp();
#pragma unroll
    for (int i = 0; i < Q; ++i) sts128(wb + lane * EPI_ROW + (((uint32_t)i ^ ((uint32_t)lane >> 1)) & 3u) * 16, regs[i]);
    __syncwarp();
    const int q = lane & 3;
#pragma unroll
    for (int it = 0; it < Q; ++it) {
        const int row = it * 8 + (lane >> 2);
        const uint4 v = lds128(wb + row * EPI_ROW + (((uint32_t)q ^ ((uint32_t)row >> 1)) & 3u) * 16);
        if ((okmask >> row) & 1u) *reinterpret_cast<uint4*>(base + rowoff[it] + q * 16) = v;
    }
}

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(THREADS, 1) gconv_tc_kernel(const TcArgs p) {
    using C = Cfg<BN, SPLIT>;
    constexpr int STAGES = C::STAGES;
    constexpr int W_TILE = C::W_TILE;
    constexpr int STAGE_BYTES = C::STAGE_BYTES;
    constexpr int TMEM_COLS = C::TMEM_COLS;
    constexpr int ACC_COLS = C::ACC_COLS;
    constexpr int NACC = C::NACC;
    constexpr uint32_t IDESC = C::IDESC;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* stage_base = smem;
    float* ep_bias = (float*)(smem + STAGES * STAGE_BYTES);
    float* ep_scale = ep_bias + 512;
    float* ep_shift = ep_scale + 512;
    int* idx_s = (int*)(ep_shift + 512);
    Barriers* bars = (Barriers*)(idx_s + 64 * 13);
    uint8_t* epi_stage = (uint8_t*)bars + 256;          // 8 x EPI_WBUF, 16-byte aligned

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    for (int g = 0; g < p.ngroups; ++g) {
        const TcGroup& G = p.grp[g];
        for (int i = threadIdx.x; i < G.Jout * G.taps; i += THREADS) idx_s[G.idx_off + i] = G.idx[i];
        if (G.omap) for (int i = threadIdx.x; i < G.Jout * (G.Cout / p.ogroup); i += THREADS) idx_s[G.omap_off + i] = G.omap[i];
    }
    // bias / next-BN tables of all Cout (<= 512) channels; the group-Fourier layers (Cout up to 2560) carry no bias or
    // activation here (both are applied in the group domain by the transform kernel)
    const bool has_ep = p.grp[0].omap == nullptr;
    for (int i = threadIdx.x; has_ep && i < p.grp[0].Cout; i += THREADS) {
        ep_bias[i] = p.bias[i];
        ep_scale[i] = p.scale ? p.scale[i] : 1.f;
        ep_shift[i] = p.shift ? p.shift[i] : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->full[s], 128 + 1);   // 128 A-producer threads + the W producer's expect_tx arrive
            mbar_init(&bars->empty[s], p.cluster ? 2 : 1);   // tcgen05.commit (of both CTAs of a cluster: the peer writes into this stage too)
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);    // tcgen05.commit
            mbar_init(&bars->tmem_empty[a], EPI_THREADS); // epilogue threads
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 13) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&bars->tmem_base)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.cluster) cluster_sync_all();           // the peer's barriers are initialised before anything is multicast to them
    const uint32_t tmem_base = bars->tmem_base;
    const int total_tiles = p.total_tiles;
    // cluster mode: CTA pair c = blockIdx.x / 2 walks the pair tiles c, c + gridDim.x / 2, ...; rank r takes row tile 2 m + r
    const int cs = p.cluster ? 2 : 1;
    const int crank = p.cluster ? (int)cluster_ctarank() : 0;
    const int tile0 = blockIdx.x / cs, tstride = gridDim.x / cs;

    if (warp < 4) {
        // ================= A producers =================
        // Lane map: 4 consecutive lanes copy the 4 16-byte chunks of ONE 64-byte row segment, so a warp-wide cp.async touches 8
        // rows; warp w owns tile rows [32w, 32w+32), lane (rsub = lane/4, c = lane%4) copies chunk c of rows 32w + 8i + rsub.
        // Completion protocol (p.flags bit 0): noinc = cp.async.mbarrier.arrive.noinc — the barrier arrival fires when this
        // thread's copies land and the thread moves on (the MMA thread fences the proxy); otherwise wait_group 0, fence, arrive.
        const bool noinc = (p.flags & 1) != 0;
        const uint32_t c = (uint32_t)(lane & 3);          // 16-byte chunk of the 64-byte row this lane copies
        const int rsub = lane >> 2;                       // row within a group of 8
        uint32_t stage = 0, phase = 0;
        for (int tile = tile0; tile < total_tiles; tile += tstride) {
            const TcGroup& G = p.grp[find_group(p, tile)];
            const int m_tile = ((tile - G.tile_begin) / G.n_tiles) * cs + crank;
            const int* idx_g = idx_s + G.idx_off;
            int rowbase[4], rowj[4];
            uint32_t okmask = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = warp * 32 + 8 * i + rsub;
                const int row = m_tile * BM + r;
                const bool ok = row < G.m_total;
                const int rr = ok ? row : 0;
                const int b = rr / G.Jout;
                rowj[i] = (rr - b * G.Jout) * G.taps;
                rowbase[i] = b * p.Jin;
                okmask |= (ok ? 1u : 0u) << i;
            }
            for (int kb = 0; kb < G.nkb; ++kb) {
                uint8_t* st_hi = stage_base + stage * STAGE_BYTES;
                // K axis = tap-major, channel-minor; a 32-wide block never straddles a tap (Cin % 32 == 0)
                const int kk0 = kb * BK + (int)c * 8;
                const int k = kk0 / p.Cin;
                const int coff = kk0 - k * p.Cin;
                mbar_wait(&bars->empty[stage], phase ^ 1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = warp * 32 + 8 * i + rsub;
                    const size_t off = ((size_t)(rowbase[i] + idx_g[rowj[i] + k])) * p.Cin + coff;
                    uint8_t* dst = st_hi + r * 64 + ((c ^ (uint32_t)((r >> 1) & 3)) << 4);
                    const bool ok = (okmask >> i) & 1u;
                    cp_async16(dst, p.a_hi + off, ok);
                    cp_async16(dst + A_TILE, p.a_lo + off, ok);
                }
                if (noinc) {
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(&bars->full[stage])) : "memory");
                } else {
                    cp_async_commit();
                    cp_async_wait<0>();
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
                    mbar_arrive(&bars->full[stage]);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 12) {
        // ================= W producer (one thread) =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tstride) {
                const TcGroup& G = p.grp[find_group(p, tile)];
                const int n_tile = (tile - G.tile_begin) % G.n_tiles;
                const uint8_t* wh = G.w_hi + (size_t)n_tile * G.nkb * W_TILE;
                const uint8_t* wl = G.w_lo + (size_t)n_tile * G.nkb * W_TILE;
                const uint32_t wbytes = (uint32_t)G.n_mma * (BK * 2);       // rows [0, n_mma) of the tile image
                const uint32_t hbytes = wbytes / 2, hoff = (uint32_t)crank * hbytes;   // cluster mode: this CTA's half of the rows
                for (int kb = 0; kb < G.nkb; ++kb) {
                    uint8_t* dst = stage_base + stage * STAGE_BYTES + 2 * A_TILE;
                    mbar_wait(&bars->empty[stage], phase ^ 1);             // cluster mode: BOTH CTAs have released this stage
                    mbar_expect_tx(&bars->full[stage], 2 * wbytes);
                    if (p.cluster) {
                        bulk_g2s_mc(dst + hoff, wh + (size_t)kb * W_TILE + hoff, hbytes, &bars->full[stage], (uint16_t)3);
                        bulk_g2s_mc(dst + W_TILE + hoff, wl + (size_t)kb * W_TILE + hoff, hbytes, &bars->full[stage], (uint16_t)3);
                    } else {
                        bulk_g2s(dst, wh + (size_t)kb * W_TILE, wbytes, &bars->full[stage]);
                        bulk_g2s(dst + W_TILE, wl + (size_t)kb * W_TILE, wbytes, &bars->full[stage]);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 13) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tstride) {
                mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                const uint32_t d_lo = SPLIT ? d_tmem + BN : d_tmem;
                const TcGroup& G = p.grp[find_group(p, tile)];
                const int nkb = G.nkb;
                // instruction descriptor with this GEMM's N (bits 17..22 hold N >> 3)
                const uint32_t idesc = (IDESC & ~(0x3Fu << 17)) | ((uint32_t)(G.n_mma >> 3) << 17);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&bars->full[stage], phase);
                    if (p.flags & 1) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // cp.async data -> UMMA (async proxy)
                    tc_fence_after();
                    const uint8_t* st = stage_base + stage * STAGE_BYTES;
                    const uint64_t a_hi = umma_desc_sw64(st), a_lo = umma_desc_sw64(st + A_TILE);
                    const uint64_t w_hi = umma_desc_sw64(st + 2 * A_TILE), w_lo = umma_desc_sw64(st + 2 * A_TILE + W_TILE);
#pragma unroll
                    for (uint32_t ks = 0; ks < BK / 16; ++ks) {
                        const uint64_t adv = (uint64_t)(ks * 2);   // 32 bytes per 16 bf16, in 16-byte units
                        const uint32_t first = (kb | (int)ks) ? 1u : 0u;
                        tc_mma(d_tmem, a_hi + adv, w_hi + adv, idesc, first);
                        tc_mma(d_lo, a_lo + adv, w_hi + adv, idesc, SPLIT ? first : 1u);
                        tc_mma(d_lo, a_hi + adv, w_lo + adv, idesc, 1u);
                    }
                    if (p.cluster) tc_commit_mc(&bars->empty[stage], (uint16_t)3);   // both CTAs learn that this CTA is done with the stage
                    else tc_commit(&bars->empty[stage]);       // stage reusable once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&bars->tmem_full[acc]);              // accumulator complete
                if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue warps 4..11 =================
        // warp % 4 selects the TMEM lane quadrant the warp may access; the two warps of a quadrant split the columns
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        uint32_t acc = 0, acc_phase = 0;
        for (int tile = tile0; tile < total_tiles; tile += tstride) {
            const TcGroup& G = p.grp[find_group(p, tile)];
            const int m_pair = (tile - G.tile_begin) / G.n_tiles;
            const int n_tile = (tile - G.tile_begin) - m_pair * G.n_tiles;
            const int m_tile = m_pair * cs + crank;
            const int* omap_s = idx_s + G.omap_off;
            const int row = m_tile * BM + q * 32 + lane;
            const bool ok = row < G.m_total;
            const int n0 = n_tile * BN;
            const float* rres = nullptr;
            if (p.resid && ok) {
                const int b = row / G.Jout;
                const int j = row - b * G.Jout;
                rres = p.resid + ((size_t)b * p.Jres + p.resid_off + (p.resid_per_j ? j : 0)) * G.Cout + n0;
            }
            mbar_wait(&bars->tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * ACC_COLS + ((uint32_t)(q * 32) << 16);
            const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
            const uint32_t wb = smem_u32(epi_stage + (warp - 4) * EPI_WBUF);
            // this lane's row as (keypoint, output row) — one division per tile; column groups are powers of two wide
            const int row_b = (ok ? row : 0) / G.Jout, row_j = (ok ? row : 0) - row_b * G.Jout;
            const int og_shift = 31 - __clz(p.ogroup), og_count = G.Cout >> og_shift;
            constexpr int NCH = BN / 32;                           // 32-column chunks of the tile
            constexpr int CH0 = NCH >= 2 ? NCH / 2 : 0;           // chunks [0,CH0) -> half 0, [CH0,NCH) -> half 1
            const int cc_lo = half == 0 ? 0 : CH0, cc_hi = half == 0 ? (NCH >= 2 ? CH0 : 0) : NCH;
#pragma unroll 1
            for (int cc = cc_lo; cc < cc_hi; ++cc) {
                uint32_t v[32];
                tmem_ld32(t_addr + cc * 32, v);     // .sync.aligned: executed by the whole warp, rows past the end included
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
                if constexpr (SPLIT) {
                    tmem_ld32(t_addr + BN + cc * 32, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] += __uint_as_float(v[i]);
                }
                if (n0 + cc * 32 < G.n_valid) {       // warp-uniform
                    if (has_ep) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) f[i] += ep_bias[n0 + cc * 32 + i];
                    }
                    if (rres) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            const float4 rv = *reinterpret_cast<const float4*>(rres + cc * 32 + i);
                            f[i] += rv.x; f[i + 1] += rv.y; f[i + 2] += rv.z; f[i + 3] += rv.w;
                        }
                    }
                    // element offset of this lane's row at this chunk's first column (32-bit: every output tensor of a pass
                    // has fewer than 2^32 elements), then the offsets of the four row groups this lane STORES (rows it*8 + lane/4)
                    uint32_t blk;
                    if (G.omap) {
                        const int n = n0 + cc * 32, grp = n >> og_shift;
                        blk = ((uint32_t)(row_b * p.out_J + omap_s[row_j * og_count + grp]) << og_shift) + (uint32_t)(n - (grp << og_shift));
                    } else {
                        blk = (uint32_t)(ok ? row : 0) * (uint32_t)G.Cout + (uint32_t)(n0 + cc * 32);
                    }
                    uint32_t ro[4];
#pragma unroll
                    for (int it = 0; it < 4; ++it) ro[it] = __shfl_sync(0xffffffffu, blk, it * 8 + (lane >> 2));
                    if (p.out_raw) {
                        uint32_t ro4[4];
#pragma unroll
                        for (int it = 0; it < 4; ++it) ro4[it] = ro[it] * 4u;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint4 r4[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                r4[i] = make_uint4(__float_as_uint(f[16 * h + 4 * i]), __float_as_uint(f[16 * h + 4 * i + 1]),
                                                   __float_as_uint(f[16 * h + 4 * i + 2]), __float_as_uint(f[16 * h + 4 * i + 3]));
                            coalesced_store<64>(wb, r4, (uint8_t*)(p.out_raw + 16 * h), ro4, lane, okmask);
                        }
                    }
                    if (p.out_act || p.out_hi) {
                        if (has_ep) {      // next layer's BN + ReLU; group-Fourier layers emit the raw value (hi/lo split only)
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                f[i] = fmaxf(fmaf(f[i], ep_scale[n0 + cc * 32 + i], ep_shift[n0 + cc * 32 + i]), 0.f);
                        }
                        if (p.out_act) {
                            uint32_t ro4[4];
#pragma unroll
                            for (int it = 0; it < 4; ++it) ro4[it] = ro[it] * 4u;
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                uint4 r4[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    r4[i] = make_uint4(__float_as_uint(f[16 * h + 4 * i]), __float_as_uint(f[16 * h + 4 * i + 1]),
                                                       __float_as_uint(f[16 * h + 4 * i + 2]), __float_as_uint(f[16 * h + 4 * i + 3]));
                                coalesced_store<64>(wb, r4, (uint8_t*)(p.out_act + 16 * h), ro4, lane, okmask);
                            }
                        }
                        if (p.out_hi) {
                            uint32_t hi[16], lo[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                                hi[i] = *reinterpret_cast<const uint32_t*>(&h2);
                                const __nv_bfloat162 l2 = __floats2bfloat162_rn(f[2 * i] - __uint_as_float(hi[i] << 16),
                                                                                f[2 * i + 1] - __uint_as_float(hi[i] & 0xffff0000u));
                                lo[i] = *reinterpret_cast<const uint32_t*>(&l2);
                            }
                            uint32_t ro2[4];
#pragma unroll
                            for (int it = 0; it < 4; ++it) ro2[it] = ro[it] * 2u;
                            uint4 r4[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) r4[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
                            coalesced_store<64>(wb, r4, (uint8_t*)p.out_hi, ro2, lane, okmask);
#pragma unroll
                            for (int i = 0; i < 4; ++i) r4[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                            coalesced_store<64>(wb, r4, (uint8_t*)p.out_lo, ro2, lane, okmask);
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&bars->tmem_empty[acc]);
            if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster) cluster_sync_all();           // no CTA leaves while its peer may still multicast into it or arrive on its barriers
    if (warp == 13) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float f[4] = {v.x, v.y, v.z, v.w};
    unsigned short h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat16 hh = __float2bfloat16_rn(f[k]);
        h[k] = __bfloat16_as_ushort(hh);
        l[k] = __bfloat16_as_ushort(__float2bfloat16_rn(f[k] - __bfloat162float(hh)));
    }
    reinterpret_cast<uint2*>(hi)[i] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    reinterpret_cast<uint2*>(lo)[i] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}

inline unsigned short f2bf(float f) {   // round-to-nearest-even, host side (weights are finite)
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (unsigned short)(u >> 16);
}
inline float bf2f(unsigned short h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

}  // namespace

// Host-side packing: W_k[c][o] fp32 -> bf16 hi/lo, laid out as [n_tile][kb] ready-to-copy tiles (BN rows x 64 B) in the
// UMMA K-major SWIZZLE_64B image (row r at r*64 bytes, 16-byte chunk j stored at position j ^ ((r >> 1) & 3)).
// K axis = tap-major, channel-minor (Cin is a multiple of 32: no padding).
static int tc_tile_n(const GLayer& L) {
    if (L.taps != YT && !L.tc_dense) return 0;
    if (!(L.cin == 32 || L.cin % BK == 0)) return 0;
    if (L.cout % 256 == 0) return 256;
    if (L.cout == 32) return 32;
    return 0;
}

int gconv_tc_pack(yoho_ctx*, GLayer& L, const std::vector<float>& w) {
    const int bn = tc_tile_n(L);
    if (!bn) return YOHO_OK;   // not a tensor-core layer
    const int ktot = L.taps * L.cin;
    const int nkb = (ktot + BK - 1) / BK, n_tiles = L.cout / bn;
    const size_t tile_elems = (size_t)bn * BK;
    const size_t elems = (size_t)n_tiles * nkb * tile_elems;
    std::vector<unsigned short> hi(elems, 0), lo(elems, 0);
    for (int nt = 0; nt < n_tiles; ++nt)
        for (int kb = 0; kb < nkb; ++kb) {
            const size_t tile = ((size_t)nt * nkb + kb) * tile_elems;
            for (int r = 0; r < bn; ++r)
                for (int i = 0; i < BK; ++i) {
                    const int kk = kb * BK + i;
                    if (kk >= ktot) continue;
                    const int k = kk / L.cin, c = kk - k * L.cin;
                    const float v = w[((size_t)k * L.cin + c) * L.cout + nt * bn + r];
                    const unsigned short h = f2bf(v);
                    const size_t pos = tile + (size_t)r * BK + (size_t)(((i >> 3) ^ ((r >> 1) & 3)) << 3) + (i & 7);
                    hi[pos] = h;
                    lo[pos] = f2bf(v - bf2f(h));
                }
        }
    YCHECK(cudaMalloc(&L.w_hi, elems * 2));
    YCHECK(cudaMalloc(&L.w_lo, elems * 2));
    YCHECK(cudaMemcpy(L.w_hi, hi.data(), elems * 2, cudaMemcpyHostToDevice));
    YCHECK(cudaMemcpy(L.w_lo, lo.data(), elems * 2, cudaMemcpyHostToDevice));
    return YOHO_OK;
}

bool gconv_tc_eligible(const GLayer& L, const GConvArgs& a) {
    return L.w_hi && L.w_lo && tc_tile_n(L) && a.act_hi && a.act_lo && a.B * a.Jout >= BM && a.Jout * L.taps <= 64 * 13;
}

int gconv_split_bf16(yoho_ctx* ctx, const float* x, void* hi, void* lo, size_t n, cudaStream_t st) {
    YARG(n % 4 == 0);
    const size_t n4 = n / 4;
    split_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n4);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

template <int BN, bool SPLIT>
static int tc_launch(yoho_ctx* ctx, TcArgs& p, cudaStream_t st) {
    // per-device attribute; cheap enough to set on every launch (one process may drive several devices)
    YCHECK(cudaFuncSetAttribute(gconv_tc_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<BN, SPLIT>::SMEM_BYTES));
    if (p.cluster) {
        // 2-CTA clusters: total_tiles counts pair tiles
        int pairs = p.total_tiles < ctx->num_sms / 2 ? p.total_tiles : ctx->num_sms / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = Cfg<BN, SPLIT>::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        YCHECK(cudaLaunchKernelEx(&cfg, gconv_tc_kernel<BN, SPLIT>, p));
        ctx->launches++;
        return YOHO_OK;
    }
    const int grid = p.total_tiles < ctx->num_sms ? p.total_tiles : ctx->num_sms;
    gconv_tc_kernel<BN, SPLIT><<<grid, THREADS, Cfg<BN, SPLIT>::SMEM_BYTES, st>>>(p);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

static void tc_fill_common(TcArgs& p, const GLayer& L, const GConvArgs& a, yoho_ctx* ctx) {
    p.a_hi = (const __nv_bfloat16*)a.act_hi; p.a_lo = (const __nv_bfloat16*)a.act_lo;
    p.bias = L.bias;
    p.Jin = a.Jin; p.Cin = L.cin;
    p.resid = a.resid; p.Jres = a.Jres; p.resid_off = a.resid_off; p.resid_per_j = a.resid_per_j;
    p.out_raw = a.out_raw; p.out_act = a.out_act;
    p.out_hi = (__nv_bfloat16*)a.out_hi; p.out_lo = (__nv_bfloat16*)a.out_lo;
    p.scale = a.scale; p.shift = a.shift;
    p.flags = ctx->tc_flags;
    p.cluster = 0;
    p.ogroup = a.omap ? a.ogroup : L.cout; p.out_J = a.out_J;
}

static void tc_fill_group(TcGroup& G, const GLayer& L, const GConvArgs& a, int bn, int tile_begin, int idx_off, int omap_off) {
    G.w_hi = (const uint8_t*)L.w_hi; G.w_lo = (const uint8_t*)L.w_lo;
    G.idx = a.idx; G.omap = a.omap;
    G.Jout = a.Jout; G.taps = L.taps; G.Cout = L.cout;
    G.nkb = (L.taps * L.cin + BK - 1) / BK;
    G.m_total = a.B * a.Jout;
    G.n_tiles = L.cout / bn;
    G.tile_begin = tile_begin; G.idx_off = idx_off; G.omap_off = omap_off;
    G.n_valid = a.n_valid > 0 ? a.n_valid : L.cout;
    G.n_mma = bn;
    if (bn == 256 && G.n_tiles == 1 && G.n_valid < bn) {            // single-tile GEMM with zero-padded columns
        G.n_mma = ((G.n_valid + 15) / 16) * 16;
        if (G.n_mma < 32) G.n_mma = 32;
    }
}

int gconv_tc_forward(yoho_ctx* ctx, const GLayer& L, const GConvArgs& a, cudaStream_t st) {
    YARG(gconv_tc_eligible(L, a));
    YARG(a.omap ? (!a.out_act && !a.resid && L.cout % a.ogroup == 0 && a.ogroup % 32 == 0 && (a.ogroup & (a.ogroup - 1)) == 0) : L.cout <= 512);
    TcArgs p;
    tc_fill_common(p, L, a, ctx);
    const int bn = tc_tile_n(L);
    p.ngroups = 1;
    tc_fill_group(p.grp[0], L, a, bn, 0, 0, 800);       // <= 25 omap entries behind the (<= 780-entry) index table
    p.total_tiles = ((p.grp[0].m_total + BM - 1) / BM) * p.grp[0].n_tiles;
    // split accumulators only where the accumulation chain is long; short-K layers (PartI layers 1 and 4, the group-Fourier
    // GEMMs) keep two accumulator buffers in flight so that their (relatively heavy) epilogue overlaps the next tile's MMAs
    const bool split = ctx->gconv_impl >= 2 && p.grp[0].nkb * BK >= ctx->split_min_k && !a.omap;
    if (bn == 256) return split ? tc_launch<256, true>(ctx, p, st) : tc_launch<256, false>(ctx, p, st);
    return split ? tc_launch<32, true>(ctx, p, st) : tc_launch<32, false>(ctx, p, st);
}

// The per-irrep GEMMs of one group-Fourier layer as ONE persistent launch: n GEMMs over the same activation tensor
// (coefficient rows selected by each irrep's idx table) writing disjoint coefficient rows of the same output.
// Groups are visited in the given order; pass the largest first so the tail of the launch is made of short tiles.
int gconv_tc_forward_grouped(yoho_ctx* ctx, const GLayer* const* Ls, const GConvArgs* as, int n, cudaStream_t st) {
    YARG(n >= 1 && n <= MAX_GROUPS);
    TcArgs p;
    tc_fill_common(p, *Ls[0], as[0], ctx);
    p.ngroups = n;
    // 2-CTA clusters sharing every weight tile (tuning flag 32768): the weight tile is 2/3 of the operand bytes of a K block
    p.cluster = (ctx->tc_flags & 32768) ? 1 : 0;
    int tiles = 0;
    for (int g = 0; g < n; ++g) {
        const GLayer& L = *Ls[g];
        const GConvArgs& a = as[g];
        YARG(gconv_tc_eligible(L, a) && tc_tile_n(L) == 256 && a.omap && !a.out_act && !a.resid);
        YARG(L.cin == Ls[0]->cin && a.ogroup == as[0].ogroup && a.act_hi == as[0].act_hi && a.out_hi == as[0].out_hi &&
             a.out_raw == as[0].out_raw && a.B == as[0].B && a.Jin == as[0].Jin && L.cout % a.ogroup == 0 && (a.ogroup & (a.ogroup - 1)) == 0);
        tc_fill_group(p.grp[g], L, a, 256, tiles, g * 32, 160 + g * 40);   // <= 25 index entries, <= 5 x 8 output-row entries
        const int m_tiles = (p.grp[g].m_total + BM - 1) / BM;
        tiles += (p.cluster ? (m_tiles + 1) / 2 : m_tiles) * p.grp[g].n_tiles;
        if (p.cluster && (p.grp[g].n_mma % 32)) p.cluster = -1;           // halves must be whole 16-row groups of the tile image
    }
    if (p.cluster < 0) { yoho_set_error("cluster mode needs UMMA N multiples of 32"); return YOHO_ERR_ARG; }
    p.total_tiles = tiles;
    return tc_launch<256, false>(ctx, p, st);
}
