// FP32 SIMT gather-GEMM: the group convolution of YOHO as an implicit GEMM.
//
//   out[(b,j), o] = bias[o] + sum_k sum_c act[b, idx[j][k], c] * W_k[c][o]      (+ residual, + next BN/ReLU)
//
// Replaces, fused: the 13x gather `data[:,:,Nei_in_SO3]` + reshape (utils/network.py:46-52,80-84), the
// Conv2d(C,O,(1,13)) (utils/network.py:18,30,35,76), the residual add (:65) and the NEXT layer's eval-mode
// BatchNorm+ReLU (utils/network.py:16-17,28-29,33-34), which commutes with the gather.
// With taps=1 and idx={0} it is the 1x1 Conv2d of the PartII head (utils/network.py:232-240).
//
// Layout: activations are [B][J][C] (channel innermost) so a GEMM row for tap k is one contiguous C-vector
// at row idx[j][k]; the gather costs nothing but address arithmetic in the cp.async producer.
// Tile 128 rows x BN cols x 16 k, 256 threads, 8 x (BN/16) outputs per thread, 3-stage cp.async pipeline.
// Accumulation order is fixed (taps outer, channels ascending, FP32 FMA), so results are deterministic.
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int THREADS = 256;
constexpr int STAGES = 3;
constexpr int AS_STRIDE = 20;  // floats per A row in smem: 16 + 4 pad (16B aligned, conflict-free LDS.128)

template <int BN>
struct SmemLayout {
    static constexpr int A_FLOATS = BM * AS_STRIDE;
    static constexpr int B_FLOATS = BK * BN;
    static constexpr int STAGE_FLOATS = A_FLOATS + B_FLOATS;
    static constexpr int IDX_INTS = 64 * 13;  // Jout*taps <= 60*13 = 780
    static constexpr size_t BYTES = (size_t)STAGES * STAGE_FLOATS * sizeof(float) + IDX_INTS * sizeof(int);
};

struct KArgs {
    const float* act;
    const float* w;
    const float* bias;
    const int* idx;
    int B, Jin, Jout, Cin, Cout, taps;
    const float* resid;
    int Jres, resid_off, resid_per_j;
    float* out_raw;
    float* out_act;
    const float* scale;
    const float* shift;
    unsigned short* out_hi;   // optional bf16 hi/lo split of out_act (feeds the tensor-core layers)
    unsigned short* out_lo;
    int n_tiles;   // Cout / BN
    int m_total;   // B * Jout
};

template <int BN>
__global__ void __launch_bounds__(THREADS, 2) gconv_f32_kernel(const KArgs p) {
    constexpr int TN = BN / 16;
    using L = SmemLayout<BN>;
    extern __shared__ __align__(16) float smem[];
    int* idx_s = reinterpret_cast<int*>(smem + STAGES * L::STAGE_FLOATS);

    const int t = threadIdx.x;
    const int tx = t & 15;
    const int ty = t >> 4;
    const int n_tile = blockIdx.x % p.n_tiles;
    const int m_tile = blockIdx.x / p.n_tiles;
    const int m0 = m_tile * BM;
    const int n0 = n_tile * BN;

    for (int i = t; i < p.Jout * p.taps; i += THREADS) idx_s[i] = p.idx[i];
    __syncthreads();

    // --- loader assignment: A chunks t and t+256 (row = chunk/4, quad = chunk%4) --------------------
    const int quad = t & 3;
    int rowbase[2], rowj[2];
    bool rowok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int r = m0 + (t >> 2) + 64 * h;
        rowok[h] = r < p.m_total;
        int rr = rowok[h] ? r : 0;
        int b = rr / p.Jout;
        rowj[h] = (rr - b * p.Jout) * p.taps;
        rowbase[h] = b * p.Jin;
    }
    const int cblocks = p.Cin / BK;
    const int nkb = p.taps * cblocks;

    auto load_stage = [&](int kb, int stage) {
        const int k = kb / cblocks;
        const int c0 = (kb - k * cblocks) * BK;
        float* As = smem + stage * L::STAGE_FLOATS;
        float* Bs = As + L::A_FLOATS;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int srow = rowbase[h] + idx_s[rowj[h] + k];
            const float* src = p.act + (size_t)srow * p.Cin + c0 + quad * 4;
            cp_async16(As + ((t >> 2) + 64 * h) * AS_STRIDE + quad * 4, src, rowok[h]);
        }
        constexpr int BCH = BK * BN / 4;  // float4 chunks of the B tile
        const float* wk = p.w + ((size_t)k * p.Cin + c0) * p.Cout + n0;
#pragma unroll
        for (int c = t; c < BCH; c += THREADS) {
            const int kk = c / (BN / 4);
            const int n4 = c - kk * (BN / 4);
            cp_async16(Bs + kk * BN + n4 * 4, wk + (size_t)kk * p.Cout + n4 * 4, true);
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkb) load_stage(s, s);
        cp_async_commit();
    }

    for (int kb = 0; kb < nkb; ++kb) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = kb + STAGES - 1;
            if (nx < nkb) load_stage(nx, nx % STAGES);
            cp_async_commit();
        }
        const float* As = smem + (kb % STAGES) * L::STAGE_FLOATS;
        const float* Bs = As + L::A_FLOATS;
#pragma unroll
        for (int kk4 = 0; kk4 < BK / 4; ++kk4) {
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * AS_STRIDE + kk4 * 4);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* brow = Bs + (kk4 * 4 + q) * BN;
                float b[TN];
                if constexpr (TN == 8) {
                    float4 b0 = *reinterpret_cast<const float4*>(brow + tx * 4);
                    float4 b1 = *reinterpret_cast<const float4*>(brow + 64 + tx * 4);
                    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                    b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
                } else if constexpr (TN == 4) {
                    float4 b0 = *reinterpret_cast<const float4*>(brow + tx * 4);
                    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                } else {
                    float2 b0 = *reinterpret_cast<const float2*>(brow + tx * 2);
                    b[0] = b0.x; b[1] = b0.y;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float av = q == 0 ? a[i].x : (q == 1 ? a[i].y : (q == 2 ? a[i].z : a[i].w));
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
                }
            }
        }
    }
    cp_async_wait<0>();

    // --- epilogue -------------------------------------------------------------------------------------
    int col[TN];
    if constexpr (TN == 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { col[j] = n0 + tx * 4 + j; col[4 + j] = n0 + 64 + tx * 4 + j; }
    } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) col[j] = n0 + tx * TN + j;
    }
    float bias[TN], sc[TN], sh[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        bias[j] = p.bias[col[j]];
        sc[j] = (p.out_act || p.out_hi) ? p.scale[col[j]] : 1.f;
        sh[j] = (p.out_act || p.out_hi) ? p.shift[col[j]] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = m0 + ty + 16 * i;
        if (r >= p.m_total) continue;
        float v[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) v[j] = acc[i][j] + bias[j];
        if (p.resid) {
            const int b = r / p.Jout;
            const int j0 = r - b * p.Jout;
            const float* rr = p.resid + ((size_t)b * p.Jres + p.resid_off + (p.resid_per_j ? j0 : 0)) * p.Cout;
#pragma unroll
            for (int j = 0; j < TN; ++j) v[j] += rr[col[j]];
        }
        const size_t o = (size_t)r * p.Cout;
        if (p.out_raw) {
            if constexpr (TN == 8) {
                *reinterpret_cast<float4*>(p.out_raw + o + col[0]) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(p.out_raw + o + col[4]) = make_float4(v[4], v[5], v[6], v[7]);
            } else if constexpr (TN == 4) {
                *reinterpret_cast<float4*>(p.out_raw + o + col[0]) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
                *reinterpret_cast<float2*>(p.out_raw + o + col[0]) = make_float2(v[0], v[1]);
            }
        }
        if (p.out_act || p.out_hi) {
            float w[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) w[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
            if (p.out_hi) {
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const __nv_bfloat16 h = __float2bfloat16_rn(w[j]);
                    p.out_hi[o + col[j]] = __bfloat16_as_ushort(h);
                    p.out_lo[o + col[j]] = __bfloat16_as_ushort(__float2bfloat16_rn(w[j] - __bfloat162float(h)));
                }
            }
            if (p.out_act) {
            if constexpr (TN == 8) {
                *reinterpret_cast<float4*>(p.out_act + o + col[0]) = make_float4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<float4*>(p.out_act + o + col[4]) = make_float4(w[4], w[5], w[6], w[7]);
            } else if constexpr (TN == 4) {
                *reinterpret_cast<float4*>(p.out_act + o + col[0]) = make_float4(w[0], w[1], w[2], w[3]);
            } else {
                *reinterpret_cast<float2*>(p.out_act + o + col[0]) = make_float2(w[0], w[1]);
            }
            }
        }
    }
}

template <int BN>
int launch(yoho_ctx* ctx, KArgs& k, cudaStream_t st) {
    using L = SmemLayout<BN>;
    // per-device attribute; cheap enough to set on every launch (one process may drive several devices)
    YCHECK(cudaFuncSetAttribute(gconv_f32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
    k.n_tiles = k.Cout / BN;
    const int m_tiles = (k.m_total + BM - 1) / BM;
    gconv_f32_kernel<BN><<<m_tiles * k.n_tiles, THREADS, L::BYTES, st>>>(k);
    ctx->launches++;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

}  // namespace

int gconv_simt_forward(yoho_ctx* ctx, const GLayer& Lr, const GConvArgs& a, cudaStream_t st) {
    YARG(Lr.w && Lr.cin % BK == 0 && Lr.cout % 32 == 0 && a.Jout * Lr.taps <= 64 * 13);
    if (a.B <= 0) return YOHO_OK;
    KArgs k;
    k.act = a.act; k.w = Lr.w; k.bias = Lr.bias; k.idx = a.idx;
    k.B = a.B; k.Jin = a.Jin; k.Jout = a.Jout; k.Cin = Lr.cin; k.Cout = Lr.cout; k.taps = Lr.taps;
    k.resid = a.resid; k.Jres = a.Jres; k.resid_off = a.resid_off; k.resid_per_j = a.resid_per_j;
    k.out_raw = a.out_raw; k.out_act = a.out_act; k.scale = a.scale; k.shift = a.shift;
    k.out_hi = (unsigned short*)a.out_hi; k.out_lo = (unsigned short*)a.out_lo;
    k.m_total = a.B * a.Jout;
    const int m_tiles = (k.m_total + BM - 1) / BM;
    // widest tile that still gives every SM at least two CTAs; narrow tiles for small-M layers
    const int want = 2 * ctx->num_sms;
    if (Lr.cout % 128 == 0 && m_tiles * (Lr.cout / 128) >= want) return launch<128>(ctx, k, st);
    if (Lr.cout % 64 == 0 && m_tiles * (Lr.cout / 64) >= want) return launch<64>(ctx, k, st);
    return launch<32>(ctx, k, st);
}
