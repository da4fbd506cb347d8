// Shared declarations of the yoho_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../include/yoho_b200.h"

#define YG 60
#define YT 13
#define YF 32

void yoho_set_error(const char* fmt, ...);

#define YCHECK(expr)                                                                            \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            yoho_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return YOHO_ERR_CUDA;                                                               \
        }                                                                                       \
    } while (0)

#define YARG(cond)                                                                 \
    do {                                                                           \
        if (!(cond)) {                                                             \
            yoho_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);  \
            return YOHO_ERR_ARG;                                                   \
        }                                                                          \
    } while (0)

// One group-convolution (or 1x1) layer, packed for the kernels.
struct GLayer {
    int cin = 0, cout = 0, taps = 0;
    int prof_class = 0;
    float* w = nullptr;       // [taps][cin][cout] fp32
    float* bias = nullptr;    // [cout]
    // tcgen05 path: bf16 hi/lo split of the weights, [taps][cout][cin] (K-major B operand)
    void* w_hi = nullptr;
    void* w_lo = nullptr;
    int tc_dense = 0;         // 1: tensor-core layer with taps == 1 (dense GEMM, e.g. PartI layer 4 "all taps at once")
};

// Folded eval-mode BatchNorm: y = x*scale + shift.
struct GBn {
    int c = 0;
    float* scale = nullptr;
    float* shift = nullptr;
};

struct yoho_ctx {
    int device = 0;
    int num_sms = 148;
    int gconv_impl = 0;
    int tc_flags = 3 | 256;         // tuning flags (include/yoho_b200.h yoho_set_tuning): bit0 noinc producers (gconv_tc.cu; bit1 ignored); 256 = all-Fourier PartI with the tcgen05 transform kernel (fourier_tc.cu)
    int split_min_k = 1024;         // tensor-core GEMM: accumulation chains of at least this many K elements use the split accumulators (tuning key 1)
    int64_t launches = 0;
    // group tables on the device
    double* d_rot = nullptr;        // [60][9] f64
    float* d_rot32 = nullptr;       // [60][9] f32 (Rgroup.astype(float32), tests/extractor.py:110)
    uint8_t* d_perm_t = nullptr;    // [60 g][60 a] = P[a][g]
    uint8_t* d_perm = nullptr;      // [60 a][60 g] = P[a][g]
    int* d_idx_full = nullptr;      // [60][13]
    int* d_idx_full_inv = nullptr;  // [60 j][13 k] = the g with N[g][k] == j (train.cu: backward-data; built on first use)
    int* d_idx_p2_init = nullptr;   // [45][13] -> 60
    int* d_idx_p2_a = nullptr;      // [13][13] -> 45
    int* d_idx_p2_b = nullptr;      // [1][13]  -> 13
    int* d_idx_one = nullptr;       // [1][1] = 0
    int hop2_zero_pos = 0;
    // PartI
    bool has_p1 = false;
    GLayer p1_in, p1_a, p1_b, p1_out;
    GLayer p1_out_cat;              // PartI layer 4 as one dense GEMM: W_cat[c][k*32+o] = W_k[c][o], 416 -> 512 columns
    int* d_idx_ident = nullptr;     // [60][1] identity
    GBn p1_bn_a, p1_bn_b, p1_bn_out;
    // PartI layers 2+3 in the group-Fourier domain (implementation 3)
    bool has_p1f = false;
    int nf = 0;
    int fd[8] = {0};
    GLayer p1f_a[8], p1f_b[8];
    int* d_fidx[8] = {nullptr};
    int* d_fomap[8] = {nullptr};
    // ... and layers 1 and 4 (optional): 32 -> d*256 and 256 -> d*32 (the latter zero-padded to 256 columns, output-row table
    // with 8 column groups of 32 per GEMM row)
    bool has_p1f_io = false;
    GLayer p1f_in[8], p1f_out[8];
    int* d_fomap_out[8] = {nullptr};
    float* d_p1_bias31 = nullptr;   // [256] bias of layer 3 + bias of layer 1 (the shortcut's bias, added in the group domain)
    float* d_F = nullptr;           // [60 m][60 g] FP32 (input transform and the finalize kernel's inverse transform)
    // forward / inverse transform matrices for the tcgen05 transform kernel: [64 out][64 in] bf16 hi/lo (row = OUTPUT index)
    void* d_fwd_hi = nullptr; void* d_fwd_lo = nullptr;     // forward: rows m (coefficient), cols g
    void* d_inv_hi = nullptr; void* d_inv_lo = nullptr;     // inverse: rows g, cols m
    // PartII
    bool has_p2 = false;
    GLayer p2_init, p2_a, p2_b, p2_fc1, p2_fc2, p2_fc3;
    GLayer p2_fc2_pad;              // head layer 512 -> 128 zero-padded to 256 columns (one tensor-core N tile); p2_bn2_pad likewise
    GBn p2_bn2_pad;
    GLayer p2_b_split[5];           // p2_b cut along its 13 taps (3,3,3,2,2): five partial GEMMs in one launch (only 22 row tiles at M = 2800)
    GBn p2_bn_init, p2_bn_a, p2_bn_b, p2_bn1, p2_bn2;
    // split-phase pair calls (pair.cu): FIFO of match counts in flight — pinned host slot + event per begun pair
    static constexpr int kPairRing = 8;
    int32_t* pair_M_pinned = nullptr;           // [kPairRing] cudaHostAlloc
    cudaEvent_t pair_ev[kPairRing] = {nullptr};
    int pair_head = 0, pair_tail = 0;            // begin pushes at head, end pops at tail
    // grow-only workspace
    void* ws = nullptr;
    size_t ws_bytes = 0;
    // optional per-layer event timing (yoho_profile_enable)
    bool prof_on = false;
    struct ProfRec { cudaEvent_t a, b; int cls; double flops; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> prof_pool;
};

int yoho_ws_reserve(yoho_ctx* ctx, size_t bytes);

// ---- generic gather-GEMM (gconv_simt.cu) ---------------------------------------------------------
struct GConvArgs {
    const float* act;        // [B][Jin][Cin], already activated if the layer has a pre-activation
    const int* idx;          // [Jout][taps] -> row in [0,Jin)
    int B, Jin, Jout;
    const float* resid;      // nullable: [B][Jres][Cout]; row(b,j) = b*Jres + resid_off + (Jres_per_j ? j : 0)
    int Jres, resid_off, resid_per_j;
    float* out_raw;          // nullable [B][Jout][Cout]   = acc + bias (+ resid)
    float* out_act;          // nullable [B][Jout][Cout]   = relu(out_raw*scale + shift)
    const float* scale;      // folded BN of the NEXT layer's pre-activation (with out_act)
    const float* shift;
    // bf16 hi/lo split of the activations (tensor-core path): inputs of this layer / outputs for the next
    const void* act_hi;      // [B][Jin][Cin] bf16
    const void* act_lo;
    void* out_hi;            // nullable [B][Jout][Cout] bf16: hi/lo of relu(out_raw*scale + shift)
    void* out_lo;
    int n_valid;             // tensor-core path: only columns < n_valid are written (0 = all)
    // tensor-core path, group-Fourier layers: remapped output rows (see gconv_tc.cu TcArgs)
    const int* omap;
    int ogroup, out_J;
};
int gconv_forward(yoho_ctx* ctx, const GLayer& L, const GConvArgs& a, cudaStream_t st);

// ---- small device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    // src-size 0 zero-fills the 16 destination bytes (used for rows past the end of the tile)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(valid ? 16 : 0));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
