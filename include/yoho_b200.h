/* yoho_b200 — C ABI of the B200-native YOHO descriptor + registration hot path.
 *
 * The reference (HpWang-whu/YOHO) has NO FFI on this path: its boundary is a set of pure-Python plugin
 * registries (name2network / name2extractor / name2matcher / name2estimator).  This header is
 * the C boundary a maintainer binds underneath those Python classes (ctypes stub: INTEGRATION.md); every
 * entry point names the reference function whose arithmetic it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.  Every data pointer is a DEVICE pointer unless
 *     the parameter name ends in _host.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - return 0 on success, negative yoho_status otherwise; yoho_last_error() gives the message (thread-local).
 *   - the context owns the packed weights, the group tables and a grow-only device workspace: steady-state
 *     calls allocate nothing.  One context per (device, stream-of-use); calls on one context are not
 *     re-entrant.
 *   - tensor layouts at the boundary are the reference's own: group features [K,32,60] float32 with the
 *     group axis innermost (tests/extractor.py:48,60), matches int64 [M,2], rotation index int64 [M],
 *     keypoints float64 [K,3], transforms float64 [3,4] row-major.
 *   - there is no CPU fallback anywhere: without a CUDA device every call returns YOHO_ERR_CUDA.
 */
#ifndef YOHO_B200_H
#define YOHO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YOHO_ABI_VERSION 3
#define YOHO_G 60      /* group order */
#define YOHO_TAPS 13   /* group-convolution kernel support */
#define YOHO_F 32      /* descriptor channels */

typedef enum {
    YOHO_OK = 0,
    YOHO_ERR_CUDA = -1,        /* a CUDA runtime call failed / no device */
    YOHO_ERR_ARG = -2,         /* bad argument */
    YOHO_ERR_NOWEIGHTS = -3,   /* forward called before the matching *_load ("No model exists", tests/extractor.py:34,122) */
    YOHO_ERR_ALLOC = -4
} yoho_status;

typedef struct yoho_ctx yoho_ctx;

/* BatchNorm2d (eval mode) parameters, host float32[C] each; eps is the reference's 1e-5. */
typedef struct {
    const float* weight_host;
    const float* bias_host;
    const float* running_mean_host;
    const float* running_var_host;
} yoho_bn_host;

/* Conv2d(C,O,(1,taps)) parameters in the reference's layout: weight [O,C,1,taps] float32, bias [O]. */
typedef struct {
    const float* weight_host;
    const float* bias_host;
} yoho_conv_host;

/* PartI_test state-dict (utils/network.py:67-105; key names in SURVEY.md §8a). */
typedef struct {
    yoho_conv_host conv_in;       /* PartI_net.Conv_in.0                           32 -> 256 */
    yoho_bn_host bn_a;            /* PartI_net.SO3_Conv_layers.0.comb_layer_in.0   256 */
    yoho_conv_host conv_a;        /* ...comb_layer_in.2                            256 -> 512 */
    yoho_bn_host bn_b;            /* ...comb_layer_out.0                           512 */
    yoho_conv_host conv_b;        /* ...comb_layer_out.2                           512 -> 256 */
    yoho_bn_host bn_out;          /* PartI_net.Conv_out.comb_layer.0               256 */
    yoho_conv_host conv_out;      /* PartI_net.Conv_out.comb_layer.2               256 -> 32 */
} yoho_part1_weights;

/* PartII_test state-dict (utils/network.py:218-278). */
typedef struct {
    yoho_bn_host bn_init;         /* Conv_init.comb_layer.0                        128 */
    yoho_conv_host conv_init;     /* Conv_init.comb_layer.2                        128 -> 256 */
    yoho_bn_host bn_a;            /* PartII_SO3_Conv_layers.0.comb_layer_in.0      256 */
    yoho_conv_host conv_a;        /* ...comb_layer_in.2                            256 -> 512 */
    yoho_bn_host bn_b;            /* ...comb_layer_out.0                           512 */
    yoho_conv_host conv_b;        /* ...comb_layer_out.2                           512 -> 256 */
    yoho_conv_host fc1;           /* PartII_To_R_FC.0   [512,256,1,1] */
    yoho_bn_host bn1;             /* PartII_To_R_FC.1   512 */
    yoho_conv_host fc2;           /* PartII_To_R_FC.3   [128,512,1,1] */
    yoho_bn_host bn2;             /* PartII_To_R_FC.4   128 */
    yoho_conv_host fc3;           /* PartII_To_R_FC.6   [4,128,1,1] */
} yoho_part2_weights;

int yoho_abi_version(void);
const char* yoho_last_error(void);

/* A0 — group tables (utils/network.py:72-74,223-226; tests/extractor.py:67,110; tests/estimator.py:283-284).
 * rotation_host float64[60*3*3], perm_host int32[60*60] (P[a][b]), nei_host int32[60*13] (N[g][k]). */
int yoho_ctx_create(int device, const double* rotation_host, const int32_t* perm_host, const int32_t* nei_host,
                    yoho_ctx** out);
int yoho_ctx_destroy(yoho_ctx* ctx);

/* Checkpoint packing: folds eval-mode BN into (scale, shift), reorders W[o,c,0,k] -> W_k[c][o], uploads.
 * Replaces torch's load_state_dict for PartI_test / PartII_test (tests/extractor.py:26-34,113-122). */
int yoho_part1_load(yoho_ctx* ctx, const yoho_part1_weights* w);
int yoho_part2_load(yoho_ctx* ctx, const yoho_part2_weights* w);

/* Group-Fourier form of the PartI layers (implementation 3).  F_host: orthogonal 60x60 transform, row m = Fourier
 * coefficient (irrep, l, j), column g = group element (yoho_b200/fourier.py).  Per real irrep (dims 1,3,3,4,5): the weights of
 * PartI layer 2 (w_a_host [d][256][d*512]) and layer 3 (w_b_host [d][512][d*256]) as d-tap gather-GEMM weights with column
 * n = i*O + o, the input-row table idx_host[j*d + l] and the output-row table omap_host[j*d + i].
 * Optional (both or neither; NULL keeps layers 1 and 4 as direct 13-tap convolutions): layer 1 (w_in_host [d][32][d*256]) and
 * layer 4 (w_out_host [d][256][d*32]) in the same form — the whole stack then stays in the Fourier domain between the
 * BatchNorm/ReLU points, including the shortcut of the residual block (the transform is linear). */
typedef struct {
    int d, off;
    const float* w_a_host;
    const float* w_b_host;
    const int32_t* idx_host;
    const int32_t* omap_host;
    const float* w_in_host;
    const float* w_out_host;
} yoho_fourier_irrep;
int yoho_part1_load_fourier(yoho_ctx* ctx, const float* F_host, int n_irreps, const yoho_fourier_irrep* irreps);

/* Implementation of the group-convolution layers: 0 = FP32 SIMT (default), 1 = tcgen05 split-BF16,
 * 2 = tcgen05 split-BF16 with the small (lo) products in a separate TMEM accumulator (shorter rounding chain),
 * 3 = as 2, with all four PartI layers evaluated in the group-Fourier domain (needs yoho_part1_load_fourier with the layer-1/4
 * weights; otherwise, and for fewer than 128 keypoints, the direct layers of implementation 2 run). */
int yoho_set_gconv_impl(yoho_ctx* ctx, int impl);

/* Tuning knobs (defaults are the measured best).  key 1 = smallest accumulation length K (taps x Cin) for which the tensor-core GEMM
 * keeps the hi*hi products and the cross products in two TMEM accumulators (shorter rounding chain, no epilogue overlap).  key 0 = flag word: 1 = non-blocking producer protocol of the tensor-core GEMM
 * (2: ignored); 256 = PartI entirely in the group-Fourier domain with the tcgen05 transform kernel (default on; cleared, or with
 * 512 set, implementation 3 runs the direct 13-tap tensor-core layers of implementation 2); 1024 = PartII last group convolution
 * as one GEMM instead of five tap-split partial GEMMs; 2048 = PartI output side (inverse transform of the layer-4 coefficients,
 * residual, norms, pools) on the tcgen05 transform machinery, 4096 = the same with shared-memory staged row accesses (both
 * parity-green, not faster: default off); 8192 = FP32 SIMT input / output side of the all-Fourier PartI instead of the register-resident
 * warp-MMA kernels (test twin); 16384 = FP32 SIMT 1-NN search instead of the tensor-core search with exact verification (test twin);
 * 32768 = the grouped tensor-core launches of PartI run as 2-CTA clusters whose CTAs take adjacent row tiles of one column
 * tile and share every weight tile (each fetches half of it, TMA multicast to both; bit-identical results, measured neutral:
 * 2.22 ms against 2.23 ms per 5000-keypoint PartI, so it stays off).
 * The warp-MMA / SIMT transform kernels of round 1 (flags 4-128) were removed. */
int yoho_set_tuning(yoho_ctx* ctx, int key, int value);

/* A1-A6 — PartI_test.forward (utils/network.py:86-105,140-147) on B keypoints.
 * x [B,32,60] -> eqv [B,32,60] (unit norm over channels per (b,g)), inv [B,32] (may be NULL),
 * desc_mean [B,32] = mean_g eqv, the matcher's descriptor (tests/matcher.py:35-36; may be NULL). */
int yoho_part1_forward(yoho_ctx* ctx, const float* x, int B, float* eqv, float* inv, float* desc_mean,
                       void* stream);

/* B1 — matcher descriptor from a stored eqv: np.mean(feats, axis=-1) in float32 (tests/matcher.py:35-36),
 * same pairwise summation order as numpy.  eqv [K,32,60] -> desc [K,32]. */
int yoho_group_mean(yoho_ctx* ctx, const float* eqv, int K, float* desc, void* stream);

/* B2 — knn_module.KNN(1)(target, source) (utils/knn_search.py:17-24,26-66,138-154), dist_type 'L2':
 * for every source row (m rows, F floats, F <= 32) the nearest target row: dist = sqrt(sum (s-t)^2 + 1e-7)
 * in float32, ties -> lowest target index.  dist [m] float32, idx [m] int64. */
int yoho_nn1(yoho_ctx* ctx, const float* source, int m, const float* target, int n, int F, float* dist,
             int64_t* idx, void* stream);

/* B1+B2 — matcher_dual.match (tests/matcher.py:37-48): both 1-NN searches in one pass over the Ka x Kb
 * distance tiles, mutual filter, pairs written in ascending fragment-0 index.
 * dA [Ka,32], dB [Kb,32]; pairs int64 [min(Ka,Kb),2] capacity; n_pairs device int32[1].
 * nnA int32[Ka] / nnB int32[Kb] (argmin of each A row in B / each B row in A) may be NULL. */
int yoho_mutual_nn(yoho_ctx* ctx, const float* dA, int Ka, const float* dB, int Kb, int64_t* pairs,
                   int32_t* n_pairs, int32_t* nnA, int32_t* nnB, void* stream);

/* C1 — extractor_dr_index.Batch_Des2R_torch (tests/extractor.py:74-78): for match m,
 * cor[a] = sum_{f,g} des1[row1(m), f, P[a][g]] * des2[row2(m), f, g], idx[m] = argmax_a (ties -> lowest a).
 * rows1/rows2: int64 row ids with element stride `row_stride` (pass the [M,2] match array with stride 2 and
 * the column offset applied to the pointer), NULL = identity.  cor_out [M,60] may be NULL.
 * The reference calls it with des1 = eqv of fragment 1, des2 = eqv of fragment 0 (tests/extractor.py:97-99). */
int yoho_rot_argmax(yoho_ctx* ctx, const float* des1, const int64_t* rows1, const float* des2,
                    const int64_t* rows2, int row_stride, int M, int64_t* idx, float* cor_out, void* stream);

/* D1-D3 — extractor_PartII.batch_create + PartII_test.forward + the quaternion/transform post-loops
 * (tests/extractor.py:125-138,185-201; utils/network.py:259-278; utils/r_eval.py:94-110).
 * fcgf0/yoho0: fragment id0 ("A") tensors [K0,32,60]; fcgf1/yoho1: fragment id1 ("B").
 * pairs int64 [M,2] (row in A, row in B); pre_idx int64 [M]; kps0/kps1 float64 [K,3] (NULL -> no trans).
 * quat [M,4] float32 (w,x,y,z); trans [M,3,4] float64 = [R(q) Rgroup[idx] | k0 - R k1] (may be NULL).
 * Evaluates only the receptive field of group element 0 (45/13/1 elements), which is exact. */
int yoho_part2_forward(yoho_ctx* ctx, const float* fcgf0, const float* fcgf1, const float* yoho0,
                       const float* yoho1, const int64_t* pairs, const int64_t* pre_idx, int M,
                       const double* kps0, const double* kps1, float* quat, double* trans, void* stream);

/* Gather matched keypoints: out0[m] = kps0[pairs[m][0]], out1[m] = kps1[pairs[m][1]] (tests/estimator.py:98-99). */
int yoho_gather_kps(yoho_ctx* ctx, const double* kps0, const double* kps1, const int64_t* pairs, int M,
                    double* out0, double* out1, void* stream);

/* E1 + draw — yohoc.DR_statictic (tests/estimator.py:34-51) and the hypothesis draws of the RANSAC loop
 * (:119-126: categorical bin, then three members WITH replacement) generated on the device from a
 * counter-based Philox4x32-10 stream.  Same distribution as the reference, not the same MT19937 stream;
 * bit-parity runs pass a host-drawn list to yoho_c_ransac instead.
 * status (device int32[1]): 0 ok, 1 = degenerate statistics (reference returns None -> identity, recalltime 50001),
 * 2 = dr_index holds a value outside [0,60) (the reference would raise IndexError, tests/estimator.py:39): no draws are made,
 * hyp is zero-filled and yoho_register_pair returns the identity like for status 1. */
int yoho_c_draw(yoho_ctx* ctx, const int64_t* dr_index, int M, int iters, uint64_t seed, int32_t* hyp,
                int32_t* status, void* stream);

/* E2-E4 — yohoc.ransac inner loop over a pre-drawn hypothesis list (tests/estimator.py:55-70,119-137).
 * k0/k1 float64 [M,3] matched keypoints; hyp int32 [iters,3] match ids; signs int8[iters] or NULL
 * (0 = sign rule of DESIGN.md, +1/-1 = forced determinant of the null-space completion, 2 = take the caller's transform
 * fixed[h] for this hypothesis); fixed float64 [iters,3,4] or NULL.  The reference's R = V U^T depends on LAPACK's
 * arbitrary sign for the null-space singular pair of the rank-2 cross-covariance (and on its arbitrary completion of a
 * rank-1 one, when a match was drawn twice): a reference-identical run passes sign(det(V U^T)) of np.linalg.svd per
 * hypothesis and LAPACK's transform for the rank-deficient triplets (yoho_b200/estimator.py does).
 * Outputs (device): T float64[12] ([I|0] if nothing scores), best_iter int32 (0-based, -1 if none),
 * n_inl int32, mask uint8[M] (inliers of the winner), counts int32[iters] (may be NULL). */
int yoho_c_ransac(yoho_ctx* ctx, const double* k0, const double* k1, int M, const int32_t* hyp,
                  const int8_t* signs, const double* fixed, int iters, double inlier_dist, double* T,
                  int32_t* best_iter, int32_t* n_inl, uint8_t* mask, int32_t* counts, void* stream);

/* Device-side random evaluation order for YOHO-O (np.random.shuffle(index), tests/estimator.py:321-323). */
int yoho_o_order(yoho_ctx* ctx, int M, uint64_t seed, int32_t* order, void* stream);

/* E5 — yohoo.ransac scoring (tests/estimator.py:321-336): trans float64 [Mt,3,4] hypotheses, order int32[H]
 * (indices into trans, NULL = 0..H-1), first strictly-best kept.  best_iter is the position in `order`. */
int yoho_o_score(yoho_ctx* ctx, const double* k0, const double* k1, int M, const double* trans,
                 const int32_t* order, int H, double inlier_dist, double* T, int32_t* best_iter,
                 int32_t* n_inl, uint8_t* mask, int32_t* counts, void* stream);

/* "Next" row — the in-memory pair pipeline as ONE call: what Evaluator_PartI/II.run_onescene
 * (tests/evaluator.py:41-47,112-117) does for one pair through files — Extract (both fragments) -> match -> PartI_Rindex ->
 * yohoc.ransac -> PartII_R_pre -> yohoo.ransac — with everything device resident and a single host synchronisation (the
 * match count).  Same stages, order and arguments as calling yoho_part1_forward x2, yoho_mutual_nn, yoho_rot_argmax,
 * yoho_gather_kps, yoho_c_draw, yoho_c_ransac, yoho_part2_forward, yoho_o_order and yoho_o_score one by one (bit-identical
 * results).  All pointers are device pointers; M-sized buffers have capacity min(Ka, Kb) rows.
 * have_part1 = 1: eqvA/eqvB/descA/descB already hold the PartI outputs of the two fragments (amortised regime:
 * tests/extractor.py:46-47 runs PartI once per fragment); 0: they are computed here.
 * Degenerate rotation statistics (c_status = 1) give T_c = [I|0], c_best = -1; M = 0 gives both transforms = [I|0]. */
typedef struct {
    const float* featA; const float* featB;      /* FCGF group features [Ka,32,60], [Kb,32,60] */
    const double* kpsA; const double* kpsB;      /* keypoints [Ka,3], [Kb,3] */
    int Ka, Kb;
    int have_part1;
    int c_iters, o_iters;                        /* YOHO-C hypotheses / YOHO-O evaluation cap (tests/estimator.py max_iter) */
    double c_dist, o_dist;                       /* ransac_c_inlinerdist / ransac_o_inlinerdist */
    uint64_t seed;                               /* Philox seed of the device-side draws (yoho_c_draw, yoho_o_order) */
    float* eqvA; float* eqvB;                    /* [K,32,60] */
    float* descA; float* descB;                  /* [K,32] matcher descriptors (mean over the group axis) */
    int64_t* pairs; int32_t* n_pairs;            /* [cap,2], [1] */
    int64_t* dr_index;                           /* [cap] */
    double* k0; double* k1;                      /* [cap,3] matched keypoints */
    int32_t* hyp; int32_t* c_status;             /* [c_iters,3], [1] */
    double* T_c; int32_t* c_best; int32_t* c_inl; uint8_t* c_mask;      /* [12], [1], [1], [cap] */
    float* quat; double* trans;                  /* [cap,4], [cap,12] */
    int32_t* order;                              /* [cap] */
    double* T_o; int32_t* o_best; int32_t* o_inl; uint8_t* o_mask;      /* [12], [1], [1], [cap] */
} yoho_pair_io;
int yoho_register_pair(yoho_ctx* ctx, const yoho_pair_io* io, int32_t* M_host, void* stream);

/* The same pair in two phases, for a SEQUENCE of pairs (a dataset: tests/evaluator.py:41-47 loops over dataset.pair_ids).
 * _begin queues PartI x2 + the mutual matching and starts the match count on its way to the host, without waiting;
 * _end(io of the OLDEST begun pair) waits for that count and queues the rest.  Calling begin(i+1) before end(i) puts ~4 ms of
 * device work between a pair's matching and the moment the host needs its count, so the single host synchronisation of a pair
 * no longer idles the device (yoho_register_pair = begin + end back to back; results are identical).  FIFO order, at most 8
 * pairs in flight, every pair with its own output buffers; one stream. */
int yoho_register_pair_begin(yoho_ctx* ctx, const yoho_pair_io* io, void* stream);
int yoho_register_pair_end(yoho_ctx* ctx, const yoho_pair_io* io, int32_t* M_host, void* stream);

/* "Next" row — the tail of the group-feature lift, YOHO_testset.py:153-166: for every group rotation g,
 * rotate the keypoints (Keys @ R_g^T, float64), 1-NN of each into that rotation's down-sampled cloud (float64 distances
 * against float32 points, first minimal index), gather the 32-d backbone feature: out[k,:,g] = feats[offsets[g] + nn, :].
 * kps float64 [K,3]; pts float32 [sum n_g,3] and feats float32 [sum n_g,32] concatenated over g; offsets DEVICE int32[61];
 * out float32 [K,32,60]; nn_out int64 [60,K] (may be NULL). */
int yoho_lift_group_features(yoho_ctx* ctx, const double* kps, int K, const float* pts, const float* feats,
                             const int32_t* offsets, float* out, int64_t* nn_out, void* stream);

/* "Next" row — evaluation metrics on the device.
 * yoho_fmr_batch replaces the per-pair body of Evaluator_PartI.Feature_match_Recall (tests/evaluator.py:57-66; also
 * utils/utils.py:221-228 evaluate_the_match): keys0/keys1 float64 [Mtot,3] are the MATCHED keypoints of all pairs,
 * concatenated; pair p owns rows offsets[p]..offsets[p+1]-1 (DEVICE int64[n_pairs+1]); gt float64 [n_pairs,4,4] (R|t with
 * R @ keys1 + t = keys0; pass a 3x4 ground truth with the row 0 0 0 1); counts int32[n_pairs] = number of matches with
 * ||keys0 - gt(keys1)|| < threshold.  The pair ratio is counts/M and FMR = mean(ratio > fmr_ratio) (host, exact). */
int yoho_fmr_batch(yoho_ctx* ctx, const double* keys0, const double* keys1, const int64_t* offsets, const double* gt,
                   int n_pairs, double threshold, int32_t* counts, void* stream);

/* Replaces the numeric body of utils/RR_cal.py evaluate_registration / benchmark for aligned lists of n pairs:
 * est, gt float64 [n,4,4]; info float64 [n,6,6] (needed only for p).  Outputs (each may be NULL): p[n] = the Redwood error
 * computeTransformationErr(inv(gt) @ est, info) BEFORE the square root (RR_cal.py:48-65,273,289; nibabel mat2quat's
 * largest-eigenvector quaternion); rre_deg[n] = rotation_error(gt_R, est_R) in degrees with the reference's float32 pi
 * (RR_cal.py:13-33); rte[n] = translation_error (RR_cal.py:35-46).  A singular gt gives p = NaN. */
int yoho_registration_errors(yoho_ctx* ctx, const double* est, const double* gt, const double* info, int n, double* p,
                             double* rre_deg, double* rte, void* stream);

/* "Next" row — training-time twins of the hot kernels, FP32, reference layouts ([B,C,60], group axis innermost;
 * weight [O,C,1,13] and bias [O] as DEVICE pointers in the reference's own layout: they change every optimiser step).
 * yoho_gconv_train_forward: y[b,o,g] = bias[o] + sum_{c,k} weight[o,c,0,k] x[b,c,N[g][k]] — Comb_Conv's gather + Conv2d(C,O,(1,13))
 * (utils/network.py:12-21,46-52,80-84) without the 13x gathered tensor; bias may be NULL.
 * yoho_gconv_train_backward: dx[b,c,j] = sum_{o,k} weight[o,c,0,k] dy[b,o,Ninv_k(j)] (every tap column of N is a permutation of
 * the group), dweight[o,c,0,k] = sum_{b,g} dy[b,o,g] x[b,c,N[g][k]], dbias[o] = sum_{b,g} dy[b,o,g]; each output may be NULL
 * (dbias is only written together with dweight).  Deterministic (fixed summation order, no atomics). */
int yoho_gconv_train_forward(yoho_ctx* ctx, const float* x, const float* weight, const float* bias, int B, int C, int O,
                             float* y, void* stream);
int yoho_gconv_train_backward(yoho_ctx* ctx, const float* x, const float* weight, const float* dy, int B, int C, int O,
                              float* dx, float* dweight, float* dbias, void* stream);
/* Gradient of the rotation correlation cor[m,a] = sum_{f,g} des1[m,f,P[a][g]] des2[m,f,g] — PartI_train.Des2DR
 * (utils/network.py:115-118) and Batch_hard_Rindex_loss.eqvloss (train/loss_val.py:27-31); the forward is yoho_rot_argmax's cor_out.
 * des1, des2, grad_des1, grad_des2 [M,32,60]; grad_cor [M,60]; either gradient may be NULL. */
int yoho_rot_correlation_backward(yoho_ctx* ctx, const float* des1, const float* des2, const float* grad_cor, int M,
                                  float* grad_des1, float* grad_des2, void* stream);

/* Launch accounting for bench.py's "gpu_launches": kernels launched by this context since creation. */
int64_t yoho_launch_count(const yoho_ctx* ctx);

/* Per-layer timing of the group-convolution launches with CUDA events on the launching stream (bench.py's
 * roofline object).  enable=1 starts/clears recording, enable=0 stops.  yoho_profile_read synchronises the
 * device and returns, for layer class c in [0, YOHO_PROF_CLASSES): total milliseconds, launches and algorithmic
 * FLOPs (2 * rows * taps * Cin * Cout per launch).  Classes: 0..3 = PartI layers 1..4, 4..6 = PartII group
 * convolutions (init, a, b), 7 = PartII 1x1 head layers, 8 = group-Fourier transforms (implementation 3; classes 1 and 2 then
 * hold the Fourier-domain GEMMs and their FLOPs are the executed ones). */
/* Test hook: one PartI/PartII group-convolution layer in isolation on FP32 activations act [B,60,Cin]
 * (full 60-element index table), raw output [B,60,Cout] = conv + bias.  layer: 0..3 = PartI layers 1..4,
 * 4..6 = PartII init/a/b.  impl as in yoho_set_gconv_impl. */
int yoho_debug_layer(yoho_ctx* ctx, int layer, int impl, const float* act, int B, float* out_raw, void* stream);

#define YOHO_PROF_CLASSES 9
int yoho_profile_enable(yoho_ctx* ctx, int enable);
int yoho_profile_read(yoho_ctx* ctx, double* ms_host, int64_t* launches_host, double* flops_host);

#ifdef __cplusplus
}
#endif
#endif /* YOHO_B200_H */
