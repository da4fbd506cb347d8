// One fragment pair, FCGF group features -> YOHO-C and YOHO-O transforms, in ONE C-ABI call (include/yoho_b200.h,
// yoho_register_pair): the same stages, in the same order and with the same arguments, as the per-stage entry points the
// Python mirror of tests/evaluator.py:41-47,112-117 would call one by one.  About twenty short kernels follow the single host
// synchronisation of a pair (the match count M sizes them); launched from here they are ~4 us apart instead of the 20-30 us of a
// Python / ctypes round trip each, so the device does not starve after the synchronisation and the host cost of a pair is one call.
#include "common.cuh"

namespace {

// Degenerate rotation statistics (DR_statictic returns None, tests/estimator.py:41-51,107-108): identity transform, no winner.
__global__ void c_finish_kernel(const int32_t* __restrict__ status, double* __restrict__ T, int32_t* __restrict__ best) {
    if (threadIdx.x == 0 && *status != 0) {
        for (int i = 0; i < 12; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;      // [I | 0], row-major 3x4
        *best = -1;
    }
}

__global__ void identity_kernel(double* __restrict__ Tc, double* __restrict__ To, int32_t* __restrict__ cb, int32_t* __restrict__ ob,
                                int32_t* __restrict__ ci, int32_t* __restrict__ oi) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < 12; ++i) { Tc[i] = (i % 5 == 0) ? 1.0 : 0.0; To[i] = Tc[i]; }
        *cb = -1; *ob = -1; *ci = 0; *oi = 0;
    }
}

}  // namespace

static int pair_check(yoho_ctx* ctx, const yoho_pair_io* io) {
    YARG(ctx && io && io->featA && io->featB && io->kpsA && io->kpsB && io->Ka >= 0 && io->Kb >= 0);
    YARG(io->eqvA && io->eqvB && io->descA && io->descB && io->pairs && io->n_pairs && io->dr_index && io->k0 && io->k1);
    YARG(io->hyp && io->c_status && io->T_c && io->c_best && io->c_inl && io->c_mask && io->quat && io->trans && io->order);
    YARG(io->T_o && io->o_best && io->o_inl && io->o_mask && io->c_iters >= 0 && io->o_iters >= 0);
    return YOHO_OK;
}

// Front half of a pair: PartI on both fragments (unless given), the mutual matching, and the match count on its way to a pinned
// host slot (asynchronous; the event marks its arrival).  Nothing here waits for the device.
extern "C" int yoho_register_pair_begin(yoho_ctx* ctx, const yoho_pair_io* io, void* stream) {
    if (int rc = pair_check(ctx, io)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    YCHECK(cudaSetDevice(ctx->device));
    if (ctx->pair_head - ctx->pair_tail >= yoho_ctx::kPairRing) {
        yoho_set_error("yoho_register_pair_begin: %d pairs already in flight (call yoho_register_pair_end)", yoho_ctx::kPairRing);
        return YOHO_ERR_ARG;
    }
    if (!ctx->pair_M_pinned) {
        YCHECK(cudaHostAlloc((void**)&ctx->pair_M_pinned, sizeof(int32_t) * yoho_ctx::kPairRing, cudaHostAllocDefault));
        for (int i = 0; i < yoho_ctx::kPairRing; ++i) YCHECK(cudaEventCreateWithFlags(&ctx->pair_ev[i], cudaEventDisableTiming));
    }
    int rc;
    if (!io->have_part1) {
        if ((rc = yoho_part1_forward(ctx, io->featA, io->Ka, io->eqvA, nullptr, io->descA, stream))) return rc;
        if ((rc = yoho_part1_forward(ctx, io->featB, io->Kb, io->eqvB, nullptr, io->descB, stream))) return rc;
    }
    const int slot = ctx->pair_head % yoho_ctx::kPairRing;
    if (io->Ka > 0 && io->Kb > 0) {
        if ((rc = yoho_mutual_nn(ctx, io->descA, io->Ka, io->descB, io->Kb, io->pairs, io->n_pairs, nullptr, nullptr, stream))) return rc;
    } else {
        YCHECK(cudaMemsetAsync(io->n_pairs, 0, sizeof(int), st));
    }
    YCHECK(cudaMemcpyAsync(ctx->pair_M_pinned + slot, io->n_pairs, sizeof(int), cudaMemcpyDeviceToHost, st));
    YCHECK(cudaEventRecord(ctx->pair_ev[slot], st));
    ctx->pair_head++;
    return YOHO_OK;
}

// Back half of the OLDEST begun pair (`io` must be the structure passed to its begin call): waits for that pair's match count —
// the one host synchronisation of a pair; when other pairs were begun in between it has long arrived and the device never idles —
// then queues rotation index, YOHO-C, PartII and YOHO-O.
extern "C" int yoho_register_pair_end(yoho_ctx* ctx, const yoho_pair_io* io, int32_t* M_host, void* stream) {
    if (int rc = pair_check(ctx, io)) return rc;
    YARG(M_host);
    cudaStream_t st = (cudaStream_t)stream;
    YCHECK(cudaSetDevice(ctx->device));
    if (ctx->pair_head == ctx->pair_tail) {
        yoho_set_error("yoho_register_pair_end without a matching yoho_register_pair_begin");
        return YOHO_ERR_ARG;
    }
    const int slot = ctx->pair_tail % yoho_ctx::kPairRing;
    ctx->pair_tail++;
    YCHECK(cudaEventSynchronize(ctx->pair_ev[slot]));
    const int M = ctx->pair_M_pinned[slot];
    int rc;
    *M_host = M;
    if (M == 0) {
        identity_kernel<<<1, 32, 0, st>>>(io->T_c, io->T_o, io->c_best, io->o_best, io->c_inl, io->o_inl);
        ctx->launches++;
        YCHECK(cudaGetLastError());
        return YOHO_OK;
    }
    // Batch_Des2R_torch(feats1[m1], feats0[m0]) (tests/extractor.py:97-99): des1 = fragment 1 rows (column 1 of the match list)
    if ((rc = yoho_rot_argmax(ctx, io->eqvB, io->pairs + 1, io->eqvA, io->pairs, 2, M, io->dr_index, nullptr, stream))) return rc;
    if ((rc = yoho_gather_kps(ctx, io->kpsA, io->kpsB, io->pairs, M, io->k0, io->k1, stream))) return rc;
    if ((rc = yoho_c_draw(ctx, io->dr_index, M, io->c_iters, io->seed, io->hyp, io->c_status, stream))) return rc;
    if ((rc = yoho_c_ransac(ctx, io->k0, io->k1, M, io->hyp, nullptr, nullptr, io->c_iters, io->c_dist, io->T_c, io->c_best, io->c_inl,
                            io->c_mask, nullptr, stream))) return rc;
    c_finish_kernel<<<1, 32, 0, st>>>(io->c_status, io->T_c, io->c_best);
    ctx->launches++;
    if ((rc = yoho_part2_forward(ctx, io->featA, io->featB, io->eqvA, io->eqvB, io->pairs, io->dr_index, M, io->kpsA, io->kpsB,
                                 io->quat, io->trans, stream))) return rc;
    if ((rc = yoho_o_order(ctx, M, io->seed, io->order, stream))) return rc;
    const int H = M < io->o_iters ? M : io->o_iters;
    if ((rc = yoho_o_score(ctx, io->k0, io->k1, M, io->trans, io->order, H, io->o_dist, io->T_o, io->o_best, io->o_inl, io->o_mask,
                           nullptr, stream))) return rc;
    YCHECK(cudaGetLastError());
    return YOHO_OK;
}

extern "C" int yoho_register_pair(yoho_ctx* ctx, const yoho_pair_io* io, int32_t* M_host, void* stream) {
    YARG(M_host);
    if (ctx && ctx->pair_head != ctx->pair_tail) {
        yoho_set_error("yoho_register_pair while split-phase pairs are in flight");
        return YOHO_ERR_ARG;
    }
    if (int rc = yoho_register_pair_begin(ctx, io, stream)) return rc;
    return yoho_register_pair_end(ctx, io, M_host, stream);
}
