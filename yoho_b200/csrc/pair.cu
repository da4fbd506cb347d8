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

extern "C" int yoho_register_pair(yoho_ctx* ctx, const yoho_pair_io* io, int32_t* M_host, void* stream) {
    YARG(ctx && io && M_host && io->featA && io->featB && io->kpsA && io->kpsB && io->Ka >= 0 && io->Kb >= 0);
    YARG(io->eqvA && io->eqvB && io->descA && io->descB && io->pairs && io->n_pairs && io->dr_index && io->k0 && io->k1);
    YARG(io->hyp && io->c_status && io->T_c && io->c_best && io->c_inl && io->c_mask && io->quat && io->trans && io->order);
    YARG(io->T_o && io->o_best && io->o_inl && io->o_mask && io->c_iters >= 0 && io->o_iters >= 0);
    cudaStream_t st = (cudaStream_t)stream;
    YCHECK(cudaSetDevice(ctx->device));
    int rc;
    if (!io->have_part1) {
        if ((rc = yoho_part1_forward(ctx, io->featA, io->Ka, io->eqvA, nullptr, io->descA, stream))) return rc;
        if ((rc = yoho_part1_forward(ctx, io->featB, io->Kb, io->eqvB, nullptr, io->descB, stream))) return rc;
    }
    int M = 0;
    if (io->Ka > 0 && io->Kb > 0) {
        if ((rc = yoho_mutual_nn(ctx, io->descA, io->Ka, io->descB, io->Kb, io->pairs, io->n_pairs, nullptr, nullptr, stream))) return rc;
        YCHECK(cudaMemcpyAsync(&M, io->n_pairs, sizeof(int), cudaMemcpyDeviceToHost, st));
        YCHECK(cudaStreamSynchronize(st));                       // the one host synchronisation of the pair
    } else {
        YCHECK(cudaMemsetAsync(io->n_pairs, 0, sizeof(int), st));
    }
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
