// Context, group tables, checkpoint packing and error plumbing of the yoho_b200 C ABI (include/yoho_b200.h).
#include <stdarg.h>
#include <math.h>
#include <vector>
#include "common.cuh"

static thread_local char g_err[1024] = "";

void yoho_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* yoho_last_error(void) { return g_err; }
extern "C" int yoho_abi_version(void) { return YOHO_ABI_VERSION; }

int yoho_ws_reserve(yoho_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return YOHO_OK;
    size_t want = bytes + (bytes >> 3);
    if (ctx->ws) {
        YCHECK(cudaDeviceSynchronize());
        YCHECK(cudaFree(ctx->ws));
        ctx->ws = nullptr;
        ctx->ws_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->ws, want);
    if (e != cudaSuccess) {
        yoho_set_error("workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
        return YOHO_ERR_ALLOC;
    }
    ctx->ws_bytes = want;
    return YOHO_OK;
}

template <typename T>
static int upload(T** dst, const std::vector<T>& v) {
    YCHECK(cudaMalloc((void**)dst, v.size() * sizeof(T)));
    YCHECK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return YOHO_OK;
}

extern "C" int yoho_ctx_create(int device, const double* rotation_host, const int32_t* perm_host, const int32_t* nei_host,
                               yoho_ctx** out) {
    YARG(rotation_host && perm_host && nei_host && out);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        yoho_set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
        return YOHO_ERR_CUDA;
    }
    YARG(device >= 0 && device < ndev);
    YCHECK(cudaSetDevice(device));
    yoho_ctx* c = new yoho_ctx();
    c->device = device;
    cudaDeviceProp prop;
    YCHECK(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    for (int i = 0; i < YG * YG; ++i) YARG(perm_host[i] >= 0 && perm_host[i] < YG);
    for (int i = 0; i < YG * YT; ++i) YARG(nei_host[i] >= 0 && nei_host[i] < YG);

    std::vector<double> rot(rotation_host, rotation_host + YG * 9);
    std::vector<float> rot32(YG * 9);
    for (int i = 0; i < YG * 9; ++i) rot32[i] = (float)rot[i];
    std::vector<uint8_t> perm(YG * YG), perm_t(YG * YG);
    for (int a = 0; a < YG; ++a)
        for (int g = 0; g < YG; ++g) {
            perm[a * YG + g] = (uint8_t)perm_host[a * YG + g];
            perm_t[g * YG + a] = (uint8_t)perm_host[a * YG + g];
        }
    std::vector<int> idx_full(nei_host, nei_host + YG * YT);
    // receptive field of g = 0 (PartII): hop1 = N[0], hop2 = ordered union of N[e], e in hop1
    std::vector<int> hop1(YT), hop2;
    for (int k = 0; k < YT; ++k) hop1[k] = nei_host[k];
    for (int e1 : hop1)
        for (int k = 0; k < YT; ++k) {
            int v = nei_host[e1 * YT + k];
            bool seen = false;
            for (int h : hop2) seen |= (h == v);
            if (!seen) hop2.push_back(v);
        }
    YARG(hop2.size() == 45);
    std::vector<int> pos(YG, -1);
    for (size_t i = 0; i < hop2.size(); ++i) pos[hop2[i]] = (int)i;
    std::vector<int> idx_init(45 * YT), idx_a(YT * YT), idx_b(YT), idx_one(1, 0);
    for (int j = 0; j < 45; ++j)
        for (int k = 0; k < YT; ++k) idx_init[j * YT + k] = nei_host[hop2[j] * YT + k];
    for (int j = 0; j < YT; ++j)
        for (int k = 0; k < YT; ++k) {
            int v = pos[nei_host[hop1[j] * YT + k]];
            YARG(v >= 0);
            idx_a[j * YT + k] = v;
        }
    for (int k = 0; k < YT; ++k) idx_b[k] = k;   // N[0][k] = hop1[k]
    c->hop2_zero_pos = pos[0];
    YARG(c->hop2_zero_pos >= 0);

    int rc = 0;
    rc |= upload(&c->d_rot, rot);
    rc |= upload(&c->d_rot32, rot32);
    rc |= upload(&c->d_perm, perm);
    rc |= upload(&c->d_perm_t, perm_t);
    rc |= upload(&c->d_idx_full, idx_full);
    rc |= upload(&c->d_idx_p2_init, idx_init);
    rc |= upload(&c->d_idx_p2_a, idx_a);
    rc |= upload(&c->d_idx_p2_b, idx_b);
    rc |= upload(&c->d_idx_one, idx_one);
    std::vector<int> idx_ident(YG);
    for (int g = 0; g < YG; ++g) idx_ident[g] = g;
    rc |= upload(&c->d_idx_ident, idx_ident);
    if (rc) { delete c; return YOHO_ERR_CUDA; }
    if ((rc = yoho_ws_reserve(c, (size_t)64 << 20))) { delete c; return rc; }
    *out = c;
    return YOHO_OK;
}

static void free_layer(GLayer& L) {
    cudaFree(L.w); cudaFree(L.bias); cudaFree(L.w_hi); cudaFree(L.w_lo);
    L = GLayer();
}
static void free_bn(GBn& b) {
    cudaFree(b.scale); cudaFree(b.shift);
    b = GBn();
}

extern "C" int yoho_ctx_destroy(yoho_ctx* c) {
    if (!c) return YOHO_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    GLayer* layers[] = {&c->p1_in, &c->p1_a, &c->p1_b, &c->p1_out, &c->p1_out_cat, &c->p2_init, &c->p2_a, &c->p2_b, &c->p2_fc1, &c->p2_fc2, &c->p2_fc3, &c->p2_fc2_pad,
                        &c->p2_b_split[0], &c->p2_b_split[1], &c->p2_b_split[2], &c->p2_b_split[3], &c->p2_b_split[4]};
    for (GLayer* l : layers) free_layer(*l);
    GBn* bns[] = {&c->p1_bn_a, &c->p1_bn_b, &c->p1_bn_out, &c->p2_bn_init, &c->p2_bn_a, &c->p2_bn_b, &c->p2_bn1, &c->p2_bn2, &c->p2_bn2_pad};
    for (GBn* b : bns) free_bn(*b);
    if (c->pair_M_pinned) {
        cudaFreeHost(c->pair_M_pinned);
        for (int i = 0; i < yoho_ctx::kPairRing; ++i) cudaEventDestroy(c->pair_ev[i]);
    }
    cudaFree(c->d_rot); cudaFree(c->d_rot32); cudaFree(c->d_perm); cudaFree(c->d_perm_t);
    cudaFree(c->d_idx_full_inv);
    cudaFree(c->d_idx_full); cudaFree(c->d_idx_p2_init); cudaFree(c->d_idx_p2_a); cudaFree(c->d_idx_p2_b); cudaFree(c->d_idx_one); cudaFree(c->d_idx_ident);
    for (int r = 0; r < 8; ++r) {
        free_layer(c->p1f_a[r]); free_layer(c->p1f_b[r]); free_layer(c->p1f_in[r]); free_layer(c->p1f_out[r]);
        cudaFree(c->d_fidx[r]); cudaFree(c->d_fomap[r]); cudaFree(c->d_fomap_out[r]);
    }
    cudaFree(c->d_F); cudaFree(c->d_p1_bias31);
    cudaFree(c->d_fwd_hi); cudaFree(c->d_fwd_lo); cudaFree(c->d_inv_hi); cudaFree(c->d_inv_lo);
    cudaFree(c->ws);
    delete c;
    return YOHO_OK;
}

int gconv_tc_pack(yoho_ctx* ctx, GLayer& L, const std::vector<float>& w_kco);   // gconv_tc.cu

// W[o][c][0][k] (reference layout) -> W_k[c][o]
static int pack_conv(yoho_ctx* ctx, GLayer& L, const yoho_conv_host& h, int cin, int cout, int taps) {
    YARG(h.weight_host && h.bias_host);
    free_layer(L);
    L.cin = cin; L.cout = cout; L.taps = taps;
    std::vector<float> w((size_t)taps * cin * cout);
    for (int o = 0; o < cout; ++o)
        for (int c = 0; c < cin; ++c)
            for (int k = 0; k < taps; ++k) w[((size_t)k * cin + c) * cout + o] = h.weight_host[((size_t)o * cin + c) * taps + k];
    std::vector<float> b(h.bias_host, h.bias_host + cout);
    if (int rc = upload(&L.w, w)) return rc;
    if (int rc = upload(&L.bias, b)) return rc;
    if (taps == YT) return gconv_tc_pack(ctx, L, w);
    return YOHO_OK;
}

// eval-mode BatchNorm2d, eps = 1e-5: y = (x - mean) / sqrt(var + eps) * weight + bias = x*scale + shift
static int pack_bn(GBn& B, const yoho_bn_host& h, int c) {
    YARG(h.weight_host && h.bias_host && h.running_mean_host && h.running_var_host);
    free_bn(B);
    B.c = c;
    std::vector<float> sc(c), sh(c);
    for (int i = 0; i < c; ++i) {
        const double s = (double)h.weight_host[i] / sqrt((double)h.running_var_host[i] + 1e-5);
        sc[i] = (float)s;
        sh[i] = (float)((double)h.bias_host[i] - (double)h.running_mean_host[i] * s);
    }
    if (int rc = upload(&B.scale, sc)) return rc;
    return upload(&B.shift, sh);
}

extern "C" int yoho_part1_load(yoho_ctx* ctx, const yoho_part1_weights* w) {
    YARG(ctx && w);
    YCHECK(cudaSetDevice(ctx->device));
    ctx->has_p1 = false;
    int rc = 0;
    if ((rc = pack_conv(ctx, ctx->p1_in, w->conv_in, 32, 256, YT))) return rc;
    if ((rc = pack_bn(ctx->p1_bn_a, w->bn_a, 256))) return rc;
    if ((rc = pack_conv(ctx, ctx->p1_a, w->conv_a, 256, 512, YT))) return rc;
    if ((rc = pack_bn(ctx->p1_bn_b, w->bn_b, 512))) return rc;
    if ((rc = pack_conv(ctx, ctx->p1_b, w->conv_b, 512, 256, YT))) return rc;
    if ((rc = pack_bn(ctx->p1_bn_out, w->bn_out, 256))) return rc;
    if ((rc = pack_conv(ctx, ctx->p1_out, w->conv_out, 256, 32, YT))) return rc;
    {   // layer 4 with the gather on the OUTPUT side: Z = a3 . W_cat (dense), y4[g] = sum_k Z[N[g][k]][k*32 : k*32+32]
        GLayer& L = ctx->p1_out_cat;
        free_layer(L);
        L.cin = 256; L.cout = 512; L.taps = 1; L.tc_dense = 1; L.prof_class = 3;
        std::vector<float> wc((size_t)256 * 512, 0.f), zb(512, 0.f);
        for (int k = 0; k < YT; ++k)
            for (int c = 0; c < 256; ++c)
                for (int o = 0; o < 32; ++o) wc[(size_t)c * 512 + k * 32 + o] = w->conv_out.weight_host[((size_t)o * 256 + c) * YT + k];
        if ((rc = upload(&L.bias, zb))) return rc;
        if ((rc = gconv_tc_pack(ctx, L, wc))) return rc;
    }
    ctx->p1_in.prof_class = 0; ctx->p1_a.prof_class = 1; ctx->p1_b.prof_class = 2; ctx->p1_out.prof_class = 3;
    {   // bias of the residual block's output plus the bias of its identity shortcut (all-Fourier path: the shortcut travels as
        // Fourier coefficients without its bias)
        std::vector<float> b31(256);
        for (int o = 0; o < 256; ++o) b31[o] = w->conv_b.bias_host[o] + w->conv_in.bias_host[o];
        cudaFree(ctx->d_p1_bias31); ctx->d_p1_bias31 = nullptr;
        if ((rc = upload(&ctx->d_p1_bias31, b31))) return rc;
    }
    ctx->has_p1 = true;
    return YOHO_OK;
}

extern "C" int yoho_part1_load_fourier(yoho_ctx* ctx, const float* F_host, int n_irreps, const yoho_fourier_irrep* irreps) {
    YARG(ctx && F_host && irreps && n_irreps >= 1 && n_irreps <= 8);
    YCHECK(cudaSetDevice(ctx->device));
    ctx->has_p1f = false;
    ctx->has_p1f_io = false;
    for (int r = 0; r < 8; ++r) {
        free_layer(ctx->p1f_a[r]); free_layer(ctx->p1f_b[r]); free_layer(ctx->p1f_in[r]); free_layer(ctx->p1f_out[r]);
        cudaFree(ctx->d_fidx[r]); cudaFree(ctx->d_fomap[r]); cudaFree(ctx->d_fomap_out[r]);
        ctx->d_fidx[r] = ctx->d_fomap[r] = ctx->d_fomap_out[r] = nullptr;
    }
    cudaFree(ctx->d_F);
    ctx->d_F = nullptr;
    {
        std::vector<float> Fv(F_host, F_host + YG * YG);
        if (int rc0 = upload(&ctx->d_F, Fv)) return rc0;
    }
    int rc = 0;
    {   // bf16 hi/lo copies with the OUTPUT index as the row: forward[m][g] = F[m][g], inverse[g][m] = F[m][g]
        auto f2bf = [](float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7FFFu + ((u >> 16) & 1u); return (unsigned short)(u >> 16); };
        auto bf2f = [](unsigned short h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; };
        std::vector<unsigned short> fh(64 * 64, 0), fl(64 * 64, 0), ih(64 * 64, 0), il(64 * 64, 0);
        for (int m = 0; m < YG; ++m)
            for (int g = 0; g < YG; ++g) {
                const float v = F_host[m * YG + g];
                const unsigned short h = f2bf(v), l = f2bf(v - bf2f(h));
                fh[m * 64 + g] = h; fl[m * 64 + g] = l;
                ih[g * 64 + m] = h; il[g * 64 + m] = l;
            }
        cudaFree(ctx->d_fwd_hi); cudaFree(ctx->d_fwd_lo); cudaFree(ctx->d_inv_hi); cudaFree(ctx->d_inv_lo);
        if ((rc = upload((unsigned short**)&ctx->d_fwd_hi, fh))) return rc;
        if ((rc = upload((unsigned short**)&ctx->d_fwd_lo, fl))) return rc;
        if ((rc = upload((unsigned short**)&ctx->d_inv_hi, ih))) return rc;
        if ((rc = upload((unsigned short**)&ctx->d_inv_lo, il))) return rc;
    }
    int total = 0, n_io = 0;
    for (int r = 0; r < n_irreps; ++r) {
        const yoho_fourier_irrep& ir = irreps[r];
        YARG(ir.d >= 1 && ir.d <= 5 && ir.w_a_host && ir.w_b_host && ir.idx_host && ir.omap_host);
        YARG((ir.w_in_host == nullptr) == (ir.w_out_host == nullptr));
        if (ir.w_in_host) {
            ++n_io;
            // layer 1: 32 -> d*256, d taps
            GLayer& Li = ctx->p1f_in[r];
            Li.cin = 32; Li.cout = ir.d * 256; Li.taps = ir.d; Li.tc_dense = 1; Li.prof_class = 0;
            std::vector<float> wi(ir.w_in_host, ir.w_in_host + (size_t)ir.d * 32 * Li.cout), zbi(Li.cout, 0.f);
            if ((rc = upload(&Li.bias, zbi))) return rc;
            if ((rc = gconv_tc_pack(ctx, Li, wi))) return rc;
            YARG(Li.w_hi && Li.w_lo);
            // layer 4: 256 -> d*32, d taps, columns zero-padded to one 256-wide tile; output-row table [j][8 column groups]
            GLayer& Lo = ctx->p1f_out[r];
            Lo.cin = 256; Lo.cout = 256; Lo.taps = ir.d; Lo.tc_dense = 1; Lo.prof_class = 3;
            const int nv = ir.d * 32;
            std::vector<float> wo((size_t)ir.d * 256 * 256, 0.f), zbo(256, 0.f);
            for (int l = 0; l < ir.d; ++l)
                for (int c = 0; c < 256; ++c)
                    memcpy(&wo[((size_t)l * 256 + c) * 256], ir.w_out_host + ((size_t)l * 256 + c) * nv, nv * sizeof(float));
            if ((rc = upload(&Lo.bias, zbo))) return rc;
            if ((rc = gconv_tc_pack(ctx, Lo, wo))) return rc;
            YARG(Lo.w_hi && Lo.w_lo);
            std::vector<int> om8(ir.d * 8, 0);
            for (int j = 0; j < ir.d; ++j)
                for (int i = 0; i < ir.d; ++i) om8[j * 8 + i] = ir.omap_host[j * ir.d + i];
            if ((rc = upload(&ctx->d_fomap_out[r], om8))) return rc;
        }
        total += ir.d * ir.d;
        ctx->fd[r] = ir.d;
        std::vector<int> idx(ir.idx_host, ir.idx_host + ir.d * ir.d), om(ir.omap_host, ir.omap_host + ir.d * ir.d);
        if ((rc = upload(&ctx->d_fidx[r], idx))) return rc;
        if ((rc = upload(&ctx->d_fomap[r], om))) return rc;
        struct { GLayer* L; const float* w; int cin, o; int cls; } two[2] = {{&ctx->p1f_a[r], ir.w_a_host, 256, 512, 1},
                                                                              {&ctx->p1f_b[r], ir.w_b_host, 512, 256, 2}};
        for (auto& e : two) {
            GLayer& L = *e.L;
            L.cin = e.cin; L.cout = ir.d * e.o; L.taps = ir.d; L.tc_dense = 1; L.prof_class = e.cls;
            std::vector<float> w(e.w, e.w + (size_t)ir.d * e.cin * L.cout), zb(L.cout, 0.f);
            if ((rc = upload(&L.bias, zb))) return rc;
            if ((rc = gconv_tc_pack(ctx, L, w))) return rc;
            YARG(L.w_hi && L.w_lo);
        }
    }
    YARG(total == YG && (n_io == 0 || n_io == n_irreps));
    ctx->nf = n_irreps;
    ctx->has_p1f = true;
    ctx->has_p1f_io = n_io == n_irreps;
    return YOHO_OK;
}

extern "C" int yoho_part2_load(yoho_ctx* ctx, const yoho_part2_weights* w) {
    YARG(ctx && w);
    YCHECK(cudaSetDevice(ctx->device));
    ctx->has_p2 = false;
    int rc = 0;
    if ((rc = pack_bn(ctx->p2_bn_init, w->bn_init, 128))) return rc;
    if ((rc = pack_conv(ctx, ctx->p2_init, w->conv_init, 128, 256, YT))) return rc;
    if ((rc = pack_bn(ctx->p2_bn_a, w->bn_a, 256))) return rc;
    if ((rc = pack_conv(ctx, ctx->p2_a, w->conv_a, 256, 512, YT))) return rc;
    if ((rc = pack_bn(ctx->p2_bn_b, w->bn_b, 512))) return rc;
    if ((rc = pack_conv(ctx, ctx->p2_b, w->conv_b, 512, 256, YT))) return rc;
    {   // the last group convolution is evaluated at g = 0 only: M rows instead of 60 M.  Cut along the taps so that one launch
        // fills the machine; the five partial sums are added in a fixed order by part2_reduce_kernel.
        const int cuts[6] = {0, 3, 6, 9, 11, 13};
        std::vector<float> wf((size_t)YT * 512 * 256);
        for (int o = 0; o < 256; ++o)
            for (int c = 0; c < 512; ++c)
                for (int k = 0; k < YT; ++k) wf[((size_t)k * 512 + c) * 256 + o] = w->conv_b.weight_host[((size_t)o * 512 + c) * YT + k];
        for (int q = 0; q < 5; ++q) {
            GLayer& L = ctx->p2_b_split[q];
            free_layer(L);
            L.cin = 512; L.cout = 256; L.taps = cuts[q + 1] - cuts[q]; L.tc_dense = 1; L.prof_class = 6;
            std::vector<float> wq(wf.begin() + (size_t)cuts[q] * 512 * 256, wf.begin() + (size_t)cuts[q + 1] * 512 * 256), zb(256, 0.f);
            if ((rc = upload(&L.bias, zb))) return rc;
            if ((rc = gconv_tc_pack(ctx, L, wq))) return rc;
        }
    }
    if ((rc = pack_conv(ctx, ctx->p2_fc1, w->fc1, 256, 512, 1))) return rc;
    if ((rc = pack_bn(ctx->p2_bn1, w->bn1, 512))) return rc;
    if ((rc = pack_conv(ctx, ctx->p2_fc2, w->fc2, 512, 128, 1))) return rc;
    if ((rc = pack_bn(ctx->p2_bn2, w->bn2, 128))) return rc;
    {   // the two hidden 1x1 layers of PartII_To_R_FC as dense tensor-core GEMMs (one tap): 256 -> 512, and 512 -> 128 zero-padded
        // to one 256-column tile (n_valid = 128 at the call)
        std::vector<float> w1((size_t)256 * 512);
        for (int o = 0; o < 512; ++o)
            for (int c = 0; c < 256; ++c) w1[(size_t)c * 512 + o] = w->fc1.weight_host[(size_t)o * 256 + c];
        ctx->p2_fc1.tc_dense = 1;
        if ((rc = gconv_tc_pack(ctx, ctx->p2_fc1, w1))) return rc;
        GLayer& L = ctx->p2_fc2_pad;
        free_layer(L);
        L.cin = 512; L.cout = 256; L.taps = 1; L.tc_dense = 1; L.prof_class = 7;
        std::vector<float> w2((size_t)512 * 256, 0.f), b2(256, 0.f), sc(256, 0.f), sh(256, 0.f);
        for (int o = 0; o < 128; ++o) {
            for (int c = 0; c < 512; ++c) w2[(size_t)c * 256 + o] = w->fc2.weight_host[(size_t)o * 512 + c];
            b2[o] = w->fc2.bias_host[o];
            const double sdv = (double)w->bn2.weight_host[o] / sqrt((double)w->bn2.running_var_host[o] + 1e-5);
            sc[o] = (float)sdv;
            sh[o] = (float)((double)w->bn2.bias_host[o] - (double)w->bn2.running_mean_host[o] * sdv);
        }
        if ((rc = upload(&L.bias, b2))) return rc;
        if ((rc = gconv_tc_pack(ctx, L, w2))) return rc;
        free_bn(ctx->p2_bn2_pad);
        ctx->p2_bn2_pad.c = 256;
        if ((rc = upload(&ctx->p2_bn2_pad.scale, sc))) return rc;
        if ((rc = upload(&ctx->p2_bn2_pad.shift, sh))) return rc;
    }
    if ((rc = pack_conv(ctx, ctx->p2_fc3, w->fc3, 128, 4, 1))) return rc;   // packed [128][4]
    ctx->p2_init.prof_class = 4; ctx->p2_a.prof_class = 5; ctx->p2_b.prof_class = 6;
    ctx->p2_fc1.prof_class = 7; ctx->p2_fc2.prof_class = 7; ctx->p2_fc3.prof_class = 7;
    ctx->has_p2 = true;
    return YOHO_OK;
}

extern "C" int yoho_set_gconv_impl(yoho_ctx* ctx, int impl) {
    YARG(ctx && impl >= 0 && impl <= 3);
    ctx->gconv_impl = impl;
    return YOHO_OK;
}

extern "C" int yoho_set_tuning(yoho_ctx* ctx, int key, int value) {
    YARG(ctx && (key == 0 || key == 1));
    if (key == 0) ctx->tc_flags = value;
    else { YARG(value >= 0); ctx->split_min_k = value; }
    return YOHO_OK;
}

extern "C" int64_t yoho_launch_count(const yoho_ctx* ctx) { return ctx ? ctx->launches : 0; }

int gconv_simt_forward(yoho_ctx* ctx, const GLayer& L, const GConvArgs& a, cudaStream_t st);
int gconv_tc_forward(yoho_ctx* ctx, const GLayer& L, const GConvArgs& a, cudaStream_t st);
bool gconv_tc_eligible(const GLayer& L, const GConvArgs& a);

static cudaEvent_t prof_event(yoho_ctx* ctx) {
    cudaEvent_t e = nullptr;
    if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

void yoho_prof_begin(yoho_ctx* ctx, int cls, double flops, cudaStream_t st) {
    if (!ctx->prof_on) return;
    yoho_ctx::ProfRec rec{};
    rec.a = prof_event(ctx); rec.b = prof_event(ctx); rec.cls = cls; rec.flops = flops;
    cudaEventRecord(rec.a, st);
    ctx->prof.push_back(rec);
}
void yoho_prof_end(yoho_ctx* ctx, cudaStream_t st) {
    if (!ctx->prof_on || ctx->prof.empty()) return;
    cudaEventRecord(ctx->prof.back().b, st);
}

int gconv_forward(yoho_ctx* ctx, const GLayer& L, const GConvArgs& a, cudaStream_t st) {
    yoho_ctx::ProfRec rec{};
    const bool prof = ctx->prof_on && a.B > 0;
    if (prof) {
        rec.a = prof_event(ctx); rec.b = prof_event(ctx); rec.cls = L.prof_class;
        rec.flops = 2.0 * (double)a.B * a.Jout * L.taps * L.cin * (a.n_valid > 0 ? a.n_valid : L.cout);   // algorithmic
        cudaEventRecord(rec.a, st);
    }
    int rc;
    if (ctx->gconv_impl >= 1 && gconv_tc_eligible(L, a)) rc = gconv_tc_forward(ctx, L, a, st);
    else if (!a.act) { yoho_set_error("group convolution: FP32 activations missing for the SIMT path"); rc = YOHO_ERR_ARG; }
    else rc = gconv_simt_forward(ctx, L, a, st);
    if (prof) { cudaEventRecord(rec.b, st); ctx->prof.push_back(rec); }
    return rc;
}

int gconv_tc_forward_grouped(yoho_ctx* ctx, const GLayer* const* Ls, const GConvArgs* as, int n, cudaStream_t st);

// All per-irrep GEMMs of a group-Fourier layer in one launch (falls back to one launch per irrep for the SIMT twin).
int gconv_forward_grouped(yoho_ctx* ctx, const GLayer* const* Ls, const GConvArgs* as, int n, cudaStream_t st) {
    double flops = 0;
    for (int g = 0; g < n; ++g) flops += 2.0 * (double)as[g].B * as[g].Jout * Ls[g]->taps * Ls[g]->cin * Ls[g]->cout;
    yoho_prof_begin(ctx, Ls[0]->prof_class, flops, st);
    const int rc = gconv_tc_forward_grouped(ctx, Ls, as, n, st);
    yoho_prof_end(ctx, st);
    return rc;
}

extern "C" int yoho_profile_enable(yoho_ctx* ctx, int enable) {
    YARG(ctx);
    YCHECK(cudaSetDevice(ctx->device));
    for (auto& r : ctx->prof) { ctx->prof_pool.push_back(r.a); ctx->prof_pool.push_back(r.b); }
    ctx->prof.clear();
    ctx->prof_on = enable != 0;
    return YOHO_OK;
}

extern "C" int yoho_profile_read(yoho_ctx* ctx, double* ms_host, int64_t* launches_host, double* flops_host) {
    YARG(ctx && ms_host && launches_host && flops_host);
    YCHECK(cudaSetDevice(ctx->device));
    YCHECK(cudaDeviceSynchronize());
    for (int c = 0; c < YOHO_PROF_CLASSES; ++c) { ms_host[c] = 0; launches_host[c] = 0; flops_host[c] = 0; }
    for (auto& r : ctx->prof) {
        float ms = 0.f;
        YCHECK(cudaEventElapsedTime(&ms, r.a, r.b));
        if (r.cls >= 0 && r.cls < YOHO_PROF_CLASSES) { ms_host[r.cls] += ms; launches_host[r.cls] += 1; flops_host[r.cls] += r.flops; }
    }
    return YOHO_OK;
}

int gconv_split_bf16(yoho_ctx* ctx, const float* x, void* hi, void* lo, size_t n, cudaStream_t st);

extern "C" int yoho_debug_layer(yoho_ctx* ctx, int layer, int impl, const float* act, int B, float* out_raw, void* stream) {
    YARG(ctx && act && out_raw && B > 0 && layer >= 0 && layer <= 6 && impl >= 0 && impl <= 2);
    YARG(layer < 4 ? ctx->has_p1 : ctx->has_p2);
    YCHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const GLayer* Ls[] = {&ctx->p1_in, &ctx->p1_a, &ctx->p1_b, &ctx->p1_out, &ctx->p2_init, &ctx->p2_a, &ctx->p2_b};
    const GLayer& L = *Ls[layer];
    GConvArgs a{};
    a.idx = ctx->d_idx_full; a.B = B; a.Jin = YG; a.Jout = YG; a.act = act; a.out_raw = out_raw;
    const int saved = ctx->gconv_impl;
    int rc;
    if (impl >= 1) {
        const size_t n = (size_t)B * YG * L.cin;
        if ((rc = yoho_ws_reserve(ctx, n * 4))) return rc;
        void* hi = ctx->ws;
        void* lo = (char*)ctx->ws + n * 2;
        if ((rc = gconv_split_bf16(ctx, act, hi, lo, n, st))) return rc;
        a.act_hi = hi; a.act_lo = lo;
        YARG(gconv_tc_eligible(L, a));
        ctx->gconv_impl = impl;
    } else {
        ctx->gconv_impl = 0;
    }
    rc = gconv_forward(ctx, L, a, st);
    ctx->gconv_impl = saved;
    return rc;
}
