"""Device engine: one `Engine` per GPU wraps a `yoho_ctx` of the C ABI and exposes every stage of the hot
path on torch CUDA tensors (torch is used for device memory and streams only — the arithmetic is in
libyoho_b200.so).  All methods enqueue on the current torch stream and do not synchronise unless they
return host values.
"""
import ctypes
import os
import threading
import numpy as np
import torch

from . import _lib
from . import group as _group


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class PairBuffers:
    """Outputs of one `yoho_register_pair` call: one device allocation + a name -> (offset, bytes, shape, dtype) table.
    `out[name]` builds the tensor view on demand; `out.M` is the match count."""

    def __init__(self, buf, ent, ext, M, keep):
        self.buf, self.ent, self.ext, self.M, self._keep = buf, ent, ext, M, keep

    def __getitem__(self, name):
        if name in self.ext:
            return self.ext[name]
        if name == "T_c":
            return self["T_co"][0]
        if name == "T_o":
            return self["T_co"][1]
        o, n, shape, dt = self.ent[name]
        return self.buf[o:o + n].view(dt).view(shape)


class Engine:
    """One C-ABI context on one GPU.  Contract (include/yoho_b200.h): a context owns ONE packed weight set per network and ONE
    grow-only workspace, and its calls are not re-entrant — use it from one thread and one CUDA stream at a time (every method
    enqueues on the current torch stream).  `get_engine` hands every caller on a device the same Engine, so two torch modules
    holding different checkpoints share the weight slot: `network.PartI_test / PartII_test` re-upload their own weights
    whenever `weights_owner` says another caller loaded the slot in between."""

    def __init__(self, device=None, so3_dir=None):
        if not torch.cuda.is_available():
            raise _lib.YohoError("yoho_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")
        self.lib = _lib.load_library()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.tables = _group.load(so3_dir)
        rot = np.ascontiguousarray(self.tables.R, np.float64)
        perm = np.ascontiguousarray(self.tables.P, np.int32)
        nei = np.ascontiguousarray(self.tables.N, np.int32)
        h = ctypes.c_void_p()
        _lib.check(self.lib.yoho_ctx_create(self.device.index, rot.ctypes.data, perm.ctypes.data, nei.ctypes.data,
                                            ctypes.byref(h)))
        self.h = h
        self.has_part1 = False
        self.has_part2 = False
        # who uploaded the weights currently inside the context (several nn.Modules share one engine per device)
        self.weights_owner = {1: None, 2: None}
        self.impl_name = None
        # group-convolution implementation: tensor cores, PartI layers 2+3 in the group-Fourier domain by default;
        # YOHO_B200_GCONV=simt|tcgen05|tcgen05_split|tcgen05_fourier selects another one
        self.set_gconv_impl(os.environ.get("YOHO_B200_GCONV", "tcgen05_fourier"))

    def close(self):
        if getattr(self, "h", None):
            self.lib.yoho_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -----------------------------------------------------------------------------------
    def load_part1(self, state_dict, fourier=True, owner=None):
        w, keep = _lib.part1_struct(state_dict)
        _lib.check(self.lib.yoho_part1_load(self.h, ctypes.byref(w)))
        self.has_part1 = True
        self.weights_owner[1] = owner
        if fourier:
            self._load_part1_fourier(_lib._to_numpy_sd(state_dict))

    def _load_part1_fourier(self, sd):
        """Group-Fourier weights of the four PartI layers (yoho_b200/fourier.py) for implementation 'tcgen05_fourier'."""
        from . import fourier
        T = fourier.build(self.tables.dir if self.tables.dir != _group._PKG_DIR else None)
        blk = "PartI_net.SO3_Conv_layers.0."
        pa = fourier.pack_layer(sd[blk + "comb_layer_in.2.weight"], T)
        pb = fourier.pack_layer(sd[blk + "comb_layer_out.2.weight"], T)
        pi = fourier.pack_layer(sd["PartI_net.Conv_in.0.weight"], T)
        po = fourier.pack_layer(sd["PartI_net.Conv_out.comb_layer.2.weight"], T)
        n = len(pa)
        arr = (_lib.yoho_fourier_irrep * n)()
        keep = []
        for r in range(n):
            wa, wb = np.ascontiguousarray(pa[r]["w"]), np.ascontiguousarray(pb[r]["w"])
            idx, om = np.ascontiguousarray(pa[r]["idx"], np.int32), np.ascontiguousarray(pa[r]["omap"], np.int32)
            wi, wo = np.ascontiguousarray(pi[r]["w"]), np.ascontiguousarray(po[r]["w"])
            keep += [wa, wb, idx, om, wi, wo]
            arr[r].w_in_host = wi.ctypes.data_as(_lib._c_f)
            arr[r].w_out_host = wo.ctypes.data_as(_lib._c_f)
            arr[r].d, arr[r].off = pa[r]["d"], pa[r]["off"]
            arr[r].w_a_host = wa.ctypes.data_as(_lib._c_f)
            arr[r].w_b_host = wb.ctypes.data_as(_lib._c_f)
            arr[r].idx_host = idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
            arr[r].omap_host = om.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
        F = np.ascontiguousarray(T["F"], np.float32)
        _lib.check(self.lib.yoho_part1_load_fourier(self.h, F.ctypes.data, n, ctypes.cast(arr, ctypes.c_void_p)))

    def load_part2(self, state_dict, owner=None):
        w, keep = _lib.part2_struct(state_dict)
        _lib.check(self.lib.yoho_part2_load(self.h, ctypes.byref(w)))
        self.has_part2 = True
        self.weights_owner[2] = owner

    def set_gconv_impl(self, impl):
        self.impl_name = impl
        _lib.check(self.lib.yoho_set_gconv_impl(self.h, {"simt": 0, "tcgen05": 1, "tcgen05_split": 2, "tcgen05_fourier": 3}.get(impl, impl)))

    DEFAULT_TUNING = 3 | 256      # csrc/common.cuh yoho_ctx::tc_flags

    def set_tuning(self, key, value):
        _lib.check(self.lib.yoho_set_tuning(self.h, int(key), int(value)))

    def launch_count(self):
        return int(self.lib.yoho_launch_count(self.h))

    def debug_layer(self, layer, impl, act, cout):
        """Test hook: one group-convolution layer on FP32 activations [B,60,Cin] -> raw [B,60,cout]."""
        act = self._f32(act)
        B = act.shape[0]
        out = self._empty((B, 60, cout), torch.float32)
        _lib.check(self.lib.yoho_debug_layer(self.h, layer, {"simt": 0, "tcgen05": 1, "tcgen05_split": 2, "tcgen05_fourier": 3}[impl], _ptr(act), B, _ptr(out), _stream()))
        return out

    def profile(self, enable):
        _lib.check(self.lib.yoho_profile_enable(self.h, 1 if enable else 0))

    def profile_read(self):
        """-> list of dict(name, ms, launches, flops) per group-convolution layer class (device sync)."""
        n = _lib.PROF_CLASSES
        ms = np.zeros(n, np.float64); ln = np.zeros(n, np.int64); fl = np.zeros(n, np.float64)
        _lib.check(self.lib.yoho_profile_read(self.h, ms.ctypes.data, ln.ctypes.data, fl.ctypes.data))
        return [dict(name=_lib.PROF_NAMES[i], ms=float(ms[i]), launches=int(ln[i]), flops=float(fl[i])) for i in range(n)]

    # ---- helpers -----------------------------------------------------------------------------------
    def _f32(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        return x.to(device=self.device, dtype=torch.float32).contiguous()

    def _f64(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        return x.to(device=self.device, dtype=torch.float64).contiguous()

    def _i64(self, x):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int64))
        return x.to(device=self.device, dtype=torch.int64).contiguous()

    def _empty(self, shape, dtype):
        return torch.empty(shape, device=self.device, dtype=dtype)

    @staticmethod
    def _check_rows(pairs, K0, K1, what):
        """Row ids that arrive from the HOST (the reference's on-disk match lists) are range-checked like the reference's own
        fancy indexing would (IndexError); device-resident lists produced by this library are trusted."""
        if isinstance(pairs, np.ndarray) and pairs.size:
            p = pairs.reshape(-1, 2)
            if p.min() < 0 or p[:, 0].max() >= K0 or p[:, 1].max() >= K1:
                raise IndexError(f"{what}: match row out of range for fragments of {K0} / {K1} keypoints")

    @staticmethod
    def _check_bins(idx, what):
        if isinstance(idx, np.ndarray) and idx.size and (idx.min() < 0 or idx.max() >= 60):
            raise IndexError(f"{what}: rotation index outside [0, 60)")

    # ---- A: PartI ----------------------------------------------------------------------------------
    def part1(self, x, want_inv=True, want_desc=True):
        """x [B,32,60] -> dict(eqv [B,32,60], inv [B,32], desc [B,32])."""
        x = self._f32(x)
        assert x.dim() == 3 and x.shape[1] == 32 and x.shape[2] == 60, "group feature must be [B,32,60]"
        B = x.shape[0]
        eqv = self._empty((B, 32, 60), torch.float32)
        inv = self._empty((B, 32), torch.float32) if want_inv else None
        desc = self._empty((B, 32), torch.float32) if want_desc else None
        if B == 0:
            return {"eqv": eqv, "inv": inv, "desc": desc}
        _lib.check(self.lib.yoho_part1_forward(self.h, _ptr(x), B, _ptr(eqv), _ptr(inv), _ptr(desc), _stream()))
        return {"eqv": eqv, "inv": inv, "desc": desc}

    # ---- B: matching -------------------------------------------------------------------------------
    def group_mean(self, eqv):
        eqv = self._f32(eqv)
        K = eqv.shape[0]
        desc = self._empty((K, 32), torch.float32)
        _lib.check(self.lib.yoho_group_mean(self.h, _ptr(eqv), K, _ptr(desc), _stream()))
        return desc

    def nn1(self, source, target):
        """For every source row [m,F] the nearest target row [n,F] -> (dist [m] f32, idx [m] i64)."""
        s, t = self._f32(source), self._f32(target)
        m, F = s.shape
        n = t.shape[0]
        dist = self._empty((m,), torch.float32)
        idx = self._empty((m,), torch.int64)
        _lib.check(self.lib.yoho_nn1(self.h, _ptr(s), m, _ptr(t), n, F, _ptr(dist), _ptr(idx), _stream()))
        return dist, idx

    def mutual_nn(self, dA, dB, want_nn=False):
        """dA [Ka,32], dB [Kb,32] -> (pairs buffer [min,2] i64, n_pairs device i32[1]) (+ nnA, nnB)."""
        dA, dB = self._f32(dA), self._f32(dB)
        Ka, Kb = dA.shape[0], dB.shape[0]
        pairs = self._empty((min(Ka, Kb), 2), torch.int64)
        n = self._empty((1,), torch.int32)
        nnA = self._empty((Ka,), torch.int32) if want_nn else None
        nnB = self._empty((Kb,), torch.int32) if want_nn else None
        _lib.check(self.lib.yoho_mutual_nn(self.h, _ptr(dA), Ka, _ptr(dB), Kb, _ptr(pairs), _ptr(n), _ptr(nnA),
                                           _ptr(nnB), _stream()))
        if want_nn:
            return pairs, n, nnA, nnB
        return pairs, n

    # ---- C: rotation index -------------------------------------------------------------------------
    def rot_argmax(self, des1, des2, pairs=None, want_cor=False):
        """idx[m] = argmax_a sum_{f,g} des1[r1(m),f,P[a][g]] des2[r2(m),f,g].
        With `pairs` [M,2] (row in fragment 0, row in fragment 1): r1 = pairs[:,1] indexes des1 (fragment 1's eqv)
        and r2 = pairs[:,0] indexes des2 (fragment 0's eqv) — the reference's call order
        (tests/extractor.py:97-99).  Without pairs the rows are matched one to one."""
        des1, des2 = self._f32(des1), self._f32(des2)
        if pairs is not None:
            self._check_rows(pairs, des2.shape[0], des1.shape[0], "rot_argmax")
            pairs = self._i64(pairs)
            M = pairs.shape[0]
            r1 = ctypes.c_void_p(pairs.data_ptr() + 8) if M else None
            r2 = ctypes.c_void_p(pairs.data_ptr()) if M else None
            stride = 2
        else:
            M = des1.shape[0]
            r1 = r2 = None
            stride = 1
        idx = self._empty((M,), torch.int64)
        cor = self._empty((M, 60), torch.float32) if want_cor else None
        if M == 0:
            return (idx, cor) if want_cor else idx
        _lib.check(self.lib.yoho_rot_argmax(self.h, _ptr(des1), r1, _ptr(des2), r2, stride, M, _ptr(idx), _ptr(cor),
                                            _stream()))
        return (idx, cor) if want_cor else idx

    # ---- D: PartII ---------------------------------------------------------------------------------
    def part2(self, fcgf0, fcgf1, yoho0, yoho1, pre_idx, pairs=None, kps0=None, kps1=None):
        """Fragment tensors [K,32,60] + pairs [M,2] (or per-match rows when pairs is None), pre_idx [M]
        -> quat [M,4] f32 and, when keypoints are given, trans [M,3,4] f64."""
        f0, f1, y0, y1 = self._f32(fcgf0), self._f32(fcgf1), self._f32(yoho0), self._f32(yoho1)
        self._check_bins(pre_idx, "part2")
        if pairs is not None:
            self._check_rows(pairs, min(f0.shape[0], y0.shape[0]), min(f1.shape[0], y1.shape[0]), "part2")
        pre = self._i64(pre_idx)
        M = pre.shape[0]
        pr = self._i64(pairs) if pairs is not None else None
        quat = self._empty((M, 4), torch.float32)
        trans = None
        k0 = k1 = None
        if kps0 is not None:
            k0, k1 = self._f64(kps0), self._f64(kps1)
            trans = self._empty((M, 3, 4), torch.float64)
        if M == 0:
            return quat, trans
        _lib.check(self.lib.yoho_part2_forward(self.h, _ptr(f0), _ptr(f1), _ptr(y0), _ptr(y1), _ptr(pr), _ptr(pre), M,
                                               _ptr(k0), _ptr(k1), _ptr(quat), _ptr(trans), _stream()))
        return quat, trans

    # ---- next row: group-feature lift tail ----------------------------------------------------------------
    def lift_group_features(self, kps, pts_list, feats_list, want_nn=False):
        """kps [K,3] f64; pts_list[g] [n_g,3] f32 (the rotation-g down-sampled cloud), feats_list[g] [n_g,32] f32
        -> group feature [K,32,60] f32 (YOHO_testset.py:153-166)."""
        assert len(pts_list) == 60 and len(feats_list) == 60
        k = self._f64(kps)
        K = k.shape[0]
        offs = np.zeros(61, np.int32)
        offs[1:] = np.cumsum([int(p.shape[0]) for p in pts_list])
        pts = torch.cat([self._f32(p).reshape(-1, 3) for p in pts_list])
        feats = torch.cat([self._f32(f).reshape(-1, 32) for f in feats_list])
        od = torch.from_numpy(offs).to(self.device)
        out = torch.zeros((K, 32, 60), device=self.device, dtype=torch.float32)
        nn = self._empty((60, K), torch.int64) if want_nn else None
        _lib.check(self.lib.yoho_lift_group_features(self.h, _ptr(k), K, _ptr(pts), _ptr(feats), _ptr(od), _ptr(out), _ptr(nn),
                                                     _stream()))
        return (out, nn) if want_nn else out

    # ---- evaluation metrics ---------------------------------------------------------
    def fmr_counts(self, keys0, keys1, offsets, gt, threshold):
        """Matched keypoints of n pairs concatenated ([Mtot,3] f64 each), offsets [n+1], gt [n,4,4] (or [n,3,4]) ->
        int32[n] matches closer than `threshold` under the ground truth (tests/evaluator.py:57-66)."""
        k0, k1 = self._f64(keys0).reshape(-1, 3), self._f64(keys1).reshape(-1, 3)
        off = self._i64(offsets)
        n = off.shape[0] - 1
        g = self._f64(gt)
        if g.dim() == 3 and g.shape[1] == 3:
            pad = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=torch.float64, device=self.device).expand(g.shape[0], 1, 4)
            g = torch.cat([g, pad], 1).contiguous()
        assert g.shape == (n, 4, 4) and k0.shape == k1.shape
        counts = torch.zeros((max(n, 0),), device=self.device, dtype=torch.int32)
        if n > 0:
            _lib.check(self.lib.yoho_fmr_batch(self.h, _ptr(k0), _ptr(k1), _ptr(off), _ptr(g), n, ctypes.c_double(threshold),
                                               _ptr(counts), _stream()))
        return counts

    def registration_errors(self, est, gt, info=None):
        """est, gt [n,4,4] f64, info [n,6,6] f64 or None -> (p or None, rre_deg, rte), utils/RR_cal.py:13-65."""
        e, g = self._f64(est), self._f64(gt)
        n = e.shape[0]
        assert e.shape == (n, 4, 4) and g.shape == (n, 4, 4)
        inf = self._f64(info) if info is not None else None
        p = self._empty((n,), torch.float64) if inf is not None else None
        re = self._empty((n,), torch.float64)
        te = self._empty((n,), torch.float64)
        if n > 0:
            _lib.check(self.lib.yoho_registration_errors(self.h, _ptr(e), _ptr(g), _ptr(inf), n, _ptr(p), _ptr(re), _ptr(te), _stream()))
        return p, re, te

    # ---- the whole pair in one call (csrc/pair.cu) ---------------------------------------------------------
    _ESIZE = {torch.float32: 4, torch.float64: 8, torch.int64: 8, torch.int32: 4, torch.uint8: 1}
    _PAIR_PTRS = ("eqvA", "eqvB", "descA", "descB", "pairs", "n_pairs", "dr_index", "k0", "k1", "hyp", "c_status", "T_c",
                  "c_best", "c_inl", "c_mask", "quat", "trans", "order", "T_o", "o_best", "o_inl", "o_mask")

    def _pair_layout(self, Ka, Kb, c_iters, have):
        """Offsets of the per-pair outputs inside ONE allocation (cached per shape: the host cost of a pair is part of its time)."""
        key = (Ka, Kb, c_iters, have)
        cache = self.__dict__.setdefault("_pair_layouts", {})
        lay = cache.get(key)
        if lay is None:
            cap = max(1, min(Ka, Kb))
            # T_co holds the two results back to back (T_c = T_co[0], T_o = T_co[1]): one D2H copy for both
            spec = [("T_co", (2, 3, 4), torch.float64), ("pairs", (cap, 2), torch.int64), ("n_pairs", (1,), torch.int32),
                    ("dr_index", (cap,), torch.int64), ("k0", (cap, 3), torch.float64), ("k1", (cap, 3), torch.float64),
                    ("hyp", (max(1, c_iters), 3), torch.int32), ("c_status", (1,), torch.int32), ("c_best", (1,), torch.int32),
                    ("c_inl", (1,), torch.int32), ("c_mask", (cap,), torch.uint8), ("quat", (cap, 4), torch.float32),
                    ("trans", (cap, 3, 4), torch.float64), ("order", (cap,), torch.int32), ("o_best", (1,), torch.int32),
                    ("o_inl", (1,), torch.int32), ("o_mask", (cap,), torch.uint8)]
            if not have:
                spec += [("eqvA", (Ka, 32, 60), torch.float32), ("eqvB", (Kb, 32, 60), torch.float32),
                         ("descA", (Ka, 32), torch.float32), ("descB", (Kb, 32), torch.float32)]
            ent, total = {}, 0
            for name, shape, dt in spec:
                n = int(np.prod(shape)) * self._ESIZE[dt]
                ent[name] = (total, n, shape, dt)
                total += (n + 255) // 256 * 256
            lay = cache[key] = (ent, max(total, 256))
        return lay

    def _pair_io(self, featA, featB, kpsA, kpsB, c_iters, o_iters, c_dist, o_dist, seed, eqvA=None, eqvB=None, descA=None,
                 descB=None):
        """Builds the yoho_pair_io of one pair: every output lives in ONE allocation sized for min(Ka, Kb) matches."""
        if (eqvA is None) != (descA is None) or (eqvB is None) != (descB is None) or (eqvA is None) != (eqvB is None):
            raise ValueError("precomputed PartI outputs must be given together: eqvA, eqvB, descA, descB")
        fa, fb, ka, kb = self._f32(featA), self._f32(featB), self._f64(kpsA), self._f64(kpsB)
        Ka, Kb = fa.shape[0], fb.shape[0]
        have = eqvA is not None
        ent, total = self._pair_layout(Ka, Kb, int(c_iters), have)
        buf = torch.empty((total,), dtype=torch.uint8, device=self.device)
        base = buf.data_ptr()
        ext = {}
        if have:
            ext = dict(eqvA=self._f32(eqvA), eqvB=self._f32(eqvB), descA=self._f32(descA), descB=self._f32(descB))
        io = _lib.yoho_pair_io()
        io.featA, io.featB, io.kpsA, io.kpsB = (x.data_ptr() or base for x in (fa, fb, ka, kb))   # empty fragment: any valid pointer
        io.Ka, io.Kb, io.have_part1 = Ka, Kb, int(have)
        io.c_iters, io.o_iters, io.c_dist, io.o_dist, io.seed = int(c_iters), int(o_iters), float(c_dist), float(o_dist), int(seed)
        for name in self._PAIR_PTRS:
            if name in ext:
                ptr = ext[name].data_ptr() or base
            elif name == "T_c":
                ptr = base + ent["T_co"][0]
            elif name == "T_o":
                ptr = base + ent["T_co"][0] + 96
            else:
                ptr = base + ent[name][0]
            setattr(io, name, ptr)
        return io, buf, ent, ext, (fa, fb, ka, kb)

    def register_pair(self, featA, featB, kpsA, kpsB, c_iters, o_iters, c_dist, o_dist, seed, eqvA=None, eqvB=None,
                      descA=None, descB=None):
        """yoho_register_pair: PartI x2 (unless eqv/desc are given), mutual matching, rotation index, YOHO-C, PartII, YOHO-O with
        ONE host round trip.  Returns a `PairBuffers`: every output lives in one allocation sized for min(Ka, Kb) matches and is
        materialised as a tensor view only when asked for (`out["quat"]`), so the throughput path pays for two views."""
        io, buf, ent, ext, keep = self._pair_io(featA, featB, kpsA, kpsB, c_iters, o_iters, c_dist, o_dist, seed, eqvA, eqvB,
                                                descA, descB)
        M = ctypes.c_int32(0)
        _lib.check(self.lib.yoho_register_pair(self.h, ctypes.byref(io), ctypes.byref(M), _stream()))
        return PairBuffers(buf, ent, ext, int(M.value), keep)

    def register_pair_begin(self, *args, **kw):
        """First phase of `register_pair` (yoho_register_pair_begin): queues PartI x2 + matching, does not wait.  Returns the
        token to hand to `register_pair_end` (pairs end in the order they were begun)."""
        tok = self._pair_io(*args, **kw)
        _lib.check(self.lib.yoho_register_pair_begin(self.h, ctypes.byref(tok[0]), _stream()))
        return tok

    def register_pair_end(self, tok):
        io, buf, ent, ext, keep = tok
        M = ctypes.c_int32(0)
        _lib.check(self.lib.yoho_register_pair_end(self.h, ctypes.byref(io), ctypes.byref(M), _stream()))
        return PairBuffers(buf, ent, ext, int(M.value), keep)

    # ---- E: estimators -----------------------------------------------------------------------------
    def gather_kps(self, kps0, kps1, pairs):
        self._check_rows(pairs, len(kps0), len(kps1), "gather_kps")
        k0, k1, pr = self._f64(kps0), self._f64(kps1), self._i64(pairs)
        M = pr.shape[0]
        o0 = self._empty((M, 3), torch.float64)
        o1 = self._empty((M, 3), torch.float64)
        _lib.check(self.lib.yoho_gather_kps(self.h, _ptr(k0), _ptr(k1), _ptr(pr), M, _ptr(o0), _ptr(o1), _stream()))
        return o0, o1

    def c_draw(self, dr_index, iters, seed):
        """-> (hyp [iters,3] i32, status i32[1]: 0 ok, 1 degenerate statistics, 2 a rotation index outside [0,60) was seen)."""
        self._check_bins(dr_index, "c_draw")
        dr = self._i64(dr_index)
        hyp = self._empty((iters, 3), torch.int32)
        status = self._empty((1,), torch.int32)
        _lib.check(self.lib.yoho_c_draw(self.h, _ptr(dr), dr.shape[0], iters, ctypes.c_uint64(seed), _ptr(hyp),
                                        _ptr(status), _stream()))
        return hyp, status

    def _est_out(self, M, n_hyp, want_counts):
        T = self._empty((3, 4), torch.float64)
        bi = self._empty((1,), torch.int32)
        ni = self._empty((1,), torch.int32)
        mask = self._empty((M,), torch.uint8)
        counts = self._empty((n_hyp,), torch.int32) if want_counts else None
        return T, bi, ni, mask, counts

    def c_ransac(self, k0, k1, hyp, dist, signs=None, want_counts=False, fixed=None):
        k0, k1 = self._f64(k0), self._f64(k1)
        if isinstance(hyp, np.ndarray):
            hyp = torch.from_numpy(np.ascontiguousarray(hyp, np.int32))
        hyp = hyp.to(device=self.device, dtype=torch.int32).contiguous()
        sg = None
        if signs is not None:
            sg = torch.as_tensor(np.ascontiguousarray(signs, np.int8)).to(self.device)
        fx = None
        if fixed is not None:
            fx = self._f64(fixed).reshape(-1, 12)
            if sg is None or fx.shape[0] != hyp.shape[0]:
                raise ValueError("fixed transforms need signs and one [3,4] entry per hypothesis")
        M, iters = k0.shape[0], hyp.shape[0]
        T, bi, ni, mask, counts = self._est_out(M, iters, want_counts)
        _lib.check(self.lib.yoho_c_ransac(self.h, _ptr(k0), _ptr(k1), M, _ptr(hyp), _ptr(sg), _ptr(fx), iters, float(dist),
                                          _ptr(T), _ptr(bi), _ptr(ni), _ptr(mask), _ptr(counts), _stream()))
        return dict(T=T, best_iter=bi, n_inl=ni, mask=mask, counts=counts)

    def o_order(self, M, seed):
        order = self._empty((M,), torch.int32)
        _lib.check(self.lib.yoho_o_order(self.h, M, ctypes.c_uint64(seed), _ptr(order), _stream()))
        return order

    def o_score(self, k0, k1, trans, dist, order=None, max_hyp=None, want_counts=False):
        k0, k1, tr = self._f64(k0), self._f64(k1), self._f64(trans)
        od = None
        H = tr.shape[0]
        if order is not None:
            if isinstance(order, np.ndarray):
                order = torch.from_numpy(np.ascontiguousarray(order, np.int32))
            od = order.to(device=self.device, dtype=torch.int32).contiguous()
            H = od.shape[0]
        if max_hyp is not None:
            H = min(H, int(max_hyp))
        M = k0.shape[0]
        T, bi, ni, mask, counts = self._est_out(M, H, want_counts)
        _lib.check(self.lib.yoho_o_score(self.h, _ptr(k0), _ptr(k1), M, _ptr(tr), _ptr(od), H, float(dist), _ptr(T),
                                         _ptr(bi), _ptr(ni), _ptr(mask), _ptr(counts), _stream()))
        return dict(T=T, best_iter=bi, n_inl=ni, mask=mask, counts=counts)


_engines = {}
_lock = threading.Lock()


def get_engine(device=None, so3_dir=None) -> Engine:
    """Process-wide engine per (device, table dir)."""
    if not torch.cuda.is_available():
        raise _lib.YohoError("yoho_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")
    dev = torch.cuda.current_device() if device is None else int(device)
    t = _group.load(so3_dir)
    # same table contents -> same engine, whichever directory they were read from
    key = (dev, hash((t.R.tobytes(), t.P.tobytes(), t.N.tobytes())))
    with _lock:
        if key not in _engines:
            _engines[key] = Engine(dev, so3_dir)
        return _engines[key]
