"""In-memory fragment-pair pipeline: FCGF group features + keypoints of two fragments -> YOHO-C and YOHO-O
transforms, device resident between stages (the reference round-trips every stage through .npy files:
tests/evaluator.py:41-47,112-117).  This is the unit of work of BASELINE.json's metric: one cold pair =
PartI on both fragments, mutual matching, rotation index, YOHO-C, PartII, YOHO-O.

One host synchronisation per pair (the match count M sizes the downstream launches).  By default the whole pair is ONE
C-ABI call (`yoho_register_pair`, csrc/pair.cu); `fused=False` walks the stages through the per-stage entry points instead.
"""
import numpy as np
import torch

from .engine import get_engine


class PairResult(dict):
    __getattr__ = dict.get


class PairPipeline:
    def __init__(self, engine=None, c_iters=1000, o_iters=1000, c_dist=0.07, o_dist=0.09, seed=0, fused=True):
        self.eng = engine or get_engine()
        if not (self.eng.has_part1 and self.eng.has_part2):
            raise RuntimeError("load PartI and PartII weights into the engine first (No model exists)")
        self.c_iters, self.o_iters = int(c_iters), int(o_iters)
        self.c_dist, self.o_dist = float(c_dist), float(o_dist)
        self.seed = int(seed)
        self.fused = bool(fused)      # one C-ABI call per pair (csrc/pair.cu); False: one call per stage (same results)
        self._copy_stream = None

    # ---- device-resident pair ------------------------------------------------------------------------
    def register(self, featA, featB, kpsA, kpsB, eqvA=None, eqvB=None, descA=None, descB=None, seed=None, lean=False):
        """All inputs CUDA tensors: feat [K,32,60] f32, kps [K,3] f64.  Pass precomputed eqv/desc to skip PartI
        (the amortised regime: one PartI pass per fragment per dataset, tests/extractor.py:46-47).
        lean=True (fused path only) returns just `M` and `T_co` ([2,3,4]: YOHO-C, YOHO-O) — the throughput callers' form, which
        skips building the per-stage tensor views on the host."""
        e = self.eng
        if self.fused:
            if seed is None:
                self.seed += 1
                seed = self.seed
            t = e.register_pair(featA, featB, kpsA, kpsB, self.c_iters, self.o_iters, self.c_dist, self.o_dist, seed,
                                eqvA=eqvA, eqvB=eqvB, descA=descA, descB=descB)
            M = t.M
            if lean:                                  # throughput path: the two transforms (one contiguous [2,3,4] block) and M
                return PairResult(M=M, T_co=t["T_co"], _buffers=t)
            out = PairResult(M=M, pairs=t["pairs"][:M], eqvA=t["eqvA"], eqvB=t["eqvB"], T_c=t["T_c"], T_o=t["T_o"])
            if M == 0:
                out.update(dr_index=t["dr_index"][:0], c_best=-1, o_best=-1)
                return out
            out.update(dr_index=t["dr_index"][:M], k0=t["k0"][:M], k1=t["k1"][:M], hyp=t["hyp"][:self.c_iters], c_status=t["c_status"],
                       c_best=t["c_best"], c_inl=t["c_inl"], c_mask=t["c_mask"][:M], quat=t["quat"][:M], trans_pre=t["trans"][:M],
                       o_best=t["o_best"], o_inl=t["o_inl"], o_mask=t["o_mask"][:M])
            return out
        if eqvA is None:
            oa = e.part1(featA, want_inv=False, want_desc=True)
            eqvA, descA = oa["eqv"], oa["desc"]
        if eqvB is None:
            ob = e.part1(featB, want_inv=False, want_desc=True)
            eqvB, descB = ob["eqv"], ob["desc"]
        pairs_buf, n_dev = e.mutual_nn(descA, descB)
        M = int(n_dev.item())                               # the one host sync of the pair
        pairs = pairs_buf[:M]
        out = PairResult(M=M, pairs=pairs, eqvA=eqvA, eqvB=eqvB)
        if M == 0:
            eye = torch.eye(4, dtype=torch.float64, device=e.device)[:3]
            out.update(dr_index=pairs.new_zeros((0,)), T_c=eye, T_o=eye.clone(), c_best=-1, o_best=-1)
            return out
        dr = e.rot_argmax(eqvB, eqvA, pairs=pairs)          # Batch_Des2R_torch(feats1[m1], feats0[m0])
        k0, k1 = e.gather_kps(kpsA, kpsB, pairs)
        if seed is None:
            self.seed += 1
            seed = self.seed
        hyp, status = e.c_draw(dr, self.c_iters, seed)
        rc = e.c_ransac(k0, k1, hyp, self.c_dist)
        # degenerate rotation statistics: the reference skips RANSAC and returns the identity (tests/estimator.py:107-108)
        eye = torch.eye(4, dtype=torch.float64, device=e.device)[:3]
        degenerate = status.to(torch.bool)
        rc["T"] = torch.where(degenerate, eye, rc["T"])
        rc["best_iter"] = torch.where(degenerate, torch.full_like(rc["best_iter"], -1), rc["best_iter"])
        quat, trans = e.part2(featA, featB, eqvA, eqvB, dr, pairs=pairs, kps0=kpsA, kps1=kpsB)
        order = e.o_order(M, seed)
        ro = e.o_score(k0, k1, trans, self.o_dist, order=order, max_hyp=self.o_iters)
        out.update(dr_index=dr, k0=k0, k1=k1, hyp=hyp, c_status=status, T_c=rc["T"], c_best=rc["best_iter"],
                   c_inl=rc["n_inl"], c_mask=rc["mask"], quat=quat, trans_pre=trans, T_o=ro["T"],
                   o_best=ro["best_iter"], o_inl=ro["n_inl"], o_mask=ro["mask"])
        return out

    def register_many(self, inputs, lookahead=1):
        """Throughput form of `register` for a SEQUENCE of device-resident pairs (a dataset's pair list): generator over
        `(featA, featB, kpsA, kpsB)` tuples yielding one lean result (`M`, `T_co` [2,3,4]) per pair, in order.  Pair i+1's PartI
        and matching are queued BEFORE the host waits for pair i's match count (yoho_register_pair_begin / _end), so the one
        host synchronisation of a pair never idles the device.  Same seeds, same results as calling `register` pair by pair."""
        e = self.eng
        assert self.fused, "register_many needs the fused pair call"
        pend = []
        for args in inputs:
            self.seed += 1
            pend.append(e.register_pair_begin(*args, self.c_iters, self.o_iters, self.c_dist, self.o_dist, self.seed))
            if len(pend) > lookahead:
                t = e.register_pair_end(pend.pop(0))
                yield PairResult(M=t.M, T_co=t["T_co"], _buffers=t)
        while pend:
            t = e.register_pair_end(pend.pop(0))
            yield PairResult(M=t.M, T_co=t["T_co"], _buffers=t)

    # ---- host-facing calls (the e2e path: host buffers in, host transforms out) -----------------------------
    @staticmethod
    def pin(featA, featB, kpsA, kpsB):
        """numpy -> pinned host tensors (do this once per pair, outside any timed region)."""
        from .hostutil import numa_local

        def p(a, dt):
            t = torch.empty(a.shape, dtype=dt, pin_memory=True)
            t.copy_(torch.from_numpy(np.ascontiguousarray(a)))
            return t
        # allocate (and first-touch) on the memory node next to the GPU: the H2D DMA of 77 MB per pair reads it from there
        with numa_local(torch.cuda.current_device() if torch.cuda.is_available() else 0) as nl:
            out = (p(featA, torch.float32), p(featB, torch.float32), p(kpsA, torch.float64), p(kpsB, torch.float64))
        PairPipeline.numa_note = nl.note
        return out

    def register_pinned(self, fa_pin, fb_pin, ka_pin, kb_pin):
        """Pinned host tensors in -> numpy transforms out.  The four H2D copies run on a side stream; PartI of fragment A
        starts as soon as A has landed, so the copy of fragment B hides behind it."""
        dev = self.eng.device
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            fa = fa_pin.to(dev, non_blocking=True)
            evA = torch.cuda.Event(); evA.record(cs)
            fb = fb_pin.to(dev, non_blocking=True)
            ka = ka_pin.to(dev, non_blocking=True)
            kb = kb_pin.to(dev, non_blocking=True)
            evB = torch.cuda.Event(); evB.record(cs)
        for t in (fa, fb, ka, kb):
            t.record_stream(main)
        main.wait_event(evA)
        oa = self.eng.part1(fa, want_inv=False, want_desc=True)
        main.wait_event(evB)
        ob = self.eng.part1(fb, want_inv=False, want_desc=True)
        r = self.register(fa, fb, ka, kb, eqvA=oa["eqv"], eqvB=ob["eqv"], descA=oa["desc"], descB=ob["desc"])
        res = torch.stack([r["T_c"], r["T_o"]]).cpu().numpy()          # D2H of the result (synchronises)
        return dict(T_c=res[0], T_o=res[1], M=r["M"])

    def register_stream(self, pinned_pairs, stats=None):
        """Throughput form of `register_pinned` for a sequence of pairs (the way a dataset is processed, tests/evaluator.py:41-47):
        generator over `(fa_pin, fb_pin, ka_pin, kb_pin)` tuples yielding one `dict(T_c, T_o, M)` per pair, in order.
        Software pipeline over three persistent device input sets: while pair n's rotation index / estimators / PartII run, pair
        n+1's PartI has already been queued behind them (split-phase pair call: the host never idles the device waiting for a
        match count) and pair n+2's inputs are crossing PCIe on the side stream; the D2H of pair n's transforms (into a pinned
        buffer) is awaited only after pair n+1 has been queued.  Every pair's inputs still come from host memory and every
        result is read back to the host.
        `stats` (a dict) switches on accounting with CUDA events on the main stream and host clocks; on return it holds lists
        (ms per pair): copy_wait = main stream stalled on the upload of a pair, period = time between the last kernels of
        consecutive pairs, and the host times upload_host / begin_host / end_host / d2h_wait_host of the four host-side steps."""
        import time as _time
        e = self.eng
        dev = e.device
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        NS = 3
        # persistent sets of device input buffers (no allocator traffic in steady state): pair n uses set n % 3; the copy of pair
        # n + 3 into the same set waits for the event recorded after pair n's last kernel was queued
        if not hasattr(self, "_in_sets") or len(self._in_sets) != NS:
            self._in_sets = [None] * NS
            self._in_free = [None] * NS
            self._res_pin = [torch.empty((2, 3, 4), dtype=torch.float64, pin_memory=True) for _ in range(NS)]
        timing = stats is not None

        def note(key, t0):
            if timing:
                stats.setdefault(key, []).append(1e3 * (_time.perf_counter() - t0))

        def upload(pp, slot):
            t0 = _time.perf_counter()
            cur = self._in_sets[slot]
            if cur is None or any(c.shape != t.shape or c.dtype != t.dtype for c, t in zip(cur, pp)):
                cur = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in pp]
                self._in_sets[slot] = cur
                cs.wait_stream(main)
            if self._in_free[slot] is not None:
                cs.wait_event(self._in_free[slot])
            with torch.cuda.stream(cs):
                for c, t in zip(cur, pp):
                    c.copy_(t, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            note("upload_host", t0)
            return cur, ev

        waits, ends = [], []

        def begin(up):
            (fa, fb, ka, kb), ev = up
            t0 = _time.perf_counter()
            if timing:
                w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                w0.record(main)
            main.wait_event(ev)
            if timing:
                w1.record(main)
                waits.append((w0, w1))
            self.seed += 1
            tok = e.register_pair_begin(fa, fb, ka, kb, self.c_iters, self.o_iters, self.c_dist, self.o_dist, self.seed)
            note("begin_host", t0)
            return tok

        def end(tok, n):
            t0 = _time.perf_counter()
            t = e.register_pair_end(tok)
            free = torch.cuda.Event(enable_timing=timing)
            free.record(main)                           # every kernel reading this input set has been queued
            self._in_free[n % NS] = free
            if timing:
                ends.append(free)
            host = self._res_pin[n % NS]
            host.copy_(t["T_co"], non_blocking=True)
            dv = torch.cuda.Event()
            dv.record(main)
            note("end_host", t0)
            return host, dv, t.M

        def finish(pend):
            host, ev, M = pend
            t0 = _time.perf_counter()
            ev.synchronize()
            note("d2h_wait_host", t0)
            res = host.numpy().copy()
            return dict(T_c=res[0], T_o=res[1], M=M)

        if not self.fused:
            raise RuntimeError("register_stream needs the fused pair call")
        it = iter(pinned_pairs)
        ups = []                                        # uploaded, not yet begun (at most 2)
        n_up = 0
        for _ in range(2):
            try:
                ups.append(upload(next(it), n_up % NS))
                n_up += 1
            except StopIteration:
                break
        if not ups:
            return
        toks = [begin(ups.pop(0))]                      # begun, not yet ended (at most 2)
        pending = None
        n = 0
        while toks:
            try:
                ups.append(upload(next(it), n_up % NS))     # pair n + 2: crosses PCIe during pair n + 1's PartI
                n_up += 1
            except StopIteration:
                pass
            if ups:
                toks.append(begin(ups.pop(0)))          # pair n + 1's PartI + matching queued before we wait for pair n's count
            cur = end(toks.pop(0), n)
            if pending is not None:
                yield finish(pending)
            pending = cur
            n += 1
        if pending is not None:
            yield finish(pending)
        if timing and ends:
            torch.cuda.synchronize()
            stats["copy_wait"] = [a.elapsed_time(b) for a, b in waits]
            stats["period"] = [ends[i - 1].elapsed_time(ends[i]) for i in range(1, len(ends))]

    def register_host(self, featA, featB, kpsA, kpsB):
        """numpy in (feat [K,32,60] f32, kps [K,3] f64) -> numpy transforms out; staging + H2D + D2H inside."""
        return self.register_pinned(*self.pin(featA, featB, kpsA, kpsB))

    @staticmethod
    def h2d_bytes(K):
        return 2 * K * 32 * 60 * 4 + 2 * K * 3 * 8

    @staticmethod
    def d2h_bytes():
        return 2 * 12 * 8 + 4     # two 3x4 f64 transforms + the match count
