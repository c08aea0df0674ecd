"""Replays the op trace of one CKKS bootstrapping (lattigo_b200/boottrace.py, derived from circuits/ckks/bootstrapping/evaluator.go:518-563)
through the C ABI on a batch of synthetic ciphertexts: uniform residues, random evaluation keys and diagonals of the right shapes
(throughput only depends on shapes). Every op of the trace is the entry point the cgo shim would call for it; what is
measured is therefore the device time of the ring / evaluator layer of a bootstrap, not its numerical output."""
from __future__ import annotations

import ctypes
from typing import Dict, List

from . import _lib, boottrace
from .lintrans import Evaluator as LtEvaluator, LinearTransformation
from .ring import RING_Q
from .rlwe import CKKSEvaluator, Evaluator, GadgetCiphertext, div_by_last_modulus_many


class BootstrapReplay:
    def __init__(self, ctx, batch: int, generator, trace: List[dict] = None):
        import torch
        self.torch = torch
        self.ctx, self.batch, self.g = ctx, batch, generator
        self.dev = torch.device("cuda", ctx.device)
        self.N = ctx.N
        self.trace = trace if trace is not None else boottrace.bootstrap_trace(logN=ctx.logN)
        self.level_top, self.levelP = len(ctx.Q) - 1, len(ctx.P) - 1
        self.nd = (self.level_top + self.levelP + 1) // (self.levelP + 1)
        self.ev = Evaluator(ctx)
        self._key_pool: Dict[int, GadgetCiphertext] = {}
        self.relin = self._new_key()
        self.ckks = CKKSEvaluator(ctx, self.relin)
        self.k_dense_to_sparse = self._new_key()
        self.k_sparse_to_dense = self._new_key()
        self.k_conj = self._new_key()
        self.lt_ev, self.lts = None, []
        keys = {}
        helper = LtEvaluator(ctx, {})
        for op in self.trace:
            if op["op"] != "lintrans":
                continue
            lvl = op["level"]
            index, r1, r2 = boottrace.bsgs_index(op["diags"], 1 << (ctx.logN - 1), op["N1"])
            for r in [x for x in r1 if x] + [x for x in r2 if x]:
                gal = helper.GaloisElement(r)
                if gal not in keys:
                    keys[gal] = self._new_key()
            vec = {d: self._rand_rows(ctx.Q[: lvl + 1] + ctx.P, ()) for d in op["diags"]}
            self.lts.append(LinearTransformation(vec, lvl, self.levelP, ctx.logN - 1, op["N1"]))
        self.lt_ev = LtEvaluator(ctx, keys)
        self.n_galois_keys = len(keys)
        self.monomial = {}                       # level -> an NTT + Montgomery "multiply by i" polynomial (X^{N/2}); any fixed poly costs the same

    def _rand_rows(self, mods, lead):
        torch = self.torch
        out = torch.empty(tuple(lead) + (len(mods), self.N), dtype=torch.int64, device=self.dev)
        for i, m in enumerate(mods):
            out[..., i, :] = torch.randint(0, m, tuple(lead) + (self.N,), generator=self.g, device=self.dev, dtype=torch.int64)
        return out

    def _new_key(self):
        return GadgetCiphertext(self.ctx, self._rand_rows(self.ctx.Q + self.ctx.P, (self.nd, 1, 2)), self.level_top, self.levelP)

    def _ct(self, level):
        return self._rand_rows(self.ctx.Q[: level + 1], (self.batch, 2))

    def _mono(self, level):
        if level not in self.monomial:
            self.monomial[level] = self._rand_rows(self.ctx.Q[: level + 1], ())
        return self.monomial[level]

    # ---- one pass over the trace; returns {phase: milliseconds} and the number of ops run -------------------------------------------
    def run(self):
        torch = self.torch
        ctx, B, N = self.ctx, self.batch, self.N
        L = _lib.lib()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        events = []
        phase = None
        ct = self._ct_cache(0)
        cts = {"main": ct}
        lt_i = 0
        for op in self.trace:
            if op["phase"] != phase:
                e = torch.cuda.Event(enable_timing=True); e.record(); events.append((op["phase"], e)); phase = op["phase"]
            lvl, kind = op["level"], op["op"]
            rq = ctx.ringQ.AtLevel(lvl)
            cur = cts.get(op.get("ct", "main"), cts["main"])
            if cur.shape[-2] != lvl + 1:
                cur = self._ct_cache(lvl)                      # synthetic operands: a ciphertext of the level the op runs at
            if kind == "keyswitch":
                c0 = ctx.new_poly(lvl + 1, B); c1 = ctx.new_poly(lvl + 1, B)
                self.ev.GadgetProduct(lvl, cur[:, 1].contiguous(), self.k_dense_to_sparse, c0, c1)
                rq.Add(c0, cur[:, 0].contiguous(), c0)
            elif kind == "intt":
                x = cur.reshape(2 * B, lvl + 1, N)
                rq.INTT(x, x)
            elif kind == "modup_centered":
                src = self._ct_cache(0)
                nq, npp = lvl + 1, self.levelP + 1
                # c0 -> Q rows (row 0 is the input itself); c1 -> a QP-stacked block, the layout of one DecomposeNTT digit
                self._ext0 = torch.empty((B, nq, N), dtype=torch.int64, device=self.dev)
                self._ext0[:, 0] = src[:, 0, 0]
                self._ext1 = torch.empty((B, nq + npp, N), dtype=torch.int64, device=self.dev)
                _lib.check(L.lgpu_modup_centered(ctx.h, ctypes.c_void_p(src.data_ptr()), 1, lvl, -1, 0, ctypes.c_void_p(self._ext0.data_ptr()), None,
                                                 B, 2 * N, nq * N, 0, st))
                _lib.check(L.lgpu_modup_centered(ctx.h, ctypes.c_void_p(src.data_ptr() + N * 8), 0, lvl, self.levelP, 1, ctypes.c_void_p(self._ext1.data_ptr()),
                                                 ctypes.c_void_p(self._ext1.data_ptr() + nq * N * 8), B, 2 * N, (nq + npp) * N, (nq + npp) * N, st))
            elif kind == "ntt_qp_per_digit":
                nq, npp = lvl + 1, self.levelP + 1
                bs = (nq + npp) * N
                self._decomp = torch.empty((self.nd, B, nq + npp, N), dtype=torch.int64, device=self.dev)
                for d in range(self.nd):                       # the reference transforms the extended c1 into every digit buffer (:691-697)
                    dst = self._decomp[d].data_ptr()
                    _lib.check(L.lgpu_ntt(ctx.h, 0, lvl, ctypes.c_void_p(self._ext1.data_ptr()), ctypes.c_void_p(dst), 0, B, bs, st))
                    _lib.check(L.lgpu_ntt(ctx.h, 1, self.levelP, ctypes.c_void_p(self._ext1.data_ptr() + nq * N * 8), ctypes.c_void_p(dst + nq * N * 8), 0, B, bs, st))
            elif kind == "ntt":
                rq.NTT(self._ext0, self._ext0)
                self._c0 = self._ext0
            elif kind == "mulscalar":
                rq.MulScalar(self._c0, 12345, self._c0)
            elif kind == "gadget_product_hoisted":
                c0 = ctx.new_poly(lvl + 1, B); c1 = ctx.new_poly(lvl + 1, B)
                self.ev.GadgetProductHoisted(lvl, self._decomp, self.k_sparse_to_dense, c0, c1)
                rq.Add(c0, self._c0, c0)
                cts["main"] = torch.stack([c0, c1], dim=1)
                del self._decomp, self._ext0, self._ext1
            elif kind == "lintrans":
                out = torch.empty_like(cur)
                self.lt_ev.Evaluate(cur, self.lts[lt_i], out)
                lt_i += 1
                cts["main"] = out
            elif kind == "rescale":
                out = torch.empty((B, 2, lvl, N), dtype=torch.int64, device=self.dev)
                div_by_last_modulus_many(ctx, RING_Q, lvl, True, True, 1, cts["main"].reshape(2 * B, lvl + 1, N), out.view(2 * B, lvl, N))
                cts["main"] = out
            elif kind == "conjugate":
                out = torch.empty_like(cur)
                self.ev.Automorphism(cur, 2 * N - 1, self.k_conj, out)
                cts["real"], cts["imag"] = out, cur
            elif kind in ("add", "add_const"):
                x = cur.reshape(2 * B, lvl + 1, N)
                for _ in range(op.get("count", 1)):
                    rq.Add(x, x, x)
            elif kind == "mul_by_i":
                x = cur.reshape(2 * B, lvl + 1, N)
                rq.MulCoeffsMontgomery(x, self._bcast(lvl, 2 * B), x)
            elif kind == "mulrelin_rescale":
                out = self.ckks.MulRelinRescaleNew(cur, cur)
                cts[op.get("ct", "main")] = out
            else:
                raise ValueError("unknown op %r" % kind)
        e = torch.cuda.Event(enable_timing=True); e.record(); events.append(("end", e))
        torch.cuda.synchronize()
        ms: Dict[str, float] = {}
        for (ph, a), (_, b) in zip(events[:-1], events[1:]):
            ms[ph] = ms.get(ph, 0.0) + a.elapsed_time(b)
        return ms, len(self.trace)

    def _ct_cache(self, level):
        if not hasattr(self, "_cts"):
            self._cts = {}
        if level not in self._cts:
            self._cts[level] = self._ct(level)
        return self._cts[level]

    def _bcast(self, level, n):
        key = ("b", level, n)
        if key not in self.monomial:
            self.monomial[key] = self._mono(level).unsqueeze(0).expand(n, level + 1, self.N).contiguous()
        return self.monomial[key]
