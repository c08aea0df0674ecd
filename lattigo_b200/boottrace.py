"""Op trace of one CKKS bootstrapping (circuits/ckks/bootstrapping/evaluator.go:518-563) for BASELINE config 5, derived from the
reference's Go by restating its integer bookkeeping (Go cannot run here): which ring / evaluator entry point is called, at which
level, with how many diagonals / rotations. Host logic only -- no device work; bench.py --workload bootstrap replays the trace
through the C ABI with synthetic operands.

    ModUp                bootstrapping/evaluator.go:616-771   (dense->sparse key switch at level 0, centred ModUp q -> QP, one hoisted
                                                                gadget product back to the dense key at the top level)
    CoeffsToSlots        dft/dft.go:240-300, :343-375          (depth-4 factorised DFT = 4 BSGS linear transformations, each + Rescale;
                                                                then one conjugation and the real / imaginary split)
    EvalMod (x2)         mod1/mod1_evaluator.go:31-143         (Chebyshev polynomial of degree 30 by Paterson-Stockmeyer, 3 double angles)
    SlotsToCoeffs        dft/dft.go:318-341                    (depth-3 factorised DFT)
Restated pieces: dft.MatrixLiteral.computeBootstrappingDFTIndexMap (dft/dft.go:544-658), lintrans.FindBestBSGSRatio / BSGSIndex
(circuits/common/lintrans/lintrans.go:321-367), PowerBasis.genPower / SplitDegree (circuits/common/polynomial/power_basis.go:34-160),
bignum.OptimalSplit (utils/bignum/polynomial.go:14-23), the Paterson-Stockmeyer giant-step merge
(circuits/common/polynomial/polynomial_evaluator.go:95-215)."""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

ENCODE, DECODE = "HomomorphicEncode", "HomomorphicDecode"


# ---- dft/dft.go:544-658 ---------------------------------------------------------------------------------------------------
def _gen_wfft_index_map(logL, level, lt_type, bitreversed):
    if (lt_type == ENCODE and not bitreversed) or (lt_type == DECODE and bitreversed):
        rot = 1 << (level - 1)
    else:
        rot = 1 << (logL - level)
    return {0, rot, (1 << logL) - rot}


def _next_level_fft_index_map(vec, logL, N, next_level, lt_type, bitreversed):
    if (lt_type == ENCODE and not bitreversed) or (lt_type == DECODE and bitreversed):
        rot = (1 << (next_level - 1)) & (N - 1)
    else:
        rot = (1 << (logL - next_level)) & (N - 1)
    out = set()
    for i in vec:
        out |= {i, (i + rot) & (N - 1), (i - rot) & (N - 1)}
    return out


def dft_index_map(logN: int, log_slots: int, depth: int, lt_type: str, repack_imag_as_real=True, bitreversed=False) -> List[List[int]]:
    """computeBootstrappingDFTIndexMap: the non-zero diagonals of each of the `depth` factor matrices."""
    level = log_slots
    merge = [0] * depth
    for i in range(depth):
        d = int(math.ceil(level / (depth - i)))
        if lt_type == ENCODE:
            merge[i] = d
        else:
            merge[depth - i - 1] = d
        level -= d
    level = log_slots
    maps = []
    for i in range(depth):
        if log_slots < logN - 1 and lt_type == DECODE and i == 0 and repack_imag_as_real:
            m = {0, 1 << log_slots}
            m = _next_level_fft_index_map(m, log_slots, 2 << log_slots, level, lt_type, bitreversed)
            nxt = level - 1
            for _ in range(merge[i] - 1):
                m = _next_level_fft_index_map(m, log_slots, 2 << log_slots, nxt, lt_type, bitreversed)
                nxt -= 1
        else:
            m = _gen_wfft_index_map(log_slots, level, lt_type, bitreversed)
            nxt = level - 1
            for _ in range(merge[i] - 1):
                m = _next_level_fft_index_map(m, log_slots, 1 << log_slots, nxt, lt_type, bitreversed)
                nxt -= 1
        maps.append(sorted(m))
        level -= merge[i]
    return maps


# ---- circuits/common/lintrans/lintrans.go:321-367 ---------------------------------------------------------------------------
def bsgs_index(diags, slots, N1):
    index: Dict[int, List[int]] = {}
    r1, r2 = set(), set()
    for rot in diags:
        rot &= slots - 1
        i1 = ((rot // N1) * N1) & (slots - 1)
        i2 = rot & (N1 - 1)
        index.setdefault(i1, []).append(i2)
        r1.add(i1); r2.add(i2)
    return index, sorted(r1), sorted(r2)


def find_best_bsgs_ratio(diags, max_n, log_max_ratio):
    max_ratio = float(1 << log_max_ratio)
    N1 = 1
    while N1 < max_n:
        _, r1, r2 = bsgs_index(diags, max_n, N1)
        nb1, nb2 = len(r1) - 1, len(r2) - 1
        if nb1 and nb2 / nb1 == max_ratio:
            return N1
        if nb1 and nb2 / nb1 > max_ratio:
            return N1 // 2
        N1 <<= 1
    return 1


# ---- polynomial evaluation: number of ciphertext-ciphertext products and their levels ---------------------------------------------
def _split_degree(n):
    if n & (n - 1) == 0:
        return n // 2, n // 2
    k = (n - 1).bit_length() - 1
    return (1 << k) - 1, n + 1 - (1 << k)


def _optimal_split(log_degree):
    s = log_degree >> 1
    a = (1 << s) + (1 << (log_degree - s)) + log_degree - s - 3
    b = (1 << (s + 1)) + (1 << (log_degree - s - 1)) + log_degree - s - 4
    return s + 1 if a > b else s


def chebyshev_eval_products(degree: int, even: bool, odd: bool) -> Tuple[Dict[int, int], int, int]:
    """Depth (in rescalings below the input) at which each power T_n is produced by PowerBasis.GenPower, the number of giant-step
    products (EvaluateMonomial: Mul + Relinearize + Rescale) and the total depth of the evaluation."""
    log_degree = degree.bit_length()
    log_split = _optimal_split(log_degree)
    depth_of: Dict[int, int] = {1: 0}

    def gen(n):
        if n in depth_of:
            return depth_of[n]
        a, b = _split_degree(n)
        d = max(gen(a), gen(b)) + 1
        c = abs(a - b)
        if c:
            gen(c)                       # Chebyshev: T_n = 2 T_a T_b - T_|a-b|
        depth_of[n] = d
        return d
    gen(1 << (log_degree - 1))
    for i in range((1 << log_split) - 1, 2, -1):
        if not (even or odd) or (i & 1 == 0 and even) or (i & 1 == 1 and odd):
            gen(i)
    split = 1 << (log_degree - log_split)        # baby-step polynomials; merged pairwise by EvaluateGiantStep
    giant = split - 1
    return depth_of, giant, log_degree


# ---- the trace ----------------------------------------------------------------------------------------------------------------
def bootstrap_trace(logN=16, log_slots=None, residual_limbs=14, stc_depth=3, evalmod_limbs=9, cts_depth=4, mod1_degree=30, double_angle=3,
                    log_bsgs_ratio=1) -> List[dict]:
    """Ops of one Evaluator.bootstrap for a default literal (circuits/ckks/bootstrapping/default_parameters.go): levels follow
    parameters.go:128-210 (SlotsToCoeffs starts at residual_max + stc_depth, EvalMod above it, CoeffsToSlots on top)."""
    if log_slots is None:
        log_slots = logN - 1
    slots = 1 << log_slots
    top = residual_limbs - 1 + stc_depth + evalmod_limbs + cts_depth          # MaxLevel of the bootstrapping parameters
    ops: List[dict] = []
    # ModUp (:616-771)
    ops.append({"phase": "ModUp", "op": "keyswitch", "level": 0, "note": "ApplyEvaluationKey(EvkDenseToSparse) at level 0"})
    ops.append({"phase": "ModUp", "op": "intt", "level": 0, "polys": 2})
    ops.append({"phase": "ModUp", "op": "modup_centered", "level": top, "note": "q -> Q for c0, q -> QP for c1 (coefficient-wise BRedAdd)"})
    ops.append({"phase": "ModUp", "op": "ntt_qp_per_digit", "level": top, "note": "ringQ.NTT / ringP.NTT of the extended c1 into every digit buffer (:691-697)"})
    ops.append({"phase": "ModUp", "op": "ntt", "level": top, "polys": 1})
    ops.append({"phase": "ModUp", "op": "mulscalar", "level": top, "polys": 1})
    ops.append({"phase": "ModUp", "op": "gadget_product_hoisted", "level": top, "note": "EvkSparseToDense (:722-724) + Add"})
    # CoeffsToSlots (dft.go:240-300)
    level = top
    for diags in dft_index_map(logN, log_slots, cts_depth, ENCODE):
        n1 = find_best_bsgs_ratio(diags, slots, log_bsgs_ratio)
        ops.append({"phase": "CoeffsToSlots", "op": "lintrans", "level": level, "diags": diags, "N1": n1})
        ops.append({"phase": "CoeffsToSlots", "op": "rescale", "level": level})
        level -= 1
    ops.append({"phase": "CoeffsToSlots", "op": "conjugate", "level": level, "note": "eval.Conjugate: one key switch (galEl = 2N - 1)"})
    ops.append({"phase": "CoeffsToSlots", "op": "add", "level": level, "count": 2, "note": "Sub / Add for the real and imaginary parts"})
    ops.append({"phase": "CoeffsToSlots", "op": "mul_by_i", "level": level})
    # EvalMod on ctReal and ctImag (mod1_evaluator.go:31-143)
    depth_of, giant, log_degree = chebyshev_eval_products(mod1_degree, even=True, odd=False)
    for which in ("real", "imag"):
        l0 = level
        ops.append({"phase": "EvalMod", "op": "add_const", "level": l0, "ct": which})
        for n, d in sorted(depth_of.items(), key=lambda kv: (kv[1], kv[0])):
            if n == 1:
                continue
            # T_n = 2 T_a T_b - T_|a-b|: MulRelinNew at the level of its deeper operand, Rescale when the power is consumed
            ops.append({"phase": "EvalMod", "op": "mulrelin_rescale", "level": l0 - (d - 1), "ct": which, "note": "power basis T_%d" % n})
        lv = l0 - max(depth_of.values())
        for g in range(giant):
            # EvaluateMonomial: Relinearize + Rescale of the odd part, Mul by X^{2^k}, Add (:162-190); the merges of one round sit on the
            # same level, one level lower per round
            rnd = int(math.floor(math.log2(g + 1))) if giant > 1 else 0
            ops.append({"phase": "EvalMod", "op": "mulrelin_rescale", "level": max(lv - (int(math.log2(giant + 1)) - 1 - rnd), 1), "ct": which,
                        "note": "Paterson-Stockmeyer giant step"})
        lv = l0 - log_degree
        for _ in range(double_angle):
            ops.append({"phase": "EvalMod", "op": "mulrelin_rescale", "level": lv, "ct": which, "note": "double angle"})
            ops.append({"phase": "EvalMod", "op": "add", "level": lv, "count": 2, "ct": which})
            lv -= 1
    level = level - evalmod_limbs
    # SlotsToCoeffs (dft.go:318-341)
    ops.append({"phase": "SlotsToCoeffs", "op": "mul_by_i", "level": level})
    ops.append({"phase": "SlotsToCoeffs", "op": "add", "level": level, "count": 1})
    for diags in dft_index_map(logN, log_slots, stc_depth, DECODE):
        n1 = find_best_bsgs_ratio(diags, slots, log_bsgs_ratio)
        ops.append({"phase": "SlotsToCoeffs", "op": "lintrans", "level": level, "diags": diags, "N1": n1})
        ops.append({"phase": "SlotsToCoeffs", "op": "rescale", "level": level})
        level -= 1
    return ops


def summarize(ops: List[dict]) -> dict:
    out: Dict[str, dict] = {}
    for o in ops:
        ph = out.setdefault(o["phase"], {})
        ph[o["op"]] = ph.get(o["op"], 0) + 1
        if o["op"] == "lintrans":
            _, r1, r2 = bsgs_index(o["diags"], 1 << 30, o["N1"])
            ph.setdefault("lintrans_shapes", []).append({"level": o["level"], "diagonals": len(o["diags"]), "N1": o["N1"],
                                                         "baby_rotations": len([r for r in r2 if r]), "giant_rotations": len([r for r in r1 if r])})
    return out
