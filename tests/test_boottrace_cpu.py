"""The bootstrapping op trace (lattigo_b200/boottrace.py) against facts stated in the reference's own sources: the rotation sets of the
factorised DFT (dft.go's worked example in the comment of computeBootstrappingDFTIndexMap: depth 2 over 2^8 slots uses 127 + 128 ... the
merge order), the default literal's level budget (default_parameters.go:118-134) and the BSGS split."""
from lattigo_b200 import boottrace as BT


def test_default_trace_level_budget_and_shapes():
    ops = BT.bootstrap_trace()
    top = 14 - 1 + 3 + 9 + 4
    assert top == 29
    lts = [o for o in ops if o["op"] == "lintrans"]
    assert [o["level"] for o in lts] == [29, 28, 27, 26, 16, 15, 14]          # CoeffsToSlots on top, SlotsToCoeffs right above the residual chain
    assert all(o["level"] >= 14 for o in ops if o["phase"] != "ModUp")
    evalmod = [o for o in ops if o["phase"] == "EvalMod" and o["op"] == "mulrelin_rescale"]
    assert len(evalmod) == 2 * 12                                            # 6 power-basis + 3 giant-step + 3 double-angle products per ciphertext
    assert min(o["level"] for o in evalmod) >= 17 and max(o["level"] for o in evalmod) == 25
    s = BT.summarize(ops)
    for sh in s["CoeffsToSlots"]["lintrans_shapes"] + s["SlotsToCoeffs"]["lintrans_shapes"]:
        assert sh["baby_rotations"] + sh["giant_rotations"] < sh["diagonals"]    # BSGS needs fewer rotations than diagonals


def test_dft_index_maps():
    # one level of the (non bit-reversed) encoding FFT has the three diagonals {0, +-2^(level-1)}
    assert BT.dft_index_map(16, 15, 15, BT.ENCODE)[0] == sorted({0, 1 << 14, (1 << 15) - (1 << 14)})
    # merging all levels of a 2^3-slot transform gives the dense matrix
    assert BT.dft_index_map(4, 3, 1, BT.ENCODE)[0] == list(range(8))
    # every factor of the default CoeffsToSlots is a stride-2^k comb: at most 2^(merged levels + 1) - 1 diagonals
    for m, depth in zip(BT.dft_index_map(16, 15, 4, BT.ENCODE), (4, 4, 4, 3)):
        assert len(m) <= (2 << depth) - 1 and 0 in m
    # sparse packing: the decode side starts with the repacking matrix (rotation by the slot count)
    assert (1 << 10) in BT.dft_index_map(16, 10, 3, BT.DECODE)[0] or any((1 << 10) & d for d in BT.dft_index_map(16, 10, 3, BT.DECODE)[0])


def test_bsgs_helpers_match_lintrans_oracle():
    from oracle import lintrans as LT
    diags = [0, 1, 2, 3, 15, 16, 17, 31]
    assert BT.bsgs_index(diags, 32, 4) == LT.bsgs_index(diags, 32, 4)
    assert BT.find_best_bsgs_ratio(list(range(32)), 1 << 15, 1) in (4, 8)
