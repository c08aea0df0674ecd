"""The moduli literals shipped in lattigo_b200/params.py are what the reference's GenModuli produces for the
named parameter sets (core/rlwe/params.go:811-862 + ring/primes.go), regenerated here with the oracle."""
from lattigo_b200 import params as presets
from oracle import oracle as O


def test_presets_match_gen_moduli():
    for name, s in presets.PRESETS.items():
        q, p = O.gen_moduli(s["logN"] + 1, s["LogQ"], s["LogP"])
        assert q == s["Q"] and p == s["P"], name
        assert len(set(q + p)) == len(q) + len(p)


def test_qi60_literals():
    q61, _ = O.gen_moduli(18, [61] * 8, [])
    assert q61 == presets.QI60[:8]
