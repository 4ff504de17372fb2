"""CPU suite: the library's host-side transcript code (csrc/transcript.hpp: Keccak-256, Merlin/STROBE,
`SolidityTranscript`, `StandardTranscript`, `from_le_bytes_mod_order`) needs no GPU.  It is checked
against the reference's Keccak KAT, merlin's published vector and the Python restatement."""
import json
import os
import random

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "transcript_vectors.json")))


def test_keccak256_kat_and_random_lengths():
    import mpc_jellyfish_b200 as jf
    import plonk_ref as P
    g = GOLD["keccak256"]
    assert jf.keccak256(g["message"].encode()).hex() == g["digest"]
    rnd = random.Random(1)
    for n in [0, 1, 31, 32, 55, 135, 136, 137, 271, 272, 273, 1000]:
        msg = bytes(rnd.randrange(256) for _ in range(n))
        assert jf.keccak256(msg) == P.keccak256(msg), n


def test_merlin_vector_through_the_standard_transcript():
    """StandardTranscript::new(label) == merlin Transcript::new(label); the challenge is 64 PRF bytes
    reduced mod r, so the published 32-byte vector is checked through the Python restatement, which
    is pinned to it (tests/test_plonk_oracle.py), on identical call sequences."""
    import mpc_jellyfish_b200 as jf
    import plonk_ref as P
    import pyref
    for field, f in (("bn254_fr", pyref.BN254_FR), ("bls12_381_fr", pyref.BLS12_381_FR)):
        for kind, cls in (("standard", P.StandardTranscript), ("solidity", P.SolidityTranscript)):
            a = jf.Transcript(kind, b"PlonkProof")
            b = cls(b"PlonkProof")
            rnd = random.Random(7)
            for step in range(40):
                if rnd.random() < 0.6:
                    label = rnd.choice([b"witness_poly_comms", b"wire_evals", b"x"])
                    msg = bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 4, 8, 32, 32, 200, 700])))
                    a.append_message(label, msg)
                    b.append_message(label, msg)
                else:
                    label = rnd.choice([b"beta", b"gamma", b"alpha", b"zeta"])
                    got = a.get_and_append_challenge(field, label)
                    want = b.get_and_append_challenge(f, label)
                    got_int = sum(int(v) << (64 * i) for i, v in enumerate(got))
                    assert f.from_mont(got_int) == want, (field, kind, step)


def test_error_type_for_unsatisfied_witness_exists():
    import mpc_jellyfish_b200 as jf
    from mpc_jellyfish_b200 import errors, _ffi
    with pytest.raises(jf.WrongQuotientPolyDegree):
        errors.raise_for_status(_ffi.JF_ERR_QUOTIENT_DEGREE, "x")
    assert issubclass(jf.WrongQuotientPolyDegree, jf.PlonkError)
