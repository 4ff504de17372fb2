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


def test_product_serializer_emits_the_published_g1_encodings(py):
    """`jf_plonk_proof_serialize` is host code.  BN254 points use ark-ec's generic compressed form, BLS12-381 points the
    ZCash / IETF form ark-bls12-381 0.4.0 emits; both are checked against the PUBLISHED encodings of G, 2 G and the identity
    (tests/golden/constants.json), not against the Python restatement."""
    import ctypes
    from mpc_jellyfish_b200 import _ffi
    consts = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "constants.json")))
    for curve_id, cv, key, L in ((0, py.BN254, "bn254_g1", 4), (1, py.BLS12_381, "bls12_381_g1", 6)):
        g = consts[key]
        pr = _ffi.PlonkProofStruct()
        pr.curve = curve_id
        fq = cv.fq

        def put(dst, slot, P):
            x, y = P
            for k, v in enumerate((fq.to_mont(x), fq.to_mont(y))):
                for i in range(L):
                    dst[2 * L * slot + L * k + i] = (v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF

        G, G2 = cv.gen, cv.mul(2, cv.gen)
        put(pr.wires_poly_comms, 0, G)
        put(pr.wires_poly_comms, 1, G2)
        put(pr.wires_poly_comms, 2, cv.neg(G))
        pr.wires_inf[3] = 1
        put(pr.wires_poly_comms, 4, G)
        for name in ("prod_perm_poly_comm", "opening_proof", "shifted_opening_proof"):
            put(getattr(pr, name), 0, G)
        for j in range(5):
            put(pr.split_quot_poly_comms, j, G)
        buf = ctypes.create_string_buffer(2048)
        n = _ffi.lib().jf_plonk_proof_serialize(ctypes.byref(pr), buf, len(buf))
        nb = 8 * L
        assert n == 8 + 5 * nb + nb + 8 + 5 * nb + 2 * nb + 8 + 5 * 32 + 8 + 4 * 32 + 32 + 1
        raw = buf.raw[:n]
        pts = [raw[8 + nb * i: 8 + nb * (i + 1)].hex() for i in range(5)]
        assert pts[0] == g["generator_compressed"] and pts[4] == g["generator_compressed"]
        assert pts[3] == g["identity_compressed"]
        if key == "bls12_381_g1":
            assert pts[1] == g["two_g_compressed"]
            assert pts[2] == "b7" + g["generator_compressed"][2:]
        else:
            assert pts[2] == g["generator_compressed"][:-2] + "80"   # -G: same x, "negative" flag in the last byte
        assert pts[1] == cv.serialize_compressed(G2).hex() and pts[2] == cv.serialize_compressed(cv.neg(G)).hex()


def test_link_proof_serializer_and_linking_challenge_on_the_host(py):
    """`jf_link_proof_serialize` and the collaborative mirror's `MultiproverLinking.challenge` are host code: `LinkingProof` bytes
    are the two compressed points in order (published encodings of G / identity), and the challenge over three points equals the
    restated SolidityTranscript's (proof_linking.rs:171-191)."""
    import ctypes
    import numpy as np
    import plonk_ref as P
    from mpc_jellyfish_b200 import _ffi
    from mpc_jellyfish_b200.multiprover import MultiproverLinking, g1_serialize_compressed
    consts = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "constants.json")))
    for curve_id, name, cv, key, L in ((0, "bn254", py.BN254, "bn254_g1", 4), (1, "bls12_381", py.BLS12_381, "bls12_381_g1", 6)):
        g = consts[key]
        fq = cv.fq

        def limbs(Pt):
            out = []
            for v in (fq.to_mont(Pt[0]), fq.to_mont(Pt[1])):
                out += [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(L)]
            return out

        raw = _ffi.LinkProofStruct()
        raw.curve = curve_id
        for i, v in enumerate(limbs(cv.gen)):
            raw.quotient_commitment[i] = v
        raw.opening_inf = 1
        buf = ctypes.create_string_buffer(128)
        n = _ffi.lib().jf_link_proof_serialize(ctypes.byref(raw), buf, len(buf))
        assert n == 16 * L
        assert buf.raw[:8 * L].hex() == g["generator_compressed"] and buf.raw[8 * L:n].hex() == g["identity_compressed"]
        assert _ffi.lib().jf_link_proof_serialize(ctypes.byref(raw), buf, 16 * L - 1) < 0
        # the three-point challenge
        pts = [cv.gen, cv.mul(2, cv.gen), None]
        xy = [(np.array(limbs(Q), dtype=np.uint64), False) if Q is not None else (np.zeros(2 * L, dtype=np.uint64), True) for Q in pts]
        assert g1_serialize_compressed(name, *xy[1]) == P.ser_g1(cv, pts[1])
        eta = MultiproverLinking.challenge(name, xy[0], xy[1], xy[2])
        want = P._link_challenge(cv, pts[0], pts[1], pts[2], "solidity")
        got = sum(int(eta[i]) << (64 * i) for i in range(4))
        assert cv.fr.from_mont(got) == want
