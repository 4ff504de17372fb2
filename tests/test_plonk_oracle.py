"""CPU suite: pin the prover oracle (oracle/plonk_ref.py).  Keccak-256 against the reference's own
KAT, Merlin / ChaCha20 / coset representatives against published vectors (tests/golden/
transcript_vectors.json), then prover <-> verifier consistency: the prover restated from prover.rs /
snark.rs must satisfy the verifier restated from verifier.rs (known-beta G1 form of the pairing check)."""
import json
import os
import random
import struct

import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "transcript_vectors.json")))


@pytest.fixture(scope="module")
def P():
    import plonk_ref
    return plonk_ref


def test_keccak256_reference_kat(P):
    g = GOLD["keccak256"]
    assert P.keccak256(g["message"].encode()).hex() == g["digest"]
    assert P.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"


def test_merlin_published_vector(P):
    g = GOLD["merlin"]
    t = P.MerlinTranscript(g["protocol"].encode())
    t.append_message(g["label"].encode(), g["data"].encode())
    assert t.challenge_bytes(g["challenge_label"].encode(), 32).hex() == g["challenge"]


def test_chacha20_and_coset_representatives(P, py):
    assert struct.pack("<16I", *P.chacha20_block([0] * 8, 0)).hex() == GOLD["chacha20_zero_key_block0"]
    want = [int(x, 16) for x in GOLD["coset_k_bn254"]]
    for n in (1 << 5, 1 << 10, 1 << 20):
        assert P.compute_coset_representatives(py.BN254_FR, 5, n) == want


def test_solidity_transcript_follows_the_code_not_the_doc(P, py):
    # state <- keccak(state|transcript|0) || keccak(state|transcript|1); the transcript is NOT cleared
    t = P.SolidityTranscript()
    t.append_message(b"ignored", b"abc")
    c1 = t.get_and_append_challenge(py.BN254_FR, b"x")
    s1 = P.keccak256(bytes(64) + b"abc\x00") + P.keccak256(bytes(64) + b"abc\x01")
    assert c1 == int.from_bytes(s1[:48], "little") % py.BN254_FR.p
    t.append_message(b"ignored", b"de")
    c2 = t.get_and_append_challenge(py.BN254_FR, b"y")
    s2 = P.keccak256(s1 + b"abcde\x00") + P.keccak256(s1 + b"abcde\x01")
    assert c2 == int.from_bytes(s2[:48], "little") % py.BN254_FR.p


@pytest.mark.parametrize("kind", ["solidity", "standard"])
@pytest.mark.parametrize("which", ["test_m2", "bench_64", "test_m20", "all_sel"])
def test_oracle_prover_satisfies_oracle_verifier(P, py, kind, which):
    cv = py.BN254
    cs = {"test_m2": lambda: P.gen_circuit_for_test(2, 3), "bench_64": lambda: P.gen_circuit_for_bench(64),
          "test_m20": lambda: P.gen_circuit_for_test(20, 1), "all_sel": lambda: P.gen_circuit_all_selectors(5)}[which]()
    if which == "all_sel":
        assert all(any(col) for col in cs.selector_evals()), "every selector column must be exercised"
    assert cs.check_satisfiability()
    beta = 0x1234567890ABCDEF1234567890ABCDEF % cv.fr.p
    pk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    rnd = random.Random(3)
    bl = [rnd.randrange(cv.fr.p) for _ in range(17)]
    proof = P.prove(cv, cs, pk, bl, kind)
    assert P.verify(cv, pk["vk"], cs.public_input(), proof, beta, kind)
    assert len(P.serialize_proof(cv, proof)) == 8 * 4 + 32 * 13 + 32 * 10 + 1
    # soundness smoke: any tampering is rejected, and so is the wrong transcript or public input
    bad = dict(proof)
    bad["wires_evals"] = [proof["wires_evals"][0] ^ 1] + proof["wires_evals"][1:]
    assert not P.verify(cv, pk["vk"], cs.public_input(), bad, beta, kind)
    other = "standard" if kind == "solidity" else "solidity"
    assert not P.verify(cv, pk["vk"], cs.public_input(), proof, beta, other)
    if cs.num_inputs():
        pi = cs.public_input()
        pi[0] = (pi[0] + 1) % cv.fr.p
        assert not P.verify(cv, pk["vk"], pi, proof, beta, kind)
    # different blinders, different proof, still accepted (zero-knowledge masking is live)
    proof2 = P.prove(cv, cs, pk, [b + 1 for b in bl], kind)
    assert proof2["wires_poly_comms"] != proof["wires_poly_comms"]
    assert P.verify(cv, pk["vk"], cs.public_input(), proof2, beta, kind)


def test_unsatisfied_witness_fails_quotient_degree(P, py):
    cv = py.BN254
    cs = P.gen_circuit_for_test(2, 3)
    pk = P.preprocess(cv, P.gen_srs(cv, 77, cs.n + 2), cs)
    cs.witness[5] = (cs.witness[5] + 1) % cv.fr.p  # break a gate
    assert not cs.check_satisfiability()
    with pytest.raises(ValueError, match="WrongQuotientPolyDegree"):
        P.prove(cv, cs, pk, list(range(1, 18)), "solidity")


@pytest.mark.parametrize("kind", ["solidity", "standard"])
@pytest.mark.parametrize("which", ["ultra_test_m2", "ultra_test_m6", "ultra_bench_100"])
def test_ultraplonk_oracle_prover_satisfies_oracle_verifier(P, py, kind, which):
    """UltraPlonk (Plookup): prover restated from prover.rs:98-190,239-297,773-888,1037-1113 and constraint_system.rs:1261-1492,
    verifier from verifier.rs:340-418,591-650,707-745 -- written from different reference files, they must agree."""
    cv = py.BN254
    cs = {"ultra_test_m2": lambda: P.gen_circuit_for_test(2, 3, ultra=True), "ultra_test_m6": lambda: P.gen_circuit_for_test(6, 1, ultra=True),
          "ultra_bench_100": lambda: P.gen_circuit_for_bench(100, ultra=True)}[which]()
    assert cs.check_satisfiability() and cs.nw == 6 and len(cs.selector_evals()) == 14
    beta = 0xABCDEF0123456789 % cv.fr.p
    pk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    rnd = random.Random(4)
    bl = [rnd.randrange(cv.fr.p) for _ in range(P.num_blinders(cs))]
    assert len(bl) == 29
    proof = P.prove(cv, cs, pk, bl, kind)
    assert P.verify(cv, pk["vk"], cs.public_input(), proof, beta, kind)
    assert len(P.serialize_proof(cv, proof)) == 1481
    # tampering with any Plookup evaluation, the lookup product commitment or a sorted-vector commitment is rejected
    for field in P.PLOOKUP_EVAL_FIELDS:
        lp = dict(proof["plookup_proof"])
        pe = dict(lp["poly_evals"])
        pe[field] = (pe[field] + 1) % cv.fr.p
        lp["poly_evals"] = pe
        assert not P.verify(cv, pk["vk"], cs.public_input(), dict(proof, plookup_proof=lp), beta, kind), field
    lp = dict(proof["plookup_proof"], prod_lookup_poly_comm=cv.gen)
    assert not P.verify(cv, pk["vk"], cs.public_input(), dict(proof, plookup_proof=lp), beta, kind)
    lp = dict(proof["plookup_proof"], h_poly_comms=[proof["plookup_proof"]["h_poly_comms"][1], proof["plookup_proof"]["h_poly_comms"][0]])
    assert not P.verify(cv, pk["vk"], cs.public_input(), dict(proof, plookup_proof=lp), beta, kind)
    # a TurboPlonk proof against an UltraPlonk key (and vice versa) is refused outright
    assert not P.verify(cv, pk["vk"], cs.public_input(), dict(proof, plookup_proof=None), beta, kind)


def test_ultraplonk_lookup_outside_the_table_is_refused(P, py):
    cv = py.BN254
    cs = P.gen_circuit_for_test(2, 3, ultra=True)
    pk = P.preprocess(cv, P.gen_srs(cv, 31337, cs.n + 2), cs)
    cs.witness[cs.wire_variables[5][0]] = 1000   # range-checked variable outside [0, 32)
    assert not cs.check_satisfiability()
    with pytest.raises(ValueError, match="sorted vector has wrong length"):
        P.prove(cv, cs, pk, list(range(1, 30)), "solidity")


@pytest.mark.parametrize("kind", ["solidity", "standard"])
@pytest.mark.parametrize("ultra", [False, True], ids=["turbo", "ultra"])
def test_batch_prove_several_instances(P, py, kind, ultra):
    """`batch_prove` (snark.rs:201-469): three instances over one domain, one transcript, ONE quotient (alpha bases 1, a, a^2 with
    a = alpha^3 | alpha^7) and one pair of opening proofs; the batch verifier restated from verifier.rs:68-254 accepts."""
    cv = py.BN254
    css = [P.gen_circuit_for_test(3, 2, ultra=ultra), P.gen_circuit_for_test(3, 1, ultra=ultra),
           P.gen_circuit_for_test(5, 3, ultra=ultra) if ultra else P.gen_circuit_for_bench(32)]
    assert len({cs.n for cs in css}) == 1
    beta = 0x7777777777777777777 % cv.fr.p
    srs = P.gen_srs(cv, beta, css[0].n + 2)
    pks = [P.preprocess(cv, srs, cs) for cs in css]
    rnd = random.Random(6)
    bl = [rnd.randrange(cv.fr.p) for _ in range(P.batch_num_blinders(css))]
    bp = P.batch_prove(cv, css, pks, bl, kind)
    vks, pis = [pk["vk"] for pk in pks], [cs.public_input() for cs in css]
    assert P.batch_verify(cv, vks, pis, bp, beta, kind)
    # instance order matters; every instance's evaluations are bound
    assert not P.batch_verify(cv, [vks[1], vks[0], vks[2]], [pis[1], pis[0], pis[2]], bp, beta, kind)
    for i in range(3):
        pe = [dict(x) for x in bp["poly_evals_vec"]]
        pe[i]["wires_evals"] = [(pe[i]["wires_evals"][0] + 1) % cv.fr.p] + pe[i]["wires_evals"][1:]
        assert not P.batch_verify(cv, vks, pis, dict(bp, poly_evals_vec=pe), beta, kind)
    # a batch of one is `prove` (structs.rs:302-331)
    one = P.batch_prove(cv, css[:1], pks[:1], bl[:2 * css[0].nw] + bl[2 * css[0].nw * 3:][:0] + [7] * (P.batch_num_blinders(css[:1]) - 2 * css[0].nw), kind)
    single = P.prove(cv, css[0], pks[0], bl[:2 * css[0].nw] + [7] * (P.num_blinders(css[0]) - 2 * css[0].nw), kind)
    assert single["wires_poly_comms"] == one["wires_poly_comms_vec"][0] and single["opening_proof"] == one["opening_proof"]
    assert P.verify(cv, vks[0], pis[0], single, beta, kind)
    # instances over different domains are refused (snark.rs:228-236)
    other = P.gen_circuit_for_bench(200, ultra=ultra)
    with pytest.raises(ValueError, match="domain size"):
        P.batch_prove(cv, [css[0], other], [pks[0], P.preprocess(cv, P.gen_srs(cv, beta, other.n + 2), other)],
                      [1] * P.batch_num_blinders([css[0], other]), kind)
