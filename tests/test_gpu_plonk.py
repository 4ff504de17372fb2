"""Prover rows: the device-resident TurboPlonk prover behind jf_plonk_preprocess /
jf_plonk_prove against the CPU restatement (oracle/plonk_ref.py) on the same circuit, witness,
masking scalars and transcript: identical verifying-key commitments and byte-identical serialized
proofs, which the restated jellyfish verifier accepts.  At sizes the Python prover cannot reach the
proof is checked by that verifier alone (size-independent)."""
import os
import random
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

BETA = 0x1D3C7A5B9E8F60412B7A6C5D4E3F20198A7B6C5D4E3F2A1B0C9D8E7F6A5B4C3


@pytest.fixture(scope="module")
def P():
    import plonk_ref
    return plonk_ref


def _blinders(co, fr, seed):
    rnd = random.Random(seed)
    ints = [rnd.randrange(fr.p) for _ in range(17)]
    return ints, co.ints_to_limbs([fr.to_mont(v) for v in ints], 4)


def _tiny(P, adds):
    cs = P.PlonkCircuit()
    a = cs.create_public_variable(5) if adds == 1 else cs.create_variable(5)
    for _ in range(adds):
        a = cs.add(a, cs.one())
    cs.finalize_for_arithmetization()
    return cs


CIRCUITS = {
    "tiny_n4": lambda P: _tiny(P, 1),      # 2 constant gates + io gate + 1 addition -> n = 4, quotient domain 32
    "tiny_n8": lambda P: _tiny(P, 4),
    "test_m2": lambda P: P.gen_circuit_for_test(2, 3),
    "bench_64": lambda P: P.gen_circuit_for_bench(64),
    "test_m20": lambda P: P.gen_circuit_for_test(20, 1),
    "bench_2^10": lambda P: P.gen_circuit_for_bench(1 << 10),
    "test_m300": lambda P: P.gen_circuit_for_test(300, 7),
    "all_selectors_m9": lambda P: P.gen_circuit_all_selectors(9),      # every selector column non-zero
    "all_selectors_m100": lambda P: P.gen_circuit_all_selectors(100),
}


@pytest.mark.parametrize("name", list(CIRCUITS))
def test_proof_bytes_match_the_cpu_restatement(ctx, co, py, P, name):
    import mpc_jellyfish_b200 as jf
    import plonk_util as U
    cv, fr = py.BN254, py.BN254_FR
    cs = CIRCUITS[name](P)
    n = cs.n
    beta = BETA % fr.p
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing("bn254", beta, n + 3)
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                     arr["pub_gate_ids"])
    opk = P.preprocess(cv, P.gen_srs(cv, beta, n + 2), cs)
    vk = U.vk_from_product(co, cv, pk, cs.k)
    assert vk["selector_comms"] == opk["vk"]["selector_comms"]
    assert vk["sigma_comms"] == opk["vk"]["sigma_comms"]
    for kind in ("solidity", "standard"):
        ints, bl = _blinders(co, fr, 11)
        proof = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, kind)
        want = P.prove(cv, cs, opk, ints, kind)
        got = U.proof_to_oracle(co, cv, proof)
        for f in ("wires_poly_comms", "prod_perm_poly_comm", "split_quot_poly_comms", "wires_evals", "wire_sigma_evals",
                  "perm_next_eval", "opening_proof", "shifted_opening_proof"):
            assert got[f] == want[f], (kind, f)
        assert proof.serialize_compressed() == P.serialize_proof(cv, want)
        assert P.verify(cv, opk["vk"], cs.public_input(), got, beta, kind)
        ch = [fr.from_mont(v) for v in co.limbs_to_ints(proof.challenges)]
        assert ch == [want["challenges"][c] for c in ("beta", "gamma", "alpha", "zeta", "v")]
    # extra transcript message (snark.rs:263-266) changes every challenge, on both sides alike
    ints, bl = _blinders(co, fr, 12)
    p2 = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity", b"extra transcript init msg")
    w2 = P.prove(cv, cs, opk, ints, "solidity", b"extra transcript init msg")
    assert p2.serialize_compressed() == P.serialize_proof(cv, w2)
    pk.free()
    key.free()


@pytest.mark.parametrize("name,kind", [("test_m2", "standard"), ("test_m20", "solidity")])
def test_bls12_381_proofs_match_the_cpu_restatement(ctx, co, py, P, name, kind):
    """The same prover over BLS12-381 (48-byte compressed points, 255-bit Fr, GENERATOR 7, its own k)."""
    import mpc_jellyfish_b200 as jf
    import plonk_util as U
    cv, fr = py.BLS12_381, py.BLS12_381_FR
    cs = {"test_m2": lambda: P.gen_circuit_for_test(2, 3, fr), "test_m20": lambda: P.gen_circuit_for_test(20, 1, fr)}[name]()
    beta = BETA % fr.p
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing("bls12_381", beta, cs.n + 3)
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                     arr["pub_gate_ids"], skip_zero_selectors=True)
    opk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    vk = U.vk_from_product(co, cv, pk, cs.k)
    assert vk["selector_comms"] == opk["vk"]["selector_comms"] and vk["sigma_comms"] == opk["vk"]["sigma_comms"]
    ints, bl = _blinders(co, fr, 21)
    proof = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, kind)
    want = P.prove(cv, cs, opk, ints, kind)
    ser = proof.serialize_compressed()
    assert len(ser) == 8 * 4 + 48 * 13 + 32 * 10 + 1 and ser == P.serialize_proof(cv, want)
    assert P.verify(cv, opk["vk"], cs.public_input(), U.proof_to_oracle(co, cv, proof), beta, kind)
    pk.free()
    key.free()


def test_cached_coset_evaluations_and_zero_selector_skip_give_the_same_proof(ctx, co, py, P):
    import mpc_jellyfish_b200 as jf
    import plonk_util as U
    fr = py.BN254_FR
    cs = P.gen_circuit_for_test(40, 2)
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing("bn254", BETA % fr.p, cs.n + 3)
    _, bl = _blinders(co, fr, 5)
    outs = []
    for cache, skip, full in ((False, False, False), (True, False, False), (False, True, False), (True, True, False),
                              (False, False, True), (True, True, True)):
        # full: all 8n points of the quotient coset in one transform per polynomial, as the reference does
        # (prover.rs:552-567), instead of six sub-cosets of n points
        pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                         arr["pub_gate_ids"], cache_coset_evals=cache, skip_zero_selectors=skip,
                                         full_quotient_coset=full)
        outs.append(jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "standard").serialize_compressed())
        # the key is reusable: a second proof with other masks differs but has the same evaluations' count
        outs.append(jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "standard").serialize_compressed())
        pk.free()
    assert len(set(outs)) == 1 and len(outs) == 12
    key.free()


def test_unsatisfied_witness_is_rejected_like_the_reference(ctx, co, py, P):
    """prover.rs:916-919 WrongQuotientPolyDegree; the context stays usable afterwards."""
    import mpc_jellyfish_b200 as jf
    import plonk_util as U
    fr = py.BN254_FR
    cs = P.gen_circuit_for_test(2, 3)
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing("bn254", 77, cs.n + 3)
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                     arr["pub_gate_ids"])
    _, bl = _blinders(co, fr, 1)
    bad = arr["witness"].copy()
    bad[5] = co.ints_to_limbs([fr.to_mont(cs.witness[5] + 1)], 4)[0]
    with pytest.raises(jf.WrongQuotientPolyDegree):
        jf.PlonkKzgSnark.prove(pk, bad, bl, "solidity")
    good = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity")
    opk = P.preprocess(py.BN254, P.gen_srs(py.BN254, 77, cs.n + 2), cs)
    assert P.verify(py.BN254, opk["vk"], cs.public_input(), U.proof_to_oracle(co, py.BN254, good), 77, "solidity")
    with pytest.raises(jf.InvalidParameters):  # commit key too short for n + 3 coefficients
        short = ctx.generate_srs_for_testing("bn254", 77, cs.n)
        jf.PlonkKzgSnark.preprocess(ctx, short, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                    arr["pub_gate_ids"])
    pk.free()
    key.free()


@pytest.mark.parametrize("adds,form", [(4, "8n coset (n = 8)"), (12, "six sub-cosets (n = 16, the smallest)"), (20, "six sub-cosets (n = 32)")])
def test_unsatisfied_witness_is_rejected_at_the_boundary_between_the_two_round_3_forms(ctx, co, py, P, adds, form):
    """n = 8 keeps the reference's 8n coset: six rows would hold exactly the quotient's 48 coefficients and leave nothing
    for the degree check.  From n = 16 on the six-row interpolant has n - 8 coefficients above the degree to betray a
    witness that does not satisfy the circuit (tests/test_subcoset_identity.py states the same in big-int terms)."""
    import mpc_jellyfish_b200 as jf
    import plonk_util as U
    fr = py.BN254_FR
    cs = P.PlonkCircuit()
    a = cs.create_variable(5)
    for _ in range(adds):
        a = cs.add(a, cs.one())
    cs.finalize_for_arithmetization()
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing("bn254", 99, cs.n + 3)
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                     arr["pub_gate_ids"])
    ints, bl = _blinders(co, fr, 8)
    for victim in (len(cs.witness) - 1, 2):  # the last sum (one gate breaks) / the input variable (its gate and the copy constraints)
        bad = arr["witness"].copy()
        bad[victim] = co.ints_to_limbs([fr.to_mont((cs.witness[victim] + 7) % fr.p)], 4)[0]
        with pytest.raises(jf.WrongQuotientPolyDegree):
            jf.PlonkKzgSnark.prove(pk, bad, bl, "solidity")
    good = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "standard")
    opk = P.preprocess(py.BN254, P.gen_srs(py.BN254, 99, cs.n + 2), cs)
    assert good.serialize_compressed() == P.serialize_proof(py.BN254, P.prove(py.BN254, cs, opk, ints, "standard")), form
    pk.free()
    key.free()


def test_numpy_circuit_builder_matches_the_restated_circuit(ctx, co, py, P):
    import bench_circuit as B
    import plonk_util as U
    cs = P.gen_circuit_for_bench(1 << 8)
    want = U.arrays_from_oracle_circuit(co, py, cs)
    got = B.bench_circuit_arrays(ctx, 8)
    assert cs.k == B.BN254_K
    for f in ("selectors", "sigmas", "k", "witness"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(got["wire_vars"], want["wire_vars"]) and got["num_vars"] == want["num_vars"]


@pytest.mark.parametrize("log_n,kind", [(14, "solidity"), (16, "standard"), (18, "solidity"), (20, "standard")])
def test_large_proofs_are_accepted_by_the_restated_verifier(ctx, co, py, P, log_n, kind):
    """BASELINE config 4 shape (the bench circuit); the verifier check is size independent."""
    import mpc_jellyfish_b200 as jf
    import bench_circuit as B
    import plonk_util as U
    cv, fr = py.BN254, py.BN254_FR
    arr = B.bench_circuit_arrays(ctx, log_n)
    beta = BETA % fr.p
    key = ctx.generate_srs_for_testing("bn254", beta, arr["n"] + 3)
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [])
    _, bl = _blinders(co, fr, log_n)
    proof = jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, kind)
    got = U.proof_to_oracle(co, cv, proof)
    vk = U.vk_from_product(co, cv, pk, B.BN254_K)
    assert P.verify(cv, vk, [], got, beta, kind)
    tampered = dict(got)
    tampered["perm_next_eval"] = (got["perm_next_eval"] + 1) % fr.p
    assert not P.verify(cv, vk, [], tampered, beta, kind)
    pk.free()
    key.free()
