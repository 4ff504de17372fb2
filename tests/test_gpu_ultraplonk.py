"""GPU suite: UltraPlonk (6 wire types, Plookup) through `jf_ultraplonk_preprocess` / `jf_ultraplonk_prove`, byte for byte against
the CPU restatement (oracle/plonk_ref.py, whose prover satisfies its restated verifier: tests/test_plonk_oracle.py) on the
reference's own UltraPlonk test circuit (plonk/src/proof_system/snark.rs:681-744: range gates, one key-value table, two lookups)
and on the bench circuit (plonk/benches/bench.rs:29-46 with `new_ultra_plonk(8)`), both transcripts, both curves."""
import os
import random
import sys

import numpy as np
import pytest

import plonk_util as U

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def P():
    import plonk_ref
    return plonk_ref


def _setup(ctx, co, py, P, cv, cs, beta, skip=False, full=False):
    import mpc_jellyfish_b200 as jf
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing(cv.name, beta, cs.n + 3)
    pk = jf.PlonkKzgSnark.preprocess_ultra(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                           arr["pub_gate_ids"], arr["range_bit_len"], arr["table_key"], arr["table_dom_sep"],
                                           arr["q_dom_sep"], skip_zero_selectors=skip, full_quotient_coset=full)
    return arr, key, pk


@pytest.mark.parametrize("kind", ["solidity", "standard"])
@pytest.mark.parametrize("which", ["test_m2", "test_m5", "bench_200", "test_m12", "bench_3000"])
def test_ultraplonk_proof_bytes_match_the_cpu_restatement(ctx, co, py, P, which, kind):
    import mpc_jellyfish_b200 as jf
    cv, fr = py.BN254, py.BN254_FR
    cs = {"test_m2": lambda: P.gen_circuit_for_test(2, 3, ultra=True), "test_m5": lambda: P.gen_circuit_for_test(5, 2, ultra=True),
          "bench_200": lambda: P.gen_circuit_for_bench(200, ultra=True), "test_m12": lambda: P.gen_circuit_for_test(12, 1, ultra=True), "bench_3000": lambda: P.gen_circuit_for_bench(3000, ultra=True)}[which]()
    assert cs.check_satisfiability()
    beta = 0x1234567890ABCDEF1234567890ABCDEF % fr.p
    arr, key, pk = _setup(ctx, co, py, P, cv, cs, beta)
    opk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    vk = U.vk_from_product(co, cv, pk, cs.k)
    assert vk["selector_comms"] == opk["vk"]["selector_comms"] and vk["sigma_comms"] == opk["vk"]["sigma_comms"]
    assert vk["plookup"] == opk["vk"]["plookup"]
    rnd = random.Random(11)
    ints = [rnd.randrange(fr.p) for _ in range(P.num_blinders(cs))]
    assert len(ints) == 29
    bl = co.ints_to_limbs([fr.to_mont(v) for v in ints], 4)
    proof = jf.PlonkKzgSnark.prove_ultra(pk, arr["witness"], bl, kind)
    want = P.prove(cv, cs, opk, ints, kind)
    got = U.proof_to_oracle(co, cv, proof)
    for name in ("wires_poly_comms", "prod_perm_poly_comm", "split_quot_poly_comms", "opening_proof", "shifted_opening_proof",
                 "wires_evals", "wire_sigma_evals", "perm_next_eval"):
        assert got[name] == want[name], name
    assert got["plookup_proof"] == want["plookup_proof"]
    ser = proof.serialize_compressed()
    assert ser == P.serialize_proof(cv, want) and len(ser) == 1481
    assert P.verify(cv, vk, cs.public_input(), got, beta, kind)
    # a second proof with other masks differs and still verifies; the extra transcript message changes the proof
    bl2 = co.ints_to_limbs([fr.to_mont((v + 1) % fr.p) for v in ints], 4)
    p2 = jf.PlonkKzgSnark.prove_ultra(pk, arr["witness"], bl2, kind)
    assert p2.serialize_compressed() != ser and P.verify(cv, vk, cs.public_input(), U.proof_to_oracle(co, cv, p2), beta, kind)
    p3 = jf.PlonkKzgSnark.prove_ultra(pk, arr["witness"], bl, kind, b"extra")
    assert p3.serialize_compressed() == P.serialize_proof(cv, P.prove(cv, cs, opk, ints, kind, b"extra"))
    pk.free()
    key.free()


def test_ultraplonk_bls12_381_and_zero_selector_skip(ctx, co, py, P):
    import mpc_jellyfish_b200 as jf
    cv, fr = py.BLS12_381, py.BLS12_381_FR
    cs = P.gen_circuit_for_test(3, 2, fr, ultra=True)
    beta = 0xFEEDFACE12345678 % fr.p
    opk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    rnd = random.Random(5)
    ints = [rnd.randrange(fr.p) for _ in range(29)]
    bl = co.ints_to_limbs([fr.to_mont(v) for v in ints], 4)
    want = P.serialize_proof(cv, P.prove(cv, cs, opk, ints, "solidity"))
    for skip, full in ((False, False), (True, False), (False, True), (True, True)):
        # seven sub-cosets of n points (default) and the reference's literal 8n coset give the same quotient polynomial
        arr, key, pk = _setup(ctx, co, py, P, cv, cs, beta, skip=skip, full=full)
        proof = jf.PlonkKzgSnark.prove_ultra(pk, arr["witness"], bl, "solidity")
        assert proof.serialize_compressed() == want, "skip=%s full=%s" % (skip, full)
        assert P.verify(cv, U.vk_from_product(co, cv, pk, cs.k), cs.public_input(), U.proof_to_oracle(co, cv, proof), beta, "solidity")
        pk.free()
        key.free()


def test_ultraplonk_errors(ctx, co, py, P):
    import mpc_jellyfish_b200 as jf
    cv, fr = py.BN254, py.BN254_FR
    cs = P.gen_circuit_for_test(2, 3, ultra=True)
    beta = 99991
    arr, key, pk = _setup(ctx, co, py, P, cv, cs, beta)
    bl = co.ints_to_limbs([fr.to_mont(7 + i) for i in range(29)], 4)
    # a range-checked variable outside [0, 32): its merged lookup value is not in the table
    bad = arr["witness"].copy()
    rv = cs.wire_variables[5][0]
    bad[rv] = co.ints_to_limbs([fr.to_mont(1000)], 4)[0]
    with pytest.raises(jf.InvalidParameters, match="sorted vector has wrong length"):
        jf.PlonkKzgSnark.prove_ultra(pk, bad, bl)
    # the key still works afterwards
    good = jf.PlonkKzgSnark.prove_ultra(pk, arr["witness"], bl)
    opk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    assert good.serialize_compressed() == P.serialize_proof(cv, P.prove(cv, cs, opk, [7 + i for i in range(29)], "solidity"))
    # an unsatisfied arithmetic gate: WrongQuotientPolyDegree, as for TurboPlonk
    bad2 = arr["witness"].copy()
    bad2[cs.num_vars() - 1] = co.ints_to_limbs([fr.to_mont(123456)], 4)[0]
    if not np.array_equal(bad2, arr["witness"]):
        with pytest.raises((jf.WrongQuotientPolyDegree, jf.InvalidParameters)):
            jf.PlonkKzgSnark.prove_ultra(pk, bad2, bl)
    # key types do not mix
    with pytest.raises(jf.InvalidParameters):
        jf.PlonkKzgSnark.prove(pk, arr["witness"], bl[:17])
    pk.free()
    key.free()


def test_numpy_ultra_bench_circuit_matches_the_restated_circuit(ctx, co, py, P):
    import bench_circuit as B
    cs = P.gen_circuit_for_bench(1 << 9, ultra=True)
    want = U.arrays_from_oracle_circuit(co, py, cs)
    got = B.bench_circuit_arrays(ctx, 9, ultra=True)
    assert cs.k == B.BN254_K + [B.BN254_K5] and cs.n == 512
    for f in ("selectors", "sigmas", "k", "witness", "table_key", "table_dom_sep", "q_dom_sep"):
        assert np.array_equal(got[f], want[f]), f
    assert np.array_equal(got["wire_vars"], want["wire_vars"]) and got["num_vars"] == want["num_vars"]
    assert got["range_bit_len"] == want["range_bit_len"] == 8


@pytest.mark.parametrize("log_n,kind", [(14, "solidity"), (18, "standard"), (20, "solidity")])
def test_large_ultraplonk_proofs_are_accepted_by_the_restated_verifier(ctx, co, py, P, log_n, kind):
    """the reference's UltraPlonk bench shape (plonk/benches/bench.rs with PlonkType::UltraPlonk) up to 2^20 gates"""
    import mpc_jellyfish_b200 as jf
    import bench_circuit as B
    cv, fr = py.BN254, py.BN254_FR
    arr = B.bench_circuit_arrays(ctx, log_n, ultra=True)
    beta = 0x1D3C7A5B9E8F60412B7A6C5D4E3F20198A7B6C5D4E3F2A1B0C9D8E7F6A5B4C3 % fr.p
    key = ctx.generate_srs_for_testing("bn254", beta, arr["n"] + 3)
    pk = jf.PlonkKzgSnark.preprocess_ultra(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [],
                                           arr["range_bit_len"], arr["table_key"], arr["table_dom_sep"], arr["q_dom_sep"])
    bl = np.random.default_rng(log_n).integers(0, 1 << 60, size=(29, 4), dtype=np.uint64)
    proof = jf.PlonkKzgSnark.prove_ultra(pk, arr["witness"], bl, kind)
    got = U.proof_to_oracle(co, cv, proof)
    vk = U.vk_from_product(co, cv, pk, B.BN254_K + [B.BN254_K5])
    assert P.verify(cv, vk, [], got, beta, kind)
    lp = dict(got["plookup_proof"])
    pe = dict(lp["poly_evals"])
    pe["h_1_next_eval"] = (pe["h_1_next_eval"] + 1) % fr.p
    lp["poly_evals"] = pe
    assert not P.verify(cv, vk, [], dict(got, plookup_proof=lp), beta, kind)
    pk.free()
    key.free()


@pytest.mark.parametrize("which", ["test_m12", "bench_3000"])
def test_ultraplonk_resident_coset_evaluations_give_the_same_proof(ctx, co, py, P, which):
    """flags & 1 for UltraPlonk keys: the coset evaluations of the 14 selector, 6 sigma and 4 table polynomials stay resident
    (24 of the 35 polynomials of round 3 are transformed once per key); alone and with the other key-side options, on the seven
    sub-cosets and on the 8n coset: byte-identical proofs, several in a row."""
    import mpc_jellyfish_b200 as jf
    cv, fr = py.BN254, py.BN254_FR
    cs = {"test_m12": lambda: P.gen_circuit_for_test(12, 1, ultra=True), "bench_3000": lambda: P.gen_circuit_for_bench(3000, ultra=True)}[which]()
    beta = 0xABCDEF1234567 % fr.p
    opk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    import plonk_util as U2
    arr = U2.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing("bn254", beta, cs.n + 3)
    rnd = random.Random(15)
    for opts in (dict(cache_coset_evals=True), dict(cache_coset_evals=True, skip_zero_selectors=True, lagrange_wire_commitments=True),
                 dict(cache_coset_evals=True, full_quotient_coset=True)):
        pk = jf.PlonkKzgSnark.preprocess_ultra(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                               arr["pub_gate_ids"], arr["range_bit_len"], arr["table_key"], arr["table_dom_sep"],
                                               arr["q_dom_sep"], **opts)
        for kind in ("solidity", "standard"):
            ints = [rnd.randrange(fr.p) for _ in range(29)]
            bl = co.ints_to_limbs([fr.to_mont(v) for v in ints], 4)
            want = P.serialize_proof(cv, P.prove(cv, cs, opk, ints, kind))
            assert jf.PlonkKzgSnark.prove_ultra(pk, arr["witness"], bl, kind).serialize_compressed() == want, (opts, kind)
        pk.free()
    key.free()
