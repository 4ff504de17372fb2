"""GPU suite: `PlonkKzgSnark::batch_prove` (several instances, one transcript, ONE quotient, one pair of opening proofs;
plonk/src/proof_system/snark.rs:201-469) through `jf_plonk_batch_prove` / `jf_ultraplonk_batch_prove`, byte for byte against the
CPU restatement and accepted by the restated batch verifier (verifier.rs:68-254)."""
import os
import random
import sys

import numpy as np
import pytest

import plonk_util as U

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.fixture(scope="module")
def P():
    import plonk_ref
    return plonk_ref


def _batch_to_oracle(co, cv, bp):
    recs = [U.proof_to_oracle(co, cv, pr) for pr in bp.proofs]
    return {"wires_poly_comms_vec": [r["wires_poly_comms"] for r in recs],
            "prod_perm_poly_comms_vec": [r["prod_perm_poly_comm"] for r in recs],
            "poly_evals_vec": [{"wires_evals": r["wires_evals"], "wire_sigma_evals": r["wire_sigma_evals"], "perm_next_eval": r["perm_next_eval"]}
                               for r in recs],
            "plookup_proofs_vec": [r["plookup_proof"] for r in recs],
            "split_quot_poly_comms": recs[0]["split_quot_poly_comms"], "opening_proof": recs[0]["opening_proof"],
            "shifted_opening_proof": recs[0]["shifted_opening_proof"]}


def _preprocess(ctx, jf, arr, key, ultra, **kw):
    if ultra:
        return jf.PlonkKzgSnark.preprocess_ultra(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                                 arr["pub_gate_ids"], arr["range_bit_len"], arr["table_key"], arr["table_dom_sep"],
                                                 arr["q_dom_sep"])
    return jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                       arr["pub_gate_ids"], **kw)


@pytest.mark.parametrize("kind", ["solidity", "standard"])
@pytest.mark.parametrize("ultra", [False, True], ids=["turbo", "ultra"])
def test_batch_proof_bytes_match_the_cpu_restatement(ctx, co, py, P, ultra, kind):
    import mpc_jellyfish_b200 as jf
    cv, fr = py.BN254, py.BN254_FR
    css = [P.gen_circuit_for_test(3, 2, ultra=ultra), P.gen_circuit_for_test(3, 1, ultra=ultra),
           P.gen_circuit_for_test(5, 3, ultra=ultra) if ultra else P.gen_circuit_for_bench(32)]
    beta = 0x55AA55AA55AA55AA1234 % fr.p
    key = ctx.generate_srs_for_testing("bn254", beta, css[0].n + 3)
    arrs = [U.arrays_from_oracle_circuit(co, py, cs) for cs in css]
    pks = [_preprocess(ctx, jf, a, key, ultra) for a in arrs]
    srs = P.gen_srs(cv, beta, css[0].n + 2)
    opks = [P.preprocess(cv, srs, cs) for cs in css]
    rnd = random.Random(21)
    ints = [rnd.randrange(fr.p) for _ in range(P.batch_num_blinders(css))]
    bl = co.ints_to_limbs([fr.to_mont(v) for v in ints], 4)
    bp = jf.PlonkKzgSnark.batch_prove(pks, [a["witness"] for a in arrs], bl, kind)
    want = P.batch_prove(cv, css, opks, ints, kind)
    got = _batch_to_oracle(co, cv, bp)
    for name in ("wires_poly_comms_vec", "prod_perm_poly_comms_vec", "poly_evals_vec", "plookup_proofs_vec", "split_quot_poly_comms",
                 "opening_proof", "shifted_opening_proof"):
        assert got[name] == want[name], name
    assert bp.serialize_compressed() == P.serialize_batch_proof(cv, want)
    vks = [U.vk_from_product(co, cv, pk, cs.k) for pk, cs in zip(pks, css)]
    pis = [cs.public_input() for cs in css]
    assert P.batch_verify(cv, vks, pis, got, beta, kind)
    assert not P.batch_verify(cv, [vks[1], vks[0], vks[2]], [pis[1], pis[0], pis[2]], got, beta, kind)
    # a batch of one is `prove`
    one_ints = ints[:2 * css[0].nw] + [9] * (P.num_blinders(css[0]) - 2 * css[0].nw)
    one_bl = co.ints_to_limbs([fr.to_mont(v) for v in one_ints], 4)
    b1 = jf.PlonkKzgSnark.batch_prove(pks[:1], [arrs[0]["witness"]], one_bl, kind)
    single = (jf.PlonkKzgSnark.prove_ultra if ultra else jf.PlonkKzgSnark.prove)(pks[0], arrs[0]["witness"], one_bl, kind)
    assert b1.proofs[0].serialize_compressed() == single.serialize_compressed() == P.serialize_proof(cv, P.prove(cv, css[0], opks[0], one_ints, kind))
    # the same key twice, keys of different sizes and mixed key types are refused
    with pytest.raises(jf.InvalidParameters):
        jf.PlonkKzgSnark.batch_prove([pks[0], pks[0]], [arrs[0]["witness"]] * 2, bl[: 2 * (len(one_ints) - css[0].nw + 1) + css[0].nw - 1], kind)
    for pk in pks:
        pk.free()
    key.free()


def test_batch_of_two_large_instances_is_accepted(ctx, co, py, P):
    """two 2^16-gate bench instances (different witnesses would need different circuits: the bench circuit's witness is fixed, so
    the second instance uses the skip-zero-selector key of the same circuit): one quotient, verifier-accepted"""
    import mpc_jellyfish_b200 as jf
    import bench_circuit as B
    cv, fr = py.BN254, py.BN254_FR
    log_n = 16
    arr = B.bench_circuit_arrays(ctx, log_n)
    beta = 0x1D3C7A5B9E8F60412B7A6C5D4E3F2019 % fr.p
    key = ctx.generate_srs_for_testing("bn254", beta, arr["n"] + 3)
    pks = [jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [])
           for _ in range(2)]
    bl = np.random.default_rng(3).integers(0, 1 << 60, size=(2 * 13 + 4, 4), dtype=np.uint64)
    bp = jf.PlonkKzgSnark.batch_prove(pks, [arr["witness"], arr["witness"]], bl, "solidity")
    got = _batch_to_oracle(co, cv, bp)
    vks = [U.vk_from_product(co, cv, pk, B.BN254_K) for pk in pks]
    assert P.batch_verify(cv, vks, [[], []], got, beta, "solidity")
    pe = [dict(x) for x in got["poly_evals_vec"]]
    pe[1]["perm_next_eval"] = (pe[1]["perm_next_eval"] + 1) % fr.p
    assert not P.batch_verify(cv, vks, [[], []], dict(got, poly_evals_vec=pe), beta, "solidity")
    for pk in pks:
        pk.free()
    key.free()
