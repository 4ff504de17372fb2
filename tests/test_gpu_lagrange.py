"""The commit key in the Lagrange basis (`jf_srs_lagrange`, csrc/lagrange.cu: an inverse DFT of the key points taken in the group) and
the wire commitments of the prover over it (`jf_plonk_preprocess` flags & 8).  With the test SRS's known beta every point has a
closed form, [L_j(beta)] G with L_j(beta) = w^j (beta^n - 1) / (n (beta - w^j)); commitments over values must equal commitments
over coefficients; proofs must not change by a byte."""
import os
import random
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu
BETA = 0x1F2E3D4C5B6A79880102030405060708090A0B0C0D0E0F


@pytest.fixture(scope="module")
def P():
    import plonk_ref
    return plonk_ref


@pytest.mark.parametrize("curve", ["bn254", "bls12_381"])
@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 9, 12])
def test_lagrange_key_points_have_their_closed_form(ctx, co, py, curve, log_n):
    cv = py.CURVES[curve]
    fr = cv.fr
    p = fr.p
    n = 1 << log_n
    beta = BETA % p
    key = ctx.generate_srs_for_testing(curve, beta, n + 3)
    lag = key.lagrange(log_n, mask_points=True)
    assert len(lag) == n + 2
    got = lag.read(0, n + 2)
    dom = py.Radix2Domain(fr, n)
    w = dom.group_gen
    bn = pow(beta, n, p)
    ninv = pow(n, -1, p)
    wj, scal = 1, []
    for j in range(n):
        scal.append(wj * (bn - 1) % p * ninv % p * pow((beta - wj) % p, -1, p) % p)
        wj = wj * w % p
    scal += [(bn - 1) % p, (bn * beta - beta) % p]          # P_n - P_0, P_(n+1) - P_1
    step = max(1, n // 64)
    idx = sorted(set(list(range(0, n, step)) + [n - 1, n, n + 1]))
    want = co.fixed_base_mul(curve, co.ints_to_limbs([scal[i] for i in idx], 4))
    assert np.array_equal(got[idx], want)
    # commit over values == commit over coefficients
    coeffs = [random.Random(log_n).randrange(p) for _ in range(n)]
    cm = co.ints_to_limbs([fr.to_mont(c) for c in coeffs], 4)
    evals = ctx.ntt(cv.fr.name, cm.copy(), log_n, False) if n > 1 else cm.copy()
    a = ctx.msm(key, cm, montgomery=True)
    b = ctx.msm(lag, evals, montgomery=True)
    assert a[1] == b[1] and np.array_equal(a[0], b[0])
    lag.free()
    key.free()


def test_lagrange_key_argument_checks(ctx):
    import mpc_jellyfish_b200 as jf
    key = ctx.generate_srs_for_testing("bn254", 5, 40)
    with pytest.raises(jf.InvalidParameters):
        key.lagrange(6)                       # 64 points needed
    key.lagrange(5).free()                    # 32 points: fine
    small = ctx.generate_srs_for_testing("bn254", 5, 33)
    with pytest.raises(jf.InvalidParameters):
        small.lagrange(5, mask_points=True)   # 34 points needed
    small.free()
    with pytest.raises(jf.DomainCreationError):
        key.lagrange(29)
    key.free()


def _blinders(co, fr, count, seed):
    rnd = random.Random(seed)
    ints = [rnd.randrange(fr.p) for _ in range(count)]
    return ints, co.ints_to_limbs([fr.to_mont(v) for v in ints], 4)


@pytest.mark.parametrize("name", ["test_m20", "bench_2^10", "all_selectors_m9", "tiny_n8"])
def test_proofs_with_lagrange_wire_commitments_are_byte_identical(ctx, co, py, P, name):
    import mpc_jellyfish_b200 as jf
    import plonk_util as U
    cv, fr = py.BN254, py.BN254_FR

    def tiny():
        cs = P.PlonkCircuit()
        a = cs.create_variable(5)
        for _ in range(4):
            a = cs.add(a, cs.one())
        cs.finalize_for_arithmetization()
        return cs
    cs = {"test_m20": lambda: P.gen_circuit_for_test(20, 1), "bench_2^10": lambda: P.gen_circuit_for_bench(1 << 10),
          "all_selectors_m9": lambda: P.gen_circuit_all_selectors(9), "tiny_n8": tiny}[name]()
    beta = BETA % fr.p
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing("bn254", beta, cs.n + 3)
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                     arr["pub_gate_ids"], lagrange_wire_commitments=True)
    opk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    for kind in ("solidity", "standard"):
        ints, bl = _blinders(co, fr, 17, 7)
        want = P.serialize_proof(cv, P.prove(cv, cs, opk, ints, kind))
        for _ in range(2):
            assert jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, kind).serialize_compressed() == want
    pk.free()
    key.free()


def test_ultraplonk_and_batch_with_lagrange_wire_commitments(ctx, co, py, P):
    import mpc_jellyfish_b200 as jf
    import plonk_util as U
    cv, fr = py.BN254, py.BN254_FR
    beta = BETA % fr.p
    cs = P.gen_circuit_for_test(6, 1, ultra=True)
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing("bn254", beta, cs.n + 3)
    pk = jf.PlonkKzgSnark.preprocess_ultra(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                           arr["pub_gate_ids"], arr["range_bit_len"], arr["table_key"], arr["table_dom_sep"],
                                           arr["q_dom_sep"], lagrange_wire_commitments=True)
    opk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    ints, bl = _blinders(co, fr, 29, 8)
    assert jf.PlonkKzgSnark.prove_ultra(pk, arr["witness"], bl, "solidity").serialize_compressed() == \
        P.serialize_proof(cv, P.prove(cv, cs, opk, ints, "solidity"))
    pk.free()
    key.free()
    # a batch of two TurboPlonk instances, one key with the Lagrange form and one without
    c1, c2 = P.gen_circuit_for_test(20, 1), P.gen_circuit_for_test(20, 5)
    a1, a2 = U.arrays_from_oracle_circuit(co, py, c1), U.arrays_from_oracle_circuit(co, py, c2)
    key = ctx.generate_srs_for_testing("bn254", beta, c1.n + 3)
    pks = [jf.PlonkKzgSnark.preprocess(ctx, key, a["selectors"], a["sigmas"], a["k"], a["wire_vars"], a["num_vars"], a["pub_gate_ids"],
                                       lagrange_wire_commitments=flag) for a, flag in ((a1, True), (a2, False))]
    osrs = P.gen_srs(cv, beta, c1.n + 2)
    opks = [P.preprocess(cv, osrs, c) for c in (c1, c2)]
    ints, bl = _blinders(co, fr, P.batch_num_blinders([c1, c2]), 9)
    got = jf.PlonkKzgSnark.batch_prove(pks, [a1["witness"], a2["witness"]], bl, "solidity")
    assert got.serialize_compressed() == P.serialize_batch_proof(cv, P.batch_prove(cv, [c1, c2], opks, ints, "solidity"))
    for k_ in pks:
        k_.free()
    key.free()


def test_lagrange_wire_commitments_at_2_pow_16(ctx, co, py, P):
    """the bench circuit at 2^16 gates: the restated verifier accepts, and the bytes equal the monomial-basis proof"""
    import mpc_jellyfish_b200 as jf
    import bench_circuit as B
    import plonk_util as U
    cv, fr = py.BN254, py.BN254_FR
    arr = B.bench_circuit_arrays(ctx, 16)
    beta = BETA % fr.p
    key = ctx.generate_srs_for_testing("bn254", beta, arr["n"] + 3)
    _, bl = _blinders(co, fr, 17, 3)
    proofs = []
    for flag in (False, True):
        pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"], [],
                                         lagrange_wire_commitments=flag)
        proofs.append(jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "solidity"))
        if flag:
            vk = U.vk_from_product(co, cv, pk, B.BN254_K)
            assert P.verify(cv, vk, [], U.proof_to_oracle(co, cv, proofs[-1]), beta, "solidity")
        pk.free()
    assert proofs[0].serialize_compressed() == proofs[1].serialize_compressed()
    key.free()


def test_bls12_381_proof_with_lagrange_wire_commitments(ctx, co, py, P):
    import mpc_jellyfish_b200 as jf
    import plonk_util as U
    cv, fr = py.BLS12_381, py.BLS12_381_FR
    cs = P.gen_circuit_for_test(20, 1, fr)
    beta = BETA % fr.p
    arr = U.arrays_from_oracle_circuit(co, py, cs)
    key = ctx.generate_srs_for_testing("bls12_381", beta, cs.n + 3)
    pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                     arr["pub_gate_ids"], lagrange_wire_commitments=True)
    opk = P.preprocess(cv, P.gen_srs(cv, beta, cs.n + 2), cs)
    ints, bl = _blinders(co, fr, 17, 17)
    assert jf.PlonkKzgSnark.prove(pk, arr["witness"], bl, "standard").serialize_compressed() == \
        P.serialize_proof(cv, P.prove(cv, cs, opk, ints, "standard"))
    pk.free()
    key.free()
