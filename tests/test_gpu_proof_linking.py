"""Proof linking on the device (jf_plonk_link_hint / jf_plonk_link_proofs / jf_plonk_link_proofs_resident) against the CPU
restatement of plonk/src/proof_system/proof_linking.rs in oracle/plonk_ref.py: the reference's own test circuits and cases
(:330-407, :530-689) -- identical hint polynomials, byte-identical `LinkingProof`s on both division paths, accepted (or, for
pairs that are not linked, rejected) by the restated verifier."""
import os
import random
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

BETA = 0x2C3D4E5F60718293A4B5C6D7E8F9011223344556677889 % (1 << 180)
MAX_DEGREE = 1000  # MAX_DEGREE_TESTING (proof_linking.rs:308)


@pytest.fixture(scope="module")
def P():
    import plonk_ref
    return plonk_ref


def _blinders(co, fr, seed):
    rnd = random.Random(seed)
    ints = [rnd.randrange(fr.p) for _ in range(17)]
    return ints, co.ints_to_limbs([fr.to_mont(v) for v in ints], 4)


class _Side:
    """one circuit proved on both sides"""

    def __init__(self, ctx, co, py, P, cv, key, osrs, cs, kind, seed, with_oracle=True):
        import mpc_jellyfish_b200 as jf
        import plonk_util as U
        fr = cv.fr
        self.cs = cs
        arr = U.arrays_from_oracle_circuit(co, py, cs)
        self.pk = jf.PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"], arr["num_vars"],
                                              arr["pub_gate_ids"])
        ints, bl = _blinders(co, fr, seed)
        self.proof, self.hint = jf.PlonkKzgSnark.prove_with_link_hint(self.pk, arr["witness"], bl, kind)
        self.oproof = U.proof_to_oracle(co, cv, self.proof)
        self.ohint = None
        if with_oracle:
            opk = P.preprocess(cv, osrs, cs)
            want, self.ohint = P.prove_with_link_hint(cv, cs, opk, ints, kind)
            assert self.proof.serialize_compressed() == P.serialize_proof(cv, want)
            got_poly = [fr.from_mont(v) for v in co.limbs_to_ints(self.hint.linking_wire_poly)]
            assert P._strip(got_poly) == self.ohint["linking_wire_poly"]
            assert U.point_to_affine(co, cv, self.hint.linking_wire_comm, self.hint.linking_wire_inf) == self.ohint["linking_wire_comm"]
        else:
            self.ohint = {"linking_wire_comm": self.oproof["wires_poly_comms"][0]}

    def free(self):
        self.pk.free()


def _link_all_ways(ctx, key, lhs, rhs, layout, kind):
    """host hints (both division paths) and the resident form: the three must agree byte for byte"""
    import mpc_jellyfish_b200 as jf
    lay = jf.GroupLayout(layout.alignment, layout.offset, layout.size)
    a = jf.PlonkKzgSnark.link_proofs(ctx, key, lhs.hint, rhs.hint, lay, kind)
    b = jf.PlonkKzgSnark.link_proofs(ctx, key, lhs.hint, rhs.hint, lay, kind, sequential_division=True)
    c = jf.PlonkKzgSnark.link_proofs_resident(lhs.pk, lhs.proof, rhs.pk, rhs.proof, lay, kind)
    assert b.path == 1
    assert a.serialize_compressed() == b.serialize_compressed() == c.serialize_compressed()
    assert np.array_equal(a.eta, b.eta) and a.path == c.path
    return a


def _to_oracle_link(co, cv, lp):
    import plonk_util as U
    return {"quotient_commitment": U.point_to_affine(co, cv, lp.quotient_commitment, lp.quotient_inf),
            "opening_proof": U.point_to_affine(co, cv, lp.opening_proof, lp.opening_inf)}


@pytest.mark.parametrize("kind", ["solidity", "standard"])
@pytest.mark.parametrize("curve", ["bn254", "bls12_381"])
def test_linking_proofs_match_the_cpu_restatement(ctx, co, py, P, curve, kind):
    cv = py.BN254 if curve == "bn254" else py.BLS12_381
    fr = cv.fr
    beta = BETA % fr.p
    key = ctx.generate_srs_for_testing(curve, beta, MAX_DEGREE + 1)
    osrs = P.gen_srs(cv, beta, MAX_DEGREE)
    rnd = random.Random(21)
    witness = [rnd.randrange(fr.p) for _ in range(10)]
    c1 = P.gen_link_test_circuit(1, witness, None, fr)
    lay = c1.get_link_group_layout(P.LINK_GROUP_NAME)
    s1 = _Side(ctx, co, py, P, cv, key, osrs, c1, kind, 1)
    s1b = _Side(ctx, co, py, P, cv, key, osrs, P.gen_link_test_circuit(1, witness, None, fr), kind, 2)
    s2 = _Side(ctx, co, py, P, cv, key, osrs, P.gen_link_test_circuit(2, witness, lay, fr), kind, 3)
    for lhs, rhs in ((s1, s1b), (s1, s2), (s2, s1)):
        lp = _link_all_ways(ctx, key, lhs, rhs, lay, kind)
        assert lp.path == 0                                       # the hints agree on the link domain: exact division
        want = P.link_proofs(cv, lhs.ohint, rhs.ohint, lay, osrs, kind)
        assert lp.serialize_compressed() == P.serialize_link_proof(cv, want)
        assert fr.from_mont(co.limbs_to_ints(lp.eta.reshape(1, 4))[0]) == want["eta"]
        assert P.verify_link_proof(cv, lhs.oproof, rhs.oproof, _to_oracle_link(co, cv, lp), lay, beta, kind)
    # a proof linked with itself: the empty quotient (proof_linking.rs:122-125) -> identity commitments
    lp = _link_all_ways(ctx, key, s1, s1, lay, kind)
    assert lp.quotient_inf and lp.opening_inf
    assert lp.serialize_compressed() == P.serialize_link_proof(cv, P.link_proofs(cv, s1.ohint, s1.ohint, lay, osrs, kind))
    for s in (s1, s1b, s2):
        s.free()
    key.free()


def test_specific_layout_and_pairs_that_are_not_linked(ctx, co, py, P):
    """test_valid_proof_link__specific_layout, test_invalid_proof_link__different_witnesses / __wrong_alignment / __wrong_offset"""
    cv, fr = py.BN254, py.BN254_FR
    beta = BETA % fr.p
    key = ctx.generate_srs_for_testing("bn254", beta, MAX_DEGREE + 1)
    osrs = P.gen_srs(cv, beta, MAX_DEGREE)
    rnd = random.Random(22)
    w1 = [rnd.randrange(fr.p) for _ in range(10)]
    layout = P.GroupLayout(8, 20, 10)
    a = _Side(ctx, co, py, P, cv, key, osrs, P.gen_link_test_circuit(1, w1, layout, fr), "solidity", 1)
    b = _Side(ctx, co, py, P, cv, key, osrs, P.gen_link_test_circuit(2, w1, layout, fr), "solidity", 2)
    lp = _link_all_ways(ctx, key, a, b, layout, "solidity")
    assert lp.path == 0
    assert lp.serialize_compressed() == P.serialize_link_proof(cv, P.link_proofs(cv, a.ohint, b.ohint, layout, osrs, "solidity"))
    assert P.verify_link_proof(cv, a.oproof, b.oproof, _to_oracle_link(co, cv, lp), layout, beta, "solidity")
    # different witnesses: a1 - a2 does not vanish on the link domain; the floor quotient is still the reference's
    w2 = list(w1)
    w2[rnd.randrange(10)] = rnd.randrange(fr.p)
    c = _Side(ctx, co, py, P, cv, key, osrs, P.gen_link_test_circuit(2, w2, layout, fr), "solidity", 3)
    bad = _link_all_ways(ctx, key, a, c, layout, "solidity")
    assert bad.path == 2                                         # remainder taken off, then the exact division
    assert bad.serialize_compressed() == P.serialize_link_proof(cv, P.link_proofs(cv, a.ohint, c.ohint, layout, osrs, "solidity"))
    assert not P.verify_link_proof(cv, a.oproof, c.oproof, _to_oracle_link(co, cv, bad), layout, beta, "solidity")
    # the right witness on a misaligned domain / at another offset
    for wrong in (P.GroupLayout(9, 20, 10), P.GroupLayout(8, 19, 10)):
        d = _Side(ctx, co, py, P, cv, key, osrs, P.gen_link_test_circuit(2, w1, wrong, fr), "solidity", 4)
        bad = _link_all_ways(ctx, key, a, d, wrong, "solidity")
        assert bad.serialize_compressed() == P.serialize_link_proof(cv, P.link_proofs(cv, a.ohint, d.ohint, wrong, osrs, "solidity"))
        assert not P.verify_link_proof(cv, a.oproof, d.oproof, _to_oracle_link(co, cv, bad), wrong, beta, "solidity")
        d.free()
    for s in (a, b, c):
        s.free()
    key.free()


def test_link_argument_checks(ctx, co, py, P):
    import mpc_jellyfish_b200 as jf
    fr = py.BN254_FR
    key = ctx.generate_srs_for_testing("bn254", 7, 64)
    poly = co.ints_to_limbs([fr.to_mont(v) for v in range(1, 35)], 4)
    pt = np.zeros(8, dtype=np.uint64)
    hint = jf.LinkingHint(poly, pt, True)
    with pytest.raises(jf.InvalidParameters):       # empty group
        jf.PlonkKzgSnark.link_proofs(ctx, key, hint, hint, jf.GroupLayout(4, 3, 0))
    with pytest.raises(jf.InvalidParameters):       # offset + size >= 2^alignment (validate_layout)
        jf.PlonkKzgSnark.link_proofs(ctx, key, hint, hint, jf.GroupLayout(4, 10, 6))
    with pytest.raises(jf.DomainCreationError):     # no 2^29-th root of unity in BN254 Fr
        jf.PlonkKzgSnark.link_proofs(ctx, key, hint, hint, jf.GroupLayout(29, 3, 2))
    big = jf.LinkingHint(co.ints_to_limbs([1] * 80, 4), pt, True)
    with pytest.raises(jf.InvalidParameters):       # degree above the commit key
        jf.PlonkKzgSnark.link_proofs(ctx, key, big, hint, jf.GroupLayout(4, 3, 2))
    # a group larger than the polynomials: the quotient is empty, the opening is that of a1 - a2 itself
    other = jf.LinkingHint(co.ints_to_limbs([fr.to_mont(v) for v in range(5, 25)], 4), pt, True)
    lp = jf.PlonkKzgSnark.link_proofs(ctx, key, hint, other, jf.GroupLayout(6, 3, 40))
    assert lp.quotient_inf and not lp.opening_inf
    key.free()


@pytest.mark.parametrize("log_n,size", [(12, 100), (14, 1000)])
def test_large_link_groups(ctx, co, py, P, log_n, size):
    """a link group of `size` values in circuits of 2^log_n gates: exact coset division == `size` linear divisions, and the restated
    verifier accepts; one changed value is caught"""
    import mpc_jellyfish_b200 as jf
    cv, fr = py.BN254, py.BN254_FR
    beta = BETA % fr.p
    n = 1 << log_n
    key = ctx.generate_srs_for_testing("bn254", beta, n + 3)
    rnd = random.Random(log_n)
    vals = [rnd.randrange(fr.p) for _ in range(size)]
    layout = P.GroupLayout(log_n - 1, 5, size)

    def circuit(values, adds, seed):
        cs = P.PlonkCircuit(fr)
        x = cs.create_public_variable(seed)
        g = cs.create_link_group("shared", layout)
        vs = [cs.create_variable_with_link_groups(v, [g]) for v in values]
        acc = x
        for i in range(adds):
            acc = cs.add(acc, vs[i % size])
        cs.finalize_for_arithmetization()
        assert cs.n in (n, n // 2) and cs.check_satisfiability()   # the two circuits live on different domains
        return cs

    a = _Side(ctx, co, py, P, cv, key, None, circuit(vals, n // 2, 3), "solidity", 1, with_oracle=False)
    b = _Side(ctx, co, py, P, cv, key, None, circuit(vals, n // 3, 4), "solidity", 2, with_oracle=False)
    lp = _link_all_ways(ctx, key, a, b, layout, "solidity")
    assert lp.path == 0
    assert P.verify_link_proof(cv, a.oproof, b.oproof, _to_oracle_link(co, cv, lp), layout, beta, "solidity")
    vals2 = list(vals)
    vals2[size // 2] = (vals2[size // 2] + 1) % fr.p
    c = _Side(ctx, co, py, P, cv, key, None, circuit(vals2, n // 3, 4), "solidity", 3, with_oracle=False)
    bad = jf.PlonkKzgSnark.link_proofs_resident(a.pk, a.proof, c.pk, c.proof, jf.GroupLayout(layout.alignment, layout.offset, layout.size))
    assert bad.path == (2 if layout.alignment <= 12 else 1)      # 2^alignment coefficients must fit one CTA's shared memory
    seq = jf.PlonkKzgSnark.link_proofs_resident(a.pk, a.proof, c.pk, c.proof, jf.GroupLayout(layout.alignment, layout.offset, layout.size),
                                                sequential_division=True)
    assert seq.path == 1 and seq.serialize_compressed() == bad.serialize_compressed()
    assert not P.verify_link_proof(cv, a.oproof, c.oproof, _to_oracle_link(co, cv, bad), layout, beta, "solidity")
    for s in (a, b, c):
        s.free()
    key.free()


@pytest.mark.parametrize("seed", range(6))
def test_random_hints_and_layouts_match_the_cpu_restatement(ctx, co, py, P, seed):
    """hints that do not come from circuits: unequal lengths, groups longer than a polynomial, size 1, alignments above and below the
    polynomial length, pairs that vanish on the group and pairs that do not -- quotient commitment, eta and opening == oracle"""
    import mpc_jellyfish_b200 as jf
    import plonk_util as U
    cv, fr = py.BN254, py.BN254_FR
    p = fr.p
    beta = BETA % p
    rnd = random.Random(1000 + seed)
    key = ctx.generate_srs_for_testing("bn254", beta, 700)
    osrs = P.gen_srs(cv, beta, 699)
    mont = lambda xs: co.ints_to_limbs([fr.to_mont(x) for x in xs], 4) if xs else np.zeros((0, 4), dtype=np.uint64)  # noqa: E731
    for case in range(6):
        align = rnd.randrange(1, 11)
        size = rnd.randrange(1, max(2, min(1 << align, 200) - 1))
        offset = rnd.randrange(0, (1 << align) - size)
        if offset + size >= (1 << align):
            continue
        lay = P.GroupLayout(align, offset, size)
        len1 = rnd.choice([0, 1, size, size + 1, rnd.randrange(1, 600), rnd.randrange(300, 690)])
        a1 = [rnd.randrange(p) for _ in range(len1)]
        mode = rnd.choice(["linked", "linked_longer", "unrelated", "equal", "shorter"])
        if mode == "equal":
            a2 = list(a1)
        elif mode == "unrelated":
            a2 = [rnd.randrange(p) for _ in range(rnd.randrange(0, 690))]
        elif mode == "shorter":
            a2 = a1[: len(a1) // 2]
        else:   # a2 = a1 + Z_D * r for a random r: equal on the group, different elsewhere
            z = [1]
            for r_ in P._link_roots(fr, lay):
                nz = [0] * (len(z) + 1)
                for i, c in enumerate(z):
                    nz[i] = (nz[i] - c * r_) % p
                    nz[i + 1] = (nz[i + 1] + c) % p
                z = nz
            rl = rnd.randrange(1, 40) if mode == "linked" else max(1, 680 - size - 1)
            rpoly = [rnd.randrange(p) for _ in range(rl)]
            prod = [0] * (len(z) + rl - 1)
            for i, zc in enumerate(z):
                for j, rc in enumerate(rpoly):
                    prod[i + j] = (prod[i + j] + zc * rc) % p
            a2 = P._poly_add(p, a1, prod)
        if max(len(a1), len(a2)) > 690:
            continue
        c1, c2 = ctx.msm(key, mont(a1), montgomery=True) if a1 else (np.zeros(8, dtype=np.uint64), True), \
            ctx.msm(key, mont(a2), montgomery=True) if a2 else (np.zeros(8, dtype=np.uint64), True)
        h1 = jf.LinkingHint(mont(a1), c1[0], bool(c1[1]))
        h2 = jf.LinkingHint(mont(a2), c2[0], bool(c2[1]))
        oh1 = {"linking_wire_poly": P._strip(a1), "linking_wire_comm": U.point_to_affine(co, cv, c1[0], c1[1])}
        oh2 = {"linking_wire_poly": P._strip(a2), "linking_wire_comm": U.point_to_affine(co, cv, c2[0], c2[1])}
        want = P.link_proofs(cv, oh1, oh2, lay, osrs, "solidity")
        jl = jf.GroupLayout(align, offset, size)
        for seq in (False, True):
            got = jf.PlonkKzgSnark.link_proofs(ctx, key, h1, h2, jl, "solidity", sequential_division=seq)
            assert got.serialize_compressed() == P.serialize_link_proof(cv, want), (seed, case, mode, lay, len(a1), len(a2), seq)
        if mode in ("linked", "linked_longer", "equal") and max(len(a1), len(a2)) > size:
            assert jf.PlonkKzgSnark.link_proofs(ctx, key, h1, h2, jl, "solidity").path == 0
        elif max(len(a1), len(a2)) > size:
            assert jf.PlonkKzgSnark.link_proofs(ctx, key, h1, h2, jl, "solidity").path in (0, 2)   # alignments here are <= 10
    key.free()
