"""CPU suite: the proof-linking restatement in oracle/plonk_ref.py.
Circuit layout: the reference's own tests (relation/src/proof_linking/linkable_circuit.rs:510-700), with its random draws replaced
by seeded sweeps.  Linking proofs: the reference's tests of plonk/src/proof_system/proof_linking.rs:530-689 -- prover restated from
:79-216, verifier from :233-299 with the pairing check in its known-beta G1 form."""
import random

import pytest


@pytest.fixture(scope="module")
def P():
    import plonk_ref
    return plonk_ref


def _assert_link_group_placement(P, cs, gid, layout):
    """assert_link_group_placement (linkable_circuit.rs:476-503)"""
    log_n = cs.n.bit_length() - 1
    spacing = 1 << (log_n - layout.alignment)
    for i, var in enumerate(cs.link_groups[gid]):
        row = layout.offset * spacing + i * spacing
        assert cs.gates[row].name == "link"
        assert cs.wire_variables[P.PROOF_LINK_WIRE_IDX][row] == var
        assert all(cs.wire_variables[j][row] == 0 for j in range(1, 5))


@pytest.mark.parametrize("seed", range(12))
def test_circuit_layout(P, py, seed):
    rnd = random.Random(seed)
    lo, hi = 6, 10
    cs = P.PlonkCircuit(py.BN254_FR)
    a1 = rnd.randrange(lo, hi)
    o1 = rnd.randrange(10, 20)
    a3 = rnd.randrange(a1, hi + 1)
    o3 = rnd.randrange(o1 * (1 << (a3 - a1)) + 1, 1 << a3)
    g1 = cs.create_link_group("group1", P.GroupLayout(a1, o1, 1))
    g2 = cs.create_link_group("group2", None)
    g3 = cs.create_link_group("group3", P.GroupLayout(a3, o3, 1))
    cs.create_public_variable(1)
    cs.create_public_variable(1)
    cs.create_variable_with_link_groups(rnd.randrange(py.BN254_FR.p), [g1, g2])
    cs.create_variable_with_link_groups(rnd.randrange(py.BN254_FR.p), [g2, g3])
    order, size = cs._generate_layout()
    assert cs.num_inputs() == 2 and cs.num_gates() == 4   # two i/o gates, two constant gates
    lay = dict(order)
    assert lay["group1"] == P.GroupLayout(a1, o1, 1)
    assert lay["group2"] == P.GroupLayout(a3, 2, 2)       # right behind the public inputs
    assert lay["group3"] == P.GroupLayout(a3, o3, 1)
    cs.finalize_for_arithmetization()
    assert cs.n == max(8, 1 << a3)
    for gid in ("group1", "group2", "group3"):
        _assert_link_group_placement(P, cs, gid, lay[gid])
    assert cs.check_satisfiability()


def test_invalid_circuit_layout(P, py):
    cs = P.PlonkCircuit(py.BN254_FR)
    g = cs.create_link_group("test", P.GroupLayout(4, 0, 1))     # a link group inside the public inputs
    cs.create_public_variable(1)
    cs.create_variable_with_link_groups(5, [g])
    with pytest.raises(ValueError, match="mangle public inputs"):
        cs._generate_layout()
    for seed in range(10):                                       # two groups at conflicting offsets
        rnd = random.Random(seed)
        cs = P.PlonkCircuit(py.BN254_FR)
        a1 = rnd.randrange(0, 10)
        o1 = rnd.randrange(0, 1 << a1)
        a2 = rnd.randrange(a1, 11)
        spacing = 1 << (a2 - a1)
        g1 = cs.create_link_group("group1", P.GroupLayout(a1, o1, 2))
        g2 = cs.create_link_group("group2", P.GroupLayout(a2, o1 * spacing + spacing, 2))
        cs.create_variable_with_link_groups(7, [g1, g2])
        cs.create_variable_with_link_groups(9, [g1, g2])
        with pytest.raises(ValueError):
            cs._generate_layout()


@pytest.mark.parametrize("seed", range(8))
def test_proof_linking_gates_keep_the_circuit(P, py, seed):
    """test_proof_linking_gates (linkable_circuit.rs:622-699): every gate and its wiring survive the layout"""
    rnd = random.Random(100 + seed)
    N = 100
    n_inputs = rnd.randrange(1, N)
    n_gates = rnd.randrange(n_inputs, 2 * n_inputs)
    n_witness = rnd.randrange(1, N)
    n_links = rnd.randrange(1, N)
    p = py.BN254_FR.p
    cs = P.PlonkCircuit(py.BN254_FR)
    for _ in range(n_inputs):
        cs.create_public_variable(rnd.randrange(p))
    for _ in range(n_witness):
        cs.create_variable(rnd.randrange(p))
    wiring = []
    for _ in range(n_gates):
        wires = [rnd.randrange(n_witness) for _ in range(5)]
        cs.insert_gate(wires, P.Gate("dummy"))
        wiring.append(wires)
    alignment = rnd.randrange(N.bit_length(), N.bit_length() + 3)
    # add_random_link_group (:454-473)
    offset = rnd.randrange(n_inputs + 1, (1 << alignment) - n_links - 1) if (1 << alignment) - n_links - 1 > n_inputs + 1 else n_inputs + 1
    if offset + n_links >= (1 << alignment):
        pytest.skip("the drawn group does not fit its alignment")
    g = cs.create_link_group("test_group", P.GroupLayout(alignment, offset, n_links))
    for _ in range(n_links):
        cs.create_variable_with_link_groups(rnd.randrange(p), [g])
    cs.finalize_for_arithmetization()
    got = sorted([cs.wire_variables[j][i] for j in range(5)] for i, gt in enumerate(cs.gates) if gt.name == "dummy")
    assert got == sorted(wiring)
    assert sum(1 for gt in cs.gates if gt.name == "io") == n_inputs
    assert sum(1 for gt in cs.gates if gt.name == "link") == n_links
    _assert_link_group_placement(P, cs, "test_group", cs.get_link_group_layout("test_group"))


# ---- linking proofs ------------------------------------------------------------------------------------------------------------
BETA = 0x0BADC0FFEE0DDF00D1234567 % (1 << 200)
MAX_DEGREE = 1000  # MAX_DEGREE_TESTING (proof_linking.rs:308)


def _proof_and_hint(P, cv, srs, which, witness, layout, kind, seed):
    cs = P.gen_link_test_circuit(which, witness, layout)
    assert cs.check_satisfiability()
    lay = cs.get_link_group_layout(P.LINK_GROUP_NAME)
    pk = P.preprocess(cv, srs, cs)
    rnd = random.Random(seed)
    proof, hint = P.prove_with_link_hint(cv, cs, pk, [rnd.randrange(cv.fr.p) for _ in range(17)], kind)
    assert P.verify(cv, pk["vk"], cs.public_input(), proof, BETA, kind)
    assert hint["linking_wire_comm"] == proof["wires_poly_comms"][0]
    return proof, hint, lay


@pytest.fixture(scope="module")
def link_srs(P, py):
    return P.gen_srs(py.BN254, BETA, MAX_DEGREE)


def _link_and_verify(P, cv, srs, h1, h2, p1, p2, layout, kind):
    lp = P.link_proofs(cv, h1, h2, layout, srs, kind)
    return P.verify_link_proof(cv, p1, p2, lp, layout, BETA, kind), lp


@pytest.mark.parametrize("kind", ["solidity", "standard"])
def test_valid_proof_link(P, py, link_srs, kind):
    cv = py.BN254
    rnd = random.Random(11)
    witness = [rnd.randrange(cv.fr.p) for _ in range(10)]
    # no layout, the same circuit twice (test_valid_proof_link__no_layout)
    p1, h1, lay = _proof_and_hint(P, cv, link_srs, 1, witness, None, kind, 1)
    p2, h2, _ = _proof_and_hint(P, cv, link_srs, 1, witness, None, kind, 2)
    ok, lp = _link_and_verify(P, cv, link_srs, h1, h2, p1, p2, lay, kind)
    assert ok and len(P.serialize_link_proof(cv, lp)) == 64
    # the quotient is the schoolbook one
    assert P.linking_quotient(cv.fr, h1["linking_wire_poly"], h2["linking_wire_poly"], lay) == \
        P.linking_quotient_schoolbook(cv.fr, h1["linking_wire_poly"], h2["linking_wire_poly"], lay)
    # different circuits, the second one laid out like the first (test_valid_proof_link__different_circuits)
    p3, h3, lay3 = _proof_and_hint(P, cv, link_srs, 2, witness, lay, kind, 3)
    assert lay3 == lay
    ok, lp = _link_and_verify(P, cv, link_srs, h1, h3, p1, p3, lay, kind)
    assert ok
    # a tampered linking proof, or the proofs the other way round with a non-zero quotient, must not pass
    bad = dict(lp)
    bad["opening_proof"] = cv.add(lp["opening_proof"], cv.gen)
    assert not P.verify_link_proof(cv, p1, p3, bad, lay, BETA, kind)
    assert not P.verify_link_proof(cv, p3, p1, lp, lay, BETA, kind)
    # identical hints: the empty quotient (proof_linking.rs:122-125), identity commitments
    ok, lp = _link_and_verify(P, cv, link_srs, h1, h1, p1, p1, lay, kind)
    assert ok and lp["quotient_commitment"] is None and lp["opening_proof"] is None


def test_valid_proof_link_specific_layout(P, py, link_srs):
    cv = py.BN254
    rnd = random.Random(12)
    witness = [rnd.randrange(cv.fr.p) for _ in range(10)]
    layout = P.GroupLayout(8, 20, 10)
    p1, h1, l1 = _proof_and_hint(P, cv, link_srs, 1, witness, layout, "solidity", 1)
    p2, h2, l2 = _proof_and_hint(P, cv, link_srs, 2, witness, layout, "solidity", 2)
    assert l1 == layout and l2 == layout
    assert len(h1["linking_wire_poly"]) == 256 + 2
    ok, _ = _link_and_verify(P, cv, link_srs, h1, h2, p1, p2, layout, "solidity")
    assert ok


def test_invalid_proof_links(P, py, link_srs):
    cv = py.BN254
    rnd = random.Random(13)
    w1 = [rnd.randrange(cv.fr.p) for _ in range(10)]
    w2 = list(w1)
    w2[rnd.randrange(10)] = rnd.randrange(cv.fr.p)
    # different witnesses: the same circuit, then different circuits
    p1, h1, lay = _proof_and_hint(P, cv, link_srs, 1, w1, None, "solidity", 1)
    p2, h2, _ = _proof_and_hint(P, cv, link_srs, 1, w2, None, "solidity", 2)
    assert not _link_and_verify(P, cv, link_srs, h1, h2, p1, p2, lay, "solidity")[0]
    p3, h3, _ = _proof_and_hint(P, cv, link_srs, 2, w2, lay, "solidity", 3)
    assert not _link_and_verify(P, cv, link_srs, h1, h3, p1, p3, lay, "solidity")[0]
    # the right witness on a misaligned domain / at another offset
    for bad in (P.GroupLayout(lay.alignment + 1, lay.offset, lay.size), P.GroupLayout(lay.alignment, lay.offset - 1, lay.size)):
        try:
            p4, h4, _ = _proof_and_hint(P, cv, link_srs, 2, w1, bad, "solidity", 4)
        except ValueError:
            continue   # the moved group collides with the public inputs: the reference fails in finalize as well
        assert not _link_and_verify(P, cv, link_srs, h1, h4, p1, p4, bad, "solidity")[0]


def _vanishing_coeffs_model(p, g, offset, s):
    """the recurrences of `Plonk::vanishing_coeffs` (csrc/plonk.cu) in exact integers: q-binomial theorem,
    prod_{i<s} (X - g^(offset+i)) = sum_k (-1)^k [s;k]_g g^(k(k-1)/2 + offset k) X^(s-k)"""
    gp = [1] * (s + 1)
    for k in range(1, s + 1):
        gp[k] = gp[k - 1] * g % p
    den = [0] + [(1 - gp[k]) % p for k in range(1, s + 1)]
    z = [0] * (s + 1)
    z[s] = 1
    binom, e, step = 1, 1, pow(g, offset, p)
    for k in range(1, s + 1):
        binom = binom * den[s - k + 1] % p * pow(den[k], -1, p) % p
        e = e * step % p
        step = step * g % p
        z[s - k] = (-binom * e if k & 1 else binom * e) % p
    return z


@pytest.mark.parametrize("alignment,offset,size", [(4, 0, 1), (4, 3, 2), (5, 11, 10), (8, 20, 10), (8, 1, 254), (10, 700, 300)])
def test_vanishing_polynomial_closed_form(P, py, alignment, offset, size):
    fr = py.BN254_FR
    p = fr.p
    lay = P.GroupLayout(alignment, offset, size)
    g = lay.domain_generator(fr)
    z = [1]
    for r in P._link_roots(fr, lay):   # the reference's product of monomials (proof_linking.rs:137-160)
        nz = [0] * (len(z) + 1)
        for i, c in enumerate(z):
            nz[i] = (nz[i] - c * r) % p
            nz[i + 1] = (nz[i + 1] + c) % p
        z = nz
    assert _vanishing_coeffs_model(p, g, offset, size) == z


@pytest.mark.parametrize("alignment,offset,size,plen", [(4, 3, 5, 40), (5, 11, 10, 34), (6, 2, 50, 300), (8, 20, 10, 258), (3, 1, 6, 9)])
def test_remainder_through_the_fold_then_exact_division(P, py, alignment, offset, size, plen):
    """the algebra of the device's path for dividends that do not vanish on the group (csrc/plonk.cu: `fold`, `poly_mod_monic`):
    Z_D divides X^(2^a) - 1, so p mod Z_D = (p mod (X^(2^a) - 1)) mod Z_D, and floor(p / Z_D) = (p - p mod Z_D) / Z_D exactly."""
    fr = py.BN254_FR
    p = fr.p
    rnd = random.Random(alignment * 1000 + size)
    lay = P.GroupLayout(alignment, offset, size)
    poly = [rnd.randrange(p) for _ in range(plen)]
    z = _vanishing_coeffs_model(p, lay.domain_generator(fr), offset, size)
    m = 1 << alignment
    folded = [0] * m
    for i, c in enumerate(poly):                 # p mod (X^m - 1)
        folded[i % m] = (folded[i % m] + c) % p
    rem = list(folded)
    for k in range(m - 1, size - 1, -1):         # schoolbook reduction mod the monic Z_D, top coefficient first
        c = rem[k]
        if c:
            for j in range(size):
                rem[k - size + j] = (rem[k - size + j] - c * z[j]) % p
    rem = rem[:size]
    # the remainder interpolates p on the group ...
    for r in P._link_roots(fr, lay):
        assert P._poly_eval(p, rem, r) == P._poly_eval(p, poly, r)
    # ... and taking it off makes the division exact, with the floor quotient as its result
    exact = P._poly_add(p, poly, [(-c) % p for c in rem])
    want = P.linking_quotient(fr, poly, [], lay)
    assert want == P.linking_quotient_schoolbook(fr, poly, [], lay)
    assert P.linking_quotient(fr, exact, [], lay) == want
    back = [0] * (len(want) + size)
    for i, a in enumerate(want):
        for j, b in enumerate(z):
            back[i + j] = (back[i + j] + a * b) % p
    assert P._strip(back) == P._strip(exact)
