"""Collaborative-prover rows: share-wise commit / open / NTT on authenticated
shares.  Mirrors the reference's own MPC tests, which compare the opened multiprover result with the
single-prover result (plonk/src/multiprover/primitives/multiprover_kzg.rs tests `test_commit`, `test_open`;
plonk/src/multiprover/proof_system/prover.rs:1316-1438): two parties hold additive shares of a polynomial
and of its MAC (mac = key * value); each runs the GPU path on its own shares; opening = adding the shares."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _share(co, py, fr, vals, mac_key, seed):
    """additive 2-party shares of `vals` and of mac_key * vals, as Montgomery limb arrays"""
    p = fr.p
    r1 = py.random_field_elems(fr, len(vals), seed=seed)
    r2 = py.random_field_elems(fr, len(vals), seed=seed + 1)
    s = [r1, [(v - a) % p for v, a in zip(vals, r1)]]
    m = [r2, [(mac_key * v - a) % p for v, a in zip(vals, r2)]]
    mont = lambda xs: co.ints_to_limbs([fr.to_mont(x) for x in xs], 4)  # noqa: E731
    return [(mont(s[i]), mont(m[i])) for i in range(2)]


def _pt(co, cv, xy, inf):
    if inf:
        return None
    x, y = co.limbs_to_ints(np.asarray(xy).reshape(2, cv.fq.limbs64))
    return (cv.fq.from_mont(x), cv.fq.from_mont(y))


@pytest.mark.parametrize("curve", ["bn254", "bls12_381"])
def test_multiprover_commit_and_open_match_single_prover(ctx, co, py, curve):
    import mpc_jellyfish_b200 as jf
    cv = py.CURVES[curve]
    fr = cv.fr
    beta, mac_key = 0xABCDEF987654321, 0x1357924680
    deg = 300
    pp = jf.UnivariateProverParam(ctx.generate_srs_for_testing(curve, beta, deg + 1))
    vals = py.random_field_elems(fr, deg + 1, seed=5)
    parties = _share(co, py, fr, vals, mac_key, 50)
    single = jf.UnivariateKzgPCS.commit(pp, jf.DensePolynomial(co.ints_to_limbs([fr.to_mont(v) for v in vals], 4)))
    C = _pt(co, cv, np.array(single.xy, dtype=np.uint64), single.infinity)
    shares = [jf.MultiproverKZG.commit(pp, jf.AuthenticatedDensePoly(s, m)) for s, m in parties]
    opened = cv.add(_pt(co, cv, shares[0].share, shares[0].share_inf), _pt(co, cv, shares[1].share, shares[1].share_inf))
    opened_mac = cv.add(_pt(co, cv, shares[0].mac, shares[0].mac_inf), _pt(co, cv, shares[1].mac, shares[1].mac_inf))
    assert opened == C == cv.mul(py.poly_eval(fr, vals, beta), cv.gen)
    assert opened_mac == cv.mul(mac_key, C)  # the MAC check of ark-mpc's open_authenticated
    # open at a public point
    z = py.random_field_elems(fr, 1, seed=77)[0]
    zl = co.ints_to_limbs([fr.to_mont(z)], 4)[0]
    proof_single, ev_single = jf.UnivariateKzgPCS.open(pp, jf.DensePolynomial(co.ints_to_limbs([fr.to_mont(v) for v in vals], 4)), z)
    outs = [jf.MultiproverKZG.open(pp, jf.AuthenticatedDensePoly(s, m), zl) for s, m in parties]
    proof_opened = cv.add(_pt(co, cv, outs[0][0].share, outs[0][0].share_inf), _pt(co, cv, outs[1][0].share, outs[1][0].share_inf))
    assert proof_opened == _pt(co, cv, np.array(proof_single.xy, dtype=np.uint64), proof_single.infinity)
    ev = sum(fr.from_mont(co.limbs_to_ints(o[1][0][None, :])[0]) for o in outs) % fr.p
    ev_mac = sum(fr.from_mont(co.limbs_to_ints(o[1][1][None, :])[0]) for o in outs) % fr.p
    assert ev == ev_single == py.poly_eval(fr, vals, z) and ev_mac == mac_key * ev % fr.p
    assert py.kzg_verify_known_beta(cv, beta, cv.gen, C, z, ev, proof_opened)
    # batch_commit == per-polynomial commit; degree check as in the reference
    both = jf.MultiproverKZG.batch_commit(pp, [jf.AuthenticatedDensePoly(s, m) for s, m in parties])
    assert all(np.array_equal(a.share, b.share) and np.array_equal(a.mac, b.mac) for a, b in zip(both, shares))
    big = np.zeros((deg + 4, 4), dtype=np.uint64)
    big[:, 0] = 1
    with pytest.raises(jf.InvalidParameters):
        jf.MultiproverKZG.commit(pp, jf.AuthenticatedDensePoly(big, big))
    pp.key.free()


@pytest.mark.parametrize("field", ["bn254_fr", "bls12_381_fr"])
def test_sharewise_ntt_matches_plain_ntt(ctx, co, py, field):
    import mpc_jellyfish_b200 as jf
    fr = py.FIELDS[field]
    log_n = 10
    n = 1 << log_n
    vals = py.random_field_elems(fr, n, seed=8)
    parties = _share(co, py, fr, vals, 0x99887766, 90)
    plain = co.ints_to_limbs([fr.to_mont(v) for v in vals], 4)
    for offset in (None, fr.generator):
        dom = jf.Radix2EvaluationDomain(ctx, field, n) if offset is None else jf.Radix2EvaluationDomain(ctx, field, n).get_coset(offset)
        for inverse in (False, True):
            want = dom.ifft(plain.copy()) if inverse else dom.fft(plain.copy())
            outs = [jf.fft_with_domain(dom, jf.AuthenticatedDensePoly(s, m), inverse=inverse) for s, m in parties]
            got = co.field_op(field, "add", outs[0].share, outs[1].share)
            assert np.array_equal(got, want)
            # MAC shares open to mac_key * result
            mac = co.field_op(field, "add", outs[0].mac, outs[1].mac)
            key_m = co.ints_to_limbs([fr.to_mont(0x99887766)], 4)
            assert np.array_equal(mac, co.field_op(field, "mul", want, np.repeat(key_m, n, axis=0)))
    # round trip on shares, zero-padded input (in_len < n)
    dom8 = jf.Radix2EvaluationDomain(ctx, field, 8 * n).get_coset(fr.generator)
    e = jf.fft_with_domain(dom8, jf.AuthenticatedDensePoly(*parties[0]))
    back = jf.ifft_with_domain(dom8, e)
    assert np.array_equal(back.share[:n], parties[0][0]) and not back.share[n:].any()


def test_kzg_batch_open_and_edge_polynomials(ctx, co, py):
    import mpc_jellyfish_b200 as jf
    cv, fr = py.BN254, py.BN254_FR
    beta = 424242
    pp = jf.UnivariateProverParam(ctx.generate_srs_for_testing("bn254", beta, 70))
    polys, pts, raw = [], [], []
    for it, deg in enumerate([0, 1, 2, 65, 33, 5000 % 60]):
        vals = py.random_field_elems(fr, deg + 1, seed=300 + it)
        raw.append(vals)
        polys.append(jf.DensePolynomial(co.ints_to_limbs([fr.to_mont(v) for v in vals], 4)))
        pts.append(py.random_field_elems(fr, 1, seed=400 + it)[0] if it != 3 else 0)  # also the point 0
    polys.append(jf.DensePolynomial(np.zeros((3, 4), np.uint64)))  # zero polynomial
    raw.append([])
    pts.append(5)
    proofs, evals = jf.UnivariateKzgPCS.batch_open(pp, polys, pts)
    for vals, z, pr, ev in zip(raw, pts, proofs, evals):
        assert ev == py.poly_eval(fr, vals, z)
        want = cv.mul(py.poly_eval(fr, py.poly_div_linear(fr, vals, z), beta), cv.gen) if len(vals) > 1 else None
        assert _pt(co, cv, np.array(pr.xy, dtype=np.uint64), pr.infinity) == want
    with pytest.raises(jf.InvalidParameters):
        jf.UnivariateKzgPCS.batch_open(pp, polys, pts[:-1])
    pp.key.free()


def _fast_shares(co, field, vals_m, key_m, seed, with_modifier):
    """two parties' (share, mac[, public modifier]) vectors of `vals_m` (Montgomery limbs), built with the C oracle so that
    2^18-element vectors take milliseconds: share_1 random, share_0 = v - share_1; likewise for mac = key * (v + modifier)."""
    n = len(vals_m)
    r1 = co.random_field_elems(field, n, seed, True)
    r2 = co.random_field_elems(field, n, seed + 1, True)
    mod = co.random_field_elems(field, n, seed + 2, True) if with_modifier else None
    base = co.field_op(field, "add", vals_m, mod) if with_modifier else vals_m
    mac = co.field_op(field, "mul", base, np.repeat(key_m[None, :], n, axis=0))
    parties = []
    for who in range(2):
        share = r1 if who == 1 else co.field_op(field, "sub", vals_m, r1)
        m = r2 if who == 1 else co.field_op(field, "sub", mac, r2)
        parties.append((share, m, mod))          # the public modifier is public: both parties hold the same vector
    return parties, mac, mod


@pytest.mark.parametrize("three", [False, True], ids=["share+mac", "share+mac+modifier"])
def test_config5_size_commit_and_sharewise_coset_ntt(ctx, co, py, three):
    """BASELINE configs[4]: the collaborative prover at n = 2^18 (BN254).  `MultiproverKZG::batch_commit` of authenticated
    polynomials of n + 2 coefficients and their share-wise coset NTT to the 8n quotient domain: the parties' results, added
    (= opened), equal the single prover's (multiprover/proof_system/prover.rs:1316-1438 assert the same per round), the MAC
    shares open to key * result, and the optional third component goes through the same linear map."""
    import mpc_jellyfish_b200 as jf
    cv, fr, field = py.BN254, py.BN254_FR, "bn254_fr"
    n = 1 << 18
    beta, mac_key = 0x5EEDBEEFCAFE1234567, 0x246813579BDF
    key_m = co.ints_to_limbs([fr.to_mont(mac_key)], 4)[0]
    pp = jf.UnivariateProverParam(ctx.generate_srs_for_testing("bn254", beta, n + 3))
    polys = [co.random_field_elems(field, n + 2, 700 + i, True) for i in range(2)]
    shared = [_fast_shares(co, field, v, key_m, 800 + 10 * i, three) for i, v in enumerate(polys)]
    beta_m = co.ints_to_limbs([fr.to_mont(beta)], 4)[0]

    def commit_of(vec_m):  # known-beta identity: commit(p) = p(beta) G
        ev = co.poly_eval(field, vec_m, beta_m)
        xy = co.fixed_base_mul("bn254", co.field_op(field, "from_mont", ev[None, :]))[0]
        return _pt(co, cv, xy, not xy.any())

    per_party = [jf.MultiproverKZG.batch_commit(pp, [jf.AuthenticatedDensePoly(*sh[0][who]) for sh in shared]) for who in range(2)]
    single = jf.UnivariateKzgPCS.batch_commit(pp, [jf.DensePolynomial(v) for v in polys])
    for i, v in enumerate(polys):
        a, b = per_party[0][i], per_party[1][i]
        opened = cv.add(_pt(co, cv, a.share, a.share_inf), _pt(co, cv, b.share, b.share_inf))
        assert opened == _pt(co, cv, np.array(single[i].xy, dtype=np.uint64), single[i].infinity) == commit_of(v)
        opened_mac = cv.add(_pt(co, cv, a.mac, a.mac_inf), _pt(co, cv, b.mac, b.mac_inf))
        assert opened_mac == commit_of(shared[i][1])
        if three:
            want_mod = commit_of(shared[i][2])
            assert _pt(co, cv, a.public_modifier, a.public_modifier_inf) == want_mod
            assert opened_mac == cv.mul(mac_key, cv.add(opened, want_mod))   # mac = key * (value + modifier)
        else:
            assert a.public_modifier is None and opened_mac == cv.mul(mac_key, opened)
    # share-wise coset NTT to the quotient domain (prover.rs:373-388): opened evaluations == the plain transform
    dom8 = jf.Radix2EvaluationDomain(ctx, field, 8 * n).get_coset(fr.generator)
    v = polys[0]
    want = dom8.fft(v)
    outs = [jf.fft_with_domain(dom8, jf.AuthenticatedDensePoly(*shared[0][0][who])) for who in range(2)]
    assert np.array_equal(co.field_op(field, "add", outs[0].share, outs[1].share), want)
    mac_want = dom8.fft(shared[0][1])
    assert np.array_equal(co.field_op(field, "add", outs[0].mac, outs[1].mac), mac_want)
    if three:
        assert np.array_equal(outs[0].public_modifier, dom8.fft(shared[0][2]))
    # ... and back: the coset iNTT of the opened evaluations returns the coefficients (prover.rs:418)
    back = jf.ifft_with_domain(dom8, outs[1])
    assert np.array_equal(back.share[: n + 2], shared[0][0][1][0]) and not back.share[n + 2:].any()
    # batch_open on shares at size (multiprover_kzg.rs:199-229)
    z = co.ints_to_limbs([fr.to_mont(0x1234567890ABCDEF1234567890)], 4)[0]
    pr = [jf.MultiproverKZG.batch_open(pp, [jf.AuthenticatedDensePoly(*shared[0][0][who])], [z]) for who in range(2)]
    proof_single, ev_single = jf.UnivariateKzgPCS.open(pp, jf.DensePolynomial(v), 0x1234567890ABCDEF1234567890)
    opened_proof = cv.add(_pt(co, cv, pr[0][0][0].share, pr[0][0][0].share_inf), _pt(co, cv, pr[1][0][0].share, pr[1][0][0].share_inf))
    assert opened_proof == _pt(co, cv, np.array(proof_single.xy, dtype=np.uint64), proof_single.infinity)
    ev = co.field_op(field, "add", pr[0][1][0][0][None, :], pr[1][1][0][0][None, :])[0]
    assert fr.from_mont(co.limbs_to_ints(ev[None, :])[0]) == ev_single
    pp.key.free()


@pytest.mark.parametrize("three", [False, True])
def test_collaborative_proof_linking_matches_the_single_prover(ctx, co, py, three):
    """multiprover/proof_system/proof_linking.rs (its own tests open the collaborative linking proof and verify it like a single
    prover's): two parties hold additive shares (+ MAC shares [+ public modifiers]) of two first-wire polynomials that agree on a
    link group; quotient shares, commitment shares, the challenge from the opened commitments, opening shares.  The opened
    (quotient commitment, opening proof) must be the single prover's `LinkingProof` bytes, and the MAC points must check."""
    import mpc_jellyfish_b200 as jf
    cv, fr = py.BN254, py.BN254_FR
    p = fr.p
    beta, mac_key = 0xFEEDFACE1234567, 0x2468ACE
    n, align, offset, size = 256, 7, 9, 20
    pp = jf.UnivariateProverParam(ctx.generate_srs_for_testing("bn254", beta, n + 3))
    a1 = py.random_field_elems(fr, n + 2, seed=91)
    a2 = list(a1)
    for j in (3, 40):                         # a2 = a1 + c_j X^j (X^(2^align) - 1): equal on every 2^align-th root of unity
        c = 1000 + j
        a2[j] = (a2[j] - c) % p
        a2[j + (1 << align)] = (a2[j + (1 << align)] + c) % p
    mont = lambda xs: co.ints_to_limbs([fr.to_mont(x) for x in xs], 4)  # noqa: E731
    pt = lambda xy, inf: _pt(co, cv, xy, inf)                           # noqa: E731
    # single prover
    key = pp.key
    c1 = ctx.msm(key, mont(a1), montgomery=True)
    c2 = ctx.msm(key, mont(a2), montgomery=True)
    layout = jf.GroupLayout(align, offset, size)
    single = jf.PlonkKzgSnark.link_proofs(ctx, key, jf.LinkingHint(mont(a1), c1[0], bool(c1[1])), jf.LinkingHint(mont(a2), c2[0], bool(c2[1])),
                                          layout, "solidity")
    assert single.path == 0
    # two parties
    def hints(vals, seed):
        out = []
        for s, m in _share(co, py, fr, vals, mac_key, seed):
            mod = None
            if three:   # public modifiers: zero vectors on both parties except one public constant folded into party 0's view
                mod = np.zeros_like(s)
            poly = jf.AuthenticatedDensePoly(s, m, mod)
            out.append(jf.MpcLinkingHint(poly, jf.MultiproverKZG.commit(pp, poly)))
        return out
    h1, h2 = hints(a1, 300), hints(a2, 400)
    opened = lambda shares, which: cv.add(*[pt(getattr(s, which), getattr(s, which + "_inf")) for s in shares])  # noqa: E731
    a1c, a2c = opened([h.linking_wire_comm for h in h1], "share"), opened([h.linking_wire_comm for h in h2], "share")
    assert a1c == pt(*c1) and a2c == pt(*c2)
    qs = [jf.MultiproverLinking.quotient(pp, h1[i], h2[i], layout) for i in range(2)]
    assert len(qs[0][0].share) == n + 2 - size
    qc, qc_mac = opened([q[1] for q in qs], "share"), opened([q[1] for q in qs], "mac")
    assert qc == pt(single.quotient_commitment, single.quotient_inf)
    assert qc_mac == cv.mul(mac_key, qc)
    # the shares of the quotient do not vanish on the group one by one, their sum does: check the opened polynomial too
    qsum = [(x + y) % p for x, y in zip(*[[fr.from_mont(v) for v in co.limbs_to_ints(q[0].share)] for q in qs])]
    import plonk_ref as P
    assert P._strip(qsum) == P.linking_quotient(fr, a1, a2, P.GroupLayout(align, offset, size))
    # public challenge from the opened points, then the opening shares
    to_xy = lambda P_: (np.concatenate([co.ints_to_limbs([cv.fq.to_mont(P_[0])], 4)[0], co.ints_to_limbs([cv.fq.to_mont(P_[1])], 4)[0]]), False)  # noqa: E731
    eta = jf.MultiproverLinking.challenge("bn254", to_xy(a1c), to_xy(a2c), to_xy(qc))
    assert np.array_equal(eta, single.eta)
    eta_i = fr.from_mont(co.limbs_to_ints(eta.reshape(1, 4))[0])
    zd = P._link_vanishing_eval(fr, eta_i, P.GroupLayout(align, offset, size))
    zd_l = co.ints_to_limbs([fr.to_mont(zd)], 4)[0]
    proofs = [jf.MultiproverLinking.identity_opening(pp, h1[i], h2[i], qs[i][0], eta, zd_l) for i in range(2)]
    w, w_mac = opened(proofs, "share"), opened(proofs, "mac")
    assert w == pt(single.opening_proof, single.opening_inf)
    assert w_mac == cv.mul(mac_key, w)
    # ... and the restated verifier accepts the opened proof
    ok = P.verify_link_proof(cv, {"wires_poly_comms": [a1c]}, {"wires_poly_comms": [a2c]},
                             {"quotient_commitment": qc, "opening_proof": w}, P.GroupLayout(align, offset, size), beta, "solidity")
    assert ok
    pp.key.free()
