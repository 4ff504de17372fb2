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
