"""The reference's own KZG tests (primitives/src/pcs/univariate_kzg/mod.rs:407-511: end_to_end,
linear polynomial, batch) restated with a known-beta key: the pairing check collapses to a G1
identity that the oracle evaluates exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _to_affine(py, cv, c):
    if c.infinity:
        return None
    L = cv.fq.limbs64
    x = sum(v << (64 * k) for k, v in enumerate(c.xy[:L]))
    y = sum(v << (64 * k) for k, v in enumerate(c.xy[L:]))
    return (cv.fq.from_mont(x), cv.fq.from_mont(y))


@pytest.mark.parametrize("curve", ["bn254", "bls12_381"])
def test_commit_open_verify(ctx, co, py, curve):
    import mpc_jellyfish_b200 as jf
    cv = py.CURVES[curve]
    fr = cv.fr
    beta = py.random_field_elems(fr, 1, seed=4)[0]
    max_deg = 60
    pp = jf.UnivariateProverParam(ctx.generate_srs_for_testing(curve, beta, max_deg + 1))
    for it in range(12):
        deg = [0, 1, 2, 50, 60, 33, 7, 19, 45, 12, 59, 3][it]
        vals = py.random_field_elems(fr, deg + 1, seed=100 + it)
        if it == 5:
            vals[0] = vals[1] = 0  # low-order zeros exercise the key offset of mod.rs:110
        poly = jf.DensePolynomial(co.ints_to_limbs([fr.to_mont(v) for v in vals], 4))
        comm = jf.UnivariateKzgPCS.commit(pp, poly)
        C = _to_affine(py, cv, comm)
        assert C == cv.mul(py.poly_eval(fr, vals, beta), cv.gen)
        z = py.random_field_elems(fr, 1, seed=200 + it)[0]
        proof, ev = jf.UnivariateKzgPCS.open(pp, poly, z)
        assert ev == py.poly_eval(fr, vals, z)
        assert py.kzg_verify_known_beta(cv, beta, cv.gen, C, z, ev, _to_affine(py, cv, proof))
    pp.key.free()


def test_batch_commit_and_errors(ctx, co, py):
    import mpc_jellyfish_b200 as jf
    cv, fr = py.BN254, py.BN254_FR
    beta = 0x1234567890ABCDEF
    pp = jf.UnivariateProverParam(ctx.generate_srs_for_testing("bn254", beta, 40))
    polys, vals_all = [], []
    for it in range(6):
        vals = py.random_field_elems(fr, [40, 1, 17, 40, 5, 30][it], seed=it)
        vals_all.append(vals)
        polys.append(jf.DensePolynomial(co.ints_to_limbs([fr.to_mont(v) for v in vals], 4)))
    polys.append(jf.DensePolynomial(np.zeros((5, 4), np.uint64)))  # the zero polynomial -> identity
    comms = jf.UnivariateKzgPCS.batch_commit(pp, polys)
    for vals, c in zip(vals_all, comms):
        assert _to_affine(py, cv, c) == cv.mul(py.poly_eval(fr, vals, beta), cv.gen)
    assert comms[-1].infinity
    assert comms == [jf.UnivariateKzgPCS.commit(pp, p) for p in polys]
    too_big = jf.DensePolynomial(co.ints_to_limbs([1] * 42, 4))  # degree 41 > 40 points
    with pytest.raises(jf.InvalidParameters):
        jf.UnivariateKzgPCS.commit(pp, too_big)
    pp.key.free()
