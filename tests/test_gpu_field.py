"""K1 parity: device Montgomery arithmetic vs the CPU oracle, bit-exact on the limbs."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FIELDS = ["bn254_fr", "bn254_fq", "bls12_381_fr", "bls12_381_fq"]


def _vectors(py, fname, n=4096, seed=11):
    f = py.FIELDS[fname]
    rnd = random.Random(seed)
    edge = [0, 1, 2, f.p - 1, f.p - 2, f.R, f.R2, f.p >> 1, (1 << (f.bits - 1)) % f.p, f.inv32]
    a = [x for x in edge for _ in edge] + [rnd.randrange(f.p) for _ in range(n)]
    b = [y for _ in edge for y in edge] + [rnd.randrange(f.p) for _ in range(n)]
    return f, a, b


@pytest.mark.parametrize("fname", FIELDS)
@pytest.mark.parametrize("op", ["mul", "add", "sub", "sqr", "neg", "to_mont", "from_mont"])
def test_field_op_matches_oracle(ctx, co, py, fname, op):
    f, a, b = _vectors(py, fname)
    A = co.ints_to_limbs(a, f.limbs64)
    B = co.ints_to_limbs(b, f.limbs64)
    got = ctx.field_op(fname, op, A, B if op in ("mul", "add", "sub") else None)
    want = co.field_op(fname, op, A, B if op in ("mul", "add", "sub") else None)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("fname", FIELDS)
def test_field_mul_matches_bigint(ctx, co, py, fname):
    f, a, b = _vectors(py, fname, n=512, seed=5)
    got = co.limbs_to_ints(ctx.field_op(fname, "mul", co.ints_to_limbs(a, f.limbs64), co.ints_to_limbs(b, f.limbs64)))
    rinv = pow(f.R, -1, f.p)
    assert got == [x * y * rinv % f.p for x, y in zip(a, b)]


@pytest.mark.parametrize("fname", FIELDS)
def test_field_inverse(ctx, co, py, fname):
    f, a, _ = _vectors(py, fname, n=256, seed=9)
    A = co.ints_to_limbs(a, f.limbs64)
    got = ctx.field_op(fname, "inv", A)
    assert np.array_equal(got, co.field_op(fname, "inv", A))
    prod = co.limbs_to_ints(ctx.field_op(fname, "mul", A, got))
    assert all(p == (f.R if x else 0) for p, x in zip(prod, a))
