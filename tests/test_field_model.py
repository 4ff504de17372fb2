"""CPU suite: a limb-level model of the even/odd-accumulator Montgomery product of csrc/field.cuh (the chains of
csrc/gen_chains.py, one Python statement per carry-chain instruction) and of its squaring variant (Fp::sqr_dedicated:
row i multiplies a_i with (a_i, 2 a^{>i}) and skips the products with j < i).  Every place where the device code
drops a carry because it "cannot happen" raises here if it does.  This is how the two-spare-bits condition of the
squaring (SQR_OK in field.cuh) was established: BLS12-381 Fr, with one spare bit, does overflow."""
import random

import pytest
M32 = (1 << 32) - 1
class Ovf(Exception): pass

def mad_even(acc, a, b, ci_holder, N, start=0):
    # acc += sum_{j even >= start} a[j]*b << 32 j ; carry -> ci_holder[0] (E[N-1])
    carry = 0
    first = True
    for j in range(0, N, 2):
        if j < start: continue
        prod = a[j] * b
        t = acc[j] + (prod & M32) + carry; acc[j] = t & M32; carry = t >> 32
        t = acc[j + 1] + (prod >> 32) + carry; acc[j + 1] = t & M32; carry = t >> 32
    if any(j >= start for j in range(0, N, 2)):
        t = ci_holder[0] + carry
        if t > M32: raise Ovf("ci overflow")
        ci_holder[0] = t

def mad_odd(acc, a, b, N):
    # acc += sum_{j odd} a[j]*b << 32 (j-1); carry-out must be zero
    carry = 0
    for j in range(1, N, 2):
        prod = a[j] * b
        t = acc[j - 1] + (prod & M32) + carry; acc[j - 1] = t & M32; carry = t >> 32
        t = acc[j] + (prod >> 32) + carry; acc[j] = t & M32; carry = t >> 32
    if carry: raise Ovf("mad_odd carry out")

def shift_mad_odd(x, o, a, b, N, start=0):
    # o[0] += x[1] (carry into chain); x = (x >> 64) + sum_{j odd >= start} a[j]*b << 32 (j-1)
    t = o[0] + x[1]; o[0] = t & M32; carry = t >> 32
    for j in range(1, N, 2):
        k = j - 1
        lo_add = x[k + 2] if k + 2 < N else 0
        hi_add = x[k + 3] if k + 3 < N else 0
        prod = a[j] * b if j >= start else 0
        t = lo_add + (prod & M32) + carry; x[k] = t & M32; carry = t >> 32
        t = hi_add + (prod >> 32) + carry; x[k + 1] = t & M32; carry = t >> 32
    if carry: raise Ovf("shift_mad_odd carry out")

def limbs(v, N): return [(v >> (32 * i)) & M32 for i in range(N)]
def val(l): return sum(x << (32 * i) for i, x in enumerate(l))

def mont(a_l, rows, p_l, inv, N):
    """rows: list of (operand array d, scalar b, start) per row i."""
    x = [0] * N; y = [0] * N
    d, b, _ = rows[0]
    for j in range(0, N, 2):
        e = d[j] * b; o = d[j + 1] * b
        x[j] = e & M32; x[j + 1] = e >> 32; y[j] = o & M32; y[j + 1] = o >> 32
    def mad_p_pair(odd_acc, even_acc, m):
        mad_odd(odd_acc, p_l, m, N)
        h = [odd_acc[N - 1]]; mad_even(even_acc, p_l, m, h, N); odd_acc[N - 1] = h[0]
    m = (x[0] * inv) & M32
    mad_p_pair(y, x, m)
    def row(prev_e, prev_o, d, b, start):
        shift_mad_odd(prev_e, prev_o, d, b, N, start)
        h = [prev_e[N - 1]]; mad_even(prev_o, d, b, h, N, start); prev_e[N - 1] = h[0]
        m = (prev_o[0] * inv) & M32
        mad_p_pair(prev_e, prev_o, m)
        assert prev_o[0] == 0
    for i in range(1, N):
        d, b, start = rows[i]
        if i % 2 == 1: row(x, y, d, b, start)
        else: row(y, x, d, b, start)
    # merge: r = (e >> 32) + o with even acc = y, odd acc = x  (N even)
    r = (val(y) >> 32) + val(x)
    if r >> (32 * N): raise Ovf("merge overflow")
    p = val(p_l)
    if r >= p: r -= p
    if r >= p: raise Ovf("not reduced")
    return r

def mul_rows(a, b, N): return [(a, b[i], 0) for i in range(N)]
def sqr_rows(a, N):
    a2 = [((a[k] << 1) & M32) | ((a[k - 1] >> 31) if k else 0) for k in range(N)]
    rows = []
    for i in range(N):
        d = list(a2); d[i] = a[i]
        if i + 1 < N: d[i + 1] = (a[i + 1] << 1) & M32
        rows.append((d, a[i], i))
    return rows


FIELDS = {
    "bn254_fq": (21888242871839275222246405745257275088696311157297823662689037894645226208583, 8),
    "bn254_fr": (21888242871839275222246405745257275088548364400416034343698204186575808495617, 8),
    "bls12_381_fr": (0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 8),
    "bls12_381_fq": (0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab, 12),
}


def _operands(p, N, count, seed):
    rnd = random.Random(seed)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (1 << (p.bit_length() - 1)) - 1, 1 << (p.bit_length() - 1)]
    edge += [sum(M32 << (32 * i) for i in range(N)) % p, int("ffffffff00000000" * N, 16) % p, int("00000000ffffffff" * N, 16) % p]
    out = [e for e in edge if e < p]
    while len(out) < count:
        if len(out) % 3 == 0:  # mostly saturated limbs
            out.append(val([M32 if rnd.random() < 0.8 else rnd.randrange(1 << 32) for _ in range(N)]) % p)
        else:
            out.append(rnd.randrange(p))
    return out


def _params(name):
    p, N = FIELDS[name]
    return p, N, limbs(p, N), (-pow(p, -1, 1 << 32)) & M32, pow(1 << (32 * N), -1, p)


@pytest.mark.parametrize("name", list(FIELDS))
def test_product_model_is_exact_and_never_drops_a_carry(name):
    p, N, p_l, inv, rinv = _params(name)
    ops = _operands(p, N, 400, 1)
    for a, b in zip(ops, reversed(ops)):
        assert mont(limbs(a, N), mul_rows(limbs(a, N), limbs(b, N), N), p_l, inv, N) == a * b * rinv % p


@pytest.mark.parametrize("name", ["bn254_fq", "bn254_fr", "bls12_381_fq"])
def test_squaring_rows_are_exact_with_two_spare_bits(name):
    p, N, p_l, inv, rinv = _params(name)
    assert p_l[N - 1] < 1 << 30                                         # SQR_OK
    for a in _operands(p, N, 1500, 2):
        assert mont(limbs(a, N), sqr_rows(limbs(a, N), N), p_l, inv, N) == a * a * rinv % p


def test_squaring_rows_overflow_with_one_spare_bit():
    """BLS12-381 Fr (255 bits): the doubled top limb times a 32-bit limb plus the reduction row exceeds 64 bits, so
    Fp<Bls12381Fr>::sqr stays mul(a, a)."""
    p, N, p_l, inv, rinv = _params("bls12_381_fr")
    assert p_l[N - 1] >= 1 << 30
    with pytest.raises(Ovf):
        for a in _operands(p, N, 200, 3):
            mont(limbs(a, N), sqr_rows(limbs(a, N), N), p_l, inv, N)


def mont2(rowsA, rowsB, p_l, inv, N):
    """(a b + c d) / R as Fp::mul_add computes it: every row adds both partial products, then one reduction row."""
    x = [0] * N
    y = [0] * N
    d, b, _ = rowsA[0]
    for j in range(0, N, 2):
        e = d[j] * b
        o = d[j + 1] * b
        x[j], x[j + 1], y[j], y[j + 1] = e & M32, e >> 32, o & M32, o >> 32

    def add_products(odd_acc, even_acc, d, b):
        mad_odd(odd_acc, d, b, N)
        h = [odd_acc[N - 1]]
        mad_even(even_acc, d, b, h, N)
        odd_acc[N - 1] = h[0]

    d2, b2, _ = rowsB[0]
    add_products(y, x, d2, b2)
    add_products(y, x, p_l, (x[0] * inv) & M32)

    def row(prev_e, prev_o, ra, rb):
        shift_mad_odd(prev_e, prev_o, ra[0], ra[1], N)
        h = [prev_e[N - 1]]
        mad_even(prev_o, ra[0], ra[1], h, N)
        prev_e[N - 1] = h[0]
        add_products(prev_e, prev_o, rb[0], rb[1])
        add_products(prev_e, prev_o, p_l, (prev_o[0] * inv) & M32)
        assert prev_o[0] == 0

    for i in range(1, N):
        if i % 2 == 1:
            row(x, y, rowsA[i], rowsB[i])
        else:
            row(y, x, rowsA[i], rowsB[i])
    r = (val(y) >> 32) + val(x)
    if r >> (32 * N):
        raise Ovf("merge overflow")
    p = val(p_l)
    if r >= p:
        r -= p
    if r >= p:
        raise Ovf("not reduced after one subtraction")
    return r


@pytest.mark.parametrize("name", ["bn254_fq", "bn254_fr", "bls12_381_fq"])
def test_two_products_under_one_reduction_are_exact_with_two_spare_bits(name):
    p, N, p_l, inv, rinv = _params(name)
    ops = _operands(p, N, 1200, 4)
    rnd = random.Random(9)
    quads = [(p - 1, p - 1, p - 1, p - 1)] + [tuple(rnd.choice(ops) for _ in range(4)) for _ in range(1200)]
    for a, b, c, d in quads:
        got = mont2(mul_rows(limbs(a, N), limbs(b, N), N), mul_rows(limbs(c, N), limbs(d, N), N), p_l, inv, N)
        assert got == (a * b + c * d) * rinv % p
