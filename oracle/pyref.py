"""Big-int CPU oracle for the KZG-commit MSM and radix-2 NTT hot path (TEST INFRASTRUCTURE ONLY).

This file is a *checker*.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``mpc-jellyfish_b200``) never does.

PARITY UNPINNED by the reference itself: mpc-jellyfish holds no golden vectors or
known-answer tests for MSM / NTT / commitments, and its arithmetic
lives in un-vendored crates (ark-ff / ark-ec / ark-poly 0.4.x, ark-bn254 /
ark-bls12-381 0.4.0) that cannot be compiled here (no Rust toolchain).  The oracle
is therefore a restatement of the *published* arkworks semantics, anchored on

* the reference's call sites
    - ``primitives/src/pcs/univariate_kzg/mod.rs:90-116``  (commit -> msm_bigint -> into_affine)
    - ``primitives/src/pcs/univariate_kzg/mod.rs:379-395`` (skip low-order zeros, into_bigint)
    - ``primitives/src/pcs/univariate_kzg/mod.rs:135-161`` (open: p/(X-z), MSM, Horner)
    - ``primitives/src/pcs/univariate_kzg/srs.rs:118-153``  (test SRS = [beta^i] g)
    - ``plonk/src/proof_system/prover.rs:54-62,545-567,672`` (domains, coset fft / ifft)
    - ``relation/src/constraint_system.rs:1162-1259``        (ifft call sites)
* public constants every implementation of these curves shares (moduli, generators,
  2-adic roots of unity, Montgomery R) -- pinned in ``tests/golden/constants.json``
  and re-derived numerically in ``tests/test_oracle.py``;
* public known-answer points (EIP-196 BN254 2*G, the BLS12-381 G1 generator).

Everything is exact Python ``int`` arithmetic; use it for N <= 2^12 (MSM) and
n <= 2^14 (NTT).  Larger sizes go through ``oracle/jf_oracle.c``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------
# Fields  (ark-bn254 0.4 / ark-bls12-381 0.4 `FrConfig` / `FqConfig`)
# --------------------------------------------------------------------------------------


@dataclass(frozen=True)
class Field:
    name: str
    p: int
    limbs64: int  # number of u64 limbs of ark-ff's BigInt<N>
    generator: int  # `GENERATOR` (multiplicative generator used for cosets)
    two_adicity: int

    @property
    def bits(self) -> int:
        return self.p.bit_length()

    @property
    def R(self) -> int:  # Montgomery radix 2^(64 N) mod p
        return pow(2, 64 * self.limbs64, self.p)

    @property
    def R2(self) -> int:
        return pow(2, 128 * self.limbs64, self.p)

    @property
    def inv64(self) -> int:  # -p^-1 mod 2^64  (ark-ff `INV`)
        return (-pow(self.p, -1, 1 << 64)) % (1 << 64)

    @property
    def inv32(self) -> int:
        return (-pow(self.p, -1, 1 << 32)) % (1 << 32)

    @property
    def two_adic_root(self) -> int:  # `TWO_ADIC_ROOT_OF_UNITY` = g^((p-1)/2^s)
        return pow(self.generator, (self.p - 1) >> self.two_adicity, self.p)

    # Montgomery helpers (ark-ff keeps `Fp(BigInt)` in Montgomery form in memory)
    def to_mont(self, a: int) -> int:
        return a * self.R % self.p

    def from_mont(self, a: int) -> int:
        return a * pow(self.R, -1, self.p) % self.p

    def to_limbs(self, a: int) -> List[int]:
        return [(a >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(self.limbs64)]

    def from_limbs(self, l: Sequence[int]) -> int:
        return sum(int(x) << (64 * i) for i, x in enumerate(l))

    def inv(self, a: int) -> int:
        return pow(a, -1, self.p)


BN254_FR = Field(
    "bn254_fr",
    21888242871839275222246405745257275088548364400416034343698204186575808495617,
    4, 5, 28)
BN254_FQ = Field(
    "bn254_fq",
    21888242871839275222246405745257275088696311157297823662689037894645226208583,
    4, 3, 1)
BLS12_381_FR = Field(
    "bls12_381_fr",
    0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    4, 7, 32)
BLS12_381_FQ = Field(
    "bls12_381_fq",
    0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    6, 2, 1)

FIELDS = {f.name: f for f in (BN254_FR, BN254_FQ, BLS12_381_FR, BLS12_381_FQ)}

# --------------------------------------------------------------------------------------
# Curves: short Weierstrass y^2 = x^3 + b (a = 0) G1 groups
# --------------------------------------------------------------------------------------

Affine = Optional[Tuple[int, int]]  # None == point at infinity


@dataclass(frozen=True)
class Curve:
    name: str
    fq: Field
    fr: Field
    b: int
    gen: Tuple[int, int]

    def is_on_curve(self, P: Affine) -> bool:
        if P is None:
            return True
        x, y = P
        p = self.fq.p
        return (y * y - x * x * x - self.b) % p == 0

    def neg(self, P: Affine) -> Affine:
        if P is None:
            return None
        return (P[0], (-P[1]) % self.fq.p)

    def add(self, P: Affine, Q: Affine) -> Affine:
        p = self.fq.p
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        y3 = (lam * (x1 - x3) - y1) % p
        return (x3, y3)

    # Jacobian arithmetic for speed (one inversion at the end)
    def _jdbl(self, P):
        X, Y, Z = P
        p = self.fq.p
        if Z == 0:
            return P
        A = X * X % p
        B = Y * Y % p
        C = B * B % p
        D = 2 * ((X + B) * (X + B) - A - C) % p
        E = 3 * A % p
        F = E * E % p
        X3 = (F - 2 * D) % p
        Y3 = (E * (D - X3) - 8 * C) % p
        Z3 = 2 * Y * Z % p
        return (X3, Y3, Z3)

    def _jadd(self, P, Q):
        p = self.fq.p
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        if Z1 == 0:
            return Q
        if Z2 == 0:
            return P
        Z1Z1 = Z1 * Z1 % p
        Z2Z2 = Z2 * Z2 % p
        U1 = X1 * Z2Z2 % p
        U2 = X2 * Z1Z1 % p
        S1 = Y1 * Z2 * Z2Z2 % p
        S2 = Y2 * Z1 * Z1Z1 % p
        if U1 == U2:
            if S1 == S2:
                return self._jdbl(P)
            return (1, 1, 0)
        H = (U2 - U1) % p
        I = 4 * H * H % p
        J = H * I % p
        r = 2 * (S2 - S1) % p
        V = U1 * I % p
        X3 = (r * r - J - 2 * V) % p
        Y3 = (r * (V - X3) - 2 * S1 * J) % p
        Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % p
        return (X3, Y3, Z3)

    def _to_jac(self, P: Affine):
        return (1, 1, 0) if P is None else (P[0], P[1], 1)

    def _to_affine(self, P) -> Affine:
        X, Y, Z = P
        p = self.fq.p
        if Z == 0:
            return None
        zi = pow(Z, -1, p)
        zi2 = zi * zi % p
        return (X * zi2 % p, Y * zi2 * zi % p)

    def mul(self, k: int, P: Affine) -> Affine:
        """k * P, double-and-add on Jacobian coordinates."""
        k %= self.fr.p
        acc = (1, 1, 0)
        base = self._to_jac(P)
        while k:
            if k & 1:
                acc = self._jadd(acc, base)
            base = self._jdbl(base)
            k >>= 1
        return self._to_affine(acc)

    def msm_naive(self, scalars: Sequence[int], points: Sequence[Affine]) -> Affine:
        """sum_i s_i * P_i over min(len) pairs -- the *value* `msm_bigint(..).into_affine()`
        returns (univariate_kzg/mod.rs:110).  O(N * 256) group ops; small N only."""
        n = min(len(scalars), len(points))
        acc = (1, 1, 0)
        for i in range(n):
            if scalars[i] == 0 or points[i] is None:
                continue
            acc = self._jadd(acc, self._to_jac(self.mul(scalars[i], points[i])))
        return self._to_affine(acc)

    def msm_pippenger(self, scalars: Sequence[int], points: Sequence[Affine], c: int = 8) -> Affine:
        """Plain (unsigned-window) bucket method, used to cross-check msm_naive at
        a few thousand points."""
        n = min(len(scalars), len(points))
        bits = self.fr.bits
        W = (bits + c - 1) // c
        total = (1, 1, 0)
        for w in reversed(range(W)):
            for _ in range(c):
                total = self._jdbl(total)
            buckets = [(1, 1, 0)] * (1 << c)
            for i in range(n):
                d = (scalars[i] >> (w * c)) & ((1 << c) - 1)
                if d and points[i] is not None:
                    buckets[d] = self._jadd(buckets[d], self._to_jac(points[i]))
            run = (1, 1, 0)
            acc = (1, 1, 0)
            for d in range((1 << c) - 1, 0, -1):
                run = self._jadd(run, buckets[d])
                acc = self._jadd(acc, run)
            total = self._jadd(total, acc)
        return self._to_affine(total)

    # ---- `serialize_compressed` of a G1 affine point (`to_bytes!`, utilities/src/macros.rs:13-18) ----
    # BN254 (ark-bn254 0.4.0) uses ark-ec's generic short-Weierstrass encoding: x little-endian, flags in the top two bits of
    # the LAST byte (bit 7: y is the lexicographically larger root, bit 6: infinity).
    # BLS12-381 (ark-bls12-381 0.4.0, the version the reference pins in primitives/Cargo.toml:13) overrides
    # `SWCurveConfig::serialize_with_mode` for G1 with the ZCash / IETF encoding (its `curves/util.rs`: `EncodingFlags`,
    # `serialize_fq`): x BIG-endian, flags in the top three bits of the FIRST byte (bit 7: compressed, bit 6: infinity,
    # bit 5: y is the larger root, only when compressed and finite).  SURVEY.md 8c described the generic form for both
    # curves; that was wrong for BLS12-381 (VERDICT r1, weak #1b).  Pinned by the published compressed generator, 2 G and
    # identity in tests/golden/constants.json.
    def serialize_compressed(self, P: Affine) -> bytes:
        nbytes = self.fq.limbs64 * 8
        if self.name == "bls12_381":
            if P is None:
                return bytes([0xC0]) + bytes(nbytes - 1)
            x, y = P
            out = bytearray(x.to_bytes(nbytes, "big"))
            out[0] |= 0x80
            if y > (self.fq.p - y):
                out[0] |= 0x20
            return bytes(out)
        if P is None:
            out = bytearray(nbytes)
            out[-1] |= 0x40
            return bytes(out)
        x, y = P
        out = bytearray(x.to_bytes(nbytes, "little"))
        if y > (self.fq.p - y):  # "negative" flag: y is the lexicographically larger root
            out[-1] |= 0x80
        return bytes(out)


BN254 = Curve("bn254", BN254_FQ, BN254_FR, 3, (1, 2))
BLS12_381 = Curve(
    "bls12_381", BLS12_381_FQ, BLS12_381_FR, 4,
    (0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
     0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1))

CURVES = {c.name: c for c in (BN254, BLS12_381)}

# --------------------------------------------------------------------------------------
# Radix-2 evaluation domains  (ark-poly 0.4.2 `Radix2EvaluationDomain`)
# --------------------------------------------------------------------------------------


class Radix2Domain:
    """`Radix2EvaluationDomain::new(k)` (+ `get_coset(offset)`).

    size = next_pow2(k); group_gen = two_adic_root^(2^(two_adicity - log2 size));
    `element(i)` = offset * group_gen^i.  Creation fails (ValueError here,
    `PlonkError::DomainCreationError` there) when log2(size) > two_adicity.
    """

    def __init__(self, field: Field, num_coeffs: int, offset: int = 1):
        size = 1
        log = 0
        while size < num_coeffs:
            size <<= 1
            log += 1
        if log > field.two_adicity:
            raise ValueError("domain too large for the field's two-adicity")
        self.f = field
        self.size = size
        self.log_size = log
        self.group_gen = pow(field.two_adic_root, 1 << (field.two_adicity - log), field.p)
        self.group_gen_inv = pow(self.group_gen, -1, field.p)
        self.size_inv = pow(size, -1, field.p)
        self.offset = offset % field.p
        self.offset_inv = pow(self.offset, -1, field.p)

    def get_coset(self, offset: int) -> "Radix2Domain":
        return Radix2Domain(self.f, self.size, offset)

    def element(self, i: int) -> int:
        return self.offset * pow(self.group_gen, i, self.f.p) % self.f.p

    # -- O(n^2) definitions ------------------------------------------------------------
    def fft_naive(self, coeffs: Sequence[int]) -> List[int]:
        p = self.f.p
        assert len(coeffs) <= self.size
        out = []
        for i in range(self.size):
            x = self.element(i)
            acc = 0
            for c in reversed(coeffs):
                acc = (acc * x + c) % p
            out.append(acc)
        return out

    # -- O(n log n) -------------------------------------------------------------------
    def _ntt(self, a: List[int], root: int) -> List[int]:
        p = self.f.p
        n = self.size
        a = list(a)
        # bit reversal then Cooley-Tukey DIT; natural-order output
        j = 0
        for i in range(1, n):
            bit = n >> 1
            while j & bit:
                j ^= bit
                bit >>= 1
            j |= bit
            if i < j:
                a[i], a[j] = a[j], a[i]
        length = 2
        while length <= n:
            w_len = pow(root, n // length, p)
            half = length >> 1
            tw = [1] * half
            for k in range(1, half):
                tw[k] = tw[k - 1] * w_len % p
            for s in range(0, n, length):
                for k in range(half):
                    u = a[s + k]
                    v = a[s + k + half] * tw[k] % p
                    a[s + k] = (u + v) % p
                    a[s + k + half] = (u - v) % p
            length <<= 1
        return a

    def fft(self, coeffs: Sequence[int]) -> List[int]:
        """out[i] = sum_j c[j] (offset * gen^i)^j, natural order; input zero-padded to size."""
        p = self.f.p
        assert len(coeffs) <= self.size
        a = list(coeffs) + [0] * (self.size - len(coeffs))
        if self.offset != 1:
            g = 1
            for j in range(len(coeffs)):
                a[j] = a[j] * g % p
                g = g * self.offset % p
        return self._ntt(a, self.group_gen)

    def ifft(self, evals: Sequence[int]) -> List[int]:
        """c[j] = offset^-j * size^-1 * sum_i e[i] gen^(-i j)."""
        p = self.f.p
        assert len(evals) <= self.size
        a = list(evals) + [0] * (self.size - len(evals))
        a = self._ntt(a, self.group_gen_inv)
        g = self.size_inv
        out = []
        for j in range(self.size):
            out.append(a[j] * g % p)
            g = g * self.offset_inv % p
        return out


# --------------------------------------------------------------------------------------
# Polynomial helpers and KZG (primitives/src/pcs/univariate_kzg/{mod,srs}.rs)
# --------------------------------------------------------------------------------------


def poly_strip(coeffs: Sequence[int]) -> List[int]:
    """`DensePolynomial::from_coefficients_vec` strips trailing (high-degree) zeros."""
    c = list(coeffs)
    while c and c[-1] == 0:
        c.pop()
    return c


def poly_eval(field: Field, coeffs: Sequence[int], x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % field.p
    return acc


def poly_div_linear(field: Field, coeffs: Sequence[int], z: int) -> List[int]:
    """Quotient of p(X) / (X - z) (remainder dropped), as `open` does at mod.rs:142-145."""
    p = field.p
    c = poly_strip(coeffs)
    if len(c) <= 1:
        return []
    q = [0] * (len(c) - 1)
    carry = 0
    for i in range(len(c) - 1, 0, -1):
        carry = (c[i] + carry * z) % p
        q[i - 1] = carry
    return poly_strip(q)


def gen_srs_for_testing(curve: Curve, beta: int, max_degree: int, g: Affine = None) -> List[Affine]:
    """powers_of_g[i] = beta^i * g, i in 0..=max_degree  (srs.rs:118-153; known beta)."""
    g = curve.gen if g is None else g
    out = []
    cur = 1
    for _ in range(max_degree + 1):
        out.append(curve.mul(cur, g))
        cur = cur * beta % curve.fr.p
    return out


def kzg_commit(curve: Curve, powers_of_g: Sequence[Affine], coeffs: Sequence[int]) -> Affine:
    """`UnivariateKzgPCS::commit` (mod.rs:90-116): degree check, skip *low-order* zero
    coefficients (mod.rs:379-388), MSM over the offset SRS slice, into_affine."""
    c = poly_strip(coeffs)
    degree = max(len(c) - 1, 0)
    if degree > len(powers_of_g):
        raise ValueError("poly degree %d is larger than allowed %d" % (degree, len(powers_of_g)))
    nz = 0
    while nz < len(c) and c[nz] == 0:
        nz += 1
    return curve.msm_naive(c[nz:], powers_of_g[nz:])


def kzg_open(curve: Curve, powers_of_g: Sequence[Affine], coeffs: Sequence[int], z: int):
    """`UnivariateKzgPCS::open` (mod.rs:135-161) -> (proof point, evaluation)."""
    w = poly_div_linear(curve.fr, coeffs, z)
    nz = 0
    while nz < len(w) and w[nz] == 0:
        nz += 1
    proof = curve.msm_naive(w[nz:], powers_of_g[nz:])
    return proof, poly_eval(curve.fr, poly_strip(coeffs), z)


def kzg_verify_known_beta(curve: Curve, beta: int, g: Affine, comm: Affine, z: int, value: int,
                          proof: Affine) -> bool:
    """The pairing check e(C - v g, h) == e(pi, (beta - z) h) of `verify` (mod.rs:195-217)
    collapses, for a *known* beta, to the G1 identity  C - v*g == (beta - z) * pi."""
    lhs = curve.add(comm, curve.neg(curve.mul(value, g)))
    rhs = curve.mul((beta - z) % curve.fr.p, proof)
    return lhs == rhs


# --------------------------------------------------------------------------------------
# Deterministic input generation shared by tests, bench and the C oracle
# (SplitMix64, seed 0x6a656c6c79666973 "jellyfis")
# --------------------------------------------------------------------------------------

SEED = 0x6A656C6C79666973
_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int = SEED):
        self.s = seed & _M64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)


def random_field_elems(field: Field, n: int, seed: int = SEED) -> List[int]:
    """Uniform canonical values in [0, p) by rejection: draw limbs64 words (little-endian),
    mask to the modulus bit length, retry while >= p."""
    rng = SplitMix64(seed)
    mask = (1 << field.bits) - 1
    out = []
    while len(out) < n:
        v = 0
        for i in range(field.limbs64):
            v |= rng.next() << (64 * i)
        v &= mask
        if v < field.p:
            out.append(v)
    return out
