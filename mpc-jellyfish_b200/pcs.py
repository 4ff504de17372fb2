"""Host-side mirror of `UnivariateKzgPCS` for the commit / batch_commit / open path
(primitives/src/pcs/univariate_kzg/mod.rs:90-161), same names, argument meaning and error
behaviour; the MSM runs on the GPU through the C ABI.

Polynomials are `DensePolynomial`s: (n, 4) uint64 coefficient arrays in ark-ff's Montgomery
layout, low degree first.  Commitments are affine G1 points `(x || y limbs, infinity)`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from .context import CommitKey, Context
from .errors import InvalidParameters
from .fields import MODULUS, array_to_ints, from_mont, ints_to_array, to_mont
from ._ffi import CURVE_FR


class DensePolynomial:
    """`ark_poly::univariate::DensePolynomial`: trailing (high-degree) zero coefficients are
    stripped on construction (`from_coefficients_vec`)."""

    def __init__(self, coeffs: np.ndarray):
        c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        nz = np.flatnonzero(c.any(axis=1))
        self.coeffs = c[: (nz[-1] + 1 if nz.size else 0)]

    def degree(self) -> int:
        return max(len(self.coeffs) - 1, 0)

    def __len__(self):
        return len(self.coeffs)


@dataclass
class UnivariateProverParam:
    """`UnivariateProverParam { powers_of_g }` (srs.rs:34-38), resident on the GPU."""
    key: CommitKey

    @property
    def ctx(self) -> Context:
        return self.key.ctx


@dataclass(frozen=True)
class Commitment:
    """`Commitment(E::G1Affine)` (structs.rs:16-19)."""
    xy: Tuple[int, ...]
    infinity: bool

    @staticmethod
    def from_raw(xy: np.ndarray, inf: bool) -> "Commitment":
        return Commitment(tuple(int(v) for v in xy), bool(inf))


def _skip_leading_zeros(poly: DensePolynomial) -> int:
    """`skip_leading_zeros_and_convert_to_bigints` (mod.rs:379-388): number of LOW-order zero
    coefficients; the Montgomery -> canonical conversion is done on the GPU."""
    nz = np.flatnonzero(poly.coeffs.any(axis=1))
    return int(nz[0]) if nz.size else len(poly.coeffs)


class UnivariateKzgPCS:
    """`UnivariateKzgPCS<E>`: commit / batch_commit / open."""

    @staticmethod
    def commit(prover_param: UnivariateProverParam, poly: DensePolynomial) -> Commitment:
        n_pts = len(prover_param.key)
        if poly.degree() > n_pts:  # mod.rs:98-104
            raise InvalidParameters("poly degree %d is larger than allowed %d" % (poly.degree(), n_pts))
        nz = _skip_leading_zeros(poly)
        xy, inf = prover_param.ctx.msm(prover_param.key, poly.coeffs[nz:], base_offset=nz, montgomery=True)
        return Commitment.from_raw(xy, inf)

    @staticmethod
    def batch_commit(prover_param: UnivariateProverParam, polys: Sequence[DensePolynomial]) -> List[Commitment]:
        n_pts = len(prover_param.key)
        for p in polys:
            if p.degree() > n_pts:
                raise InvalidParameters("poly degree %d is larger than allowed %d" % (p.degree(), n_pts))
        offs = [_skip_leading_zeros(p) for p in polys]
        out, infs = prover_param.ctx.msm_batch(prover_param.key, [p.coeffs[o:] for p, o in zip(polys, offs)], offs,
                                               montgomery=True)
        return [Commitment.from_raw(out[i], infs[i]) for i in range(len(polys))]

    @staticmethod
    def open(prover_param: UnivariateProverParam, polynomial: DensePolynomial, point: int):
        """-> (proof commitment, evaluation as a canonical int).  `point` is a canonical int.
        Witness polynomial p / (X - z), its commitment and the evaluation all run on the GPU
        (`jf_kzg_open`; mod.rs:135-161)."""
        field = CURVE_FR[prover_param.key.curve]
        z = ints_to_array([to_mont(field, point % MODULUS[field])], 4)
        xy, inf, ev = prover_param.ctx.kzg_open(prover_param.key, [polynomial.coeffs], z)
        return Commitment.from_raw(xy[0], inf[0]), from_mont(field, array_to_ints(ev)[0])

    @staticmethod
    def batch_open(prover_param: UnivariateProverParam, polynomials: Sequence[DensePolynomial], points: Sequence[int]):
        """`batch_open` (mod.rs:165-194): one opening per (polynomial, point) pair; one library call."""
        if len(polynomials) != len(points):  # mod.rs:171-177
            raise InvalidParameters("poly length %d is different from points length %d" % (len(polynomials), len(points)))
        field = CURVE_FR[prover_param.key.curve]
        z = ints_to_array([to_mont(field, pt % MODULUS[field]) for pt in points], 4)
        xy, inf, ev = prover_param.ctx.kzg_open(prover_param.key, [p.coeffs for p in polynomials], z)
        return ([Commitment.from_raw(xy[i], inf[i]) for i in range(len(points))],
                [from_mont(field, v) for v in array_to_ints(ev)])
