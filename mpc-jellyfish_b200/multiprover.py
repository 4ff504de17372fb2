"""Host-side mirror of the collaborative (MPC) prover's share-wise use of the two kernels
(SURVEY §8 rows a13 / a14):

  MultiproverKZG::{commit, batch_commit, open}  plonk/src/multiprover/primitives/multiprover_kzg.rs:128-197
  fft_with_domain / ifft_with_domain            plonk/src/multiprover/proof_system/prover.rs:373-388,418
  AuthenticatedScalarResult::ifft               plonk/src/multiprover/proof_system/constraint_system.rs:930,955,978

An authenticated value is an additive share plus an additive share of its MAC (ark-mpc
`AuthenticatedScalarResult { share, mac, .. }`).  MSM and NTT are linear, so each party applies them
to its share vector and to its MAC vector locally -- here as one batch of two on the GPU -- and the
results are valid shares of the plain result.  Opening shares (the network exchange of ark-mpc) is
not on this path; `open_shares` below is the arithmetic the reference's tests use to check results.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from .context import Context
from .domain import Radix2EvaluationDomain
from .errors import InvalidParameters
from .pcs import UnivariateProverParam


@dataclass
class AuthenticatedDensePoly:
    """`AuthenticatedDensePoly`: coefficient shares and their MAC shares, (n, 4) Montgomery limbs each."""
    share: np.ndarray
    mac: np.ndarray

    def __post_init__(self):
        self.share = np.ascontiguousarray(self.share, dtype=np.uint64).reshape(-1, 4)
        self.mac = np.ascontiguousarray(self.mac, dtype=np.uint64).reshape(-1, 4)
        if self.share.shape != self.mac.shape:
            raise InvalidParameters("share and MAC vectors differ in length")

    def degree(self) -> int:
        return max(len(self.share) - 1, 0)


@dataclass
class AuthenticatedPointShare:
    """One party's share of a commitment / opening proof: (share point, MAC point), affine x || y."""
    share: np.ndarray
    share_inf: bool
    mac: np.ndarray
    mac_inf: bool


class MultiproverKZG:
    @staticmethod
    def commit(prover_params: UnivariateProverParam, poly: AuthenticatedDensePoly) -> AuthenticatedPointShare:
        return MultiproverKZG.batch_commit(prover_params, [poly])[0]

    @staticmethod
    def batch_commit(prover_params: UnivariateProverParam, polys: Sequence[AuthenticatedDensePoly]) -> List[AuthenticatedPointShare]:
        n_pts = len(prover_params.key)
        for p in polys:
            if p.degree() > n_pts:  # multiprover_kzg.rs:132-136
                raise InvalidParameters("Polynomial degree exceeds supported degree")
        vecs = []
        for p in polys:
            vecs += [p.share, p.mac]
        out, inf = prover_params.ctx.msm_batch(prover_params.key, vecs, None, montgomery=True)
        return [AuthenticatedPointShare(out[2 * i], inf[2 * i], out[2 * i + 1], inf[2 * i + 1]) for i in range(len(polys))]

    @staticmethod
    def open(prover_params: UnivariateProverParam, poly: AuthenticatedDensePoly, point: np.ndarray
             ) -> Tuple[AuthenticatedPointShare, Tuple[np.ndarray, np.ndarray]]:
        """`point` is public (4 Montgomery limbs).  -> (proof share, (evaluation share, evaluation MAC share))."""
        if poly.degree() - 1 > len(prover_params.key):  # multiprover_kzg.rs:176-181
            raise InvalidParameters("Polynomial degree exceeds supported degree")
        z = np.ascontiguousarray(point, dtype=np.uint64).reshape(1, 4)
        xy, inf, ev = prover_params.ctx.kzg_open(prover_params.key, [poly.share, poly.mac], np.repeat(z, 2, axis=0))
        return AuthenticatedPointShare(xy[0], inf[0], xy[1], inf[1]), (ev[0], ev[1])


def fft_with_domain(domain: Radix2EvaluationDomain, poly: AuthenticatedDensePoly, inverse: bool = False) -> AuthenticatedDensePoly:
    """Share-wise (coset) NTT / iNTT of an authenticated vector: one batch of two on the GPU."""
    n = domain.size
    if len(poly.share) > n:
        raise InvalidParameters("input of length %d exceeds the domain size %d" % (len(poly.share), n))
    buf = np.zeros((2, n, 4), dtype=np.uint64)
    buf[0, : len(poly.share)] = poly.share
    buf[1, : len(poly.mac)] = poly.mac
    out = domain.batch_fft(buf, inverse=inverse, in_len=len(poly.share))
    return AuthenticatedDensePoly(out[0], out[1])


def ifft_with_domain(domain: Radix2EvaluationDomain, evals: AuthenticatedDensePoly) -> AuthenticatedDensePoly:
    return fft_with_domain(domain, evals, inverse=True)
