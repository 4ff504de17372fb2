"""Host-side mirror of the collaborative (MPC) prover's share-wise use of the two kernels
(SURVEY §8 rows a13 / a14):

  MultiproverKZG::{commit, batch_commit, open}  plonk/src/multiprover/primitives/multiprover_kzg.rs:128-197
  fft_with_domain / ifft_with_domain            plonk/src/multiprover/proof_system/prover.rs:373-388,418
  AuthenticatedScalarResult::ifft               plonk/src/multiprover/proof_system/constraint_system.rs:930,955,978

An authenticated value is an additive share plus an additive share of its MAC (ark-mpc
`AuthenticatedScalarResult { share, mac, public_modifier }`; the reference builds shares as `ScalarShare::new(share, mac)`,
multiprover/proof_system/prover.rs:985-1013).  ark-mpc is an unpinned git dependency (Cargo.toml:16): revisions that carry
the third component -- the public modifier, the running sum of public constants folded into the value, which the MAC check
needs -- push it through `msm_authenticated` / `fft_with_domain` as a third vector.  MSM and NTT are linear, so each party
applies them to every component vector locally -- here as ONE batch of two or three on the GPU -- and the results are
valid shares of the plain result.  Opening shares (the network exchange of ark-mpc) is not on this path; the tests add
the parties' results, which is the arithmetic the reference's own tests use (`open_authenticated`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .context import Context
from .domain import Radix2EvaluationDomain
from .errors import InvalidParameters
from .pcs import UnivariateProverParam


@dataclass
class AuthenticatedDensePoly:
    """`AuthenticatedDensePoly`: coefficient shares, their MAC shares and (optionally: the three-component form of
    ark-mpc's authenticated scalars) their public modifiers, (n, 4) Montgomery limbs each."""
    share: np.ndarray
    mac: np.ndarray
    public_modifier: Optional[np.ndarray] = None

    def __post_init__(self):
        self.share = np.ascontiguousarray(self.share, dtype=np.uint64).reshape(-1, 4)
        self.mac = np.ascontiguousarray(self.mac, dtype=np.uint64).reshape(-1, 4)
        if self.share.shape != self.mac.shape:
            raise InvalidParameters("share and MAC vectors differ in length")
        if self.public_modifier is not None:
            self.public_modifier = np.ascontiguousarray(self.public_modifier, dtype=np.uint64).reshape(-1, 4)
            if self.public_modifier.shape != self.share.shape:
                raise InvalidParameters("share and public-modifier vectors differ in length")

    def degree(self) -> int:
        return max(len(self.share) - 1, 0)

    def components(self) -> List[np.ndarray]:
        """the vectors every linear operation is applied to, in ark-mpc's field order"""
        return [self.share, self.mac] + ([self.public_modifier] if self.public_modifier is not None else [])


@dataclass
class AuthenticatedPointShare:
    """One party's share of a commitment / opening proof: (share point, MAC point[, public-modifier point]), affine x || y."""
    share: np.ndarray
    share_inf: bool
    mac: np.ndarray
    mac_inf: bool
    public_modifier: Optional[np.ndarray] = None
    public_modifier_inf: bool = True


def _point_shares(polys: Sequence[AuthenticatedDensePoly], out: np.ndarray, inf: Sequence[bool]) -> List[AuthenticatedPointShare]:
    res, i = [], 0
    for p in polys:
        k = len(p.components())
        res.append(AuthenticatedPointShare(out[i], inf[i], out[i + 1], inf[i + 1], out[i + 2] if k == 3 else None,
                                           inf[i + 2] if k == 3 else True))
        i += k
    return res


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
            vecs += p.components()
        # one call for every component of every polynomial (the reference loops over `commit`, and each `commit` rebuilds the
        # key as `Vec<CurvePoint>`, multiprover_kzg.rs:232-234; here the key stays resident)
        out, inf = prover_params.ctx.msm_batch(prover_params.key, vecs, None, montgomery=True)
        return _point_shares(polys, out, inf)

    @staticmethod
    def open(prover_params: UnivariateProverParam, poly: AuthenticatedDensePoly, point: np.ndarray
             ) -> Tuple[AuthenticatedPointShare, Tuple[np.ndarray, np.ndarray]]:
        """`point` is public (4 Montgomery limbs).  -> (proof share, (evaluation share, evaluation MAC share))."""
        if poly.degree() - 1 > len(prover_params.key):  # multiprover_kzg.rs:176-181
            raise InvalidParameters("Polynomial degree exceeds supported degree")
        proofs, evals = MultiproverKZG.batch_open(prover_params, [poly], [point])
        return proofs[0], evals[0]

    @staticmethod
    def batch_open(prover_params: UnivariateProverParam, polys: Sequence[AuthenticatedDensePoly], points: Sequence[np.ndarray]
                   ) -> Tuple[List[AuthenticatedPointShare], List[Tuple[np.ndarray, ...]]]:
        """`MultiproverKZG::batch_open` (multiprover_kzg.rs:199-229): every (polynomial, public point) pair, all components,
        in one `jf_kzg_open` call.  -> (proof shares, per polynomial the evaluation of each component)."""
        if len(polys) != len(points):
            raise InvalidParameters("poly length %d is different from points length %d" % (len(polys), len(points)))
        vecs, zs = [], []
        for poly, point in zip(polys, points):
            if poly.degree() - 1 > len(prover_params.key):  # multiprover_kzg.rs:176-181
                raise InvalidParameters("Polynomial degree exceeds supported degree")
            z = np.ascontiguousarray(point, dtype=np.uint64).reshape(1, 4)
            comps = poly.components()
            vecs += comps
            zs += [z] * len(comps)
        xy, inf, ev = prover_params.ctx.kzg_open(prover_params.key, vecs, np.concatenate(zs, axis=0))
        proofs = _point_shares(polys, xy, inf)
        evals, i = [], 0
        for poly in polys:
            k = len(poly.components())
            evals.append(tuple(ev[i + j] for j in range(k)))
            i += k
        return proofs, evals


def fft_with_domain(domain: Radix2EvaluationDomain, poly: AuthenticatedDensePoly, inverse: bool = False) -> AuthenticatedDensePoly:
    """Share-wise (coset) NTT / iNTT of an authenticated vector: one batch of two (three) on the GPU."""
    n = domain.size
    if len(poly.share) > n:
        raise InvalidParameters("input of length %d exceeds the domain size %d" % (len(poly.share), n))
    comps = poly.components()
    buf = np.zeros((len(comps), n, 4), dtype=np.uint64)
    for i, c in enumerate(comps):
        buf[i, : len(c)] = c
    out = domain.batch_fft(buf, inverse=inverse, in_len=len(poly.share))
    return AuthenticatedDensePoly(out[0], out[1], out[2] if len(comps) == 3 else None)


def ifft_with_domain(domain: Radix2EvaluationDomain, evals: AuthenticatedDensePoly) -> AuthenticatedDensePoly:
    return fft_with_domain(domain, evals, inverse=True)


# ---- collaborative proof linking (plonk/src/multiprover/proof_system/proof_linking.rs:96-246) --------------------------------
@dataclass
class MpcLinkingHint:
    """`MpcLinkingHint` (multiprover/proof_system/structs.rs:206-230): this party's share of the first wire polynomial and of its
    commitment."""
    linking_wire_poly: AuthenticatedDensePoly
    linking_wire_comm: AuthenticatedPointShare


class MultiproverLinking:
    """The party-local arithmetic of `MultiproverPlonkKzgSnark::link_proofs`.  The protocol opens the quotient commitment before the
    challenge eta exists, so it comes in two steps around that opening (ark-mpc's network; here the tests add the parties' points):

        q, q_comm = MultiproverLinking.quotient(params, lhs_hint, rhs_hint, layout)          # every party
        eta = MultiproverLinking.challenge(ctx, curve, a1_comm, a2_comm, opened_q_comm)      # public points -> public scalar
        proof = MultiproverLinking.identity_opening(params, lhs_hint, rhs_hint, q, eta, layout)

    The opened (quotient commitment, opening proof) equal the single prover's `LinkingProof` under `SolidityTranscript`
    (`MpcTranscript` wraps it: multiprover/primitives/mpc_transcript.rs:27-52)."""

    @staticmethod
    def _sub(ctx: Context, field: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        n = max(len(a), len(b))
        pa, pb = np.zeros((n, 4), dtype=np.uint64), np.zeros((n, 4), dtype=np.uint64)
        pa[: len(a)], pb[: len(b)] = a, b
        return ctx.field_op(field, "sub", pa, pb) if n else pa

    @staticmethod
    def quotient(prover_params: UnivariateProverParam, lhs: MpcLinkingHint, rhs: MpcLinkingHint, layout
                 ) -> Tuple[AuthenticatedDensePoly, AuthenticatedPointShare]:
        """compute_linking_quotient (:127-138) share-wise -- division by the public Z_D is linear -- and `MultiproverKZG::commit`."""
        import ctypes
        from . import _ffi
        ctx = prover_params.ctx
        field = _ffi.CURVE_FR[prover_params.key.curve]
        c1, c2 = lhs.linking_wire_poly.components(), rhs.linking_wire_poly.components()
        if len(c1) != len(c2):
            raise InvalidParameters("the two hints carry different share forms")
        diffs = [MultiproverLinking._sub(ctx, field, a, b) for a, b in zip(c1, c2)]
        k = len(diffs)
        outs = [np.zeros((max(len(d) - layout.size, 0), 4), dtype=np.uint64) for d in diffs]
        ptrs = (_ffi.c_u64p * k)(*[d.ctypes.data_as(_ffi.c_u64p) for d in diffs])
        optrs = (_ffi.c_u64p * k)(*[o.ctypes.data_as(_ffi.c_u64p) for o in outs])
        lens = (ctypes.c_size_t * k)(*[len(d) for d in diffs])
        ctx._check(ctx._lib.jf_poly_div_link_domain(ctx._h, _ffi.FIELDS[field], ptrs, lens, k, layout.alignment, layout.offset,
                                                    layout.size, 0, optrs))
        q = AuthenticatedDensePoly(outs[0], outs[1], outs[2] if k == 3 else None)
        return q, MultiproverKZG.commit(prover_params, q)

    @staticmethod
    def challenge(curve: str, a1_comm, a2_comm, quotient_comm) -> np.ndarray:
        """compute_quotient_challenge (:176-195) on the OPENED commitments: (xy, is_infinity) each -> eta, 4 Montgomery limbs."""
        from .plonk import Transcript
        from . import _ffi
        tr = Transcript("solidity", b"MpcPlonkLinkingProof")
        for label, (xy, inf) in ((b"linking_wire_comms", a1_comm), (b"linking_wire_comms", a2_comm), (b"quotient_comm", quotient_comm)):
            tr.append_message(label, g1_serialize_compressed(curve, xy, inf))
        return tr.get_and_append_challenge(_ffi.CURVE_FR[curve], b"eta")

    @staticmethod
    def identity_opening(prover_params: UnivariateProverParam, lhs: MpcLinkingHint, rhs: MpcLinkingHint, quotient: AuthenticatedDensePoly,
                         eta: np.ndarray, vanishing_eval: np.ndarray) -> AuthenticatedPointShare:
        """compute_identity_opening (:203-220): a1 - a2 - q Z_D(eta), opened at the public eta.  `vanishing_eval` = Z_D(eta)
        (4 Montgomery limbs; public, O(size) host products: compute_vanishing_poly_eval, :155-170)."""
        from . import _ffi
        ctx = prover_params.ctx
        field = _ffi.CURVE_FR[prover_params.key.curve]
        comps = []
        for a, b, q in zip(lhs.linking_wire_poly.components(), rhs.linking_wire_poly.components(), quotient.components()):
            d = MultiproverLinking._sub(ctx, field, a, b)
            if len(q):
                zq = ctx.field_op(field, "mul", q, np.broadcast_to(np.asarray(vanishing_eval, dtype=np.uint64), q.shape).copy())
                d = MultiproverLinking._sub(ctx, field, d, zq)
            comps.append(d)
        ident = AuthenticatedDensePoly(comps[0], comps[1], comps[2] if len(comps) == 3 else None)
        proof, _ = MultiproverKZG.open(prover_params, ident, eta)
        return proof


def g1_serialize_compressed(curve: str, xy: np.ndarray, inf: bool) -> bytes:
    """`to_bytes!` of a commitment (the library's host serialiser: ark-ec's form for BN254, the ZCash form for BLS12-381)."""
    import ctypes
    from . import _ffi
    raw = _ffi.LinkProofStruct()
    raw.curve = _ffi.CURVES[curve]
    L = _ffi.CURVE_FQ_LIMBS[curve]
    for i in range(2 * L):
        raw.quotient_commitment[i] = int(xy[i]) if not inf else 0
    raw.quotient_inf = int(bool(inf))
    raw.opening_inf = 1
    buf = ctypes.create_string_buffer(128)
    n = _ffi.lib().jf_link_proof_serialize(ctypes.byref(raw), buf, len(buf))
    if n < 0:
        raise InvalidParameters("point serialization failed (%d)" % n)
    return buf.raw[: n // 2]
