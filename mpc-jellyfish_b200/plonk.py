"""Host-side mirror of `PlonkKzgSnark::{preprocess, prove}` for one TurboPlonk instance
(plonk/src/proof_system/snark.rs:529-611, 613-640 -> 201-469) and of the `PlonkTranscript`
implementations (plonk/src/transcript/{solidity,standard}.rs).  Everything here is a thin ctypes
wrapper: the five prover rounds, the polynomial algebra, the MSMs / NTTs and the Fiat-Shamir
transcript all run inside libjf_b200.so (csrc/plonk.cu, csrc/transcript.hpp).

The caller hands over what `relation::PlonkCircuit` produces after `finalize_for_arithmetization`:
selector columns, the extended permutation, the coset representatives k, the wire -> variable map and
the io gate ids; per proof, the witness and the 17 masking scalars.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from .context import CommitKey, Context
from .errors import InvalidParameters

TRANSCRIPT_KINDS = {"solidity": 0, "standard": 1}
NUM_WIRE_TYPES = 5
NUM_SELECTORS = 13
NUM_BLINDERS = 17
# UltraPlonk: the range / lookup wire, the q_lookup selector, 3 + 3 + 3 more masking scalars (h1, h2, lookup product) and one
# more split-quotient randomizer
ULTRA_WIRE_TYPES = 6
ULTRA_SELECTORS = 14
ULTRA_BLINDERS = 29
PLOOKUP_EVAL_FIELDS = ["range_table_eval", "key_table_eval", "table_dom_sep_eval", "q_dom_sep_eval", "h_1_eval", "q_lookup_eval",
                       "prod_next_eval", "range_table_next_eval", "key_table_next_eval", "table_dom_sep_next_eval", "h_1_next_eval",
                       "h_2_next_eval", "q_lookup_next_eval", "w_3_next_eval", "w_4_next_eval"]  # structs.rs:496-541


def keccak256(data: bytes) -> bytes:
    """sha3 crate `Keccak256` (host code of the library; no GPU needed)."""
    out = ctypes.create_string_buffer(32)
    _ffi.lib().jf_keccak256(data, len(data), out)
    return out.raw


class Transcript:
    """`SolidityTranscript` / `StandardTranscript` (host code of the library; no GPU needed)."""

    def __init__(self, kind: str, label: bytes = b"PlonkProof"):
        self._lib = _ffi.lib()
        self._h = self._lib.jf_transcript_new(TRANSCRIPT_KINDS[kind], label)
        if not self._h:
            raise InvalidParameters("unknown transcript kind %r" % kind)

    def append_message(self, label: bytes, msg: bytes):
        self._lib.jf_transcript_append(self._h, label, msg, len(msg))

    def get_and_append_challenge(self, field: str, label: bytes) -> np.ndarray:
        """-> 4 Montgomery limbs of the challenge."""
        out = np.zeros(4, dtype=np.uint64)
        rc = self._lib.jf_transcript_challenge(self._h, _ffi.FIELDS[field], label, out.ctypes.data_as(_ffi.c_u64p))
        if rc != _ffi.JF_OK:
            raise InvalidParameters("transcript challenge failed (%d)" % rc)
        return out

    def __del__(self):
        try:
            if self._h:
                self._lib.jf_transcript_free(self._h)
                self._h = None
        except Exception:
            pass


@dataclass
class Proof:
    """`Proof<E>` (structs.rs:62-84).  Points: (count, 2L) Montgomery x || y + infinity flags;
    evaluations: (count, 4) Montgomery limbs."""
    curve: str
    wires_poly_comms: np.ndarray
    wires_inf: List[bool]
    prod_perm_poly_comm: np.ndarray
    prod_perm_inf: bool
    split_quot_poly_comms: np.ndarray
    split_inf: List[bool]
    opening_proof: np.ndarray
    opening_inf: bool
    shifted_opening_proof: np.ndarray
    shifted_opening_inf: bool
    wires_evals: np.ndarray
    wire_sigma_evals: np.ndarray
    perm_next_eval: np.ndarray
    challenges: np.ndarray  # beta, gamma, alpha, zeta, v  (UltraPlonk: tau first)
    _raw: object = None
    # `plookup_proof: Some(PlookupProof)` of an UltraPlonk proof (structs.rs:255-265): None for TurboPlonk
    h_poly_comms: Optional[np.ndarray] = None
    h_inf: Optional[List[bool]] = None
    prod_lookup_poly_comm: Optional[np.ndarray] = None
    prod_lookup_inf: bool = False
    plookup_evals: Optional[np.ndarray] = None  # (15, 4), `PlookupEvaluations` in declaration order

    def serialize_compressed(self) -> bytes:
        buf = ctypes.create_string_buffer(4096)
        fn = _ffi.lib().jf_plonk_proof_serialize if self.h_poly_comms is None else _ffi.lib().jf_ultraplonk_proof_serialize
        n = fn(ctypes.byref(self._raw), buf, len(buf))
        if n < 0:
            raise InvalidParameters("proof serialization failed (%d)" % n)
        return buf.raw[:n]


@dataclass
class GroupLayout:
    """`GroupLayout` (relation/src/proof_linking/mod.rs:17-53): `size` proof-linking gates on the 2^alignment-th roots of unity,
    starting at `offset`."""
    alignment: int
    offset: int
    size: int


@dataclass
class LinkingHint:
    """`LinkingHint<E>` (structs.rs:88-97): the first wire polynomial after masking ((n + 2, 4) Montgomery limbs) and its commitment."""
    linking_wire_poly: np.ndarray
    linking_wire_comm: np.ndarray
    linking_wire_inf: bool = False


@dataclass
class LinkingProof:
    """`LinkingProof<E>` (proof_linking.rs:32-39)."""
    curve: str
    quotient_commitment: np.ndarray
    quotient_inf: bool
    opening_proof: np.ndarray
    opening_inf: bool
    eta: np.ndarray   # the opening challenge, Montgomery limbs (diagnostics)
    path: int         # 0: exact division on a coset, 2: the same after taking the remainder off, 1: successive linear divisions
    _raw: object = None

    def serialize_compressed(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        n = _ffi.lib().jf_link_proof_serialize(ctypes.byref(self._raw), buf, len(buf))
        if n < 0:
            raise InvalidParameters("linking proof serialization failed (%d)" % n)
        return buf.raw[:n]


def _link_proof_from_raw(curve: str, raw) -> LinkingProof:
    L = _ffi.CURVE_FQ_LIMBS[curve]
    return LinkingProof(curve, np.array(raw.quotient_commitment[:2 * L], dtype=np.uint64), bool(raw.quotient_inf),
                        np.array(raw.opening_proof[:2 * L], dtype=np.uint64), bool(raw.opening_inf),
                        np.array(raw.eta, dtype=np.uint64), int(raw.path), raw)


class ProvingKey:
    """`ProvingKey<E>` resident on the GPU (selector / sigma polynomials, commit key, vk commitments)."""

    def __init__(self, ctx: Context, key: CommitKey, handle, n: int, num_vars: int, num_inputs: int, k: np.ndarray, ultra: bool = False):
        self.ctx, self.key, self._h, self.n, self.num_vars, self.num_inputs, self.k = ctx, key, handle, n, num_vars, num_inputs, k
        self.ultra = ultra
        L = _ffi.CURVE_FQ_LIMBS[key.curve]
        ns, nw = (ULTRA_SELECTORS, ULTRA_WIRE_TYPES) if ultra else (NUM_SELECTORS, NUM_WIRE_TYPES)
        total = ns + nw + (4 if ultra else 0)
        xy = np.zeros((total, 2 * L), dtype=np.uint64)
        inf = (ctypes.c_int * total)()
        ctx._check(ctx._lib.jf_plonk_vk_commitments(ctx._h, handle, xy.ctypes.data_as(_ffi.c_u64p), inf))
        self.selector_comms, self.sigma_comms = xy[:ns], xy[ns:ns + nw]
        self.selector_inf = [bool(v) for v in inf[:ns]]
        self.sigma_inf = [bool(v) for v in inf[ns:ns + nw]]
        # PlookupVerifyingKey: range table, key table, table dom sep, q dom sep (snark.rs:573-594)
        self.plookup_comms = xy[ns + nw:] if ultra else None
        self.plookup_inf = [bool(v) for v in inf[ns + nw:]] if ultra else None

    def shard_commits(self, comm, key_slice: Optional[CommitKey], slice_start: int = 0, shard_round3: bool = True):
        """ONE proof on several GPUs (`jf_plonk_pk_shard_commits`): from now on every rank runs the same prove call and commits
        only coefficients [slice_start, slice_start + len(key_slice)); with shard_round3 the sub-cosets of round 3 are dealt out over
        the ranks as well.  comm = None returns to one-GPU operation.  `comm`: `mpc_jellyfish_b200.sharded.Comm`."""
        self._comm, self._key_slice = comm, key_slice   # keep them alive
        self.ctx._check(self.ctx._lib.jf_plonk_pk_shard_commits(self.ctx._h, self._h, comm._h if comm is not None else None,
                                                                key_slice._h if key_slice is not None else None, slice_start,
                                                                int(shard_round3)))

    def free(self):
        if self._h is not None and self.ctx._h:
            self.ctx._lib.jf_plonk_pk_free(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _proof_from_raw(curve: str, raw, ultra: bool) -> Proof:
    L = _ffi.CURVE_FQ_LIMBS[curve]
    nw = ULTRA_WIRE_TYPES if ultra else NUM_WIRE_TYPES

    def pts(arr, count):
        return np.array(arr[: count * 2 * L], dtype=np.uint64).reshape(count, 2 * L)

    pr = Proof(
        curve=curve,
        wires_poly_comms=pts(raw.wires_poly_comms, nw), wires_inf=[bool(v) for v in raw.wires_inf],
        prod_perm_poly_comm=pts(raw.prod_perm_poly_comm, 1)[0], prod_perm_inf=bool(raw.prod_perm_inf),
        split_quot_poly_comms=pts(raw.split_quot_poly_comms, nw), split_inf=[bool(v) for v in raw.split_inf],
        opening_proof=pts(raw.opening_proof, 1)[0], opening_inf=bool(raw.opening_inf),
        shifted_opening_proof=pts(raw.shifted_opening_proof, 1)[0], shifted_opening_inf=bool(raw.shifted_opening_inf),
        wires_evals=np.array(raw.wires_evals, dtype=np.uint64).reshape(nw, 4),
        wire_sigma_evals=np.array(raw.wire_sigma_evals, dtype=np.uint64).reshape(nw - 1, 4),
        perm_next_eval=np.array(raw.perm_next_eval, dtype=np.uint64),
        challenges=np.array(raw.challenges, dtype=np.uint64).reshape(-1, 4), _raw=raw)
    if ultra:
        pr.h_poly_comms = pts(raw.h_poly_comms, 2)
        pr.h_inf = [bool(v) for v in raw.h_inf]
        pr.prod_lookup_poly_comm = pts(raw.prod_lookup_poly_comm, 1)[0]
        pr.prod_lookup_inf = bool(raw.prod_lookup_inf)
        pr.plookup_evals = np.array(raw.plookup_evals, dtype=np.uint64).reshape(15, 4)
    return pr


@dataclass
class BatchProof:
    """`BatchProof<E>` (structs.rs:271-292): per-instance records plus the shared split-quotient commitments and opening
    proofs (present in every record)."""
    proofs: List[Proof]
    _raw: object = None
    ultra: bool = False

    def __len__(self):
        return len(self.proofs)

    def serialize_compressed(self) -> bytes:
        buf = ctypes.create_string_buffer(4096 * (len(self.proofs) + 1))
        fn = _ffi.lib().jf_ultraplonk_batch_proof_serialize if self.ultra else _ffi.lib().jf_plonk_batch_proof_serialize
        n = fn(self._raw, len(self.proofs), buf, len(buf))
        if n < 0:
            raise InvalidParameters("batch proof serialization failed (%d)" % n)
        return buf.raw[:n]


class PlonkKzgSnark:
    @staticmethod
    def batch_prove(pks: Sequence["ProvingKey"], witnesses: Sequence[np.ndarray], blinders: np.ndarray, transcript: str = "solidity",
                    extra_transcript_init_msg: Optional[bytes] = None) -> BatchProof:
        """`PlonkKzgSnark::batch_prove` (snark.rs:201-469).  blinders: (count * 13 + 4, 4) for TurboPlonk, (count * 24 + 5, 4) for
        UltraPlonk, in the reference's prng order (all wire masks, [all h masks,] all z masks, [all lookup-product masks,] split)."""
        if not pks or len(pks) != len(witnesses):
            raise InvalidParameters("the number of circuits != the number of proving keys (or zero)")
        ctx, ultra, count = pks[0].ctx, pks[0].ultra, len(pks)
        nw = ULTRA_WIRE_TYPES if ultra else NUM_WIRE_TYPES
        per = 2 * nw + 3 + (9 if ultra else 0)
        b = np.ascontiguousarray(blinders, dtype=np.uint64)
        if b.shape != (count * per + nw - 1, 4):
            raise InvalidParameters("blinders must be (%d, 4)" % (count * per + nw - 1))
        ws = [np.ascontiguousarray(w, dtype=np.uint64) for w in witnesses]
        for pk, w in zip(pks, ws):
            if pk.ultra != ultra:
                raise InvalidParameters("inconsistent plonk circuit types")
            if w.shape != (pk.num_vars, 4):
                raise InvalidParameters("witness must be (num_vars, 4)")
        handles = (ctypes.c_void_p * count)(*[pk._h for pk in pks])
        wptrs = (_ffi.c_u64p * count)(*[w.ctypes.data_as(_ffi.c_u64p) for w in ws])
        raw = ((_ffi.UltraPlonkProofStruct if ultra else _ffi.PlonkProofStruct) * count)()
        fn = ctx._lib.jf_ultraplonk_batch_prove if ultra else ctx._lib.jf_plonk_batch_prove
        ctx._check(fn(ctx._h, handles, count, wptrs, b.ctypes.data_as(_ffi.c_u64p), TRANSCRIPT_KINDS[transcript],
                      extra_transcript_init_msg, len(extra_transcript_init_msg) if extra_transcript_init_msg else 0, raw))
        return BatchProof([_proof_from_raw(pks[0].key.curve, raw[i], ultra) for i in range(count)], raw, ultra)

    @staticmethod
    def preprocess(ctx: Context, key: CommitKey, selector_evals: np.ndarray, sigma_evals: np.ndarray, k: np.ndarray,
                   wire_variables: np.ndarray, num_vars: int, pub_input_gate_ids: Sequence[int] = (),
                   cache_coset_evals: bool = False, skip_zero_selectors: bool = False,
                   full_quotient_coset: bool = False, lagrange_wire_commitments: bool = False) -> ProvingKey:
        """selector_evals (13, n, 4), sigma_evals (5, n, 4) = the extended permutation, k (5, 4):
        Montgomery limbs; wire_variables (5, n) uint32."""
        sel = np.ascontiguousarray(selector_evals, dtype=np.uint64)
        sig = np.ascontiguousarray(sigma_evals, dtype=np.uint64)
        kk = np.ascontiguousarray(k, dtype=np.uint64)
        wv = np.ascontiguousarray(wire_variables, dtype=np.uint32)
        if sel.ndim != 3 or sel.shape[0] != NUM_SELECTORS or sel.shape[2] != 4:
            raise InvalidParameters("selector_evals must be (13, n, 4)")
        n = sel.shape[1]
        if n & (n - 1) or sig.shape != (NUM_WIRE_TYPES, n, 4) or kk.shape != (NUM_WIRE_TYPES, 4) or wv.shape != (NUM_WIRE_TYPES, n):
            raise InvalidParameters("inconsistent proving-key shapes")
        gids = np.ascontiguousarray(list(pub_input_gate_ids), dtype=np.uint32)
        h = ctypes.c_void_p()
        ctx._check(ctx._lib.jf_plonk_preprocess(
            ctx._h, key._h, n.bit_length() - 1, sel.ctypes.data_as(_ffi.c_u64p), sig.ctypes.data_as(_ffi.c_u64p),
            kk.ctypes.data_as(_ffi.c_u64p), wv.ctypes.data_as(_ffi.c_u32p), num_vars,
            gids.ctypes.data_as(_ffi.c_u32p) if len(gids) else None, len(gids), int(cache_coset_evals) | (2 if skip_zero_selectors else 0) | (4 if full_quotient_coset else 0)
            | (8 if lagrange_wire_commitments else 0), ctypes.byref(h)))
        return ProvingKey(ctx, key, h, n, num_vars, len(gids), kk)

    @staticmethod
    def prove(pk: ProvingKey, witness: np.ndarray, blinders: np.ndarray, transcript: str = "solidity",
              extra_transcript_init_msg: Optional[bytes] = None) -> Proof:
        """witness (num_vars, 4), blinders (17, 4): Montgomery limbs."""
        ctx = pk.ctx
        w = np.ascontiguousarray(witness, dtype=np.uint64)
        b = np.ascontiguousarray(blinders, dtype=np.uint64)
        if w.shape != (pk.num_vars, 4) or b.shape != (NUM_BLINDERS, 4):
            raise InvalidParameters("witness must be (num_vars, 4) and blinders (17, 4)")
        raw = _ffi.PlonkProofStruct()
        ctx._check(ctx._lib.jf_plonk_prove(ctx._h, pk._h, w.ctypes.data_as(_ffi.c_u64p), b.ctypes.data_as(_ffi.c_u64p),
                                           TRANSCRIPT_KINDS[transcript], extra_transcript_init_msg,
                                           len(extra_transcript_init_msg) if extra_transcript_init_msg else 0,
                                           ctypes.byref(raw)))
        L = _ffi.CURVE_FQ_LIMBS[pk.key.curve]

        def pts(arr, count):
            return np.array(arr[: count * 2 * L], dtype=np.uint64).reshape(count, 2 * L)

        return Proof(
            curve=pk.key.curve,
            wires_poly_comms=pts(raw.wires_poly_comms, 5), wires_inf=[bool(v) for v in raw.wires_inf],
            prod_perm_poly_comm=pts(raw.prod_perm_poly_comm, 1)[0], prod_perm_inf=bool(raw.prod_perm_inf),
            split_quot_poly_comms=pts(raw.split_quot_poly_comms, 5), split_inf=[bool(v) for v in raw.split_inf],
            opening_proof=pts(raw.opening_proof, 1)[0], opening_inf=bool(raw.opening_inf),
            shifted_opening_proof=pts(raw.shifted_opening_proof, 1)[0], shifted_opening_inf=bool(raw.shifted_opening_inf),
            wires_evals=np.array(raw.wires_evals, dtype=np.uint64).reshape(5, 4),
            wire_sigma_evals=np.array(raw.wire_sigma_evals, dtype=np.uint64).reshape(4, 4),
            perm_next_eval=np.array(raw.perm_next_eval, dtype=np.uint64),
            challenges=np.array(raw.challenges, dtype=np.uint64).reshape(5, 4), _raw=raw)

    # ---- proof linking (plonk/src/proof_system/proof_linking.rs) --------------------------------------------------------------
    @staticmethod
    def prove_with_link_hint(pk: ProvingKey, witness: np.ndarray, blinders: np.ndarray, transcript: str = "solidity"):
        """`PlonkKzgSnark::prove_with_link_hint` (snark.rs:81-114) -> (Proof, LinkingHint); TurboPlonk only."""
        if pk.ultra:
            raise InvalidParameters("only TurboPlonk supports link groups")
        proof = PlonkKzgSnark.prove(pk, witness, blinders, transcript)
        ctx = pk.ctx
        poly = np.zeros((pk.n + 2, 4), dtype=np.uint64)
        got = ctypes.c_size_t(0)
        ctx._check(ctx._lib.jf_plonk_link_hint(ctx._h, pk._h, poly.ctypes.data_as(_ffi.c_u64p), pk.n + 2, ctypes.byref(got)))
        return proof, LinkingHint(poly[:got.value], proof.wires_poly_comms[0].copy(), proof.wires_inf[0])

    @staticmethod
    def link_proofs(ctx: Context, key: CommitKey, lhs: LinkingHint, rhs: LinkingHint, layout: GroupLayout,
                    transcript: str = "solidity", sequential_division: bool = False) -> LinkingProof:
        """`PlonkKzgSnark::link_proofs` (proof_linking.rs:79-112) from two hints in host memory."""
        a1 = np.ascontiguousarray(lhs.linking_wire_poly, dtype=np.uint64).reshape(-1, 4)
        a2 = np.ascontiguousarray(rhs.linking_wire_poly, dtype=np.uint64).reshape(-1, 4)
        c1 = np.ascontiguousarray(lhs.linking_wire_comm, dtype=np.uint64)
        c2 = np.ascontiguousarray(rhs.linking_wire_comm, dtype=np.uint64)
        raw = _ffi.LinkProofStruct()
        ctx._check(ctx._lib.jf_plonk_link_proofs(
            ctx._h, key._h, a1.ctypes.data_as(_ffi.c_u64p), len(a1), c1.ctypes.data_as(_ffi.c_u64p), int(lhs.linking_wire_inf),
            a2.ctypes.data_as(_ffi.c_u64p), len(a2), c2.ctypes.data_as(_ffi.c_u64p), int(rhs.linking_wire_inf),
            layout.alignment, layout.offset, layout.size, TRANSCRIPT_KINDS[transcript], int(sequential_division), ctypes.byref(raw)))
        return _link_proof_from_raw(key.curve, raw)

    @staticmethod
    def link_proofs_resident(lhs_pk: ProvingKey, lhs_proof: Proof, rhs_pk: ProvingKey, rhs_proof: Proof, layout: GroupLayout,
                             transcript: str = "solidity", sequential_division: bool = False) -> LinkingProof:
        """The same with both wire polynomials still in the workspaces of the two proving keys (no polynomial crosses PCIe):
        `lhs_proof` / `rhs_proof` must be the LAST proofs made with those keys."""
        ctx = lhs_pk.ctx
        raw = _ffi.LinkProofStruct()
        ctx._check(ctx._lib.jf_plonk_link_proofs_resident(
            ctx._h, lhs_pk._h, ctypes.byref(lhs_proof._raw), rhs_pk._h, ctypes.byref(rhs_proof._raw), layout.alignment, layout.offset,
            layout.size, TRANSCRIPT_KINDS[transcript], int(sequential_division), ctypes.byref(raw)))
        return _link_proof_from_raw(lhs_pk.key.curve, raw)

    # ---- UltraPlonk (`PlonkCircuit::new_ultra_plonk`; Plookup argument) ----------------------------------------------------
    @staticmethod
    def preprocess_ultra(ctx: Context, key: CommitKey, selector_evals: np.ndarray, sigma_evals: np.ndarray, k: np.ndarray,
                         wire_variables: np.ndarray, num_vars: int, pub_input_gate_ids: Sequence[int], range_bit_len: int,
                         table_key_evals: np.ndarray, table_dom_sep_evals: np.ndarray, q_dom_sep_evals: np.ndarray,
                         skip_zero_selectors: bool = False, full_quotient_coset: bool = False,
                         lagrange_wire_commitments: bool = False, cache_coset_evals: bool = False) -> ProvingKey:
        """selector_evals (14, n, 4) = `all_selectors()` with q_lookup last, sigma_evals (6, n, 4), k (6, 4), wire_variables (6, n)
        with the range wire last; table_key / table_dom_sep / q_dom_sep: (n, 4) per-gate columns (constraint_system.rs:873-888)."""
        sel = np.ascontiguousarray(selector_evals, dtype=np.uint64)
        sig = np.ascontiguousarray(sigma_evals, dtype=np.uint64)
        kk = np.ascontiguousarray(k, dtype=np.uint64)
        wv = np.ascontiguousarray(wire_variables, dtype=np.uint32)
        if sel.ndim != 3 or sel.shape[0] != ULTRA_SELECTORS or sel.shape[2] != 4:
            raise InvalidParameters("selector_evals must be (14, n, 4)")
        n = sel.shape[1]
        cols = [np.ascontiguousarray(c, dtype=np.uint64) for c in (table_key_evals, table_dom_sep_evals, q_dom_sep_evals)]
        if n & (n - 1) or sig.shape != (ULTRA_WIRE_TYPES, n, 4) or kk.shape != (ULTRA_WIRE_TYPES, 4) or wv.shape != (ULTRA_WIRE_TYPES, n) \
                or any(c.shape != (n, 4) for c in cols):
            raise InvalidParameters("inconsistent proving-key shapes")
        gids = np.ascontiguousarray(list(pub_input_gate_ids), dtype=np.uint32)
        h = ctypes.c_void_p()
        ctx._check(ctx._lib.jf_ultraplonk_preprocess(
            ctx._h, key._h, n.bit_length() - 1, sel.ctypes.data_as(_ffi.c_u64p), sig.ctypes.data_as(_ffi.c_u64p),
            kk.ctypes.data_as(_ffi.c_u64p), wv.ctypes.data_as(_ffi.c_u32p), num_vars,
            gids.ctypes.data_as(_ffi.c_u32p) if len(gids) else None, len(gids), range_bit_len,
            cols[0].ctypes.data_as(_ffi.c_u64p), cols[1].ctypes.data_as(_ffi.c_u64p), cols[2].ctypes.data_as(_ffi.c_u64p),
            int(cache_coset_evals) | (2 if skip_zero_selectors else 0) | (4 if full_quotient_coset else 0)
            | (8 if lagrange_wire_commitments else 0), ctypes.byref(h)))
        return ProvingKey(ctx, key, h, n, num_vars, len(gids), kk, ultra=True)

    @staticmethod
    def prove_ultra(pk: ProvingKey, witness: np.ndarray, blinders: np.ndarray, transcript: str = "solidity",
                    extra_transcript_init_msg: Optional[bytes] = None) -> Proof:
        """witness (num_vars, 4), blinders (29, 4): Montgomery limbs."""
        ctx = pk.ctx
        w = np.ascontiguousarray(witness, dtype=np.uint64)
        b = np.ascontiguousarray(blinders, dtype=np.uint64)
        if w.shape != (pk.num_vars, 4) or b.shape != (ULTRA_BLINDERS, 4):
            raise InvalidParameters("witness must be (num_vars, 4) and blinders (29, 4)")
        raw = _ffi.UltraPlonkProofStruct()
        ctx._check(ctx._lib.jf_ultraplonk_prove(ctx._h, pk._h, w.ctypes.data_as(_ffi.c_u64p), b.ctypes.data_as(_ffi.c_u64p),
                                                TRANSCRIPT_KINDS[transcript], extra_transcript_init_msg,
                                                len(extra_transcript_init_msg) if extra_transcript_init_msg else 0,
                                                ctypes.byref(raw)))
        L = _ffi.CURVE_FQ_LIMBS[pk.key.curve]

        def pts(arr, count):
            return np.array(arr[: count * 2 * L], dtype=np.uint64).reshape(count, 2 * L)

        return Proof(
            curve=pk.key.curve,
            wires_poly_comms=pts(raw.wires_poly_comms, 6), wires_inf=[bool(v) for v in raw.wires_inf],
            prod_perm_poly_comm=pts(raw.prod_perm_poly_comm, 1)[0], prod_perm_inf=bool(raw.prod_perm_inf),
            split_quot_poly_comms=pts(raw.split_quot_poly_comms, 6), split_inf=[bool(v) for v in raw.split_inf],
            opening_proof=pts(raw.opening_proof, 1)[0], opening_inf=bool(raw.opening_inf),
            shifted_opening_proof=pts(raw.shifted_opening_proof, 1)[0], shifted_opening_inf=bool(raw.shifted_opening_inf),
            wires_evals=np.array(raw.wires_evals, dtype=np.uint64).reshape(6, 4),
            wire_sigma_evals=np.array(raw.wire_sigma_evals, dtype=np.uint64).reshape(5, 4),
            perm_next_eval=np.array(raw.perm_next_eval, dtype=np.uint64),
            challenges=np.array(raw.challenges, dtype=np.uint64).reshape(6, 4), _raw=raw,
            h_poly_comms=pts(raw.h_poly_comms, 2), h_inf=[bool(v) for v in raw.h_inf],
            prod_lookup_poly_comm=pts(raw.prod_lookup_poly_comm, 1)[0], prod_lookup_inf=bool(raw.prod_lookup_inf),
            plookup_evals=np.array(raw.plookup_evals, dtype=np.uint64).reshape(15, 4))
