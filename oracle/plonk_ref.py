"""CPU restatement of the TurboPlonk prover / verifier of mpc-jellyfish (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s checker legs may import this file;
the product (``mpc-jellyfish_b200``) never does.  PARITY UNPINNED by the reference for proofs
(no golden proof exists and the Rust crates cannot be built here); what IS pinned:

* Keccak-256 against the reference's own known-answer test
  (``plonk/src/transcript/solidity.rs:80-96``, ``test_solidity_keccak``);
* Merlin/STROBE-128 against the merlin crate's published ``equivalence_simple`` vector;
* ChaCha20 against RFC 7539 / the all-zero-key keystream;
* the proof itself through a restatement of the reference verifier (``verifier.rs``) whose final
  pairing check e(A,[beta]_2) = e(B,[1]_2) collapses to the G1 identity beta*A == B because the
  test SRS has a known beta (``primitives/src/pcs/univariate_kzg/srs.rs:118-153``).

What follows which reference code (TurboPlonk, one instance, 5 wire types, no Plookup):
  PlonkCircuit            relation/src/constraint_system.rs:193-225,464-501,630-666,743-778,913-1003,1150-1259
                          relation/src/proof_linking/linkable_circuit.rs:136-230,294-314 (no link groups)
                          relation/src/gates/arithmetic.rs, relation/src/traits.rs:140-262,640-667
  coset representatives   relation/src/constants.rs:30-79 (ChaCha20 rng, zero seed, ark-ff Fp::rand)
  preprocess              plonk/src/proof_system/snark.rs:529-611
  prove                   plonk/src/proof_system/snark.rs:201-469 + prover.rs (all rounds)
  transcripts             plonk/src/transcript/{mod,solidity,standard}.rs
  verify                  plonk/src/proof_system/verifier.rs:57-254,317-805
All field values in this file are canonical Python ints.
"""
from __future__ import annotations

import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np

import pyref
from pyref import BN254, Curve, Field, Radix2Domain

GATE_WIDTH = 4
NUM_WIRE_TYPES = 5
N_SELECTORS = 13  # q_lc[4], q_mul[2], q_hash[4], q_o, q_c, q_ecc

# ======================================================================================
# Keccak-f[1600], Keccak-256 (sha3 crate `Keccak256`: original padding 0x01 .. 0x80)
# ======================================================================================
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def keccak_f1600(state: bytearray) -> None:
    """In-place permutation of a 200-byte state (lanes little-endian, A[x][y] at 8*(x+5y))."""
    A = [[int.from_bytes(state[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little") for y in range(5)] for x in range(5)]
    for rnd in range(24):
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= _RC[rnd]
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = A[x][y].to_bytes(8, "little")


def keccak256(data: bytes) -> bytes:
    rate = 136
    st = bytearray(200)
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    for off in range(0, len(msg), rate):
        for i in range(rate):
            st[i] ^= msg[off + i]
        keccak_f1600(st)
    return bytes(st[:32])


# ======================================================================================
# STROBE-128 / Merlin (merlin crate 3.x `Transcript`), as used by StandardTranscript
# ======================================================================================
class _Strobe128:
    R = 166
    I, A, C, T, M, K = 1, 2, 4, 8, 16, 32

    def __init__(self, protocol_label: bytes):
        st = bytearray(200)
        st[0:6] = bytes([1, self.R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.state, self.pos, self.pos_begin, self.cur_flags = st, 0, 0, 0
        self.meta_ad(protocol_label, False)

    def _run_f(self):
        self.state[self.pos] ^= self.pos_begin
        self.state[self.pos + 1] ^= 0x04
        self.state[self.R + 1] ^= 0x80
        keccak_f1600(self.state)
        self.pos, self.pos_begin = 0, 0

    def _absorb(self, data: bytes):
        for b in data:
            self.state[self.pos] ^= b
            self.pos += 1
            if self.pos == self.R:
                self._run_f()

    def _squeeze(self, n: int) -> bytes:
        out = bytearray()
        for _ in range(n):
            out.append(self.state[self.pos])
            self.state[self.pos] = 0
            self.pos += 1
            if self.pos == self.R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags: int, more: bool):
        if more:
            assert self.cur_flags == flags
            return
        assert flags & self.T == 0
        old_begin = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old_begin, flags]))
        if flags & (self.C | self.K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data: bytes, more: bool):
        self._begin_op(self.M | self.A, more)
        self._absorb(data)

    def ad(self, data: bytes, more: bool):
        self._begin_op(self.A, more)
        self._absorb(data)

    def prf(self, n: int, more: bool) -> bytes:
        self._begin_op(self.I | self.A | self.C, more)
        return self._squeeze(n)


class MerlinTranscript:
    def __init__(self, label: bytes):
        self.strobe = _Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label: bytes, message: bytes):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", len(message)), True)
        self.strobe.ad(message, False)

    def challenge_bytes(self, label: bytes, n: int) -> bytes:
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", n), True)
        return self.strobe.prf(n, False)


# ======================================================================================
# ChaCha20 rng (rand_chacha 0.3 `ChaChaRng` = ChaCha20Rng) + ark-ff 0.4 `Fp::rand`
# ======================================================================================
def chacha20_block(key_words: Sequence[int], counter: int, stream: int = 0) -> List[int]:
    """16 output words; state = consts | key | 64-bit counter | 64-bit stream id."""
    M = 0xFFFFFFFF
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + \
         [counter & M, (counter >> 32) & M, stream & M, (stream >> 32) & M]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & M; x[d] ^= x[a]; x[d] = ((x[d] << 16) | (x[d] >> 16)) & M
        x[c] = (x[c] + x[d]) & M; x[b] ^= x[c]; x[b] = ((x[b] << 12) | (x[b] >> 20)) & M
        x[a] = (x[a] + x[b]) & M; x[d] ^= x[a]; x[d] = ((x[d] << 8) | (x[d] >> 24)) & M
        x[c] = (x[c] + x[d]) & M; x[b] ^= x[c]; x[b] = ((x[b] << 7) | (x[b] >> 25)) & M

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & M for a, b in zip(x, st)]


class ChaCha20Rng:
    """`ChaChaRng::from_seed(seed)`: key = seed, counter and stream start at 0; `next_u64` takes
    two consecutive keystream words (low word first)."""

    def __init__(self, seed: bytes = bytes(32)):
        self.key = list(struct.unpack("<8I", seed))
        self.counter = 0
        self.buf: List[int] = []

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = chacha20_block(self.key, self.counter)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)


def fp_rand(field: Field, rng) -> int:
    """ark-ff 0.4 `impl Distribution<Fp> for Standard`: draw N u64 limbs, mask the top limb down to
    the modulus bit length, reject if >= modulus; the limbs ARE the Montgomery representation.
    Returns the canonical value."""
    nbits = field.p.bit_length()
    shave = 64 * field.limbs64 - nbits
    mask = (1 << 64) - 1 if shave == 0 else ((1 << 64) - 1) >> shave
    while True:
        limbs = [rng.next_u64() for _ in range(field.limbs64)]
        limbs[-1] &= mask
        v = field.from_limbs(limbs)
        if v < field.p:
            return field.from_mont(v)


def compute_coset_representatives(field: Field, num_wire_types: int, coset_size: int) -> List[int]:
    """relation/src/constants.rs:30-79."""
    p = field.p
    rng = ChaCha20Rng(bytes(32))
    ks, pows = [], []
    for i in range(num_wire_types):
        if i == 0:
            ks.append(1)
            pows.append(1)
            continue
        while True:
            nxt = fp_rand(field, rng)
            pw = pow(nxt, coset_size, p)
            if all(pow(prev, -1, p) * pw % p != 1 for prev in pows):
                break
        ks.append(nxt)
        pows.append(pw)
    return ks


# ======================================================================================
# Serialization (`to_bytes!` = ark-serialize `serialize_compressed`)
# ======================================================================================
def ser_fr(field: Field, x: int) -> bytes:
    return int(x % field.p).to_bytes(8 * field.limbs64, "little")


def ser_g1(curve: Curve, P) -> bytes:
    return curve.serialize_compressed(P)


def from_le_bytes_mod_order(field: Field, b: bytes) -> int:
    return int.from_bytes(b, "little") % field.p


# ======================================================================================
# Transcripts (plonk/src/transcript)
# ======================================================================================
class SolidityTranscript:
    """solidity.rs:31-78 -- labels ignored; the byte vector is never cleared and the challenge is
    not re-appended (the code, not its doc comment, is authoritative)."""

    def __init__(self, label: bytes = b"PlonkProof"):
        self.transcript = bytearray()
        self.state = bytes(64)

    def append_message(self, label: bytes, msg: bytes):
        self.transcript += msg

    def get_and_append_challenge(self, field: Field, label: bytes) -> int:
        base = self.state + bytes(self.transcript)
        self.state = keccak256(base + b"\x00") + keccak256(base + b"\x01")
        return from_le_bytes_mod_order(field, self.state[:48])


class StandardTranscript:
    """standard.rs:18-46 -- Merlin; challenge = 64 PRF bytes mod r, then re-appended under the label."""

    def __init__(self, label: bytes = b"PlonkProof"):
        self.t = MerlinTranscript(label)

    def append_message(self, label: bytes, msg: bytes):
        self.t.append_message(label, msg)

    def get_and_append_challenge(self, field: Field, label: bytes) -> int:
        c = from_le_bytes_mod_order(field, self.t.challenge_bytes(label, 64))
        self.t.append_message(label, ser_fr(field, c))
        return c


TRANSCRIPTS = {"solidity": SolidityTranscript, "standard": StandardTranscript}


def append_vk_and_pub_input(tr, curve: Curve, vk: dict, pub_input: Sequence[int]):
    """transcript/mod.rs:45-102."""
    fr = curve.fr
    tr.append_message(b"field size in bits", struct.pack("<I", fr.p.bit_length()))
    tr.append_message(b"domain size", struct.pack("<Q", vk["domain_size"]))
    tr.append_message(b"input size", struct.pack("<Q", vk["num_inputs"]))
    for k in vk["k"]:
        tr.append_message(b"wire subsets separators", ser_fr(fr, k))
    for c in vk["selector_comms"]:
        tr.append_message(b"selector commitments", ser_g1(curve, c))
    for c in vk["sigma_comms"]:
        tr.append_message(b"sigma commitments", ser_g1(curve, c))
    for x in pub_input:
        tr.append_message(b"public input", ser_fr(fr, x))


# ======================================================================================
# Circuit (TurboPlonk subset of `PlonkCircuit`)
# ======================================================================================
class Gate:
    def __init__(self, name, q_lc=(0, 0, 0, 0), q_mul=(0, 0), q_hash=(0, 0, 0, 0), q_o=0, q_c=0, q_ecc=0,
                 q_lookup=0, q_dom_sep=0, table_key=0, table_dom_sep=0):
        self.name, self.q_lc, self.q_mul, self.q_hash, self.q_o, self.q_c, self.q_ecc = name, q_lc, q_mul, q_hash, q_o, q_c, q_ecc
        # UltraPlonk (relation/src/gates/lookup.rs): lookup selector, domain separators and the table key of a LookupGate
        self.q_lookup, self.q_dom_sep, self.table_key, self.table_dom_sep = q_lookup, q_dom_sep, table_key, table_dom_sep

    def selectors(self) -> List[int]:
        return list(self.q_lc) + list(self.q_mul) + list(self.q_hash) + [self.q_o, self.q_c, self.q_ecc]


def ConstantGate(c): return Gate("const", q_c=c, q_o=1)
def AdditionGate(): return Gate("add", q_lc=(1, 1, 0, 0), q_o=1)
def SubtractionGate(p): return Gate("sub", q_lc=(1, p - 1, 0, 0), q_o=1)
def MultiplicationGate(): return Gate("mul", q_mul=(1, 0), q_o=1)
def EqualityGate(p): return Gate("eq", q_lc=(1, p - 1, 0, 0), q_o=1)
def IoGate(): return Gate("io", q_o=1)
def ConstantAdditionGate(c): return Gate("const_add", q_lc=(1, 0, 0, 0), q_c=c, q_o=1)
def LookupGate(q_dom_sep, table_dom_sep, table_key):
    return Gate("lookup", q_lookup=1, q_dom_sep=q_dom_sep, table_dom_sep=table_dom_sep, table_key=table_key)
def PaddingGate(): return Gate("pad")  # all selectors zero in this fork (relation/src/gates/mod.rs)
def ProofLinkingGate(): return Gate("link", q_mul=(1, 0))  # a * 0 = 0 (relation/src/gates/mod.rs:86-99)


PROOF_LINK_WIRE_IDX = 0  # relation/src/proof_linking/linkable_circuit.rs:22


class GroupLayout:
    """relation/src/proof_linking/mod.rs:17-53: `size` proof-linking gates on the 2^alignment-th roots of unity, from `offset`."""

    def __init__(self, alignment: int, offset: int, size: int):
        self.alignment, self.offset, self.size = alignment, offset, size

    def range_in_nth_roots(self, n: int) -> Tuple[int, int]:
        assert n >= self.alignment, "Group alignment must be <= n"
        spacing = 1 << (n - self.alignment)
        start = self.offset * spacing
        return start, start + max(self.size - 1, 0) * spacing

    def domain_generator(self, field: Field) -> int:
        return Radix2Domain(field, 1 << self.alignment).group_gen

    def __eq__(self, o): return (self.alignment, self.offset, self.size) == (o.alignment, o.offset, o.size)
    def __repr__(self): return "GroupLayout(alignment=%d, offset=%d, size=%d)" % (self.alignment, self.offset, self.size)


RANGE_WIRE_ID = 5  # relation/src/constraint_system.rs:77-85


def _next_pow2(x: int) -> int:
    n = 1
    while n < x:
        n <<= 1
    return n


def _place_group(size: int, n_inputs: int, alignment: int, gid: str, placed: list) -> bool:
    """place_group_with_alignment (linkable_circuit.rs:251-290): the first gap behind the public inputs that takes the group"""
    ranges = sorted(l.range_in_nth_roots(alignment) for _, l in placed)
    offset = n_inputs
    for idx, (start, end) in enumerate(ranges):
        if offset + size <= start:
            placed.insert(idx, (gid, GroupLayout(alignment, offset, size)))
            return True
        offset = end + 1
    if offset + size < (1 << alignment):
        placed.append((gid, GroupLayout(alignment, offset, size)))
        return True
    return False


class PlonkCircuit:
    def __init__(self, field: Field = pyref.BN254_FR, range_bit_len: Optional[int] = None):
        """`new_turbo_plonk()`; with `range_bit_len`: `new_ultra_plonk(range_bit_len)` (constraint_system.rs:193-240):
        a sixth wire type (the range / lookup wire), the q_lookup selector, the range table {0..2^bits-1} and
        key-value tables inserted with `create_table_and_lookup_variables`."""
        self.f = field
        self.range_bit_len = range_bit_len
        self.ultra = range_bit_len is not None
        self.nw = NUM_WIRE_TYPES + (1 if self.ultra else 0)
        self.witness = [0, 1]
        self.gates: List[Gate] = []
        self.wire_variables: List[List[int]] = [[] for _ in range(NUM_WIRE_TYPES + 1)]
        self.pub_input_gate_ids: List[int] = []
        self.num_table_elems = 0
        self.table_gate_ids: List[Tuple[int, int]] = []
        self.n = 1  # eval domain size; 1 == not finalized
        self.link_groups: dict = {}         # id -> member variables, in allocation order (constraint_system.rs:166)
        self.link_group_layouts: dict = {}  # id -> GroupLayout, where one is fixed (:172)
        self.enforce_constant(0, 0)
        self.enforce_constant(1, 1)

    # -- proof linking (constraint_system.rs:283-298,514-595; TurboPlonk only) ------------------------------
    def create_link_group(self, gid: str, layout: Optional["GroupLayout"] = None) -> str:
        assert not self.ultra, "only TurboPlonk supports link groups"
        self.link_groups[gid] = []
        if layout is not None:
            self.link_group_layouts[gid] = layout
        return gid

    def create_variable_with_link_groups(self, val: int, groups: Sequence[str]) -> int:
        assert self.n == 1
        v = self.create_variable(val)
        for g in groups:
            if g not in self.link_groups:
                raise ValueError("link group %s not found" % g)
            self.link_groups[g].append(v)
        return v

    def get_link_group_layout(self, gid: str) -> Optional["GroupLayout"]:
        return self.link_group_layouts.get(gid)

    def num_links(self) -> int:
        return sum(len(g) for g in self.link_groups.values())

    def _generate_layout(self):
        """`LinkableCircuit::generate_layout` (linkable_circuit.rs:136-178) -> (group layouts sorted by range, circuit size).
        Groups without a layout are placed in creation order (the reference walks a HashMap: any order with one such group)."""
        placed = [(g, l) for g, l in self.link_group_layouts.items() if g in self.link_groups]
        unplaced = [g for g in self.link_groups if g not in self.link_group_layouts]
        n_links = self.num_links()
        alignment = max([l.alignment for _, l in placed] + [(_next_pow2(n_links)).bit_length() - 1])   # min_alignment (:104-118)
        placed.sort(key=lambda t: t[1].range_in_nth_roots(alignment))
        inputs = self.num_inputs()
        for gid in unplaced:
            size = len(self.link_groups[gid])
            while not _place_group(size, inputs, alignment, gid, placed):
                alignment += 1
        layouts = dict(placed)
        # CircuitLayout::circuit_size / circuit_alignment (mod.rs:74-93)
        max_alignment = max(l.alignment for l in layouts.values()) if layouts else 1
        link_gates = sum(l.size for l in layouts.values())
        size = max(_next_pow2(self.num_gates() + link_gates), 1 << max_alignment)
        log_size = size.bit_length() - 1
        # validate_layout (linkable_circuit.rs:352-399)
        for gid, l in layouts.items():
            if l.size == 0:
                raise ValueError("Link group %s (layout = %r) is empty" % (gid, l))
            if l.offset + l.size >= (1 << l.alignment):
                raise ValueError("Link group %s (layout = %r) exceeds its alignment" % (gid, l))
            if l.range_in_nth_roots(log_size)[0] < inputs:
                raise ValueError("Link group %s (layout = %r) would mangle public inputs" % (gid, l))
        order = sorted(layouts.items(), key=lambda t: t[1].range_in_nth_roots(max_alignment))
        for (g1, l1), (g2, l2) in zip(order, order[1:]):
            r1, r2 = l1.range_in_nth_roots(log_size), l2.range_in_nth_roots(log_size)
            if max(r1[0], r2[0]) <= min(r1[1], r2[1]):
                raise ValueError("Link group %s (layout = %r) overlaps with group %s (layout = %r)" % (g1, l1, g2, l2))
        self.link_group_layouts.update(layouts)
        return order, size

    def _apply_layout(self, order, size):
        """`LinkableCircuit::apply_layout` (linkable_circuit.rs:184-238): inputs first, the proof-linking gates at their rows with
        the circuit's own gates in between, padding to `size`."""
        old_gates = iter(self.gates)
        old_vars = [iter(w) for w in self.wire_variables[:NUM_WIRE_TYPES]]
        gates: List[Gate] = []
        wires: List[List[int]] = [[] for _ in range(NUM_WIRE_TYPES)]

        def place(cnt):   # place_gates (:296-318)
            for _ in range(cnt):
                gates.append(next(old_gates, None) or PaddingGate())
                for j in range(NUM_WIRE_TYPES):
                    wires[j].append(next(old_vars[j], 0))

        place(self.num_inputs())
        log_size = size.bit_length() - 1
        for gid, l in order:
            start, _ = l.range_in_nth_roots(log_size)
            if start < len(gates):
                raise ValueError("link group %s does not fit behind the gates placed before it" % gid)
            place(start - len(gates))
            spacing = 1 << (log_size - l.alignment)
            for var in self.link_groups[gid]:   # append_group (:322-345)
                gates.append(ProofLinkingGate())
                for j in range(NUM_WIRE_TYPES):
                    wires[j].append(var if j == PROOF_LINK_WIRE_IDX else 0)
                place(spacing - 1)
        if len(gates) > size:
            raise ValueError("the circuit's gates do not fit its layout")
        place(size - len(gates))
        if next(old_gates, None) is not None:
            raise ValueError("the circuit's gates do not fit its layout")
        self.gates = gates
        for j in range(NUM_WIRE_TYPES):
            self.wire_variables[j] = wires[j]

    # -- UltraPlonk construction -----------------------------------------------------------------------
    def range_size(self) -> int:
        return 1 << self.range_bit_len

    def add_range_check_variable(self, var: int):
        assert self.ultra and self.n == 1
        self.wire_variables[RANGE_WIRE_ID].append(var)

    def add_constant(self, x: int, c: int) -> int:
        y = self.create_variable(self.witness[x] + c)
        self.insert_gate([x, 1, 0, 0, y], ConstantAdditionGate(c % self.f.p))
        return y

    def create_table_and_lookup_variables(self, lookup_vars, table_vars):
        """relation/src/gadgets/ultraplonk/lookup_table.rs:21-57"""
        assert self.ultra
        cnt = max(len(lookup_vars), len(table_vars))
        self.table_gate_ids.append((self.num_gates(), cnt))
        table_ctr = len(self.table_gate_ids)
        for i in range(cnt):
            q_dom_sep, key, v0, v1 = (table_ctr,) + tuple(lookup_vars[i]) if i < len(lookup_vars) else (0, 0, 0, 0)
            t_dom_sep, t_key, t0, t1 = (table_ctr, i) + tuple(table_vars[i]) if i < len(table_vars) else (0, 0, 0, 0)
            self.insert_gate([key, v0, v1, t0, t1], LookupGate(q_dom_sep, t_dom_sep, t_key))
        self.num_table_elems += cnt

    # -- construction ------------------------------------------------------------------------
    def zero(self): return 0
    def one(self): return 1
    def num_gates(self): return len(self.gates)
    def num_vars(self): return len(self.witness)
    def num_inputs(self): return len(self.pub_input_gate_ids)

    def create_variable(self, val: int) -> int:
        self.witness.append(val % self.f.p)
        return len(self.witness) - 1

    def insert_gate(self, wire_vars: Sequence[int], gate: Gate):
        assert self.n == 1 and len(wire_vars) == 5
        for j, v in enumerate(wire_vars):
            self.wire_variables[j].append(v)
        self.gates.append(gate)

    def set_variable_public(self, var: int):
        self.pub_input_gate_ids.append(self.num_gates())
        self.insert_gate([0, 0, 0, 0, var], IoGate())

    def create_public_variable(self, val: int) -> int:
        v = self.create_variable(val)
        self.set_variable_public(v)
        return v

    def enforce_constant(self, var: int, c: int): self.insert_gate([0, 0, 0, 0, var], ConstantGate(c % self.f.p))
    def enforce_equal(self, a: int, b: int): self.insert_gate([a, b, 0, 0, 0], EqualityGate(self.f.p))
    def add_gate(self, a, b, c): self.insert_gate([a, b, 0, 0, c], AdditionGate())
    def sub_gate(self, a, b, c): self.insert_gate([a, b, 0, 0, c], SubtractionGate(self.f.p))
    def mul_gate(self, a, b, c): self.insert_gate([a, b, 0, 0, c], MultiplicationGate())

    def add(self, a, b):
        c = self.create_variable(self.witness[a] + self.witness[b])
        self.add_gate(a, b, c)
        return c

    def sub(self, a, b):
        c = self.create_variable(self.witness[a] - self.witness[b])
        self.sub_gate(a, b, c)
        return c

    def mul(self, a, b):
        c = self.create_variable(self.witness[a] * self.witness[b])
        self.mul_gate(a, b, c)
        return c

    def lc_gate(self, wires: Sequence[int], coeffs: Sequence[int]):
        """LinCombGate (relation/src/traits.rs:258-276): q_lc = coeffs, q_o = 1"""
        self.insert_gate(list(wires), Gate("lc", q_lc=tuple(c % self.f.p for c in coeffs), q_o=1))

    def lc(self, wires_in: Sequence[int], coeffs: Sequence[int]) -> int:
        y = self.create_variable(sum(self.witness[v] * c for v, c in zip(wires_in, coeffs)))
        self.lc_gate(list(wires_in) + [y], coeffs)
        return y

    def sum(self, elems: Sequence[int]) -> int:
        """relation/src/traits.rs:369-408: z_0 = x_0, z_i = z_{i-1} + x_{3i-2} + x_{3i-1} + x_{3i}, the last step lands on `sum`"""
        assert elems
        total = self.create_variable(sum(self.witness[v] for v in elems))
        rate = GATE_WIDTH - 1
        padded_len = -(-(len(elems) - 1) // rate) * rate + 1
        padded = list(elems) + [0] * (padded_len - len(elems))
        accum = padded[0]
        for i in range(1, padded_len // rate):
            accum = self.lc([accum, padded[rate * i - 2], padded[rate * i - 1], padded[rate * i]], [1, 1, 1, 1])
        self.lc_gate([accum, padded[padded_len - 3], padded[padded_len - 2], padded[padded_len - 1], total], [1, 1, 1, 1])
        return total

    def public_input(self) -> List[int]:
        return [self.witness[self.wire_variables[GATE_WIDTH][g]] for g in self.pub_input_gate_ids]

    # -- finalize (TurboPlonk without link groups; UltraPlonk) ------------------------------------------
    def finalize_for_arithmetization(self):
        if self.n != 1:
            return
        order = []
        if not self.ultra:
            # generate_layout + apply_layout (constraint_system.rs:972-980); without link groups: next_power_of_two(n_gates)
            order, n_gates = self._generate_layout()
        else:  # range gates and lookup gates need separate slots (constraint_system.rs:981-987)
            n_gates = max(self.num_gates(),
                          max(self.range_size(), len(self.wire_variables[RANGE_WIRE_ID])) + self.num_table_elems + 1)
        n = 1
        while n < n_gates:
            n <<= 1
        if self.ultra:
            # pad (:675-686) first, then rearrange (:630-666)
            while len(self.gates) < n:
                self.gates.append(PaddingGate())
            for j in range(self.nw):
                w = self.wire_variables[j]
                w.extend([0] * (n - len(w)))
        # rearrange_gates: io gates to the front (constraint_system.rs:630-645)
        for gate_id, io_gate_id in enumerate(list(self.pub_input_gate_ids)):
            if io_gate_id > gate_id:
                self.gates[gate_id], self.gates[io_gate_id] = self.gates[io_gate_id], self.gates[gate_id]
                for j in range(NUM_WIRE_TYPES):
                    w = self.wire_variables[j]
                    w[gate_id], w[io_gate_id] = w[io_gate_id], w[gate_id]
                self.pub_input_gate_ids[gate_id] = gate_id
        if self.ultra:
            # lookup gates to the rear, relative order kept, never the very last slot (:647-664)
            cur = n - 2
            for table_gate_id, table_size in reversed(self.table_gate_ids):
                for gate_id in reversed(range(table_gate_id, table_gate_id + table_size)):
                    if gate_id < cur:
                        self.gates[gate_id], self.gates[cur] = self.gates[cur], self.gates[gate_id]
                        for j in range(NUM_WIRE_TYPES):
                            w = self.wire_variables[j]
                            w[gate_id], w[cur] = w[cur], w[gate_id]
                        cur -= 1
        else:
            # proof-linking gates at their rows; pad with PaddingGate / variable 0 (linkable_circuit.rs:184-238,294-314)
            self._apply_layout(order, n)
        self.n = n
        self.domain = Radix2Domain(self.f, n)
        self._compute_wire_permutation()
        self.k = compute_coset_representatives(self.f, self.nw, n)
        p = self.f.p
        g = self.domain.group_gen
        elems = [1] * n
        for j in range(1, n):
            elems[j] = elems[j - 1] * g % p
        self.extended_id_permutation = [ki * e % p for ki in self.k for e in elems]

    def _compute_wire_permutation(self):
        n = self.n
        var_map: List[List[Tuple[int, int]]] = [[] for _ in range(self.num_vars())]
        for wire_id in range(self.nw):
            for gate_id, var in enumerate(self.wire_variables[wire_id]):
                var_map[var].append((wire_id, gate_id))
        self.wire_permutation = [(0, 0)] * (self.nw * n)
        for wires in var_map:
            if wires:
                cyc = wires + [wires[0]]
                for a, b in zip(cyc[:-1], cyc[1:]):
                    self.wire_permutation[a[0] * n + a[1]] = b

    # -- arithmetization inputs ----------------------------------------------------------------------
    def selector_evals(self) -> List[List[int]]:
        """all_selectors (constraint_system.rs:890-905): the 13 TurboPlonk columns, then q_lookup when lookups are supported"""
        cols = [[0] * self.n for _ in range(N_SELECTORS + (1 if self.ultra else 0))]
        for i, g in enumerate(self.gates):
            for s, v in enumerate(g.selectors()):
                cols[s][i] = v % self.f.p
            if self.ultra:
                cols[N_SELECTORS][i] = g.q_lookup % self.f.p
        return cols

    # -- Plookup arithmetization (constraint_system.rs:1261-1492) ------------------------------------------
    def q_lookup(self): return [g.q_lookup % self.f.p for g in self.gates]
    def q_dom_sep(self): return [g.q_dom_sep % self.f.p for g in self.gates]
    def table_key_vec(self): return [g.table_key % self.f.p for g in self.gates]
    def table_dom_sep_vec(self): return [g.table_dom_sep % self.f.p for g in self.gates]

    def range_table(self) -> List[int]:
        assert self.n >= self.range_size(), "Domain size < range size"
        return list(range(self.range_size())) + [0] * (self.n - self.range_size())

    def _wit(self, wire: int, i: int) -> int:
        return self.witness[self.wire_variables[wire][i]]

    def merged_lookup_table(self, tau: int) -> List[int]:
        p = self.f.p
        rt, tk, td, ql = self.range_table(), self.table_key_vec(), self.table_dom_sep_vec(), self.q_lookup()
        return [(rt[i] + ql[i] * tau % p * ((td[i] + tau * ((tk[i] + tau * ((self._wit(3, i) + tau * self._wit(4, i)) % p)) % p)) % p)) % p
                for i in range(self.n)]

    def merged_lookup_wire_value(self, tau: int, i: int, ql, qd) -> int:
        p = self.f.p
        return (self._wit(RANGE_WIRE_ID, i)
                + ql[i] * tau % p * ((qd[i] + tau * ((self._wit(0, i) + tau * ((self._wit(1, i) + tau * self._wit(2, i)) % p)) % p)) % p)) % p

    def lookup_sorted_vec(self, tau: int, merged_table: Sequence[int]) -> List[int]:
        """compute_lookup_sorted_vec_polynomials (:1370-1418): the lookup values merged into the table, in table order."""
        n = self.n
        ql, qd = self.q_lookup(), self.q_dom_sep()
        counts: dict = {}
        for i in range(n - 1):
            e = self.merged_lookup_wire_value(tau, i, ql, qd)
            counts[e] = counts.get(e, 0) + 1
        out = []
        for e in merged_table:
            if e in counts:
                out.extend([e] * (1 + counts.pop(e)))
            else:
                out.append(e)
        if len(out) != 2 * n - 1:
            raise ValueError("The sorted vector has wrong length, some lookup variables might be outside the table")
        return out

    def lookup_prod_vec(self, tau: int, beta: int, gamma: int, merged_table: Sequence[int], sorted_vec: Sequence[int]) -> List[int]:
        """compute_lookup_prod_polynomial (:1311-1368), evaluations on the domain"""
        p, n = self.f.p, self.n
        ql, qd = self.q_lookup(), self.q_dom_sep()
        bp1 = (1 + beta) % p
        gb = gamma * bp1 % p
        prod = [1]
        for j in range(n - 2):
            lw = self.merged_lookup_wire_value(tau, j, ql, qd)
            a = bp1 * ((gamma + lw) % p) % p * ((gb + merged_table[j] + beta * merged_table[j + 1]) % p) % p
            b = (gb + sorted_vec[j] + beta * sorted_vec[j + 1]) % p * ((gb + sorted_vec[n - 1 + j] + beta * sorted_vec[n + j]) % p) % p
            prod.append(prod[-1] * a % p * pow(b, -1, p) % p)
        prod.append(1)
        return prod

    def extended_permutation(self) -> List[int]:
        n = self.n
        return [self.extended_id_permutation[w * n + g] for (w, g) in self.wire_permutation]

    def wire_values(self) -> List[List[int]]:
        return [[self.witness[v] for v in self.wire_variables[j]] for j in range(self.nw)]

    def check_satisfiability(self) -> bool:
        p = self.f.p
        pi = [0] * self.n
        for g in self.pub_input_gate_ids:
            pi[g] = self.witness[self.wire_variables[GATE_WIDTH][g]]
        for i, g in enumerate(self.gates):
            w = [self.witness[self.wire_variables[j][i]] for j in range(5)]
            v = (g.q_c + pi[i] + sum(q * x for q, x in zip(g.q_lc, w)) + g.q_mul[0] * w[0] * w[1] + g.q_mul[1] * w[2] * w[3]
                 + g.q_ecc * w[0] * w[1] * w[2] * w[3] * w[4] + sum(q * pow(x, 5, p) for q, x in zip(g.q_hash, w))
                 - g.q_o * w[4]) % p
            if v:
                return False
        if self.ultra:  # constraint_system.rs:405-449
            nrange = len(self.wire_variables[RANGE_WIRE_ID])
            if any(self._wit(RANGE_WIRE_ID, i) >= self.range_size() for i in range(nrange)):
                return False
            table = {(0, 0, 0, 0)}
            for i, g in enumerate(self.gates):
                if g.q_lookup:
                    table.add((g.table_dom_sep % p, g.table_key % p, self._wit(3, i), self._wit(4, i)))
            for i, g in enumerate(self.gates):
                if g.q_lookup and (g.q_dom_sep % p, self._wit(0, i), self._wit(1, i), self._wit(2, i)) not in table:
                    return False
        return True


def gen_circuit_all_selectors(m: int, field: Field = pyref.BN254_FR) -> PlonkCircuit:
    """A satisfiable circuit that drives every one of the 13 selector columns (q_lc0-3, q_mul0-1, q_hash0-3,
    q_o, q_c, q_ecc) with non-trivial values, so that every term of the quotient / linearisation formulas
    (prover.rs:696-708, 963-1003) is exercised.  Gates are generic `insert_gate` calls with the selector
    values of the reference's gate types (relation/src/gates/arithmetic.rs: LinCombGate, MulAddGate,
    FifthRootGate, and an ecc-style degree-5 product gate)."""
    p = field.p
    cs = PlonkCircuit(field)
    pub = cs.create_public_variable(11)
    v = [cs.create_variable(3 + 2 * i) for i in range(4 * m)]
    for i in range(m):
        a, b, c, d = v[4 * i:4 * i + 4]
        wa, wb, wc, wd = (cs.witness[x] for x in (a, b, c, d))
        # LinCombGate: q_lc = (2, 3, 5, 7), q_o = 1
        e = cs.create_variable(2 * wa + 3 * wb + 5 * wc + 7 * wd)
        cs.insert_gate([a, b, c, d, e], Gate("lincomb", q_lc=(2, 3, 5, 7), q_o=1))
        # MulAddGate: q_mul = (4, 9), q_o = 1
        f = cs.create_variable(4 * wa * wb + 9 * wc * wd)
        cs.insert_gate([a, b, c, d, f], Gate("muladd", q_mul=(4, 9), q_o=1))
        # hash-style gate: q_hash = (1, 2, 3, 4), q_c = 6, q_o = 1
        h = cs.create_variable(pow(wa, 5, p) + 2 * pow(wb, 5, p) + 3 * pow(wc, 5, p) + 4 * pow(wd, 5, p) + 6)
        cs.insert_gate([a, b, c, d, h], Gate("hash", q_hash=(1, 2, 3, 4), q_c=6, q_o=1))
        # ecc-style gate: q_ecc * w0 w1 w2 w3 w4 + q_lc0 w0 + q_c = q_o w4 with q_o = 0:  the product is pinned
        # by the constant: choose w4 = e, q_ecc = 5, q_c = -(5 a b c d e + 8 a)
        we = cs.witness[e]
        cs.insert_gate([a, b, c, d, e], Gate("ecc", q_lc=(8, 0, 0, 0), q_ecc=5, q_c=(-(5 * wa * wb * wc * wd * we + 8 * wa)) % p))
    s2 = cs.add(pub, v[0])
    cs.mul_gate(s2, v[1], cs.create_variable(cs.witness[s2] * cs.witness[v[1]]))
    cs.finalize_for_arithmetization()
    return cs


def gen_circuit_for_bench(num_gates: int, field: Field = pyref.BN254_FR, ultra: bool = False) -> PlonkCircuit:
    """plonk/benches/bench.rs:29-46 (`new_turbo_plonk()` / `new_ultra_plonk(RANGE_BIT_LEN = 8)`)."""
    cs = PlonkCircuit(field, 8 if ultra else None)
    a = cs.zero()
    for _ in range(num_gates - 10):
        a = cs.add(a, cs.one())
    cs.finalize_for_arithmetization()
    return cs


def gen_circuit_for_test(m: int, a0: int, field: Field = pyref.BN254_FR, ultra: bool = False) -> PlonkCircuit:
    """plonk/src/proof_system/snark.rs:681-744 (both branches; UltraPlonk: range_bit_len = 5, a0 <= m + 1)."""
    cs = PlonkCircuit(field, 5 if ultra else None)
    a = [cs.create_variable(i) for i in range(a0, a0 + 4 * m)]
    b = [cs.create_public_variable(m * 2), cs.create_public_variable(a0 * 2 + m * 4 - 1)]
    c = cs.create_public_variable((cs.witness[b[1]] + cs.witness[a[0]]) * (cs.witness[b[1]] - cs.witness[a[0]]))
    acc = cs.zero()
    for e in a:
        acc = cs.add(acc, e)
    b_mul = cs.mul(b[0], b[1])
    cs.enforce_equal(acc, b_mul)
    p1 = cs.add(b[1], a[0])
    m1 = cs.sub(b[1], a[0])
    cs.mul_gate(p1, m1, c)
    cs.enforce_constant(b[0], m * 2)
    if ultra:
        # range gates: a_i in {0..31} for i < m, b0 in the table; one key-value table with two lookups (:722-738)
        for var in a[:m]:
            cs.add_range_check_variable(var)
        cs.add_range_check_variable(b[0])
        table_vars = [(a[0], a[2]), (a[1], a[3]), (b[0], a[0])]
        key0 = cs.one()
        key1 = cs.create_variable(2)
        two_m = cs.create_public_variable(m * 2)
        a1 = cs.add_constant(a[0], 1)
        a3 = cs.add_constant(a[0], 3)
        cs.create_table_and_lookup_variables([(key0, a1, a3), (key1, two_m, a[0])], table_vars)
    cs.finalize_for_arithmetization()
    return cs


# ======================================================================================
# Polynomial helpers (ark-poly DensePolynomial semantics; coefficient lists, low degree first)
# ======================================================================================
def _strip(c: List[int]) -> List[int]:
    c = list(c)
    while c and c[-1] == 0:
        c.pop()
    return c


def _poly_add(p: int, a: Sequence[int], b: Sequence[int]) -> List[int]:
    n = max(len(a), len(b))
    return _strip([((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % p for i in range(n)])


def _poly_scale(p: int, a: Sequence[int], s: int) -> List[int]:
    return _strip([x * s % p for x in a])


def _poly_eval(p: int, c: Sequence[int], x: int) -> int:
    acc = 0
    for v in reversed(c):
        acc = (acc * x + v) % p
    return acc


def _div_linear(p: int, c: Sequence[int], z: int) -> List[int]:
    """quotient of c(X) / (X - z), remainder dropped (ark-poly `/`)."""
    if len(c) < 2:
        return []
    q = [0] * (len(c) - 1)
    carry = 0
    for i in range(len(c) - 1, 0, -1):
        carry = (c[i] + carry * z) % p
        q[i - 1] = carry
    return _strip(q)


class _Backend:
    """NTT / MSM through the C oracle when present (sizes >= 2^8), else exact Python."""

    def __init__(self, curve: Curve):
        self.curve, self.fr = curve, curve.fr
        try:
            import coracle
            coracle.build()
            self.co = coracle
        except Exception:  # pragma: no cover
            self.co = None

    def ntt(self, vals: Sequence[int], log_n: int, inverse: bool, offset: int = 1) -> List[int]:
        n = 1 << log_n
        f = self.fr
        assert len(vals) <= n
        if self.co is None or n < 64:
            dom = Radix2Domain(f, n, offset)
            return dom.ifft(list(vals) + [0] * (n - len(vals))) if inverse else dom.fft(list(vals))
        co = self.co
        R = f.R
        arr = co.ints_to_limbs([v * R % f.p for v in vals] + [0] * (n - len(vals)), 4)
        off = None if offset == 1 else co.ints_to_limbs([offset * R % f.p], 4)[0]
        out = co.ntt(f.name, arr, log_n, inverse, off)
        ri = pow(R, -1, f.p)
        return [v * ri % f.p for v in co.limbs_to_ints(out)]

    def commit(self, srs_limbs, srs_points, coeffs: Sequence[int]):
        """UnivariateKzgPCS::commit (mod.rs:90-116): skip low-order zeros, msm_bigint, into_affine."""
        c = list(coeffs)
        nz = 0
        while nz < len(c) and c[nz] == 0:
            nz += 1
        c = c[nz:]
        if not c:
            return None
        if self.co is not None and srs_limbs is not None:
            xy, inf = self.co.msm(self.curve.name, srs_limbs[nz:], self.co.ints_to_limbs(c, 4))
            if inf:
                return None
            fq = self.curve.fq
            x, y = self.co.limbs_to_ints(xy.reshape(2, fq.limbs64))
            return (fq.from_mont(x), fq.from_mont(y))
        return self.curve.msm_pippenger(c, srs_points[nz:nz + len(c)])


# ======================================================================================
# preprocess / prove / verify
# ======================================================================================
def quotient_domain_size(n: int, num_wire_types: int = NUM_WIRE_TYPES) -> int:
    """domain_size_ratio (plonk/src/constants.rs:18-20) then GeneralEvaluationDomain::new -> radix 2 (prover.rs:54-62)."""
    ratio = (num_wire_types * (n + 1) + 2) // n + 1
    m = 1
    while m < n * ratio:
        m <<= 1
    return m


def gen_srs(curve: Curve, beta: int, max_degree: int):
    """powers_of_g = [beta^i] g, i <= max_degree (srs.rs:118-153, known beta).  Returns
    (limb array or None, python points or None): the C oracle generates large keys."""
    try:
        import coracle
        coracle.build()
        lim = coracle.gen_srs(curve.name, coracle.ints_to_limbs([beta], 4)[0], max_degree + 1)
        return lim, None
    except Exception:  # pragma: no cover
        return None, pyref.gen_srs_for_testing(curve, beta, max_degree)


def preprocess(curve: Curve, srs, cs: PlonkCircuit) -> dict:
    """snark.rs:529-611: selector / sigma (and, for UltraPlonk, the four Plookup table) polynomials (ifft) and their commitments."""
    be = _Backend(curve)
    n = cs.n
    log_n = n.bit_length() - 1
    srs_limbs, srs_points = srs
    selectors = [_strip(be.ntt(col, log_n, True)) for col in cs.selector_evals()]
    ext = cs.extended_permutation()
    sigmas = [_strip(be.ntt(ext[i * n:(i + 1) * n], log_n, True)) for i in range(cs.nw)]
    vk = {
        "domain_size": n, "num_inputs": cs.num_inputs(),
        "selector_comms": [be.commit(srs_limbs, srs_points, s) for s in selectors],
        "sigma_comms": [be.commit(srs_limbs, srs_points, s) for s in sigmas],
        "k": list(cs.k), "plookup": None,
    }
    pk = {"selectors": selectors, "sigmas": sigmas, "vk": vk, "srs": srs, "n": n, "plookup": None}
    if cs.ultra:
        lk = {"range_table_poly": _strip(be.ntt(cs.range_table(), log_n, True)),
              "key_table_poly": _strip(be.ntt(cs.table_key_vec(), log_n, True)),
              "table_dom_sep_poly": _strip(be.ntt(cs.table_dom_sep_vec(), log_n, True)),
              "q_dom_sep_poly": _strip(be.ntt(cs.q_dom_sep(), log_n, True))}
        pk["plookup"] = lk
        vk["plookup"] = {"range_table_comm": be.commit(srs_limbs, srs_points, lk["range_table_poly"]),
                         "key_table_comm": be.commit(srs_limbs, srs_points, lk["key_table_poly"]),
                         "table_dom_sep_comm": be.commit(srs_limbs, srs_points, lk["table_dom_sep_poly"]),
                         "q_dom_sep_comm": be.commit(srs_limbs, srs_points, lk["q_dom_sep_poly"])}
    return pk


def num_blinders(cs_or_nw, ultra: Optional[bool] = None) -> int:
    """field elements `prove` draws from the prng: 2 per wire polynomial, 3 for z, nw - 1 split-quotient randomizers and, for
    UltraPlonk, 3 each for h1, h2 and the lookup product (17 / 29)."""
    nw = cs_or_nw if isinstance(cs_or_nw, int) else cs_or_nw.nw
    ultra = (nw == 6) if ultra is None else ultra
    return 2 * nw + 3 + (nw - 1) + (9 if ultra else 0)


def _merged_table(p, tau, rng, key, ql, w3, w4, tds):
    """eval_merged_table (structs.rs:926-939)"""
    return (rng + ql * tau % p * ((tds + tau * ((key + tau * ((w3 + tau * w4) % p)) % p)) % p)) % p


def _merged_lookup(p, tau, wr, w0, w1, w2, ql, qds):
    """eval_merged_lookup_witness (structs.rs:943-956)"""
    return (wr + ql * tau % p * ((qds + tau * ((w0 + tau * ((w1 + tau * w2) % p)) % p)) % p)) % p


PLOOKUP_EVAL_FIELDS = ["range_table_eval", "key_table_eval", "table_dom_sep_eval", "q_dom_sep_eval", "h_1_eval", "q_lookup_eval",
                       "prod_next_eval", "range_table_next_eval", "key_table_next_eval", "table_dom_sep_next_eval", "h_1_next_eval",
                       "h_2_next_eval", "q_lookup_next_eval", "w_3_next_eval", "w_4_next_eval"]  # structs.rs:496-541, declaration order


def _plookup_evals_vec(e):       # structs.rs:545-554
    return [e["range_table_eval"], e["key_table_eval"], e["h_1_eval"], e["q_lookup_eval"], e["table_dom_sep_eval"], e["q_dom_sep_eval"]]


def _plookup_next_evals_vec(e):  # structs.rs:557-569
    return [e["prod_next_eval"], e["range_table_next_eval"], e["key_table_next_eval"], e["h_1_next_eval"], e["h_2_next_eval"],
            e["q_lookup_next_eval"], e["w_3_next_eval"], e["w_4_next_eval"], e["table_dom_sep_next_eval"]]


def _append_plookup_evals(tr, fr, e):
    """transcript/mod.rs:168-201: only six of the fifteen evaluations enter the transcript"""
    for label, key in ((b"lookup_table_eval", "range_table_eval"), (b"h_1_eval", "h_1_eval"), (b"prod_next_eval", "prod_next_eval"),
                       (b"lookup_table_next_eval", "range_table_next_eval"), (b"h_1_next_eval", "h_1_next_eval"),
                       (b"h_2_next_eval", "h_2_next_eval")):
        tr.append_message(label, ser_fr(fr, e[key]))


def prove(curve: Curve, cs: PlonkCircuit, pk: dict, blinders: Sequence[int], transcript: str = "solidity",
          extra_msg: Optional[bytes] = None) -> dict:
    """`PlonkKzgSnark::prove` = batch_prove_internal for ONE instance (snark.rs:624-651), TurboPlonk or UltraPlonk.  `blinders`: the
    field elements the reference draws from its prng, in consumption order: nw x (b0, b1) for the wire
    polynomials, [3 for h1, 3 for h2,] 3 for z, [3 for the lookup product,] nw - 1 split-quotient randomizers: 17 / 29."""
    bp = batch_prove(curve, [cs], [pk], blinders, transcript, extra_msg)
    return {   # `From<BatchProof> for Proof` (structs.rs:302-317)
        "wires_poly_comms": bp["wires_poly_comms_vec"][0], "prod_perm_poly_comm": bp["prod_perm_poly_comms_vec"][0],
        "split_quot_poly_comms": bp["split_quot_poly_comms"], "opening_proof": bp["opening_proof"],
        "shifted_opening_proof": bp["shifted_opening_proof"], "wires_evals": bp["poly_evals_vec"][0]["wires_evals"],
        "wire_sigma_evals": bp["poly_evals_vec"][0]["wire_sigma_evals"], "perm_next_eval": bp["poly_evals_vec"][0]["perm_next_eval"],
        "plookup_proof": bp["plookup_proofs_vec"][0], "challenges": bp["challenges"]}


def prove_with_link_hint(curve: Curve, cs: PlonkCircuit, pk: dict, blinders: Sequence[int], transcript: str = "solidity"):
    """`PlonkKzgSnark::prove_with_link_hint` (snark.rs:81-114) -> (proof, LinkingHint): the hint is the first wire polynomial
    (after masking) and its commitment (structs.rs:88-97)."""
    bp = batch_prove(curve, [cs], [pk], blinders, transcript, None)
    proof = {
        "wires_poly_comms": bp["wires_poly_comms_vec"][0], "prod_perm_poly_comm": bp["prod_perm_poly_comms_vec"][0],
        "split_quot_poly_comms": bp["split_quot_poly_comms"], "opening_proof": bp["opening_proof"],
        "shifted_opening_proof": bp["shifted_opening_proof"], "wires_evals": bp["poly_evals_vec"][0]["wires_evals"],
        "wire_sigma_evals": bp["poly_evals_vec"][0]["wire_sigma_evals"], "perm_next_eval": bp["poly_evals_vec"][0]["perm_next_eval"],
        "plookup_proof": bp["plookup_proofs_vec"][0], "challenges": bp["challenges"]}
    hint = {"linking_wire_poly": list(bp["wire_polys_vec"][0][PROOF_LINK_WIRE_IDX]),
            "linking_wire_comm": bp["wires_poly_comms_vec"][0][PROOF_LINK_WIRE_IDX]}
    return proof, hint


# ======================================================================================
# Proof linking (plonk/src/proof_system/proof_linking.rs)
# ======================================================================================
def _link_roots(fr: Field, layout: GroupLayout) -> List[int]:
    """the roots of the linking domain's vanishing polynomial: g^offset .. g^(offset + size - 1), g the 2^alignment-th root of
    unity (proof_linking.rs:137-160)"""
    g = layout.domain_generator(fr)
    r = pow(g, layout.offset, fr.p)
    out = []
    for _ in range(layout.size):
        out.append(r)
        r = r * g % fr.p
    return out


def _poly_sub(p: int, a: Sequence[int], b: Sequence[int]) -> List[int]:
    return _poly_add(p, a, [(-x) % p for x in b])


def linking_quotient(fr: Field, a1: Sequence[int], a2: Sequence[int], layout: GroupLayout) -> List[int]:
    """compute_linking_quotient (proof_linking.rs:116-135): (a1 - a2) / Z_D, remainder dropped.  ark-poly's `/` is schoolbook long
    division; dividing by the linear factors of Z_D one after the other yields the same quotient (the remainders of the steps
    combine to one of degree < size)."""
    p = fr.p
    if _strip(a1) == _strip(a2):
        return []
    q = _poly_sub(p, a1, a2)
    for r in _link_roots(fr, layout):
        q = _div_linear(p, q, r)
    return q


def linking_quotient_schoolbook(fr: Field, a1: Sequence[int], a2: Sequence[int], layout: GroupLayout) -> List[int]:
    """the literal form (test only): build Z_D by multiplying its monomials (:137-160), then long division"""
    p = fr.p
    z = [1]
    for r in _link_roots(fr, layout):
        nz = [0] * (len(z) + 1)
        for i, c in enumerate(z):
            nz[i] = (nz[i] - c * r) % p
            nz[i + 1] = (nz[i + 1] + c) % p
        z = nz
    rem = _poly_sub(p, a1, a2)
    if len(rem) < len(z):
        return []
    q = [0] * (len(rem) - len(z) + 1)
    for k in range(len(q) - 1, -1, -1):   # Z_D is monic
        c = rem[k + len(z) - 1]
        q[k] = c
        if c:
            for i, zc in enumerate(z):
                rem[k + i] = (rem[k + i] - c * zc) % p
    return _strip(q)


def _link_challenge(curve: Curve, a1_comm, a2_comm, q_comm, transcript: str) -> int:
    """compute_quotient_challenge (:171-191): a fresh transcript over the two wire commitments and the quotient commitment"""
    tr = TRANSCRIPTS[transcript](b"PlonkLinkingProof")
    for c in (a1_comm, a2_comm):
        tr.append_message(b"linking_wire_comms", ser_g1(curve, c))
    tr.append_message(b"quotient_comm", ser_g1(curve, q_comm))
    return tr.get_and_append_challenge(curve.fr, b"eta")


def _link_vanishing_eval(fr: Field, eta: int, layout: GroupLayout) -> int:
    ev = 1
    for r in _link_roots(fr, layout):   # compute_vanishing_poly_eval (:162-177)
        ev = ev * (eta - r) % fr.p
    return ev


def link_proofs(curve: Curve, lhs_hint: dict, rhs_hint: dict, layout: GroupLayout, srs, transcript: str = "solidity") -> dict:
    """`PlonkKzgSnark::link_proofs` (proof_linking.rs:79-112) -> LinkingProof {quotient_commitment, opening_proof}.
    srs: (limb array | None, python points | None) as returned by gen_srs."""
    fr, p = curve.fr, curve.fr.p
    be = _Backend(curve)
    srs_limbs, srs_points = srs
    a1, a2 = lhs_hint["linking_wire_poly"], rhs_hint["linking_wire_poly"]
    quotient = linking_quotient(fr, a1, a2, layout)
    q_comm = be.commit(srs_limbs, srs_points, quotient)
    eta = _link_challenge(curve, lhs_hint["linking_wire_comm"], rhs_hint["linking_wire_comm"], q_comm, transcript)
    # compute_identity_opening (:193-216): a1 - a2 - q * Z_D(eta), opened at eta (`UnivariateKzgPCS::open`, mod.rs:135-161)
    zd = _link_vanishing_eval(fr, eta, layout)
    identity = _poly_sub(p, _poly_sub(p, a1, a2), _poly_scale(p, quotient, zd))
    opening = be.commit(srs_limbs, srs_points, _div_linear(p, identity, eta))
    return {"quotient_commitment": q_comm, "opening_proof": opening, "eta": eta}


def verify_link_proof(curve: Curve, proof1: dict, proof2: dict, link_proof: dict, layout: GroupLayout, beta_srs: int,
                      transcript: str = "solidity") -> bool:
    """`verify_link_proof` (proof_linking.rs:233-283); `UnivariateKzgPCS::verify` (mod.rs:218-243) e(C - v g, h) == e(pi, [beta - z] h)
    is evaluated in G1 with the known trapdoor: C - v g == (beta - z) pi, here with v = 0, z = eta."""
    fr, p = curve.fr, curve.fr.p
    a1c, a2c = proof1["wires_poly_comms"][PROOF_LINK_WIRE_IDX], proof2["wires_poly_comms"][PROOF_LINK_WIRE_IDX]
    qc = link_proof["quotient_commitment"]
    eta = _link_challenge(curve, a1c, a2c, qc, transcript)
    zd = _link_vanishing_eval(fr, eta, layout)
    # compute_identity_commitment (:285-299)
    ident = curve.add(curve.add(a1c, curve.neg(a2c)), curve.neg(curve.mul(zd, qc)))
    return ident == curve.mul((beta_srs - eta) % p, link_proof["opening_proof"])


def serialize_link_proof(curve: Curve, lp: dict) -> bytes:
    """`LinkingProof<E>` CanonicalSerialize, compressed (proof_linking.rs:32-39): the quotient commitment, then the opening proof"""
    return ser_g1(curve, lp["quotient_commitment"]) + ser_g1(curve, lp["opening_proof"])


LINK_GROUP_NAME = "test_group"  # proof_linking.rs:310


def gen_link_test_circuit(which: int, witness: Sequence[int], layout: Optional[GroupLayout], field: Field = pyref.BN254_FR) -> PlonkCircuit:
    """`gen_test_circuit1` (a sum) / `gen_test_circuit2` (a product) of proof_linking.rs:330-407, finalized."""
    cs = PlonkCircuit(field)
    p = field.p
    if which == 1:
        expected = cs.create_public_variable(sum(witness) % p)
    else:
        prod = 1
        for w in witness:
            prod = prod * w % p
        expected = cs.create_public_variable(prod)
    for w in witness:               # a few public inputs, for spacing
        cs.create_public_variable(w)
    group = cs.create_link_group(LINK_GROUP_NAME, layout)
    wvars = [cs.create_variable_with_link_groups(w, [group]) for w in witness]
    for w in witness:               # a few witnesses that are not linked
        cs.create_variable(w * w)
    if which == 1:
        cs.enforce_equal(cs.sum(wvars), expected)
    else:
        prod = cs.one()
        for v in wvars:
            prod = cs.mul(prod, v)
        cs.enforce_equal(prod, expected)
    cs.finalize_for_arithmetization()
    return cs


def batch_num_blinders(circuits) -> int:
    nw, ultra = circuits[0].nw, circuits[0].ultra
    return len(circuits) * (2 * nw + 3 + (9 if ultra else 0)) + (nw - 1)


def batch_prove(curve: Curve, circuits: Sequence[PlonkCircuit], pks: Sequence[dict], blinders: Sequence[int],
                transcript: str = "solidity", extra_msg: Optional[bytes] = None) -> dict:
    """batch_prove_internal (snark.rs:201-469): several instances over the same domain share one transcript, ONE quotient
    polynomial (instance i enters with alpha_base_i = (alpha^3 | alpha^7)^i) and ONE pair of opening proofs.  `blinders` in the
    order the reference's single prng is consumed: the wire masks of every instance, [the h1 / h2 masks of every instance,] the z
    masks of every instance, [the lookup-product masks,] then the nw - 1 split-quotient randomizers."""
    fr, p = curve.fr, curve.fr.p
    be = _Backend(curve)
    ninst = len(circuits)
    assert ninst >= 1 and len(pks) == ninst
    n = pks[0]["n"]
    nw, ultra = circuits[0].nw, circuits[0].ultra
    for cs, pk in zip(circuits, pks):   # snark.rs:226-260
        if cs.n != n or pk["n"] != n:
            raise ValueError("circuit / proving key domain size differs from the expected domain size")
        if cs.nw != nw or cs.ultra != ultra or (pk["plookup"] is not None) != ultra:
            raise ValueError("inconsistent plonk circuit types")
    log_n = n.bit_length() - 1
    m = quotient_domain_size(n, nw)
    log_m = m.bit_length() - 1
    ratio = m // n
    dom = Radix2Domain(fr, n)
    qdom = Radix2Domain(fr, m)
    g = dom.group_gen
    g_inv = pow(g, -1, p)
    srs_limbs, srs_points = pks[0]["srs"]
    bl = list(blinders)
    assert len(bl) == batch_num_blinders(circuits)

    def commit(c):
        return be.commit(srs_limbs, srs_points, c)

    def mask(poly: List[int], hb: int) -> List[int]:
        # mask_polynomial (prover.rs:463-486): rand(hb) * (X^n - 1) + poly
        r = [bl.pop(0) for _ in range(hb + 1)]
        out = list(poly) + [0] * (n + hb + 1 - len(poly))
        for i, b in enumerate(r):
            out[i] = (out[i] - b) % p
            out[n + i] = (out[n + i] + b) % p
        return _strip(out)

    tr = TRANSCRIPTS[transcript](b"PlonkProof")
    if extra_msg is not None:
        tr.append_message(b"extra info", extra_msg)
    for cs, pk in zip(circuits, pks):
        append_vk_and_pub_input(tr, curve, pk["vk"], cs.public_input())
    I = [dict() for _ in range(ninst)]   # per-instance oracles

    # ---- round 1 ---------------------------------------------------------------------------------
    for cs, st in zip(circuits, I):
        st["wvals"] = cs.wire_values()
        st["wire_polys"] = [mask(_strip(be.ntt(w, log_n, True)), 1) for w in st["wvals"]]
        st["wires_comms"] = [commit(wp) for wp in st["wire_polys"]]
        pi_vec = [0] * n
        for gid in cs.pub_input_gate_ids:
            pi_vec[gid] = cs.witness[cs.wire_variables[GATE_WIDTH][gid]]
        st["pi_poly"] = _strip(be.ntt(pi_vec, log_n, True))
        for c in st["wires_comms"]:
            tr.append_message(b"witness_poly_comms", ser_g1(curve, c))

    # ---- round 1.5 (Plookup; the challenge is squeezed even without it) -------------------------------
    tau = tr.get_and_append_challenge(fr, b"tau")
    if ultra:
        for cs, st in zip(circuits, I):
            st["merged_table"] = cs.merged_lookup_table(tau)
            st["sorted_vec"] = cs.lookup_sorted_vec(tau, st["merged_table"])
            sv = st["sorted_vec"]
            st["h_polys"] = [mask(_strip(be.ntt(sv[:n], log_n, True)), 2), mask(_strip(be.ntt(sv[n - 1:], log_n, True)), 2)]
            st["h_comms"] = [commit(h) for h in st["h_polys"]]
            for c in st["h_comms"]:
                tr.append_message(b"h_poly_comms", ser_g1(curve, c))

    # ---- round 2 ---------------------------------------------------------------------------------
    beta = tr.get_and_append_challenge(fr, b"beta")
    gamma = tr.get_and_append_challenge(fr, b"gamma")
    for cs, st in zip(circuits, I):
        ext_id, wperm, wvals = cs.extended_id_permutation, cs.wire_permutation, st["wvals"]
        prod = [1]
        for j in range(n - 1):
            a = b = 1
            for i in range(nw):
                tmp = (wvals[i][j] + gamma) % p
                a = a * (tmp + beta * ext_id[i * n + j]) % p
                pi_, pj_ = wperm[i * n + j]
                b = b * (tmp + beta * ext_id[pi_ * n + pj_]) % p
            prod.append(prod[-1] * a % p * pow(b, -1, p) % p)
        st["z_poly"] = mask(_strip(be.ntt(prod, log_n, True)), 2)
        st["z_comm"] = commit(st["z_poly"])
        tr.append_message(b"perm_poly_comms", ser_g1(curve, st["z_comm"]))

    # ---- round 2.5 (Plookup product) -------------------------------------------------------------------
    if ultra:
        for cs, st in zip(circuits, I):
            st["pl_poly"] = mask(_strip(be.ntt(cs.lookup_prod_vec(tau, beta, gamma, st["merged_table"], st["sorted_vec"]), log_n, True)), 2)
            st["pl_comm"] = commit(st["pl_poly"])
            tr.append_message(b"plookup_poly_comms", ser_g1(curve, st["pl_comm"]))

    # ---- round 3 ---------------------------------------------------------------------------------
    alpha = tr.get_and_append_challenge(fr, b"alpha")
    G = fr.generator
    z_h_inv = [pow((pow(G * qdom.element(i) % p, n, p) - 1) % p, -1, p) for i in range(ratio)]
    cfft = lambda poly: be.ntt(poly, log_m, False, G)  # noqa: E731  coset.fft
    alpha2 = alpha * alpha % p
    alpha3 = alpha2 * alpha % p
    alpha7 = alpha3 * alpha3 % p * alpha % p
    alpha_step = alpha7 if ultra else alpha3
    n_f = n % p
    bp1 = (1 + beta) % p
    gb = gamma * bp1 % p
    quot = [0] * m
    wq = qdom.group_gen
    alpha_base = 1
    for cs, pk, st in zip(circuits, pks, I):
        k = pk["vk"]["k"]
        sel_c = [cfft(s_) for s_ in pk["selectors"]]
        sig_c = [cfft(s_) for s_ in pk["sigmas"]]
        w_c = [cfft(wp) for wp in st["wire_polys"]]
        z_c = cfft(st["z_poly"])
        pi_c = cfft(st["pi_poly"])
        if ultra:
            lk = pk["plookup"]
            tds_c, qds_c = cfft(lk["table_dom_sep_poly"]), cfft(lk["q_dom_sep_poly"])
            rng_c, key_c = cfft(lk["range_table_poly"]), cfft(lk["key_table_poly"])
            h1_c, h2_c, pl_c = cfft(st["h_polys"][0]), cfft(st["h_polys"][1]), cfft(st["pl_poly"])
            ql_c = sel_c[N_SELECTORS]
        x = G  # eval point g * w_m^i
        for i in range(m):
            inx = (i + ratio) % m
            w = [w_c[j][i] for j in range(nw)]
            q = [sel_c[s_][i] for s_ in range(N_SELECTORS)]
            t_circ = (q[11] + pi_c[i] + q[0] * w[0] + q[1] * w[1] + q[2] * w[2] + q[3] * w[3]
                      + q[4] * w[0] * w[1] + q[5] * w[2] * w[3] + q[12] * w[0] * w[1] * w[2] * w[3] * w[4]
                      + q[6] * pow(w[0], 5, p) + q[7] * pow(w[1], 5, p) + q[8] * pow(w[2], 5, p) + q[9] * pow(w[3], 5, p)
                      - q[10] * w[4]) % p
            zx, zxw = z_c[i], z_c[inx]
            r1 = zx
            r2 = zxw
            for j in range(nw):
                r1 = r1 * (w[j] + k[j] * x % p * beta + gamma) % p
                r2 = r2 * (w[j] + sig_c[j][i] * beta + gamma) % p
            t1 = (t_circ + alpha * (r1 - r2)) % p
            t2 = alpha2 * (zx - 1) % p * pow(n_f * (x - 1) % p, -1, p) % p
            if ultra:  # compute_quotient_plookup_contribution (prover.rs:773-888)
                lag_n = g_inv * pow(n_f * (x - g_inv) % p, -1, p) % p
                lag_1 = pow(n_f * (x - 1) % p, -1, p)
                mt_x = _merged_table(p, tau, rng_c[i], key_c[i], ql_c[i], w[3], w[4], tds_c[i])
                mt_xw = _merged_table(p, tau, rng_c[inx], key_c[inx], ql_c[inx], w_c[3][inx], w_c[4][inx], tds_c[inx])
                ml_x = _merged_lookup(p, tau, w[5], w[0], w[1], w[2], ql_c[i], qds_c[i])
                ap = alpha3
                t2 = (t2 + ap * ((h1_c[i] - h2_c[inx]) * lag_n % p)) % p
                ap = ap * alpha % p
                t2 = (t2 + ap * ((pl_c[i] - 1) * lag_1 % p)) % p
                ap = ap * alpha % p
                t2 = (t2 + ap * ((pl_c[i] - 1) * lag_n % p)) % p
                ap = ap * alpha % p
                term = (x - g_inv) * (pl_c[i] * bp1 % p * ((gamma + ml_x) % p) % p * ((gb + mt_x + beta * mt_xw) % p)
                                      - pl_c[inx] * ((gb + h1_c[i] + beta * h1_c[inx]) % p) % p * ((gb + h2_c[i] + beta * h2_c[inx]) % p)) % p
                t1 = (t1 + ap * term) % p
            quot[i] = (quot[i] + alpha_base * ((t1 * z_h_inv[i % ratio] + t2) % p)) % p
            x = x * wq % p
        alpha_base = alpha_base * alpha_step % p
    quot_poly = _strip(be.ntt(quot, log_m, True, G))
    expected_degree = nw * (n + 1) + 2
    if len(quot_poly) - 1 != expected_degree:
        raise ValueError("WrongQuotientPolyDegree(%d, %d)" % (len(quot_poly) - 1, expected_degree))
    split = []
    for i in range(nw):
        end = (i + 1) * (n + 2) if i < nw - 1 else len(quot_poly)
        split.append(_strip(quot_poly[i * (n + 2):end]))
    last = 0
    for i in range(nw - 1):
        now = bl.pop(0)
        split[i][0] = (split[i][0] - last) % p
        assert len(split[i]) == n + 2
        split[i].append(now)
        last = now
    split[-1][0] = (split[-1][0] - last) % p
    split_comms = [commit(sp) for sp in split]
    for c in split_comms:
        tr.append_message(b"quot_poly_comms", ser_g1(curve, c))

    # ---- round 4 ---------------------------------------------------------------------------------
    zeta = tr.get_and_append_challenge(fr, b"zeta")
    zg = zeta * g % p
    for pk, st in zip(pks, I):
        st["wires_evals"] = [_poly_eval(p, wp, zeta) for wp in st["wire_polys"]]
        st["sigma_evals"] = [_poly_eval(p, sg, zeta) for sg in pk["sigmas"][:nw - 1]]
        st["perm_next_eval"] = _poly_eval(p, st["z_poly"], zg)
        for e in st["wires_evals"]:
            tr.append_message(b"wire_evals", ser_fr(fr, e))
        for e in st["sigma_evals"]:
            tr.append_message(b"wire_sigma_evals", ser_fr(fr, e))
        tr.append_message(b"perm_next_eval", ser_fr(fr, st["perm_next_eval"]))
    for pk, st in zip(pks, I):
        st["plookup_evals"] = None
        if ultra:  # round 4.5: compute_plookup_evaluations (prover.rs:239-297)
            lk, h_polys, pl_poly, wire_polys = pk["plookup"], st["h_polys"], st["pl_poly"], st["wire_polys"]
            qlp = pk["selectors"][N_SELECTORS]
            st["plookup_evals"] = {
                "range_table_eval": _poly_eval(p, lk["range_table_poly"], zeta), "key_table_eval": _poly_eval(p, lk["key_table_poly"], zeta),
                "h_1_eval": _poly_eval(p, h_polys[0], zeta), "q_lookup_eval": _poly_eval(p, qlp, zeta),
                "table_dom_sep_eval": _poly_eval(p, lk["table_dom_sep_poly"], zeta), "q_dom_sep_eval": _poly_eval(p, lk["q_dom_sep_poly"], zeta),
                "prod_next_eval": _poly_eval(p, pl_poly, zg), "range_table_next_eval": _poly_eval(p, lk["range_table_poly"], zg),
                "key_table_next_eval": _poly_eval(p, lk["key_table_poly"], zg), "h_1_next_eval": _poly_eval(p, h_polys[0], zg),
                "h_2_next_eval": _poly_eval(p, h_polys[1], zg), "q_lookup_next_eval": _poly_eval(p, qlp, zg),
                "w_3_next_eval": _poly_eval(p, wire_polys[3], zg), "w_4_next_eval": _poly_eval(p, wire_polys[4], zg),
                "table_dom_sep_next_eval": _poly_eval(p, lk["table_dom_sep_poly"], zg)}
            _append_plookup_evals(tr, fr, st["plookup_evals"])

    # linearization polynomial (snark.rs:403-428; prover.rs:302-360,963-1113)
    vanish = (pow(zeta, n, p) - 1) % p
    zeta_n2 = (vanish + 1) * zeta % p * zeta % p
    r_quot = list(split[0])
    coeff = 1
    for sp in split[1:]:
        coeff = coeff * zeta_n2 % p
        r_quot = _poly_add(p, r_quot, _poly_scale(p, sp, coeff))
    lin = _poly_scale(p, r_quot, (-vanish) % p)
    lagrange_1 = vanish * pow(n_f * (zeta - 1) % p, -1, p) % p
    alpha_base = 1
    for pk, st in zip(pks, I):
        k = pk["vk"]["k"]
        we, sigma_evals, perm_next_eval = st["wires_evals"], st["sigma_evals"], st["perm_next_eval"]
        sel = pk["selectors"]
        r_circ = []
        for s_idx, sc in ((0, we[0]), (1, we[1]), (2, we[2]), (3, we[3]), (4, we[0] * we[1]), (5, we[2] * we[3]),
                          (6, pow(we[0], 5, p)), (7, pow(we[1], 5, p)), (8, pow(we[2], 5, p)), (9, pow(we[3], 5, p)),
                          (12, we[0] * we[1] * we[2] * we[3] * we[4]), (10, -we[4])):
            r_circ = _poly_add(p, r_circ, _poly_scale(p, sel[s_idx], sc % p))
        r_circ = _poly_add(p, r_circ, sel[11])
        c1 = alpha
        for j in range(nw):
            c1 = c1 * (we[j] + beta * k[j] % p * zeta + gamma) % p
        c1 = (c1 + alpha2 * lagrange_1) % p
        r_perm = _poly_scale(p, st["z_poly"], c1)
        c2 = alpha * beta % p * perm_next_eval % p
        for j in range(nw - 1):
            c2 = c2 * (we[j] + beta * sigma_evals[j] + gamma) % p
        r_perm = _poly_add(p, r_perm, _poly_scale(p, pk["sigmas"][nw - 1], (-c2) % p))
        non_quot = _poly_add(p, r_circ, r_perm)
        if ultra:  # compute_lin_poly_plookup_contribution (prover.rs:1037-1113)
            pe = st["plookup_evals"]
            alpha4 = alpha2 * alpha2 % p
            alpha5, alpha6 = alpha4 * alpha % p, alpha4 * alpha2 % p
            lagrange_n = vanish * g_inv % p * pow(n_f * (zeta - g_inv) % p, -1, p) % p
            mt = _merged_table(p, tau, pe["range_table_eval"], pe["key_table_eval"], pe["q_lookup_eval"], we[3], we[4], pe["table_dom_sep_eval"])
            mtn = _merged_table(p, tau, pe["range_table_next_eval"], pe["key_table_next_eval"], pe["q_lookup_next_eval"], pe["w_3_next_eval"],
                                pe["w_4_next_eval"], pe["table_dom_sep_next_eval"])
            ml = _merged_lookup(p, tau, we[5], we[0], we[1], we[2], pe["q_lookup_eval"], pe["q_dom_sep_eval"])
            zmg = (zeta - g_inv) % p
            cpl = (alpha4 * lagrange_1 + alpha5 * lagrange_n
                   + alpha6 * zmg % p * bp1 % p * ((gamma + ml) % p) % p * ((gb + mt + beta * mtn) % p)) % p
            r_lookup = _poly_scale(p, st["pl_poly"], cpl)
            ch2 = (-alpha6) % p * zmg % p * pe["prod_next_eval"] % p * ((gb + pe["h_1_eval"] + beta * pe["h_1_next_eval"]) % p) % p
            r_lookup = _poly_add(p, r_lookup, _poly_scale(p, st["h_polys"][1], ch2))
            non_quot = _poly_add(p, non_quot, r_lookup)
        lin = _poly_add(p, lin, _poly_scale(p, non_quot, alpha_base))
        alpha_base = alpha_base * alpha_step % p

    # ---- round 5 ---------------------------------------------------------------------------------
    v = tr.get_and_append_challenge(fr, b"v")

    def batched_witness(polys, r, point):
        acc, coeff = [], 1
        for poly in polys:
            acc = _poly_add(p, acc, _poly_scale(p, poly, coeff))
            coeff = coeff * r % p
        return commit(_div_linear(p, acc, point))

    open_polys = [lin]
    shifted_polys = []
    for pk, st in zip(pks, I):
        open_polys += st["wire_polys"] + pk["sigmas"][:nw - 1]
        shifted_polys += [st["z_poly"]]
        if ultra:  # plookup_open_polys_ref / plookup_shifted_open_polys_ref (prover.rs:427-460)
            lk, h_polys = pk["plookup"], st["h_polys"]
            open_polys += [lk["range_table_poly"], lk["key_table_poly"], h_polys[0], pk["selectors"][N_SELECTORS], lk["table_dom_sep_poly"],
                           lk["q_dom_sep_poly"]]
            shifted_polys += [st["pl_poly"], lk["range_table_poly"], lk["key_table_poly"], h_polys[0], h_polys[1], pk["selectors"][N_SELECTORS],
                              st["wire_polys"][3], st["wire_polys"][4], lk["table_dom_sep_poly"]]
    opening = batched_witness(open_polys, v, zeta)
    shifted = batched_witness(shifted_polys, v, zg)
    return {
        "wires_poly_comms_vec": [st["wires_comms"] for st in I], "prod_perm_poly_comms_vec": [st["z_comm"] for st in I],
        "poly_evals_vec": [{"wires_evals": st["wires_evals"], "wire_sigma_evals": st["sigma_evals"], "perm_next_eval": st["perm_next_eval"]}
                           for st in I],
        "plookup_proofs_vec": [({"h_poly_comms": st["h_comms"], "prod_lookup_poly_comm": st["pl_comm"], "poly_evals": st["plookup_evals"]}
                                if ultra else None) for st in I],
        "split_quot_poly_comms": split_comms, "opening_proof": opening, "shifted_opening_proof": shifted,
        "challenges": {"tau": tau, "beta": beta, "gamma": gamma, "alpha": alpha, "zeta": zeta, "v": v},
        "wire_polys_vec": [st["wire_polys"] for st in I],   # `Oracles::wire_polys` (snark.rs:466-468): not part of the proof
    }


def _ser_plookup_proof(curve, fr, lp) -> bytes:
    out = bytearray()
    if lp is None:
        return b"\x00"
    out += b"\x01"
    out += struct.pack("<Q", len(lp["h_poly_comms"]))
    for c in lp["h_poly_comms"]:
        out += ser_g1(curve, c)
    out += ser_g1(curve, lp["prod_lookup_poly_comm"])
    for name in PLOOKUP_EVAL_FIELDS:
        out += ser_fr(fr, lp["poly_evals"][name])
    return bytes(out)


def serialize_batch_proof(curve: Curve, bp: dict) -> bytes:
    """`BatchProof<E>` CanonicalSerialize (compressed), field order of structs.rs:271-292."""
    fr = curve.fr
    out = bytearray()
    out += struct.pack("<Q", len(bp["wires_poly_comms_vec"]))
    for comms in bp["wires_poly_comms_vec"]:
        out += struct.pack("<Q", len(comms))
        for c in comms:
            out += ser_g1(curve, c)
    out += struct.pack("<Q", len(bp["prod_perm_poly_comms_vec"]))
    for c in bp["prod_perm_poly_comms_vec"]:
        out += ser_g1(curve, c)
    out += struct.pack("<Q", len(bp["poly_evals_vec"]))
    for pe in bp["poly_evals_vec"]:
        out += struct.pack("<Q", len(pe["wires_evals"]))
        for e in pe["wires_evals"]:
            out += ser_fr(fr, e)
        out += struct.pack("<Q", len(pe["wire_sigma_evals"]))
        for e in pe["wire_sigma_evals"]:
            out += ser_fr(fr, e)
        out += ser_fr(fr, pe["perm_next_eval"])
    out += struct.pack("<Q", len(bp["plookup_proofs_vec"]))
    for lp in bp["plookup_proofs_vec"]:
        out += _ser_plookup_proof(curve, fr, lp)
    out += struct.pack("<Q", len(bp["split_quot_poly_comms"]))
    for c in bp["split_quot_poly_comms"]:
        out += ser_g1(curve, c)
    out += ser_g1(curve, bp["opening_proof"])
    out += ser_g1(curve, bp["shifted_opening_proof"])
    return bytes(out)


def serialize_proof(curve: Curve, proof: dict) -> bytes:
    """`Proof<E>` CanonicalSerialize (compressed), field order of structs.rs:62-84."""
    fr = curve.fr
    out = bytearray()
    out += struct.pack("<Q", len(proof["wires_poly_comms"]))
    for c in proof["wires_poly_comms"]:
        out += ser_g1(curve, c)
    out += ser_g1(curve, proof["prod_perm_poly_comm"])
    out += struct.pack("<Q", len(proof["split_quot_poly_comms"]))
    for c in proof["split_quot_poly_comms"]:
        out += ser_g1(curve, c)
    out += ser_g1(curve, proof["opening_proof"])
    out += ser_g1(curve, proof["shifted_opening_proof"])
    out += struct.pack("<Q", len(proof["wires_evals"]))
    for e in proof["wires_evals"]:
        out += ser_fr(fr, e)
    out += struct.pack("<Q", len(proof["wire_sigma_evals"]))
    for e in proof["wire_sigma_evals"]:
        out += ser_fr(fr, e)
    out += ser_fr(fr, proof["perm_next_eval"])
    lp = proof.get("plookup_proof")
    if lp is None:
        out += b"\x00"  # plookup_proof: None
    else:               # Some(PlookupProof { h_poly_comms: Vec, prod_lookup_poly_comm, poly_evals }) (structs.rs:255-265,496-541)
        out += b"\x01"
        out += struct.pack("<Q", len(lp["h_poly_comms"]))
        for c in lp["h_poly_comms"]:
            out += ser_g1(curve, c)
        out += ser_g1(curve, lp["prod_lookup_poly_comm"])
        for name in PLOOKUP_EVAL_FIELDS:
            out += ser_fr(fr, lp["poly_evals"][name])
    return bytes(out)


def verify(curve: Curve, vk: dict, pub_input: Sequence[int], proof: dict, beta_srs: int, transcript: str = "solidity",
           extra_msg: Optional[bytes] = None) -> bool:
    """`PlonkKzgSnark::verify` (one instance): `From<Proof> for BatchProof` (structs.rs:319-331), then `batch_verify`."""
    bp = {"wires_poly_comms_vec": [proof["wires_poly_comms"]], "prod_perm_poly_comms_vec": [proof["prod_perm_poly_comm"]],
          "poly_evals_vec": [{"wires_evals": proof["wires_evals"], "wire_sigma_evals": proof["wire_sigma_evals"],
                              "perm_next_eval": proof["perm_next_eval"]}],
          "plookup_proofs_vec": [proof.get("plookup_proof")], "split_quot_poly_comms": proof["split_quot_poly_comms"],
          "opening_proof": proof["opening_proof"], "shifted_opening_proof": proof["shifted_opening_proof"]}
    return batch_verify(curve, [vk], [pub_input], bp, beta_srs, transcript, extra_msg)


def batch_verify(curve: Curve, vks: Sequence[dict], pub_inputs: Sequence[Sequence[int]], bp: dict, beta_srs: int,
                 transcript: str = "solidity", extra_msg: Optional[bytes] = None) -> bool:
    """verifier.rs: prepare_pcs_info (:68-193) + batch_verify_opening_proofs (:195-254) for ONE batch proof over several instances,
    TurboPlonk or UltraPlonk; the pairing check e(A,[x]_2) == e(B,[1]_2) is evaluated as x*A == B with the known trapdoor
    x = beta_srs."""
    fr, p = curve.fr, curve.fr.p
    ninst = len(vks)
    if ninst == 0 or len(pub_inputs) != ninst or len(bp["prod_perm_poly_comms_vec"]) != ninst:
        return False
    n = vks[0]["domain_size"]
    dom = Radix2Domain(fr, n)
    g = dom.group_gen
    g_inv = pow(g, -1, p)
    for vk, pi, lp, wc in zip(vks, pub_inputs, bp["plookup_proofs_vec"], bp["wires_poly_comms_vec"]):
        if (vk.get("plookup") is not None) != (lp is not None):   # verifier.rs:97
            return False
        if vk["domain_size"] != n or len(pi) != vk["num_inputs"] or len(wc) != len(vk["k"]):
            return False
    # compute_challenges (verifier.rs:256-318)
    tr = TRANSCRIPTS[transcript](b"PlonkProof")
    if extra_msg is not None:
        tr.append_message(b"extra info", extra_msg)
    for vk, pi in zip(vks, pub_inputs):
        append_vk_and_pub_input(tr, curve, vk, pi)
    for comms in bp["wires_poly_comms_vec"]:
        for c in comms:
            tr.append_message(b"witness_poly_comms", ser_g1(curve, c))
    tau = tr.get_and_append_challenge(fr, b"tau")
    for lp in bp["plookup_proofs_vec"]:
        if lp is not None:
            for c in lp["h_poly_comms"]:
                tr.append_message(b"h_poly_comms", ser_g1(curve, c))
    beta = tr.get_and_append_challenge(fr, b"beta")
    gamma = tr.get_and_append_challenge(fr, b"gamma")
    for c in bp["prod_perm_poly_comms_vec"]:
        tr.append_message(b"perm_poly_comms", ser_g1(curve, c))
    for lp in bp["plookup_proofs_vec"]:
        if lp is not None:
            tr.append_message(b"plookup_poly_comms", ser_g1(curve, lp["prod_lookup_poly_comm"]))
    alpha = tr.get_and_append_challenge(fr, b"alpha")
    for c in bp["split_quot_poly_comms"]:
        tr.append_message(b"quot_poly_comms", ser_g1(curve, c))
    zeta = tr.get_and_append_challenge(fr, b"zeta")
    for pe in bp["poly_evals_vec"]:
        for e in pe["wires_evals"]:
            tr.append_message(b"wire_evals", ser_fr(fr, e))
        for e in pe["wire_sigma_evals"]:
            tr.append_message(b"wire_sigma_evals", ser_fr(fr, e))
        tr.append_message(b"perm_next_eval", ser_fr(fr, pe["perm_next_eval"]))
    for lp in bp["plookup_proofs_vec"]:
        if lp is not None:
            _append_plookup_evals(tr, fr, lp["poly_evals"])
    v = tr.get_and_append_challenge(fr, b"v")
    tr.append_message(b"open_proof", ser_g1(curve, bp["opening_proof"]))
    tr.append_message(b"shifted_open_proof", ser_g1(curve, bp["shifted_opening_proof"]))
    u = tr.get_and_append_challenge(fr, b"u")

    alpha2 = alpha * alpha % p
    alpha3, alpha4 = alpha2 * alpha % p, alpha2 * alpha2 % p
    alpha5, alpha6 = alpha4 * alpha % p, alpha4 * alpha2 % p
    alpha7 = alpha6 * alpha % p
    step = alpha7 if vks[0].get("plookup") is not None else alpha3   # verifier.rs:131-138
    alpha_bases = [1]
    for _ in range(ninst - 1):
        alpha_bases.append(alpha_bases[-1] * step % p)
    vanish = (pow(zeta, n, p) - 1) % p
    n_f = n % p
    lagrange_1 = vanish * pow(n_f * (zeta - 1) % p, -1, p) % p
    lagrange_n = vanish * g_inv % p * pow(n_f * (zeta - g_inv) % p, -1, p) % p
    bp1 = (1 + beta) % p
    gb = gamma * bp1 % p

    lin_const = 0
    sb: List[Tuple[int, object]] = []
    for vk, pub_input, pe, lp, z_comm, ab in zip(vks, pub_inputs, bp["poly_evals_vec"], bp["plookup_proofs_vec"],
                                                  bp["prod_perm_poly_comms_vec"], alpha_bases):
        k = vk["k"]
        nw = len(k)
        we, se, pne = pe["wires_evals"], pe["wire_sigma_evals"], pe["perm_next_eval"]
        # evaluate_pi_poly (verifier.rs:845-881)
        pi_eval = 0
        if vanish:
            vdn = pow(n_f, -1, p) * vanish % p
            for i, val in enumerate(pub_input):
                e = dom.element(i)
                pi_eval = (pi_eval + vdn * e % p * pow((zeta - e) % p, -1, p) % p * val) % p
        # compute_lin_poly_constant_term (verifier.rs:340-418)
        tmp = (pi_eval - alpha2 * lagrange_1) % p
        acc = alpha * pne % p * ((gamma + we[nw - 1]) % p) % p
        for j in range(nw - 1):
            acc = acc * ((gamma + we[j] + beta * se[j]) % p) % p
        tmp = (tmp - acc) % p
        if lp is not None:
            ev_ = lp["poly_evals"]
            pc = (lagrange_n * ((ev_["h_1_eval"] - ev_["h_2_next_eval"] - alpha2) % p) - alpha * lagrange_1
                  - alpha3 * ((zeta - g_inv) % p) % p * ev_["prod_next_eval"] % p
                  * ((gb + ev_["h_1_eval"] + beta * ev_["h_1_next_eval"]) % p) % p * ((gb + beta * ev_["h_2_next_eval"]) % p)) % p
            tmp = (tmp + alpha3 * pc) % p
        lin_const = (lin_const + ab * tmp) % p
        # linearization_scalars_and_bases (verifier.rs:513-656)
        coeff = alpha2 * lagrange_1 % p
        c = alpha
        for j in range(nw):
            c = c * ((beta * k[j] % p * zeta + gamma + we[j]) % p) % p
        sb.append(((coeff + c) * ab % p, z_comm))
        c = alpha * beta % p * pne % p
        for j in range(nw - 1):
            c = c * ((beta * se[j] + gamma + we[j]) % p) % p
        sb.append(((-c * ab) % p, vk["sigma_comms"][nw - 1]))
        qs = [we[0], we[1], we[2], we[3], we[0] * we[1] % p, we[2] * we[3] % p, pow(we[0], 5, p), pow(we[1], 5, p),
              pow(we[2], 5, p), pow(we[3], 5, p), (-we[4]) % p, 1, we[0] * we[1] * we[2] * we[3] * we[4] % p]
        for s_, cm in zip(qs, vk["selector_comms"]):   # 13 scalars: the q_lookup commitment (14th) is not part of [D]
            sb.append((s_ * ab % p, cm))
        if lp is not None:
            ev_ = lp["poly_evals"]
            ml = _merged_lookup(p, tau, we[5], we[0], we[1], we[2], ev_["q_lookup_eval"], ev_["q_dom_sep_eval"])
            mt = _merged_table(p, tau, ev_["range_table_eval"], ev_["key_table_eval"], ev_["q_lookup_eval"], we[3], we[4], ev_["table_dom_sep_eval"])
            mtn = _merged_table(p, tau, ev_["range_table_next_eval"], ev_["key_table_next_eval"], ev_["q_lookup_next_eval"], ev_["w_3_next_eval"],
                                ev_["w_4_next_eval"], ev_["table_dom_sep_next_eval"])
            cpl = (alpha4 * lagrange_1 + alpha5 * lagrange_n
                   + alpha6 * ((zeta - g_inv) % p) % p * bp1 % p * ((gamma + ml) % p) % p * ((gb + mt + beta * mtn) % p)) % p
            sb.append((cpl * ab % p, lp["prod_lookup_poly_comm"]))
            ch2 = alpha6 * ((g_inv - zeta) % p) % p * ev_["prod_next_eval"] % p * ((gb + ev_["h_1_eval"] + beta * ev_["h_1_next_eval"]) % p) % p
            sb.append((ch2 * ab % p, lp["h_poly_comms"][1]))
    zeta_n2 = (1 + vanish) * zeta % p * zeta % p
    coeff = (-vanish) % p
    sb.append((coeff, bp["split_quot_poly_comms"][0]))
    for cm in bp["split_quot_poly_comms"][1:]:
        coeff = coeff * zeta_n2 % p
        sb.append((coeff, cm))
    # aggregate_poly_commitments / aggregate_evaluations (verifier.rs:421-511,673-745): the powers of v run on across the instances
    v_base, uv_base = v, u
    buf, evals = [], []
    for vk, wires, pe, lp, z_comm in zip(vks, bp["wires_poly_comms_vec"], bp["poly_evals_vec"], bp["plookup_proofs_vec"],
                                         bp["prod_perm_poly_comms_vec"]):
        nw = len(wires)
        for cm in wires:
            buf.append(v_base); sb.append((v_base, cm)); v_base = v_base * v % p
        for cm in vk["sigma_comms"][:nw - 1]:
            buf.append(v_base); sb.append((v_base, cm)); v_base = v_base * v % p
        buf.append(uv_base); sb.append((uv_base, z_comm)); uv_base = uv_base * v % p
        evals += list(pe["wires_evals"]) + list(pe["wire_sigma_evals"]) + [pe["perm_next_eval"]]
        if lp is not None:
            lvk = vk["plookup"]
            for cm in (lvk["range_table_comm"], lvk["key_table_comm"], lp["h_poly_comms"][0], vk["selector_comms"][N_SELECTORS],
                       lvk["table_dom_sep_comm"], lvk["q_dom_sep_comm"]):
                buf.append(v_base); sb.append((v_base, cm)); v_base = v_base * v % p
            for cm in (lp["prod_lookup_poly_comm"], lvk["range_table_comm"], lvk["key_table_comm"], lp["h_poly_comms"][0], lp["h_poly_comms"][1],
                       vk["selector_comms"][N_SELECTORS], wires[3], wires[4], lvk["table_dom_sep_comm"]):
                buf.append(uv_base); sb.append((uv_base, cm)); uv_base = uv_base * v % p
            evals += _plookup_evals_vec(lp["poly_evals"]) + _plookup_next_evals_vec(lp["poly_evals"])
    ev = (-lin_const) % p
    assert len(buf) == len(evals)
    for b_, e in zip(buf, evals):
        ev = (ev + e * b_) % p
    # batch_verify_opening_proofs with one batch proof (r = 1)
    A = curve.add(bp["opening_proof"], curve.mul(u, bp["shifted_opening_proof"]))
    B = None
    for s_, cm in sb:
        B = curve.add(B, curve.mul(s_ % p, cm))
    B = curve.add(B, curve.mul(zeta, bp["opening_proof"]))
    B = curve.add(B, curve.mul(u * (zeta * g % p) % p, bp["shifted_opening_proof"]))
    B = curve.add(B, curve.mul((-ev) % p, curve.gen))
    return curve.mul(beta_srs % p, A) == B
