"""Host-side constants and limb helpers (pure Python ints; no computation on the hot path)."""
from __future__ import annotations

import numpy as np

MODULUS = {
    "bn254_fr": 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    "bn254_fq": 21888242871839275222246405745257275088696311157297823662689037894645226208583,
    "bls12_381_fr": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    "bls12_381_fq": 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
}
LIMBS = {"bn254_fr": 4, "bn254_fq": 4, "bls12_381_fr": 4, "bls12_381_fq": 6}
GENERATOR = {"bn254_fr": 5, "bls12_381_fr": 7}      # `Fr::GENERATOR`, the coset offset of prover.rs:545
TWO_ADICITY = {"bn254_fr": 28, "bls12_381_fr": 32}


def mont_r(field: str) -> int:
    return pow(2, 64 * LIMBS[field], MODULUS[field])


def to_mont(field: str, x: int) -> int:
    return x * mont_r(field) % MODULUS[field]


def from_mont(field: str, x: int) -> int:
    return x * pow(mont_r(field), -1, MODULUS[field]) % MODULUS[field]


def int_to_limbs(x: int, limbs: int) -> np.ndarray:
    return np.array([(x >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(limbs)], dtype=np.uint64)


def limbs_to_int(a) -> int:
    return sum(int(v) << (64 * k) for k, v in enumerate(np.asarray(a, dtype=np.uint64).reshape(-1)))


def ints_to_array(vals, limbs: int) -> np.ndarray:
    out = np.zeros((len(vals), limbs), dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(limbs):
            out[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def array_to_ints(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64)
    flat = arr.reshape(-1, arr.shape[-1])
    return [sum(int(flat[i, k]) << (64 * k) for k in range(flat.shape[1])) for i in range(flat.shape[0])]
