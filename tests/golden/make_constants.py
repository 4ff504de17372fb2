#!/usr/bin/env python3
"""Writes tests/golden/constants.json: public constants and known-answer values that pin the
oracle.  The reference holds no golden vectors for this path; these are values
every implementation of BN254 / BLS12-381 shares.  The *expected* numbers below are typed in
from the public specifications (EIP-196/197 for BN254, the BLS12-381 spec / zkcrypto for
BLS12-381, arkworks' `FrConfig`/`FqConfig` constants); tests/test_oracle.py recomputes them
with oracle/pyref.py and with the C oracle and compares.

    python tests/golden/make_constants.py
"""
import json
import os

GOLDEN = {
    "bn254_fr": {
        "modulus": "21888242871839275222246405745257275088548364400416034343698204186575808495617",
        "generator": 5, "two_adicity": 28,
        "two_adic_root_of_unity": "19103219067921713944291392827692070036145651957329286315305642004821462161904",
        "inv64": "0xc2e1f593efffffff",
        "R": "0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb",
        "R2": "0x0216d0b17f4e44a58c49833d53bb808553fe3ab1e35c59e31bb8e645ae216da7",
    },
    "bn254_fq": {
        "modulus": "21888242871839275222246405745257275088696311157297823662689037894645226208583",
        "inv64": "0x87d20782e4866389",
        "R": "0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d",
        "R2": "0x06d89f71cab8351f47ab1eff0a417ff6b5e71911d44501fbf32cfc5b538afa89",
    },
    "bls12_381_fr": {
        "modulus": "0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001",
        "generator": 7, "two_adicity": 32,
        "two_adic_root_of_unity": "10238227357739495823651030575849232062558860180284477541189508159991286009131",
        "inv64": "0xfffffffeffffffff",
        "R": "0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe",
        "R2": "0x0748d9d99f59ff1105d314967254398f2b6cedcb87925c23c999e990f3f29c6d",
    },
    "bls12_381_fq": {
        "modulus": "0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab",
        "inv64": "0x89f3fffcfffcfffd",
        "R": "0x15f65ec3fa80e4935c071a97a256ec6d77ce5853705257455f48985753c758baebf4000bc40c0002760900000002fffd",
        "R2": "0x11988fe592cae3aa9a793e85b519952d67eb88a9939d83c08de5476c4c95b6d50a76e6a609d104f1f4df1f341c341746",
    },
    "bn254_g1": {
        "b": 3, "generator": ["1", "2"],
        # EIP-196 ecMul test vector: 2 * (1, 2)
        "two_g": ["1368015179489954701390400359078579693043519447331113978918064868415326638035",
                  "9918110051302171585080402603319702774565515993150576347155970296011118125764"],
        # ark-serialize compressed encoding of the generator: x little-endian, y = 2 is the smaller root
        "generator_compressed": "0100000000000000000000000000000000000000000000000000000000000000",
        "identity_compressed": "0000000000000000000000000000000000000000000000000000000000000040",
    },
    "bls12_381_g1": {
        "b": 4,
        "generator": [
            "0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb",
            "0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1"],
        # ZCash / IETF compressed encoding (what ark-bls12-381 0.4.0 emits for G1): the published generator, the public keys
        # of the secret keys 2 and 3 in the BLS signature test suites (= 2 G, 3 G), and the identity
        "generator_compressed": "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb",
        "two_g_compressed": "a572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e",
        "three_g_compressed": "89ece308f9d1f0131765212deca99697b112d61f9be9a5f1f3780a51335b3ff981747a0b2ca2179b96d2c0c9024e5224",
        "identity_compressed": "c0" + "00" * 47,
    },
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "constants.json")
    with open(out, "w") as f:
        json.dump(GOLDEN, f, indent=1, sort_keys=True)
    print("wrote", out)
