"""Host-side mirror of ark-poly's `Radix2EvaluationDomain` for the calls the prover makes
(plonk/src/proof_system/prover.rs:54-62,545-567,672; relation/src/constraint_system.rs:1172-1257):
`new`, `get_coset`, `fft` / `fft_in_place`, `ifft` / `ifft_in_place`, `element`, `size`.
Vectors are (len, 4) uint64 arrays of Montgomery limbs; the transforms run on the GPU."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .context import Context
from .errors import DomainCreationError, InvalidParameters
from .fields import GENERATOR, MODULUS, TWO_ADICITY, int_to_limbs, to_mont


class Radix2EvaluationDomain:
    def __init__(self, ctx: Context, field: str, num_coeffs: int, offset: Optional[int] = None):
        """`Radix2EvaluationDomain::new(num_coeffs)`: size = next power of two; raises
        DomainCreationError where arkworks returns None (log2 size > TWO_ADICITY)."""
        if field not in TWO_ADICITY:
            raise InvalidParameters("no FFT domain over %s" % field)
        size, log = 1, 0
        while size < num_coeffs:
            size <<= 1
            log += 1
        if log > TWO_ADICITY[field]:
            raise DomainCreationError("domain of size 2^%d exceeds the two-adicity of %s" % (log, field))
        self.ctx, self.field, self.size, self.log_size = ctx, field, size, log
        p = MODULUS[field]
        self.offset = 1 if offset is None else offset % p
        root = pow(GENERATOR[field], (p - 1) >> TWO_ADICITY[field], p)
        self.group_gen = pow(root, 1 << (TWO_ADICITY[field] - log), p)
        self._off_limbs = None if self.offset == 1 else int_to_limbs(to_mont(field, self.offset), 4)

    def get_coset(self, offset: int) -> "Radix2EvaluationDomain":
        """`domain.get_coset(offset)`; the prover uses offset = Fr::GENERATOR (prover.rs:545)."""
        return Radix2EvaluationDomain(self.ctx, self.field, self.size, offset)

    def element(self, i: int) -> int:
        p = MODULUS[self.field]
        return self.offset * pow(self.group_gen, i, p) % p

    def _padded(self, v: np.ndarray):
        v = np.asarray(v, dtype=np.uint64).reshape(-1, 4)
        if len(v) > self.size:
            # arkworks folds longer inputs; the prover never does that: refuse loudly
            raise InvalidParameters("input of length %d exceeds the domain size %d" % (len(v), self.size))
        buf = np.zeros((self.size, 4), dtype=np.uint64)
        buf[: len(v)] = v
        return buf, len(v)

    def fft(self, coeffs: np.ndarray) -> np.ndarray:
        """out[i] = p(offset * g^i), natural order (`domain.fft(&coeffs)`)."""
        buf, n_in = self._padded(coeffs)
        return self.ctx.ntt(self.field, buf, self.log_size, False, self._off_limbs, in_len=n_in)

    def ifft(self, evals: np.ndarray) -> np.ndarray:
        buf, n_in = self._padded(evals)
        return self.ctx.ntt(self.field, buf, self.log_size, True, self._off_limbs, in_len=n_in)

    def fft_in_place(self, v: np.ndarray) -> np.ndarray:
        """v must already have `size` rows (arkworks resizes; numpy arrays cannot grow in place)."""
        return self.ctx.ntt(self.field, v, self.log_size, False, self._off_limbs)

    def ifft_in_place(self, v: np.ndarray) -> np.ndarray:
        return self.ctx.ntt(self.field, v, self.log_size, True, self._off_limbs)

    def batch_fft(self, vs: np.ndarray, inverse: bool = False, in_len: Optional[int] = None) -> np.ndarray:
        """The `par_iter` over polynomials of prover.rs:552-562 as one call: vs is (batch, size, 4)."""
        return self.ctx.ntt(self.field, vs, self.log_size, inverse, self._off_limbs, in_len=in_len)
