"""Error types mirroring the reference's (primitives/src/pcs/errors.rs:17-34,
plonk/src/errors.rs:16-49) so that callers and tests read like the reference's."""


class PCSError(Exception):
    """Base of `jf_primitives::pcs::errors::PCSError`."""


class InvalidParameters(PCSError):
    """`PCSError::InvalidParameters(String)`"""


class UpstreamError(PCSError):
    """`PCSError::UpstreamError` -- here: a CUDA / NCCL failure surfaced through the C ABI."""


class PlonkError(Exception):
    """Base of `mpc_plonk::errors::PlonkError`."""


class DomainCreationError(PlonkError):
    """`PlonkError::DomainCreationError` -- domain size exceeds the field's two-adicity."""


class WrongQuotientPolyDegree(PlonkError):
    """`SnarkError::WrongQuotientPolyDegree` (prover.rs:916-919): the witness does not satisfy the circuit."""


def raise_for_status(rc: int, msg: str):
    from . import _ffi
    if rc == _ffi.JF_OK:
        return
    if rc in (_ffi.JF_ERR_INVALID_ARG, _ffi.JF_ERR_SCALAR_RANGE):
        raise InvalidParameters(msg)
    if rc == _ffi.JF_ERR_DOMAIN_TOO_LARGE:
        raise DomainCreationError(msg)
    if rc == _ffi.JF_ERR_QUOTIENT_DEGREE:
        raise WrongQuotientPolyDegree(msg)
    raise UpstreamError("%s (status %d)" % (msg, rc))
