"""mpc-jellyfish_b200: B200-native MSM / NTT core of the mpc-jellyfish PLONK prover.

The product is `libjf_b200.so` (hand-written sm_100a CUDA behind the C ABI of
include/jf_b200.h).  This package is the Python host side: a ctypes binding plus mirrors of
the reference's `UnivariateKzgPCS` and `Radix2EvaluationDomain` call surfaces.  Nothing in
here computes on the CPU; without the CUDA library every entry point raises.
"""
from . import _ffi
from .context import CommitKey, Context
from .domain import Radix2EvaluationDomain
from .errors import (DomainCreationError, InvalidParameters, PCSError, PlonkError, UpstreamError,
                     WrongQuotientPolyDegree)
from .multiprover import (AuthenticatedDensePoly, AuthenticatedPointShare, MpcLinkingHint, MultiproverKZG, MultiproverLinking,
                          fft_with_domain, ifft_with_domain)
from .plonk import (BatchProof, GroupLayout, LinkingHint, LinkingProof, PlonkKzgSnark, Proof, ProvingKey, Transcript,
                    keccak256)
from .sharded import Comm, Group, GroupKey, ShardedMsm, combine_partials, poly_owner, shard_range
from .pcs import Commitment, DensePolynomial, UnivariateKzgPCS, UnivariateProverParam

__all__ = [
    "Context", "CommitKey", "Radix2EvaluationDomain", "UnivariateKzgPCS", "UnivariateProverParam",
    "DensePolynomial", "Commitment", "PCSError", "InvalidParameters", "UpstreamError", "PlonkError",
    "DomainCreationError", "WrongQuotientPolyDegree", "PlonkKzgSnark", "Proof", "BatchProof", "GroupLayout", "LinkingHint", "LinkingProof", "ProvingKey", "Transcript", "keccak256",
    "MultiproverKZG", "MultiproverLinking", "MpcLinkingHint", "AuthenticatedDensePoly", "AuthenticatedPointShare", "fft_with_domain", "ifft_with_domain",
    "ShardedMsm", "Comm", "Group", "GroupKey", "combine_partials", "poly_owner", "shard_range",
]
