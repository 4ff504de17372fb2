//! Collaborative prover (feature `ark-mpc`): `CurvePoint::msm_authenticated` on the GPU
//! (plonk/src/multiprover/primitives/multiprover_kzg.rs:128-143,171-197).
//!
//! An `AuthenticatedScalarResult` is (share, mac, public_modifier); `msm_authenticated` in ark-mpc allocates ONE gate whose
//! inputs are the 3 n component results and whose outputs are the three MSMs over the public points.  This module allocates the
//! same gate, but its body hands the three scalar vectors to `try_commit_authenticated` (one `jf_msm_batch` over the resident
//! key) instead of running three CPU MSMs -- the values, and therefore everything the fabric opens or MAC-checks afterwards,
//! are identical.
//!
//! ark-mpc is an UNPINNED git dependency of the reference (Cargo.toml:16, no Cargo.lock), so the gate-allocation calls below
//! (`MpcFabric::new_batch_gate_op`, `ResultValue::{Scalar, Point}`, the public fields of the authenticated result types) are
//! written against its public API as of the revision the reference's tests were written for; NOT COMPILED here (no Rust
//! toolchain in the build image).  If a later revision renames them, only this file changes.
use crate::try_commit_authenticated;
use ark_ec::{pairing::Pairing, CurveGroup};
use ark_mpc::{
    algebra::{AuthenticatedPointResult, AuthenticatedScalarResult, CurvePoint, CurvePointResult, MpcPointResult, Scalar},
    ResultValue,
};

/// `None`: the pairing is not implemented on the GPU (the caller keeps `CurvePoint::msm_authenticated`).
pub fn msm_authenticated<E: Pairing>(scalars: &[AuthenticatedScalarResult<E::G1>], powers_of_g: &[E::G1Affine])
                                     -> Option<AuthenticatedPointResult<E::G1>> {
    // probe once with an empty batch: is this curve implemented?
    crate::try_commit_authenticated::<E>(powers_of_g, &[])?.ok()?;
    let n = scalars.len().min(powers_of_g.len());
    if n == 0 { return None; }
    let fabric = scalars[0].fabric().clone();
    // inputs in ark-mpc's own order: all shares, all macs, all modifiers
    let mut ids = Vec::with_capacity(3 * n);
    ids.extend(scalars[..n].iter().map(|s| s.share.id()));
    ids.extend(scalars[..n].iter().map(|s| s.mac.id()));
    ids.extend(scalars[..n].iter().map(|s| s.public_modifier.id()));
    // the key slice is 'static for the life of the `ProvingKey`; the gate may run on an executor thread
    let pts: &'static [E::G1Affine] = unsafe { core::slice::from_raw_parts(powers_of_g.as_ptr(), n) };
    let outs: Vec<CurvePointResult<E::G1>> = fabric.new_batch_gate_op(ids, 3 /* arity */, move |args| {
        let vals: Vec<E::ScalarField> = args.into_iter().map(|a| Scalar::<E::G1>::from(a).inner()).collect();
        let comps = [&vals[..n], &vals[n..2 * n], &vals[2 * n..]];
        let points = try_commit_authenticated::<E>(pts, &comps)
            .expect("curve checked above")
            .expect("B200 MSM failed");   // no CPU fallback for an implemented curve
        points.into_iter().map(|p| ResultValue::Point(CurvePoint::from(p.into()))).collect()
    });
    Some(AuthenticatedPointResult {
        share: MpcPointResult::new_shared(outs[0].clone()),
        mac: MpcPointResult::new_shared(outs[1].clone()),
        public_modifier: outs[2].clone(),
    })
}
