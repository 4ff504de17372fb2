//! Safe, arkworks-typed wrappers over `jf-b200-sys`, generic over the pairing: `ark_bn254::Bn254` and
//! `ark_bls12_381::Bls12_381` implement [`GpuCurve`]; the entry points the patched reference calls
//! (`rust/patches/*.diff`) dispatch on the pairing type at run time so that the reference's
//! `impl<E: Pairing> PolynomialCommitmentScheme for UnivariateKzgPCS<E>` keeps its signature.
//! NOT COMPILED in the build image (no Rust toolchain there): `tests/test_abi.py` keeps the `-sys` crate in step with
//! the header, and the Python binding exercises the same symbols.
//!
//! Layout facts this file relies on (ark-ff 0.4 / ark-ec 0.4):
//!   * `Fp<MontBackend<_, N>, N>` is a newtype over `BigInt<N>([u64; N])` in Montgomery form, so `&[Fr]` can be passed as
//!     `*const u64` (4 limbs per element) without conversion;
//!   * `short_weierstrass::Affine { x, y, infinity: bool }` is passed with its real stride and the byte offset of
//!     `infinity`; the library repacks once in `jf_srs_load`.
use ark_ec::{pairing::Pairing, AffineRepr};
use ark_ff::{BigInt, Zero};
use core::ffi::c_int;
use jf_b200_sys as sys;
use std::{
    any::{Any, TypeId},
    collections::HashMap,
    ffi::CStr,
    marker::PhantomData,
    mem::MaybeUninit,
    ptr,
    sync::{Arc, Mutex, OnceLock},
};

#[cfg(feature = "ark-mpc")]
pub mod mpc;

#[derive(Debug)]
pub enum GpuError { InvalidParameters(String), Upstream(String), DomainCreation, WrongQuotientPolyDegree }

/// A pairing whose G1 / Fr the library implements.
pub trait GpuCurve: Pairing {
    const CURVE: c_int;   // JF_BN254 / JF_BLS12_381
    const FR: c_int;      // JF_BN254_FR / JF_BLS12_381_FR
    const L: usize;       // u64 limbs per base-field element (4 / 6)
    fn point_from_mont(xy: &[u64], infinity: bool) -> Self::G1Affine;
    /// x || y Montgomery limbs (2 L of the 12 words used) and the infinity flag: the form commitments cross the ABI in
    fn point_to_mont(p: &Self::G1Affine) -> ([u64; 12], bool);
    fn fr_from_mont(limbs: [u64; 4]) -> Self::ScalarField;
    /// (size_of::<G1Affine>(), byte offset of `infinity` inside it)
    fn affine_layout() -> (usize, usize);
}

macro_rules! impl_gpu_curve {
    ($pairing:ty, $fq:ty, $fr:ty, $aff:ty, $curve:expr, $frid:expr, $l:expr) => {
        impl GpuCurve for $pairing {
            const CURVE: c_int = $curve;
            const FR: c_int = $frid;
            const L: usize = $l;
            fn point_from_mont(xy: &[u64], infinity: bool) -> $aff {
                if infinity { return <$aff>::identity(); }
                let (mut x, mut y) = ([0u64; $l], [0u64; $l]);
                x.copy_from_slice(&xy[..$l]);
                y.copy_from_slice(&xy[$l..2 * $l]);
                <$aff>::new_unchecked(<$fq>::new_unchecked(BigInt::<$l>(x)), <$fq>::new_unchecked(BigInt::<$l>(y)))
            }
            fn point_to_mont(p: &$aff) -> ([u64; 12], bool) {
                let mut xy = [0u64; 12];
                if !p.infinity {
                    xy[..$l].copy_from_slice(&(p.x.0).0);      // Fp(BigInt<N>): the Montgomery representation as stored
                    xy[$l..2 * $l].copy_from_slice(&(p.y.0).0);
                }
                (xy, p.infinity)
            }
            fn fr_from_mont(limbs: [u64; 4]) -> $fr { <$fr>::new_unchecked(BigInt::<4>(limbs)) }
            fn affine_layout() -> (usize, usize) {
                let probe = <$aff>::identity();
                (core::mem::size_of::<$aff>(), (&probe.infinity as *const bool as usize) - (&probe as *const $aff as usize))
            }
        }
    };
}
impl_gpu_curve!(ark_bn254::Bn254, ark_bn254::Fq, ark_bn254::Fr, ark_bn254::G1Affine, sys::JF_BN254, sys::JF_BN254_FR, 4);
impl_gpu_curve!(ark_bls12_381::Bls12_381, ark_bls12_381::Fq, ark_bls12_381::Fr, ark_bls12_381::G1Affine, sys::JF_BLS12_381,
                sys::JF_BLS12_381_FR, 6);

fn map_status(rc: c_int, msg: String) -> GpuError {
    match rc {
        sys::JF_ERR_INVALID_ARG | sys::JF_ERR_SCALAR_RANGE => GpuError::InvalidParameters(msg), // PCSError::InvalidParameters
        sys::JF_ERR_DOMAIN_TOO_LARGE => GpuError::DomainCreation,                              // PlonkError::DomainCreationError
        sys::JF_ERR_QUOTIENT_DEGREE => GpuError::WrongQuotientPolyDegree,                      // SnarkError::WrongQuotientPolyDegree
        _ => GpuError::Upstream(msg),                                                          // PCSError::UpstreamError (CUDA, NCCL)
    }
}

// ---- one GPU ---------------------------------------------------------------------------------------------------------
pub struct Gpu { ctx: *mut sys::jf_ctx }
unsafe impl Send for Gpu {}
unsafe impl Sync for Gpu {}   // every entry point locks the context

impl Gpu {
    pub fn new(device: i32) -> Result<Self, GpuError> {
        let mut ctx = ptr::null_mut();
        match unsafe { sys::jf_ctx_create(device, &mut ctx) } {
            sys::JF_OK => Ok(Self { ctx }),
            _ => Err(GpuError::Upstream("no usable CUDA device (the library has no CPU fallback)".into())),
        }
    }
    fn check(&self, rc: c_int) -> Result<(), GpuError> {
        if rc == sys::JF_OK { return Ok(()); }
        Err(map_status(rc, unsafe { CStr::from_ptr(sys::jf_last_error(self.ctx)) }.to_string_lossy().into_owned()))
    }
}
impl Drop for Gpu { fn drop(&mut self) { unsafe { sys::jf_ctx_destroy(self.ctx) } } }

/// The process-wide device the patched call sites use (`JF_B200_DEVICE`, default 0).
pub fn global() -> Result<&'static Gpu, GpuError> {
    static GPU: OnceLock<Result<Gpu, String>> = OnceLock::new();
    GPU.get_or_init(|| {
        let dev = std::env::var("JF_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        Gpu::new(dev).map_err(|e| format!("{e:?}"))
    }).as_ref().map_err(|e| GpuError::Upstream(e.clone()))
}

/// `UnivariateProverParam::powers_of_g`, resident on the GPU with its window tables.
pub struct GpuCommitKey<E: GpuCurve> { gpu: &'static Gpu, srs: *mut sys::jf_srs, len: usize, _e: PhantomData<E> }
unsafe impl<E: GpuCurve> Send for GpuCommitKey<E> {}
unsafe impl<E: GpuCurve> Sync for GpuCommitKey<E> {}

fn low_zeros<F: Zero>(c: &[F]) -> usize { c.iter().take_while(|x| x.is_zero()).count() }   // mod.rs:382-385

impl<E: GpuCurve> GpuCommitKey<E> {
    pub fn load(gpu: &'static Gpu, powers_of_g: &[E::G1Affine]) -> Result<Self, GpuError> {
        let mut srs = ptr::null_mut();
        let (stride, inf_off) = E::affine_layout();
        gpu.check(unsafe {
            sys::jf_srs_load(gpu.ctx, E::CURVE, powers_of_g.as_ptr() as *const _, powers_of_g.len(), stride, inf_off as _, 0, 1, &mut srs)
        })?;
        Ok(Self { gpu, srs, len: powers_of_g.len(), _e: PhantomData })
    }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }

    /// `UnivariateKzgPCS::commit` (mod.rs:90-116): the degree check stays with the caller's `PCSError`; the Montgomery ->
    /// canonical conversion (`convert_to_bigints`, :390-395) happens on the GPU.
    pub fn commit(&self, coeffs: &[E::ScalarField]) -> Result<E::G1Affine, GpuError> {
        let nz = low_zeros(coeffs);
        let (mut xy, mut inf) = ([0u64; 12], 0);
        self.gpu.check(unsafe {
            sys::jf_msm(self.gpu.ctx, self.srs, nz, coeffs[nz..].as_ptr() as *const u64, coeffs.len() - nz, 1, xy.as_mut_ptr(), &mut inf)
        })?;
        Ok(E::point_from_mont(&xy, inf != 0))
    }
    /// `batch_commit` (mod.rs:119-131): the rayon `par_iter` becomes one call (uploads overlap kernels, groups of MSMs share
    /// one bucket reduction).
    pub fn batch_commit(&self, polys: &[&[E::ScalarField]]) -> Result<Vec<E::G1Affine>, GpuError> {
        let offs: Vec<usize> = polys.iter().map(|p| low_zeros(p)).collect();
        let ptrs: Vec<*const u64> = polys.iter().zip(&offs).map(|(p, &o)| p[o..].as_ptr() as *const u64).collect();
        let lens: Vec<usize> = polys.iter().zip(&offs).map(|(p, &o)| p.len() - o).collect();
        let w = 2 * E::L;
        let mut xy = vec![0u64; w * polys.len()];
        let mut inf = vec![0 as c_int; polys.len()];
        self.gpu.check(unsafe {
            sys::jf_msm_batch(self.gpu.ctx, self.srs, ptrs.as_ptr(), lens.as_ptr(), offs.as_ptr(), polys.len(), 1, xy.as_mut_ptr(),
                              inf.as_mut_ptr())
        })?;
        Ok((0..polys.len()).map(|i| E::point_from_mont(&xy[w * i..w * (i + 1)], inf[i] != 0)).collect())
    }
    /// `open` (mod.rs:135-161) / `batch_open` (:165-194): witness polynomials, their commitments and the evaluations on the
    /// GPU, all (polynomial, point) pairs in one call.
    pub fn batch_open(&self, polys: &[&[E::ScalarField]], points: &[E::ScalarField])
                      -> Result<(Vec<E::G1Affine>, Vec<E::ScalarField>), GpuError> {
        if polys.len() != points.len() {
            return Err(GpuError::InvalidParameters(format!("poly length {} is different from points length {}", polys.len(), points.len())));
        }
        let ptrs: Vec<*const u64> = polys.iter().map(|p| p.as_ptr() as *const u64).collect();
        let lens: Vec<usize> = polys.iter().map(|p| p.len()).collect();
        let w = 2 * E::L;
        let (mut xy, mut inf, mut ev) = (vec![0u64; w * polys.len()], vec![0 as c_int; polys.len()], vec![0u64; 4 * polys.len()]);
        self.gpu.check(unsafe {
            sys::jf_kzg_open(self.gpu.ctx, self.srs, ptrs.as_ptr(), lens.as_ptr(), polys.len(), points.as_ptr() as *const u64,
                             xy.as_mut_ptr(), inf.as_mut_ptr(), ev.as_mut_ptr())
        })?;
        Ok(((0..polys.len()).map(|i| E::point_from_mont(&xy[w * i..w * (i + 1)], inf[i] != 0)).collect(),
            (0..polys.len()).map(|i| E::fr_from_mont([ev[4 * i], ev[4 * i + 1], ev[4 * i + 2], ev[4 * i + 3]])).collect()))
    }
    pub fn open(&self, coeffs: &[E::ScalarField], z: &E::ScalarField) -> Result<(E::G1Affine, E::ScalarField), GpuError> {
        let (mut p, mut e) = self.batch_open(&[coeffs], core::slice::from_ref(z))?;
        Ok((p.remove(0), e.remove(0)))
    }

    /// `MultiproverKZG::commit` on authenticated shares (multiprover_kzg.rs:128-143): `msm_authenticated` is linear, so the
    /// share vector, the MAC vector and -- in ark-mpc revisions that carry it -- the public-modifier vector are three MSMs
    /// over the same resident key.  Returns one point per component, in the order given.
    pub fn commit_authenticated(&self, components: &[&[E::ScalarField]]) -> Result<Vec<E::G1Affine>, GpuError> {
        // no low-zero skipping here: the components of a share are random-looking and must stay aligned
        let ptrs: Vec<*const u64> = components.iter().map(|p| p.as_ptr() as *const u64).collect();
        let lens: Vec<usize> = components.iter().map(|p| p.len()).collect();
        let w = 2 * E::L;
        let mut xy = vec![0u64; w * components.len()];
        let mut inf = vec![0 as c_int; components.len()];
        self.gpu.check(unsafe {
            sys::jf_msm_batch(self.gpu.ctx, self.srs, ptrs.as_ptr(), lens.as_ptr(), ptr::null(), components.len(), 1, xy.as_mut_ptr(),
                              inf.as_mut_ptr())
        })?;
        Ok((0..components.len()).map(|i| E::point_from_mont(&xy[w * i..w * (i + 1)], inf[i] != 0)).collect())
    }
    /// `MultiproverKZG::open` (multiprover_kzg.rs:171-197): the public point `z`, every component of the shared polynomial.
    pub fn open_authenticated(&self, components: &[&[E::ScalarField]], z: &E::ScalarField)
                              -> Result<(Vec<E::G1Affine>, Vec<E::ScalarField>), GpuError> {
        let zs = vec![*z; components.len()];
        self.batch_open(components, &zs)
    }
    /// The key in the Lagrange basis of the size-2^log_n domain (`jf_srs_lagrange`): `commit` over it takes a polynomial's VALUES on
    /// the domain (zero / small values cost nothing / little).  About a second at 2^20, once per key and domain.
    pub fn lagrange(&self, log_n: u32, mask_points: bool) -> Result<Self, GpuError> {
        let mut srs = ptr::null_mut();
        self.gpu.check(unsafe { sys::jf_srs_lagrange(self.gpu.ctx, self.srs, log_n, mask_points as c_int, &mut srs) })?;
        Ok(Self { gpu: self.gpu, srs, len: (1usize << log_n) + if mask_points { 2 } else { 0 }, _e: PhantomData })
    }
    /// floor(p / Z_D) for the component vectors of a share (`compute_linking_quotient` of the collaborative prover,
    /// multiprover/proof_system/proof_linking.rs:127-138): division by the public vanishing polynomial of a link group is linear.
    pub fn div_link_domain(&self, polys: &[&[E::ScalarField]], layout: (usize, usize, usize)) -> Result<Vec<Vec<E::ScalarField>>, GpuError> {
        let lens: Vec<usize> = polys.iter().map(|p| p.len()).collect();
        let ptrs: Vec<*const u64> = polys.iter().map(|p| p.as_ptr() as *const u64).collect();
        let mut outs: Vec<Vec<u64>> = lens.iter().map(|&l| vec![0u64; 4 * l.saturating_sub(layout.2).max(1)]).collect();
        let optrs: Vec<*mut u64> = outs.iter_mut().map(|o| o.as_mut_ptr()).collect();
        self.gpu.check(unsafe {
            sys::jf_poly_div_link_domain(self.gpu.ctx, E::FR, ptrs.as_ptr(), lens.as_ptr(), polys.len(), layout.0 as u32, layout.1, layout.2,
                                         0, optrs.as_ptr())
        })?;
        Ok(outs.iter().zip(&lens).map(|(o, &l)| {
            (0..l.saturating_sub(layout.2)).map(|i| E::fr_from_mont([o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]])).collect()
        }).collect())
    }

    /// `PlonkKzgSnark::link_proofs` (plonk/src/proof_system/proof_linking.rs:79-112) from the two hints' wire polynomials and
    /// commitments: -> (quotient commitment, opening proof).  `layout` = (alignment, offset, size) of the `GroupLayout`.
    #[allow(clippy::too_many_arguments)]
    pub fn link_proofs(&self, a1: &[E::ScalarField], a1_comm: &E::G1Affine, a2: &[E::ScalarField], a2_comm: &E::G1Affine,
                       layout: (usize, usize, usize), solidity_transcript: bool) -> Result<(E::G1Affine, E::G1Affine), GpuError> {
        let (c1, i1) = E::point_to_mont(a1_comm);
        let (c2, i2) = E::point_to_mont(a2_comm);
        let mut out = MaybeUninit::<sys::jf_link_proof>::uninit();
        self.gpu.check(unsafe {
            sys::jf_plonk_link_proofs(self.gpu.ctx, self.srs, a1.as_ptr() as *const u64, a1.len(), c1.as_ptr(), i1 as c_int,
                                      a2.as_ptr() as *const u64, a2.len(), c2.as_ptr(), i2 as c_int, layout.0 as u32, layout.1, layout.2,
                                      if solidity_transcript { 0 } else { 1 }, 0, out.as_mut_ptr())
        })?;
        let lp = unsafe { out.assume_init() };
        let w = 2 * E::L;
        Ok((E::point_from_mont(&lp.quotient_commitment[..w], lp.quotient_inf != 0), E::point_from_mont(&lp.opening_proof[..w], lp.opening_inf != 0)))
    }
}
impl<E: GpuCurve> Drop for GpuCommitKey<E> { fn drop(&mut self) { unsafe { sys::jf_srs_free(self.gpu.ctx, self.srs) } } }

// ---- what the patched reference calls (generic over ANY pairing; None = curve not implemented on the GPU) --------------
type KeyCache = Mutex<HashMap<(TypeId, usize, usize), Arc<dyn Any + Send + Sync>>>;
fn key_cache() -> &'static KeyCache { static C: OnceLock<KeyCache> = OnceLock::new(); C.get_or_init(Default::default) }

/// The resident copy of `powers_of_g`, uploaded (and its window tables built) the first time this slice is seen.  The key of
/// a `ProvingKey` is immutable across proofs (plonk/src/proof_system/structs.rs:583), so the slice address + length identify it;
/// `forget_commit_key` drops the entry when the parameters are dropped.  This is also what replaces the per-call
/// `convert_prover_params` of the collaborative prover (multiprover_kzg.rs:232-234).
pub fn cached_commit_key<E: GpuCurve>(powers_of_g: &[E::G1Affine]) -> Result<Arc<GpuCommitKey<E>>, GpuError> {
    let id = (TypeId::of::<E>(), powers_of_g.as_ptr() as usize, powers_of_g.len());
    let mut cache = key_cache().lock().unwrap();
    if let Some(k) = cache.get(&id) { return Ok(k.clone().downcast::<GpuCommitKey<E>>().unwrap()); }
    let key = Arc::new(GpuCommitKey::<E>::load(global()?, powers_of_g)?);
    cache.insert(id, key.clone());
    Ok(key)
}
pub fn forget_commit_key<E: GpuCurve>(powers_of_g: &[E::G1Affine]) {
    key_cache().lock().unwrap().remove(&(TypeId::of::<E>(), powers_of_g.as_ptr() as usize, powers_of_g.len()));
}

/// View `&[E::X]` as `&[C::X]` when `E` and `C` are the same pairing.
fn same<E: Pairing, C: GpuCurve, T: 'static, U: 'static>(x: &[T]) -> Option<&[U]> {
    if TypeId::of::<E>() == TypeId::of::<C>() && TypeId::of::<T>() == TypeId::of::<U>() {
        Some(unsafe { core::slice::from_raw_parts(x.as_ptr() as *const U, x.len()) })
    } else { None }
}
fn back<A: 'static, B: 'static>(a: A) -> B {   // identity when A == B (checked)
    assert_eq!(TypeId::of::<A>(), TypeId::of::<B>());
    let b = unsafe { core::ptr::read(&a as *const A as *const B) };
    core::mem::forget(a);
    b
}

macro_rules! dispatch {
    ($E:ty, $pts:expr, |$C:ident, $p:ident| $body:expr) => {{
        if let Some($p) = same::<$E, ark_bn254::Bn254, <$E as Pairing>::G1Affine, ark_bn254::G1Affine>($pts) {
            type $C = ark_bn254::Bn254;
            Some($body)
        } else if let Some($p) = same::<$E, ark_bls12_381::Bls12_381, <$E as Pairing>::G1Affine, ark_bls12_381::G1Affine>($pts) {
            type $C = ark_bls12_381::Bls12_381;
            Some($body)
        } else { None }
    }};
}

/// `commit` for the patched `UnivariateKzgPCS::<E>::commit`.  `None`: E is neither BN254 nor BLS12-381 (the caller keeps its
/// arkworks path for, e.g., BW6-761); `Some(Err(_))`: GPU failure -- there is no silent fallback for an implemented curve.
pub fn try_commit<E: Pairing>(powers_of_g: &[E::G1Affine], coeffs: &[E::ScalarField]) -> Option<Result<E::G1Affine, GpuError>> {
    dispatch!(E, powers_of_g, |C, pts| {
        let c = same::<E, C, E::ScalarField, <C as Pairing>::ScalarField>(coeffs).unwrap();
        cached_commit_key::<C>(pts).and_then(|k| k.commit(c)).map(back::<<C as Pairing>::G1Affine, E::G1Affine>)
    })
}
pub fn try_batch_commit<E: Pairing>(powers_of_g: &[E::G1Affine], polys: &[&[E::ScalarField]]) -> Option<Result<Vec<E::G1Affine>, GpuError>> {
    dispatch!(E, powers_of_g, |C, pts| {
        let ps: Vec<&[<C as Pairing>::ScalarField]> = polys.iter().map(|p| same::<E, C, E::ScalarField, _>(p).unwrap()).collect();
        cached_commit_key::<C>(pts).and_then(|k| k.batch_commit(&ps)).map(back::<Vec<<C as Pairing>::G1Affine>, Vec<E::G1Affine>>)
    })
}
#[allow(clippy::type_complexity)]
pub fn try_batch_open<E: Pairing>(powers_of_g: &[E::G1Affine], polys: &[&[E::ScalarField]], points: &[E::ScalarField])
                                  -> Option<Result<(Vec<E::G1Affine>, Vec<E::ScalarField>), GpuError>> {
    dispatch!(E, powers_of_g, |C, pts| {
        let ps: Vec<&[<C as Pairing>::ScalarField]> = polys.iter().map(|p| same::<E, C, E::ScalarField, _>(p).unwrap()).collect();
        let zs = same::<E, C, E::ScalarField, <C as Pairing>::ScalarField>(points).unwrap();
        cached_commit_key::<C>(pts).and_then(|k| k.batch_open(&ps, zs))
            .map(back::<(Vec<<C as Pairing>::G1Affine>, Vec<<C as Pairing>::ScalarField>), (Vec<E::G1Affine>, Vec<E::ScalarField>)>)
    })
}
/// `CurvePoint::msm_authenticated(&poly.coeffs, &srs)` of the collaborative prover: one point per component vector.
pub fn try_commit_authenticated<E: Pairing>(powers_of_g: &[E::G1Affine], components: &[&[E::ScalarField]])
                                            -> Option<Result<Vec<E::G1Affine>, GpuError>> {
    dispatch!(E, powers_of_g, |C, pts| {
        let cs: Vec<&[<C as Pairing>::ScalarField]> = components.iter().map(|p| same::<E, C, E::ScalarField, _>(p).unwrap()).collect();
        cached_commit_key::<C>(pts).and_then(|k| k.commit_authenticated(&cs)).map(back::<Vec<<C as Pairing>::G1Affine>, Vec<E::G1Affine>>)
    })
}

/// `link_proofs` for the patched `PlonkKzgSnark::<E>::link_proofs`: quotient, its commitment, the challenge and the opening in
/// one call.  `solidity_transcript`: T is `SolidityTranscript` (else `StandardTranscript`).
#[allow(clippy::too_many_arguments, clippy::type_complexity)]
pub fn try_link_proofs<E: Pairing>(powers_of_g: &[E::G1Affine], a1: &[E::ScalarField], a1_comm: &E::G1Affine, a2: &[E::ScalarField],
                                   a2_comm: &E::G1Affine, layout: (usize, usize, usize), solidity_transcript: bool)
                                   -> Option<Result<(E::G1Affine, E::G1Affine), GpuError>> {
    dispatch!(E, powers_of_g, |C, pts| {
        let p1 = same::<E, C, E::ScalarField, <C as Pairing>::ScalarField>(a1).unwrap();
        let p2 = same::<E, C, E::ScalarField, <C as Pairing>::ScalarField>(a2).unwrap();
        let c1 = same::<E, C, E::G1Affine, <C as Pairing>::G1Affine>(core::slice::from_ref(a1_comm)).unwrap();
        let c2 = same::<E, C, E::G1Affine, <C as Pairing>::G1Affine>(core::slice::from_ref(a2_comm)).unwrap();
        cached_commit_key::<C>(pts).and_then(|k| k.link_proofs(p1, &c1[0], p2, &c2[0], layout, solidity_transcript))
            .map(back::<(<C as Pairing>::G1Affine, <C as Pairing>::G1Affine), (E::G1Affine, E::G1Affine)>)
    })
}

// ---- NTT ---------------------------------------------------------------------------------------------------------------
/// `domain.fft` / `domain.ifft` / `coset.fft` / `coset.ifft` on `batch` vectors of 2^log_n elements, in place
/// (prover.rs:552-567,672; relation/src/constraint_system.rs:1172-1257).  The vectors cross PCIe both ways: a caller that can keep
/// its polynomials resident should use the prover entry points (or `jf_ntt_device`) instead -- see INTEGRATION.md.
pub fn ntt<E: GpuCurve>(gpu: &Gpu, data: &mut [E::ScalarField], in_len: usize, log_n: u32, inverse: bool,
                        coset_offset: Option<&E::ScalarField>, batch: usize) -> Result<(), GpuError> {
    let n = 1usize << log_n;
    assert!(data.len() >= batch * n);
    gpu.check(unsafe {
        sys::jf_ntt(gpu.ctx, E::FR, data.as_mut_ptr() as *mut u64, in_len, log_n, inverse as c_int,
                    coset_offset.map_or(ptr::null(), |g| g as *const E::ScalarField as *const u64), batch, n)
    })
}

/// `polys` coefficient vectors (`in_len` <= 2n each, contiguous) evaluated on the cosets `offsets[r] * <w_n>`:
/// `out[(p * rows + r) * n + i] = poly_p(offsets[r] * w_n^i)`.  With `offsets[r] = g * w_8n^r` the rows are the residue
/// classes mod 8 of the 8n-point `coset.fft` of prover.rs:552-567; six rows determine the quotient polynomial.
pub fn ntt_cosets<E: GpuCurve>(gpu: &Gpu, polys: &[E::ScalarField], in_len: usize, log_n: u32, offsets: &[E::ScalarField],
                               out: &mut [E::ScalarField]) -> Result<(), GpuError> {
    let (n, count) = (1usize << log_n, polys.len() / in_len.max(1));
    assert!(polys.len() == count * in_len && out.len() >= count * offsets.len() * n);
    gpu.check(unsafe {
        sys::jf_ntt_cosets(gpu.ctx, E::FR, polys.as_ptr() as *const u64, in_len, in_len, count, log_n, 0,
                           offsets.as_ptr() as *const u64, offsets.len() as c_int, out.as_mut_ptr() as *mut u64)
    })
}

// ---- several GPUs, one process -------------------------------------------------------------------------------------------
/// `jf_group`: the commit key split by point range over the listed devices, batched transforms dealt out by polynomial.
pub struct GpuGroup { g: *mut sys::jf_group }
unsafe impl Send for GpuGroup {}
unsafe impl Sync for GpuGroup {}
impl GpuGroup {
    pub fn new(devices: &[i32]) -> Result<Self, GpuError> {
        let mut g = ptr::null_mut();
        match unsafe { sys::jf_group_create(devices.as_ptr(), devices.len() as c_int, &mut g) } {
            sys::JF_OK => Ok(Self { g }),
            rc => Err(map_status(rc, "jf_group_create failed".into())),
        }
    }
    fn check(&self, rc: c_int) -> Result<(), GpuError> {
        if rc == sys::JF_OK { return Ok(()); }
        Err(map_status(rc, unsafe { CStr::from_ptr(sys::jf_group_last_error(self.g)) }.to_string_lossy().into_owned()))
    }
    pub fn load_key<E: GpuCurve>(&self, powers_of_g: &[E::G1Affine]) -> Result<GroupCommitKey<'_, E>, GpuError> {
        let mut srs = ptr::null_mut();
        let (stride, inf_off) = E::affine_layout();
        self.check(unsafe {
            sys::jf_group_srs_load(self.g, E::CURVE, powers_of_g.as_ptr() as *const _, powers_of_g.len(), stride, inf_off as _, 0, 1, &mut srs)
        })?;
        Ok(GroupCommitKey { group: self, srs, _e: PhantomData })
    }
    /// the 25 coset transforms of prover.rs:552-567 (a rayon `par_iter` over polynomials) spread over the GPUs
    pub fn ntt<E: GpuCurve>(&self, data: &mut [E::ScalarField], in_len: usize, log_n: u32, inverse: bool,
                            coset_offset: Option<&E::ScalarField>, batch: usize) -> Result<(), GpuError> {
        let n = 1usize << log_n;
        assert!(data.len() >= batch * n);
        self.check(unsafe {
            sys::jf_group_ntt(self.g, E::FR, data.as_mut_ptr() as *mut u64, in_len, log_n, inverse as c_int,
                              coset_offset.map_or(ptr::null(), |g| g as *const E::ScalarField as *const u64), batch, n)
        })
    }
}
impl Drop for GpuGroup { fn drop(&mut self) { unsafe { sys::jf_group_destroy(self.g) } } }

pub struct GroupCommitKey<'a, E: GpuCurve> { group: &'a GpuGroup, srs: *mut sys::jf_group_srs, _e: PhantomData<E> }
impl<E: GpuCurve> GroupCommitKey<'_, E> {
    /// `commit` with the key range-sharded over the group's GPUs (one partial sum per GPU, joined on the host)
    pub fn commit(&self, coeffs: &[E::ScalarField]) -> Result<E::G1Affine, GpuError> {
        let nz = low_zeros(coeffs);
        let (mut xy, mut inf) = ([0u64; 12], 0);
        self.group.check(unsafe {
            sys::jf_group_msm(self.group.g, self.srs, nz, coeffs[nz..].as_ptr() as *const u64, coeffs.len() - nz, 1, xy.as_mut_ptr(), &mut inf)
        })?;
        Ok(E::point_from_mont(&xy, inf != 0))
    }
}
impl<E: GpuCurve> Drop for GroupCommitKey<'_, E> { fn drop(&mut self) { unsafe { sys::jf_group_srs_free(self.group.g, self.srs) } } }

// ---- prover rounds -------------------------------------------------------------------------------------------------------
/// `ProvingKey` resident on the GPU; `prove` == `PlonkKzgSnark::prove` for one TurboPlonk instance (snark.rs:201-469).
pub struct GpuProvingKey<E: GpuCurve> { key: Arc<GpuCommitKey<E>>, pk: *mut sys::jf_plonk_pk }
impl<E: GpuCurve> GpuProvingKey<E> {
    #[allow(clippy::too_many_arguments)]
    pub fn preprocess(key: Arc<GpuCommitKey<E>>, log_n: u32, selectors: &[E::ScalarField], extended_perm: &[E::ScalarField],
                      k: &[E::ScalarField; 5], wire_variables: &[u32], num_vars: usize, io_gate_ids: &[u32], flags: i32)
                      -> Result<Self, GpuError> {
        let mut pk = ptr::null_mut();
        key.gpu.check(unsafe {
            sys::jf_plonk_preprocess(key.gpu.ctx, key.srs, log_n, selectors.as_ptr() as *const u64, extended_perm.as_ptr() as *const u64,
                                     k.as_ptr() as *const u64, wire_variables.as_ptr(), num_vars, io_gate_ids.as_ptr(),
                                     io_gate_ids.len(), flags, &mut pk)
        })?;
        Ok(Self { key, pk })
    }
    /// `blinders`: 17 elements drawn with `Fr::rand(prng)` in the reference's order (prover.rs:483-484, 946-957).
    pub fn prove(&self, witness: &[E::ScalarField], blinders: &[E::ScalarField; 17], solidity_transcript: bool, extra: Option<&[u8]>)
                 -> Result<sys::jf_plonk_proof, GpuError> {
        let mut out = MaybeUninit::<sys::jf_plonk_proof>::uninit();
        self.key.gpu.check(unsafe {
            sys::jf_plonk_prove(self.key.gpu.ctx, self.pk, witness.as_ptr() as *const u64, blinders.as_ptr() as *const u64,
                                if solidity_transcript { 0 } else { 1 }, extra.map_or(ptr::null(), |e| e.as_ptr()),
                                extra.map_or(0, |e| e.len()), out.as_mut_ptr())
        })?;
        Ok(unsafe { out.assume_init() })   // 13 points + 10 scalars -> mpc_plonk::proof_system::structs::Proof<E>
    }
    /// the points of a `jf_plonk_proof` as arkworks values: (wires, prod_perm, split_quot, opening, shifted_opening)
    #[allow(clippy::type_complexity)]
    pub fn proof_points(p: &sys::jf_plonk_proof) -> (Vec<E::G1Affine>, E::G1Affine, Vec<E::G1Affine>, E::G1Affine, E::G1Affine) {
        let w = 2 * E::L;
        ((0..5).map(|j| E::point_from_mont(&p.wires_poly_comms[w * j..w * (j + 1)], p.wires_inf[j] != 0)).collect(),
         E::point_from_mont(&p.prod_perm_poly_comm[..w], p.prod_perm_inf != 0),
         (0..5).map(|j| E::point_from_mont(&p.split_quot_poly_comms[w * j..w * (j + 1)], p.split_inf[j] != 0)).collect(),
         E::point_from_mont(&p.opening_proof[..w], p.opening_inf != 0),
         E::point_from_mont(&p.shifted_opening_proof[..w], p.shifted_opening_inf != 0))
    }
}
impl<E: GpuCurve> Drop for GpuProvingKey<E> { fn drop(&mut self) { unsafe { sys::jf_plonk_pk_free(self.key.gpu.ctx, self.pk) } } }

impl<E: GpuCurve> GpuProvingKey<E> {
    /// `LinkingHint::linking_wire_poly` of the last proof made with this key (snark.rs:96-100): n + 2 coefficients.
    pub fn link_hint(&self, n: usize) -> Result<Vec<E::ScalarField>, GpuError> {
        let mut buf = vec![0u64; 4 * (n + 2)];
        let mut len = 0usize;
        self.key.gpu.check(unsafe { sys::jf_plonk_link_hint(self.key.gpu.ctx, self.pk, buf.as_mut_ptr(), n + 2, &mut len) })?;
        Ok((0..len).map(|i| E::fr_from_mont([buf[4 * i], buf[4 * i + 1], buf[4 * i + 2], buf[4 * i + 3]])).collect())
    }
    /// `link_proofs` with both wire polynomials still in HBM (the last proofs made with `self` and `rhs`).
    pub fn link_proofs_resident(&self, lhs_proof: &sys::jf_plonk_proof, rhs: &Self, rhs_proof: &sys::jf_plonk_proof,
                                layout: (usize, usize, usize), solidity_transcript: bool) -> Result<(E::G1Affine, E::G1Affine), GpuError> {
        let mut out = MaybeUninit::<sys::jf_link_proof>::uninit();
        self.key.gpu.check(unsafe {
            sys::jf_plonk_link_proofs_resident(self.key.gpu.ctx, self.pk, lhs_proof, rhs.pk, rhs_proof, layout.0 as u32, layout.1, layout.2,
                                               if solidity_transcript { 0 } else { 1 }, 0, out.as_mut_ptr())
        })?;
        let lp = unsafe { out.assume_init() };
        let w = 2 * E::L;
        Ok((E::point_from_mont(&lp.quotient_commitment[..w], lp.quotient_inf != 0), E::point_from_mont(&lp.opening_proof[..w], lp.opening_inf != 0)))
    }
    /// ONE proof on several GPUs (one process per GPU): every rank calls `prove` with the same witness and blinders afterwards;
    /// commitments are split by point range and round 3 by sub-coset (`jf_plonk_pk_shard_commits`).  `comm` and `key_slice` must
    /// outlive the sharded use of the key; `unshard` returns to one-GPU operation.
    ///
    /// # Safety
    /// `comm` must be a live `jf_comm` created on this key's context.
    pub unsafe fn shard(&self, comm: *mut sys::jf_comm, key_slice: &GpuCommitKey<E>, slice_start: usize, shard_round3: bool) -> Result<(), GpuError> {
        self.key.gpu.check(sys::jf_plonk_pk_shard_commits(self.key.gpu.ctx, self.pk, comm, key_slice.srs, slice_start, shard_round3 as c_int))
    }
    pub fn unshard(&self) -> Result<(), GpuError> {
        self.key.gpu.check(unsafe { sys::jf_plonk_pk_shard_commits(self.key.gpu.ctx, self.pk, ptr::null_mut(), ptr::null(), 0, 0) })
    }
}

impl<E: GpuCurve> GpuProvingKey<E> {
    /// `PlonkKzgSnark::batch_prove` (snark.rs:201-469) for TurboPlonk keys: one transcript, ONE quotient, one pair of opening
    /// proofs.  `blinders`: `keys.len() * 13 + 4` elements in the reference's prng order (all wire masks, all z masks, split).
    /// Record i of the result holds instance i's parts; the shared parts are replicated in every record.
    pub fn batch_prove(keys: &[&Self], witnesses: &[&[E::ScalarField]], blinders: &[E::ScalarField], solidity_transcript: bool,
                       extra: Option<&[u8]>) -> Result<Vec<sys::jf_plonk_proof>, GpuError> {
        assert!(!keys.is_empty() && keys.len() == witnesses.len() && blinders.len() == keys.len() * 13 + 4);
        let pks: Vec<*mut sys::jf_plonk_pk> = keys.iter().map(|k| k.pk).collect();
        let ws: Vec<*const u64> = witnesses.iter().map(|w| w.as_ptr() as *const u64).collect();
        let mut out: Vec<MaybeUninit<sys::jf_plonk_proof>> = (0..keys.len()).map(|_| MaybeUninit::uninit()).collect();
        keys[0].key.gpu.check(unsafe {
            sys::jf_plonk_batch_prove(keys[0].key.gpu.ctx, pks.as_ptr(), pks.len(), ws.as_ptr(), blinders.as_ptr() as *const u64,
                                      if solidity_transcript { 0 } else { 1 }, extra.map_or(ptr::null(), |e| e.as_ptr()),
                                      extra.map_or(0, |e| e.len()), out.as_mut_ptr() as *mut sys::jf_plonk_proof)
        })?;
        Ok(out.into_iter().map(|p| unsafe { p.assume_init() }).collect())
    }
}

/// UltraPlonk (`PlonkCircuit::new_ultra_plonk`): the same key type on the C side, built from 14 selector columns, 6 sigma rows,
/// 6 `k`, 6 wire rows and the three per-gate Plookup columns; proofs carry the `PlookupProof`.
pub struct GpuUltraProvingKey<E: GpuCurve> { key: Arc<GpuCommitKey<E>>, pk: *mut sys::jf_plonk_pk }
impl<E: GpuCurve> GpuUltraProvingKey<E> {
    #[allow(clippy::too_many_arguments)]
    pub fn preprocess(key: Arc<GpuCommitKey<E>>, log_n: u32, selectors: &[E::ScalarField], extended_perm: &[E::ScalarField],
                      k: &[E::ScalarField; 6], wire_variables: &[u32], num_vars: usize, io_gate_ids: &[u32], range_bit_len: u32,
                      table_key: &[E::ScalarField], table_dom_sep: &[E::ScalarField], q_dom_sep: &[E::ScalarField], flags: i32)
                      -> Result<Self, GpuError> {
        let mut pk = ptr::null_mut();
        key.gpu.check(unsafe {
            sys::jf_ultraplonk_preprocess(key.gpu.ctx, key.srs, log_n, selectors.as_ptr() as *const u64, extended_perm.as_ptr() as *const u64,
                                          k.as_ptr() as *const u64, wire_variables.as_ptr(), num_vars, io_gate_ids.as_ptr(),
                                          io_gate_ids.len(), range_bit_len, table_key.as_ptr() as *const u64,
                                          table_dom_sep.as_ptr() as *const u64, q_dom_sep.as_ptr() as *const u64, flags, &mut pk)
        })?;
        Ok(Self { key, pk })
    }
    /// `blinders`: 29 elements in the reference's order (6 x 2 wires, 3 + 3 for h1 / h2, 3 for z, 3 for the lookup product, 5 split).
    pub fn prove(&self, witness: &[E::ScalarField], blinders: &[E::ScalarField; 29], solidity_transcript: bool, extra: Option<&[u8]>)
                 -> Result<sys::jf_ultraplonk_proof, GpuError> {
        let mut out = MaybeUninit::<sys::jf_ultraplonk_proof>::uninit();
        self.key.gpu.check(unsafe {
            sys::jf_ultraplonk_prove(self.key.gpu.ctx, self.pk, witness.as_ptr() as *const u64, blinders.as_ptr() as *const u64,
                                     if solidity_transcript { 0 } else { 1 }, extra.map_or(ptr::null(), |e| e.as_ptr()),
                                     extra.map_or(0, |e| e.len()), out.as_mut_ptr())
        })?;
        Ok(unsafe { out.assume_init() })
    }
}
impl<E: GpuCurve> Drop for GpuUltraProvingKey<E> { fn drop(&mut self) { unsafe { sys::jf_plonk_pk_free(self.key.gpu.ctx, self.pk) } } }
