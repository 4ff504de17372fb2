//! Safe, arkworks-typed wrappers.  BN254 is spelled out; BLS12-381 is identical with 6-limb base-field
//! elements (`JF_BLS12_381`, `JF_BLS12_381_FR`).  NOT COMPILED in the build image (no Rust toolchain).
//!
//! Layout facts this file relies on (ark-ff 0.4 / ark-ec 0.4):
//!   * `Fp<MontBackend<_, 4>, 4>` is `#[repr(transparent)]`-like over `BigInt<4>([u64; 4])` in Montgomery form,
//!     so `&[Fr]` can be passed as `*const u64` (4 limbs per element) without conversion;
//!   * `short_weierstrass::Affine { x, y, infinity: bool }` is passed with its real stride and the byte offset
//!     of `infinity`; the library repacks once in `jf_srs_load`.
use ark_bn254::{Bn254, Fq, Fr, G1Affine};
use ark_ec::AffineRepr;
use ark_ff::{BigInt, Field, PrimeField, Zero};
use core::ffi::c_int;
use jf_b200_sys as sys;
use std::{ffi::CStr, mem::MaybeUninit, ptr};

#[derive(Debug)]
pub enum GpuError { InvalidParameters(String), Upstream(String), DomainCreation, WrongQuotientPolyDegree }

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
        let msg = unsafe { CStr::from_ptr(sys::jf_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(match rc {
            sys::JF_ERR_INVALID_ARG | sys::JF_ERR_SCALAR_RANGE => GpuError::InvalidParameters(msg), // PCSError::InvalidParameters
            sys::JF_ERR_DOMAIN_TOO_LARGE => GpuError::DomainCreation,                              // PlonkError::DomainCreationError
            sys::JF_ERR_QUOTIENT_DEGREE => GpuError::WrongQuotientPolyDegree,                      // SnarkError::WrongQuotientPolyDegree
            _ => GpuError::Upstream(msg),                                                          // PCSError::UpstreamError
        })
    }
}
impl Drop for Gpu { fn drop(&mut self) { unsafe { sys::jf_ctx_destroy(self.ctx) } } }

fn fq_from_mont(l: &[u64]) -> Fq { Fq::new_unchecked(BigInt::<4>([l[0], l[1], l[2], l[3]])) }
fn point(xy: &[u64], inf: c_int) -> G1Affine {
    if inf != 0 { G1Affine::identity() } else { G1Affine::new_unchecked(fq_from_mont(&xy[..4]), fq_from_mont(&xy[4..8])) }
}

/// `UnivariateProverParam::powers_of_g`, resident on the GPU with its window tables.
pub struct GpuCommitKey<'g> { gpu: &'g Gpu, srs: *mut sys::jf_srs }
impl<'g> GpuCommitKey<'g> {
    pub fn load(gpu: &'g Gpu, powers_of_g: &[G1Affine]) -> Result<Self, GpuError> {
        let mut srs = ptr::null_mut();
        let probe = G1Affine::identity();
        let inf_off = (&probe.infinity as *const bool as usize) - (&probe as *const G1Affine as usize);
        gpu.check(unsafe {
            sys::jf_srs_load(gpu.ctx, sys::JF_BN254, powers_of_g.as_ptr() as *const _, powers_of_g.len(),
                             core::mem::size_of::<G1Affine>(), inf_off as _, 0, 1, &mut srs)
        })?;
        Ok(Self { gpu, srs })
    }
    /// `UnivariateKzgPCS::commit` (mod.rs:90-116): degree check stays with the caller's `PCSError`.
    pub fn commit(&self, coeffs: &[Fr]) -> Result<G1Affine, GpuError> {
        let nz = coeffs.iter().take_while(|c| c.is_zero()).count();          // mod.rs:382-385
        let (mut xy, mut inf) = ([0u64; 8], 0);
        self.gpu.check(unsafe {
            sys::jf_msm(self.gpu.ctx, self.srs, nz, coeffs[nz..].as_ptr() as *const u64, coeffs.len() - nz, 1,
                        xy.as_mut_ptr(), &mut inf)
        })?;
        Ok(point(&xy, inf))
    }
    /// `batch_commit` (mod.rs:119-131): the rayon `par_iter` becomes one call.
    pub fn batch_commit(&self, polys: &[&[Fr]]) -> Result<Vec<G1Affine>, GpuError> {
        let offs: Vec<usize> = polys.iter().map(|p| p.iter().take_while(|c| c.is_zero()).count()).collect();
        let ptrs: Vec<*const u64> = polys.iter().zip(&offs).map(|(p, &o)| p[o..].as_ptr() as *const u64).collect();
        let lens: Vec<usize> = polys.iter().zip(&offs).map(|(p, &o)| p.len() - o).collect();
        let mut xy = vec![0u64; 8 * polys.len()];
        let mut inf = vec![0 as c_int; polys.len()];
        self.gpu.check(unsafe {
            sys::jf_msm_batch(self.gpu.ctx, self.srs, ptrs.as_ptr(), lens.as_ptr(), offs.as_ptr(), polys.len(), 1,
                              xy.as_mut_ptr(), inf.as_mut_ptr())
        })?;
        Ok((0..polys.len()).map(|i| point(&xy[8 * i..8 * i + 8], inf[i])).collect())
    }
    /// `open` (mod.rs:135-161): witness polynomial, its commitment and the evaluation on the GPU.
    pub fn open(&self, coeffs: &[Fr], z: &Fr) -> Result<(G1Affine, Fr), GpuError> {
        let (p, l) = (coeffs.as_ptr() as *const u64, coeffs.len());
        let (mut xy, mut inf, mut ev) = ([0u64; 8], 0, [0u64; 4]);
        self.gpu.check(unsafe {
            sys::jf_kzg_open(self.gpu.ctx, self.srs, &p, &l, 1, z as *const Fr as *const u64, xy.as_mut_ptr(), &mut inf,
                             ev.as_mut_ptr())
        })?;
        Ok((point(&xy, inf), Fr::new_unchecked(BigInt::<4>(ev))))
    }
}
impl Drop for GpuCommitKey<'_> { fn drop(&mut self) { unsafe { sys::jf_srs_free(self.gpu.ctx, self.srs) } } }

/// `domain.fft` / `domain.ifft` / `coset.fft` / `coset.ifft` on `batch` vectors of 2^log_n elements, in place.
pub fn ntt(gpu: &Gpu, data: &mut [Fr], in_len: usize, log_n: u32, inverse: bool, coset_offset: Option<&Fr>, batch: usize)
           -> Result<(), GpuError> {
    let n = 1usize << log_n;
    assert!(data.len() >= batch * n);
    gpu.check(unsafe {
        sys::jf_ntt(gpu.ctx, sys::JF_BN254_FR, data.as_mut_ptr() as *mut u64, in_len, log_n, inverse as c_int,
                    coset_offset.map_or(ptr::null(), |g| g as *const Fr as *const u64), batch, n)
    })
}

/// `polys` coefficient vectors (`in_len` <= 2n each, contiguous) evaluated on the cosets `offsets[r] * <w_n>`:
/// `out[(p * rows + r) * n + i] = poly_p(offsets[r] * w_n^i)`.  With `offsets[r] = g * w_8n^r` the rows are the residue
/// classes mod 8 of the 8n-point `coset.fft` of prover.rs:552-567; six rows determine the quotient polynomial.
pub fn ntt_cosets(gpu: &Gpu, polys: &[Fr], in_len: usize, log_n: u32, offsets: &[Fr], out: &mut [Fr]) -> Result<(), GpuError> {
    let (n, count) = (1usize << log_n, polys.len() / in_len.max(1));
    assert!(polys.len() == count * in_len && out.len() >= count * offsets.len() * n);
    gpu.check(unsafe {
        sys::jf_ntt_cosets(gpu.ctx, sys::JF_BN254_FR, polys.as_ptr() as *const u64, in_len, in_len, count, log_n, 0,
                           offsets.as_ptr() as *const u64, offsets.len() as c_int, out.as_mut_ptr() as *mut u64)
    })
}

/// `ProvingKey` resident on the GPU; `prove` == `PlonkKzgSnark::prove` for one TurboPlonk instance.
pub struct GpuProvingKey<'g> { gpu: &'g Gpu, pk: *mut sys::jf_plonk_pk }
impl<'g> GpuProvingKey<'g> {
    #[allow(clippy::too_many_arguments)]
    pub fn preprocess(gpu: &'g Gpu, ck: &GpuCommitKey<'g>, log_n: u32, selectors: &[Fr], extended_perm: &[Fr], k: &[Fr; 5],
                      wire_variables: &[u32], num_vars: usize, io_gate_ids: &[u32]) -> Result<Self, GpuError> {
        let mut pk = ptr::null_mut();
        gpu.check(unsafe {
            sys::jf_plonk_preprocess(gpu.ctx, ck.srs, log_n, selectors.as_ptr() as *const u64, extended_perm.as_ptr() as *const u64,
                                     k.as_ptr() as *const u64, wire_variables.as_ptr(), num_vars, io_gate_ids.as_ptr(),
                                     io_gate_ids.len(), 2 /* skip zero selectors */, &mut pk)
        })?;
        Ok(Self { gpu, pk })
    }
    /// `blinders`: 17 elements drawn with `Fr::rand(prng)` in the reference's order (prover.rs:483-484, 946-957).
    pub fn prove(&self, witness: &[Fr], blinders: &[Fr; 17], solidity_transcript: bool, extra: Option<&[u8]>)
                 -> Result<sys::jf_plonk_proof, GpuError> {
        let mut out = MaybeUninit::<sys::jf_plonk_proof>::uninit();
        self.gpu.check(unsafe {
            sys::jf_plonk_prove(self.gpu.ctx, self.pk, witness.as_ptr() as *const u64, blinders.as_ptr() as *const u64,
                                if solidity_transcript { 0 } else { 1 }, extra.map_or(ptr::null(), |e| e.as_ptr()),
                                extra.map_or(0, |e| e.len()), out.as_mut_ptr())
        })?;
        Ok(unsafe { out.assume_init() })   // 13 points + 10 scalars -> mpc_plonk::proof_system::structs::Proof<Bn254>
    }
}
impl Drop for GpuProvingKey<'_> { fn drop(&mut self) { unsafe { sys::jf_plonk_pk_free(self.gpu.ctx, self.pk) } } }

#[allow(dead_code)]
fn _type_anchors(_: Bn254, _: fn(&Fr) -> <Fr as PrimeField>::BigInt, _: fn(&Fr) -> Option<Fr>) {}
#[allow(dead_code)]
fn _inv(x: &Fr) -> Option<Fr> { x.inverse() }
