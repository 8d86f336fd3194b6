//! Drop-in replacement for the hot path of `snark-bn254-verifier` on NVIDIA B200.
//!
//! `Groth16Verifier::verify` and `PlonkVerifier::verify` keep the reference's signatures and outcomes
//! (reference `verifier/src/lib.rs:44-49`, `:69-74`): `Ok(true)` / `Ok(false)` / `Err(..)`, and a panic where the
//! reference's `unwrap()` on a parser error (or substrate-bn's affine conversion) panics.  A single proof is a batch
//! of one through the same CUDA kernels -- there is no CPU fallback.  `verify_batch` and `verify_many` are new.
//!
//! gnark proof / VK framing stays here on the host side only as byte slices: `libbn254v` decompresses VK points on
//! the host once per VK (cached by sha256(vk) inside the library) and decodes + validates proof bytes on the device.
//!
//! This crate is source only in this repository (no Rust toolchain in its build environment); see INTEGRATION.md.
mod ffi;

use bn::Fr;
use core::ffi::CStr;
use thiserror::Error;

/// reference `verifier/src/groth16/error.rs`
#[derive(Error, Debug, PartialEq, Eq)]
pub enum Groth16Error {
    #[error("Prepare inputs failed")]
    PrepareInputsFailed,
}

/// reference `verifier/src/plonk/error.rs` (the variants `verify_plonk` can return)
#[derive(Error, Debug, PartialEq, Eq)]
pub enum PlonkError {
    #[error("Bsb22 commitment number mismatch")]
    Bsb22CommitmentMismatch,
    #[error("Invalid witness")]
    InvalidWitness,
    #[error("Inverse not found")]
    InverseNotFound,
    #[error("Opening linear polynomial mismatch")]
    OpeningPolyMismatch,
    #[error("Invalid number of digests")]
    InvalidNumberOfDigests,
    #[error("Pairing check failed")]
    PairingCheckFailed,
}

/// Library-level failure (no CUDA device, CUDA error, bad argument): not an outcome of the reference.
#[derive(Error, Debug)]
#[error("bn254v error {code}: {message}")]
pub struct LibraryError {
    pub code: i32,
    pub message: String,
}

fn check(rc: i32) -> Result<(), LibraryError> {
    if rc == 0 {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(ffi::bn254v_last_error()) }.to_string_lossy().into_owned();
    Err(LibraryError { code: rc, message })
}

fn status_name(s: u8) -> String {
    unsafe { CStr::from_ptr(ffi::bn254v_status_name(s as i32)) }.to_string_lossy().into_owned()
}

/// Selects the CUDA devices batches are sharded over (`None`: all visible).  Optional: the first call initialises.
pub fn init(devices: Option<&[i32]>) -> Result<i32, LibraryError> {
    let (p, n) = devices.map_or((core::ptr::null(), 0), |d| (d.as_ptr(), d.len() as i32));
    check(unsafe { ffi::bn254v_init(p, n) })?;
    Ok(unsafe { ffi::bn254v_device_count() })
}

fn fr_be(x: &Fr, out: &mut Vec<u8>) {
    let mut b = [0u8; 32];
    x.into_u256().to_big_endian(&mut b).expect("32 bytes");
    out.extend_from_slice(&b);
}

/// ragged records -> one strided buffer + per-record lengths
fn pack(proofs: &[&[u8]]) -> (Vec<u8>, usize, Vec<u32>) {
    let stride = proofs.iter().map(|p| p.len()).max().unwrap_or(0).max(1);
    let mut buf = vec![0u8; stride * proofs.len()];
    let mut lens = Vec::with_capacity(proofs.len());
    for (i, p) in proofs.iter().enumerate() {
        buf[i * stride..i * stride + p.len()].copy_from_slice(p);
        lens.push(p.len() as u32);
    }
    (buf, stride, lens)
}

fn pack_inputs(inputs: &[&[Fr]]) -> (Vec<u8>, i32) {
    let k = inputs.first().map_or(0, |x| x.len());
    assert!(inputs.iter().all(|x| x.len() == k), "all proofs of a batch take the same number of public inputs");
    let mut out = Vec::with_capacity(32 * k * inputs.len());
    for xs in inputs {
        for x in xs.iter() {
            fr_be(x, &mut out);
        }
    }
    (out, k as i32)
}

/// VK handle from the library's cache (keyed by sha256(vk), kind, sign convention).  A VK that does not parse panics,
/// as the reference's `load_*_verifying_key_from_bytes(vk).unwrap()` does (`verifier/src/lib.rs:46,71`).
fn vk_handle(kind: i32, vk: &[u8], sign_mode: i32) -> *const ffi::bn254v_vk {
    let mut h: *const ffi::bn254v_vk = core::ptr::null();
    let rc = unsafe { ffi::bn254v_vk_cache_get(kind, vk.as_ptr(), vk.len(), sign_mode, &mut h) };
    if rc == ffi::E_VK_PARSE {
        panic!("called `Result::unwrap()` on an `Err` value: malformed verifying key");
    }
    check(rc).expect("bn254v");
    h
}

fn groth16_outcome(s: u8) -> Result<bool, Groth16Error> {
    match s {
        ffi::OK_TRUE => Ok(true),
        ffi::OK_FALSE => Ok(false),
        ffi::ERR_PREPARE_INPUTS => Err(Groth16Error::PrepareInputsFailed),
        s => panic!("{}", status_name(s)), // the reference panics here (parser unwrap / bn affine conversion)
    }
}

fn plonk_outcome(s: u8) -> Result<bool, PlonkError> {
    match s {
        ffi::OK_TRUE => Ok(true), // verify_plonk never returns Ok(false) (verifier/src/plonk/verify.rs:316)
        ffi::ERR_BSB22_MISMATCH => Err(PlonkError::Bsb22CommitmentMismatch),
        ffi::ERR_INVALID_WITNESS => Err(PlonkError::InvalidWitness),
        ffi::ERR_INVERSE_NOT_FOUND => Err(PlonkError::InverseNotFound),
        ffi::ERR_OPENING_POLY_MISMATCH => Err(PlonkError::OpeningPolyMismatch),
        ffi::ERR_INVALID_NUMBER_OF_DIGESTS => Err(PlonkError::InvalidNumberOfDigests),
        ffi::ERR_PAIRING_CHECK_FAILED => Err(PlonkError::PairingCheckFailed),
        s => panic!("{}", status_name(s)),
    }
}

/// A verifier for Groth16 zero-knowledge proofs (reference `verifier/src/lib.rs:37-49`).
#[derive(Debug)]
pub struct Groth16Verifier;

impl Groth16Verifier {
    /// Same signature and outcomes as the reference; a batch of one on the GPU.
    pub fn verify(proof: &[u8], vk: &[u8], public_inputs: &[Fr]) -> Result<bool, Groth16Error> {
        groth16_outcome(Self::verify_batch_status(&[proof], vk, &[public_inputs]).expect("bn254v")[0])
    }

    /// Many proofs against one VK: one outcome per proof, in order.  Panic statuses are returned as `Err(status name)`
    /// of the outer `Result` only through `verify_batch_status`; here a panic status panics like the reference would
    /// have for that proof.
    pub fn verify_batch(proofs: &[&[u8]], vk: &[u8], inputs: &[&[Fr]]) -> Vec<Result<bool, Groth16Error>> {
        Self::verify_batch_status(proofs, vk, inputs).expect("bn254v").into_iter().map(groth16_outcome).collect()
    }

    /// The raw status bytes (`enum bn254v_status`), for callers that want to tell the panic classes apart.
    pub fn verify_batch_status(proofs: &[&[u8]], vk: &[u8], inputs: &[&[Fr]]) -> Result<Vec<u8>, LibraryError> {
        assert_eq!(proofs.len(), inputs.len());
        let h = vk_handle(ffi::KIND_GROTH16, vk, 0);
        let (buf, stride, lens) = pack(proofs);
        let (inputs_be, k) = pack_inputs(inputs);
        let mut status = vec![255u8; proofs.len()];
        check(unsafe {
            ffi::bn254v_groth16_verify_batch(h, buf.as_ptr(), stride, lens.as_ptr(), inputs_be.as_ptr(), k, proofs.len(),
                                             status.as_mut_ptr(), core::ptr::null())
        })?;
        Ok(status)
    }

    /// OPT-IN, not in the reference: `true` iff every proof of the batch is valid (soundness error <= 2^-126 over
    /// scalars the library draws from the OS CSPRNG), for about half the GPU time of `verify_batch`.  `false` says
    /// nothing about which proof fails -- call `verify_batch` then.
    pub fn batch_all_valid(proofs: &[&[u8]], vk: &[u8], inputs: &[&[Fr]]) -> Result<bool, LibraryError> {
        assert_eq!(proofs.len(), inputs.len());
        let h = vk_handle(ffi::KIND_GROTH16, vk, 0);
        let (buf, stride, lens) = pack(proofs);
        let (inputs_be, k) = pack_inputs(inputs);
        let mut all_valid = 0u8;
        check(unsafe {
            ffi::bn254v_groth16_batch_all_valid(h, buf.as_ptr(), stride, lens.as_ptr(), inputs_be.as_ptr(), k,
                                                core::ptr::null(), proofs.len(), &mut all_valid, core::ptr::null_mut())
        })?;
        Ok(all_valid == 1)
    }
}

/// A verifier for Plonk zero-knowledge proofs (reference `verifier/src/lib.rs:62-74`).
#[derive(Debug)]
pub struct PlonkVerifier;

impl PlonkVerifier {
    /// Same signature and outcomes as the reference; a batch of one on the GPU.  The scalar of
    /// `kzg::batch_verify_multi_points` is drawn inside the library from the OS CSPRNG, where the reference calls
    /// `Fr::random(&mut OsRng)` (`verifier/src/plonk/kzg.rs:149-154`).
    pub fn verify(proof: &[u8], vk: &[u8], public_inputs: &[Fr]) -> Result<bool, PlonkError> {
        plonk_outcome(Self::verify_batch_status(&[proof], vk, &[public_inputs]).expect("bn254v")[0])
    }

    pub fn verify_batch(proofs: &[&[u8]], vk: &[u8], inputs: &[&[Fr]]) -> Vec<Result<bool, PlonkError>> {
        Self::verify_batch_status(proofs, vk, inputs).expect("bn254v").into_iter().map(plonk_outcome).collect()
    }

    pub fn verify_batch_status(proofs: &[&[u8]], vk: &[u8], inputs: &[&[Fr]]) -> Result<Vec<u8>, LibraryError> {
        assert_eq!(proofs.len(), inputs.len());
        let h = vk_handle(ffi::KIND_PLONK, vk, 0);
        let (buf, stride, lens) = pack(proofs);
        let (inputs_be, k) = pack_inputs(inputs);
        let mut status = vec![255u8; proofs.len()];
        check(unsafe {
            // rnd_be = NULL: production path, the library draws the batch-opening scalars (never caller-chosen)
            ffi::bn254v_plonk_verify_batch(h, buf.as_ptr(), stride, lens.as_ptr(), inputs_be.as_ptr(), k, core::ptr::null(),
                                           proofs.len(), status.as_mut_ptr(), core::ptr::null())
        })?;
        Ok(status)
    }
}

/// Proof system of one item of a mixed batch.
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Kind {
    Groth16,
    Plonk,
}

/// What one `verify` call of the reference receives.
#[derive(Debug, Clone, Copy)]
pub struct Item<'a> {
    pub kind: Kind,
    pub proof: &'a [u8],
    pub vk: &'a [u8],
    pub public_inputs: &'a [Fr],
}

/// Mixed batch over any number of verifying keys and both proof systems (`bn254v_verify_many`): the library groups the
/// items by (kind, sha256(vk), number of inputs) and runs one device batch per group.  Returns the status bytes in
/// item order (`enum bn254v_status`; 23 = the item's VK does not parse).
pub fn verify_many(items: &[Item<'_>]) -> Result<Vec<u8>, LibraryError> {
    let mut inputs_be: Vec<Vec<u8>> = Vec::with_capacity(items.len());
    for it in items {
        let mut b = Vec::with_capacity(32 * it.public_inputs.len());
        for x in it.public_inputs {
            fr_be(x, &mut b);
        }
        inputs_be.push(b);
    }
    let raw: Vec<ffi::bn254v_item> = items
        .iter()
        .zip(&inputs_be)
        .map(|(it, inp)| ffi::bn254v_item {
            kind: if it.kind == Kind::Groth16 { ffi::KIND_GROTH16 } else { ffi::KIND_PLONK },
            n_inputs: it.public_inputs.len() as i32,
            proof: it.proof.as_ptr(),
            proof_len: it.proof.len(),
            vk: it.vk.as_ptr(),
            vk_len: it.vk.len(),
            inputs_be: inp.as_ptr(),
        })
        .collect();
    let mut status = vec![255u8; items.len()];
    check(unsafe { ffi::bn254v_verify_many(raw.as_ptr(), raw.len(), 0, core::ptr::null(), status.as_mut_ptr()) })?;
    Ok(status)
}
