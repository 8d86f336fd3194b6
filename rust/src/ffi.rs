//! `extern "C"` declarations of include/bn254v.h (the verifier's C ABI), one to one.
#![allow(non_camel_case_types, dead_code)]

#[repr(C)]
pub struct bn254v_vk {
    _private: [u8; 0],
}

#[repr(C)]
pub struct bn254v_debug {
    pub g1_out: *mut u8,
    pub fr_out: *mut u8,
    pub miller_out: *mut u8,
    pub gt_out: *mut u8,
}

/// One item of a mixed batch (`bn254v_item`).
#[repr(C)]
pub struct bn254v_item {
    pub kind: i32,
    pub n_inputs: i32,
    pub proof: *const u8,
    pub proof_len: usize,
    pub vk: *const u8,
    pub vk_len: usize,
    pub inputs_be: *const u8,
}

pub const KIND_GROTH16: i32 = 0;
pub const KIND_PLONK: i32 = 1;

// enum bn254v_status
pub const OK_TRUE: u8 = 0;
pub const OK_FALSE: u8 = 1;
pub const ERR_PREPARE_INPUTS: u8 = 2;
pub const ERR_BSB22_MISMATCH: u8 = 3;
pub const ERR_INVALID_WITNESS: u8 = 4;
pub const ERR_INVERSE_NOT_FOUND: u8 = 5;
pub const ERR_OPENING_POLY_MISMATCH: u8 = 6;
pub const ERR_INVALID_NUMBER_OF_DIGESTS: u8 = 7;
pub const ERR_PAIRING_CHECK_FAILED: u8 = 8;
pub const PANIC_FIRST: u8 = 16;

// enum bn254v_error
pub const E_VK_PARSE: i32 = -4;

extern "C" {
    pub fn bn254v_init(devices: *const i32, n_devices: i32) -> i32;
    pub fn bn254v_shutdown();
    pub fn bn254v_device_count() -> i32;
    pub fn bn254v_last_error() -> *const core::ffi::c_char;
    pub fn bn254v_status_name(status: i32) -> *const core::ffi::c_char;
    pub fn bn254v_groth16_vk_load(vk: *const u8, len: usize, sign_mode: i32, out: *mut *mut bn254v_vk) -> i32;
    pub fn bn254v_plonk_vk_load(vk: *const u8, len: usize, out: *mut *mut bn254v_vk) -> i32;
    pub fn bn254v_vk_free(vk: *mut bn254v_vk);
    pub fn bn254v_vk_n_public(vk: *const bn254v_vk) -> i32;
    pub fn bn254v_groth16_verify_batch(
        vk: *const bn254v_vk, proofs: *const u8, proof_stride: usize, proof_len: *const u32, inputs_be: *const u8,
        n_inputs: i32, n: usize, status: *mut u8, dbg: *const bn254v_debug,
    ) -> i32;
    pub fn bn254v_groth16_batch_all_valid(
        vk: *const bn254v_vk, proofs: *const u8, proof_stride: usize, proof_len: *const u32, inputs_be: *const u8,
        n_inputs: i32, rnd16: *const u8, n: usize, all_valid: *mut u8, status: *mut u8,
    ) -> i32;
    pub fn bn254v_plonk_verify_batch(
        vk: *const bn254v_vk, proofs: *const u8, proof_stride: usize, proof_len: *const u32, inputs_be: *const u8,
        n_inputs: i32, rnd_be: *const u8, n: usize, status: *mut u8, dbg: *const bn254v_debug,
    ) -> i32;
    pub fn bn254v_pairing_product_batch(
        g1: *const u8, g2: *const u8, k: i32, n: usize, is_one: *mut u8, miller_out: *mut u8, gt_out: *mut u8,
    ) -> i32;
    pub fn bn254v_vk_cache_get(kind: i32, vk: *const u8, len: usize, sign_mode: i32, out: *mut *const bn254v_vk) -> i32;
    pub fn bn254v_vk_cache_size() -> usize;
    pub fn bn254v_vk_cache_clear();
    pub fn bn254v_verify_many(items: *const bn254v_item, n: usize, sign_mode: i32, rnd_be: *const u8, status: *mut u8) -> i32;
}
