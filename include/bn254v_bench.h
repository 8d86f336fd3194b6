/* bn254v_bench -- measurement and test support of libbn254v (NOT part of the verifier's drop-in surface).
 *
 * Device-resident batches for kernel-only timing (bench.py `value`: inputs already in HBM when the timed region
 * starts), per-stage device times, synthetic workload generators for BASELINE.json configs 2 and 4, the integer
 * multiply-add issue-rate probe that is the roofline denominator, and the launch counter.
 * The verifier itself is include/bn254v.h.
 */
#ifndef BN254V_BENCH_H
#define BN254V_BENCH_H

#include "bn254v.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- device-resident batches -------------------------------------------------------------------
 * Uploads once; *_batch_verify runs the verification kernels on the staged data without host copies.
 * *kernel_ms (nullable) receives the device time (CUDA events on the launching stream, max over devices);
 * status / is_one (nullable) is copied back after the timed region.                                  */
typedef struct bn254v_batch bn254v_batch;
int bn254v_groth16_batch_upload(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                const uint8_t* inputs_be, int n_inputs, size_t n, bn254v_batch** out);
int bn254v_groth16_batch_verify(const bn254v_vk* vk, bn254v_batch* batch, uint8_t* status, float* kernel_ms);
int bn254v_plonk_batch_upload(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                              const uint8_t* inputs_be, int n_inputs, const uint8_t* rnd_be, size_t n,
                              bn254v_batch** out);
int bn254v_plonk_batch_verify(const bn254v_vk* vk, bn254v_batch* batch, uint8_t* status, float* kernel_ms);
int bn254v_pairing_batch_upload(const uint8_t* g1, const uint8_t* g2, int k, size_t n, bn254v_batch** out);
int bn254v_pairing_batch_verify(bn254v_batch* batch, uint8_t* is_one, float* kernel_ms);
void bn254v_batch_free(bn254v_batch* batch);

/* Device times (CUDA events, device slot 0) of the stages of the last *_batch_verify call, in milliseconds:
 *   Groth16: [0] Miller-loop launch, [1] final-exponentiation launch (0 when the batch ran as one fused launch)
 *   PlonK:   [0] stage A, [1] MSM terms 0, [2] stage C, [3] MSM terms 1, [4] stage E (pairing), summed over chunks
 *   pairing: [0] the kernel
 * Returns the number of values written (<= cap).                                                          */
int bn254v_last_stage_ms(float* out, int cap);
/* Groth16 only (kept from round 1): the two values above. */
int bn254v_last_kernel_split(float* miller_ms, float* finish_ms);

/* ---- synthetic workloads (BASELINE.json configs 2 and 4; generated on device) ------------------
 * Trapdoor-simulated Groth16 instance set: writes the gnark VK bytes (*vk_len in: capacity, out:
 * length), n proofs of 256 bytes and n * n_public * 32 input bytes; 50 % of the proofs are
 * corrupted (expected[i] = BN254V_OK_TRUE or BN254V_OK_FALSE).  Same PRNG definition as the
 * oracle's generator (oracle/bn254_oracle.py Groth16Trapdoor) so both sides can be compared.     */
int bn254v_groth16_synth(uint64_t seed, int n_public, int sign_mode, size_t first_index, size_t n,
                         uint8_t* vk_bytes, size_t* vk_len, uint8_t* proofs, uint8_t* inputs_be,
                         uint8_t* expected);
/* Random k-pair sets P_j = s_j G1, Q_j = t_j G2; odd-indexed sets are solved so the product is 1. */
int bn254v_pairing_synth(uint64_t seed, int k, size_t first_index, size_t n, uint8_t* g1, uint8_t* g2,
                         uint8_t* expected_is_one);

/* Dependent-free IMAD.WIDE.U32 stream on device 0: returns achieved multiply-adds per second
 * (the int32 roofline denominator).  iters >= 1.                                 */
int bn254v_imad_peak(int iters, double* wide_mac_per_s, double* lo_mac_per_s, float* sm_clock_mhz);
/* Test support: the host half of bn254v_groth16_batch_all_valid -- s = sum r_i and t_j = sum r_i x_ij (mod r) of m proofs
 * as (1 + n_inputs) big-endian 32-byte scalars (r_i from the proof's 16 scalar bytes; csrc/groth16_agg.cuh).  No device. */
void bn254v_agg_host_sums(const uint8_t* rnd16, const uint8_t* inputs_be, int n_inputs, size_t m, uint8_t* scal_be);
/* Test support: the ChaCha20 block function (RFC 8439 2.3) and the key stream (nonce 0, counter from 0) with which the
 * library expands one getrandom(2) seed into the per-proof scalars of bn254v_groth16_batch_all_valid.  No device.     */
void bn254v_chacha20_block(const uint8_t* key32, uint32_t counter, const uint8_t* nonce12, uint8_t* out64);
void bn254v_chacha20_expand(const uint8_t* key32, size_t n, uint8_t* out);
/* Number of kernel launches issued by this library since init (for bench.py's gpu_launches).    */
uint64_t bn254v_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BN254V_BENCH_H */
