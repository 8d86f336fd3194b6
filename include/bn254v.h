/* bn254v -- B200-native batched BN254 Groth16 / PlonK verifier: C ABI.
 *
 * This is the drop-in boundary for the hot path of succinctlabs/snark-bn254-verifier.  The
 * reference has no FFI of its own; its entry points are two Rust associated functions
 *
 *     Groth16Verifier::verify(proof:&[u8], vk:&[u8], public_inputs:&[Fr]) -> Result<bool,Groth16Error>
 *                                                                 (verifier/src/lib.rs:44-49)
 *     PlonkVerifier::verify  (proof:&[u8], vk:&[u8], public_inputs:&[Fr]) -> Result<bool,PlonkError>
 *                                                                 (verifier/src/lib.rs:69-74)
 *
 * A Rust shim binds the functions below (INTEGRATION.md shows the `extern "C"` block and the
 * `verify` / `verify_batch` wrappers).  Every function takes plain pointers and sizes; all byte
 * strings are in the gnark wire format the reference parses (big-endian field elements).
 *
 * Per-proof outcome: one status byte (enum bn254v_status) that encodes the three things the
 * reference can do -- return Ok(bool), return Err(..), or panic (`unwrap` on a parser error).
 * Function return value: 0 on success, a negative bn254v_error otherwise (library-level failure:
 * bad arguments, CUDA error, no device).  There is NO CPU fallback: without a CUDA device every
 * compute entry point returns BN254V_E_NO_DEVICE.
 *
 * Threading: handles are immutable after load; one batch call per device at a time.
 */
#ifndef BN254V_H
#define BN254V_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- per-proof status ------------------------------------------------------------------- */
enum bn254v_status {
  BN254V_OK_TRUE = 0,                /* Ok(true)                                                     */
  BN254V_OK_FALSE = 1,               /* Ok(false)         groth16/verify.rs:73-77                    */
  BN254V_ERR_PREPARE_INPUTS = 2,     /* Err(PrepareInputsFailed)          groth16/verify.rs:54-56    */
  BN254V_ERR_BSB22_MISMATCH = 3,     /* Err(Bsb22CommitmentMismatch)      plonk/verify.rs:52-54      */
  BN254V_ERR_INVALID_WITNESS = 4,    /* Err(InvalidWitness)               plonk/verify.rs:57-59      */
  BN254V_ERR_INVERSE_NOT_FOUND = 5,  /* Err(InverseNotFound)              plonk/verify.rs:106        */
  BN254V_ERR_OPENING_POLY_MISMATCH = 6, /* Err(OpeningPolyMismatch)       plonk/verify.rs:212-214    */
  BN254V_ERR_INVALID_NUMBER_OF_DIGESTS = 7, /* Err(InvalidNumberOfDigests) plonk/kzg.rs:95-97       */
  BN254V_ERR_PAIRING_CHECK_FAILED = 8, /* Err(PairingCheckFailed)         plonk/kzg.rs:185-187       */
  BN254V_PANIC_FIELD_NOT_MEMBER = 16, /* Fq/Fr::from_slice >= modulus, unwrap panics  converter.rs:85 */
  BN254V_PANIC_NOT_ON_CURVE = 17,    /* AffineG::new -> NotOnCurve, unwrap panics     converter.rs:87 */
  BN254V_PANIC_NOT_IN_SUBGROUP = 18, /* AffineG2::new -> NotInSubgroup                converter.rs:152*/
  BN254V_PANIC_IDENTITY = 19,        /* bn: "Unable to convert G1 to AffineG1" (identity intermediate)*/
  BN254V_PANIC_SHORT_BUFFER = 20,    /* slice index out of range                groth16/converter.rs:15 */
  BN254V_PANIC_DIV_BY_ZERO = 21,     /* Fr `/=` by zero                          plonk/verify.rs:157  */
  BN254V_PANIC_INDEX_OUT_OF_RANGE = 22, /* claimed_values[1..5] missing           plonk/verify.rs:166-170 */
  BN254V_PANIC_VK_PARSE = 23,        /* VK parser error unwrapped (verify_many items)  lib.rs:46,71       */
  BN254V_STATUS_UNSET = 255
};

/* ---- library-level errors ---------------------------------------------------------------- */
enum bn254v_error {
  BN254V_SUCCESS = 0,
  BN254V_E_BAD_ARG = -1,
  BN254V_E_NO_DEVICE = -2,     /* no CUDA device / driver: the library never falls back to the CPU */
  BN254V_E_CUDA = -3,          /* a CUDA call failed; see bn254v_last_error()                      */
  BN254V_E_VK_PARSE = -4,      /* VK bytes malformed (the reference would panic in the VK parser)  */
  BN254V_E_UNSUPPORTED = -5    /* VK shape outside compiled limits (n_public, n_qcp)               */
};

typedef struct bn254v_vk bn254v_vk; /* opaque verifying-key handle (host struct + per-device tables) */

/* Optional per-proof debug outputs (canonical, big-endian); any pointer may be NULL.
 * Used by the parity tests to compare MSM outputs and Fq12 values bit-exactly with the oracle. */
typedef struct bn254v_debug {
  uint8_t* g1_out;     /* n * n_g1_out * 64: Groth16: L (prepare_inputs);  PlonK: lin digest, folded digest,
                          pairing G1 #0, pairing G1 #1 */
  uint8_t* fr_out;     /* n * n_fr_out * 32: PlonK: gamma, beta, alpha, zeta, kzg gamma, PI, const_lin, hashed BSB22[0] */
  uint8_t* miller_out; /* n * 384: canonical Fq12 Miller value                                   */
  uint8_t* gt_out;     /* n * 384: canonical Fq12 after final exponentiation                     */
} bn254v_debug;

#define BN254V_GROTH16_N_G1_OUT 1
#define BN254V_PLONK_N_G1_OUT 4
#define BN254V_PLONK_N_FR_OUT 8

/* ---- lifecycle ----------------------------------------------------------------------------- */
/* Selects devices (NULL/0 = all visible).  Replaces nothing in the reference (it has no state). */
int bn254v_init(const int* devices, int n_devices);
void bn254v_shutdown(void);
int bn254v_device_count(void);           /* devices selected by bn254v_init */
const char* bn254v_last_error(void);
const char* bn254v_status_name(int status);

/* ---- verifying keys --------------------------------------------------------------------------
 * Host parses the gnark framing and decompresses points (replaces
 * load_groth16_verifying_key_from_bytes verifier/src/groth16/converter.rs:28-89 and
 * load_plonk_verifying_key_from_bytes verifier/src/plonk/converter.rs:18-119), then uploads the
 * VK and precomputes VK-constant data on each device (G2 line tables, e(alpha,-beta), SHA midstate).
 * sign_mode 0 = the reference's equation e(A,B) e(L,gamma) e(C,-delta) == e(alpha,-beta_file);
 * sign_mode 1 = gnark's   e(A,B) e(L,-gamma) e(C,-delta) == e(alpha, beta_file).                */
int bn254v_groth16_vk_load(const uint8_t* vk_bytes, size_t len, int sign_mode, bn254v_vk** out);
int bn254v_plonk_vk_load(const uint8_t* vk_bytes, size_t len, bn254v_vk** out);
void bn254v_vk_free(bn254v_vk* vk);
int bn254v_vk_n_public(const bn254v_vk* vk);

/* ---- batch verification ------------------------------------------------------------------------
 * proofs: n records of `proof_stride` bytes, each starting with the gnark raw proof
 *   Groth16: A.x|A.y | B.x1|B.x0|B.y1|B.y0 | C.x|C.y  (>= 256 bytes used; groth16/converter.rs:14-26)
 *   PlonK:   layout of plonk/converter.rs:121-178 (904 bytes for the SP1 v2.0.0 circuit shape)
 * proof_len[i] (nullable) = valid bytes of record i (a short record reproduces the reference's
 *   slice-index panic as BN254V_PANIC_SHORT_BUFFER); NULL means every record is proof_stride long.
 * inputs_be: n * n_inputs * 32 bytes, big-endian Fr (must be < r, else PANIC_FIELD_NOT_MEMBER --
 *   the reference's caller builds them with Fr::from_slice).
 * Proofs are sharded by index over the devices chosen at init; the call is synchronous.          */
int bn254v_groth16_verify_batch(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs,
                                size_t n, uint8_t* status, const bn254v_debug* dbg);

/* OPT-IN aggregate check (no counterpart in the reference, which verifies proof by proof,
 * verifier/src/groth16/verify.rs:65-78): "is EVERY proof of the batch valid?", for one single-pair Miller loop per
 * proof and ONE three-pair Miller loop + final exponentiation per batch instead of a full pairing check per proof --
 * about half the work when the answer is yes.  It CHANGES SEMANTICS and is never used by bn254v_groth16_verify_batch:
 *   *all_valid = 1  <=>  every record is well-formed (same checks as the per-proof path: lengths, coordinates < p,
 *                        points on their curves, B in G2, public inputs < r and != 0) AND the random linear
 *                        combination  prod_i [e(A_i,B_i) e(L_i,gamma') e(C_i,delta') / e(alpha,beta')]^(r_i) == 1  holds.
 *                        If all proofs are valid this always holds; if one is not, it holds with probability
 *                        <= 2^-126 over the library's choice of the r_i (2^127 values each).
 *   *all_valid = 0       says nothing about WHICH proof fails: call bn254v_groth16_verify_batch to find out.
 * status (nullable): per record, the per-proof path's status if the record is malformed (PANIC_* / ERR_PREPARE_INPUTS),
 *   BN254V_OK_TRUE if it is well-formed and went into the aggregate -- NOT a verdict on that proof.  Not reproduced:
 *   substrate-bn's panic on an identity PARTIAL sum inside prepare_inputs (needs a discrete-log relation in the VK).
 * rnd16: NULL in production (AFTER it has the proofs the library draws one 32-byte seed from getrandom(2) and expands it
 *   with ChaCha20 into 16 bytes per proof); n * 16 caller-supplied bytes exist for reproducible tests only -- scalars
 *   known to the prover void the check.
 * Each device checks its own shard; the answer is the AND.                                          */
int bn254v_groth16_batch_all_valid(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                                   const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs,
                                   const uint8_t* rnd16, size_t n, uint8_t* all_valid, uint8_t* status);

/* rnd_be: the per-proof scalar the reference draws from OsRng inside kzg::batch_verify_multi_points
 * (verifier/src/plonk/kzg.rs:149-154).
 *   NULL (the production path): for every call the library draws a fresh 32-byte seed from the operating system's
 *     CSPRNG (getrandom(2)) and expands it with ChaCha20 into n 32-byte scalars, redrawing any that is 0 mod r,
 *     exactly where the reference calls Fr::random(OsRng).
 *   non-NULL (tests / reproducible runs only): n * 32 bytes, reduced mod r on device.  SECURITY: the scalar separates
 *     the two openings of the batched KZG check; if it is 0 mod r, or known to the prover before the proof is fixed
 *     (a constant, a reused array, a seeded PRNG), the z-shifted opening Z(omega zeta) = zu is effectively unchecked
 *     and proofs can be forged.  Never pass caller-chosen values in production. */
int bn254v_plonk_verify_batch(const bn254v_vk* vk, const uint8_t* proofs, size_t proof_stride,
                              const uint32_t* proof_len, const uint8_t* inputs_be, int n_inputs,
                              const uint8_t* rnd_be, size_t n, uint8_t* status, const bn254v_debug* dbg);

/* Raw k-pair pairing-product check (bn::pairing_batch, k <= 4): g1 n*k*64 (x|y), g2 n*k*128
 * (x1|x0|y1|y0).  Points are trusted (no validation: bn::pairing_batch takes group elements).  The identity is
 * encoded as all-zero bytes (64 for G1, 128 for G2); a pair with an identity member is skipped, as
 * substrate-bn's pairing_batch skips it (a set of skipped pairs only has Miller value 1).
 * is_one[i] = 1 iff the product of pairings is the identity of GT.                              */
int bn254v_pairing_product_batch(const uint8_t* g1, const uint8_t* g2, int k, size_t n,
                                 uint8_t* is_one, uint8_t* miller_out, uint8_t* gt_out);

/* ---- VK cache and mixed batches ------------------------------------------------------------------
 * The reference parses the VK on every call (verifier/src/lib.rs:46,71 -> groth16/converter.rs:28-89,
 * plonk/converter.rs:18-119).  Here a VK costs a host decompression plus device precomputation (line tables,
 * e(alpha,beta'), window tables), so the library keeps the handles it has built, keyed by
 * sha256(vk bytes) -- SP1's *_vkey_hash -- kind and sign_mode.  A cached handle is owned by the library: do not
 * pass it to bn254v_vk_free; bn254v_vk_cache_clear / bn254v_shutdown release them.                          */
enum bn254v_kind { BN254V_KIND_GROTH16 = 0, BN254V_KIND_PLONK = 1 };
int bn254v_vk_cache_get(int kind, const uint8_t* vk_bytes, size_t len, int sign_mode, const bn254v_vk** out);
size_t bn254v_vk_cache_size(void);
void bn254v_vk_cache_clear(void);

/* One item of a mixed batch: what one call of Groth16Verifier::verify / PlonkVerifier::verify receives
 * (verifier/src/lib.rs:44,69).  Items that share (kind, vk bytes, n_inputs) are verified as one device batch. */
typedef struct bn254v_item {
  int kind;                 /* enum bn254v_kind                                             */
  int n_inputs;             /* number of public inputs                                      */
  const uint8_t* proof;     /* gnark raw proof                                              */
  size_t proof_len;
  const uint8_t* vk;        /* gnark VK bytes (identical pointers are recognised without hashing) */
  size_t vk_len;
  const uint8_t* inputs_be; /* n_inputs * 32 bytes, big-endian Fr                           */
} bn254v_item;
/* Verifies n items over any number of verifying keys and both proof systems; status[i] belongs to items[i].
 * A VK that does not parse gives BN254V_PANIC_VK_PARSE for its items (the reference unwraps the VK parser).
 * rnd_be: NULL (production) or n * 32 bytes, used by the PlonK items only (see bn254v_plonk_verify_batch).    */
int bn254v_verify_many(const bn254v_item* items, size_t n, int sign_mode, const uint8_t* rnd_be, uint8_t* status);

#ifdef __cplusplus
}
#endif
#endif /* BN254V_H */
