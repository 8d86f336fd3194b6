/* Plain C (gcc) driver of the bn254v C ABI: the reference-side binding a maintainer adds is a thin FFI over exactly
 * these calls, so this program is the language-neutral proof that the boundary works without
 * Python / ctypes.  TEST PROGRAM (tests/c_abi), not part of the product.
 *
 *   abi_driver                 -> without a CUDA device: every compute entry point fails with BN254V_E_NO_DEVICE
 *                                 (prints "NO-DEVICE-OK"); with a device: runs the checks below, prints "C-ABI-OK".
 * Checks (device): trapdoor Groth16 workload (bn254v_bench.h generator) through bn254v_vk_cache_get +
 * bn254v_groth16_verify_batch, statuses equal to the generator's expectation; the same proofs as a mixed batch over two
 * VKs through bn254v_verify_many (a proof under the other VK is rejected, an unparsable VK gives PANIC_VK_PARSE);
 * ragged records; a pairing-product set with an identity pair.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bn254v.h"
#include "bn254v_bench.h"

#define CHECK(cond)                                                                  \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, bn254v_last_error()); \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int main(void) {
  int rc = bn254v_init(NULL, 0);
  if (rc == BN254V_E_NO_DEVICE) {
    uint8_t vk[64] = {0}, st[1];
    bn254v_vk* h = NULL;
    CHECK(bn254v_groth16_vk_load(vk, sizeof vk, 0, &h) == BN254V_E_NO_DEVICE);
    CHECK(bn254v_pairing_product_batch(vk, vk, 1, 1, st, NULL, NULL) == BN254V_E_NO_DEVICE);
    CHECK(strcmp(bn254v_status_name(BN254V_PANIC_VK_PARSE), "PANIC_VK_PARSE") == 0);
    printf("NO-DEVICE-OK\n");
    return 0;
  }
  CHECK(rc == BN254V_SUCCESS);
  CHECK(bn254v_device_count() >= 1);

  enum { N = 300 };
  static uint8_t vk1[4096], vk2[4096], proofs[N * 256], inputs[N * 64], expected[N], status[N + 8];
  size_t vk1_len = sizeof vk1, vk2_len = sizeof vk2;
  CHECK(bn254v_groth16_synth(11, 2, 0, 0, N, vk1, &vk1_len, proofs, inputs, expected) == 0);
  CHECK(bn254v_groth16_synth(12, 2, 0, 0, 0, vk2, &vk2_len, NULL, NULL, NULL) == 0);

  const bn254v_vk* h1 = NULL;
  const bn254v_vk* h1b = NULL;
  CHECK(bn254v_vk_cache_get(BN254V_KIND_GROTH16, vk1, vk1_len, 0, &h1) == 0);
  CHECK(bn254v_vk_cache_get(BN254V_KIND_GROTH16, vk1, vk1_len, 0, &h1b) == 0);
  CHECK(h1 == h1b && bn254v_vk_cache_size() == 1 && bn254v_vk_n_public(h1) == 2);
  CHECK(bn254v_groth16_verify_batch(h1, proofs, 256, NULL, inputs, 2, N, status, NULL) == 0);
  int n_true = 0;
  for (int i = 0; i < N; i++) {
    CHECK(status[i] == expected[i]);
    n_true += status[i] == BN254V_OK_TRUE;
  }
  CHECK(n_true == N / 2);

  /* opt-in aggregate check: yes for the valid half packed together, no for the whole (half-corrupted) batch */
  {
    static uint8_t vproofs[N * 256], vinputs[N * 64];
    int nv = 0;
    for (int i = 0; i < N; i++)
      if (expected[i] == BN254V_OK_TRUE) {
        memcpy(vproofs + 256 * nv, proofs + 256 * i, 256);
        memcpy(vinputs + 64 * nv, inputs + 64 * i, 64);
        nv++;
      }
    uint8_t yes = 7, no = 7;
    CHECK(bn254v_groth16_batch_all_valid(h1, vproofs, 256, NULL, vinputs, 2, NULL, nv, &yes, NULL) == 0);
    CHECK(bn254v_groth16_batch_all_valid(h1, proofs, 256, NULL, inputs, 2, NULL, N, &no, status) == 0);
    CHECK(yes == 1 && no == 0);
    for (int i = 0; i < N; i++) CHECK(status[i] == BN254V_OK_TRUE); /* well-formed, in the aggregate: not a verdict */
  }

  /* ragged records: the second one is cut short */
  uint32_t lens[3] = {256, 100, 256};
  CHECK(bn254v_groth16_verify_batch(h1, proofs, 256, lens, inputs, 2, 3, status, NULL) == 0);
  CHECK(status[0] == expected[0] && status[1] == BN254V_PANIC_SHORT_BUFFER && status[2] == expected[2]);

  /* mixed batch: items alternate between the two VKs; an item with an unparsable VK; wrong input count */
  enum { M = 64 };
  bn254v_item items[M + 2];
  int first_valid = -1;
  for (int i = 0; i < M; i++) {
    items[i].kind = BN254V_KIND_GROTH16;
    items[i].n_inputs = 2;
    items[i].proof = proofs + 256 * i, items[i].proof_len = 256;
    items[i].vk = (i & 1) ? vk2 : vk1, items[i].vk_len = (i & 1) ? vk2_len : vk1_len;
    items[i].inputs_be = inputs + 64 * i;
    if (first_valid < 0 && expected[i] == BN254V_OK_TRUE) first_valid = i;
  }
  items[M] = items[first_valid];
  items[M].vk_len = 100; /* truncated VK */
  items[M + 1] = items[first_valid];
  items[M + 1].n_inputs = 1;
  CHECK(bn254v_verify_many(items, M + 2, 0, NULL, status) == 0);
  for (int i = 0; i < M; i++) CHECK(status[i] == ((i & 1) ? BN254V_OK_FALSE : expected[i]));
  CHECK(status[M] == BN254V_PANIC_VK_PARSE && status[M + 1] == BN254V_ERR_PREPARE_INPUTS);
  CHECK(bn254v_vk_cache_size() == 2);

  /* pairing products: e(P, Q) e(-P, Q) = 1 is not expressible without negation here; use the generator's solved sets
   * and mask one pair of an unsolved set as the identity: the set {identity} alone is 1 */
  static uint8_t g1[2 * 64], g2[2 * 128], one[1], exp1[1];
  CHECK(bn254v_pairing_synth(5, 2, 1, 1, g1, g2, exp1) == 0); /* index 1: solved, product is 1 */
  CHECK(bn254v_pairing_product_batch(g1, g2, 2, 1, one, NULL, NULL) == 0);
  CHECK(one[0] == 1 && exp1[0] == 1);
  memset(g1, 0, 64); /* first pair -> identity: e(P2, Q2) alone is not 1 */
  CHECK(bn254v_pairing_product_batch(g1, g2, 2, 1, one, NULL, NULL) == 0);
  CHECK(one[0] == 0);
  memset(g2 + 128, 0, 128); /* second pair -> identity as well: empty product */
  CHECK(bn254v_pairing_product_batch(g1, g2, 2, 1, one, NULL, NULL) == 0);
  CHECK(one[0] == 1);

  bn254v_vk_cache_clear();
  CHECK(bn254v_vk_cache_size() == 0);
  bn254v_shutdown();
  printf("C-ABI-OK\n");
  return 0;
}
