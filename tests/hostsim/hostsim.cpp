// Host build of the kernels' __host__ __device__ code (carry flag emulated, see field.cuh).
// TEST-ONLY: lets the CPU test-suite exercise the exact per-proof routines the CUDA kernels run,
// against the oracle, in a container without a GPU.  Not part of the product library.
#include <string.h>

#include <vector>

#include "../../snark-bn254-verifier_b200/csrc/groth16.cuh"

using namespace bn254;

extern "C" {

void hs_fp_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  Fp x, y;
  memcpy(x.v, a, 32);
  memcpy(y.v, b, 32);
  Fp z = fe_mul(x, y);
  memcpy(r, z.v, 32);
}
// the dedicated squaring of the G1 arithmetic (fe_sqr_short) in both fields; a may be any value below 2 m
void hs_fe_sqr_short(int which, uint32_t* r, const uint32_t* a) {
  if (which == 0) {
    Fp x;
    memcpy(x.v, a, 32);
    Fp z = fe_sqr_short(x);
    memcpy(r, z.v, 32);
  } else {
    Fr x;
    memcpy(x.v, a, 32);
    Fr z = fe_sqr_short(x);
    memcpy(r, z.v, 32);
  }
}
// the plain 512-bit square under fe_sqr_short (any 256-bit a): 16 LE words
void hs_fe_sqr_wide(uint32_t* t16, const uint32_t* a) {
  Fp x;
  memcpy(x.v, a, 32);
  fe_sqr_wide(t16, x);
}
// 1 / a by divsteps (fe_inv) and by the Fermat chain (fe_inv_fermat); which: 0 = Fq, 1 = Fr.  Montgomery words in and out.
void hs_fe_inv(int which, const uint32_t* a, uint32_t* inv_divsteps, uint32_t* inv_fermat) {
  if (which == 0) {
    Fp x;
    memcpy(x.v, a, 32);
    Fp y = fe_inv(x), z = fe_inv_fermat(x);
    memcpy(inv_divsteps, y.v, 32), memcpy(inv_fermat, z.v, 32);
  } else {
    Fr x;
    memcpy(x.v, a, 32);
    Fr y = fe_inv(x), z = fe_inv_fermat(x);
    memcpy(inv_divsteps, y.v, 32), memcpy(inv_fermat, z.v, 32);
  }
}
void hs_fr_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  Fr x, y;
  memcpy(x.v, a, 32);
  memcpy(y.v, b, 32);
  Fr z = fe_mul(x, y);
  memcpy(r, z.v, 32);
}

int hs_pairing_product(int k, const uint8_t* g1, const uint8_t* g2, uint8_t* miller_out, uint8_t* gt_out) {
  switch (k) {
    case 1: return pairing_product_one<1>(g1, g2, miller_out, gt_out);
    case 2: return pairing_product_one<2>(g1, g2, miller_out, gt_out);
    case 3: return pairing_product_one<3>(g1, g2, miller_out, gt_out);
    case 4: return pairing_product_one<4>(g1, g2, miller_out, gt_out);
  }
  return -1;
}

// final exponentiation of a canonical Fq12 value (work counting)
void hs_final_exp_only(const uint8_t* f_be) {
  Fp12 f;
  Fp2* cs[6] = {&f.c0.c0, &f.c0.c1, &f.c0.c2, &f.c1.c0, &f.c1.c1, &f.c1.c2};
  for (int i = 0; i < 6; i++) {
    fp_load_be(cs[i]->c0, f_be + 64 * i);
    fp_load_be(cs[i]->c1, f_be + 64 * i + 32);
  }
  unsigned long long before = 0;
#ifdef BN254_COUNT_MULS
  before = fe_mac_counter();
  fe_mac_counter() = 0;
#endif
  final_exponentiation(f, f);
  (void)before;
}
int hs_g2_check(const uint8_t* g2) {
  G2Aff q;
  return load_g2_checked(q, g2);
}
// the three subgroup predicates on an (unchecked) G2 point: bit 0 = 63-bit test, bit 1 = 6x^2 test, bit 2 = end point
// of the Miller loop's point chain
int hs_g2_subgroup_both(const uint8_t* g2) {
  G2Aff q;
  load_g2_unchecked(q, g2);
  return (g2_in_subgroup<false>(q) ? 1 : 0) | (g2_in_subgroup_6x2(q) ? 2 : 0) | (g2_in_subgroup_ate(q) ? 4 : 0);
}
int hs_g1_check(const uint8_t* g1) {
  G1Aff p;
  return load_g1_checked(p, g1);
}

// vk_points: alpha(64) | beta'(128) | gamma'(128) | delta'(128) | ic[n_ic](64 each), uncompressed BE
void* hs_groth16_vk_new(const uint8_t* vk_points, int n_ic) {
  Groth16VkDev* vk = new Groth16VkDev();
  vk->n_ic = n_ic;
  load_g1_unchecked(vk->alpha, vk_points);
  load_g2_unchecked(vk->beta, vk_points + 64);
  load_g2_unchecked(vk->gamma, vk_points + 192);
  load_g2_unchecked(vk->delta, vk_points + 320);
  for (int i = 0; i < n_ic; i++) load_g1_unchecked(vk->ic[i], vk_points + 448 + 64 * i);
  groth16_vk_prepare(*vk);
  vk->ic_table = nullptr;
  return vk;
}
// same, with the fixed-base IC tables the CUDA vk_load builds
void* hs_groth16_vk_new_tables(const uint8_t* vk_points, int n_ic) {
  Groth16VkDev* vk = (Groth16VkDev*)hs_groth16_vk_new(vk_points, n_ic);
  G1Aff* tab = new G1Aff[(size_t)(n_ic - 1) * BN_IC_WINDOWS * BN_IC_ENTRIES];
  for (int b = 0; b + 1 < n_ic; b++)
    for (int w = 0; w < BN_IC_WINDOWS; w++)
      groth16_ic_table_slice(tab + ((size_t)b * BN_IC_WINDOWS + w) * BN_IC_ENTRIES, vk->ic[b + 1], w);
  vk->ic_table = tab;
  return vk;
}
void hs_groth16_vk_free(void* vk) {
  delete[] ((Groth16VkDev*)vk)->ic_table;
  delete (Groth16VkDev*)vk;
}
// multiply-adds of the part of a proof that k_groth16_prepare runs (decode, validation, prepare_inputs)
unsigned long long hs_groth16_prepare_macs(void* vk, const uint8_t* proof, uint32_t proof_len, const uint8_t* inputs, int n_inputs) {
#ifdef BN254_COUNT_MULS
  const unsigned long long before = fe_mac_counter();
  G1Aff A, C, L;
  G2Aff B;
  groth16_parse_one(A, B, C, L, *(Groth16VkDev*)vk, proof, proof_len, inputs, n_inputs);
  return fe_mac_counter() - before;
#else
  return 0;
#endif
}
void hs_groth16_vk_target(void* vk, uint8_t* out) { fp12_to_bytes(out, ((Groth16VkDev*)vk)->target); }

int hs_groth16_verify(void* vk, const uint8_t* proof, uint32_t proof_len, const uint8_t* inputs, int n_inputs,
                      uint8_t* L, uint8_t* miller, uint8_t* gt) {
  Groth16Debug dbg{L, miller, gt};
  return groth16_verify_one(*(Groth16VkDev*)vk, proof, proof_len, inputs, n_inputs, dbg);
}
}

// ---- PlonK (host build of plonk.cuh) ------------------------------------------------------------
#include "../../snark-bn254-verifier_b200/csrc/gnark_host.h"
#include "../../snark-bn254-verifier_b200/csrc/plonk.cuh"

extern "C" {
void* hs_plonk_vk_new(const uint8_t* vk_bytes, size_t len) {
  gnark::PlonkVkHost h;
  if (gnark::parse_plonk_vk(h, vk_bytes, len)) return nullptr;
  PlonkVkDev* hv = new PlonkVkDev();
  memset(hv, 0, sizeof *hv);
  hv->size = h.size;
  hv->n_public = (int)h.nb_public;
  hv->n_qcp = (int)h.qcp.size();
  hv->size_inv = fe_to_mont(h.size_inv);
  hv->generator = fe_to_mont(h.generator);
  hv->coset_shift = fe_to_mont(h.coset_shift);
  for (int i = 0; i < hv->n_qcp; i++) hv->w_pow_cci[i] = fr_pow_u64(hv->generator, h.nb_public + h.cci[i]);
  for (int i = 0; i < 3; i++) hv->s[i] = h.s[i];
  hv->ql = h.ql, hv->qr = h.qr, hv->qm = h.qm, hv->qo = h.qo, hv->qk = h.qk, hv->g1 = h.g1;
  hv->g2[0] = h.g2[0], hv->g2[1] = h.g2[1];
  for (int i = 0; i < hv->n_qcp; i++) hv->qcp[i] = h.qcp[i];
  sha256_init(hv->gamma_prefix);
  sha_bytes(hv->gamma_prefix, "gamma", 5);
  const G1Aff* pts[8] = {&hv->s[0], &hv->s[1], &hv->s[2], &hv->ql, &hv->qr, &hv->qm, &hv->qo, &hv->qk};
  uint8_t b[64];
  for (int i = 0; i < 8; i++) {
    store_g1(b, *pts[i]);
    sha256_update(hv->gamma_prefix, b, 64);
  }
  for (int i = 0; i < hv->n_qcp; i++) {
    store_g1(b, hv->qcp[i]);
    sha256_update(hv->gamma_prefix, b, 64);
  }
  store_g1(hv->kzg_vk_bytes, hv->s[0]);
  store_g1(hv->kzg_vk_bytes + 64, hv->s[1]);
  for (int i = 0; i < hv->n_qcp; i++) store_g1(hv->kzg_vk_bytes + 128 + 64 * i, hv->qcp[i]);
  plonk_vk_prepare(*hv);
  return hv;
}
// attach the fixed-base window tables the CUDA vk_load builds (slow on the host: ~75 k affine conversions)
void hs_plonk_vk_add_tables(void* vkp) {
  PlonkVkDev* vk = (PlonkVkDev*)vkp;
  const int nf = BN_PLONK_N_FIXED(vk->n_qcp);
  G1Aff* tab = new G1Aff[(size_t)nf * BN_IC_WINDOWS * BN_IC_ENTRIES];
  for (int b = 0; b < nf; b++)
    for (int w = 0; w < BN_IC_WINDOWS; w++)
      groth16_ic_table_slice(tab + ((size_t)b * BN_IC_WINDOWS + w) * BN_IC_ENTRIES, plonk_fixed_base(*vk, b), w);
  vk->fixed_tables = tab;
}
void hs_plonk_vk_free(void* vk) {
  delete[] ((PlonkVkDev*)vk)->fixed_tables;
  delete (PlonkVkDev*)vk;
}
int hs_plonk_verify(void* vk, const uint8_t* proof, uint32_t len, const uint8_t* inputs, int n_inputs, const uint8_t* rnd,
                    uint8_t* g1, uint8_t* fr, uint8_t* miller, uint8_t* gt) {
  PlonkDebug dbg{g1, fr, miller, gt};
  return plonk_verify_one(*(PlonkVkDev*)vk, proof, len, inputs, n_inputs, rnd, dbg);
}
// the same with the MSM terms of each stage evaluated jointly (shared doublings: the large-batch form)
int hs_plonk_verify_joint(void* vk, const uint8_t* proof, uint32_t len, const uint8_t* inputs, int n_inputs,
                          const uint8_t* rnd, uint8_t* g1, uint8_t* fr, uint8_t* miller, uint8_t* gt) {
  PlonkDebug dbg{g1, fr, miller, gt};
  return plonk_verify_one(*(PlonkVkDev*)vk, proof, len, inputs, n_inputs, rnd, dbg, true);
}
// limb multiply-adds of each stage of the staged PlonK path (the five kernels of k_plonk.cu): out[0..4] = stage A,
// terms 0, stage C, terms 1, stage E.  Returns the final status (a proof rejected in stage A leaves out[1..4] = 0).
int hs_plonk_stage_macs_form(void* vkp, const uint8_t* pr, uint32_t len, const uint8_t* inputs, int n_inputs,
                             const uint8_t* rnd, unsigned long long* out, int joint);
int hs_plonk_stage_macs(void* vkp, const uint8_t* pr, uint32_t len, const uint8_t* inputs, int n_inputs,
                        const uint8_t* rnd, unsigned long long* out) {
  return hs_plonk_stage_macs_form(vkp, pr, len, inputs, n_inputs, rnd, out, 0);
}
// joint != 0: the large-batch form (MSM terms of a sum evaluated jointly)
int hs_plonk_stage_macs_form(void* vkp, const uint8_t* pr, uint32_t len, const uint8_t* inputs, int n_inputs,
                             const uint8_t* rnd, unsigned long long* out, int joint) {
  const PlonkVkDev& vk = *(PlonkVkDev*)vkp;
  PlonkDebug dbg{nullptr, nullptr, nullptr, nullptr};
  for (int i = 0; i < 5; i++) out[i] = 0;
#ifdef BN254_COUNT_MULS
  PlonkWork w;
  fe_mac_counter() = 0;
  int st = plonk_stage_a(w, vk, pr, len, inputs, n_inputs, dbg);
  out[0] = fe_mac_counter(), fe_mac_counter() = 0;
  if (st != BN254V_OK_TRUE) return st;
  for (int t = 0; t < plonk_n_items(vk.n_qcp, 0, joint != 0); t++) {
    if (joint) plonk_item_joint(w, vk, pr, 0, t);
    else plonk_term(w, vk, pr, 0, t);
  }
  out[1] = fe_mac_counter(), fe_mac_counter() = 0;
  st = plonk_stage_c(w, vk, pr, rnd, dbg, joint != 0);
  out[2] = fe_mac_counter(), fe_mac_counter() = 0;
  if (st != BN254V_OK_TRUE) return st;
  for (int t = 0; t < plonk_n_items(vk.n_qcp, 1, joint != 0); t++) {
    if (joint) plonk_item_joint(w, vk, pr, 1, t);
    else plonk_term(w, vk, pr, 1, t);
  }
  out[3] = fe_mac_counter(), fe_mac_counter() = 0;
  st = plonk_stage_e(w, vk, pr, dbg, true, joint != 0);
  out[4] = fe_mac_counter(), fe_mac_counter() = 0;
  return st;
#else
  return -1;
#endif
}
void hs_fp2_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {  // 16 words each: c0 | c1 (Montgomery)
  Fp2 x, y;
  memcpy(x.c0.v, a, 32), memcpy(x.c1.v, a + 8, 32), memcpy(y.c0.v, b, 32), memcpy(y.c1.v, b + 8, 32);
  Fp2 z = mul(x, y);
  memcpy(r, z.c0.v, 32), memcpy(r + 8, z.c1.v, 32);
}
void hs_fp6_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {  // 48 words each: c0.c0 | c0.c1 | c1.c0 | ... (Montgomery)
  Fp6 x, y, z, zl;
  memcpy(&x, a, 192), memcpy(&y, b, 192);
  mul(z, x, y);
  mul_lazy(zl, x, y);  // the opt-in lazily reduced multiplier must give the same element
  memcpy(r, &z, 192);
  if (memcmp(&zl, &z, 192) != 0) memset(r, 0xff, 192);
  mul_lazy(x, x, y);  // aliased destination
  if (memcmp(&x, &z, 192) != 0) memset(r, 0xff, 192);
}
void hs_fp2_mul_xi(uint32_t* r, const uint32_t* a) {
  Fp2 x;
  memcpy(x.c0.v, a, 32), memcpy(x.c1.v, a + 8, 32);
  Fp2 z = mul_xi(x);
  memcpy(r, z.c0.v, 32), memcpy(r + 8, z.c1.v, 32);
}
void hs_fp_halve(uint32_t* r, const uint32_t* a) {
  Fp x;
  memcpy(x.v, a, 32);
  Fp z = fe_halve(x);
  memcpy(r, z.v, 32);
}
void hs_fp2_sqr(uint32_t* r, const uint32_t* a) {
  Fp2 x;
  memcpy(x.c0.v, a, 32), memcpy(x.c1.v, a + 8, 32);
  Fp2 z = sqr(x);
  memcpy(r, z.c0.v, 32), memcpy(r + 8, z.c1.v, 32);
}
unsigned long long hs_mul_count(int reset) {  // limb multiply-adds since the last reset
#ifdef BN254_COUNT_MULS
  unsigned long long c = fe_mac_counter();
  if (reset) fe_mac_counter() = 0;
  return c;
#else
  (void)reset;
  return 0;
#endif
}
// [k] P through the GLV / windowed path of the PlonK term kernels (k: 32 bytes BE, < r); 0 on identity
int hs_g1_mul_w4(uint8_t* out64, const uint8_t* pt64, const uint8_t* k_be) {
  G1Aff p, r;
  load_g1_unchecked(p, pt64);
  Fr k;
  fe_from_be_bytes(k, k_be);
  if (!to_affine(r, g1_mul_w4(p, k.v))) return 0;
  store_g1(out64, r);
  return 1;
}
void hs_sha256(const uint8_t* data, uint32_t len, uint8_t* out) {
  Sha256 s;
  sha256_init(s);
  sha256_update(s, data, len);
  sha256_final(s, out);
}
}

// ---- three lanes per proof (csrc/trio.cuh): the host build runs the three lanes in lock-step ---------------------
#include "../../snark-bn254-verifier_b200/csrc/trio.cuh"

extern "C" {
// a, b: Fq12 values as 96 Montgomery words each (c0.c0 c0.c1 c0.c2 c1.c0 c1.c1 c1.c2).  Runs every sliced Fq12
// operation next to the sequential one; returns a bit mask of the operations whose results differ (0 = all equal):
// 1 mul, 2 sqr, 4 cyclotomic_sqr(a), 8 inv, 16/32/64 frobenius 1/2/3, 128 aliased mul/sqr, 256 conj.
int hs_trio_fp12_ops(const uint32_t* aw, const uint32_t* bw) {
  Fp12 a, b, want, got;
  memcpy(&a, aw, 384), memcpy(&b, bw, 384);
  const trio::S12 sa = trio::fp12s_load(a), sb = trio::fp12s_load(b);
  trio::S12 sr;
  int bad = 0;
  mul(want, a, b), trio::fp12s_mul(sr, sa, sb), trio::fp12s_store(got, sr);
  if (memcmp(&want, &got, 384)) bad |= 1;
  sqr(want, a), trio::fp12s_sqr(sr, sa), trio::fp12s_store(got, sr);
  if (memcmp(&want, &got, 384)) bad |= 2;
  cyclotomic_sqr(want, a), trio::fp12s_cyclotomic_sqr(sr, sa), trio::fp12s_store(got, sr);
  if (memcmp(&want, &got, 384)) bad |= 4;
  inv(want, a), trio::fp12s_inv(sr, sa), trio::fp12s_store(got, sr);
  if (memcmp(&want, &got, 384)) bad |= 8;
  frobenius<1>(want, a), trio::fp12s_frobenius<1>(sr, sa), trio::fp12s_store(got, sr);
  if (memcmp(&want, &got, 384)) bad |= 16;
  frobenius<2>(want, a), trio::fp12s_frobenius<2>(sr, sa), trio::fp12s_store(got, sr);
  if (memcmp(&want, &got, 384)) bad |= 32;
  frobenius<3>(want, a), trio::fp12s_frobenius<3>(sr, sa), trio::fp12s_store(got, sr);
  if (memcmp(&want, &got, 384)) bad |= 64;
  {  // destination aliases an operand
    trio::S12 t = sa;
    trio::fp12s_mul(t, t, sb), trio::fp12s_sqr(t, t), trio::fp12s_cyclotomic_sqr(t, t), trio::fp12s_store(got, t);
    mul(want, a, b), sqr(want, want), cyclotomic_sqr(want, want);
    if (memcmp(&want, &got, 384)) bad |= 128;
  }
  conj(want, a), trio::fp12s_store(got, trio::fp12s_conj(sa));
  if (memcmp(&want, &got, 384)) bad |= 256;
  return bad;
}
// final exponentiation of a canonical Fq12 value, sliced: writes the canonical result
void hs_trio_final_exp(const uint8_t* f_be, uint8_t* out_be) {
  Fp12 f;
  Fp2* cs[6] = {&f.c0.c0, &f.c0.c1, &f.c0.c2, &f.c1.c0, &f.c1.c1, &f.c1.c2};
  for (int i = 0; i < 6; i++) {
    fp_load_be(cs[i]->c0, f_be + 64 * i);
    fp_load_be(cs[i]->c1, f_be + 64 * i + 32);
  }
  trio::S12 r;
  trio::fp12s_final_exponentiation(r, trio::fp12s_load(f));
  trio::fp12s_store(f, r);
  fp12_to_bytes(out_be, f);
}
// Miller value of the two-pair KZG check of a PlonK VK for G1 points pf (2 x 64 bytes), sliced and sequential:
// returns 0 when equal; writes the sliced canonical bytes
int hs_trio_plonk_miller(void* vkp, const uint8_t* pf_be, uint8_t* out_be) {
  const PlonkVkDev& vk = *(PlonkVkDev*)vkp;
  G1Aff pf[2];
  load_g1_unchecked(pf[0], pf_be);
  load_g1_unchecked(pf[1], pf_be + 64);
  Fp12 want, got;
  miller_loop_pairtab<0>(want, nullptr, nullptr, pf, vk.g2_pairs);
  trio::S12 f;
  trio::miller_loop_pairtab0_s(f, pf, vk.g2_pairs);
  trio::fp12s_store(got, f);
  fp12_to_bytes(out_be, got);
  return memcmp(&want, &got, 384) ? 1 : 0;
}
}

extern "C" {
static int line_differs(const trio::S12& ls, const Line3& l) {
  Fp12 got, want;
  memset(&want, 0, sizeof want);
  want.c0.c0 = l.x0, want.c0.c2 = l.x2, want.c1.c1 = l.x4;
  trio::fp12s_store(got, ls);
  return memcmp(&got, &want, 384) != 0;
}
static int point_differs(const trio::V2& r, const G2Jac& p) {
  return memcmp(&r.l[0], &p.x, 64) || memcmp(&r.l[1], &p.y, 64) || memcmp(&r.l[2], &p.z, 64);
}
// G2 steps of the Miller loop, sliced against sequential: doubling, +Q, doubling, -Q, line product, end-point test
// chain.  q: 128 bytes (any point of E'(Fq2)), p: 64 bytes.  Returns a bit mask of mismatching steps (0 = equal).
int hs_trio_g2_steps(const uint8_t* q_be, const uint8_t* p_be) {
  G2Aff q;
  G1Aff p;
  load_g2_unchecked(q, q_be);
  load_g1_unchecked(p, p_be);
  G2Jac r = to_jac(q);
  trio::V2 rs = trio::v_const(q.x, q.y, fp2_one());
  const trio::V1 px = trio::v_bcast(p.x), py = trio::v_bcast(p.y);
  const trio::V2 qx = trio::v_bcast(q.x), qy = trio::v_bcast(q.y);
  Line3 l1, l2;
  trio::S12 s1, s2, sm;
  int bad = 0;
  doubling_step_at(l1, r, &p), trio::doubling_step_s(s1, rs, px, py);
  if (line_differs(s1, l1) || point_differs(rs, r)) bad |= 1;
  addition_step_at(l2, r, q, false, &p), trio::addition_step_s(s2, rs, qx, qy, px, py);
  if (line_differs(s2, l2) || point_differs(rs, r)) bad |= 2;
  Line5 m;
  mul_lines(m, l1, l2), trio::mul_lines_s(sm, s1, s2);
  {
    Fp12 got, want;
    memset(&want, 0, sizeof want);
    want.c0.c0 = m.m0, want.c0.c1 = m.m1, want.c0.c2 = m.m2, want.c1.c0 = m.n0, want.c1.c1 = m.n1;
    trio::fp12s_store(got, sm);
    if (memcmp(&got, &want, 384)) bad |= 4;
  }
  doubling_step_at(l1, r, &p), trio::doubling_step_s(s1, rs, px, py);
  if (line_differs(s1, l1) || point_differs(rs, r)) bad |= 8;
  addition_step_at(l2, r, q, true, &p), trio::addition_step_s(s2, rs, qx, trio::v_neg(qy), px, py);
  if (line_differs(s2, l2) || point_differs(rs, r)) bad |= 16;
  return bad;
}
// Groth16 Miller loop (variable pair + pair table of the VK), sliced against sequential.  a: 64, b: 128, lc: 2 x 64 bytes
// (L, C).  Returns 0 when the Fq12 value and the G2 verdict agree; *in_g2_out receives the verdict.
int hs_trio_groth16_miller(void* vkp, const uint8_t* a_be, const uint8_t* b_be, const uint8_t* lc_be, int* in_g2_out) {
  const Groth16VkDev& vk = *(Groth16VkDev*)vkp;
  G1Aff a, pf[2];
  G2Aff b;
  load_g1_unchecked(a, a_be), load_g2_unchecked(b, b_be);
  load_g1_unchecked(pf[0], lc_be), load_g1_unchecked(pf[1], lc_be + 64);
  Fp12 want, got;
  bool g_want = false, g_got = false;
  miller_loop_pairtab<1>(want, &a, &b, pf, vk.gd_pairs, &g_want);
  trio::S12 f;
  trio::miller_loop_pairtab1_s(f, a, b, pf, vk.gd_pairs, &g_got);
  trio::fp12s_store(got, f);
  if (in_g2_out) *in_g2_out = g_got;
  return (memcmp(&want, &got, 384) ? 1 : 0) | (g_want != g_got ? 2 : 0);
}
// k-pair product Miller value, sliced against sequential (skip: identity mask)
int hs_trio_pairing_miller(int k, const uint8_t* g1, const uint8_t* g2, uint32_t skip) {
  G1Aff p[4];
  G2Aff q[4];
  for (int j = 0; j < k; j++) load_g1_unchecked(p[j], g1 + 64 * j), load_g2_unchecked(q[j], g2 + 128 * j);
  Fp12 want, got;
  trio::S12 f;
  switch (k) {
    case 1: miller_loop<1, 0>(want, p, q, nullptr, nullptr, skip), trio::miller_loop_var_s<1>(f, p, q, skip); break;
    case 2: miller_loop<2, 0>(want, p, q, nullptr, nullptr, skip), trio::miller_loop_var_s<2>(f, p, q, skip); break;
    case 3: miller_loop<3, 0>(want, p, q, nullptr, nullptr, skip), trio::miller_loop_var_s<3>(f, p, q, skip); break;
    default: miller_loop<4, 0>(want, p, q, nullptr, nullptr, skip), trio::miller_loop_var_s<4>(f, p, q, skip); break;
  }
  trio::fp12s_store(got, f);
  return memcmp(&want, &got, 384) ? 1 : 0;
}
}

// ---- opt-in aggregate Groth16 check (csrc/groth16_agg.cuh) --------------------------------------------------------
#include "../../snark-bn254-verifier_b200/csrc/groth16_agg.cuh"

extern "C" {
// the IC_0 / alpha window tables the CUDA vk_load builds for the aggregate check (vk from hs_groth16_vk_new*)
void hs_groth16_vk_add_agg_tables(void* vkp) {
  Groth16VkDev* vk = (Groth16VkDev*)vkp;
  G1Aff* tab = new G1Aff[(size_t)2 * BN_IC_WINDOWS * BN_IC_ENTRIES];
  for (int w = 0; w < BN_IC_WINDOWS; w++) {
    groth16_ic_table_slice(tab + (size_t)w * BN_IC_ENTRIES, vk->ic[0], w);
    groth16_ic_table_slice(tab + ((size_t)BN_IC_WINDOWS + w) * BN_IC_ENTRIES, vk->alpha, w);
  }
  vk->agg_table = tab;  // (leaked with the VK: test process)
}
// [a + b lambda] P0 and P1 through the aggregate check's double-scalar routine (a, b: 8 bytes LE each); 0 on identity
int hs_g1_mul_glv64_2(uint8_t* out128, const uint8_t* pts128, const uint8_t* a8, const uint8_t* b8) {
  G1Aff p[2], r;
  load_g1_unchecked(p[0], pts128);
  load_g1_unchecked(p[1], pts128 + 64);
  uint32_t a[2], b[2];
  for (int k = 0; k < 2; k++) {
    a[k] = (uint32_t)a8[4 * k] | ((uint32_t)a8[4 * k + 1] << 8) | ((uint32_t)a8[4 * k + 2] << 16) | ((uint32_t)a8[4 * k + 3] << 24);
    b[k] = (uint32_t)b8[4 * k] | ((uint32_t)b8[4 * k + 1] << 8) | ((uint32_t)b8[4 * k + 2] << 16) | ((uint32_t)b8[4 * k + 3] << 24);
  }
  G1Jac res[2];
  g1_mul_glv64_2(res, p, a, b);
  for (int v = 0; v < 2; v++) {
    if (!to_affine(r, res[v])) return 0;
    store_g1(out128 + 64 * v, r);
  }
  return 1;
}
// multiply-adds of one proof's share of the aggregate check: [r] C, then validation + [r] A + the single-pair Miller loop
unsigned long long hs_groth16_agg_proof_macs(void* vkp, const uint8_t* proof, uint32_t len, const uint8_t* inputs, int n_inputs,
                                             const uint8_t* rnd16, unsigned long long* prepare_part) {
#ifdef BN254_COUNT_MULS
  const unsigned long long before = fe_mac_counter();
  G1Aff rA;
  G2Aff B;
  G1Jac g;
  const int st = groth16_agg_prepare_one(rA, B, g, *(Groth16VkDev*)vkp, proof, len, inputs, n_inputs, rnd16);
  const unsigned long long mid = fe_mac_counter();
  Fp12 f;
  groth16_agg_miller_one(f, *(Groth16VkDev*)vkp, rA, B, st == BN254V_OK_TRUE);
  if (prepare_part) *prepare_part = mid - before;
  return fe_mac_counter() - before;
#else
  return 0;
#endif
}
// n records of `stride` bytes, n x n_inputs x 32 input bytes, n x 16 scalar bytes, (1 + n_inputs) x 32 batch scalars.
// Folds in groups of `per` like the CUDA reduction kernel.  status_out[n]; f_out: n x 384 (per-proof Miller values) or
// null.  Returns the batch verdict (1 = all valid).
int hs_groth16_agg(void* vkp, const uint8_t* proofs, size_t stride, const uint32_t* lens, int n, const uint8_t* inputs,
                   int n_inputs, const uint8_t* rnd16, const uint8_t* scal_be, int per, uint8_t* status_out, uint8_t* f_out) {
  const Groth16VkDev& vk = *(Groth16VkDev*)vkp;
  std::vector<Fp12> f(n);
  std::vector<G1Jac> g(n);
  bool all_ok = true;
  for (int i = 0; i < n; i++) {
    const uint32_t len = lens ? lens[i] : (uint32_t)stride;
    G1Aff rA;
    G2Aff B;
    int st = groth16_agg_prepare_one(rA, B, g[i], vk, proofs + stride * i, len, inputs + (size_t)32 * n_inputs * i, n_inputs,
                                     rnd16 + 16 * i);
    if (!groth16_agg_miller_one(f[i], vk, rA, B, st == BN254V_OK_TRUE) && st == BN254V_OK_TRUE) st = BN254V_PANIC_NOT_IN_SUBGROUP;
    status_out[i] = (uint8_t)st;
    if (st != BN254V_OK_TRUE) all_ok = false;
    if (f_out) fp12_to_bytes(f_out + 384 * (size_t)i, f[i]);
  }
  size_t cur = n;
  while (cur > 1) {  // the shape of the CUDA trees: `per` to one per pass
    size_t nxt = (cur + per - 1) / per;
    std::vector<Fp12> f2(nxt);
    std::vector<G1Jac> g2(nxt);
    for (size_t t = 0; t < nxt; t++) {
      f2[t] = f[t * per], g2[t] = g[t * per];
      for (size_t k = t * per + 1; k < cur && k < (t + 1) * per; k++) mul(f2[t], f2[t], f[k]), g2[t] = jac_add(g2[t], g[k]);
    }
    f.swap(f2), g.swap(g2);
    cur = nxt;
  }
  Fp12 fp;
  const bool verdict = groth16_agg_fprime(fp, g[0], vk, scal_be) && groth16_agg_final(f[0], fp);
  return all_ok && verdict ? 1 : 0;
}
}
