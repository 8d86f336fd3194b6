// Lane-pair execution of the G2 / pairing arithmetic: TWO adjacent lanes per proof.
//
// One proof per thread leaves a 2^16-proof batch at 3.5 warps per SMSP, and a pairing's working set (Fq12 accumulator,
// G2 point, temporaries) at ~6 KB of stack per thread.  Here every Fq2 value (a0 + a1 u) is split across the two lanes
// of a pair -- the even lane holds a0, the odd lane a1 -- so the same batch runs on twice as many warps, each thread
// carries half the state, and additions/subtractions cost one Fq operation per lane.  Cross terms travel by
// __shfl_xor(.., 1):
//   mul   lane0: a0 b0 + a1 (p - b1)     lane1: a1 b0 + a0 b1      two wide products + one Montgomery reduction each
//   sqr   lane0: (a0 + a1)(a0 - a1)      lane1: (a0 + a0) a1       one Montgomery multiplication each
// The Fq6 / Fq12 tower, the line steps, the Miller loop and the final exponentiation are the SAME source as the scalar
// build (tower_body.inc, pairing_body.inc), included here against this Fp2 type, so both builds compute the same field
// elements; G1 arithmetic (Fq only) is done redundantly by both lanes with the scalar code.
//
// Rules: the two lanes of a pair always take the same branches (all predicates are exchanged before use) and shuffles
// name only the pair, so pairs may diverge from each other; the block-wide phase barriers sit in code every thread
// executes, so these kernels never return early -- a malformed proof keeps running on substitute data and only its
// status differs.  Device code only.
#pragma once
#if defined(__CUDACC__)
#include "groth16.cuh"
#include "plonk.cuh"

#pragma push_macro("HD")
#pragma push_macro("HDN")
#undef HD
#undef HDN
#define HD __device__ __forceinline__
#define HDN static __device__ __noinline__

namespace bn254 {
namespace lp {

typedef bn254::Fp2 Fp2F;    // stored (full) forms, as in the VK structs and line tables
typedef bn254::Fp12 Fp12F;
typedef bn254::Line LineF;
typedef bn254::G1Aff G1Aff;

HD int lane_h() { return threadIdx.x & 1; }  // 0: holds c0, 1: holds c1
// Shuffles name only the two lanes of the pair: every branch in this code is uniform within a pair, but pairs of one
// warp may diverge from each other (e.g. the special cases of a point addition), which a full-warp mask would not allow.
HD unsigned pair_mask() { return 3u << (threadIdx.x & 30); }
HD Fp xchg(const Fp& a) {
  const unsigned m = pair_mask();
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(m, a.v[i], 1);
  return r;
}
HD Fp sel(bool odd, const Fp& if_odd, const Fp& if_even) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = odd ? if_odd.v[i] : if_even.v[i];
  return r;
}

struct Fp2 {
  Fp c;  // this lane's component
};
HD Fp2 ld2(const Fp2F& s) { return Fp2{((const Fp*)&s)[lane_h()]}; }
HD Fp2 fp2_zero() { return Fp2{fe_zero<FpCfg>()}; }
HD Fp2 fp2_one() { return Fp2{sel(lane_h(), fe_zero<FpCfg>(), fe_one<FpCfg>())}; }
HD Fp2 add(const Fp2& a, const Fp2& b) { return Fp2{fe_add(a.c, b.c)}; }
HD Fp2 sub(const Fp2& a, const Fp2& b) { return Fp2{fe_sub(a.c, b.c)}; }
HD Fp2 neg(const Fp2& a) { return Fp2{fe_neg(a.c)}; }
HD Fp2 dbl(const Fp2& a) { return Fp2{fe_dbl(a.c)}; }
HD Fp2 conj(const Fp2& a) { return Fp2{sel(lane_h(), fe_neg(a.c), a.c)}; }
HD bool is_zero(const Fp2& a) {
  int z = fe_is_zero(a.c);
  return z & __shfl_xor_sync(pair_mask(), z, 1);
}
HD bool eq(const Fp2& a, const Fp2& b) {
  int z = fe_eq(a.c, b.c);
  return z & __shfl_xor_sync(pair_mask(), z, 1);
}
// lane0: a0 b0 - a1 b1 = a0 b0 + a1 (p - b1);  lane1: a1 b0 + a0 b1.  Both are own*U + other*V with
// (U, V) = (b_own, p - b_other) on lane0 and (b_other, b_own) on lane1; the sum of the two wide products is < 2 p^2.
HDN Fp2 mul(Fp2 a, Fp2 b) {
  const bool h = lane_h();
  Fp oa = xchg(a.c), ob = xchg(b.c);
  Fp U = sel(h, ob, b.c), V = sel(h, b.c, fe_mod_minus(ob));
  uint32_t T0[16], T1[16];
  fe_mul_wide(T0, a.c, U);
  fe_mul_wide(T1, oa, V);
  T0[0] = cc::add_cc(T0[0], T1[0]);
#pragma unroll
  for (int k = 1; k < 15; k++) T0[k] = cc::addc_cc(T0[k], T1[k]);
  T0[15] = cc::addc(T0[15], T1[15]);
  return Fp2{fe_redc_wide<FpCfg>(T0)};
}
HDN Fp2 sqr(Fp2 a) {
  const bool h = lane_h();
  Fp oa = xchg(a.c);
  Fp X = fe_add_nr(sel(h, oa, a.c), oa);       // lane0: a0 + a1, lane1: 2 a0
  Fp Y = sel(h, a.c, fe_sub(a.c, oa));         // lane0: a0 - a1, lane1: a1
  return Fp2{fe_mul(X, Y)};
}
HD Fp2 scale(const Fp2& a, const Fp& k) { return Fp2{fe_mul(a.c, k)}; }
HD Fp2 scale2_add(const Fp2& a, const Fp& j, const Fp2& b, const Fp& k) { return Fp2{fe_mul2_add(a.c, j, b.c, k)}; }
// (9 + u)(a0 + a1 u) = (9 a0 - a1) + (9 a1 + a0) u
HDN Fp2 mul_xi(Fp2 a) {
  Fp oa = xchg(a.c);
  return Fp2{fe_mul9_add(a.c, sel(lane_h(), oa, fe_mod_minus(oa)))};
}
HDN Fp2 inv(Fp2 a) {
  Fp s = fe_sqr(a.c);
  Fp n = fe_inv(fe_add(s, xchg(s)));  // 1 / (a0^2 + a1^2), computed by both lanes
  Fp r = fe_mul(a.c, n);
  return Fp2{sel(lane_h(), fe_neg(r), r)};
}
HD Fp2 fp2_halve(const Fp2& a) { return Fp2{fe_halve(a.c)}; }
HD Fp2 lane_const(const Fp& c0, const Fp& c1) { return Fp2{sel(lane_h(), c1, c0)}; }
HD Fp2 fp2_b2() {
  Fp c0, c1;
  BN_LOAD_FP(c0, K::b2, 0);
  BN_LOAD_FP(c1, K::b2, 1);
  return lane_const(c0, c1);
}
template <int KK>
HD Fp2 frob_coeff(int i) {
  bn254::Fp2 f = bn254::frob_coeff<KK>(i);
  return lane_const(f.c0, f.c1);
}
template <int KK>
HD Fp frob_coeff_fp(int i) { return bn254::frob_coeff<KK>(i).c0; }

#undef BN_HAVE_FP6_MUL_LAZY  // the lazily reduced Fq6 multiplication is written for the scalar Fq2 layout
#include "tower_body.inc"

HD Fp12 ld12(const Fp12F& s) {
  return Fp12{Fp6{ld2(s.c0.c0), ld2(s.c0.c1), ld2(s.c0.c2)}, Fp6{ld2(s.c1.c0), ld2(s.c1.c1), ld2(s.c1.c2)}};
}
// canonical serialisation: each lane writes its own component of the six Fq2 coefficients
HDN void fp12_to_bytes(uint8_t* out, const Fp12& a) {
  const Fp2* cs[6] = {&a.c0.c0, &a.c0.c1, &a.c0.c2, &a.c1.c0, &a.c1.c1, &a.c1.c2};
  for (int i = 0; i < 6; i++) fe_to_be_bytes(out + 64 * i + 32 * lane_h(), fe_from_mont(cs[i]->c));
}

}  // namespace lp

// curve.cuh's generic templates (Aff / Jac / jac_double / jac_add / scalar_mul ...) work on lp::Fp2 through
// argument-dependent lookup; they only need these three constants.
// (the primary templates are __host__ __device__, hence the guards: the lane-pair type only exists in device code)
#if defined(__CUDA_ARCH__)
#define BN_LP_DEV(expr) return expr
#else
#define BN_LP_DEV(expr) return lp::Fp2()
#endif
template <> __host__ __device__ __forceinline__ lp::Fp2 f_one<lp::Fp2>() { BN_LP_DEV(lp::fp2_one()); }
template <> __host__ __device__ __forceinline__ lp::Fp2 f_zero<lp::Fp2>() { BN_LP_DEV(lp::fp2_zero()); }
template <> __host__ __device__ __forceinline__ lp::Fp2 curve_b<lp::Fp2>() { BN_LP_DEV(lp::fp2_b2()); }
#undef BN_LP_DEV

namespace lp {

typedef bn254::Aff<Fp2> G2Aff;
typedef bn254::Jac<Fp2> G2Jac;
typedef bn254::G2Aff G2AffF;

HD G2Aff ld_g2(const G2AffF& q) { return G2Aff{ld2(q.x), ld2(q.y)}; }
HD G2Aff g2_psi(const G2Aff& q) {
  return G2Aff{mul(conj(q.x), frob_coeff<1>(2)), mul(conj(q.y), frob_coeff<1>(3))};
}
HD G2Jac g2_psi_jac(const G2Jac& p) {
  return G2Jac{mul(conj(p.x), frob_coeff<1>(2)), mul(conj(p.y), frob_coeff<1>(3)), conj(p.z)};
}
HD bool jac_eq(const G2Jac& a, const G2Jac& b) {
  bool ia = is_identity(a), ib = is_identity(b);
  if (ia || ib) return ia && ib;
  Fp2 za2 = sqr(a.z), zb2 = sqr(b.z);
  bool e1 = eq(mul(a.x, zb2), mul(b.x, za2));
  bool e2 = eq(mul(a.y, mul(zb2, b.z)), mul(b.y, mul(za2, a.z)));
  return e1 && e2;
}
// [x+1]P + psi([x]P) + psi^2([x]P) == psi^3([2x]P)   (see curve_body.inc)
HDN bool g2_in_subgroup(const G2Aff& q) {
  const uint32_t k[8] = {0x4a6909f1u, 0x44e992b4u, 0, 0, 0, 0, 0, 0};
  G2Jac a = scalar_mul<Fp2, true>(q, k);
  G2Jac b = g2_psi_jac(a);
  G2Jac c = g2_psi_jac(b);
  G2Jac d = g2_psi_jac(c);
  G2Jac lhs = jac_add(jac_add(jac_add_mixed(a, q), b), c);
  return jac_eq(lhs, jac_double(d));
}

struct Line {
  Fp2 ell_0, ell_vw, ell_vv;
};
HD Line lane_line(const LineF& l) { return Line{ld2(l.ell_0), ld2(l.ell_vw), ld2(l.ell_vv)}; }

typedef bn254::LinePairKF LinePairKF;
#define BN_LD_LINE(l) lane_line(l)
#define BN_LD_FP2(x) ld2(x)
#include "pairing_body.inc"
#undef BN_LD_LINE
#undef BN_LD_FP2

// ---- wire format: x1 | x0 | y1 | y0, each lane reads its own component
HD bool load_g2_lane(G2Aff& q, const uint8_t* b) {
  const int o = lane_h() ? 0 : 32;
  bool ok = fp_load_be(q.x.c, b + o);
  ok = fp_load_be(q.y.c, b + 64 + o) && ok;
  int z = ok;
  return z & __shfl_xor_sync(pair_mask(), z, 1);
}
HD G2Aff g2_generator_lane() {
  Fp c0, c1, d0, d1;
  BN_LOAD_FP(c0, K::g2_gen, 0);
  BN_LOAD_FP(c1, K::g2_gen, 1);
  BN_LOAD_FP(d0, K::g2_gen, 2);
  BN_LOAD_FP(d1, K::g2_gen, 3);
  return G2Aff{lane_const(c0, c1), lane_const(d0, d1)};
}
HD G1Aff g1_generator_lane() {
  G1Aff g;
  BN_LOAD_FP(g.x, K::g1_gen, 0);
  BN_LOAD_FP(g.y, K::g1_gen, 1);
  return g;
}

// ---- Groth16, one proof per lane pair.  Same checks, same order and same values as bn254::groth16_verify_one; a
// failed check records the status and continues on substitute data (the generators) instead of returning.
HD int groth16_verify_pair(const Groth16VkDev& vk, const uint8_t* proof, uint32_t proof_len, const uint8_t* inputs_be,
                           int n_inputs, const Groth16Debug& dbg) {
  int st = BN254V_OK_TRUE;
  const bool h = lane_h();
  const bool short_buf = proof_len < 256;
  if (short_buf) st = BN254V_PANIC_SHORT_BUFFER;
  G1Aff A = g1_generator_lane(), C = A;
  G2Aff B = g2_generator_lane();
  if (!short_buf) {  // uniform within the pair; reads stay inside the record
    int s = load_g1_checked(A, proof);
    if (s != BN254V_OK_TRUE) {
      if (st == BN254V_OK_TRUE) st = s;
      A = g1_generator_lane();
    }
    G2Aff Bp;
    bool ok = load_g2_lane(Bp, proof + 64);
    s = !ok ? BN254V_PANIC_FIELD_NOT_MEMBER : (!on_curve(Bp) ? BN254V_PANIC_NOT_ON_CURVE : BN254V_OK_TRUE);
    if (s == BN254V_OK_TRUE) B = Bp;
    else if (st == BN254V_OK_TRUE) st = s;
  }
  // the subgroup test contains block barriers: every thread runs it, on the generator when B was rejected above
  bool in_g2 = g2_in_subgroup(B);
  if (!in_g2) {
    if (st == BN254V_OK_TRUE) st = BN254V_PANIC_NOT_IN_SUBGROUP;
    B = g2_generator_lane();
  }
  if (!short_buf) {
    int s = load_g1_checked(C, proof + 192);
    if (s != BN254V_OK_TRUE) {
      if (st == BN254V_OK_TRUE) st = s;
      C = g1_generator_lane();
    }
  }
  G1Aff L;
  {
    int s = bn254::groth16_prepare_inputs(L, vk, inputs_be, n_inputs);
    if (s != BN254V_OK_TRUE) {
      if (st == BN254V_OK_TRUE) st = s;
      L = g1_generator_lane();
    }
  }
  if (dbg.L && !h) store_g1(dbg.L, L);

  G1Aff pf[2] = {L, C};
  const LineF* tabs[2] = {vk.gamma_lines, vk.delta_lines};
  Fp12 f;
  miller_loop<1, 2>(f, &A, &B, pf, tabs);
  if (dbg.miller) fp12_to_bytes(dbg.miller, f);
  final_exponentiation(f, f);
  if (dbg.gt) fp12_to_bytes(dbg.gt, f);
  bool same = eq(f, ld12(vk.target));
  if (st != BN254V_OK_TRUE) return st;
  return same ? BN254V_OK_TRUE : BN254V_OK_FALSE;
}

// ---- raw k-pair product (bn::pairing_batch), one set per lane pair
template <int KP>
HD bool pairing_product_pair(const uint8_t* g1, const uint8_t* g2, uint8_t* miller_out, uint8_t* gt_out) {
  G1Aff p[KP];
  G2Aff q[KP];
  for (int j = 0; j < KP; j++) {
    load_g1_unchecked(p[j], g1 + 64 * j);
    load_g2_lane(q[j], g2 + 128 * j);
  }
  Fp12 f;
  miller_loop<KP, 0>(f, p, q, nullptr, nullptr);
  if (miller_out) fp12_to_bytes(miller_out, f);
  final_exponentiation(f, f);
  if (gt_out) fp12_to_bytes(gt_out, f);
  return eq(f, fp12_one());
}

// ---- PlonK stage E (plonk.cuh): the G1 sums are Fq-only and done by both lanes, the 2-pair pairing is lane-split
HD int plonk_stage_e_pair(PlonkWork& w, const PlonkVkDev& vk, const uint8_t* pr, const PlonkDebug& dbg) {
  const int nq = vk.n_qcp, b = 5 + nq;
  const bool h = lane_h();
  int st = BN254V_OK_TRUE;
  G1Jac fd = to_jac(w.lin);
  for (int t = 0; t < b; t++) fd = jac_add(fd, w.part[t]);
  if (is_identity(fd)) st = BN254V_PANIC_IDENTITY;
  if (dbg.g1 && !h) {
    G1Aff t;
    if (to_affine(t, fd)) store_g1(dbg.g1 + 64, t);
  }
  G1Aff bh;
  load_g1_unchecked(bh, pr + 448);
  G1Jac fq = jac_add_mixed(w.part[b + 0], bh);
  if (is_identity(fq) && st == BN254V_OK_TRUE) st = BN254V_PANIC_IDENTITY;
  G1Jac fdg = jac_add(fd, w.part[b + 1]);
  if (is_identity(fdg) && st == BN254V_OK_TRUE) st = BN254V_PANIC_IDENTITY;
  G1Jac fec = w.part[b + 2];
  if (is_identity(fec) && st == BN254V_OK_TRUE) st = BN254V_PANIC_IDENTITY;
  fec.y = bn254::neg(fec.y);
  fdg = jac_add(fdg, fec);
  if (is_identity(fdg) && st == BN254V_OK_TRUE) st = BN254V_PANIC_IDENTITY;
  G1Jac fpq = jac_add(w.part[b + 3], w.part[b + 4]);
  if (is_identity(fpq) && st == BN254V_OK_TRUE) st = BN254V_PANIC_IDENTITY;
  fdg = jac_add(fdg, fpq);
  if (is_identity(fdg) && st == BN254V_OK_TRUE) st = BN254V_PANIC_IDENTITY;
  fq.y = bn254::neg(fq.y);
  G1Aff pf[2];
  pf[0] = pf[1] = g1_generator_lane();  // substitutes when a sum was the identity (status already recorded)
  if (st == BN254V_OK_TRUE) {
    to_affine(pf[0], fdg);
    to_affine(pf[1], fq);
    if (dbg.g1 && !h) {
      store_g1(dbg.g1 + 128, pf[0]);
      store_g1(dbg.g1 + 192, pf[1]);
    }
  }
  const LineF* tabs[2] = {vk.g2_lines[0], vk.g2_lines[1]};
  Fp12 f;
  miller_loop<0, 2>(f, nullptr, nullptr, pf, tabs);
  if (dbg.miller && st == BN254V_OK_TRUE) fp12_to_bytes(dbg.miller, f);
  final_exponentiation(f, f);
  if (dbg.gt && st == BN254V_OK_TRUE) fp12_to_bytes(dbg.gt, f);
  bool one = eq(f, fp12_one());
  if (st != BN254V_OK_TRUE) return st;
  return one ? BN254V_OK_TRUE : BN254V_ERR_PAIRING_CHECK_FAILED;
}

}  // namespace lp
}  // namespace bn254

#pragma pop_macro("HDN")
#pragma pop_macro("HD")
#endif  // __CUDACC__
