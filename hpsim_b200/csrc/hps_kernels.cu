// hps_kernels.cu -- __global__ kernels + their launch wrappers (see hps_launch.h).
#include <algorithm>
#include <cstdlib>
#include "hps_kernels.cuh"
#include "hps_launch.h"

namespace hps {

enum { SEG_NS = 16, MOM_NS = 12, LL_NS = 4, STAT_NS = 8 };

// ================================================================================================
// K1: fused segment sweep
// ================================================================================================
// slots: 0 n_good, 1 sum phi (good), 2 MAX phi over lloss==0, 3 n_loss, 4 sum x (loss==0),
//        5 sum y (loss==0), 6 n_lloss, 7 sum phi (loss==0) right after OP_FREQ_CHANGE
__device__ void finalize_segment(int fin, const double* t, DevStats* st, u32 n_global) {
  if (fin & FIN_FREQ) {
    // Beam::ChangeFrequency -> UpdateAvgPhi() (beam.cpp:1666; divides by N - num_loss, stale count)
    st->v[HPS_STAT_PHI_AVG] = t[7] / (double)(u32)(n_global - st->num_loss);
  }
  switch (fin & 3) {
    case FIN_KICK: {
      // Beam::UpdateLongitudinalLoss steps 1-2 (beam.cpp:776-779); finalisers beam_kernel.cu:1420-1427
      const u32 ng = (u32)t[0];
      st->num_good = ng;
      const double avg = ng > 0 ? t[1] / ng : -99999;
      st->v[HPS_STAT_PHI_AVG_GOOD] = avg;
      st->num_lloss = (u32)t[6];
      // beam_kernel.cu:1556: a particle is marked when phi > <phi>_good + 2pi + pi; if nobody with
      // lloss==0 crosses, steps 3-6 of the chain change nothing and the lloss pass is skipped.
      st->need_lloss_pass = (t[2] > avg + K_TWOPI + K_PI) ? 1u : 0u;
      break;
    }
    case FIN_APERTURE: {
      // Beam::UpdateLoss + UpdateAvgXY (simulation_engine.cu:311-312; beam_kernel.cu:1385-1394)
      const u32 nl = (u32)t[3];
      st->num_loss = nl;
      const double surv = (double)(u32)(n_global - nl);
      st->v[HPS_STAT_X_AVG] = t[4] / surv;
      st->v[HPS_STAT_Y_AVG] = t[5] / surv;
      break;
    }
    default: break;
  }
}

// one op applied to the N register-resident particles of a thread.  The type dispatch and the shared-memory
// reads of the op's parameters happen once per op, not once per particle, and the per-particle bodies sit
// back to back so that their independent fp64 chains can overlap.
// MATH: 0 = the reference's expression order with nvcc's `/` and sqrt, 1 = the fast maps, 2 = the reference's
// expression order with the branch-free exact arithmetic of hps_math.cuh (same bits as 0)
// TYPE / CCL: the op type and the RF-gap flavour when the launch's op sequence is known at compile time (the pattern
// sweeps below: the type switch and the DTL / CCL branches inside the gap maps fold away); -1 = read them from the record.
template <bool FULL, int N, int MATH, int TYPE = -1, int CCL = -1>
__device__ __forceinline__ void apply_op(Pt (&pt)[N], const u32 (&ll)[N], const u32 (&gidx)[N], const bool (&valid)[N],
                                         const Op& op, const OpC& opc, const KickCtx& kc, const double* cxy,
                                         const double2* __restrict__ field, const BeamView& b, int compat,
                                         double (&phi_fc)[N], bool (&live_fc)[N], double inv_mass) {
  const double* q = op.p;
  const double* c = opc.c;
  const double mass = b.mass, charge = b.charge;
  const bool ccl = CCL >= 0 ? (CCL != 0) : (op.flag != 0);
#define EACH for (int j = 0; j < N; ++j) if (valid[j])
#define EACH_ALIVE for (int j = 0; j < N; ++j) if (valid[j] && pt[j].loss == 0)
  switch (TYPE >= 0 ? TYPE : op.type) {
    case OP_DRIFT:
#pragma unroll
      EACH_ALIVE { if (MATH == 1) op_drift_f(pt[j], q, c, op.elem, inv_mass); else if (MATH == 2) op_drift<MathNR>(pt[j], q, c, op.elem, mass, inv_mass); else op_drift<MathLib>(pt[j], q, c, op.elem, mass, inv_mass); }
      break;
    case OP_QUAD_HALF:
#pragma unroll
      EACH_ALIVE { if (MATH == 1) op_quad_half_f(pt[j], q, c, op.elem, inv_mass); else if (MATH == 2) op_quad_half<MathNR>(pt[j], q, c, op.elem, mass, inv_mass); else op_quad_half<MathLib>(pt[j], q, c, op.elem, mass, inv_mass); }
      break;
    case OP_GAP_FIRST:
#pragma unroll
      EACH_ALIVE { if (MATH == 1) op_gap_first_f(pt[j], q, c, ccl, inv_mass); else if (MATH == 2) op_gap_first<MathNR>(pt[j], q, c, ccl, mass, inv_mass); else op_gap_first<MathLib>(pt[j], q, c, ccl, mass, inv_mass); }
      break;
    case OP_GAP_SECOND:
#pragma unroll
      EACH_ALIVE { if (MATH == 1) op_gap_second_f(pt[j], q, c, ccl, inv_mass); else if (MATH == 2) op_gap_second<MathNR>(pt[j], q, c, ccl, mass, inv_mass); else op_gap_second<MathLib>(pt[j], q, c, ccl, mass, inv_mass); }
      break;
    // the rare, code-heavy maps (7 dipoles and 2 bunchers in the whole linac) live in the FULL variant only,
    // which keeps the hot variant's SASS inside the instruction cache
    case OP_DIPOLE_FIRST:
      if (FULL) {
#pragma unroll
        EACH_ALIVE op_dipole(pt[j], q, false, mass);
      }
      break;
    case OP_DIPOLE_SECOND:
      if (FULL) {
#pragma unroll
        EACH_ALIVE op_dipole(pt[j], q, true, mass);
      }
      break;
    case OP_BUNCHER:
      if (FULL) {
#pragma unroll
        EACH_ALIVE op_buncher(pt[j], q, op.elem, mass, charge);
      }
      break;
    case OP_ROTATION: {  // simulation_kernel.cu:511-529
      const double cosa = c[R_COS], sina = c[R_SIN];
#pragma unroll
      EACH_ALIVE {
        const double x = pt[j].x, y = pt[j].y, xp = pt[j].xp, yp = pt[j].yp;
        pt[j].x = x * cosa + y * sina; pt[j].y = y * cosa - x * sina;
        pt[j].xp = xp * cosa + yp * sina; pt[j].yp = yp * cosa - xp * sina;
      }
      break;
    }
    case OP_STEERER:  // simulation_kernel.cu:120-140
#pragma unroll
      EACH_ALIVE {
        const double gm1 = pt[j].w / mass;
        const double pm = sqrt(gm1 * (gm1 + 2.0)) * mass / K_CLIGHT;
        if (q[0] > 1e-10) pt[j].xp += q[0] * charge / pm;
        if (q[1] > 1e-10) pt[j].yp += q[1] * charge / pm;
      }
      break;
    case OP_CAPERTURE:  // simulation_kernel.cu:543-556 (reference particle exempt, H1)
#pragma unroll
      EACH_ALIVE if (gidx[j] > 0) {
        const double x = pt[j].x - cxy[0], y = pt[j].y - cxy[1];
        if (x * x + y * y > q[0] * q[0]) pt[j].loss = op.elem;
      }
      break;
    case OP_RAPERTURE: {  // simulation_kernel.cu:558-575
      const double xmin = cxy[0] - q[0], xmax = cxy[0] + q[1];
      const double ymin = cxy[1] - q[3], ymax = cxy[1] + q[2];
#pragma unroll
      EACH_ALIVE if (gidx[j] > 0) {
        if (pt[j].x < xmin || pt[j].x > xmax || pt[j].y < ymin || pt[j].y > ymax) pt[j].loss = op.elem;
      }
      break;
    }
    case OP_SC_KICK:
      // H2: the reference launches the kick <<<grid, blck/4>>> with `while(indx > 0 && ...)`, so
      // global thread 0 exits at once and particles k*NextPow2(N)/4 never see a kick.
#pragma unroll
      // (np2_quarter is a power of two)
      EACH if (!(gidx[j] == 0 || (compat && (gidx[j] & (b.np2_quarter - 1u)) == 0u))) {
        if (MATH == 1) op_sc_kick_f(pt[j], ll[j], kc, field, inv_mass); else if (MATH == 2) op_sc_kick<MathNR>(pt[j], ll[j], kc, field, mass, inv_mass); else op_sc_kick<MathLib>(pt[j], ll[j], kc, field, mass, inv_mass);
      }
      break;
    case OP_FREQ_CHANGE:  // beam_kernel.cu:1531-1538: every particle, lost or not
#pragma unroll
      EACH {
        pt[j].phi *= q[0];
        phi_fc[j] = pt[j].phi;
        live_fc[j] = pt[j].loss == 0;
      }
      break;
    default: break;
  }
#undef EACH
#undef EACH_ALIVE
}

// The epilogue's partial sums live in SHARED memory, one column per thread ([slot][thread], conflict-free): the sweep's
// body holds two particles and the temporaries of the RF-gap map in registers right at the 128-register budget of two
// resident blocks per SM; 16 accumulators carried across that body would cost 32 more.
// The sums of one loop iteration (up to two particles) are formed in registers AFTER the maps -- the maps' temporaries are
// dead by then -- and folded into the thread's column with one batched read-modify-write: all loads, then all stores
// (a chain of dependent shared-memory round trips per slot cost 80 us per 16 Mi-particle sweep).
__device__ __forceinline__ double lds_f64(const double* p) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((u32)__cvta_generic_to_shared(p)));
  return v;
}
__device__ __forceinline__ void sts_f64(double* p, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(p)), "d"(v));
}
template <int NS>
__device__ __forceinline__ void acc_flush(double* col, const double (&c)[NS], unsigned max_mask) {
  double t[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) t[s] = lds_f64(col + s * HPS_BLOCK);
#pragma unroll
  for (int s = 0; s < NS; ++s) t[s] = ((max_mask >> s) & 1u) ? fmax(t[s], c[s]) : t[s] + c[s];
#pragma unroll
  for (int s = 0; s < NS; ++s) sts_f64(col + s * HPS_BLOCK, t[s]);
}

enum { SEGL_NS = 8 };  // slots of the legacy epilogues (FIN_KICK / FIN_APERTURE), see finalize_segment
__device__ __forceinline__ void seg_accumulate(double (&c)[SEGL_NS], const Pt& p, u32 ll, double phi_fc, bool live_fc) {
  const bool alive = p.loss == 0, good = alive && ll == 0;
  if (good) { c[0] += 1.0; c[1] += p.phi; }
  if (ll == 0) c[2] = fmax(c[2], p.phi);
  else c[6] += 1.0;
  if (alive) { c[4] += p.x; c[5] += p.y; }
  else c[3] += 1.0;
  if (live_fc) c[7] += phi_fc;
}

// ---- fused moments (FIN_KICK_FUSED) ---------------------------------------------------------------------------------
// Beam::UpdateStatForSpaceCharge (beam.cpp:1003-1016) is a chain of dependent reductions: <phi>_good -> longitudinal-loss
// marking against <phi>_good + 3 pi -> <phi>_good again -> relative phases about it -> their moments.  The reference (and
// round 1 of this code) re-reads the particles for every link.  Here the sweep that moves the particles also sums
//   u = rel_phi(phi, P)      about a PREDICTED pivot P (a pilot particle at (<phi>_good, <W>_good) of the previous kick,
//                            carried through the segment's maps by thread 0 of every block),
// and the last block re-centres exactly once the true R = <phi>_good is known: with delta = R - P,
//   sum phi_r = sum u - n delta,   sum phi_r^2 = sum u^2 - 2 delta sum u + n delta^2
// for every particle whose classification cannot change within |delta| <= m.  The others -- within m of the 2 pi wrap of
// the relative phase, or within m of the marking threshold P + 3 pi -- are appended to a side list (DeferList) and
// resolved one by one with the true R, their contributions summed in two-limb 64-bit FIXED POINT (2^-80 resolution) so
// that the (arbitrary) list order cannot change a bit of the result.  A prediction that misses by more than m, or a list
// overflow, hands the kick to the stand-alone moment kernel (k_moments, DevStats::moments_fallback).
// slots: 0 packed counts A: n_good before marking (all particles) | n_lloss != 0 (all) << 21 | deferred << 42
//        1 packed counts B: n alive, non-deferred | n good, non-deferred << 21
//        2 sum phi over good non-deferred   3 sum u alive non-deferred   4 sum u   5 sum u^2   6 sum x   7 sum x^2   8 sum y
//        9 sum y^2   10 sum w (good, non-deferred)   11 MAX (x-x_avg)^2+(y-y_avg)^2 over loss == 0 (all)
//        12 sum phi right after OP_FREQ_CHANGE (loss == 0; FIN_FREQ only)
// The counts are integers packed into the bit pattern of a slot (21-bit fields hold a thread's share of up to 2^32 particles;
// the block / grid folds unpack them into doubles).
enum { SEG_ACC = 13 };  // columns of the epilogue's shared-memory accumulator (>= KF_NACC, SEGL_NS)
enum { KF_CNTA = 0, KF_CNTB, KF_SPHI, KF_SU_ALIVE, KF_SU, KF_SU2, KF_SX, KF_SX2, KF_SY, KF_SY2, KF_SW, KF_R2MAX, KF_SPHI_FC, KF_NACC };
// totals after the grid fold (doubles): the unpacked counts follow the sums
enum { KT_NGOOD = 0, KT_NLLOSS, KT_NDEFER, KT_NALIVE_M, KT_NGOOD_M, KT_SPHI, KT_SU_ALIVE, KT_SU, KT_SU2, KT_SX, KT_SX2, KT_SY, KT_SY2,
       KT_SW, KT_R2MAX, KT_SPHI_FC };
struct FusedCtx {
  double pivot;
  double m;         // margin of the 2 pi wrap of the relative phase: |<phi>_good after marking - pivot| must stay below it
  double mc;        // margin of the marking threshold: |<phi>_good before marking - pivot| must stay below it
  double cand_thr, amb_thr, xa, ya;
};
struct FusedAcc {  // sums of one loop iteration (registers)
  unsigned long long ca, cb;
  double s[KF_NACC];
};

__device__ __forceinline__ void seg_accumulate_fused(FusedAcc& c, const Pt& p, u32 ll, double phi_fc, bool live_fc, const FusedCtx& f,
                                                     DeferList* dl, u32 local_idx) {
  const bool alive = p.loss == 0, good = alive && ll == 0;
  c.ca += (good ? 1ull : 0ull) | ((ll != 0 ? 1ull : 0ull) << 21);
  if (live_fc) c.s[KF_SPHI_FC] += phi_fc;
  // beam_kernel.cu:1556 marks `phi > <phi>_good + 3 pi && lloss != 1` whatever the transverse loss flag says
  const bool cand = ll != 1u && (p.phi - f.pivot > f.cand_thr);
  double u = 0.0;
  bool amb = false;
  if (alive) {
    u = rel_phi(p.phi, f.pivot);
    amb = !(fabs(u) <= f.amb_thr);  // also catches NaN
    const double tx = p.x - f.xa, ty = p.y - f.ya;
    c.s[KF_R2MAX] = fmax(c.s[KF_R2MAX], tx * tx + ty * ty);
  }
  if (cand || amb) {
    const u32 k = atomicAdd(&dl->count, 1u);
    if (k < HPS_DEFER_CAP) {
      DeferEntry& e = dl->e[k];
      e.phi = p.phi; e.x = p.x; e.y = p.y; e.w = p.w;
      e.tag = (unsigned long long)local_idx | ((unsigned long long)((alive ? 1u : 0u) | (ll != 0 ? 2u : 0u) | (cand ? 4u : 0u)) << 32);
    }
    c.ca += 1ull << 42;
  } else if (alive) {
    c.cb += 1ull | ((good ? 1ull : 0ull) << 21);
    c.s[KF_SU_ALIVE] += u;
    if (good) {
      c.s[KF_SPHI] += p.phi;
      c.s[KF_SU] += u; c.s[KF_SU2] = fma(u, u, c.s[KF_SU2]);
      c.s[KF_SX] += p.x; c.s[KF_SX2] = fma(p.x, p.x, c.s[KF_SX2]);
      c.s[KF_SY] += p.y; c.s[KF_SY2] = fma(p.y, p.y, c.s[KF_SY2]);
      c.s[KF_SW] += p.w;
    }
  }
}
// folds the iteration's sums into the thread's shared-memory column (fc: the FIN_FREQ slot is in use)
__device__ __forceinline__ void fused_flush(double* col, const FusedAcc& c, bool fc) {
  const int NF = KF_NACC - 1;
  double t[NF];
#pragma unroll
  for (int s = 0; s < NF; ++s) t[s] = lds_f64(col + s * HPS_BLOCK);
  t[KF_CNTA] = __longlong_as_double(__double_as_longlong(t[KF_CNTA]) + (long long)c.ca);
  t[KF_CNTB] = __longlong_as_double(__double_as_longlong(t[KF_CNTB]) + (long long)c.cb);
#pragma unroll
  for (int s = KF_SPHI; s < NF; ++s) t[s] = s == KF_R2MAX ? fmax(t[s], c.s[s]) : t[s] + c.s[s];
#pragma unroll
  for (int s = 0; s < NF; ++s) sts_f64(col + s * HPS_BLOCK, t[s]);
  if (fc) sts_f64(col + KF_SPHI_FC * HPS_BLOCK, lds_f64(col + KF_SPHI_FC * HPS_BLOCK) + c.s[KF_SPHI_FC]);
}
// the thread's column -> the 16 doubles that go through the block / grid fold
__device__ __forceinline__ void fused_unpack(const double* col, double (&v)[16]) {
  const unsigned long long ca = (unsigned long long)__double_as_longlong(col[KF_CNTA * HPS_BLOCK]);
  const unsigned long long cb = (unsigned long long)__double_as_longlong(col[KF_CNTB * HPS_BLOCK]);
  v[KT_NGOOD] = (double)(ca & 0x1fffffull); v[KT_NLLOSS] = (double)((ca >> 21) & 0x1fffffull); v[KT_NDEFER] = (double)(ca >> 42);
  v[KT_NALIVE_M] = (double)(cb & 0x1fffffull); v[KT_NGOOD_M] = (double)((cb >> 21) & 0x1fffffull);
  v[KT_SPHI] = col[KF_SPHI * HPS_BLOCK]; v[KT_SU_ALIVE] = col[KF_SU_ALIVE * HPS_BLOCK]; v[KT_SU] = col[KF_SU * HPS_BLOCK];
  v[KT_SU2] = col[KF_SU2 * HPS_BLOCK]; v[KT_SX] = col[KF_SX * HPS_BLOCK]; v[KT_SX2] = col[KF_SX2 * HPS_BLOCK];
  v[KT_SY] = col[KF_SY * HPS_BLOCK]; v[KT_SY2] = col[KF_SY2 * HPS_BLOCK]; v[KT_SW] = col[KF_SW * HPS_BLOCK];
  v[KT_R2MAX] = col[KF_R2MAX * HPS_BLOCK]; v[KT_SPHI_FC] = col[KF_SPHI_FC * HPS_BLOCK];
}

// The pilot: one ON-AXIS particle at (<phi>_good, <W>_good) of the previous kick through the segment's maps, by thread 0
// of every block ahead of the sweep.  Only a pivot comes out of it, so it runs the fast forms of the maps specialised
// for x = x' = y = y' = 0 (I0 = 1, I1 = 0: no Bessel functions, no transverse update) -- about 1 us of one thread.
__device__ __noinline__ double pilot_phase(const Op* ops, const OpC* opcs, int nops, double phi, double w, double mass, double charge) {
  const double inv_mass = 1.0 / mass;
  for (int k = 0; k < nops; ++k) {
    const Op& op = ops[k];
    const double* q = op.p;
    const double* c = opcs[k].c;
    switch (op.type) {
      case OP_DRIFT: phi += c[D_PH_OVER_WL] * inv_beta_of(w, inv_mass); break;
      case OP_QUAD_HALF: phi += c[Q_PH_OVER_WL] * inv_beta_of(w, inv_mass); break;
      case OP_GAP_FIRST: {
        const double ib = inv_beta_of(w, inv_mass);
        if (q[G_AMP] == 0.0) phi += c[GF_DD_OVER_WL] * ib;
        else if (op.flag == 0) phi += (c[GC_BETA_IN] * ib) * c[GC_PI_CLBL_DPS] + c[GF_DCLQ_OVER_WL] * ib;
        else phi += (1.0 - c[GC_BETAG_CCL] * (c[GF_IBETA_IN] - ib)) * c[GC_PI_CLBL_DPS];
        break;
      }
      case OP_GAP_SECOND: {  // op_gap_second_f on the axis
        const bool ccl = op.flag != 0;
        const double ib1 = inv_beta_of(w, inv_mass), beta = 1.0 / ib1;
        double bb = beta;
        if (beta > 1.1 * q[G_FBC]) bb = q[G_FBC];
        else if (beta < q[G_FBMIN]) bb = q[G_FBMIN];
        const double t_fac = q[G_T0] + bb * (q[G_T1] + bb * (q[G_T2] + bb * (q[G_T3] + bb * (q[G_T4] + bb * q[G_T5]))));
        const double coef = bb * bb * c[GF_INV_TTF_DEN];
        const double tp_fac = coef * (q[G_T1] + bb * (2 * q[G_T2] + bb * (3 * q[G_T3] + bb * (4 * q[G_T4] + 5 * bb * q[G_T5]))));
        const double sp_fac = -coef * (q[G_S1] + bb * (2 * q[G_S2] + bb * (3 * q[G_S3] + bb * (4 * q[G_S4] + 5 * bb * q[G_S5]))));
        if (q[G_AMP] == 0.0 || (ccl && beta < q[G_FBMIN])) { phi += c[GF_DD_OVER_WL] * ib1; break; }
        double tw = t_fac;
        if (ccl) tw = t_fac - K_TWOPI * tp_fac * (c[GC_BETA_IN] * ib1 - 1.0) * c[GF_BETAG_OVER_BIN];
        double phi1 = phi + q[G_PHSHIFT];
        if (ccl) phi1 += c[GC_PH_CONST] - c[GF_DG_OVER_WL] * ib1;
        // (a pivot needs 1e-4 rad, not 1e-16: the phase is reduced in fp64 and its sine / cosine come from the SFU)
        auto fast_sincos = [](double a, double& sn, double& cs) {
          const double r = a - K_TWOPI * rint(a * (1.0 / K_TWOPI));
          float fs, fc;
          __sincosf((float)r, &fs, &fc);
          sn = (double)fs; cs = (double)fc;
        };
        double sinp1, cosp1;
        fast_sincos(phi1, sinp1, cosp1);
        const double prome1 = -c[GC_CON1] * (tp_fac * (sinp1 - c[GC_SINPS]) + sp_fac * (cosp1 - c[GC_COSPS]));
        fast_sincos(phi1 + prome1, sinp1, cosp1);
        w += c[GC_AQ_EFL] * tw * cosp1;
        const double ib = inv_beta_of(w, inv_mass);
        const double prome = -2.0 * c[GC_CON1] * tp_fac * (sinp1 - c[GC_SINPS]);
        if (ccl) phi += (1.0 - c[GC_BETAG_CCL] * (c[GF_IBETA_OUT] - ib)) * c[GC_PI_CLBL_DPS] + prome;
        else phi += (c[GC_BETA_OUT] * ib) * c[GC_PI_CLBL_DPS] + prome + c[GF_DCLQ_OVER_WL] * ib;
        break;
      }
      case OP_DIPOLE_FIRST: case OP_DIPOLE_SECOND: case OP_BUNCHER: {
        Pt p = {0.0, 0.0, 0.0, 0.0, phi, w, 0u};
        if (op.type == OP_BUNCHER) op_buncher(p, q, op.elem, mass, charge); else op_dipole(p, q, op.type == OP_DIPOLE_SECOND, mass);
        phi = p.phi; w = p.w;
        break;
      }
      case OP_FREQ_CHANGE: phi *= q[0]; break;
      default: break;  // kick (no field on the axis at the bunch centre), rotation, steerer, apertures
    }
  }
  return phi;
}

// Two-limb fixed point: a summand v, pre-scaled by a power of two into |v| < 8, becomes hi = rint(v 2^40) and
// lo = rint((v - hi 2^-40) 2^80) -- the remainder is exact in fp64 -- and the limbs are summed as 64-bit integers:
// 2^19 summands (8 ranks x 2^16 deferred particles) stay below 2^63, the resolution is 2^-80 of the pre-scaled value
// (finer than one ulp of any summand that matters), and integer addition does not care about the order.
struct Fix2 {
  long long hi, lo;
};
__device__ __forceinline__ void fix_add(long long& hi, long long& lo, double v) {
  const double h = rint(v * 0x1p40);
  hi += (long long)h;
  lo += __double2ll_rn(fma(-h, 0x1p-40, v) * 0x1p80);
}
__device__ __forceinline__ double fix_value(long long hi, long long lo) { return (double)hi * 0x1p-40 + (double)lo * 0x1p-80; }

// sum of NV 64-bit integers per thread over the block -> s_out[0..NV) (shared); every thread of the block calls it
template <int NV>
__device__ __forceinline__ void block_sum_i64(long long (&v)[NV], long long* s_warp /* shared [8][NV] */, long long* s_out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int s = 0; s < NV; ++s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[s] += __shfl_xor_sync(0xffffffffu, v[s], o);
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NV; ++s) s_warp[wid * NV + s] = v[s];
  }
  __syncthreads();
  if ((int)threadIdx.x < NV) {
    long long a = 0;
    for (int k = 0; k < nw; ++k) a += s_warp[k * NV + threadIdx.x];
    s_out[threadIdx.x] = a;
  }
  __syncthreads();
}

__device__ void scheff_policy(DevStats* st, ScDev* sc, const KickParams& kp, double r2max);
__device__ void finalize_moments(const double* t, double ref, u32 n_global, DevStats* st, ScDev* sc, const KickParams& kp, int do_policy);

// pre-scales (exact powers of two) that bring a deferred particle's summands into |v| < 8
#define PS_DPHI 0x1p-9   /* |phi - P| < 2^12 rad (checked) */
#define PS_W 0x1p-7      /* w < 2^10 MeV (checked) */
#define PS_PR2 0x1p-1    /* phi_r^2 <= pi^2 */

// Last block of a FIN_KICK_FUSED sweep; every thread of the block calls it with this rank's totals in t[0..16) (shared,
// KT_* layout).  Ends with DevStats / ScDev exactly as Beam::UpdateStatForSpaceCharge + Scheff::Start's policy leave them,
// or with st->moments_fallback != 0 (1: redo the whole chain, 2: longitudinal losses are marked, redo the moment sums only).
// This block is the serial tail of every kick, so it avoids L2 round trips: DevStats and ScDev are copied to shared
// memory up front, every scalar step works on the copies and they are written back at the end.  (Keeping a thread's
// side-list entries in registers across the three passes was tried: no gain -- the tail runs once per launch from a cold
// instruction cache, so its code size matters more than its loads.)
__device__ void fused_kick_finalize(const double* t, const FusedCtx& f, const BeamView& b, DevStats* st_g, ScDev* sc_g, double* bins,
                                    const KickParams& kp, DeferList* dl, const XrView& xr, int fin_mode) {
  constexpr int NI = 20;
  __shared__ long long s_warp[(HPS_BLOCK / 32) * NI];
  __shared__ long long s_i[NI];
  __shared__ double s_g[HPS_XR_NS];   // round 1: totals over all ranks
  __shared__ double s_x[HPS_XR_NS];   // rounds 2 and 3
  __shared__ double s_R1, s_R2, s_SG0, s_nG1, s_nnew;
  __shared__ u32 s_code, s_Kall, s_ncand;
  __shared__ DevStats s_st;
  __shared__ ScDev s_sc;
  DevStats* const st = &s_st;
  ScDev* const sc = &s_sc;
  const bool p2p = xr.mode == XR_P2P;
  const double P = f.pivot;
  unsigned long long tq = threadIdx.x == 0 ? global_ns() : 0ull;
#define FIN_STAMP(k) if (threadIdx.x == 0) { const unsigned long long tn = global_ns(); st->fin_ns[k] += tn - tq; tq = tn; }
  const u32 cnt = __ldcg(&dl->count);
  {  // shared copies of the two state structs (8-byte words; one L2 round trip, overlapped with pass 1's loads)
    const int nw_st = (int)(sizeof(DevStats) / 8), nw_sc = (int)(sizeof(ScDev) / 8);
    const unsigned long long* gs = reinterpret_cast<const unsigned long long*>(st_g);
    const unsigned long long* gc = reinterpret_cast<const unsigned long long*>(sc_g);
    unsigned long long* ls = reinterpret_cast<unsigned long long*>(&s_st);
    unsigned long long* lc = reinterpret_cast<unsigned long long*>(&s_sc);
    for (int k = threadIdx.x; k < nw_st + nw_sc; k += blockDim.x) {
      if (k < nw_st) ls[k] = __ldcg(gs + k); else lc[k - nw_st] = __ldcg(gc + (k - nw_st));
    }
  }
  const u32 K = min(cnt, HPS_DEFER_CAP);
  // ---- pass 1: the deferred members of the good set, the candidates
  {
    long long v[5] = {0, 0, 0, 0, (threadIdx.x == 0 && cnt > HPS_DEFER_CAP) ? 1 : 0};  // n good, sum (phi - P) hi lo, candidates, bad
    auto one = [&](u32 fl, double phi) {
      const double d = phi - P;
      if ((fl & 3u) == 1u) {
        v[0] += 1;
        if (!(fabs(d) < 4096.0)) v[4] += 1; else fix_add(v[1], v[2], d * PS_DPHI);
      }
      if (fl & 4u) v[3] += 1;
    };
#pragma unroll 1
    for (u32 k = threadIdx.x; k < K; k += blockDim.x) one((u32)(__ldcg(&dl->e[k].tag) >> 32), __ldcg(&dl->e[k].phi));
    if (cnt != 0u) block_sum_i64<5>(v, s_warp, s_i);  // (block-uniform condition)
    else if (threadIdx.x < 5) s_i[threadIdx.x] = 0;
  }
  __syncthreads();
  FIN_STAMP(0)
  if (threadIdx.x < 16) s_g[threadIdx.x] = t[threadIdx.x];
  if (threadIdx.x == 0) {
    s_g[16] = (double)s_i[0]; s_g[17] = __longlong_as_double(s_i[1]); s_g[18] = __longlong_as_double(s_i[2]);
    s_g[19] = (double)s_i[3]; s_g[20] = (double)s_i[4]; s_g[21] = (double)K;
  }
  __syncthreads();
  if (p2p) xr_exchange_p2p(s_g, 22, 1u << KT_R2MAX, xr, (1u << 17) | (1u << 18));
  if (threadIdx.x == 0) {
    // Beam::ChangeFrequency -> UpdateAvgPhi() (beam.cpp:1666; divides by N - num_loss, stale count)
    if (fin_mode & FIN_FREQ) st->v[HPS_STAT_PHI_AVG] = s_g[KT_SPHI_FC] / (double)(u32)(b.n_global - st->num_loss);
    const double n_g0 = s_g[KT_NGOOD], nl = s_g[16];
    const double sg0 = s_g[KT_SPHI] + nl * P + fix_value(__double_as_longlong(s_g[17]), __double_as_longlong(s_g[18])) * (1.0 / PS_DPHI);
    const double r1 = n_g0 > 0.0 ? sg0 / n_g0 : -99999.0;
    u32 code = 0u;
    if (s_g[20] != 0.0 || n_g0 != s_g[KT_NGOOD_M] + nl) { code = 1u; st->n_fb[0]++; }   // list overflow / out-of-range entry
    // the candidates are the particles ABOVE pivot + 3 pi - mc: they include everything above the marking threshold
    // <phi>_good + 3 pi as long as <phi>_good has not come out BELOW pivot - mc
    else if (!(n_g0 > 0.0) || !(r1 - P >= -f.mc)) { code = 1u; st->n_fb[1]++; }
    s_R1 = r1; s_R2 = r1; s_SG0 = sg0; s_nG1 = n_g0; s_nnew = 0.0; s_code = code; s_Kall = (u32)s_g[21]; s_ncand = (u32)s_g[19];
  }
  __syncthreads();
  FIN_STAMP(1)
  if (s_code == 0u && s_ncand != 0u) {
    // ---- pass 2: Beam::UpdateLongitudinalLoss (beam_kernel.cu:1549-1559) on the candidates
    const double thr = s_R1 + K_TWOPI + K_PI;
    long long v[4] = {0, 0, 0, 0};  // newly non-zero lloss, marked members of the good set, their sum (phi - P) hi lo
#pragma unroll 1
    for (u32 k = threadIdx.x; k < K; k += blockDim.x) {
      const unsigned long long tag = __ldcg(&dl->e[k].tag);
      const u32 fl = (u32)(tag >> 32);
      const double phi = __ldcg(&dl->e[k].phi);
      if ((fl & 4u) && phi > thr) {
        b.lloss[(u32)tag] = 1u;
        if (!(fl & 2u)) v[0] += 1;
        if ((fl & 3u) == 1u) { v[1] += 1; fix_add(v[2], v[3], (phi - P) * PS_DPHI); }
        dl->e[k].tag = tag | (2ull << 32);
      }
    }
    if (K != 0u) block_sum_i64<4>(v, s_warp, s_i);
    else if (threadIdx.x < 4) s_i[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) { s_x[0] = (double)s_i[0]; s_x[1] = (double)s_i[1]; s_x[2] = __longlong_as_double(s_i[2]); s_x[3] = __longlong_as_double(s_i[3]); }
    __syncthreads();
    if (p2p) xr_exchange_p2p(s_x, 4, 0u, xr, (1u << 2) | (1u << 3));
    if (threadIdx.x == 0) {
      const double n_mark = s_x[1];
      const double s_mark = n_mark * P + fix_value(__double_as_longlong(s_x[2]), __double_as_longlong(s_x[3])) * (1.0 / PS_DPHI);
      const double n_g1 = s_g[KT_NGOOD] - n_mark;
      s_nnew = s_x[0];
      s_nG1 = n_g1;
      s_R2 = n_g1 > 0.0 ? (s_SG0 - s_mark) / n_g1 : -99999.0;
      if (s_x[0] > 0.0) sc->n_lloss_pass++;
      if (!(n_g1 > 0.0)) s_code = 2u;
    }
    __syncthreads();
  }
  FIN_STAMP(2)
  const double R2 = s_R2, delta = R2 - P;
  if (threadIdx.x == 0 && s_code == 0u && !(fabs(delta) <= f.m)) { s_code = 2u; st->n_fb[2]++; }  // (marks are written: only the sums are redone)
  __syncthreads();
  // ---- pass 3: the deferred particles' relative phases about the true <phi>_good, summed in fixed point
  // v: 0 n alive, 1 n good, 2 bad, 3-4 sum pr alive, 5-6 sum pr, 7-8 pr^2, 9-10 x, 11-12 x^2, 13-14 y, 15-16 y^2, 17-18 w
  if (s_code == 0u && s_Kall != 0u) {
    long long v[NI];
#pragma unroll
    for (int k = 0; k < NI; ++k) v[k] = 0;
    auto one = [&](u32 fl, double phi, double x, double y, double w) {
      if (!(fl & 1u)) return;
      const double pr = rel_phi(phi, R2);
      v[0] += 1; fix_add(v[3], v[4], pr);
      if (fl & 2u) return;
      if (!(fabs(x) < 2.0 && fabs(y) < 2.0 && fabs(w) < 1024.0)) { v[2] += 1; return; }
      v[1] += 1;
      fix_add(v[5], v[6], pr); fix_add(v[7], v[8], pr * pr * PS_PR2);
      fix_add(v[9], v[10], x); fix_add(v[11], v[12], x * x); fix_add(v[13], v[14], y); fix_add(v[15], v[16], y * y);
      fix_add(v[17], v[18], w * PS_W);
    };
#pragma unroll 1
    for (u32 k = threadIdx.x; k < K; k += blockDim.x) {
      const DeferEntry& e = dl->e[k];
      one((u32)(__ldcg(&e.tag) >> 32), __ldcg(&e.phi), __ldcg(&e.x), __ldcg(&e.y), __ldcg(&e.w));
    }
    if (K != 0u) block_sum_i64<NI>(v, s_warp, s_i);
    else if (threadIdx.x < NI) s_i[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < NI) s_x[threadIdx.x] = __longlong_as_double(s_i[threadIdx.x]);
    __syncthreads();
    if (p2p) xr_exchange_p2p(s_x, NI, 0u, xr, (1u << NI) - 1u);
  } else {
    if (threadIdx.x < NI) s_x[threadIdx.x] = 0.0;  // bit pattern of the integer 0
    __syncthreads();
  }
  FIN_STAMP(3)
  if (threadIdx.x == 0) {
    auto I = [&](int k) { return __double_as_longlong(s_x[k]); };
    auto F = [&](int k) { return fix_value(I(k), I(k + 1)); };
    u32 code = s_code;
    if (code == 0u && (I(2) != 0 || s_nG1 != s_g[KT_NGOOD_M] + (double)I(1))) code = 2u;
    if (code != 1u) {  // what the first links of the chain leave behind (beam.cpp:776-782, 1007-1009)
      st->num_good = (u32)s_nG1;
      st->v[HPS_STAT_PHI_AVG_GOOD] = R2;
      st->num_lloss = (u32)(s_g[KT_NLLOSS] + s_nnew);
    }
    st->need_lloss_pass = 0u;
    st->moments_fallback = code;
    const double err = fabs(delta), err1 = fabs(s_R1 - P);
    // the margins follow the prediction errors: generous against the last miss, decaying slowly
    st->pivot_mc = (err1 == err1) ? fmin(0.5, fmax(fmax(1.0e-3, 0.5 * f.mc), 4.0 * err1)) : 0.1;
    if (code == 0u) {
      const double nm = s_g[KT_NGOOD_M], na = s_g[KT_NALIVE_M];
      double m9[9];
      m9[0] = s_g[KT_SU_ALIVE] - na * delta + F(3);
      m9[1] = s_g[KT_SU] - nm * delta + F(5);
      m9[2] = s_g[KT_SU2] - 2.0 * delta * s_g[KT_SU] + nm * delta * delta + F(7) * (1.0 / PS_PR2);
      m9[3] = s_g[KT_SX] + F(9);
      m9[4] = s_g[KT_SX2] + F(11);
      m9[5] = s_g[KT_SY] + F(13);
      m9[6] = s_g[KT_SY2] + F(15);
      m9[7] = s_g[KT_SW] + F(17) * (1.0 / PS_W);
      m9[8] = s_g[KT_R2MAX];
      finalize_moments(m9, R2, b.n_global, st, sc, kp, 1);
      st->n_fused++;
      st->n_deferred += (unsigned long long)s_Kall;
      st->pivot_err_max = fmax(st->pivot_err_max, fmax(err, err1));
      // (a miss is usually a burst of markings that moved <phi>_good in one kick: it says little about the next kick,
      //  and a wide margin defers a share 2 m / 2 pi of all out-of-bucket particles -- so: small factor, fast decay)
      // (floor 1e-3 rad: the prediction error of a quiet kick is 1e-5 .. 1e-3, and a wrap band of +-1e-3 rad defers only
      //  3e-4 of the out-of-bucket particles -- with a floor of 1e-4, 15-20 % of the DTL kicks missed the margin)
      st->pivot_m = fmin(0.25, fmax(fmax(1.0e-3, 0.25 * f.m), 4.0 * err));
    } else {
      // (the wrap margin is left alone: a miss is a burst of markings that moved <phi>_good within ONE kick -- measured:
      //  growing the margin after it defers so many particles on the next kicks that the side list overflows)
      st->n_fallback++;
      if (!(err == err)) st->pivot_m = 0.05;
    }
    dl->count = 0u;
    s_code = code;
  }
  __syncthreads();
  FIN_STAMP(4)
  if (s_code == 0u && sc->remesh) {  // ResetBinTbl (space_charge.cu:92-105)
    const u32 nb = kp.nr * kp.nz;
    for (u32 k = threadIdx.x; k < nb; k += blockDim.x) bins[k] = 0.0;
  }
  FIN_STAMP(5)
#undef FIN_STAMP
  __syncthreads();
  {  // write the state back (diag_ns / diag_t0 are updated by the caller on the global copy right after)
    const int nw_st = (int)(sizeof(DevStats) / 8), nw_sc = (int)(sizeof(ScDev) / 8);
    unsigned long long* gs = reinterpret_cast<unsigned long long*>(st_g);
    unsigned long long* gc = reinterpret_cast<unsigned long long*>(sc_g);
    const unsigned long long* ls = reinterpret_cast<const unsigned long long*>(&s_st);
    const unsigned long long* lc = reinterpret_cast<const unsigned long long*>(&s_sc);
    for (int k = threadIdx.x; k < nw_st + nw_sc; k += blockDim.x) {
      if (k < nw_st) gs[k] = ls[k]; else gc[k - nw_st] = lc[k - nw_st];
    }
  }
  __syncthreads();
}

// ---- per-thread asynchronous staging of the NEXT particle pair ------------------------------------------------------
// At 16 Mi particles the beam (0.94 GB) streams from HBM: a thread loads its pair (~1-2 us under load), computes for
// ~10 us, stores, loads the next pair -- with 4 warps per scheduler that load latency is exposed (ncu: long_scoreboard
// 20-30 % of the stall samples, profiles/r2a_ncu_k_segment_16mi.txt), and a register prefetch of 28 words does not fit the
// 128-register budget.  Each thread therefore copies its OWN next pair into its own shared-memory slots with cp.async
// (LDGSTS: no registers, no block barrier -- a thread only ever reads what it copied itself) while it works on the
// current pair, and picks it up with cp.async.wait_group at the top of the next iteration.
struct PairStage {
  double2 d[6][HPS_BLOCK];  // x xp y yp phi w
  uint2 f[2][HPS_BLOCK];    // loss lloss
};
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
// HPS_EVICT_FIRST: the particle stream of a staged sweep (a beam larger than the L2) is marked evict-first in the L2
// (cp.async with an L2 cache-hint policy, streaming stores).  It does not come back before the whole beam has passed, and
// unmarked it pushes everything else out of the L2 -- including the INSTRUCTIONS of the sweep's serial tail (block fold,
// side list, Scheff policy), which one block executes once per launch: measured, the policy phase of the finaliser takes
// 2.5 us while the beam fits the L2 and 6.8 us once it streams.
#ifndef HPS_EVICT_FIRST
#define HPS_EVICT_FIRST 1
#endif
__device__ __forceinline__ unsigned long long evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
#if HPS_EVICT_FIRST
__device__ __forceinline__ void cp_async16(void* smem, const void* g, unsigned long long pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(g), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g, unsigned long long pol) {
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(g), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream(double2* p, double a, double b) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_stream(uint2* p, u32 a, u32 b) {
  asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
#else
__device__ __forceinline__ void cp_async16(void* smem, const void* g, unsigned long long) { cp_async16(smem, g); }
__device__ __forceinline__ void cp_async8(void* smem, const void* g, unsigned long long) { cp_async8(smem, g); }
__device__ __forceinline__ void st_stream(double2* p, double a, double b) { *p = make_double2(a, b); }
__device__ __forceinline__ void st_stream(uint2* p, u32 a, u32 b) { *p = make_uint2(a, b); }
#endif
__device__ __forceinline__ void stage_issue(PairStage* s, const BeamView& b, u32 i2, unsigned long long pol) {
  const int t = threadIdx.x;
  cp_async16(&s->d[0][t], reinterpret_cast<const double2*>(b.x) + i2, pol);
  cp_async16(&s->d[1][t], reinterpret_cast<const double2*>(b.xp) + i2, pol);
  cp_async16(&s->d[2][t], reinterpret_cast<const double2*>(b.y) + i2, pol);
  cp_async16(&s->d[3][t], reinterpret_cast<const double2*>(b.yp) + i2, pol);
  cp_async16(&s->d[4][t], reinterpret_cast<const double2*>(b.phi) + i2, pol);
  cp_async16(&s->d[5][t], reinterpret_cast<const double2*>(b.w) + i2, pol);
  cp_async8(&s->f[0][t], reinterpret_cast<const uint2*>(b.loss) + i2, pol);
  cp_async8(&s->f[1][t], reinterpret_cast<const uint2*>(b.lloss) + i2, pol);
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_issue(PairStage* s, const BeamView& b, u32 i2) {
  const int t = threadIdx.x;
  cp_async16(&s->d[0][t], reinterpret_cast<const double2*>(b.x) + i2);
  cp_async16(&s->d[1][t], reinterpret_cast<const double2*>(b.xp) + i2);
  cp_async16(&s->d[2][t], reinterpret_cast<const double2*>(b.y) + i2);
  cp_async16(&s->d[3][t], reinterpret_cast<const double2*>(b.yp) + i2);
  cp_async16(&s->d[4][t], reinterpret_cast<const double2*>(b.phi) + i2);
  cp_async16(&s->d[5][t], reinterpret_cast<const double2*>(b.w) + i2);
  cp_async8(&s->f[0][t], reinterpret_cast<const uint2*>(b.loss) + i2);
  cp_async8(&s->f[1][t], reinterpret_cast<const uint2*>(b.lloss) + i2);
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// the same for the one-particle-per-thread sweeps (8- and 4-byte copies; the slots live in the PairStage area)
struct OneStage {
  double d[6][HPS_BLOCK];  // x xp y yp phi w
  u32 f[2][HPS_BLOCK];     // loss lloss
};
__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((u32)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void stage_issue1(OneStage* s, const BeamView& b, u32 i) {
  const int t = threadIdx.x;
  cp_async8(&s->d[0][t], b.x + i);
  cp_async8(&s->d[1][t], b.xp + i);
  cp_async8(&s->d[2][t], b.y + i);
  cp_async8(&s->d[3][t], b.yp + i);
  cp_async8(&s->d[4][t], b.phi + i);
  cp_async8(&s->d[5][t], b.w + i);
  cp_async4(&s->f[0][t], b.loss + i);
  cp_async4(&s->f[1][t], b.lloss + i);
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// PPT = particles per thread.  PPT == 2 moves two adjacent particles with one 128-bit access per
// array; PPT == 1 halves the register footprint and doubles the resident warps (the maps are chains of
// dependent fp64 operations, so the sweep is latency-bound before it is bandwidth-bound).
// (A TMA-staged, double-buffered variant of the PPT == 2 loop was measured slower -- 35.0 vs 32.0 ms of sweep time
// per C2 pass: the sweep is not load-bound and the per-tile block barrier re-aligns the warps -- and was removed.)
// SEQ: the launch's op sequence when it is one of the ten that make up 99.8 % of the sweeps of a full-linac run
// (tests/host_maps.cu `plan` with HPS_PLAN_PATTERNS=1: 4 855 x [kick, CCL gap second half, next CCL gap first half],
// 3 030 x [kick, drift, drift], ... of 9 108), packed 8 bits per op: type | (gap flavour + 1) << 5.  The sweep is then
// compiled for exactly that sequence -- no loop, no type switch, the DTL / CCL branches of the gap maps folded, the
// scheduler free to overlap the tail of one map with the head of the next.  0 = any sequence (loop + switch).
// Measured on C4 (16 Mi particles): 754 -> 687 us per sweep with the first two patterns alone.
#define SQ(type, flavour) ((type) | (((flavour) + 1) << 5))
#define SEQ2(a, b) ((a) | ((b) << 8))
#define SEQ3(a, b, c) ((a) | ((b) << 8) | ((c) << 16))
enum {
  SEQ_ANY = 0,
  SEQ_CCL_CELL = SEQ3(SQ(OP_SC_KICK, -1), SQ(OP_GAP_SECOND, 1), SQ(OP_GAP_FIRST, 1)),
  SEQ_DRIFT_STEP = SEQ3(SQ(OP_SC_KICK, -1), SQ(OP_DRIFT, -1), SQ(OP_DRIFT, -1)),
  SEQ_DRIFT_QUAD = SEQ3(SQ(OP_SC_KICK, -1), SQ(OP_DRIFT, -1), SQ(OP_QUAD_HALF, -1)),
  SEQ_QUAD_DRIFT = SEQ3(SQ(OP_SC_KICK, -1), SQ(OP_QUAD_HALF, -1), SQ(OP_DRIFT, -1)),
  SEQ_QUAD_GAP1D = SEQ3(SQ(OP_SC_KICK, -1), SQ(OP_QUAD_HALF, -1), SQ(OP_GAP_FIRST, 0)),
  SEQ_GAP2D_QUAD = SEQ3(SQ(OP_SC_KICK, -1), SQ(OP_GAP_SECOND, 0), SQ(OP_QUAD_HALF, -1)),
  SEQ_APER_DRIFT = SEQ2(SQ(OP_CAPERTURE, -1), SQ(OP_DRIFT, -1)),
  SEQ_KICK_DRIFT = SEQ2(SQ(OP_SC_KICK, -1), SQ(OP_DRIFT, -1)),
  SEQ_KICK_GAP2C = SEQ2(SQ(OP_SC_KICK, -1), SQ(OP_GAP_SECOND, 1)),
  SEQ_APER_GAP1C = SEQ2(SQ(OP_CAPERTURE, -1), SQ(OP_GAP_FIRST, 1)),
};
template <int PPT, int MINB, bool FULL, int MATH, int SEQ = SEQ_ANY>
__global__ void __launch_bounds__(HPS_BLOCK, MINB)
k_segment(BeamView b, const Op* __restrict__ ops, const OpC* __restrict__ opcs, int nops, int fin_mode, int dirty, int compat,
          DevStats* st, ScDev* sc, const double2* __restrict__ field, KickParams kp, double* partials,
          u32* ticket, double* bins, DeferList* dl, int stage, const __grid_constant__ XrView xr) {
  __shared__ KickCtx s_k;
  __shared__ FusedCtx s_f;
  __shared__ double s_cxy[2];
  __shared__ double s_total[SEG_NS];
  __shared__ double s_scratch[HPS_BLOCK];
  // dynamic shared memory, sized per launch (segment_smem): [SEG_ACC][HPS_BLOCK] partial sums of the epilogue, one column
  // per thread | the cp.async stage of the next particle pair (stage != 0 only) | the segment's op records and constants
  extern __shared__ double s_acc[];
  Op* const s_ops = reinterpret_cast<Op*>(s_acc + SEG_ACC * HPS_BLOCK + (stage ? sizeof(PairStage) / sizeof(double) : 0));
  OpC* const s_opc = reinterpret_cast<OpC*>(s_ops + nops);

  pdl_launch_dependents();
  pdl_wait();  // the op constants (k_prepare_ops) and everything below come from preceding kernels
  const bool fused = (fin_mode & 3) == FIN_KICK_FUSED;
  // the first particle pair starts its way from HBM now: its latency hides behind the prologue (op staging, kick
  // factors, pilot particle)
  if (stage) {
    const u32 first = blockIdx.x * blockDim.x + threadIdx.x;
    if (PPT == 2) { if (first < ((b.n + 1) >> 1)) stage_issue(reinterpret_cast<PairStage*>(s_acc + SEG_ACC * HPS_BLOCK), b, first, evict_first_policy()); }
    else if (SEQ != SEQ_ANY) { if (first < b.n) stage_issue1(reinterpret_cast<OneStage*>(s_acc + SEG_ACC * HPS_BLOCK), b, first); }
  }
  {  // stage the segment's op records (block-uniform) in shared memory
    const int nw = nops * (int)(sizeof(Op) / sizeof(double));
    const double* src = reinterpret_cast<const double*>(ops);
    double* dst = reinterpret_cast<double*>(s_ops);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
    const int nc = nops * HPS_OPC_N;
    const double* csrc = reinterpret_cast<const double*>(opcs);
    double* cdst = reinterpret_cast<double*>(s_opc);
    for (int i = threadIdx.x; i < nc; i += blockDim.x) cdst[i] = csrc[i];
  }
  double* const col = s_acc + threadIdx.x;  // this thread's column of partial sums
  if (fin_mode != FIN_NONE) {
#pragma unroll
    for (int s = 0; s < SEG_ACC; ++s) col[s * HPS_BLOCK] = 0.0;
    if (!fused) col[2 * HPS_BLOCK] = -DBL_MAX;
  }
  if (threadIdx.x == 0) {
    s_cxy[0] = st->v[HPS_STAT_X_AVG];
    s_cxy[1] = st->v[HPS_STAT_Y_AVG];
    s_k.phimc = st->v[HPS_STAT_PHI_AVG_R];
    s_k.xbar = st->v[HPS_STAT_X_AVG_GOOD];
    s_k.ybar = st->v[HPS_STAT_Y_AVG_GOOD];
    s_k.epsq = st->v[HPS_STAT_X_SIG_GOOD] / st->v[HPS_STAT_Y_SIG_GOOD];
    s_k.phi_ref = st->v[HPS_STAT_PHI_AVG_GOOD];
    if (sc) {
      s_k.dr = sc->dr; s_k.dz = sc->dz; s_k.beta = sc->beta;
      s_k.ratio_r = sc->ratio_r; s_k.ratio_z = sc->ratio_z; s_k.ratio_q = sc->ratio_q; s_k.ratio_g = sc->ratio_g;
      s_k.adj_bunch = sc->adj_bunch; s_k.num_on_mesh = sc->num_on_mesh_all;
    }
    s_k.nr = kp.nr; s_k.nz = kp.nz;
    if (sc && kp.mesh_w != 0.0 && sc->nr != 0u) { s_k.nr = sc->nr; s_k.nz = sc->nz; }
    // a segment holds at most one kick (its first op); its block-uniform factors are derived here
    for (int k = 0; k < nops; ++k)
      if (ops[k].type == OP_SC_KICK) { kick_prepare(s_k, ops[k].p, b.mass, b.n_global); break; }
  }
  __syncthreads();
  unsigned long long t_start = 0ull;
  if (fused) {
    if (threadIdx.x == 0) {
      t_start = global_ns();
      atomicMin(&st->diag_t0, t_start);
      FusedCtx f;
      f.pivot = pilot_phase(s_ops, s_opc, nops, st->v[HPS_STAT_PHI_AVG_GOOD], st->v[HPS_STAT_W_AVG_GOOD], b.mass, b.charge);
      double m = st->pivot_m, mc = st->pivot_mc;
      if (!(m >= 1.0e-4 && m <= 0.25)) m = 0.05;
      if (!(mc >= 1.0e-4 && mc <= 0.5)) mc = 0.1;
      f.m = m; f.mc = mc;
      f.cand_thr = K_TWOPI + K_PI - mc;
      f.amb_thr = K_PI - m;
      f.xa = s_cxy[0]; f.ya = s_cxy[1];
      s_f = f;
      if (blockIdx.x == 0) st->diag_ns[4] += global_ns() - t_start;
    }
    __syncthreads();
  }

  const double inv_mass = 1.0 / b.mass;

  if (PPT == 2) {
    const u32 npairs = (b.n + 1) >> 1;
    PairStage* stg = reinterpret_cast<PairStage*>(s_acc + SEG_ACC * HPS_BLOCK);
    const u32 stride = gridDim.x * blockDim.x;
    const u32 first = blockIdx.x * blockDim.x + threadIdx.x;
    for (u32 i2 = first; i2 < npairs; i2 += stride) {
      double2 X, XP, Y, YP, PH, W;
      uint2 L, LL;
      if (stage) {
        const int t = threadIdx.x;
        stage_wait();
        X = stg->d[0][t]; XP = stg->d[1][t]; Y = stg->d[2][t]; YP = stg->d[3][t]; PH = stg->d[4][t]; W = stg->d[5][t];
        L = stg->f[0][t]; LL = stg->f[1][t];
        // the next copy may overwrite the slots only after these reads have returned: its index depends on the last of them
        u32 zero;
        asm volatile("and.b32 %0, %1, 0;" : "=r"(zero) : "r"(LL.y ^ __double2hiint(W.y)));
        const u32 nx = i2 + stride + zero;
        if (nx < npairs) stage_issue(stg, b, nx, evict_first_policy());
      } else {
        X = reinterpret_cast<const double2*>(b.x)[i2];
        XP = reinterpret_cast<const double2*>(b.xp)[i2];
        Y = reinterpret_cast<const double2*>(b.y)[i2];
        YP = reinterpret_cast<const double2*>(b.yp)[i2];
        PH = reinterpret_cast<const double2*>(b.phi)[i2];
        W = reinterpret_cast<const double2*>(b.w)[i2];
        L = reinterpret_cast<const uint2*>(b.loss)[i2];
        LL = reinterpret_cast<const uint2*>(b.lloss)[i2];
      }
      Pt p[2] = {{X.x, XP.x, Y.x, YP.x, PH.x, W.x, L.x}, {X.y, XP.y, Y.y, YP.y, PH.y, W.y, L.y}};
      const u32 ll[2] = {LL.x, LL.y};
      const bool valid[2] = {true, 2 * i2 + 1 < b.n};
      const u32 g0 = b.offset + 2 * i2 * b.gstride;
      const u32 g[2] = {g0, g0 + b.gstride};
      double fc[2] = {0.0, 0.0};
      bool lf[2] = {false, false};
      // epilogue sums of this iteration (registers; see acc_flush)
      FusedAcc fa;
      fa.ca = 0ull; fa.cb = 0ull;
#pragma unroll
      for (int s = 0; s < KF_NACC; ++s) fa.s[s] = 0.0;
      double la[SEGL_NS] = {0.0, 0.0, -DBL_MAX, 0.0, 0.0, 0.0, 0.0, 0.0};
#ifndef HPS_SEQ_PARTICLE_MAJOR
#define HPS_SEQ_PARTICLE_MAJOR 0
#endif
      if (SEQ != SEQ_ANY && !HPS_SEQ_PARTICLE_MAJOR) {
        // Pattern sweep: the op sequence is compiled in, both particles of the pair go through each map back to back so
        // that their independent fp64 chains overlap.
#define AP2(TYPE, CCLF, K) apply_op<FULL, 2, MATH, TYPE, CCLF>(p, ll, g, valid, s_ops[K], s_opc[K], s_k, s_cxy, field, b, compat, fc, lf, inv_mass)
        AP2((SEQ & 31), ((SEQ >> 5) & 3) - 1, 0);
        if ((SEQ >> 8) & 31) AP2(((SEQ >> 8) & 31), ((SEQ >> 13) & 3) - 1, 1);
        if ((SEQ >> 16) & 31) AP2(((SEQ >> 16) & 31), ((SEQ >> 21) & 3) - 1, 2);
#undef AP2
      } else if (SEQ != SEQ_ANY) {
        // Particle-major form (build experiment, -DHPS_SEQ_PARTICLE_MAJOR=1): ONE copy of the op sequence in the
        // instruction stream, run for the first and then for the second particle of the pair.  Measured on C4 (16 Mi
        // particles): 871 vs 669 us per sweep -- the halved instruction footprint does not pay for the lost overlap of
        // the two particles' dependent fp64 chains at 16 warps per SM.
#pragma unroll 1
        for (int rep = 0; rep < 2; ++rep) {
          if (rep == 1 && !valid[1]) break;
          Pt q[1] = {rep ? p[1] : p[0]};
          const u32 ql[1] = {rep ? ll[1] : ll[0]};
          const u32 qg[1] = {rep ? g[1] : g[0]};
          const bool qv[1] = {true};
          double qfc[1] = {0.0};
          bool qlf[1] = {false};
#define AP1(TYPE, CCLF, K) apply_op<FULL, 1, MATH, TYPE, CCLF>(q, ql, qg, qv, s_ops[K], s_opc[K], s_k, s_cxy, field, b, compat, qfc, qlf, inv_mass)
          AP1((SEQ & 31), ((SEQ >> 5) & 3) - 1, 0);
          if ((SEQ >> 8) & 31) AP1(((SEQ >> 8) & 31), ((SEQ >> 13) & 3) - 1, 1);
          if ((SEQ >> 16) & 31) AP1(((SEQ >> 16) & 31), ((SEQ >> 21) & 3) - 1, 2);
#undef AP1
          if (fused) seg_accumulate_fused(fa, q[0], ql[0], qfc[0], qlf[0], s_f, dl, 2 * i2 + rep);
          else if (fin_mode != FIN_NONE) seg_accumulate(la, q[0], ql[0], qfc[0], qlf[0]);
          if (rep) p[1] = q[0]; else p[0] = q[0];
        }
      } else {
#define AP(TYPE, CCLF, K) apply_op<FULL, 2, MATH, TYPE, CCLF>(p, ll, g, valid, s_ops[K], s_opc[K], s_k, s_cxy, field, b, compat, fc, lf, inv_mass)
        for (int k = 0; k < nops; ++k) AP(-1, -1, k);
#undef AP
      }
      if (nops > 0) {
        if (stage) {  // a streaming beam: see HPS_EVICT_FIRST
          if (dirty & 1) st_stream(reinterpret_cast<double2*>(b.x) + i2, p[0].x, p[1].x);
          if (dirty & 2) st_stream(reinterpret_cast<double2*>(b.xp) + i2, p[0].xp, p[1].xp);
          if (dirty & 4) st_stream(reinterpret_cast<double2*>(b.y) + i2, p[0].y, p[1].y);
          if (dirty & 8) st_stream(reinterpret_cast<double2*>(b.yp) + i2, p[0].yp, p[1].yp);
          if (dirty & 16) st_stream(reinterpret_cast<double2*>(b.phi) + i2, p[0].phi, p[1].phi);
          if (dirty & 32) st_stream(reinterpret_cast<double2*>(b.w) + i2, p[0].w, p[1].w);
          if ((dirty & 64) && (p[0].loss != L.x || p[1].loss != L.y)) st_stream(reinterpret_cast<uint2*>(b.loss) + i2, p[0].loss, p[1].loss);
        } else {
          if (dirty & 1) reinterpret_cast<double2*>(b.x)[i2] = make_double2(p[0].x, p[1].x);
          if (dirty & 2) reinterpret_cast<double2*>(b.xp)[i2] = make_double2(p[0].xp, p[1].xp);
          if (dirty & 4) reinterpret_cast<double2*>(b.y)[i2] = make_double2(p[0].y, p[1].y);
          if (dirty & 8) reinterpret_cast<double2*>(b.yp)[i2] = make_double2(p[0].yp, p[1].yp);
          if (dirty & 16) reinterpret_cast<double2*>(b.phi)[i2] = make_double2(p[0].phi, p[1].phi);
          if (dirty & 32) reinterpret_cast<double2*>(b.w)[i2] = make_double2(p[0].w, p[1].w);
          if ((dirty & 64) && (p[0].loss != L.x || p[1].loss != L.y))
            reinterpret_cast<uint2*>(b.loss)[i2] = make_uint2(p[0].loss, p[1].loss);
        }
      }
      // the epilogue's sums of this pair: after the stores have been issued (the maps' temporaries are dead by now)
      if (!(SEQ != SEQ_ANY && HPS_SEQ_PARTICLE_MAJOR)) {
        if (fused) {
          seg_accumulate_fused(fa, p[0], ll[0], fc[0], lf[0], s_f, dl, 2 * i2);
          if (valid[1]) seg_accumulate_fused(fa, p[1], ll[1], fc[1], lf[1], s_f, dl, 2 * i2 + 1);
        } else if (fin_mode != FIN_NONE) {
          seg_accumulate(la, p[0], ll[0], fc[0], lf[0]);
          if (valid[1]) seg_accumulate(la, p[1], ll[1], fc[1], lf[1]);
        }
      }
      if (fused) fused_flush(col, fa, (fin_mode & FIN_FREQ) != 0);
      else if (fin_mode != FIN_NONE) acc_flush<SEGL_NS>(col, la, 1u << 2);
    }
  } else {
    OneStage* stg1 = reinterpret_cast<OneStage*>(s_acc + SEG_ACC * HPS_BLOCK);
    const bool staged = stage && SEQ != SEQ_ANY;  // (the generic one-particle sweeps serve the rare full-op-set launches)
    const u32 stride1 = gridDim.x * blockDim.x;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += stride1) {
      u32 l0, ll0;
      Pt p[1];
      if (staged) {
        const int t = threadIdx.x;
        stage_wait();
        p[0].x = stg1->d[0][t]; p[0].xp = stg1->d[1][t]; p[0].y = stg1->d[2][t]; p[0].yp = stg1->d[3][t];
        p[0].phi = stg1->d[4][t]; p[0].w = stg1->d[5][t];
        l0 = stg1->f[0][t]; ll0 = stg1->f[1][t];
        p[0].loss = l0;
        // the next copy may overwrite the slots only after these reads have returned: its index depends on the last of them
        u32 zero;
        asm volatile("and.b32 %0, %1, 0;" : "=r"(zero) : "r"(ll0 ^ l0 ^ __double2hiint(p[0].w) ^ __double2hiint(p[0].phi) ^ __double2hiint(p[0].x) ^
                                                            __double2hiint(p[0].xp) ^ __double2hiint(p[0].y) ^ __double2hiint(p[0].yp)));
        const u32 nx = i + stride1 + zero;
        if (nx < b.n) stage_issue1(stg1, b, nx);
      } else {
        l0 = b.loss[i]; ll0 = b.lloss[i];
        p[0].x = b.x[i]; p[0].xp = b.xp[i]; p[0].y = b.y[i]; p[0].yp = b.yp[i]; p[0].phi = b.phi[i]; p[0].w = b.w[i]; p[0].loss = l0;
      }
      const u32 ll[1] = {ll0};
      const bool valid[1] = {true};
      const u32 g[1] = {b.offset + i * b.gstride};
      double fc[1] = {0.0};
      bool lf[1] = {false};
      if (SEQ != SEQ_ANY) {
#define AP1(TYPE, CCLF, K) apply_op<FULL, 1, MATH, TYPE, CCLF>(p, ll, g, valid, s_ops[K], s_opc[K], s_k, s_cxy, field, b, compat, fc, lf, inv_mass)
        AP1((SEQ & 31), ((SEQ >> 5) & 3) - 1, 0);
        if ((SEQ >> 8) & 31) AP1(((SEQ >> 8) & 31), ((SEQ >> 13) & 3) - 1, 1);
        if ((SEQ >> 16) & 31) AP1(((SEQ >> 16) & 31), ((SEQ >> 21) & 3) - 1, 2);
#undef AP1
      } else
      for (int k = 0; k < nops; ++k) apply_op<FULL, 1, MATH>(p, ll, g, valid, s_ops[k], s_opc[k], s_k, s_cxy, field, b, compat, fc, lf, inv_mass);
      if (nops > 0) {
        if (dirty & 1) b.x[i] = p[0].x;
        if (dirty & 2) b.xp[i] = p[0].xp;
        if (dirty & 4) b.y[i] = p[0].y;
        if (dirty & 8) b.yp[i] = p[0].yp;
        if (dirty & 16) b.phi[i] = p[0].phi;
        if (dirty & 32) b.w[i] = p[0].w;
        if ((dirty & 64) && p[0].loss != l0) b.loss[i] = p[0].loss;
      }
      if (fused) {
        FusedAcc c;
        c.ca = 0ull; c.cb = 0ull;
#pragma unroll
        for (int s = 0; s < KF_NACC; ++s) c.s[s] = 0.0;
        seg_accumulate_fused(c, p[0], ll[0], fc[0], lf[0], s_f, dl, i);
        fused_flush(col, c, (fin_mode & FIN_FREQ) != 0);
      } else if (fin_mode != FIN_NONE) {
        double c[SEGL_NS] = {0.0, 0.0, -DBL_MAX, 0.0, 0.0, 0.0, 0.0, 0.0};
        seg_accumulate(c, p[0], ll[0], fc[0], lf[0]);
        acc_flush<SEGL_NS>(col, c, 1u << 2);
      }
    }
  }
  if (fin_mode == FIN_NONE) return;
  double v[SEG_NS];
  if (fused) {
    const unsigned long long t_loop = threadIdx.x == 0 ? global_ns() : 0ull;
    fused_unpack(col, v);
    if (grid_reduce<SEG_NS>(v, 1u << KT_R2MAX, partials, ticket, s_total, s_scratch, xr, false)) {
      const unsigned long long t_red = threadIdx.x == 0 ? global_ns() : 0ull;
      fused_kick_finalize(s_total, s_f, b, st, sc, bins, kp, dl, xr, fin_mode);
      if (threadIdx.x == 0) {
        const unsigned long long t_fin = global_ns(), t0 = __ldcg(&st->diag_t0);
        st->diag_ns[0] += t_loop - t0; st->diag_ns[1] += t_red - t_loop; st->diag_ns[2] += t_fin - t_red;
        st->diag_ns[3] += t_start - t0; st->diag_ns[5] += 1ull;
        st->diag_t0 = ~0ull;
      }
    }
    return;
  }
#pragma unroll
  for (int s = 0; s < SEG_NS; ++s) v[s] = s < SEGL_NS ? col[s * HPS_BLOCK] : 0.0;
  if (grid_reduce<SEG_NS>(v, 1u << 2, partials, ticket, s_total, s_scratch, xr)) {
    if (threadIdx.x == 0) finalize_segment(fin_mode, s_total, st, b.n_global);
  }
}
#define SEG_ARGS BeamView, const Op*, const OpC*, int, int, int, int, DevStats*, ScDev*, const double2*, KickParams, double*, u32*, double*, DeferList*, int, const __grid_constant__ XrView
typedef void (*SegKernel)(SEG_ARGS);
// variant -> kernel.  20 (default): 2 particles/thread, reference-order maps with exact branch-free arithmetic;
// 0: 2 particles/thread, fast maps; 1 / 2: fast maps, 1 particle/thread at 3 / 4 blocks per SM;
// 10, 11: variants 0, 1 with the reference's expression order and nvcc's operators (MATH = 0); 20, 21: the same
// with the branch-free exact arithmetic (MATH = 2); -1 / -2 / -3: the full op set (dipole, buncher)
static SegKernel seg_kernel(int variant) {
  switch (variant) {
    case 1: return k_segment<1, 3, false, 1>;
    case 2: return k_segment<1, 4, false, 1>;
    case 10: return k_segment<2, 2, false, 0>;
    case 11: return k_segment<1, 3, false, 0>;
    case 20: return k_segment<2, 2, false, 2>;

    case 21: return k_segment<1, 3, false, 2>;
    case -1: return k_segment<1, 3, true, 1>;
    case -2: return k_segment<1, 3, true, 0>;
    case -3: return k_segment<1, 3, true, 2>;
    case 0: return k_segment<2, 2, false, 1>;
    default: return k_segment<2, 2, false, 2>;
  }
}
// the pattern sweeps of the default variant (2 particles / thread, reference-order maps, exact arithmetic)
// (the drift sequences need few registers: 3 resident blocks per SM instead of 2 hide more of their memory latency)
struct SeqKernel { int seq; int minb; SegKernel fn; int ppt; };
static const SeqKernel g_seq_kernels[] = {
#define SK(S, B) {S, B, k_segment<2, B, false, 2, S>, 2}
#define SK1(S, B) {S, B, k_segment<1, B, false, 2, S>, 1}
    SK(SEQ_CCL_CELL, 2), SK(SEQ_DRIFT_STEP, 2), SK(SEQ_DRIFT_QUAD, 2), SK(SEQ_QUAD_DRIFT, 2), SK(SEQ_QUAD_GAP1D, 2), SK(SEQ_GAP2D_QUAD, 2),
    SK(SEQ_APER_DRIFT, 2), SK(SEQ_KICK_DRIFT, 2), SK(SEQ_KICK_GAP2C, 2), SK(SEQ_APER_GAP1C, 2),
    SK(SEQ_DRIFT_STEP, 3), SK(SEQ_APER_DRIFT, 3), SK(SEQ_KICK_DRIFT, 3),  // HPSIM_PATTERNS=2
#ifdef HPS_EXP_PPT1  // build experiment: one particle per thread at 4 blocks per SM for the RF-gap patterns (HPSIM_PATTERNS=4)
    SK1(SEQ_CCL_CELL, HPS_EXP_PPT1), SK1(SEQ_QUAD_GAP1D, HPS_EXP_PPT1), SK1(SEQ_GAP2D_QUAD, HPS_EXP_PPT1), SK1(SEQ_KICK_GAP2C, HPS_EXP_PPT1), SK1(SEQ_APER_GAP1C, HPS_EXP_PPT1),
#endif
#undef SK1
#undef SK
};
static const SeqKernel* seq_kernel(const Op* host_ops, int nops, int want_minb) {
  if (nops < 2 || nops > 3 || !host_ops) return nullptr;
  int code = 0;
  for (int k = 0; k < nops; ++k) {
    const int t = host_ops[k].type;
    const bool gap = t == OP_GAP_FIRST || t == OP_GAP_SECOND;
    code |= SQ(t, gap ? (host_ops[k].flag != 0 ? 1 : 0) : -1) << (8 * k);
  }
  const SeqKernel* hit = nullptr;
  for (const SeqKernel& e : g_seq_kernels)
    if (e.seq == code && e.minb <= want_minb && (!hit || e.minb > hit->minb)) hit = &e;  // the densest variant allowed
  return hit;
}
static int seg_full_variant(int variant) { return variant >= 20 ? -3 : (variant >= 10 ? -2 : -1); }
static bool seg_pairs(int variant) { return !(variant == 1 || variant == 2 || variant == 11 || variant == 21 || variant < 0); }

// one thread per op record of the uploaded program (once per simulate() call)
__global__ void k_prepare_ops(const Op* __restrict__ ops, OpC* opcs, int nops, double mass, double charge) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nops) prepare_op(ops[i], opcs[i], mass, charge);
}

// ================================================================================================
// K2: longitudinal-loss pass (only does work when K1 found a particle beyond <phi>_good + 3 pi)
// ================================================================================================
// Beam::UpdateLongitudinalLoss steps 3-4 + the recount of UpdateStatForSpaceCharge
// (beam.cpp:776-782, 1007-1009; beam_kernel.cu:1549-1559).  slots: 0 n_good 1 sum phi good 2 n_lloss
__device__ void finalize_lloss(const double* t, DevStats* st, ScDev* sc) {
  const u32 ng = (u32)t[0];
  st->num_good = ng;
  st->v[HPS_STAT_PHI_AVG_GOOD] = ng > 0 ? t[1] / ng : -99999;
  st->num_lloss = (u32)t[2];
  st->need_lloss_pass = 0;
  if (sc) sc->n_lloss_pass++;
}

// the sweep of the longitudinal-loss pass; every thread of the grid calls it
__device__ __forceinline__ void lloss_sweep(const BeamView& b, DevStats* st, ScDev* sc, double* partials, u32* ticket,
                                            double* s_total, double* s_scratch, const XrView& xr) {
  const double thr = st->v[HPS_STAT_PHI_AVG_GOOD] + K_TWOPI + K_PI;
  double acc[LL_NS] = {0.0, 0.0, 0.0, 0.0};
  const u32 npairs = (b.n + 1) >> 1;
  for (u32 i2 = blockIdx.x * blockDim.x + threadIdx.x; i2 < npairs; i2 += gridDim.x * blockDim.x) {
    const double2 PH = ld_f64x2(b.phi, i2);
    const uint2 L = ld_u32x2(b.loss, i2);
    uint2 LL = ld_u32x2(b.lloss, i2);
    const bool v1 = 2 * i2 + 1 < b.n;
    bool changed = false;
    if (PH.x > thr && LL.x != 1u) { LL.x = 1u; changed = true; }
    if (v1 && PH.y > thr && LL.y != 1u) { LL.y = 1u; changed = true; }
    if (changed) reinterpret_cast<uint2*>(b.lloss)[i2] = LL;
    if (L.x == 0 && LL.x == 0) { acc[0] += 1.0; acc[1] += PH.x; }
    if (LL.x != 0) acc[2] += 1.0;
    if (v1) {
      if (L.y == 0 && LL.y == 0) { acc[0] += 1.0; acc[1] += PH.y; }
      if (LL.y != 0) acc[2] += 1.0;
    }
  }
  if (grid_reduce<LL_NS>(acc, 0u, partials, ticket, s_total, s_scratch, xr)) {
    if (threadIdx.x == 0) finalize_lloss(s_total, st, sc);
  }
}

// stand-alone form: the Beam::Update* API path and XR_NCCL (whose finaliser is a separate kernel)
__global__ void __launch_bounds__(HPS_BLOCK)
k_lloss_pass(BeamView b, DevStats* st, ScDev* sc, double* partials, u32* ticket, const __grid_constant__ XrView xr) {
  __shared__ double s_total[LL_NS];
  __shared__ double s_scratch[HPS_BLOCK];
  if (st->need_lloss_pass == 0) return;
  lloss_sweep(b, st, sc, partials, ticket, s_total, s_scratch, xr);
}

// Barrier over the whole grid (all blocks are co-resident: the launch wrappers cap the grid at the
// kernel's occupancy x SM count).  bar[0] = arrivals, bar[1] = generation.
__device__ __forceinline__ void grid_barrier(u32* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const u32 gen = *(volatile u32*)&bar[1];
    __threadfence();
    if (atomicAdd(&bar[0], 1u) == gridDim.x - 1) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(&bar[1], 1u);
    } else {
      while (*(volatile u32*)&bar[1] == gen) {}
    }
    __threadfence();
  }
  __syncthreads();
}

// ================================================================================================
// K3: moment pass + the host policy of Scheff::Start evaluated on the device
// ================================================================================================
// slots: 0 sum phi_r (loss==0)  1 sum phi_r (good)  2 sum phi_r^2 (good)  3 sum x  4 sum x^2  5 sum y
//        6 sum y^2  7 sum w (all good)  8 MAX (x-x_avg)^2+(y-y_avg)^2 over loss==0
__device__ void scheff_policy(DevStats* st, ScDev* sc, const KickParams& kp, double r2max) {
  // space_charge.cpp:82-139.  This is HOST arithmetic in the reference (x86-64, no FMA), hence the
  // explicit round-to-nearest intrinsics: nvcc must not contract these into FMAs.  Divisions and square roots are
  // the IEEE-exact branch-free sequences of hps_math.cuh (this runs in ONE thread on the critical path of every kick).
  const u32 good_cnt = st->num_good;
  const double sigphi = st->v[HPS_STAT_PHI_SIG_GOOD];
  const double sigx = st->v[HPS_STAT_X_SIG_GOOD], sigy = st->v[HPS_STAT_Y_SIG_GOOD];
  const double ke = st->v[HPS_STAT_W_AVG_GOOD];
  const double gamma = __dadd_rn(div_nr(ke, kp.mass), 1.0);
  const double beta = sqrt_nr(__dsub_rn(1.0, div_nr(1.0, __dmul_rn(gamma, gamma))));
  const double lmbd = div_nr(299.792458, kp.freq);
  const double sigz = __dmul_rn(__dmul_rn(div_nr(__dmul_rn(sigphi, 0.5), 3.14159265), beta), lmbd);
  const double sigr = sqrt_nr(__dadd_rn(__dmul_rn(sigx, sigx), __dmul_rn(sigy, sigy)));
  const double cratior = __dmul_rn(sigr, sigz);
  const double cratioz = __dmul_rn(__dmul_rn(__dmul_rn(sigz, sigz), gamma), gamma);
  const double cratior_old = __dmul_rn(sc->sigr_old, sc->sigz_old);
  const double cratioz_old = __dmul_rn(__dmul_rn(__dmul_rn(sc->sigz_old, sc->sigz_old), sc->gm_old), sc->gm_old);
  u32 adj = (kp.adj_bunch_w == 0.0 || ke < kp.adj_bunch_w) ? kp.adj_bunch_cfg : 0u;
  sc->adj_bunch = adj;
  // space_charge.cpp:107-117: above the cutoff energy the mesh is halved and the kick interval quadrupled (the host
  // reads mesh_hi back before it plans the next drift).  SetMeshSize frees and re-creates ZEROED tables: until the
  // next re-mesh the kick then interpolates a zero field (mesh_resized -> the caller clears the field table).
  u32 nr = kp.nr, nz = kp.nz;
  sc->mesh_resized = 0u;
  if (kp.mesh_w != 0.0) {
    const u32 cur_nr = sc->nr ? sc->nr : kp.nr, cur_nz = sc->nr ? sc->nz : kp.nz;
    nr = cur_nr; nz = cur_nz;
    if (ke > kp.mesh_w) { nr = kp.nr / 2; nz = kp.nz / 2; sc->mesh_hi = 1u; }
    else if (ke <= kp.mesh_w) { nr = kp.nr; nz = kp.nz; sc->mesh_hi = 0u; }
    if (nr != cur_nr || nz != cur_nz) sc->mesh_resized = 1u;
    sc->nr = nr; sc->nz = nz;
  }
  sc->beta = beta;
  sc->n_kicks++;
  const bool remesh = cratior_old == 0.0 || cratioz_old == 0.0 ||
                      div_nr(fabs(__dsub_rn(cratior, cratior_old)), cratior_old) > kp.remesh_val ||
                      div_nr(fabs(__dsub_rn(cratioz, cratioz_old)), cratioz_old) > kp.remesh_val;
  sc->remesh = remesh ? 1 : 0;
  if (remesh) {
    sc->n_remesh++;
    sc->sigr_old = sigr; sc->sigz_old = sigz; sc->gm_old = gamma; sc->good_cnt_old = good_cnt;
    sc->ratio_r = 1.0; sc->ratio_z = 1.0; sc->ratio_q = 1.0; sc->ratio_g = 1.0;
    sc->num_on_mesh = 0u;
    sc->abs_phi_max_bits = 0ull;
    // UpdateMeshKernelCall (space_charge.cu:122-138) + UpdateScheffMeshSizeKernel (space_charge_kernel.cu:19-32)
    const double bl = beta * K_CLIGHT / kp.freq;
    if (adj != 0) {
      const double rmax = sqrt(r2max);  // Beam::UpdateMaxRPhi, with the stale x_avg / y_avg (H5)
      st->v[HPS_STAT_R_MAX] = rmax;
      sc->dr = 1.0 * rmax / nr;
      sc->dz = bl / nz;
    } else {
      // H6: the reference picks the sigma by comparing device POINTERS (x_sig_good > y_sig_good is false
      // for its allocation order) -> always y_sig_good.  H8: the 3*sigma > pi clamp is dead code.
      sc->dr = 3.0 * sigy / nr;
      sc->dz = 3.0 * bl * sigphi / (nz * K_PI);
    }
  } else {
    sc->ratio_r = sigr / sc->sigr_old;
    sc->ratio_z = sigz / sc->sigz_old;
    sc->ratio_q = sc->good_cnt_old ? (double)(good_cnt / sc->good_cnt_old) : 0.0;  // H7: uint division
    sc->ratio_g = gamma / sc->gm_old;
  }
}

// beam.cpp:1010-1015; finalisers beam_kernel.cu:1396-1446.  H4: surv uses the stale num_loss.
__device__ void finalize_moments(const double* t, double ref, u32 n_global, DevStats* st, ScDev* sc, const KickParams& kp,
                                 int do_policy) {
  const double surv = (double)(u32)(n_global - st->num_loss);
  const u32 ng = st->num_good;
  st->phi_r_ref = ref;
  st->v[HPS_STAT_PHI_AVG_R] = t[0] / surv;
  {
    const double avg = t[1] / ng, avg2 = t[2] / ng;
    st->v[HPS_STAT_PHI_SIG_GOOD] = sqrt(avg2 - avg * avg);
  }
  {
    const double xavg = t[3] / surv, x2avg = t[4] / surv, yavg = t[5] / surv, y2avg = t[6] / surv;
    st->v[HPS_STAT_X_AVG_GOOD] = xavg;
    st->v[HPS_STAT_X_SIG_GOOD] = sqrt(x2avg - xavg * xavg);
    st->v[HPS_STAT_Y_AVG_GOOD] = yavg;
    st->v[HPS_STAT_Y_SIG_GOOD] = sqrt(y2avg - yavg * yavg);
  }
  st->v[HPS_STAT_W_AVG_GOOD] = ng > 0 ? t[7] / ng : -99999;
  if (do_policy) scheff_policy(st, sc, kp, t[8]);
}

// mesh size in force (energy-dependent mesh, SetMeshSizeCutoffW): the configured size unless the cutoff is on
__device__ __forceinline__ KickParams with_mesh_in_force(const KickParams& kp, const ScDev* sc) {
  KickParams kq = kp;
  if (kp.mesh_w != 0.0) {
    const u32 nr = __ldcg(&sc->nr), nz = __ldcg(&sc->nz);
    if (nr != 0u) { kq.nr = nr; kq.nz = nz; }
  }
  return kq;
}

// loads of values another block may have written earlier in the SAME launch (L1 is not coherent)
__device__ __forceinline__ double ldv(const double* p) { return __ldcg(p); }
__device__ __forceinline__ u32 ldv(const u32* p) { return __ldcg(p); }
__device__ __forceinline__ int ldv(const int* p) { return __ldcg(p); }

// body of K3; every thread of the grid calls it.  s_total: MOM_NS doubles, s_scratch: blockDim.x doubles (shared).
__device__ __forceinline__ void moments_phase(const BeamView& b, DevStats* st, ScDev* sc, double* bins, const KickParams& kp,
                                              int do_policy, int with_lloss, double* partials, u32* ticket, double* s_total,
                                              double* s_scratch, const XrView& xr) {
  // with_lloss: the longitudinal-loss pass (K2) runs in front of the moments inside this launch when K1
  // asked for it -- the common "nothing to mark" case then costs nothing instead of an empty kernel.  The
  // flag only changes in K1's finaliser, so every block sees the same value for the whole launch.
  if (with_lloss && *(volatile u32*)&st->need_lloss_pass != 0u) {
    lloss_sweep(b, st, sc, partials, ticket, s_total, s_scratch, xr);
    grid_barrier(ticket + 2);
  }
  const double ref = *(volatile double*)&st->v[HPS_STAT_PHI_AVG_GOOD];
  const double xa = st->v[HPS_STAT_X_AVG], ya = st->v[HPS_STAT_Y_AVG];
  double acc[MOM_NS];
#pragma unroll
  for (int s = 0; s < MOM_NS; ++s) acc[s] = 0.0;
  const u32 npairs = (b.n + 1) >> 1;
  // Software-pipelined sweep: the loads of the NEXT pair are issued before the current pair is accumulated, so
  // one L2 round trip per iteration is all that is ever exposed.  (Left to itself ptxas issued the last of the six
  // loads of an iteration only after the first results had been consumed -- two dependent round trips, 15 us
  // per sweep instead of 4.)  A particle that fails a filter contributes an exact +0.0.
  struct MomPair { uint2 L, LL; double2 PH, X, Y, W; };
  auto load_pair = [&](u32 i2) {
    MomPair p;
    p.L = ld_u32x2(b.loss, i2, b.streaming); p.LL = ld_u32x2(b.lloss, i2, b.streaming);
    p.PH = ld_f64x2(b.phi, i2, b.streaming); p.X = ld_f64x2(b.x, i2, b.streaming); p.Y = ld_f64x2(b.y, i2, b.streaming);
    p.W = ld_f64x2(b.w, i2, b.streaming);
    return p;
  };
  const u32 stride = gridDim.x * blockDim.x;
  u32 i2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i2 < npairs) {
    MomPair cur = load_pair(i2);
    for (;;) {
      const u32 nx = i2 + stride;
      const bool more = nx < npairs;
      MomPair nxt = cur;
      if (more) nxt = load_pair(nx);
      const double x[2] = {cur.X.x, cur.X.y}, y[2] = {cur.Y.x, cur.Y.y}, ph[2] = {cur.PH.x, cur.PH.y}, w[2] = {cur.W.x, cur.W.y};
      const u32 l[2] = {cur.L.x, cur.L.y}, ll[2] = {cur.LL.x, cur.LL.y};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const bool alive = l[j] == 0 && (j == 0 || 2 * i2 + 1 < b.n);
        const bool good = alive && ll[j] == 0;
        const double pr = rel_phi(ph[j], ref);
        acc[0] += alive ? pr : 0.0;
        const double tx = x[j] - xa, ty = y[j] - ya;
        const double r2 = tx * tx + ty * ty;
        acc[8] = alive ? fmax(acc[8], r2) : acc[8];
        acc[1] += good ? pr : 0.0;
        acc[2] = good ? fma(pr, pr, acc[2]) : acc[2];  // fused like the reference build's sum2 += v * v
        acc[3] += good ? x[j] : 0.0;
        acc[4] = good ? fma(x[j], x[j], acc[4]) : acc[4];
        acc[5] += good ? y[j] : 0.0;
        acc[6] = good ? fma(y[j], y[j], acc[6]) : acc[6];
        acc[7] += good ? w[j] : 0.0;
      }
      if (!more) break;
      cur = nxt;
      i2 = nx;
    }
  }
  if (!grid_reduce<MOM_NS>(acc, 1u << 8, partials, ticket, s_total, s_scratch, xr)) return;
  if (threadIdx.x == 0) finalize_moments(s_total, ref, b.n_global, st, sc, kp, do_policy);
  __syncthreads();
  if (do_policy && sc->remesh) {  // ResetBinTbl (space_charge.cu:92-105)
    const u32 nb = kp.nr * kp.nz;
    for (u32 k = threadIdx.x; k < nb; k += blockDim.x) bins[k] = 0.0;
  }
}

// Fallback of the fused-moment epilogue of K1 (DevStats::moments_fallback): code 1 -> the whole chain again (good count
// and <phi>_good, longitudinal-loss pass, moments); code 2 -> losses are marked, moments only.  The flag is written by
// K1's finaliser only, so every block reads the same value for the whole launch.  Every thread of the grid calls it;
// s_scratch: blockDim.x doubles.
__device__ __forceinline__ void moments_fallback_phase(u32 code, const BeamView& b, DevStats* st, ScDev* sc, double* bins, const KickParams& kp,
                                                       double* partials, u32* ticket, double* s_total, double* s_scratch, const XrView& xr) {
  int with_lloss = 0;
  if (code == 1u) {
    // first link again: Beam::UpdateLongitudinalLoss steps 1-2 (what K1's FIN_KICK epilogue computes)
    double acc[LL_NS] = {0.0, 0.0, -DBL_MAX, 0.0};  // n_good, sum phi good, MAX phi over lloss == 0, n_lloss
    const u32 npairs = (b.n + 1) >> 1;
    for (u32 i2 = blockIdx.x * blockDim.x + threadIdx.x; i2 < npairs; i2 += gridDim.x * blockDim.x) {
      const double2 PH = ld_f64x2(b.phi, i2);
      const uint2 L = ld_u32x2(b.loss, i2);
      const uint2 LL = ld_u32x2(b.lloss, i2);
      const double ph[2] = {PH.x, PH.y};
      const u32 l[2] = {L.x, L.y}, ll[2] = {LL.x, LL.y};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j == 1 && !(2 * i2 + 1 < b.n)) continue;
        if (l[j] == 0 && ll[j] == 0) { acc[0] += 1.0; acc[1] += ph[j]; }
        if (ll[j] == 0) acc[2] = fmax(acc[2], ph[j]); else acc[3] += 1.0;
      }
    }
    if (grid_reduce<LL_NS>(acc, 1u << 2, partials, ticket, s_total, s_scratch, xr)) {
      if (threadIdx.x == 0) {
        const double t8[8] = {s_total[0], s_total[1], s_total[2], 0.0, 0.0, 0.0, s_total[3], 0.0};
        finalize_segment(FIN_KICK, t8, st, b.n_global);
        __threadfence();
      }
    }
    grid_barrier(ticket + 2);
    with_lloss = 1;
  }
  moments_phase(b, st, sc, bins, kp, 1, with_lloss, partials, ticket, s_total, s_scratch, xr);
}

// mode 0: the stand-alone link of the chain (Beam::Update* API path, XR_NCCL, HPSIM_FUSED_MOMENTS=0).
// mode 1: only when the fused-moment epilogue of the sweep before asked for it (DevStats::moments_fallback).
// (Folding the fallback into k_deposit instead -- one launch less per kick -- was measured: the extra code path costs the
//  deposit loop registers at its 64-register budget, 586 -> 751 ms per C4 pass, more than the empty launch.)
__global__ void __launch_bounds__(HPS_BLOCK)
k_moments(BeamView b, DevStats* st, ScDev* sc, double* bins, KickParams kp, int do_policy, int with_lloss, int mode, double* partials,
          u32* ticket, const __grid_constant__ XrView xr) {
  __shared__ double s_total[MOM_NS];
  __shared__ double s_scratch[HPS_BLOCK];
  pdl_launch_dependents();
  pdl_wait();
  if (mode == 1) {
    const u32 code = *(volatile u32*)&st->moments_fallback;
    if (code != 0u) moments_fallback_phase(code, b, st, sc, bins, kp, partials, ticket, s_total, s_scratch, xr);
    return;
  }
  moments_phase(b, st, sc, bins, kp, do_policy, with_lloss, partials, ticket, s_total, s_scratch, xr);
}

// ================================================================================================
// K4: r-z ring charge deposit, privatised in shared memory, fixed-point (order-independent) sums
// ================================================================================================
// space_charge_kernel.cu:123-224.  The reference does 4 global atomicAdds per particle into a 16-32 KB
// table (float, or a double CAS loop) -- the result depends on the order the atomics land in.  Here each
// block accumulates a private copy in shared memory and flushes its non-zero bins once, and the sums are
// 64-bit FIXED-POINT (weights scaled by a power of two, |error| <= 2^-43 per weight at 1 Mi particles):
// integer addition is associative, so the mesh is bit-identical from run to run, for any grid and for
// any split of the beam over GPUs.  sm_100 has no native 64-bit shared-memory atomic (both
// atomicAdd(double) and atomicAdd(u64) on shared memory compile to ATOMS.CAST.SPIN loops), so a bin is
// two u32 words updated with native ATOMS.ADD: low word first, the carry rides on the high-word add.
// The scale is 2^(62 - ceil(log2 N)): the largest power of two for which even N unit weights in one bin
// cannot overflow 63 bits (2^46 at 64 Ki particles, 2^42 at 1 Mi, 2^38 at 16 Mi).
__host__ __device__ inline double bin_scale(u32 n_global) {
#ifdef __CUDA_ARCH__
  // e = ceil(log2 N) without the loop (every thread of the field solve evaluates this: the loop was 17 % of its
  // instructions, profiles/r2a_ncu_k_field_16mi.txt)
  const int e = n_global <= 1u ? 0 : 32 - __clz((int)(n_global - 1u));
#else
  int e = 0;
  while (e < 32 && (1ull << e) < (unsigned long long)n_global) ++e;
#endif
  return (double)(1ull << (62 - e));
}

__device__ __forceinline__ void bin_add(u32* s_lo, u32* s_hi, int k, double w, double scale) {
  const unsigned long long q = (unsigned long long)__double2ll_rn(w * scale);  // two's complement
  const u32 lo = (u32)q;
  u32 hi = (u32)(q >> 32);
  if (lo) {
    const u32 old = atomicAdd(&s_lo[k], lo);
    hi += (u32)(old + lo < old);
  }
  if (hi) atomicAdd(&s_hi[k], hi);
}

struct DepCtx {
  double phimc, xbar, ybar, epsq, ref, dr, dzg, hl, c2, drdr, scale;
  double r_epsq, r_dr, r_dzg, r_drdr;  // correctly rounded reciprocals for udiv (hps_math.cuh)
  unsigned long long* past;            // counter of the particles dropped by the row-past-the-table rule below
  int nx, nz;
};

// one particle of space_charge_kernel.cu:123-224; returns 1 when it landed on the mesh
__device__ __forceinline__ u32 deposit_one(const DepCtx& d, u32 l, u32 ll, double x, double y, double phi, u32* s_lo,
                                           u32* s_hi, double& pmax) {
  if (l != 0) return 0;
  const double pr = rel_phi(phi, d.ref);
  {  // Beam::UpdateMaxRPhi's phase half (beam_kernel.cu:836-842); only reported, never fed back
    double tp = pr - d.phimc;
    tp = tp > 0.0 ? tp : -tp;
    pmax = pmax < tp ? tp : pmax;
  }
  if (ll != 0) return 0;
  const int nx = d.nx, nz = d.nz;
  const double dr = d.dr, dzg = d.dzg, drdr = d.drdr;
  // divisions and square roots: the IEEE-exact branch-free forms (same bits as `/` and sqrt, hps_math.cuh)
  const double rsq_sh = udiv((x - d.xbar) * (x - d.xbar), d.epsq, d.r_epsq) + (y - d.ybar) * (y - d.ybar) * d.epsq;
  const int i = (int)(udiv(sqrt_nr(rsq_sh), dr, d.r_dr) + 1);
  if (!(i <= nx)) return 0;
  double ddz = pr - d.phimc;
  while (ddz > K_PI || ddz < -K_PI) {
    if (ddz > 0.0) ddz -= K_TWOPI; else ddz += K_TWOPI;
  }
  const double z = -d.c2 * ddz;
  if (!(fabs(z) < d.hl)) return 0;
  const double zz = z + d.hl;
  const int jm1 = (int)(udiv(zz, dzg, d.r_dzg) + 1);
  int i1 = i + 1;
  if (rsq_sh < 0.5 * drdr * (double)(i * i + (i - 1) * (i - 1))) i1 = i - 1;
  if (i1 < 1) i1 = 1;
  if (i1 > nx) i1 = nx;
  int j1 = jm1 + 1;
  if (zz < (jm1 - 1 + 0.5) * dzg) j1 = jm1 - 1;
  if (j1 < 1) j1 = 1;
  if (j1 > nz) j1 = nz;
  double a;
  if (i1 == i) a = 1.0;
  else {
    const double sqrtmp = sqrt_nr(udiv(4.0 * rsq_sh, drdr, d.r_drdr) - 1.0);
    const double rminsq = 0.25 * drdr * (sqrtmp - 1.0) * (sqrtmp - 1.0);
    const double rmaxsq = rminsq + sqrtmp * drdr;
    if (i1 < i) a = div_nr(rmaxsq - (i - 1) * (i - 1) * drdr, rmaxsq - rminsq);
    else a = div_nr((i1 - 1) * (i1 - 1) * drdr - rminsq, rmaxsq - rminsq);
  }
  double cc = 1.0;
  if (j1 != jm1) {  // the divisor is +dzg or -dzg: x / (-dzg) == -(x / dzg) exactly
    const double q = udiv(zz - (j1 - 1 + 0.5) * dzg, dzg, d.r_dzg);
    cc = jm1 > j1 ? q : -q;
  }
  const double bb = 1.0 - a, dd = 1.0 - cc;
  // jm1 can reach nz+1 when zz/dzg rounds up to nz exactly; the reference then writes one row past
  // its table.  Such a particle is dropped here instead (it carries weight cc = 1 into nowhere).
  int k = (jm1 - 1) * nx + i;
  if (jm1 >= 1 && jm1 <= nz) {
    bin_add(s_lo, s_hi, k - 1, a * cc, d.scale);
    if (i1 != i) bin_add(s_lo, s_hi, k + (i1 - i) - 1, bb * cc, d.scale);
  } else atomicAdd(d.past, 1ull);  // counted (hps_scheff_phase_cycles out[7]); tests assert that it stays negligible
  if (j1 != jm1) {  // dd == 0 otherwise: the reference adds exact zeros
    k = (j1 - 1) * nx + i;
    bin_add(s_lo, s_hi, k - 1, a * dd, d.scale);
    if (i1 != i) bin_add(s_lo, s_hi, k + (i1 - i) - 1, bb * dd, d.scale);
  }
  return 1;
}

// body of K4 up to the flush of the block-private mesh; s_words: 2 * nr * nz u32 of shared memory
__device__ __forceinline__ void deposit_phase(const BeamView& b, DevStats* st, ScDev* sc, double* bins_raw, const KickParams& kp,
                                              u32* s_words) {
  unsigned long long* bins = reinterpret_cast<unsigned long long*>(bins_raw);
  const u32 nb = kp.nr * kp.nz;
  u32* s_lo = s_words;
  u32* s_hi = s_words + nb;
  for (u32 k = threadIdx.x; k < 2 * nb; k += blockDim.x) s_words[k] = 0u;
  __shared__ DepCtx s_d;  // block-uniform factors, derived once per block with the reference's expression order
  if (threadIdx.x == 0) {
    DepCtx d;
    d.nx = (int)kp.nr; d.nz = (int)kp.nz;
    d.scale = bin_scale(kp.n_global);
    d.phimc = ldv(&st->v[HPS_STAT_PHI_AVG_R]); d.xbar = ldv(&st->v[HPS_STAT_X_AVG_GOOD]); d.ybar = ldv(&st->v[HPS_STAT_Y_AVG_GOOD]);
    d.epsq = ldv(&st->v[HPS_STAT_X_SIG_GOOD]) / ldv(&st->v[HPS_STAT_Y_SIG_GOOD]);
    d.ref = ldv(&st->v[HPS_STAT_PHI_AVG_GOOD]);
    const double beta = ldv(&sc->beta);
    const double gm = rsqrt(1.0 - beta * beta);
    d.dr = ldv(&sc->dr);
    d.dzg = ldv(&sc->dz) * gm;
    const double wave_len = K_CLIGHT / kp.freq;
    d.hl = d.nz * 0.5 * d.dzg;
    d.c2 = beta * gm * wave_len / K_TWOPI;
    d.drdr = d.dr * d.dr;
    d.r_epsq = 1.0 / d.epsq; d.r_dr = 1.0 / d.dr; d.r_dzg = 1.0 / d.dzg; d.r_drdr = 1.0 / d.drdr;
    d.past = &sc->phase_clk[7];
    s_d = d;
  }
  __syncthreads();
  u32 cnt = 0;
  double pmax = 0.0;
  // two adjacent particles per thread, every array read with one 128/64-bit access; software-pipelined like the
  // moment sweep: the next pair's loads are in flight while the current pair is deposited
  const u32 npairs = (b.n + 1) >> 1;
  struct DepPair { uint2 L, LL; double2 PH, X, Y; };
  auto load_pair = [&](u32 i2) {
    DepPair p;
    p.L = ld_u32x2(b.loss, i2, b.streaming); p.LL = ld_u32x2(b.lloss, i2, b.streaming);
    p.PH = ld_f64x2(b.phi, i2, b.streaming); p.X = ld_f64x2(b.x, i2, b.streaming); p.Y = ld_f64x2(b.y, i2, b.streaming);
    return p;
  };
  const u32 stride = gridDim.x * blockDim.x;
  u32 i2 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i2 < npairs) {
    DepPair cur = load_pair(i2);
    for (;;) {
      const u32 nx = i2 + stride;
      const bool more = nx < npairs;
      DepPair nxt = cur;
      if (more) nxt = load_pair(nx);
      cnt += deposit_one(s_d, cur.L.x, cur.LL.x, cur.X.x, cur.Y.x, cur.PH.x, s_lo, s_hi, pmax);
      if (2 * i2 + 1 < b.n) cnt += deposit_one(s_d, cur.L.y, cur.LL.y, cur.X.y, cur.Y.y, cur.PH.y, s_lo, s_hi, pmax);
      if (!more) break;
      cur = nxt;
      i2 = nx;
    }
  }
  __syncthreads();
  for (u32 k = threadIdx.x; k < nb; k += blockDim.x) {
    const unsigned long long v = ((unsigned long long)s_hi[k] << 32) + s_lo[k];
    if (v) atomicAdd(&bins[k], v);  // RED.ADD.64 at the L2: integer, order-independent
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    pmax = fmax(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (cnt) atomicAdd(&sc->num_on_mesh, cnt);
    atomicMax(&sc->abs_phi_max_bits, (unsigned long long)__double_as_longlong(pmax));
  }
}

__global__ void __launch_bounds__(512, 2)
k_deposit(BeamView b, DevStats* st, ScDev* sc, double* bins_raw, KickParams kp, u32* ticket, const __grid_constant__ XrView xr) {
  extern __shared__ u32 s_words[];  // [nb] low words, [nb] high words
  pdl_launch_dependents();
  pdl_wait();
  if (!sc->remesh) return;
  const KickParams kq = with_mesh_in_force(kp, sc);
  deposit_phase(b, st, sc, bins_raw, kq, s_words);
  if (xr.mode != XR_NCCL) return;
  // XR_NCCL only: the rank's last block appends the on-mesh count and max |phi| to the mesh, which the host
  // then all-gathers.  (XR_P2P: k_field itself pushes the mesh to the peers, spread over all its blocks.)
  unsigned long long* bins = reinterpret_cast<unsigned long long*>(bins_raw);
  const u32 nb = kq.nr * kq.nz;
  __shared__ u32 s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) {
    bins[nb] = (unsigned long long)__ldcg(&sc->num_on_mesh);
    bins[nb + 1] = __ldcg(&sc->abs_phi_max_bits);
    *ticket = 0u;
  }
}

// ================================================================================================
// K5: ring-to-ring Green's function + field at the mesh nodes, one block per receiving radius
// ================================================================================================
// space_charge_kernel.cu:39-100 (Green table) and :472-568 (field, the -DDOUBLE_PRECISION variant).
// The reference materialises the 2.1 MB Green table in HBM and re-reads it from every block; here
// block rp computes its own [nr][nz] slice straight into shared memory and consumes it in place.
__device__ __forceinline__ void ellip_ke(double g, double& ee, double& ek) {
  const double h = log(g);
  ee = 1.0 + g * (0.4630106 - 0.2452740 * h + g * (0.1077857 - 0.04125321 * h));
  ek = 1.38629436 - 0.5 * h + g * (0.1119697 - 0.1213486 * h + g * (0.07253230 - 0.02887472 * h));
}

// Geometry of the field solve inside one block: RS source rings, TR = blockDim / RS threads per ring.  For the
// convolution a ring's threads are L lanes (4 adjacent nodes each, L = power of two >= nz / 4) x S source segments.
struct FieldGeom {
  u32 RS, TR, L, S, seg, HQ;
};
__host__ __device__ inline FieldGeom field_geom(u32 nz, u32 RS, u32 block) {
  FieldGeom g;
  g.RS = RS;
  g.TR = block / RS;
  g.L = 1;
  while (g.L * 4 < nz) g.L <<= 1;
  g.S = g.TR / g.L;  // 0: mesh too long for this block shape
  if (g.S > nz) g.S = nz;
  g.seg = g.S ? (nz + g.S - 1) / g.S : 0;
  g.HQ = (2 * nz + 3) / 4 + 1;
  return g;
}
__host__ __device__ inline size_t field_geom_bytes(const FieldGeom& g, u32 nz, u32 block) {
  return (size_t)g.RS * nz * sizeof(double) + (size_t)g.RS * 4 * g.HQ * sizeof(double2) + (size_t)g.RS * g.S * nz * sizeof(double2) +
         (size_t)block * sizeof(double2);
}
struct FieldSmem {  // carve-up of the dynamic shared memory of the field solve
  double* s_b;     // [RS][nz] charge of the block's source rings
  double2* s_h;    // [RS][4][HQ] Green's function of each ring as a Toeplitz kernel h[zs - zp + nz], de-interleaved
                   //   by (index mod 4) so that lanes owning 4 adjacent nodes each read consecutive words
  double2* s_q;    // [RS][S][nz] partial fields per source segment at the nodes zp < nz
  double2* s_red;  // [blockDim.x]
  FieldGeom g;
  __device__ FieldSmem(double* mem, u32 RS, u32 nz) {
    g = field_geom(nz, RS, blockDim.x);
    s_b = mem;
    s_h = reinterpret_cast<double2*>(mem + RS * nz);
    s_q = s_h + RS * 4 * g.HQ;
    s_red = s_q + RS * g.S * nz;
  }
  // h[mp], mp = zs - zp + nz in [0, 2 nz): (Er, -Ez) of offset mp - nz for sources downstream of the node (zs >= zp),
  // (Er, +Ez) of offset nz - mp - 1 for sources upstream (space_charge_kernel.cu:516-540)
  __device__ __forceinline__ double2& h(u32 kk, u32 mp) const { return s_h[(kk * 4 + (mp & 3u)) * g.HQ + (mp >> 2)]; }
};

// Sharded beam, XR_P2P: every block stores its share of the rank-local mesh (+ on-mesh count and max |phi|)
// into all peers' boxes as LL words (sequence number = the host's launch counter of the field solve,
// identical on all ranks).
__device__ __forceinline__ void field_push_mesh(const ScDev* sc, const double* bins, const KickParams& kp, u32 mesh_seq, u32 blk,
                                                u32 nblk, const XrView& xr) {
  const u32 nb = kp.nr * kp.nz;
  const unsigned long long* lb = reinterpret_cast<const unsigned long long*>(bins);
  const int par = (int)(mesh_seq & 1u);
  for (u32 k = blk * blockDim.x + threadIdx.x; k < nb + 2; k += nblk * blockDim.x) {
    const unsigned long long v = k < nb ? __ldcg(&lb[k]) : (k == nb ? (unsigned long long)ldv(&sc->num_on_mesh) : __ldcg(&sc->abs_phi_max_bits));
    for (int p = 0; p < xr.nranks; ++p) {
      unsigned long long* w = &xr.peer[p]->bins[par][xr.rank][2 * k];
      ll_send(w, (u32)v, mesh_seq);
      ll_send(w + 1, (u32)(v >> 32), mesh_seq);
    }
  }
}

// Green's function entry (ring rs0+kk -> radius irp, axial offset (t + 1/2) cells) into s_h; does not depend on
// the charge mesh (space_charge_kernel.cu:39-100)
__device__ __forceinline__ void field_green(const ScDev* sc, const KickParams& kp, u32 irp, u32 rs0, u32 kk, u32 t, bool live,
                                            const FieldSmem& m) {
  if (!live) return;
  const u32 nz = kp.nz;
  const double beta = ldv(&sc->beta);
  const double gm = rsqrt(1.0 - beta * beta);
  const double dr = ldv(&sc->dr);
  const double dzg = ldv(&sc->dz) * gm;
  const double wave_len = K_CLIGHT / kp.freq;
  const double pl = beta * gm * wave_len;
  const double c1 = 5721.67 * (kp.current_frac / kp.mass) / (kp.freq * 1e6 * (double)kp.n_global);
  const double rp = irp * dr;
  const int adj = (int)ldv(&sc->adj_bunch);
  const u32 jz = t, k = rs0 + kk;
  const double rs = sqrt(0.5 * (k * k + (k + 1) * (k + 1))) * dr;
  const double zs = 0.5 * dzg, zp = (jz + 1) * dzg;
  const double d = zp - zs;
  double d2 = d * d;
  const double c = (rp - rs) * (rp - rs), bq = (rp + rs) * (rp + rs);
  const double rsrp = rs * rp;
  double a = 4.0 * rsrp / (bq + d2);
  double ee, ek;
  ellip_ke(1.0 - a, ee, ek);
  double er1 = 0.0;
  a = sqrt(bq + d2);
  const double rs2mrp2 = rs * rs - rp * rp;
  if (rp != 0.0) er1 = (ek - (rs2mrp2 + d2) * ee / (c + d2)) / (2.0 * rp * a);
  double ez1 = d * ee / (a * (c + d2));
  for (int i = 1; i <= adj; ++i) {
    double xi = i;
    for (int j = 0; j < 2; ++j) {
      const double d1 = d - xi * pl;
      d2 = d1 * d1;
      a = 4.0 * rsrp / (bq + d2);
      ellip_ke(1.0 - a, ee, ek);
      a = sqrt(bq + d2);
      if (rp != 0.0) er1 += (ek - (rs2mrp2 + d2) * ee / (c + d2)) / (2.0 * rp * a);
      ez1 += d1 * ee / (a * (c + d2));
      xi = -xi;
    }
  }
  m.h(kk, nz + t) = make_double2(c1 * er1, -(c1 * ez1));
  m.h(kk, nz - t - 1) = make_double2(c1 * er1, c1 * ez1);
}

// Gathers the charge of the block's source rings (folding the R rank-local meshes of a sharded beam), then the
// field of those rings at the nodes of radius irp; the last block of each radius folds the chunk partials.
// Call with the Green slice of (irp, ch) already in m.s_g (a __syncthreads follows inside).
__device__ __forceinline__ void field_solve(DevStats* st, ScDev* sc, const double* __restrict__ bins, double2* field, double2* part,
                                            u32* tickets, const KickParams& kp, u32 irp, u32 ch, u32 nch, u32 RS, u32 kk, u32 t,
                                            bool live, u32 mesh_seq, const FieldSmem& m, const XrView& xr) {
  const u32 nr = kp.nr, nz = kp.nz;
  const u32 rs0 = ch * RS;
  const u32 nrs = min(RS, nr - rs0);
  const double inv_scale = 1.0 / bin_scale(kp.n_global);          // fixed-point mesh, see k_deposit
  double* s_b = m.s_b;
  double2* s_red = m.s_red;
  if (xr.nranks <= 1) {
    if (irp == 0 && ch == 0 && threadIdx.x == 0) {
      sc->num_on_mesh_all = ldv(&sc->num_on_mesh);
      if (ldv(&sc->adj_bunch) != 0) st->v[HPS_STAT_ABS_PHI_MAX] = __longlong_as_double((long long)__ldcg(&sc->abs_phi_max_bits));
    }
    const long long* ibins = reinterpret_cast<const long long*>(bins);  // fixed point, see k_deposit
    if (live) s_b[kk * nz + t] = (double)__ldcg(&ibins[t * nr + rs0 + kk]) * inv_scale;
  } else {
    // sharded beam: fold the R rank-local meshes (integer sums: exact, order-independent)
    const u32 nb = nr * nz;
    const unsigned long long* slot[HPS_XR_MAXR];
    const bool ll = xr.mode == XR_P2P;  // LL words in this rank's own box, polled in place
    const u32 seq32 = mesh_seq;
    if (ll) {
      const XrBox* own = xr.peer[xr.rank];
      for (int r = 0; r < xr.nranks; ++r) slot[r] = own->bins[(int)(mesh_seq & 1u)][r];
    } else {
      for (int r = 0; r < xr.nranks; ++r) slot[r] = reinterpret_cast<const unsigned long long*>(xr.bins_gather) + (size_t)r * (nb + 2);
    }
    auto value = [&](int r, u32 k) -> unsigned long long {
      return ll ? ll_recv64(slot[r] + 2 * (size_t)k, seq32, xr.err) : slot[r][k];
    };
    const bool stamp = ll && irp == 0 && ch == 0 && threadIdx.x == 0;
    const unsigned long long t_wait = stamp ? global_ns() : 0ull;
    if (live) {
      unsigned long long a = 0ull;
      for (int r = 0; r < xr.nranks; ++r) a += value(r, t * nr + rs0 + kk);
      s_b[kk * nz + t] = (double)(long long)a * inv_scale;
    }
    if (stamp) { xr.stats[2] += global_ns() - t_wait; xr.stats[3] += 1ull; }
    if (irp == 0 && ch == 0 && threadIdx.x == 0) {
      u32 cnt = 0;
      double pm = 0.0;
      for (int r = 0; r < xr.nranks; ++r) {
        cnt += (u32)value(r, nb);
        pm = fmax(pm, __longlong_as_double((long long)value(r, nb + 1)));
      }
      sc->num_on_mesh_all = cnt;
      if (ldv(&sc->adj_bunch) != 0) st->v[HPS_STAT_ABS_PHI_MAX] = pm;
    }
  }
  __syncthreads();

  const size_t nodes_per_rp = nz + 1;
  double2* my_part = part + ((size_t)irp * nch + ch) * nodes_per_rp;
  // Nodes zp < nz.  The field of ring kk at node zp is sum_zs v[zs] h[zs - zp + nz]: a Toeplitz product.  A thread
  // owns 4 adjacent nodes and one segment of the sources and slides a 4-wide window of h through registers: one
  // shared-memory load per source step serves 8 multiply-adds (the one-node-per-thread form re-read h for every
  // multiply-add pair and was bound by shared-memory bandwidth: 12.6 us of the 19 us solve).
  {
    const FieldGeom& g = m.g;
    const u32 ck = threadIdx.x / g.TR, cr = threadIdx.x % g.TR, cs = cr / g.L, cl = cr % g.L;
    if (ck < nrs && cs < g.S && 4 * cl < nz) {
      const double* bk = s_b + ck * nz;
      const u32 z0 = cs * g.seg, z1 = min(nz, z0 + g.seg);
      const int top = 2 * (int)nz - 1;
      auto H = [&](int mp) -> double2 { return m.h(ck, (u32)max(0, min(mp, top))); };
      int base = (int)z0 - 4 * (int)cl + (int)nz;  // index of h for the first of the thread's 4 nodes at source z0
      double2 w0 = H(base), w1 = H(base - 1), w2 = H(base - 2), w3 = H(base - 3);
      double ax0 = 0.0, ax1 = 0.0, ax2 = 0.0, ax3 = 0.0, az0 = 0.0, az1 = 0.0, az2 = 0.0, az3 = 0.0;
      u32 zs = z0;
#define FSTEP(A, B, C, D)                                              \
  {                                                                    \
    const double2 nw = H(base + 1);                                    \
    const double v = bk[zs];                                           \
    ax0 += v * A.x; az0 += v * A.y; ax1 += v * B.x; az1 += v * B.y;    \
    ax2 += v * C.x; az2 += v * C.y; ax3 += v * D.x; az3 += v * D.y;    \
    D = nw; ++base; ++zs;                                              \
  }
      while (zs < z1) {
        FSTEP(w0, w1, w2, w3) if (zs >= z1) break;
        FSTEP(w3, w0, w1, w2) if (zs >= z1) break;
        FSTEP(w2, w3, w0, w1) if (zs >= z1) break;
        FSTEP(w1, w2, w3, w0)
      }
#undef FSTEP
      double2* q = m.s_q + (size_t)(ck * g.S + cs) * nz + 4 * cl;
      q[0] = make_double2(ax0, az0);
      if (4 * cl + 1 < nz) q[1] = make_double2(ax1, az1);
      if (4 * cl + 2 < nz) q[2] = make_double2(ax2, az2);
      if (4 * cl + 3 < nz) q[3] = make_double2(ax3, az3);
    }
  }
  // node (zp = nz, rp): every source bin lies upstream (h index = zs); thread (kk, t) owns zs = t
  double2 last = make_double2(0.0, 0.0);
  if (live) {
    const double v = s_b[kk * nz + t];
    const double2 hv = m.h(kk, t);
    last = make_double2(v * hv.x, v * hv.y);
  }
  s_red[threadIdx.x] = last;
  __syncthreads();
  for (u32 s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      s_red[threadIdx.x].x += s_red[threadIdx.x + s].x;
      s_red[threadIdx.x].y += s_red[threadIdx.x + s].y;
    }
    __syncthreads();
  }
  if (threadIdx.x < nz) {  // fold the segment and ring partials of this chunk in a fixed order
    const u32 np = nrs * m.g.S;
    double2 a = m.s_q[threadIdx.x];
    for (u32 k = 1; k < np; ++k) { a.x += m.s_q[(size_t)k * nz + threadIdx.x].x; a.y += m.s_q[(size_t)k * nz + threadIdx.x].y; }
    my_part[threadIdx.x] = a;
  }
  if (threadIdx.x == 0) my_part[nz] = s_red[0];

  // last block of this rp folds the chunk partials (fixed order)
  __shared__ u32 s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const u32 tk = atomicAdd(&tickets[irp], 1u);
    s_last = (tk == nch - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (u32 node = threadIdx.x; node <= nz; node += blockDim.x) {
    double tx = 0.0, tz = 0.0;
    for (u32 c2 = 0; c2 < nch; ++c2) {
      const double2 v = __ldcg(&part[((size_t)irp * nch + c2) * nodes_per_rp + node]);
      tx += v.x; tz += v.y;
    }
    field[node * (nr + 1) + irp] = make_double2(tx, tz);
  }
  if (threadIdx.x == 0) tickets[irp] = 0u;
}

// Work split: block (rp, ch) owns receiving radius rp and RS consecutive source rings (RS = blockDim.x / T,
// T = power of two >= nz).  Thread (kk, t) computes ONE entry of the [RS][nz] Green slice (ring kk, axial
// offset t) into shared memory, then the partial field of ring kk at node (zp = t, rp) with two independent
// accumulation chains; the RS ring partials are folded in ring order, the chunk partial is stored, and the
// last block of each rp (ticket) adds the chunk partials in a fixed order -- deterministic, and
// (nr+1) x ceil(nr/RS) blocks of RS*T threads fill the GPU.
__global__ void __launch_bounds__(1024)
k_field(DevStats* st, ScDev* sc, const double* __restrict__ bins, double2* field, double2* part, u32* tickets,
        KickParams kp, int T, u32 mesh_seq, const __grid_constant__ XrView xr) {
  extern __shared__ double s_mem[];
  pdl_launch_dependents();
  pdl_wait();
  const KickParams kq = with_mesh_in_force(kp, sc);  // the grid is sized for the configured (largest) mesh
  if (!sc->remesh) {
    // SetMeshSize left the reference with a fresh, zeroed field table that this kick interpolates as it is
    if (kp.mesh_w != 0.0 && sc->mesh_resized) {
      const u32 nf = (kp.nr + 1) * (kp.nz + 1), nblk = gridDim.x * gridDim.y, blk = blockIdx.y * gridDim.x + blockIdx.x;
      for (u32 k = blk * blockDim.x + threadIdx.x; k < nf; k += nblk * blockDim.x) field[k] = make_double2(0.0, 0.0);
    }
    return;
  }
  const u32 RS = blockDim.x / T;
  const u32 irp = blockIdx.x, ch = blockIdx.y, nch = (kq.nr + RS - 1) / RS;
  if (irp > kq.nr || ch >= nch) return;
  const u32 rs0 = ch * RS;
  const u32 nrs = min(RS, kq.nr - rs0);
  const u32 kk = threadIdx.x / T, t = threadIdx.x % T;
  const bool live = kk < nrs && t < kq.nz;
  const FieldSmem m(s_mem, RS, kq.nz);
  // XR_P2P: the mesh is pushed first; the Green's function slice does not depend on the charge mesh and is
  // computed next, so the peers' meshes travel over NVLink meanwhile; then the R meshes are polled and folded.
  if (xr.mode == XR_P2P) field_push_mesh(sc, bins, kq, mesh_seq, ch * (kq.nr + 1) + irp, (kq.nr + 1) * nch, xr);
  field_green(sc, kq, irp, rs0, kk, t, live, m);
  field_solve(st, sc, bins, field, part, tickets, kq, irp, ch, nch, RS, kk, t, live, mesh_seq, m, xr);
}

// ================================================================================================
// K3 + K4 + K5 in ONE launch: the whole space-charge chain between two segment sweeps
// ================================================================================================
// Three short, latency-bound kernels (21 + 27 + 24 us per kick on C2, profiles/r1_launches_summary.txt) become
// one resident grid with two grid barriers:
//   moments (+ lloss pass) -> last block: finalisers + Scheff::Start policy -> BARRIER ->
//   [remesh kicks only] deposit into the block-private mesh, flush -> Green slice of the block's first
//   field item (independent of the mesh: it hides the flush and the barrier) -> BARRIER -> field solve.
// A sharded beam (XR_P2P) needs the complete rank-local mesh before it can be pushed to the peers, so there
// the barrier sits before the push and the Green slice hides the NVLink transfer instead.
// All blocks are co-resident (grid <= occupancy x SMs, hps_capi.cpp).  blockDim.x = BLK >= RS * T: 1024 threads
// (one block per SM) by default -- half as many block-private meshes to flush with L2 atomics, half as many
// partials and barrier arrivals as with 512-thread blocks.
template <int BLK>
__global__ void __launch_bounds__(BLK, 1024 / BLK)
k_scheff(BeamView b, DevStats* st, ScDev* sc, double* bins, double2* field, double2* part, u32* tickets, KickParams kp, int T,
         u32 mesh_seq, double* partials, u32* ticket, const __grid_constant__ XrView xr) {
  extern __shared__ __align__(16) unsigned char s_dyn[];  // deposit: 2*nr*nz u32; field: FieldSmem (whichever is larger)
  __shared__ double s_total[MOM_NS];
  __shared__ double s_scratch[BLK];
  pdl_launch_dependents();
  pdl_wait();
  const bool clk = blockIdx.x == 0 && threadIdx.x == 0;
  long long t0 = clock64(), t1;
#define PHASE(k) if (clk) { t1 = clock64(); sc->phase_clk[k] += (unsigned long long)(t1 - t0); t0 = t1; }
  moments_phase(b, st, sc, bins, kp, 1, 1, partials, ticket, s_total, s_scratch, xr);
  PHASE(0)
  grid_barrier(ticket + 2);
  PHASE(1)
  if (ldv(&sc->remesh) == 0) return;

  deposit_phase(b, st, sc, bins, kp, reinterpret_cast<u32*>(s_dyn));
  __syncthreads();  // the block-private mesh has been flushed: the shared memory is free again
  PHASE(2)

  const u32 RS = min((u32)(HPS_FIELD_RS * BLK / 512), blockDim.x / (u32)T);
  const u32 nch = (kp.nr + RS - 1) / RS;
  const u32 nitems = (kp.nr + 1) * nch;
  const u32 kk = threadIdx.x / T, t = threadIdx.x % T;
  const FieldSmem m(reinterpret_cast<double*>(s_dyn), RS, kp.nz);
  u32 item = blockIdx.x;
  auto live_of = [&](u32 it) { const u32 rs0 = (it / (kp.nr + 1)) * RS; return kk < RS && kk < min(RS, kp.nr - rs0) && t < kp.nz; };
  if (xr.mode == XR_P2P) {
    grid_barrier(ticket + 2);
    field_push_mesh(sc, bins, kp, mesh_seq, blockIdx.x, gridDim.x, xr);
    PHASE(4)
    if (item < nitems) field_green(sc, kp, item % (kp.nr + 1), (item / (kp.nr + 1)) * RS, kk, t, live_of(item), m);
    PHASE(3)
  } else {
    if (item < nitems) field_green(sc, kp, item % (kp.nr + 1), (item / (kp.nr + 1)) * RS, kk, t, live_of(item), m);
    PHASE(3)
    grid_barrier(ticket + 2);
    PHASE(4)
  }
  for (bool first = true; item < nitems; item += gridDim.x, first = false) {
    const u32 irp = item % (kp.nr + 1), ch = item / (kp.nr + 1);
    const bool live = live_of(item);
    if (!first) {
      __syncthreads();  // the previous item's shared memory has been consumed
      field_green(sc, kp, irp, ch * RS, kk, t, live, m);
    }
    field_solve(st, sc, bins, field, part, tickets, kp, irp, ch, nch, RS, kk, t, live, mesh_seq, m, xr);
  }
  PHASE(5)
#undef PHASE
}

// ================================================================================================
// generic statistics kernel behind hps_beam_update() (the reference's Beam::Update* family)
// ================================================================================================
__device__ __forceinline__ double pick(const BeamView& b, int var, u32 i, double phi_ref) {
  switch (var) {
    case 0: return b.x[i];
    case 1: return b.xp[i];
    case 2: return b.y[i];
    case 3: return b.yp[i];
    case 4: return b.phi[i];
    case 5: return b.w[i];
    default: return rel_phi(b.phi[i], phi_ref);
  }
}

__device__ void finalize_stat(const double* t, const BeamView& b, DevStats* st, const StatReq& rq) {
  const double surv = (double)(u32)(b.n_global - st->num_loss);
  switch (rq.kind) {
    case KS_COUNT_LOSS: st->num_loss = (u32)t[0]; break;
    case KS_COUNT_LLOSS: case KS_MARK_LLOSS: st->num_lloss = (u32)t[0]; break;
    case KS_COUNT_GOOD: st->num_good = (u32)t[0]; break;
    case KS_AVG:  // beam_kernel.cu:1412-1427
      if (!rq.good_only) st->v[rq.out[0]] = t[1] / surv;
      else st->v[rq.out[0]] = st->num_good > 0 ? t[1] / st->num_good : -99999;
      break;
    case KS_SIG: {  // beam_kernel.cu:1429-1446
      const double d = rq.good_only ? (double)st->num_good : surv;
      const double avg = t[1] / d, avg2 = t[2] / d;
      st->v[rq.out[0]] = sqrt(avg2 - avg * avg);
      break;
    }
    case KS_AVG_XY:  // beam_kernel.cu:1385-1394
      st->v[HPS_STAT_X_AVG] = t[1] / surv;
      st->v[HPS_STAT_Y_AVG] = t[3] / surv;
      break;
    case KS_EMIT_T: {  // beam_kernel.cu:1333-1354; out: avg sig pavg psig emit
      const double x = t[1] / surv, x2 = t[2] / surv, xp = t[3] / surv, xp2 = t[4] / surv, xxp = t[5] / surv;
      const double gamma1 = rq.w_ref / b.mass;
      const double betagamma = sqrt(gamma1 * (gamma1 + 2.0));
      st->v[rq.out[0]] = x;
      st->v[rq.out[1]] = sqrt(x2 - x * x);
      st->v[rq.out[2]] = xp;
      st->v[rq.out[3]] = sqrt(xp2 - xp * xp);
      st->v[rq.out[4]] = sqrt((x2 - x * x) * (xp2 - xp * xp) - (xxp - x * xp) * (xxp - x * xp)) * betagamma;
      break;
    }
    case KS_EMIT_Z: {  // beam_kernel.cu:1356-1375
      const double phi = t[1] / surv, phi2 = t[2] / surv, w = t[3] / surv, w2 = t[4] / surv, phi_w = t[5] / surv;
      const double phi_var = phi2 - phi * phi, w_var = w2 - w * w, cov = phi_w - phi * w;
      st->v[HPS_STAT_PHI_AVG_R] = phi;
      st->v[HPS_STAT_W_AVG] = w;
      st->v[HPS_STAT_PHI_SIG_R] = sqrt(phi_var);
      st->v[HPS_STAT_W_SIG] = sqrt(w_var);
      st->v[HPS_STAT_Z_EMIT] = sqrt(phi_var * w_var - cov * cov);
      break;
    }
    case KS_MAX_R_PHI:  // beam_kernel.cu:1377-1383
      st->v[HPS_STAT_R_MAX] = sqrt(t[6]);
      st->v[HPS_STAT_ABS_PHI_MAX] = t[7];
      break;
    default: break;
  }
}

// slots: 0 count(filter passes) 1 sum a 2 sum a^2 3 sum b 4 sum b^2 5 sum ab 6 MAX 7 MAX
__global__ void __launch_bounds__(HPS_BLOCK)
k_stat(BeamView b, DevStats* st, StatReq rq, double* partials, u32* ticket, const __grid_constant__ XrView xr) {
  __shared__ double s_total[STAT_NS];
  __shared__ double s_scratch[HPS_BLOCK];
  const double phi_ref = st->phi_r_ref;
  const double xa = st->v[HPS_STAT_X_AVG], ya = st->v[HPS_STAT_Y_AVG], pa = st->v[HPS_STAT_PHI_AVG_R];
  const double thr = st->v[HPS_STAT_PHI_AVG_GOOD] + K_TWOPI + K_PI;
  double acc[STAT_NS];
#pragma unroll
  for (int s = 0; s < STAT_NS; ++s) acc[s] = 0.0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += gridDim.x * blockDim.x) {
    const u32 l = b.loss[i];
    u32 ll = b.lloss[i];
    if (rq.kind == KS_MARK_LLOSS) {  // beam_kernel.cu:1549-1559
      if (b.phi[i] > thr && ll != 1u) { ll = 1u; b.lloss[i] = 1u; }
      if (ll != 0) acc[0] += 1.0;
      continue;
    }
    if (rq.kind == KS_COUNT_LOSS) { if (l != 0) acc[0] += 1.0; continue; }
    if (rq.kind == KS_COUNT_LLOSS) { if (ll != 0) acc[0] += 1.0; continue; }
    const bool pass = rq.good_only ? (l + ll == 0) : (l == 0);
    if (!pass) continue;
    acc[0] += 1.0;
    if (rq.kind == KS_COUNT_GOOD) continue;
    if (rq.kind == KS_MAX_R_PHI) {  // beam_kernel.cu:792-932
      const double tx = b.x[i] - xa, ty = b.y[i] - ya;
      double tp = rel_phi(b.phi[i], phi_ref) - pa;
      tp = tp > 0.0 ? tp : -tp;
      acc[6] = fmax(acc[6], tx * tx + ty * ty);
      acc[7] = fmax(acc[7], tp);
      continue;
    }
    const double a = pick(b, rq.var_a, i, phi_ref);
    acc[1] += a; acc[2] += a * a;
    if (rq.var_b >= 0) {
      const double c = pick(b, rq.var_b, i, phi_ref);
      acc[3] += c; acc[4] += c * c; acc[5] += a * c;
    }
  }
  if (!grid_reduce<STAT_NS>(acc, (1u << 6) | (1u << 7), partials, ticket, s_total, s_scratch, xr)) return;
  if (threadIdx.x == 0) finalize_stat(s_total, b, st, rq);
}

// ================================================================================================
// Update2DPlotData (simulation_engine.cu:34-82) in one sweep: Beam::UpdateLoss + Beam::UpdateEmittance (beam.cpp:981-990:
// horizontal, vertical, relative phase about the reference particle, longitudinal) + one row of the monitor table
// ================================================================================================
// slots: 0 n_loss | x: 1 sum 2 sum^2 3 sum x' 4 sum x'^2 5 sum x x' | y: 6..10 | 11 sum phi_r 12 phi_r^2 13 sum w 14 w^2 15 phi_r w
enum { MON_NS = 16 };
__global__ void __launch_bounds__(HPS_BLOCK)
k_monitor(BeamView b, DevStats* st, MonView mon, double* partials, u32* ticket, const __grid_constant__ XrView xr) {
  __shared__ double s_total[MON_NS];
  __shared__ double s_scratch[HPS_BLOCK];
  __shared__ double s_ref[2];
  // the reference particle (global particle 0): phase for UpdateRelativePhi (beam.cu:1122-1130), energy for the
  // emittance normalisation (beam_kernel.cu:1345) and <W> - W_ref.  Sharded beam: its owner tells the others.
  if (xr.nranks > 1) {
    if (blockIdx.x == 0) {
      if (threadIdx.x < 2) s_ref[threadIdx.x] = (b.offset == 0 && b.n > 0) ? (threadIdx.x == 0 ? b.phi[0] : b.w[0]) : 0.0;
      __syncthreads();
      if (xr.mode == XR_P2P) xr_exchange_p2p(s_ref, 2, 0u, xr);
      if (threadIdx.x == 0) { st->phi_r_ref = s_ref[0]; st->mon_ref_w = s_ref[1]; __threadfence(); }
    }
    grid_barrier(ticket + 2);
    if (threadIdx.x == 0) { s_ref[0] = ldv(&st->phi_r_ref); s_ref[1] = ldv(&st->mon_ref_w); }
  } else if (threadIdx.x == 0) {
    s_ref[0] = b.phi[0]; s_ref[1] = b.w[0];
  }
  __syncthreads();
  const double phi0 = s_ref[0], w0 = s_ref[1];
  double acc[MON_NS];
#pragma unroll
  for (int s = 0; s < MON_NS; ++s) acc[s] = 0.0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += gridDim.x * blockDim.x) {
    if (b.loss[i] != 0) { acc[0] += 1.0; continue; }
    const double x = b.x[i], xp = b.xp[i], y = b.y[i], yp = b.yp[i], w = b.w[i];
    const double pr = rel_phi(b.phi[i], phi0);
    acc[1] += x; acc[2] += x * x; acc[3] += xp; acc[4] += xp * xp; acc[5] += x * xp;
    acc[6] += y; acc[7] += y * y; acc[8] += yp; acc[9] += yp * yp; acc[10] += y * yp;
    acc[11] += pr; acc[12] += pr * pr; acc[13] += w; acc[14] += w * w; acc[15] += pr * w;
  }
  if (!grid_reduce<MON_NS>(acc, 0u, partials, ticket, s_total, s_scratch, xr)) return;
  if (threadIdx.x != 0) return;
  const double* t = s_total;
  st->num_loss = (u32)t[0];
  st->phi_r_ref = phi0;
  StatReq rq;
  rq.kind = KS_EMIT_T; rq.w_ref = w0; rq.good_only = 0; rq.var_a = 0; rq.var_b = 1;
  rq.out[0] = HPS_STAT_X_AVG; rq.out[1] = HPS_STAT_X_SIG; rq.out[2] = HPS_STAT_XP_AVG; rq.out[3] = HPS_STAT_XP_SIG; rq.out[4] = HPS_STAT_X_EMIT;
  finalize_stat(t, b, st, rq);
  rq.out[0] = HPS_STAT_Y_AVG; rq.out[1] = HPS_STAT_Y_SIG; rq.out[2] = HPS_STAT_YP_AVG; rq.out[3] = HPS_STAT_YP_SIG; rq.out[4] = HPS_STAT_Y_EMIT;
  finalize_stat(t + 5, b, st, rq);
  rq.kind = KS_EMIT_Z;
  finalize_stat(t + 10, b, st, rq);
  // the row (simulation_engine.cu:36-81)
  const u32 row = mon.counters[0], loss_num = mon.counters[1];
  const u32 nl = st->num_loss;
  const double loss_local = (double)(u32)(nl - loss_num);  // `GetLossNum() - loss_num`: unsigned arithmetic
  const u32 loss_new = (u32)((double)loss_num + loss_local);
  mon.counters[1] = loss_new;
  mon.counters[0] = row + 1u;
  if (row < mon.cap) {
    auto put = [&](int f, double v) { mon.data[(size_t)f * mon.cap + row] = v; };
    put(HPS_MON_LOSS_LOCAL, log10(loss_local / (double)b.n_global + 1e-6));
    put(HPS_MON_LOSS_RATIO, (double)loss_new / (double)b.n_global);
    put(HPS_MON_XAVG, st->v[HPS_STAT_X_AVG]); put(HPS_MON_XSIG, st->v[HPS_STAT_X_SIG]);
    put(HPS_MON_XPAVG, st->v[HPS_STAT_XP_AVG]); put(HPS_MON_XPSIG, st->v[HPS_STAT_XP_SIG]);
    put(HPS_MON_YAVG, st->v[HPS_STAT_Y_AVG]); put(HPS_MON_YSIG, st->v[HPS_STAT_Y_SIG]);
    put(HPS_MON_YPAVG, st->v[HPS_STAT_YP_AVG]); put(HPS_MON_YPSIG, st->v[HPS_STAT_YP_SIG]);
    put(HPS_MON_PHIAVG, st->v[HPS_STAT_PHI_AVG_R]); put(HPS_MON_PHISIG, st->v[HPS_STAT_PHI_SIG_R]);
    put(HPS_MON_WSIG, st->v[HPS_STAT_W_SIG]);
    put(HPS_MON_XEMIT, st->v[HPS_STAT_X_EMIT]); put(HPS_MON_YEMIT, st->v[HPS_STAT_Y_EMIT]); put(HPS_MON_ZEMIT, st->v[HPS_STAT_Z_EMIT]);
    put(HPS_MON_WAVG, st->v[HPS_STAT_W_AVG] - w0);
  }
}

// ================================================================================================
// XR_NCCL only: fold the all-gathered rank partials in rank order and run the finaliser that the
// XR_P2P / single-GPU path runs inside the reducing kernel's last block
// ================================================================================================
__global__ void __launch_bounds__(HPS_BLOCK)
k_xr_finalize(const __grid_constant__ XrView xr, FinReq f, BeamView b, DevStats* st, ScDev* sc, double* bins) {
  __shared__ double s_t[HPS_XR_NS];
  if (f.kind == FK_LLOSS && st->need_lloss_pass == 0) return;  // k_lloss_pass did not run either
  if ((int)threadIdx.x < f.ns) {
    const int s = threadIdx.x;
    const bool mx = (f.max_mask >> s) & 1u;
    double a = xr.recv[s];
    for (int r = 1; r < xr.nranks; ++r) {
      const double t = xr.recv[r * HPS_XR_NS + s];
      a = mx ? fmax(a, t) : a + t;
    }
    s_t[s] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    switch (f.kind) {
      case FK_SEGMENT: finalize_segment(f.fin_mode, s_t, st, b.n_global); break;
      case FK_LLOSS: finalize_lloss(s_t, st, sc); break;
      case FK_MOMENTS: finalize_moments(s_t, st->v[HPS_STAT_PHI_AVG_GOOD], b.n_global, st, sc, f.kp, f.do_policy); break;
      case FK_STAT: finalize_stat(s_t, b, st, f.rq); break;
      default: break;
    }
  }
  __syncthreads();
  if (f.kind == FK_MOMENTS && f.do_policy && sc->remesh) {
    const u32 nb = f.kp.nr * f.kp.nz;
    for (u32 k = threadIdx.x; k < nb; k += blockDim.x) bins[k] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
// Beam::UpdateRelativePhi (beam.cu:1122-1130): only the reference phase is recorded
__global__ void k_set_phi_ref(BeamView b, DevStats* st, int use_good, double phi0) {
  st->phi_r_ref = use_good ? st->v[HPS_STAT_PHI_AVG_GOOD] : phi0;
}
__global__ void k_copy_stat(DevStats* st, int dst, int src) { st->v[dst] = st->v[src]; }
// beam_kernel.cu:1489-1496 (global particle 0 is never shifted)
__global__ void k_shift(double* a, double v, u32 n, u32 offset) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && offset + i > 0) a[i] += v;
}
// beam_kernel.cu:1464-1480
__global__ void k_cut(const double* a, u32* loss, double lo, double hi, u32 n, u32 offset) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && offset + i > 0 && loss[i] == 0) {
    const double x = a[i];
    if (x < hi && x > lo) loss[i] = 99999u;
  }
}
__global__ void k_rel_phi_out(BeamView b, const DevStats* st, double* out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < b.n) out[i] = rel_phi(b.phi[i], st->phi_r_ref);
}

// ================================================================================================
// launch wrappers
// ================================================================================================

// Launch with programmatic stream serialization (see pdl_wait in hps_kernels.cuh) when c.pdl is set.
template <typename... KArgs, typename... Args>
static cudaError_t launch_chain(const LaunchCfg& c, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = c.stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = c.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

void launch_prepare_ops(const LaunchCfg& c, const Op* ops, OpC* opcs, int nops, double mass, double charge) {
  if (nops > 0) k_prepare_ops<<<(nops + 63) / 64, 64, 0, c.stream>>>(ops, opcs, nops, mass, charge);
}
size_t segment_smem(int nops, int stage) {
  return (size_t)SEG_ACC * HPS_BLOCK * sizeof(double) + (stage ? sizeof(PairStage) : 0) + (size_t)nops * (sizeof(Op) + sizeof(OpC));
}
void launch_segment(const LaunchCfg& c, const BeamView& b, const Op* host_ops, const Op* ops, const OpC* opcs, int nops,
                    int fin_mode, int dirty, int compat, DevStats* st, ScDev* sc, const double2* field,
                    const KickParams& kp, double* bins, DeferList* dl) {
  bool full = false;
  for (int k = 0; k < nops; ++k) full |= host_ops[k].type == OP_DIPOLE_FIRST || host_ops[k].type == OP_DIPOLE_SECOND || host_ops[k].type == OP_BUNCHER;
  const unsigned grid1 = c.grid_seg1, grid2 = c.grid_seg;
#define SEG_CALL b, ops, opcs, nops, fin_mode, dirty, compat, st, sc, field, kp, c.partials, c.ticket, bins, dl, c.stage, c.xr
  const size_t smem = segment_smem(nops, c.stage);
  // pattern sweeps (default variant only): the op sequence is compiled in
  const SeqKernel* pat = (c.seg_variant == 20 && c.patterns && !full) ? seq_kernel(host_ops, nops, c.patterns >= 4 ? 4 : (c.patterns >= 2 ? 3 : 2)) : nullptr;
  if (full) launch_chain(c, seg_kernel(seg_full_variant(c.seg_variant)), dim3(grid1), dim3(HPS_BLOCK), smem, SEG_CALL);
  else if (pat) {
    const unsigned pairs = pat->ppt == 2 ? (b.n + 1) / 2 : b.n;
    const unsigned grid = std::max(1u, std::min((pairs + HPS_BLOCK - 1) / HPS_BLOCK, (unsigned)(c.num_sms * pat->minb)));
    launch_chain(c, pat->fn, dim3(grid), dim3(HPS_BLOCK), smem, SEG_CALL);
  } else launch_chain(c, seg_kernel(c.seg_variant), dim3(grid2), dim3(HPS_BLOCK), smem, SEG_CALL);
#undef SEG_CALL
}
void launch_lloss_pass(const LaunchCfg& c, const BeamView& b, DevStats* st, ScDev* sc) {
  k_lloss_pass<<<c.grid_pairs, HPS_BLOCK, 0, c.stream>>>(b, st, sc, c.partials, c.ticket, c.xr);
}
void launch_moments(const LaunchCfg& c, const BeamView& b, DevStats* st, ScDev* sc, double* bins, const KickParams& kp,
                    int do_policy, int with_lloss, int mode) {
  launch_chain(c, k_moments, dim3(c.grid_pairs), dim3(HPS_BLOCK), 0, b, st, sc, bins, kp, do_policy, with_lloss, mode, c.partials, c.ticket, c.xr);
}
int moments_occupancy() {
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_moments, HPS_BLOCK, 0);
  return nb;
}
size_t deposit_smem(const KickParams& kp) { return (size_t)kp.nr * kp.nz * 2 * sizeof(u32); }
int field_threads(const KickParams& kp) {  // T: power of two >= nz (the tree reduction needs it)
  if (kp.nz == 0 || kp.nz > 1024) return 0;
  int t = 32;
  while (t < (int)kp.nz) t <<= 1;
  return t;
}
static unsigned field_rs(const KickParams& kp) {  // source rings per block: RS * T threads <= 1024
  const int t = field_threads(kp);
  return t ? std::max(1, std::min(HPS_FIELD_RS, 1024 / t)) : 1;
}
size_t field_smem(const KickParams& kp, int /*threads*/) {
  const unsigned rs = field_rs(kp), block = rs * (unsigned)field_threads(kp);
  return field_geom_bytes(field_geom(kp.nz, rs, block), kp.nz, block);
}
unsigned field_chunks(const KickParams& kp) { return (kp.nr + field_rs(kp) - 1) / field_rs(kp); }
// sets the dynamic-shared-memory opt-in of the chain kernels for a mesh (done at SCHEFF allocation, so that no
// function attribute has to change while a stream capture is in progress)
// (the attribute is per device: it is set for the CURRENT device every time a SCHEFF is allocated there, never cached
//  process-wide -- a second device would otherwise run meshes above 48 KB without the opt-in)
cudaError_t prepare_chain_attributes(const KickParams& kp) {
  if (deposit_smem(kp) > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_deposit, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
  }
  if (field_smem(kp, field_threads(kp)) > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_field, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
cudaError_t launch_deposit(const LaunchCfg& c, const BeamView& b, DevStats* st, ScDev* sc, double* bins,
                           const KickParams& kp) {
  const size_t smem = deposit_smem(kp);  // (> 48 KB: opted in by prepare_chain_attributes when the SCHEFF was allocated)
  launch_chain(c, k_deposit, dim3(c.grid_deposit), dim3(512), smem, b, st, sc, bins, kp, c.ticket + 1, c.xr);
  return cudaSuccess;
}
cudaError_t launch_field(const LaunchCfg& c, DevStats* st, ScDev* sc, const double* bins, double2* field, double2* part,
                         u32* tickets, const KickParams& kp, u32 mesh_seq) {
  const int T = field_threads(kp);
  const size_t smem = field_smem(kp, T);
  launch_chain(c, k_field, dim3(kp.nr + 1, field_chunks(kp)), dim3(T * field_rs(kp)), smem, st, sc, bins, field, part, tickets, kp, T, mesh_seq, c.xr);
  return cudaSuccess;
}
// ---- fused chain (k_scheff): BLK threads >= RS * T per block
static int g_scheff_blk = 0;
int scheff_block() {
  if (!g_scheff_blk) {
    const char* v = getenv("HPSIM_SCHEFF_BLOCK");
    g_scheff_blk = (v && atoi(v) == 512) ? 512 : 1024;
  }
  return g_scheff_blk;
}
static unsigned scheff_rs(const KickParams& kp) {
  const int t = field_threads(kp), blk = scheff_block();
  return (t && t <= blk) ? std::max(1, std::min(HPS_FIELD_RS * blk / 512, blk / t)) : 0;
}
size_t scheff_smem(const KickParams& kp) {
  const unsigned rs = scheff_rs(kp);
  if (!rs) return 0;
  const FieldGeom g = field_geom(kp.nz, rs, (u32)scheff_block());
  return std::max(field_geom_bytes(g, kp.nz, (u32)scheff_block()), deposit_smem(kp));
}
bool scheff_supported(const KickParams& kp) {
  const unsigned rs = scheff_rs(kp);
  return rs != 0 && field_geom(kp.nz, rs, (u32)scheff_block()).S != 0 && scheff_smem(kp) <= 200 * 1024;
}
unsigned scheff_items(const KickParams& kp) { const unsigned rs = scheff_rs(kp); return rs ? (kp.nr + 1) * ((kp.nr + rs - 1) / rs) : 0; }
int scheff_occupancy(const KickParams& kp) {
  const size_t smem = scheff_smem(kp);
  const bool big = scheff_block() == 1024;
  if (smem > 48 * 1024) {
    cudaError_t e = big ? cudaFuncSetAttribute(k_scheff<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                        : cudaFuncSetAttribute(k_scheff<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return 0;
  }
  int nb = 0;
  if (big) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_scheff<1024>, 1024, smem);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_scheff<512>, 512, smem);
  return nb;
}
cudaError_t launch_scheff(const LaunchCfg& c, unsigned grid, const BeamView& b, DevStats* st, ScDev* sc, double* bins, double2* field,
                          double2* part, u32* tickets, const KickParams& kp, u32 mesh_seq) {
  if (scheff_block() == 1024)
    return launch_chain(c, k_scheff<1024>, dim3(grid), dim3(1024), scheff_smem(kp), b, st, sc, bins, field, part, tickets, kp,
                        field_threads(kp), mesh_seq, c.partials, c.ticket, c.xr);
  return launch_chain(c, k_scheff<512>, dim3(grid), dim3(512), scheff_smem(kp), b, st, sc, bins, field, part, tickets, kp, field_threads(kp),
                      mesh_seq, c.partials, c.ticket, c.xr);
}
void launch_stat(const LaunchCfg& c, const BeamView& b, DevStats* st, const StatReq& rq) {
  k_stat<<<c.grid_single, HPS_BLOCK, 0, c.stream>>>(b, st, rq, c.partials, c.ticket, c.xr);
}
void launch_monitor(const LaunchCfg& c, const BeamView& b, DevStats* st, const MonView& mon) {
  k_monitor<<<c.grid_single, HPS_BLOCK, 0, c.stream>>>(b, st, mon, c.partials, c.ticket, c.xr);
}
void launch_xr_finalize(const LaunchCfg& c, const FinReq& f, const BeamView& b, DevStats* st, ScDev* sc, double* bins) {
  k_xr_finalize<<<1, HPS_BLOCK, 0, c.stream>>>(c.xr, f, b, st, sc, bins);
}
int reduce_slots(int kind) {
  switch (kind) {
    case FK_SEGMENT: return SEG_NS;
    case FK_LLOSS: return LL_NS;
    case FK_MOMENTS: return MOM_NS;
    default: return STAT_NS;
  }
}
unsigned reduce_max_mask(int kind) {
  switch (kind) {
    case FK_SEGMENT: return 1u << 2;
    case FK_MOMENTS: return 1u << 8;
    case FK_STAT: return (1u << 6) | (1u << 7);
    default: return 0u;
  }
}
void launch_set_phi_ref(const LaunchCfg& c, const BeamView& b, DevStats* st, int use_good, double phi0) {
  k_set_phi_ref<<<1, 1, 0, c.stream>>>(b, st, use_good, phi0);
}
void launch_copy_stat(const LaunchCfg& c, DevStats* st, int dst, int src) { k_copy_stat<<<1, 1, 0, c.stream>>>(st, dst, src); }
void launch_shift(const LaunchCfg& c, double* a, double v, u32 n, u32 offset) {
  k_shift<<<(n + 255) / 256, 256, 0, c.stream>>>(a, v, n, offset);
}
void launch_cut(const LaunchCfg& c, const double* a, u32* loss, double lo, double hi, u32 n, u32 offset) {
  k_cut<<<(n + 255) / 256, 256, 0, c.stream>>>(a, loss, lo, hi, n, offset);
}
void launch_rel_phi_out(const LaunchCfg& c, const BeamView& b, const DevStats* st, double* out) {
  k_rel_phi_out<<<(b.n + 255) / 256, 256, 0, c.stream>>>(b, st, out);
}
int segment_occupancy(int variant) {
  // static + dynamic shared memory of the sweep exceed the 48 KB default: opt in once per variant and device
  const size_t smem = segment_smem(HPS_MAX_SEG_OPS, 1);  // the largest launch
  cudaFuncSetAttribute(seg_kernel(variant), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (variant == 20)
    for (const SeqKernel& e : g_seq_kernels) cudaFuncSetAttribute(e.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, seg_kernel(variant), HPS_BLOCK, smem);
  return nb;
}
int segment_full_variant(int variant) { return seg_full_variant(variant); }
bool segment_pairs(int variant) { return seg_pairs(variant); }

}  // namespace hps
