// hpsim_oracle.cpp -- Oracle B: CPU restatement of HPSim's tracking + SCHEFF hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (hpsim_b200/) may call into
// this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs load the library built from it.
//
// It implements the same C ABI as the CUDA product (include/hpsim_b200.h) so a
// single test script can drive either implementation, but it is organised the
// way the reference is organised -- one array sweep per reference kernel, one
// host routine per reference launcher -- so that it is an independent check of
// the product's fused/planned formulation.  Every routine cites the reference
// file:line it restates (paths relative to /root/reference/src).
//
// Parity pinning: the reference ships no golden vectors or unit tests
// (SURVEY.md section 4), so this restatement is pinned against the reference's own
// CUDA code compiled unmodified for sm_100 (oracle/_ref/ref_driver_dp, built by
// oracle/Makefile) on the GPU box: tests/test_gpu_parity.py runs both on the
// same seeded beam.  Until that test has run this file is "parity unpinned".
//
// Arithmetic notes: glibc libm stands in for CUDA libdevice (sincos, sinh, cosh,
// exp, log, fmod, tan; rsqrt(x) -> 1/sqrt(x)); compiled with -ffp-contract=off.
// Reductions accumulate in long double in particle order (the reference uses a
// <=64-block tree; both are within ~1e-16 relative of the exact sum).
#include "../include/hpsim_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <strings.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef uint32_t uint;
// util/inc/constant.h:7-11
const double CLIGHT = 299.792458;
const double PI = 3.14159265358979323846;
const double TWOPI = 3.14159265358979323846 * 2.0;

// util/src/utility.cpp:7-16
uint NextPow2(uint v) {
  --v; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return ++v;
}
// CUDA's double->int conversion saturates and maps NaN to 0; x86's does not.
int cuda_d2i(double v) {
  if (v != v) return 0;
  if (v >= 2147483647.0) return 2147483647;
  if (v <= -2147483648.0) return -2147483647 - 1;
  return (int)v;
}
inline double rsqrt_(double v) { return 1.0 / std::sqrt(v); }

struct Stats {  // the 33 one-element device scalars of Beam (beam/inc/beam.h:200-262)
  double v[HPS_STAT_COUNT];
  uint num_loss, num_lloss, num_good;
};

struct Saved {
  bool valid = false;
  std::vector<double> c[6];
  std::vector<uint> loss, lloss;
  double design_w, freq, current;
};

}  // namespace

// Optional sharding of the oracle over several PROCESSES (tests only: tests/test_sharded_oracle.py drives
// two ranks over torch.distributed/gloo).  It restates, on the CPU, which quantities a particle-sharded
// beam has to exchange and where: every sum / count / max below goes through xr().
typedef void (*hps_oracle_allreduce_fn)(double* buf, int count, int op /* 0 sum, 1 max */, void* user);

struct hps_monitor {  // util/inc/plot_data.h:43-66 + monitor_num / loss_num (simulation_engine.cu:32-33)
  hps_ctx* ctx;
  uint cap = 0;
  std::vector<double> data;  // [HPS_MON_NFIELD][cap]
  uint monitor_num = 0, loss_num = 0;
};

struct hps_ctx {
  hps_monitor* monitor = nullptr;  // param->graphics_on / plot_data
  std::string err;
  int compat = 1;
  int nthreads = 0;
  int rank = 0, nranks = 1;
  hps_oracle_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  // engine statics (simulation/src/simulation_engine.cu:19-33)
  uint prev_end = 0xffffffffu;
  double design_w = 0.0;
  double wavelen = 0.0;  // d_wavelen, simulation_kernel.cu:13
  uint ccl_cell_num = 0;
};

struct hps_beam {
  hps_ctx* ctx;
  uint n;                     // local particles
  uint ng;                    // particles of the whole beam (== n when unsharded)
  uint off = 0, stride = 1;   // local particle i is global particle off + i * stride
  uint gi(uint i) const { return off + i * stride; }
  double mass, charge, current, freq, design_w;
  std::vector<double> x, xp, y, yp, phi, phi_r, w;
  std::vector<uint> loss, lloss;
  Stats s;
  Saved slot[2];
};

struct hps_lattice {
  hps_ctx* ctx;
  std::vector<hps_elem> e;
};

struct hps_scheff {
  hps_ctx* ctx;
  // SpaceCharge members (simulation/inc/space_charge.h:88-125)
  uint nx, nz, nx_init, nz_init;
  uint adj_bunch;
  double fraction = 1.0, interval = 0.01, interval_init = 0.01;
  double adj_bunch_w = 0.0, mesh_w = 0.0;
  double sigr_old = 0.0, sigz_old = 0.0, gm_old = 0.0;
  uint good_cnt_old = 0;
  double remesh_val = 0.05;
  // device state of space_charge_kernel.cu:12-15
  double dr = 0.0, dz = 0.0;
  uint num_on_mesh = 0;
  std::vector<double> bins;   // [nz][nr]
  std::vector<double> green;  // [nr+1][nr][nz][2]
  std::vector<double> field;  // [nz+1][nr+1][2]
  uint64_t kicks = 0, remesh = 0;
};

namespace {

static std::string g_err;
int fail(hps_ctx* c, int code, const std::string& m) {
  if (c) c->err = m; else g_err = m;
  return code;
}

// cross-rank fold of `count` doubles (no-op for a single process)
void xr(const hps_beam* b, double* v, int count, int op = 0) {
  if (b->ctx->nranks > 1) b->ctx->allreduce(v, count, op, b->ctx->allreduce_user);
}
double xr1(const hps_beam* b, double v, int op = 0) { xr(b, &v, 1, op); return v; }
// value held by the owner of global particle 0, seen by every rank
double from_ref_particle(const hps_beam* b, const std::vector<double>& a) {
  return xr1(b, (b->off == 0 && b->n > 0) ? a[0] : 0.0);
}

// ------------------------------------------------------------------ reductions
// beam/inc/beam_kernel.cu:84-166 (XReduceKernel) + finalisers :1412-1427
enum Filter { F_LOSS, F_GOOD };
inline bool keep(const hps_beam* b, uint i, Filter f) {
  return f == F_LOSS ? b->loss[i] == 0 : (b->loss[i] + b->lloss[i]) == 0;
}
long double sum1(const hps_beam* b, const std::vector<double>& a, Filter f) {
  long double s = 0;
  const long n = b->n;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (long i = 0; i < n; ++i)
    if (keep(b, (uint)i, f)) s += a[i];
  if (b->ctx->nranks > 1) return (long double)xr1(b, (double)s);
  return s;
}
void sum2(const hps_beam* b, const std::vector<double>& a, Filter f, long double* s1, long double* s2) {
  long double p = 0, q = 0;
  const long n = b->n;
#pragma omp parallel for reduction(+ : p, q) schedule(static)
  for (long i = 0; i < n; ++i)
    if (keep(b, (uint)i, f)) { p += a[i]; q += (long double)(a[i] * a[i]); }
  if (b->ctx->nranks > 1) { double t[2] = {(double)p, (double)q}; xr(b, t, 2); p = t[0]; q = t[1]; }
  *s1 = p; *s2 = q;
}

// beam.cu:493-516 + beam_kernel.cu:270-347
void UpdateLoss(hps_beam* b) {
  uint c = 0;
  for (uint i = 0; i < b->n; ++i) c += b->loss[i] != 0;
  b->s.num_loss = (uint)xr1(b, c);
}
void UpdateLlossCount(hps_beam* b) {
  uint c = 0;
  for (uint i = 0; i < b->n; ++i) c += b->lloss[i] != 0;
  b->s.num_lloss = (uint)xr1(b, c);
}
// beam.cu:1233-1255 + beam_kernel.cu:180-259
void UpdateGoodParticleCount(hps_beam* b) {
  uint c = 0;
  for (uint i = 0; i < b->n; ++i) c += (b->loss[i] + b->lloss[i]) == 0;
  b->s.num_good = (uint)xr1(b, c);
}
// beam.cu:411-436; finalisers beam_kernel.cu:1412-1427
void UpdateAvg(hps_beam* b, const std::vector<double>& a, int stat, bool good_only) {
  double s = (double)sum1(b, a, good_only ? F_GOOD : F_LOSS);
  if (!good_only) {
    double surv = (double)(uint)(b->ng - b->s.num_loss);
    b->s.v[stat] = s / surv;
  } else {
    b->s.v[stat] = b->s.num_good > 0 ? s / b->s.num_good : -99999;
  }
}
// beam.cu:651-678; finalisers beam_kernel.cu:1429-1446
void UpdateSig(hps_beam* b, const std::vector<double>& a, int stat, bool good_only) {
  if (!(b->ng > 2)) return;  // beam.cpp:866-955 guards
  long double s1, s2;
  sum2(b, a, good_only ? F_GOOD : F_LOSS, &s1, &s2);
  double d = good_only ? (double)b->s.num_good : (double)(uint)(b->ng - b->s.num_loss);
  double avg = (double)s1 / d, avg2 = (double)s2 / d;
  b->s.v[stat] = std::sqrt(avg2 - avg * avg);
}
// beam.cu:1122-1130 + beam_kernel.cu:1506-1522
void UpdateRelativePhi(hps_beam* b, bool use_good) {
  const double ref = use_good ? b->s.v[HPS_STAT_PHI_AVG_GOOD] : from_ref_particle(b, b->phi);
  const long n = b->n;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    double tmp = b->phi[i] - ref + PI, sign = 1.0;
    if (tmp < 0) { sign = -1.0; tmp = -tmp; }
    b->phi_r[i] = sign * (std::fmod(tmp, TWOPI) - PI);
  }
}
// beam.cpp:776-782 + beam_kernel.cu:1549-1559
void UpdateLongitudinalLoss(hps_beam* b) {
  UpdateGoodParticleCount(b);
  UpdateAvg(b, b->phi, HPS_STAT_PHI_AVG_GOOD, true);
  const double thr = b->s.v[HPS_STAT_PHI_AVG_GOOD] + TWOPI + PI;
  for (uint i = 0; i < b->n; ++i)
    if (b->phi[i] > thr) b->lloss[i] = 1;
  UpdateLlossCount(b);
}
// beam.cu:739-761 + beam_kernel.cu:612-723, 1385-1394
void UpdateAvgXY(hps_beam* b) {
  double sx = (double)sum1(b, b->x, F_LOSS), sy = (double)sum1(b, b->y, F_LOSS);
  double surv = (double)(uint)(b->ng - b->s.num_loss);
  b->s.v[HPS_STAT_X_AVG] = sx / surv;
  b->s.v[HPS_STAT_Y_AVG] = sy / surv;
}
// beam.cu:919-939 + beam_kernel.cu:981-1115, 1396-1410  (H4: divides by N - num_loss)
void UpdateAvgSigXY(hps_beam* b) {
  long double x1, x2, y1, y2;
  sum2(b, b->x, F_GOOD, &x1, &x2);
  sum2(b, b->y, F_GOOD, &y1, &y2);
  double surv = (double)(uint)(b->ng - b->s.num_loss);
  double xa = (double)x1 / surv, x2a = (double)x2 / surv, ya = (double)y1 / surv, y2a = (double)y2 / surv;
  b->s.v[HPS_STAT_X_AVG_GOOD] = xa;
  b->s.v[HPS_STAT_X_SIG_GOOD] = std::sqrt(x2a - xa * xa);
  b->s.v[HPS_STAT_Y_AVG_GOOD] = ya;
  b->s.v[HPS_STAT_Y_SIG_GOOD] = std::sqrt(y2a - ya * ya);
}
// beam.cu:829-852 + beam_kernel.cu:792-932, 1377-1383  (H5: stale x_avg / y_avg)
void UpdateMaxRPhi(hps_beam* b) {
  const double xa = b->s.v[HPS_STAT_X_AVG], ya = b->s.v[HPS_STAT_Y_AVG], pa = b->s.v[HPS_STAT_PHI_AVG_R];
  double r2m = 0, pm = 0;
  for (uint i = 0; i < b->n; ++i)
    if (b->loss[i] == 0) {
      double tx = b->x[i] - xa, ty = b->y[i] - ya;
      double r2 = tx * tx + ty * ty;
      double tp = b->phi_r[i] - pa;
      tp = tp > 0.0 ? tp : -tp;
      r2m = r2m < r2 ? r2 : r2m;
      pm = pm < tp ? tp : pm;
    }
  r2m = xr1(b, r2m, 1);
  pm = xr1(b, pm, 1);
  b->s.v[HPS_STAT_R_MAX] = std::sqrt(r2m);
  b->s.v[HPS_STAT_ABS_PHI_MAX] = pm;
}
// beam.cu:1007-1092 + beam_kernel.cu:1189-1375
void emit_sums(const hps_beam* b, const std::vector<double>& u, const std::vector<double>& up, long double o[5]) {
  long double a = 0, a2 = 0, c = 0, c2 = 0, ac = 0;
  for (uint i = 0; i < b->n; ++i)
    if (b->loss[i] == 0) {
      a += u[i]; a2 += (long double)(u[i] * u[i]);
      c += up[i]; c2 += (long double)(up[i] * up[i]);
      ac += (long double)(u[i] * up[i]);
    }
  if (b->ctx->nranks > 1) {
    double t[5] = {(double)a, (double)a2, (double)c, (double)c2, (double)ac};
    xr(b, t, 5);
    a = t[0]; a2 = t[1]; c = t[2]; c2 = t[3]; ac = t[4];
  }
  o[0] = a; o[1] = a2; o[2] = c; o[3] = c2; o[4] = ac;
}
void UpdateTransverseEmittance(hps_beam* b, const std::vector<double>& u, const std::vector<double>& up,
                               int avg, int sig, int pavg, int psig, int emit) {
  long double o[5];
  emit_sums(b, u, up, o);
  double surv = (double)(uint)(b->ng - b->s.num_loss);
  double x = (double)o[0] / surv, x2 = (double)o[1] / surv, xp = (double)o[2] / surv,
         xp2 = (double)o[3] / surv, xxp = (double)o[4] / surv;
  double gamma1 = from_ref_particle(b, b->w) / b->mass;
  double betagamma = std::sqrt(gamma1 * (gamma1 + 2.0));
  b->s.v[avg] = x;
  b->s.v[sig] = std::sqrt(x2 - x * x);
  b->s.v[pavg] = xp;
  b->s.v[psig] = std::sqrt(xp2 - xp * xp);
  b->s.v[emit] = std::sqrt((x2 - x * x) * (xp2 - xp * xp) - (xxp - x * xp) * (xxp - x * xp)) * betagamma;
}
void UpdateEmittance(hps_beam* b) {  // beam.cpp:981-990
  if (!(b->ng > 2)) return;
  UpdateTransverseEmittance(b, b->x, b->xp, HPS_STAT_X_AVG, HPS_STAT_X_SIG, HPS_STAT_XP_AVG, HPS_STAT_XP_SIG, HPS_STAT_X_EMIT);
  UpdateTransverseEmittance(b, b->y, b->yp, HPS_STAT_Y_AVG, HPS_STAT_Y_SIG, HPS_STAT_YP_AVG, HPS_STAT_YP_SIG, HPS_STAT_Y_EMIT);
  UpdateRelativePhi(b, false);
  long double o[5];
  emit_sums(b, b->phi_r, b->w, o);
  double surv = (double)(uint)(b->ng - b->s.num_loss);
  double phi = (double)o[0] / surv, phi2 = (double)o[1] / surv, w = (double)o[2] / surv,
         w2 = (double)o[3] / surv, phi_w = (double)o[4] / surv;
  double phi_var = phi2 - phi * phi, w_var = w2 - w * w, cov = phi_w - phi * w;
  b->s.v[HPS_STAT_PHI_AVG_R] = phi;
  b->s.v[HPS_STAT_W_AVG] = w;
  b->s.v[HPS_STAT_PHI_SIG_R] = std::sqrt(phi_var);
  b->s.v[HPS_STAT_W_SIG] = std::sqrt(w_var);
  b->s.v[HPS_STAT_Z_EMIT] = std::sqrt(phi_var * w_var - cov * cov);
}
// beam.cpp:1003-1016
void UpdateStatForSpaceCharge(hps_beam* b) {
  if (!(b->ng > 2)) return;
  UpdateLongitudinalLoss(b);
  UpdateGoodParticleCount(b);
  UpdateAvg(b, b->phi, HPS_STAT_PHI_AVG_GOOD, true);
  UpdateRelativePhi(b, true);
  UpdateAvg(b, b->phi_r, HPS_STAT_PHI_AVG_R, false);
  UpdateSig(b, b->phi_r, HPS_STAT_PHI_SIG_GOOD, true);
  UpdateAvgSigXY(b);
  UpdateAvg(b, b->w, HPS_STAT_W_AVG_GOOD, true);
}
// beam.cpp:655-661
void InitPhiAvgGood(hps_beam* b) {
  UpdateLoss(b);
  UpdateAvg(b, b->phi, HPS_STAT_PHI_AVG, false);
  b->s.v[HPS_STAT_PHI_AVG_GOOD] = b->s.v[HPS_STAT_PHI_AVG];
}
// beam.cpp:1658-1669
void ChangeFrequency(hps_beam* b, double freq) {
  b->current *= freq / b->freq;
  const double ratio = freq / b->freq;
  for (uint i = 0; i < b->n; ++i) b->phi[i] *= ratio;
  b->freq = freq;
  UpdateAvg(b, b->phi, HPS_STAT_PHI_AVG, false);
  UpdateRelativePhi(b, false);
  UpdateAvg(b, b->phi_r, HPS_STAT_PHI_AVG_R, false);
}

// ---------------------------------------------------------------- element maps
// simulation/inc/simulation_kernel.cu:16-73
double Bessi0c(double rx) {
  double ans, y, ax = std::fabs(rx);
  if (ax < 3.75) {
    y = rx / 3.75; y = y * y;
    ans = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
  } else {
    y = 3.75 / ax;
    ans = (std::exp(ax) * rsqrt_(ax)) * (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 +
          y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
  }
  return ans;
}
double Bessi1c(double rx) {
  double y, ans, ax = std::fabs(rx);
  if (ax < 3.75) {
    y = rx / 3.75; y = y * y;
    ans = ax * (0.5 + y * (0.87890594 + y * (0.51498869 + y * (0.15084934 + y * (0.2658733e-1 + y * (0.301532e-2 + y * 0.32411e-3))))));
  } else {
    y = 3.75 / ax;
    ans = 0.2282967e-1 + y * (-0.2895312e-1 + y * (0.1787654e-1 - y * 0.420059e-2));
    ans = 0.39894228 + y * (-0.3988024e-1 + y * (-0.362018e-2 + y * (0.163801e-2 + y * (-0.1031555e-1 + y * ans))));
    ans = ans * (std::exp(ax) * rsqrt_(ax));
  }
  return rx < 0.0 ? -ans : ans;
}
double Bessi1p(double rx) { return rx != 0.0 ? Bessi0c(rx) - Bessi1c(rx) / rx : 0.5; }

// simulation_kernel.cu:87-118
void PartialDrift(hps_beam* b, double wavelen, double hl, double aper, uint idx) {
  const long n = b->n;
  const double mass = b->mass;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    if (b->loss[i] != 0) continue;
    double aper2 = aper * aper;
    double x = b->x[i], y = b->y[i];
    x += b->xp[i] * hl;
    y += b->yp[i] * hl;
    if (aper2 != 0.0 && x * x + y * y > aper2) b->loss[i] = idx;
    else {
      b->x[i] = x; b->y[i] = y;
      double gm1 = b->w[i] / mass;
      double bt = std::sqrt(gm1 * (gm1 + 2.0)) / (gm1 + 1.0);
      b->phi[i] += TWOPI * hl / (wavelen * bt);
    }
  }
}
// simulation_kernel.cu:120-140
void SteererMap(hps_beam* b, double blh, double blv) {
  for (uint i = 0; i < b->n; ++i) {
    if (b->loss[i] != 0) continue;
    double gm1 = b->w[i] / b->mass;
    double btgm = std::sqrt(gm1 * (gm1 + 2.0));
    double p = btgm * b->mass / CLIGHT;
    if (blh > 1e-10) b->xp[i] += blh * b->charge / p;
    if (blv > 1e-10) b->yp[i] += blv * b->charge / p;
  }
}
// simulation_kernel.cu:142-218
void HalfQuad(hps_beam* b, double wavelen, double length, double aper, double gradient, uint idx) {
  double aper2 = aper; aper2 *= aper2;
  const double qg = gradient * b->charge, mass = b->mass;
  const long n = b->n;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    if (b->loss[i] != 0) continue;
    double x = b->x[i], y = b->y[i], xp = b->xp[i], yp = b->yp[i], w = b->w[i];
    if (qg == 0) {
      b->x[i] = x + xp * length * 0.5;
      b->y[i] = y + yp * length * 0.5;
    } else {
      double gamma1 = w / mass;
      double bg = std::sqrt(gamma1 * (gamma1 + 2.0) * (1.0 + xp * xp + yp * yp));
      double absqg = qg > 0.0 ? qg : -qg;
      double k = std::sqrt(absqg / (bg * mass) * CLIGHT);
      double inv_k = 1.0 / k;
      double kds = length * 0.25 * k;
      double s, c, sh = std::sinh(kds), ch = std::cosh(kds);
      sincos(kds, &s, &c);
      double xx, yy, sovk = s * inv_k, shovk = sh * inv_k, sk = s * k, shk = sh * k;
      for (int rep = 0; rep < 2; ++rep) {
        if (qg > 0.0) {
          xx = c * x + sovk * xp; xp = -sk * x + c * xp; x = xx;
          yy = ch * y + shovk * yp; yp = shk * y + ch * yp; y = yy;
        } else {
          xx = ch * x + shovk * xp; xp = shk * x + ch * xp; x = xx;
          yy = c * y + sovk * yp; yp = -sk * y + c * yp; y = yy;
        }
      }
      b->x[i] = x; b->y[i] = y; b->xp[i] = xp; b->yp[i] = yp;
    }
    // note: in the qg == 0 branch the test uses the PRE-drift x,y (reference quirk)
    if (aper2 != 0 && x * x + y * y > aper2) b->loss[i] = idx;
    else {
      double gm1 = w / mass;
      double bt = std::sqrt(gm1 * (gm1 + 2.0)) / (gm1 + 1.0);
      b->phi[i] += TWOPI * length * 0.5 / (wavelen * bt);
    }
  }
}

struct Gap {  // beamline/inc/rfgap_parameter.h:4-32
  double length, frequency, cell_length_over_beta_lambda, amplitude, phase_ref, phase_shift, energy_out,
      beta_center, dg, t, tp, sp, fit_beta_center, fit_beta_min, fit_t0, fit_t1, fit_t2, fit_t3, fit_t4,
      fit_t5, fit_s1, fit_s2, fit_s3, fit_s4, fit_s5;
};
// simulation_kernel.cu:226-288.  Returns the new d_wavelen (every thread writes it, :252).
double GapFirstHalf(hps_beam* b, double wavelen_in, double design_w_in, double phi_in, const Gap& gap,
                    double length, double qlen1, double qlen2, bool ccl) {
  const long n = b->n;
  const double mass = b->mass;
  const double wave_len = CLIGHT / gap.frequency;
  (void)wavelen_in;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    if (b->loss[i] != 0) continue;
    double x = b->x[i], y = b->y[i], xp = b->xp[i], yp = b->yp[i], phi = b->phi[i], w = b->w[i];
    double cell_len = length + qlen1 * 0.5 + qlen2 * 0.5;
    double dd1 = 0.5 * cell_len - gap.dg - 0.5 * qlen1;
    b->x[i] = x + dd1 * xp;
    b->y[i] = y + dd1 * yp;
    if (gap.amplitude == 0.0) {
      double gm1 = w / mass;
      double bt = std::sqrt(gm1 * (gm1 + 2.0)) / (gm1 + 1.0);
      b->phi[i] += TWOPI * dd1 / (wave_len * bt);
    } else {
      double gam = design_w_in / mass;
      double beta_in = std::sqrt((gam + 2.0) * gam) / (gam + 1.0);
      gam = w / mass;
      double beta = std::sqrt((gam + 2.0) * gam) / (gam + 1.0);
      if (!ccl) {
        double dps = gap.phase_ref - phi_in;
        double d_cell_len = cell_len - gap.beta_center * wave_len;
        double dphi_qlen1 = TWOPI * 0.5 * qlen1 / (beta * wave_len);
        double dphi_clen = 0.0;
        if (d_cell_len > 1e-15 || d_cell_len < -1e-15) dphi_clen = PI * d_cell_len / (beta * wave_len);
        b->phi[i] = phi + (1.0 - (beta - beta_in) / beta) * (PI * gap.cell_length_over_beta_lambda + dps) +
                    dphi_clen - dphi_qlen1;
      } else {
        double dps = gap.phase_ref - phi_in;
        double betag = cell_len / (gap.cell_length_over_beta_lambda * wave_len);
        b->phi[i] = phi + (1.0 - (beta - beta_in) * betag / (beta * beta_in)) *
                              (PI * gap.cell_length_over_beta_lambda + dps);
      }
    }
  }
  // the write happens only inside the loss==0 branch; with at least one live particle it lands
  bool any = false;
  for (uint i = 0; i < b->n && !any; ++i) any = b->loss[i] == 0;
  if (b->ctx->nranks > 1) any = xr1(b, any ? 1.0 : 0.0, 1) != 0.0;
  return any ? wave_len : wavelen_in;
}
// simulation_kernel.cu:290-308
void TransitTimeFactors(double* t, double* tp, double* sp, double rbeta, const Gap& g) {
  double beta = rbeta;
  const double betag = g.fit_beta_center;
  if (rbeta > 1.1 * betag) beta = betag;
  else if (rbeta < g.fit_beta_min) beta = g.fit_beta_min;
  *t = g.fit_t0 + beta * (g.fit_t1 + beta * (g.fit_t2 + beta * (g.fit_t3 + beta * (g.fit_t4 + beta * g.fit_t5))));
  double coef = beta * beta / (TWOPI * betag);
  *tp = coef * (g.fit_t1 + beta * (2 * g.fit_t2 + beta * (3 * g.fit_t3 + beta * (4 * g.fit_t4 + 5 * beta * g.fit_t5))));
  *sp = -coef * (g.fit_s1 + beta * (2 * g.fit_s2 + beta * (3 * g.fit_s3 + beta * (4 * g.fit_s4 + 5 * beta * g.fit_s5))));
}
// simulation_kernel.cu:310-497
void GapSecondHalf(hps_beam* b, double wavelen, double design_w_in, double phi_in, const Gap& gap,
                   double length, uint ccl_cell_num, double qlen1, double qlen2, bool ccl) {
  const long n = b->n;
  const double mass = b->mass;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    if (b->loss[i] != 0) continue;
    double cell_len = length + qlen1 * 0.5 + qlen2 * 0.5;
    double dd2 = 0.5 * cell_len + gap.dg - qlen2 * 0.5;
    double x = b->x[i], y = b->y[i], xp = b->xp[i], yp = b->yp[i];
    double inv_mass = 1.0 / mass;
    double phi = b->phi[i], w = b->w[i];
    double gm_in = design_w_in * inv_mass + 1.0;
    double beta_in = std::sqrt(1.0 - 1.0 / (gm_in * gm_in));
    double gam = w * inv_mass;
    double bg1 = std::sqrt(gam * (gam + 2.0));
    double beta = bg1 / (gam + 1.0);
    double t_fac, tp_fac, sp_fac;
    TransitTimeFactors(&t_fac, &tp_fac, &sp_fac, beta, gap);
    if (gap.amplitude == 0.0 || (ccl && beta < gap.fit_beta_min)) {
      b->x[i] = x + dd2 * xp;
      b->y[i] = y + dd2 * yp;
      double gm1 = b->w[i] / mass;
      double bt = std::sqrt(gm1 * (gm1 + 2.0)) / (gm1 + 1.0);
      b->phi[i] += TWOPI * dd2 / (wavelen * bt);
      continue;
    }
    double wave_len = CLIGHT / gap.frequency;
    double cl = cell_len / gap.cell_length_over_beta_lambda;
    double betag_ccl = cl / wave_len;
    double betag = ccl ? gap.beta_center : betag_ccl;
    double betag2 = betag * betag;
    double gmg = 1.0 / std::sqrt(1.0 - betag2);
    double inv_gmg2 = 1.0 / (gmg * gmg);
    double q = b->charge; q = q > 0.0 ? q : -q;
    double gm_out = gap.energy_out * inv_mass + 1.0;
    double beta_out = std::sqrt(1.0 - 1.0 / (gm_out * gm_out));
    double efl = gap.amplitude * cell_len;
    double concomm = q * efl / (mass * betag2);
    double concomm1 = concomm * inv_gmg2 / gmg;
    double con1 = PI * concomm1;
    double con2nex = concomm * inv_gmg2;
    double con5nex, tw, inv_fkis, inv_gmgcl;
    if (ccl) {
      inv_gmgcl = 1.0 / (gmg * betag * wave_len);
      inv_fkis = betag * wave_len / TWOPI;
      con5nex = concomm * (t_fac * inv_fkis - betag_ccl * wave_len * tp_fac * inv_gmg2);
      tw = t_fac - TWOPI * tp_fac * (beta_in / beta - 1.0) * betag_ccl / beta_in;
    } else {
      inv_gmgcl = 1.0 / (gmg * cl);
      inv_fkis = cl / TWOPI;
      con5nex = concomm * (t_fac * inv_fkis - beta_in * wave_len * tp_fac * inv_gmg2);
      tw = t_fac;
    }
    double con5nex1 = concomm1 * t_fac;
    double con5nex2 = -con2nex * t_fac * inv_fkis;
    double fkr = TWOPI * inv_gmgcl;
    double sinps, cosps;
    sincos(gap.phase_ref, &sinps, &cosps);
    double phi1, sinp1, cosp1;
    if (ccl) phi1 = phi + gap.phase_shift - TWOPI * gap.dg / (beta * wave_len) + ccl_cell_num * PI;
    else phi1 = phi + gap.phase_shift;
    sincos(phi1, &sinp1, &cosp1);
    double prome1 = -con1 * (tp_fac * (sinp1 - sinps) + sp_fac * (cosp1 - cosps));
    sincos(phi1 + prome1, &sinp1, &cosp1);
    double rs_p = x * x + y * y;
    double rrp = x * xp + y * yp;
    double ris = std::sqrt(rs_p), rpris = 0.0, bes1covr;
    double fkrris = fkr * ris;
    double b0c = Bessi0c(fkrris), b1c = Bessi1c(fkrris), b1p = Bessi1p(fkrris);
    if (ris != 0.0) bes1covr = b1c / ris; else bes1covr = fkr * 0.5;
    double delte = q * efl * (tw * b0c * cosp1 + ((tw - TWOPI * tp_fac) / gmg * bes1covr + tw / gmg * fkr * b1p -
                                                 tw * gmg * bes1covr) * rrp * sinp1);
    w += delte;
    gam = w * inv_mass;
    double bg = std::sqrt(gam * (gam + 2.0));
    beta = bg / (gam + 1.0);
    double bgr = bg1 / bg;
    double onema = 1.0 - (con5nex * bes1covr + con5nex1 * b1p + con5nex2 * bes1covr) * cosp1;
    double dovbl = con2nex * bes1covr * tw * sinp1;
    double bgrovonema = bgr / onema;
    xp = -dovbl * x + bgrovonema * xp;
    x = x * onema + dd2 * xp;
    yp = -dovbl * y + bgrovonema * yp;
    y = y * onema + dd2 * yp;
    if (rs_p > 0.0) rpris = rrp / ris;
    double prome = 2.0 * con1 * (-tp_fac * b0c + ris * tw * b1c * inv_gmgcl) * (sinp1 - sinps) -
                   efl * inv_mass * tw * b1c * rpris * (cosp1 - cosps);
    double phi_new;
    if (ccl) {
      double dps = phi_in - gap.phase_ref;
      phi_new = phi + (1.0 - (beta - beta_out) * betag_ccl / (beta * beta_out)) *
                          (PI * gap.cell_length_over_beta_lambda + dps) + prome;
    } else {
      double d_cell_len = cell_len - gap.beta_center * wave_len;
      double dphi_qlen2 = TWOPI * 0.5 * qlen2 / (beta * wave_len);
      double dphi_clen = 0.0;
      if (d_cell_len > 1e-15 || d_cell_len < -1e-15) dphi_clen = PI * d_cell_len / (beta * wave_len);
      double dps = gap.phase_ref - phi_in;
      phi_new = phi + (1.0 - (beta - beta_out) / beta) * (PI * gap.cell_length_over_beta_lambda + dps) + prome +
                dphi_clen - dphi_qlen2;
    }
    if (w <= 0.0 || x != x || y != y || xp != xp || yp != yp || w != w || phi_new != phi_new) {
      b->loss[i] = 33333333;
    } else {
      b->x[i] = x; b->y[i] = y; b->xp[i] = xp; b->yp[i] = yp; b->w[i] = w; b->phi[i] = phi_new;
    }
  }
}
// simulation_kernel.cu:511-529
void RotationMap(hps_beam* b, double angle) {
  const double cosa = std::cos(angle), sina = std::sin(angle);
  for (uint i = 0; i < b->n; ++i) {
    if (b->loss[i] != 0) continue;
    double x = b->x[i], y = b->y[i], xp = b->xp[i], yp = b->yp[i];
    b->x[i] = x * cosa + y * sina;
    b->y[i] = y * cosa - x * sina;
    b->xp[i] = xp * cosa + yp * sina;
    b->yp[i] = yp * cosa - xp * sina;
  }
}
// simulation_kernel.cu:543-575 (particle 0 exempt, centred on the beam centroid)
void CircularAperture(hps_beam* b, double aper, uint idx) {
  const double cx = b->s.v[HPS_STAT_X_AVG], cy = b->s.v[HPS_STAT_Y_AVG];
  for (uint i = 0; i < b->n; ++i) {
    if (b->gi(i) == 0 || b->loss[i] != 0) continue;
    double aper2 = aper * aper;
    double x = b->x[i] - cx, y = b->y[i] - cy;
    if (x * x + y * y > aper2) b->loss[i] = idx;
  }
}
void RectangularAperture(hps_beam* b, double xl, double xr, double yt, double yb, uint idx) {
  const double cx = b->s.v[HPS_STAT_X_AVG], cy = b->s.v[HPS_STAT_Y_AVG];
  for (uint i = 0; i < b->n; ++i) {
    if (b->gi(i) == 0 || b->loss[i] != 0) continue;
    double xmin = cx - xl, xmax = cx + xr, ymin = cy - yb, ymax = cy + yt;
    double x = b->x[i], y = b->y[i];
    if (x < xmin || x > xmax || y < ymin || y > ymax) b->loss[i] = idx;
  }
}
// simulation_kernel.cu:603-805
struct Dip { double radius, angle, edge_in, edge_out, half_gap, k1, k2, field_index, kinetic_energy; };
void DipoleHalf(hps_beam* b, double wavelen, const Dip& d, bool second) {
  const long n = b->n;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    if (b->loss[i] != 0) continue;
    double x = b->x[i], xp = b->xp[i], y = b->y[i], yp = b->yp[i];
    double h = 1.0 / d.radius;
    double hf_g = d.half_gap, k1 = d.k1, k2 = d.k2, findx = d.field_index, design_w = d.kinetic_energy;
    if (!second) {
      double b1 = d.edge_in;
      double c1 = 2.0 * hf_g * k1 * h;
      double sinb, inv_cosb;
      sincos(b1, &sinb, &inv_cosb);
      inv_cosb = 1.0 / inv_cosb;
      double psi = c1 * (1.0 + sinb * sinb) * inv_cosb * (1.0 - c1 * k2 * sinb * inv_cosb);
      double r21e = std::tan(b1) * h;
      double r43e = std::tan(b1 - psi) * h;
      xp += r21e * x;
      yp -= r43e * y;
    }
    double h2 = h * h;
    double dl = 0.5 * d.radius * d.angle;
    double yk2 = findx * h2;
    double yk = std::sqrt(std::fabs(yk2));
    double inv_yk = 1.0 / yk;
    double ykl = yk * dl;
    double xk2 = h2 - yk2;
    double inv_xk2 = 1.0 / xk2;
    double xk = std::sqrt(std::fabs(xk2));
    double inv_xk = 1.0 / xk;
    double xkl = xk * dl;
    double cxkl = 1.0, sxokx = dl;
    if (xk2 > 0.0) { cxkl = std::cos(xkl); sxokx = std::sin(xkl) * inv_xk; }
    else if (xk2 < 0.0) { cxkl = std::cosh(xkl); sxokx = std::sinh(xkl) * inv_xk; }
    double cykl = 1.0, syoky = dl;
    if (yk2 > 0.0) { cykl = std::cos(ykl); syoky = std::sin(ykl) * inv_yk; }
    else if (yk2 < 0.0) { cykl = std::cosh(ykl); syoky = std::sinh(ykl) * inv_yk; }
    double r11 = cxkl, r12 = sxokx, r21 = -xk2 * sxokx, r22 = cxkl;
    double r33 = cykl, r34 = syoky, r43 = -yk2 * syoky, r44 = cykl;
    double r26 = h * sxokx, r16 = h * (1.0 - cxkl) * inv_xk2;
    double r51 = -r26, r52 = -r16;
    double gamma = design_w / b->mass + 1.0;
    double beta2 = 1.0 - 1.0 / (gamma * gamma);
    double beta = std::sqrt(beta2);
    double r56 = -h2 * (dl * beta2 - sxokx) * inv_xk2 + dl * (1.0 - h2 * inv_xk2) / (gamma * gamma);
    double dpop = gamma / (design_w * (gamma + 1.0)) * (b->w[i] - design_w);
    if (!second) {
      b->x[i] = r11 * x + r12 * xp + r16 * dpop;
      b->xp[i] = r21 * x + r22 * xp + r26 * dpop;
      b->y[i] = r33 * y + r34 * yp;
      b->yp[i] = r43 * y + r44 * yp;
      b->phi[i] += TWOPI * (-(r51 * x + r52 * xp + r56 * dpop) + dl) / (beta * wavelen);
    } else {
      b->phi[i] += TWOPI * (-(r51 * x + r52 * xp + r56 * dpop) + dl) / (beta * wavelen);
      double x_new = r11 * x + r12 * xp + r16 * dpop;
      xp = r21 * x + r22 * xp + r26 * dpop;
      double y_new = r33 * y + r34 * yp;
      yp = r43 * y + r44 * yp;
      b->x[i] = x_new; b->y[i] = y_new;
      double b2 = d.edge_out;
      double c1 = 2.0 * hf_g * k1 * h;
      double sinb = std::sin(b2), inv_cosb = 1.0 / std::cos(b2);
      double psi = c1 * (1.0 + sinb * sinb) * inv_cosb * (1.0 - c1 * k2 * sinb * inv_cosb);
      double r21e = std::tan(b2) * h, r43e = std::tan(b2 - psi) * h;
      b->xp[i] = xp + r21e * x_new;
      b->yp[i] = yp - r43e * y_new;
    }
  }
}
// simulation_kernel.cu:807-871
void BuncherMap(hps_beam* b, double wavelen, double w_ref, double freq, double voltage, double phase, double aper,
                double freq_ratio, uint idx) {
  const double mass = b->mass;
  for (uint i = 0; i < b->n; ++i) {
    if (b->loss[i] != 0) continue;
    double x = b->x[i], y = b->y[i], xp = b->xp[i], yp = b->yp[i];
    double phi = b->phi[i] * freq_ratio, w = b->w[i];
    double r2 = x * x + y * y, aper2 = aper * aper;
    if (aper2 != 0.0 && r2 > aper2) { b->loss[i] = idx; continue; }
    double h = freq * wavelen / CLIGHT;
    double gm1 = w_ref / mass, gm = gm1 + 1.0;
    double btgm = std::sqrt(gm1 * (gm1 + 2.0));
    double beta = btgm / gm;
    double k = h * TWOPI / (btgm * wavelen);
    double k2 = k * k;
    double q = b->charge; q = q > 0.0 ? q : -q;
    double con1 = TWOPI * h * voltage * q / (beta * wavelen * mass);
    double arg = k2 * r2 * 0.25;
    double ph = h * phi + phase;
    double dw = q * voltage * std::cos(ph) * (1.0 + arg * (1.0 + arg * (0.25 + arg / 36.0)));
    double wb_par = w + 0.5 * dw, wf = w + dw;
    if (wb_par <= 0.0 || wf <= 0.0) { b->loss[i] = idx; continue; }
    double gmtmp = wb_par / mass;
    double bav = std::sqrt(gmtmp * (gmtmp + 2.0)) / (gmtmp + 1.0);
    gmtmp = w / mass;
    double bg = std::sqrt(gmtmp * (gmtmp + 2.0));
    gmtmp = wf / mass;
    double rbgf = rsqrt_(gmtmp * (gmtmp + 2.0));
    double del = -con1 * std::sin(ph) * (1.0 - bav * beta) * (0.5 + arg * (0.25 + arg / 24.0)) / bav;
    b->xp[i] = (xp * bg + del * x) * rbgf;
    b->yp[i] = (yp * bg + del * y) * rbgf;
    b->w[i] = wf;
  }
}

// ---------------------------------------------------------------------- SCHEFF
void AllocTables(hps_scheff* s) {  // space_charge.cu:27-61
  s->bins.assign((size_t)s->nx * s->nz, 0.0);
  s->green.assign((size_t)s->nx * s->nz * (s->nx + 1) * 2, 0.0);
  s->field.assign((size_t)(s->nz + 1) * (s->nx + 1) * 2, 0.0);
}
void SetMeshSize(hps_scheff* s, uint nx, uint nz) {  // space_charge.cpp:62-74
  if (nx != s->nx || nz != s->nz) { s->nx = nx; s->nz = nz; AllocTables(s); }
}
// space_charge_kernel.cu:39-100
void ellip(double g, double* ee, double* ek) {
  double h = std::log(g);
  *ee = 1.0 + g * (0.4630106 - 0.2452740 * h + g * (0.1077857 - 0.04125321 * h));
  *ek = 1.38629436 - 0.5 * h + g * (0.1119697 - 0.1213486 * h + g * (0.07253230 - 0.02887472 * h));
}
void UpdateGreenTable(hps_scheff* s, double freq, double beta, uint num_part, double current_ov_mc2, uint adj_bunch) {
  const uint nr = s->nx, nz = s->nz;
  const double gm = rsqrt_(1.0 - beta * beta);
  const double dr = s->dr, dzg = s->dz * gm;
  const long total = (long)(nr + 1) * nr * nz;
#pragma omp parallel for schedule(static)
  for (long cnt = 0; cnt < total; ++cnt) {
    uint jz = cnt % nz, k = (cnt / nz) % nr, irp = cnt / ((long)nz * nr);
    double rs = std::sqrt(0.5 * (k * k + (k + 1) * (k + 1))) * dr;
    double zs = 0.5 * dzg, zp = (jz + 1) * dzg, rp = irp * dr;
    double d = zp - zs, d2 = d * d;
    double c = (rp - rs) * (rp - rs), bb = (rp + rs) * (rp + rs);
    double rsrp = rs * rp;
    double a = 4.0 * rsrp / (bb + d2);
    double g = 1.0 - a, ee, ek;
    ellip(g, &ee, &ek);
    double er1 = 0.0;
    a = std::sqrt(bb + d2);
    double rs2mrp2 = rs * rs - rp * rp;
    if (rp != 0.0) er1 = (ek - (rs2mrp2 + d2) * ee / (c + d2)) / (2.0 * rp * a);
    double ez1 = d * ee / (a * (c + d2));
    double wave_len = CLIGHT / freq;
    double pl = beta * gm * wave_len;
    for (int i = 1; i <= (int)adj_bunch; ++i) {
      double xi = i;
      for (int j = 0; j < 2; ++j) {
        double d1 = d - xi * pl;
        d2 = d1 * d1;
        a = 4.0 * rsrp / (bb + d2);
        g = 1.0 - a;
        ellip(g, &ee, &ek);
        a = std::sqrt(bb + d2);
        if (rp != 0.0) er1 += (ek - (rs2mrp2 + d2) * ee / (c + d2)) / (2.0 * rp * a);
        ez1 += d1 * ee / (a * (c + d2));
        xi = -xi;
      }
    }
    double c1 = 5721.67 * current_ov_mc2 / (freq * 1e6 * (double)num_part);
    s->green[2 * cnt] = c1 * er1;
    s->green[2 * cnt + 1] = c1 * ez1;
  }
}
// space_charge_kernel.cu:123-224 (sequential, deterministic; the reference uses atomics)
void Deposit(hps_scheff* s, const hps_beam* b, double beta, double freq) {
  const uint nx = s->nx, nz = s->nz;
  s->num_on_mesh = 0;
  const double phimc = b->s.v[HPS_STAT_PHI_AVG_R], xbar = b->s.v[HPS_STAT_X_AVG_GOOD], ybar = b->s.v[HPS_STAT_Y_AVG_GOOD];
  const double epsq = b->s.v[HPS_STAT_X_SIG_GOOD] / b->s.v[HPS_STAT_Y_SIG_GOOD];
  for (uint np = 0; np < b->n; ++np) {
    if (!(b->loss[np] == 0 && b->lloss[np] == 0)) continue;
    double x = b->x[np], y = b->y[np], z = b->phi_r[np];
    double gm = rsqrt_(1.0 - beta * beta);
    double dr = s->dr, dzg = s->dz * gm;
    double wave_len = CLIGHT / freq;
    double hl = nz * 0.5 * dzg;
    double c2 = beta * gm * wave_len / TWOPI;
    double rsq_sh = (x - xbar) * (x - xbar) / epsq + (y - ybar) * (y - ybar) * epsq;
    int i = cuda_d2i(std::sqrt(rsq_sh) / dr + 1);
    if (!(i <= (int)nx)) continue;
    double ddz = z - phimc;
    while (ddz > PI || ddz < -PI) { if (ddz > 0.0) ddz -= TWOPI; else ddz += TWOPI; }
    z = -c2 * ddz;
    if (!(std::fabs(z) < hl)) continue;
    double zz = z + hl;
    int jm1 = cuda_d2i(zz / dzg + 1);
    int i1 = i + 1;
    double drdr = dr * dr;
    if (rsq_sh < 0.5 * drdr * (double)(i * i + (i - 1) * (i - 1))) i1 = i - 1;
    if (i1 < 1) i1 = 1;
    if (i1 > (int)nx) i1 = nx;
    int j1 = jm1 + 1;
    if (zz < (jm1 - 1 + 0.5) * dzg) j1 = jm1 - 1;
    if (j1 < 1) j1 = 1;
    if (j1 > (int)nz) j1 = nz;
    double a, cc;
    if (i1 == i) a = 1.0;
    else {
      double sqrtmp = std::sqrt(4.0 * rsq_sh / drdr - 1.0);
      double rminsq = 0.25 * drdr * (sqrtmp - 1.0) * (sqrtmp - 1.0);
      double rmaxsq = rminsq + sqrtmp * drdr;
      if (i1 < i) a = (rmaxsq - (i - 1) * (i - 1) * drdr) / (rmaxsq - rminsq);
      else a = ((i1 - 1) * (i1 - 1) * drdr - rminsq) / (rmaxsq - rminsq);
    }
    cc = 1.0;
    if (j1 != jm1) cc = (zz - (j1 - 1 + 0.5) * dzg) / ((jm1 - j1) * dzg);
    double bb = 1.0 - a, d = 1.0 - cc;
    int k = (jm1 - 1) * nx + i;
    if (jm1 <= (int)nz) {  // jm1 == nz+1 (zz/dzg rounding up to nz) writes past the table in the reference
      s->bins[k - 1] += a * cc;
      k += i1 - i;
      s->bins[k - 1] += bb * cc;
    }
    k = (j1 - 1) * nx + i;
    s->bins[k - 1] += a * d;
    k += i1 - i;
    s->bins[k - 1] += bb * d;
    s->num_on_mesh += 1;
  }
  if (b->ctx->nranks > 1) {  // the r-z charge mesh and the on-mesh count are the only SCHEFF data that cross ranks
    xr(b, s->bins.data(), (int)s->bins.size());
    s->num_on_mesh = (uint)xr1(b, s->num_on_mesh);
  }
}
// space_charge_kernel.cu:472-568 (the -DDOUBLE_PRECISION field solve)
void UpdateFieldTable(hps_scheff* s) {
  const uint nr = s->nx, nz = s->nz, fld_sz = nr * nz;
  const long nodes = (long)(nz + 1) * (nr + 1);
#pragma omp parallel for schedule(static)
  for (long node = 0; node < nodes; ++node) {
    uint zp = node / (nr + 1), rp = node % (nr + 1);
    double tx = 0.0, tz = 0.0;
    for (uint bin = 0; bin < fld_sz; ++bin) {
      double v = s->bins[bin];
      if (v == 0.0) continue;
      uint rs = bin % nr, zs = bin / nr, dz;
      double sign = 1.0;
      if (zp < nz && zs >= zp) { dz = zs - zp; sign = -1.0; }
      else dz = zp - zs - 1;
      size_t gi = (size_t)dz + (size_t)rs * nz + (size_t)rp * fld_sz;
      tx += v * s->green[2 * gi];
      tz += sign * v * s->green[2 * gi + 1];
    }
    s->field[2 * node] += tx;
    s->field[2 * node + 1] += tz;
  }
}
// space_charge_kernel.cu:573-706
void ApplyKick(hps_scheff* s, hps_beam* b, double current, double freq, double beta, uint adj_bunch, double length,
               double ratio_r, double ratio_z, double ratio_q, double ratio_gm) {
  const uint nx = s->nx, nz = s->nz, n = b->ng, nloc = b->n;
  // H2: launched <<<grid, blck/4>>> with `while(indx > 0 && ...)`: global thread 0 never runs
  const uint np2 = NextPow2(n);
  const uint stride = np2 >= 4 ? np2 / 4 : 1;
  const double phimc = b->s.v[HPS_STAT_PHI_AVG_R], xbar = b->s.v[HPS_STAT_X_AVG_GOOD], ybar = b->s.v[HPS_STAT_Y_AVG_GOOD];
  const double epsq = b->s.v[HPS_STAT_X_SIG_GOOD] / b->s.v[HPS_STAT_Y_SIG_GOOD];
  const double mass = b->mass;
  const long ln = nloc;
#pragma omp parallel for schedule(static)
  for (long li = 0; li < ln; ++li) {
    const uint indx = (uint)li;
    const uint gindx = b->gi(indx);  // the quirks are tied to the index in the whole beam
    if (gindx == 0) continue;
    if (b->ctx->compat && gindx % stride == 0) continue;
    if (!(b->loss[indx] == 0 && b->lloss[indx] == 0)) continue;
    double x = b->x[indx], y = b->y[indx], xp = b->xp[indx], yp = b->yp[indx], phi = b->phi_r[indx], w = b->w[indx];
    double gmg = rsqrt_(1.0 - beta * beta);
    double dr = s->dr * ratio_r;
    double dzg = s->dz * gmg * ratio_z;
    double wave_len = CLIGHT / freq;
    double hl = nz * 0.5 * dzg;
    double c1 = 5721.67 * current / (freq * 1e6 * (double)n * mass);
    double c2 = beta * gmg * wave_len / TWOPI;
    double c3 = length / (beta * gmg);
    double xfac = 2.0 / (epsq + 1.0);
    double yfac = epsq * xfac;
    double gm = w / mass + 1.0;
    double bgz = std::sqrt((gm - 1.0) * (gm + 1.0));
    double bgx = bgz * xp, bgy = bgz * yp;
    double bgzstar = gmg * (bgz - beta * gm);
    double gstar;
    double xxbar = x - xbar, yybar = y - ybar;
    double r = std::sqrt(xxbar * xxbar / epsq + yybar * yybar * epsq);
    if (r == 0.0) r = 1.0e-8;
    double xovr = xxbar * xfac / r, yovr = yybar * yfac / r;
    double ddz = phi - phimc;
    while (ddz > PI || ddz < -PI) { if (ddz > 0.0) ddz -= TWOPI; else ddz += TWOPI; }
    double z = -c2 * ddz;
    double cbgr, cbgzs;
    if (r > nx * dr || std::fabs(z) > hl) {
      z = -c2 * (phi - phimc);
      double d = std::sqrt(z * z + r * r);
      double invd3 = 1.0 / (d * d * d);
      double rod3 = r * invd3, zod3 = z * invd3;
      double pl = beta * gmg * wave_len;
      for (int nb = 1; nb <= (int)adj_bunch; ++nb) {
        int xi = nb;
        for (int lr = 1; lr <= 2; ++lr) {
          double sdist = z + xi * pl;
          d = std::sqrt(sdist * sdist + r * r);
          invd3 = 1.0 / (d * d * d);
          rod3 += r * invd3;
          zod3 += sdist * invd3;
          xi = -xi;
        }
      }
      double cbgcoef = c1 * c3 * PI * 0.5 * s->num_on_mesh;
      cbgr = cbgcoef * rod3;
      cbgzs = cbgcoef * zod3;
    } else {
      uint i = 1 + cuda_d2i(r / dr);
      double a = r / dr - (i - 1.0), bb = 1.0 - a;
      uint j = 1 + cuda_d2i((z + hl) / dzg);
      double c = (z + hl) / dzg - (j - 1.0), d = 1.0 - c;
      uint l = (j - 1) * (nx + 1) + i;
      uint m = l + (nx + 1);
      uint tbl_sz = (nx + 1) * (nz + 1);
      if (m >= tbl_sz) m = tbl_sz - 1;
      double er_ratio = ratio_q / (ratio_r * ratio_z * ratio_gm * ratio_gm);
      double ez_ratio = ratio_q / (ratio_z * ratio_z);
      const double* f = s->field.data();
      cbgr = er_ratio * c3 * (d * (a * f[2 * l] + bb * f[2 * (l - 1)]) + c * (a * f[2 * m] + bb * f[2 * (m - 1)]));
      cbgzs = ez_ratio * c3 * (d * (a * f[2 * l + 1] + bb * f[2 * (l - 1) + 1]) + c * (a * f[2 * m + 1] + bb * f[2 * (m - 1) + 1]));
    }
    bgx += cbgr * xovr;
    bgy += cbgr * yovr;
    bgzstar += cbgzs;
    gstar = 1.0 + 0.5 * bgzstar * bgzstar;
    double bgzf = gmg * (bgzstar + beta * gstar);
    if (bgzf <= 0.0) b->loss[indx] = 9999;
    else {
      b->xp[indx] = bgx / bgzf;
      b->yp[indx] = bgy / bgzf;
      b->w[indx] *= 1.0 + (gm + 1.0) * (bgzf - bgz) / (gm * bgz);
    }
  }
}
// space_charge.cpp:82-139 + space_charge.cu:122-138 (host policy of one kick)
void ScheffStart(hps_scheff* s, hps_beam* b, double length) {
  UpdateStatForSpaceCharge(b);
  uint good_cnt = b->s.num_good;
  double sigphi = b->s.v[HPS_STAT_PHI_SIG_GOOD];
  double sigx = b->ng < 2 ? 0.0 : b->s.v[HPS_STAT_X_SIG_GOOD];
  double sigy = b->ng < 2 ? 0.0 : b->s.v[HPS_STAT_Y_SIG_GOOD];
  double ke = b->s.v[HPS_STAT_W_AVG_GOOD];
  double gamma = (double)ke / b->mass + 1.0;
  double beta = std::sqrt(1.0 - 1.0 / (gamma * gamma));
  double lmbd = 299.792458 / b->freq;
  double sigz = sigphi * 0.5 / 3.14159265 * beta * lmbd;
  double sigr = std::sqrt(sigx * sigx + sigy * sigy);
  double cratior = sigr * sigz;
  double cratioz = sigz * sigz * gamma * gamma;
  double cratior_old = s->sigr_old * s->sigz_old;
  double cratioz_old = s->sigz_old * s->sigz_old * s->gm_old * s->gm_old;
  double adj_bunch_d;
  if (s->adj_bunch_w == 0.0 || ke < s->adj_bunch_w) adj_bunch_d = s->adj_bunch; else adj_bunch_d = 0;
  if (s->mesh_w != 0.0) {
    if (ke > s->mesh_w && (s->nx != s->nx_init / 2 || s->nz != s->nz_init / 2)) SetMeshSize(s, s->nx_init / 2, s->nz_init / 2);
    else if (ke <= s->mesh_w && (s->nx != s->nx_init || s->nz != s->nz_init)) SetMeshSize(s, s->nx_init, s->nz_init);
    if (ke > s->mesh_w && s->interval != s->interval_init * 4) s->interval = s->interval_init * 4;
    else if (ke <= s->mesh_w && s->interval != s->interval_init) s->interval = s->interval_init;
  }
  const uint adj_bunch = (uint)adj_bunch_d;
  const double freq = b->freq, frac = s->fraction;
  s->kicks++;
  if (cratior_old == 0.0 || cratioz_old == 0.0 || std::fabs(cratior - cratior_old) / cratior_old > s->remesh_val ||
      std::fabs(cratioz - cratioz_old) / cratioz_old > s->remesh_val) {
    s->remesh++;
    s->sigr_old = sigr; s->sigz_old = sigz; s->gm_old = gamma; s->good_cnt_old = good_cnt;
    std::fill(s->bins.begin(), s->bins.end(), 0.0);
    std::fill(s->field.begin(), s->field.end(), 0.0);
    // UpdateMeshKernelCall + UpdateScheffMeshSizeKernel (space_charge_kernel.cu:19-32)
    double bl = beta * CLIGHT / freq;
    if (adj_bunch != 0) {
      UpdateMaxRPhi(b);
      s->dr = 1.0 * b->s.v[HPS_STAT_R_MAX] / s->nx;
      s->dz = bl / s->nz;
    } else {
      // H6: the reference compares device POINTERS (x_sig_good > y_sig_good); allocation order makes it false
      s->dr = 3.0 * b->s.v[getenv("HPS_SIGX") ? HPS_STAT_X_SIG_GOOD : HPS_STAT_Y_SIG_GOOD] / s->nx;
      s->dz = 3.0 * bl * b->s.v[HPS_STAT_PHI_SIG_GOOD] / (s->nz * PI);  // H8: clamp is dead code
    }
    UpdateGreenTable(s, freq, beta, b->ng, b->current * frac / b->mass, adj_bunch);
    Deposit(s, b, beta, freq);
    UpdateFieldTable(s);
    ApplyKick(s, b, b->current * frac, freq, beta, adj_bunch, length, 1, 1, 1, 1.0);
  } else {
    ApplyKick(s, b, b->current * frac, freq, beta, adj_bunch, length, sigr / s->sigr_old, sigz / s->sigz_old,
              s->good_cnt_old ? (double)(good_cnt / s->good_cnt_old) : 0.0, gamma / s->gm_old);  // H7: uint division
  }
}

// ---------------------------------------------------------------------- engine
void Reset(hps_ctx* c, hps_beam* b) {  // simulation_engine.cu:101-111
  if (c->monitor) { c->monitor->monitor_num = 0; c->monitor->loss_num = 0; }
  c->ccl_cell_num = 0;
  c->design_w = b->design_w;
  c->wavelen = CLIGHT / b->freq;
}
Gap MakeGap(const hps_elem& e) { Gap g; std::memcpy(&g, e.p, sizeof(Gap)); return g; }
bool IsType(const hps_lattice* l, long i, uint t) { return i >= 0 && i < (long)l->e.size() && l->e[i].type == t; }

// simulation_engine.cu:163-175
void PrepGap(hps_ctx* c, hps_beam* b, const Gap& g) {
  if (b->freq != g.frequency) {
    c->wavelen = CLIGHT / g.frequency;
    ChangeFrequency(b, g.frequency);
  }
}

// simulation_engine.cu:34-82
void Update2DPlotData(hps_ctx* c, hps_beam* b) {
  hps_monitor* m = c->monitor;
  UpdateLoss(b);
  double loss_local = (double)(uint)(b->s.num_loss - m->loss_num);
  m->loss_num = (uint)((double)m->loss_num + loss_local);
  loss_local /= b->ng;
  loss_local = std::log10(loss_local + 1e-6);
  const double loss_ratio = (double)m->loss_num / b->ng;
  UpdateEmittance(b);
  const double w_avg = b->s.v[HPS_STAT_W_AVG] - from_ref_particle(b, b->w);
  const uint row = m->monitor_num++;
  if (row >= m->cap) return;  // (the reference writes past its arrays)
  auto put = [&](int f, double v) { m->data[(size_t)f * m->cap + row] = v; };
  put(HPS_MON_LOSS_LOCAL, loss_local); put(HPS_MON_LOSS_RATIO, loss_ratio);
  put(HPS_MON_XAVG, b->s.v[HPS_STAT_X_AVG]); put(HPS_MON_XSIG, b->s.v[HPS_STAT_X_SIG]);
  put(HPS_MON_XPAVG, b->s.v[HPS_STAT_XP_AVG]); put(HPS_MON_XPSIG, b->s.v[HPS_STAT_XP_SIG]);
  put(HPS_MON_YAVG, b->s.v[HPS_STAT_Y_AVG]); put(HPS_MON_YSIG, b->s.v[HPS_STAT_Y_SIG]);
  put(HPS_MON_YPAVG, b->s.v[HPS_STAT_YP_AVG]); put(HPS_MON_YPSIG, b->s.v[HPS_STAT_YP_SIG]);
  put(HPS_MON_PHIAVG, b->s.v[HPS_STAT_PHI_AVG_R]); put(HPS_MON_PHISIG, b->s.v[HPS_STAT_PHI_SIG_R]);
  put(HPS_MON_WSIG, b->s.v[HPS_STAT_W_SIG]);
  put(HPS_MON_XEMIT, b->s.v[HPS_STAT_X_EMIT]); put(HPS_MON_YEMIT, b->s.v[HPS_STAT_Y_EMIT]); put(HPS_MON_ZEMIT, b->s.v[HPS_STAT_Z_EMIT]);
  put(HPS_MON_WAVG, w_avg);
}

int VisitElement(hps_ctx* c, hps_beam* b, hps_lattice* lat, hps_scheff* sc, uint id) {
  const hps_elem& e = lat->e[id];
  const double* p = e.p;
  switch (e.type) {
    case HPS_ELEM_DRIFT: {  // simulation_engine.cu:125-145
      int nk = 1;
      double length = p[0];
      if (sc) nk = (int)std::ceil(length / sc->interval);
      double hl = 0.5 * length / nk;
      for (int k = 0; k < nk; ++k) {
        PartialDrift(b, c->wavelen, hl, p[1], id);
        if (sc) ScheffStart(sc, b, 2.0 * hl);
        PartialDrift(b, c->wavelen, hl, p[1], id);
      }
      break;
    }
    case HPS_ELEM_QUAD:  // :147-161
      HalfQuad(b, c->wavelen, p[0], p[1], p[2], id);
      if (c->monitor && (e.flags & HPS_FLAG_MONITOR)) Update2DPlotData(c, b);
      if (sc) ScheffStart(sc, b, p[0]);
      HalfQuad(b, c->wavelen, p[0], p[1], p[2], id);
      break;
    case HPS_ELEM_RFGAP_DTL: {  // :182-241
      Gap g = MakeGap(e);
      PrepGap(c, b, g);
      double phi_in = 0.0;
      const long n = lat->e.size();
      if (id == 1 || (id > 1 && !IsType(lat, (long)id - 2, HPS_ELEM_RFGAP_DTL))) {
        if ((long)id + 2 < n && IsType(lat, id + 2, HPS_ELEM_RFGAP_DTL))
          phi_in = g.phase_ref - 0.5 * (lat->e[id + 2].p[4] - g.phase_ref);
        else if ((long)id + 1 >= n)
          return fail(c, HPS_ERR_LATTICE, "DTL gap without a following quad");
        else phi_in = g.phase_ref;
      } else if (id > 1) {
        phi_in = 0.5 * (lat->e[id - 2].p[4] + g.phase_ref);
      } else return fail(c, HPS_ERR_LATTICE, "a quad is missing before the DTL gap");
      if (!IsType(lat, (long)id - 1, HPS_ELEM_QUAD) || !IsType(lat, (long)id + 1, HPS_ELEM_QUAD))
        return fail(c, HPS_ERR_LATTICE, "DTL gap needs quad neighbours");  // reference: null deref
      double q1 = lat->e[id - 1].p[0], q2 = lat->e[id + 1].p[0];
      c->wavelen = GapFirstHalf(b, c->wavelen, c->design_w, phi_in, g, p[0], q1, q2, false);
      if (sc) ScheffStart(sc, b, p[0]);
      GapSecondHalf(b, c->wavelen, c->design_w, phi_in, g, p[0], 0, q1, q2, false);
      c->design_w = g.energy_out;
      b->design_w = c->design_w;
      break;
    }
    case HPS_ELEM_RFGAP_CCL: {  // :243-293
      Gap g = MakeGap(e);
      PrepGap(c, b, g);
      if (id == 0 || !IsType(lat, (long)id - 1, HPS_ELEM_RFGAP_CCL)) c->ccl_cell_num = 0;
      double phi_in, phi_out;
      bool next_ccl = IsType(lat, (long)id + 1, HPS_ELEM_RFGAP_CCL);
      if (c->ccl_cell_num == 0) {
        if (!next_ccl) return fail(c, HPS_ERR_LATTICE, "single-cell CCL tank");  // reference: bad cast
        double cur = g.phase_ref, nxt = lat->e[id + 1].p[4];
        phi_in = cur - 0.5 * (nxt - cur);
        phi_out = 0.5 * (nxt + cur);
      } else if (!next_ccl) {
        double cur = g.phase_ref, prv = lat->e[id - 1].p[4];
        phi_in = 0.5 * (cur + prv);
        phi_out = cur + 0.5 * (cur - prv);
      } else {
        double cur = g.phase_ref;
        phi_in = 0.5 * (cur + lat->e[id - 1].p[4]);
        phi_out = 0.5 * (cur + lat->e[id + 1].p[4]);
      }
      c->wavelen = GapFirstHalf(b, c->wavelen, c->design_w, phi_in, g, p[0], 0.0, 0.0, true);
      if (sc) ScheffStart(sc, b, p[0]);
      GapSecondHalf(b, c->wavelen, c->design_w, phi_out, g, p[0], c->ccl_cell_num++, 0.0, 0.0, true);
      c->design_w = g.energy_out;
      b->design_w = c->design_w;
      break;
    }
    case HPS_ELEM_ROTATION: RotationMap(b, p[0]); break;  // :295-299
    case HPS_ELEM_SPCH_COMP: if (sc) sc->fraction = p[0]; break;  // :301-305
    case HPS_ELEM_CAPERTURE:  // :307-316
      if (!(e.flags & HPS_FLAG_ON)) break;
      UpdateLoss(b); UpdateAvgXY(b);
      CircularAperture(b, p[0], id);
      break;
    case HPS_ELEM_RAPERTURE:  // :318-329
      if (!(e.flags & HPS_FLAG_ON)) break;
      UpdateLoss(b); UpdateAvgXY(b);
      RectangularAperture(b, p[0], p[1], p[2], p[3], id);
      break;
    case HPS_ELEM_DIAGNOSTICS:  // :331-335
      if (c->monitor && (e.flags & HPS_FLAG_MONITOR)) Update2DPlotData(c, b);
      break;
    case HPS_ELEM_STEERER:  // :337-344
      if (p[0] > 1e-10 || p[1] > 1e-10) SteererMap(b, p[0], p[1]);
      break;
    case HPS_ELEM_BUNCHER: {  // :346-359
      if (!(e.flags & HPS_FLAG_ON)) break;
      double bfreq = p[1];
      c->wavelen = CLIGHT / bfreq;
      BuncherMap(b, c->wavelen, c->design_w, bfreq, p[0], p[2], p[3], bfreq / b->freq, id);
      break;
    }
    case HPS_ELEM_DIPOLE: {  // :361-373
      Dip d; std::memcpy(&d, p, sizeof(Dip));
      DipoleHalf(b, c->wavelen, d, false);
      if (sc) ScheffStart(sc, b, d.radius * d.angle);
      DipoleHalf(b, c->wavelen, d, true);
      break;
    }
    default: return fail(c, HPS_ERR_LATTICE, "unknown element type");
  }
  return HPS_OK;
}

void SaveSlot(hps_beam* b, int k) {
  Saved& s = b->slot[k];
  const std::vector<double>* src[6] = {&b->x, &b->xp, &b->y, &b->yp, &b->phi, &b->w};
  for (int i = 0; i < 6; ++i) s.c[i] = *src[i];
  s.loss = b->loss; s.lloss = b->lloss;
  s.design_w = b->design_w; s.freq = b->freq; s.current = b->current;
  s.valid = true;
}

}  // namespace

// =========================================================================== C ABI
extern "C" {

int hps_abi_version(void) { return HPS_ABI_VERSION; }
hps_ctx* hps_ctx_create(int device) { (void)device; return new hps_ctx(); }
void hps_ctx_destroy(hps_ctx* c) { delete c; }
const char* hps_last_error(const hps_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }
int hps_ctx_set_compat(hps_ctx* c, int v) { c->compat = v; return HPS_OK; }
int hps_ctx_set_math(hps_ctx*, int) { return HPS_OK; }
int hps_ctx_get_math(hps_ctx*) { return HPS_MATH_REFERENCE; }  // the oracle only knows the reference's expression order
int hps_comm_unique_id(void*) { return HPS_ERR_COMM; }
int hps_ctx_comm_init(hps_ctx* c, int, int nranks, const void*) {
  return nranks == 1 ? HPS_OK : fail(c, HPS_ERR_COMM, "the CPU oracle shards through hps_oracle_set_comm (tests only)");
}
// oracle-only: shard beams created afterwards over `nranks` processes; `fn` folds a vector over all ranks
int hps_oracle_set_comm(hps_ctx* c, int rank, int nranks, hps_oracle_allreduce_fn fn, void* user) {
  if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !fn)) return fail(c, HPS_ERR_ARG, "bad communicator");
  c->rank = rank; c->nranks = nranks; c->allreduce = fn; c->allreduce_user = user;
  return HPS_OK;
}
int hps_ctx_comm_mode(hps_ctx*) { return 0; }
int hps_ctx_comm_stats(hps_ctx*, uint64_t out[4]) { for (int k = 0; k < 4; ++k) out[k] = 0; return HPS_OK; }
int hps_ctx_synchronize(hps_ctx*) { return HPS_OK; }
void* hps_host_alloc(size_t bytes) { return std::malloc(bytes); }
void hps_host_free(void* p) { std::free(p); }
uint64_t hps_ctx_launch_count(hps_ctx*) { return 0; }
int hps_ctx_set_profile(hps_ctx*, int) { return HPS_OK; }
int hps_ctx_event_record(hps_ctx*, int) { return HPS_OK; }
int hps_ctx_event_elapsed(hps_ctx*, int, int, double* ms) { if (ms) *ms = 0.0; return HPS_OK; }
int hps_ctx_flush_l2(hps_ctx*) { return HPS_OK; }
int hps_ctx_get_profile(hps_ctx*, double* ms, uint64_t* n) {
  for (int k = 0; k < HPS_PROF_NCLASS; ++k) { ms[k] = 0.0; n[k] = 0; }
  return HPS_OK;
}
// oracle-only: OpenMP thread count (0 = runtime default)
int hps_oracle_set_threads(hps_ctx* c, int n) {
  c->nthreads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
  return HPS_OK;
}
int hps_oracle_get_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

hps_beam* hps_beam_create(hps_ctx* c, uint64_t n, double mass, double charge, double current) {
  if (!c || n == 0 || n > 0xffffffffull) { fail(c, HPS_ERR_ARG, "bad particle count"); return nullptr; }
  hps_beam* b = new hps_beam();
  b->ctx = c; b->ng = (uint)n; b->mass = mass; b->charge = charge; b->current = current;
  // interleaved ownership like the product (hps_beam_create in hpsim_b200/csrc/hps_capi.cpp)
  b->off = (uint)c->rank; b->stride = (uint)c->nranks;
  b->n = n > (uint64_t)c->rank ? (uint)((n - c->rank + c->nranks - 1) / c->nranks) : 0;
  b->freq = 0.0; b->design_w = 0.0;
  for (auto* v : {&b->x, &b->xp, &b->y, &b->yp, &b->phi, &b->phi_r, &b->w}) v->assign(b->n, 0.0);
  b->loss.assign(b->n, 0); b->lloss.assign(b->n, 0);
  std::memset(&b->s, 0, sizeof(b->s));
  return b;
}
void hps_beam_destroy(hps_beam* b) { delete b; }
int hps_beam_local_range(const hps_beam* b, uint64_t* off, uint64_t* nl) { *off = b->off; *nl = b->n; return HPS_OK; }
uint64_t hps_beam_local_stride(const hps_beam* b) { return b->stride; }

int hps_beam_upload(hps_beam* b, const double* x, const double* xp, const double* y, const double* yp,
                    const double* phi, const double* w, const uint32_t* loss, const uint32_t* lloss) {
  const size_t n = b->n;
  std::copy(x, x + n, b->x.begin()); std::copy(xp, xp + n, b->xp.begin());
  std::copy(y, y + n, b->y.begin()); std::copy(yp, yp + n, b->yp.begin());
  std::copy(phi, phi + n, b->phi.begin()); std::copy(w, w + n, b->w.begin());
  const bool own0 = b->off == 0 && n > 0;  // the reference particle lives on the owner of global index 0
  if (own0) b->x[0] = b->xp[0] = b->y[0] = b->yp[0] = 0.0;  // beam.cpp:344-345
  if (loss) { std::copy(loss, loss + n, b->loss.begin()); if (own0) b->loss[0] = 0; } else std::fill(b->loss.begin(), b->loss.end(), 0);
  if (lloss) { std::copy(lloss, lloss + n, b->lloss.begin()); if (own0) b->lloss[0] = 0; } else std::fill(b->lloss.begin(), b->lloss.end(), 0);
  InitPhiAvgGood(b);
  return HPS_OK;
}
int hps_beam_download(hps_beam* b, double* x, double* xp, double* y, double* yp, double* phi, double* w,
                      uint32_t* loss, uint32_t* lloss) {
  UpdateLoss(b);
  if (x) std::copy(b->x.begin(), b->x.end(), x);
  if (xp) std::copy(b->xp.begin(), b->xp.end(), xp);
  if (y) std::copy(b->y.begin(), b->y.end(), y);
  if (yp) std::copy(b->yp.begin(), b->yp.end(), yp);
  if (phi) std::copy(b->phi.begin(), b->phi.end(), phi);
  if (w) std::copy(b->w.begin(), b->w.end(), w);
  if (loss) std::copy(b->loss.begin(), b->loss.end(), loss);
  if (lloss) std::copy(b->lloss.begin(), b->lloss.end(), lloss);
  return HPS_OK;
}
int hps_beam_download_phi_relative(hps_beam* b, int use_good, double* out) {
  UpdateRelativePhi(b, use_good != 0);
  std::copy(b->phi_r.begin(), b->phi_r.end(), out);
  return HPS_OK;
}

// beam.cpp:481-537 (glibc rand(); the reference is linked against the same libc)
int hps_beam_set_dc(hps_beam* b, double ax, double bx, double ex, double ay, double by, double ey, double dphi,
                    double sync_phi, double sync_w, uint32_t seed) {
  const uint n = b->ng;  // every rank draws the whole rand() sequence and keeps its own particles
  if (seed == 0) std::srand(std::time(NULL)); else std::srand(seed);
  double rs = 1.0;
  double gx = (1.0 + ax * ax) / bx, xmx = std::sqrt(ex / gx), xpmx = std::sqrt(ex * gx), dx = -ax / gx;
  double gy = (1.0 + ay * ay) / by, ymx = std::sqrt(ey / gy), ypmx = std::sqrt(ey * gy), dy = -ay / gy;
  for (uint g = 0; g < n; ++g) {
    double r1, r2, r3, r4, r5;
    do {
      do { r1 = 2.0 * std::rand() / RAND_MAX - 1.0; r2 = 2.0 * std::rand() / RAND_MAX - 1.0; } while (r1 * r1 + r2 * r2 > 1.0);
      do { r3 = 2.0 * std::rand() / RAND_MAX - 1.0; r4 = 2.0 * std::rand() / RAND_MAX - 1.0; } while (r3 * r3 + r4 * r4 > 1.0);
    } while (r1 * r1 + r2 * r2 + r3 * r3 + r4 * r4 > 1.0);
    r1 *= rs; r2 *= rs; r3 *= rs; r4 *= rs;
    r5 = 2.0 * (double)(g + 1.0) / (double)n - 1.0;
    if (g < b->off || (g - b->off) % b->stride != 0) continue;
    const uint i = (g - b->off) / b->stride;
    b->x[i] = 0.01 * (r1 * xmx + dx * r2 * xpmx);
    b->xp[i] = r2 * xpmx;
    b->y[i] = 0.01 * (r3 * ymx + dy * r4 * ypmx);
    b->yp[i] = r4 * ypmx;
    b->phi[i] = r5 * dphi + sync_phi;
    b->w[i] = sync_w;
  }
  if (b->off == 0 && b->n > 0) b->x[0] = b->xp[0] = b->y[0] = b->yp[0] = 0.0;
  std::fill(b->loss.begin(), b->loss.end(), 0);
  std::fill(b->lloss.begin(), b->lloss.end(), 0);
  InitPhiAvgGood(b);
  return HPS_OK;
}
// beam.cpp:391-463
int hps_beam_set_waterbag(hps_beam* b, double ax, double bx, double ex, double ay, double by, double ey, double az,
                          double bz, double ez, double sync_phi, double sync_w, double freq, uint32_t seed) {
  const uint n = b->ng;  // every rank draws the whole rand() sequence and keeps its own particles
  if (seed == 0) std::srand(std::time(NULL)); else std::srand(seed);
  double rs = 1.0;
  double gx = (1.0 + ax * ax) / bx, xmx = std::sqrt(ex / gx), xpmx = std::sqrt(ex * gx), dx = -ax / gx;
  double gy = (1.0 + ay * ay) / by, ymx = std::sqrt(ey / gy), ypmx = std::sqrt(ey * gy), dy = -ay / gy;
  double gz = (1.0 + az * az) / bz, zmx = std::sqrt(ez / gz), zpmx = ez / zmx, dz = -az / gz;
  double wave_len = 29979.2458 / freq;
  b->freq = freq;
  double gm = sync_w / b->mass + 1.0;
  double beta = std::sqrt(1.0 - 1.0 / (gm * gm));
  for (uint g = 0; g < n; ++g) {
    double r1, r2, r3, r4, r5, r6;
    do {
      do { r1 = 2.0 * std::rand() / RAND_MAX - 1.0; r2 = 2.0 * std::rand() / RAND_MAX - 1.0; } while (r1 * r1 + r2 * r2 > 1.0);
      do { r3 = 2.0 * std::rand() / RAND_MAX - 1.0; r4 = 2.0 * std::rand() / RAND_MAX - 1.0; } while (r3 * r3 + r4 * r4 > 1.0);
      do { r5 = 2.0 * std::rand() / RAND_MAX - 1.0; r6 = 2.0 * std::rand() / RAND_MAX - 1.0; } while (r5 * r5 + r6 * r6 > 1.0);
    } while (r1 * r1 + r2 * r2 + r3 * r3 + r4 * r4 + r5 * r5 + r6 * r6 > 1.0);
    r1 *= rs; r2 *= rs; r3 *= rs; r4 *= rs; r5 *= rs; r6 *= rs;
    if (g < b->off || (g - b->off) % b->stride != 0) continue;
    const uint i = (g - b->off) / b->stride;
    b->x[i] = 0.01 * (r1 * xmx + dx * r2 * xpmx);
    b->xp[i] = r2 * xpmx;
    b->y[i] = 0.01 * (r3 * ymx + dy * r4 * ypmx);
    b->yp[i] = r4 * ypmx;
    b->phi[i] = -(r5 * zmx + dz * zpmx * r6) * 2.0 * M_PI / (beta * wave_len) + sync_phi;
    b->w[i] = 2.0 * sync_w * zpmx * r6 + sync_w;
  }
  if (b->off == 0 && b->n > 0) b->x[0] = b->xp[0] = b->y[0] = b->yp[0] = 0.0;
  std::fill(b->loss.begin(), b->loss.end(), 0);
  std::fill(b->lloss.begin(), b->lloss.end(), 0);
  InitPhiAvgGood(b);
  return HPS_OK;
}

int hps_beam_save(hps_beam* b, int slot) {
  if (slot < 0 || slot > 1) return fail(b->ctx, HPS_ERR_ARG, "slot must be 0 or 1");
  SaveSlot(b, slot);
  return HPS_OK;
}
int hps_beam_restore(hps_beam* b, int slot) {  // beam.cpp:603-647
  if (slot < 0 || slot > 1) return fail(b->ctx, HPS_ERR_ARG, "slot must be 0 or 1");
  Saved& s = b->slot[slot];
  if (!s.valid) return fail(b->ctx, HPS_ERR_STATE, "no beam has been saved in this slot");
  b->x = s.c[0]; b->xp = s.c[1]; b->y = s.c[2]; b->yp = s.c[3]; b->phi = s.c[4]; b->w = s.c[5];
  b->loss = s.loss; b->lloss = s.lloss;
  b->design_w = s.design_w; b->freq = s.freq; b->current = s.current;
  UpdateAvg(b, b->phi, HPS_STAT_PHI_AVG_GOOD, true);  // H21: with the pre-restore num_good
  UpdateRelativePhi(b, true);
  return HPS_OK;
}
int hps_beam_set_ref_w(hps_beam* b, double w) { if (b->off == 0 && b->n > 0) b->w[0] = w; b->design_w = w; return HPS_OK; }
int hps_beam_set_ref_phi(hps_beam* b, double phi) { if (b->off == 0 && b->n > 0) b->phi[0] = phi; return HPS_OK; }
int hps_beam_set_frequency(hps_beam* b, double f) { b->freq = f; return HPS_OK; }
double hps_beam_get_ref_w(hps_beam* b) { return from_ref_particle(b, b->w); }
double hps_beam_get_ref_phi(hps_beam* b) { return from_ref_particle(b, b->phi); }
double hps_beam_get_frequency(const hps_beam* b) { return b->freq; }
double hps_beam_get_mass(const hps_beam* b) { return b->mass; }
double hps_beam_get_charge(const hps_beam* b) { return b->charge; }
double hps_beam_get_current(const hps_beam* b) { return b->current; }
uint64_t hps_beam_get_size(const hps_beam* b) { return b->ng; }

int hps_beam_shift(hps_beam* b, int coord, double v) {  // beam.cpp:1595-1642 + beam_kernel.cu:1489-1496
  std::vector<double>* a[6] = {&b->x, &b->xp, &b->y, &b->yp, &b->phi, &b->w};
  if (coord < 0 || coord > 5) return fail(b->ctx, HPS_ERR_ARG, "bad coordinate");
  for (uint i = 0; i < b->n; ++i) if (b->gi(i) != 0) (*a[coord])[i] += v;
  if (coord == HPS_PHI) InitPhiAvgGood(b);
  return HPS_OK;
}
int hps_beam_cut(hps_beam* b, int coord, double lo, double hi) {  // beam.cpp:1574-1588 + beam_kernel.cu:1464-1480
  std::vector<double>* a;
  switch (coord) {
    case HPS_X: a = &b->x; break; case HPS_Y: a = &b->y; break;
    case HPS_W: a = &b->w; break; case HPS_PHI: a = &b->phi; break;
    default: return fail(b->ctx, HPS_ERR_ARG, "cut supports x, y, w, phi");
  }
  for (uint i = 0; i < b->n; ++i)
    if (b->gi(i) != 0 && b->loss[i] == 0 && (*a)[i] < hi && (*a)[i] > lo) b->loss[i] = 99999;
  return HPS_OK;
}

int hps_beam_update(hps_beam* b, int id) {
  switch (id) {
    case HPS_UPD_LOSS: UpdateLoss(b); break;
    case HPS_UPD_LLOSS: UpdateLongitudinalLoss(b); break;
    case HPS_UPD_GOOD_COUNT: UpdateGoodParticleCount(b); break;
    case HPS_UPD_AVG_X: UpdateAvg(b, b->x, HPS_STAT_X_AVG, false); break;
    case HPS_UPD_AVG_X_GOOD: UpdateAvg(b, b->x, HPS_STAT_X_AVG_GOOD, true); break;
    case HPS_UPD_AVG_XP: UpdateAvg(b, b->xp, HPS_STAT_XP_AVG, false); break;
    case HPS_UPD_AVG_Y: UpdateAvg(b, b->y, HPS_STAT_Y_AVG, false); break;
    case HPS_UPD_AVG_Y_GOOD: UpdateAvg(b, b->y, HPS_STAT_Y_AVG_GOOD, true); break;
    case HPS_UPD_AVG_YP: UpdateAvg(b, b->yp, HPS_STAT_YP_AVG, false); break;
    case HPS_UPD_AVG_PHI: UpdateAvg(b, b->phi, HPS_STAT_PHI_AVG, false); break;
    case HPS_UPD_AVG_PHI_GOOD: UpdateAvg(b, b->phi, HPS_STAT_PHI_AVG_GOOD, true); break;
    case HPS_UPD_AVG_PHI_R: UpdateAvg(b, b->phi_r, HPS_STAT_PHI_AVG_R, false); break;
    case HPS_UPD_AVG_PHI_R_GOOD: UpdateAvg(b, b->phi_r, HPS_STAT_PHI_AVG_R, true); break;
    case HPS_UPD_AVG_W: UpdateAvg(b, b->w, HPS_STAT_W_AVG, false); break;
    case HPS_UPD_AVG_W_GOOD: UpdateAvg(b, b->w, HPS_STAT_W_AVG_GOOD, true); break;
    case HPS_UPD_SIG_X: UpdateSig(b, b->x, HPS_STAT_X_SIG, false); break;
    case HPS_UPD_SIG_X_GOOD: UpdateSig(b, b->x, HPS_STAT_X_SIG_GOOD, true); break;
    case HPS_UPD_SIG_XP: UpdateSig(b, b->xp, HPS_STAT_XP_SIG, false); break;
    case HPS_UPD_SIG_Y: UpdateSig(b, b->y, HPS_STAT_Y_SIG, false); break;
    case HPS_UPD_SIG_Y_GOOD: UpdateSig(b, b->y, HPS_STAT_Y_SIG_GOOD, true); break;
    case HPS_UPD_SIG_YP: UpdateSig(b, b->yp, HPS_STAT_YP_SIG, false); break;
    case HPS_UPD_SIG_PHI: UpdateSig(b, b->phi, HPS_STAT_PHI_SIG, false); break;
    case HPS_UPD_SIG_PHI_R: UpdateSig(b, b->phi_r, HPS_STAT_PHI_SIG_R, false); break;
    case HPS_UPD_SIG_PHI_R_GOOD: UpdateSig(b, b->phi_r, HPS_STAT_PHI_SIG_GOOD, true); break;
    case HPS_UPD_SIG_W: UpdateSig(b, b->w, HPS_STAT_W_SIG, false); break;
    case HPS_UPD_SIG_W_GOOD: UpdateSig(b, b->w, HPS_STAT_W_SIG_GOOD, true); break;
    case HPS_UPD_EMITTANCE: UpdateEmittance(b); break;
    case HPS_UPD_STAT_FOR_SPCH: UpdateStatForSpaceCharge(b); break;
    case HPS_UPD_AVG_XY: UpdateAvgXY(b); break;
    case HPS_UPD_MAX_R_PHI: UpdateMaxRPhi(b); break;
    case HPS_UPD_REL_PHI: UpdateRelativePhi(b, false); break;
    case HPS_UPD_REL_PHI_GOOD: UpdateRelativePhi(b, true); break;
    default: return fail(b->ctx, HPS_ERR_ARG, "unknown update id");
  }
  return HPS_OK;
}
int hps_beam_get(hps_beam* b, int id, double* out) {
  if (id < 0 || id >= HPS_STAT_COUNT) return fail(b->ctx, HPS_ERR_ARG, "unknown stat id");
  if (id == HPS_STAT_NUM_LOSS) *out = b->s.num_loss;
  else if (id == HPS_STAT_NUM_LLOSS) *out = b->s.num_lloss;
  else if (id == HPS_STAT_NUM_GOOD) *out = b->s.num_good;
  else *out = b->s.v[id];
  return HPS_OK;
}

hps_lattice* hps_lattice_create(hps_ctx* c, const hps_elem* e, uint32_t n) {
  hps_lattice* l = new hps_lattice();
  l->ctx = c; l->e.assign(e, e + n);
  return l;
}
void hps_lattice_destroy(hps_lattice* l) { delete l; }
uint32_t hps_lattice_size(const hps_lattice* l) { return (uint32_t)l->e.size(); }
int64_t hps_lattice_index(const hps_lattice* l, const char* name) {
  for (size_t i = 0; i < l->e.size(); ++i)
    if (strcasecmp(l->e[i].name, name) == 0) return (int64_t)i;
  return -1;
}
int hps_lattice_get_elem(const hps_lattice* l, uint32_t i, hps_elem* out) {
  if (i >= l->e.size()) return HPS_ERR_ARG;
  *out = l->e[i];
  return HPS_OK;
}
int hps_lattice_update_elem(hps_lattice* l, uint32_t i, const hps_elem* in) {
  if (i >= l->e.size()) return HPS_ERR_ARG;
  l->e[i] = *in;
  return HPS_OK;
}

hps_scheff* hps_scheff_create(hps_ctx* c, uint32_t nr, uint32_t nz, int adj_bunch) {
  hps_scheff* s = new hps_scheff();
  s->ctx = c; s->nx = s->nx_init = nr; s->nz = s->nz_init = nz; s->adj_bunch = (uint)adj_bunch;
  AllocTables(s);
  return s;
}
void hps_scheff_destroy(hps_scheff* s) { delete s; }
int hps_scheff_set(hps_scheff* s, int p, double v) {
  switch (p) {
    case HPS_SC_INTERVAL: s->interval = v; s->interval_init = v; break;
    case HPS_SC_ADJ_BUNCH: s->adj_bunch = (uint)(int)v; break;
    case HPS_SC_ADJ_BUNCH_CUTOFF_W: s->adj_bunch_w = v; break;
    case HPS_SC_MESH_CUTOFF_W: s->mesh_w = v; break;
    case HPS_SC_REMESH_THRESHOLD: s->remesh_val = v; break;
    case HPS_SC_FRACTION: s->fraction = v; break;
    default: return fail(s->ctx, HPS_ERR_ARG, "parameter is read-only or unknown");
  }
  return HPS_OK;
}
int hps_scheff_get(hps_scheff* s, int p, double* v) {
  switch (p) {
    case HPS_SC_INTERVAL: *v = s->interval; break;
    case HPS_SC_ADJ_BUNCH: *v = s->adj_bunch; break;
    case HPS_SC_ADJ_BUNCH_CUTOFF_W: *v = s->adj_bunch_w; break;
    case HPS_SC_MESH_CUTOFF_W: *v = s->mesh_w; break;
    case HPS_SC_REMESH_THRESHOLD: *v = s->remesh_val; break;
    case HPS_SC_FRACTION: *v = s->fraction; break;
    case HPS_SC_NR: *v = s->nx; break;
    case HPS_SC_NZ: *v = s->nz; break;
    default: return fail(s->ctx, HPS_ERR_ARG, "unknown parameter");
  }
  return HPS_OK;
}
int hps_scheff_set_mesh_size(hps_scheff* s, uint32_t nr, uint32_t nz) {
  // wrap_spacecharge.cpp set_mesh_size -> Scheff::SetMeshSize (nx_init_/nz_init_ keep the constructor values)
  SetMeshSize(s, nr, nz);
  return HPS_OK;
}
hps_monitor* hps_monitor_create(hps_ctx* c, uint32_t capacity) {
  if (!c || capacity == 0) return nullptr;
  hps_monitor* m = new hps_monitor();
  m->ctx = c; m->cap = capacity;
  m->data.assign((size_t)HPS_MON_NFIELD * capacity, 0.0);
  return m;
}
void hps_monitor_destroy(hps_monitor* m) {
  if (!m) return;
  if (m->ctx->monitor == m) m->ctx->monitor = nullptr;
  delete m;
}
int hps_ctx_set_monitor(hps_ctx* c, hps_monitor* m) { c->monitor = m; return HPS_OK; }
int hps_monitor_download(hps_monitor* m, double* out, uint32_t* count) {
  if (out) std::memcpy(out, m->data.data(), m->data.size() * sizeof(double));
  if (count) *count = m->monitor_num;
  return HPS_OK;
}
int hps_beam_moment_stats(hps_beam*, uint64_t out[18], double* e, double* m) {  // the oracle runs the reference's chain as it is
  if (out) for (int k = 0; k < 18; ++k) out[k] = 0;
  if (e) *e = 0.0;
  if (m) *m = 0.0;
  return HPS_OK;
}
int hps_scheff_phase_cycles(hps_scheff*, uint64_t out[8]) { for (int i = 0; i < 8; ++i) out[i] = 0; return HPS_OK; }
int hps_scheff_download_tables(hps_scheff* s, double* bins, double* field, double* dr_dz) {
  if (bins) std::copy(s->bins.begin(), s->bins.end(), bins);
  if (field) std::copy(s->field.begin(), s->field.end(), field);
  if (dr_dz) { dr_dz[0] = s->dr; dr_dz[1] = s->dz; }
  return HPS_OK;
}

int hps_engine_reset(hps_ctx* c, hps_beam* b) { Reset(c, b); return HPS_OK; }

int hps_simulate(hps_ctx* c, hps_beam* b, hps_lattice* lat, hps_scheff* sc, uint32_t first, uint32_t last,
                 hps_run_stats* out) {
  if (!c || !b || !lat) return fail(c, HPS_ERR_ARG, "null handle");
  if (last >= lat->e.size()) return fail(c, HPS_ERR_ARG, "element index out of range");
  // simulation_engine.cpp:73-95
  if (first < c->prev_end) Reset(c, b);
  c->prev_end = last;
  uint64_t k0 = sc ? sc->kicks : 0, r0 = sc ? sc->remesh : 0;
  auto t0 = std::chrono::steady_clock::now();
  for (uint i = 0; i <= last; ++i)
    if (i >= first || lat->e[i].type == HPS_ELEM_SPCH_COMP) {
      int rc = VisitElement(c, b, lat, sc, i);
      if (rc != HPS_OK) return rc;
    }
  if (c->monitor) {  // simulation_engine.cpp:93-94 -> Beam::UpdateStatForPlotting (beam.cpp:1024-1029)
    UpdateStatForSpaceCharge(b);
    UpdateSig(b, b->w, HPS_STAT_W_SIG_GOOD, true);
    UpdateAvg(b, b->phi_r, HPS_STAT_PHI_AVG_R, true);
  }
  auto t1 = std::chrono::steady_clock::now();
  if (out) {
    std::memset(out, 0, sizeof(*out));
    out->particles = b->ng;
    out->elements = last >= first ? last - first + 1 : 0;
    out->kicks = sc ? sc->kicks - k0 : 0;
    out->remesh = sc ? sc->remesh - r0 : 0;
    out->device_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  }
  return HPS_OK;
}

}  // extern "C"
