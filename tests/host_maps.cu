// host_maps.cu -- TEST INFRASTRUCTURE (no GPU needed).
//
// The per-particle element maps of hpsim_b200/csrc/hps_kernels.cuh are plain fp64 arithmetic and compile for the
// host.  This program runs every op record the planner produces for a lattice range through BOTH forms of each
// map -- the reference's expression order (MATH = 0; restates simulation_kernel.cu:16-871 and
// space_charge_kernel.cu:573-706) and the fast form the product launches by default (MATH = 1: reciprocals of
// block-uniform quantities, rsqrt for sqrt-then-divide) -- on random particles around the design energy, and
// reports the largest deviation.  tests/test_host_maps.py builds it with nvcc (host code only) and asserts
// that the deviation stays at rounding level and that the loss flags agree.
//
//   host_maps <lattice.lat> <first index> <last index> <particles per op> <seed>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../hpsim_b200/csrc/hps_kernels.cuh"
#include "../hpsim_b200/csrc/hps_plan.h"

using namespace hps;

static bool read_lattice(const char* path, std::vector<hps_elem>& out) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  char magic[4];
  uint32_t ver = 0, n = 0;
  bool ok = fread(magic, 1, 4, f) == 4 && fread(&ver, 4, 1, f) == 1 && fread(&n, 4, 1, f) == 1 && memcmp(magic, "HPSL", 4) == 0;
  if (ok) {
    out.resize(n);
    ok = fread(out.data(), sizeof(hps_elem), n, f) == n;
  }
  fclose(f);
  return ok;
}

struct Dev {
  double x = 0, xp = 0, y = 0, yp = 0, phi = 0, w = 0;
  long loss_mismatch = 0, n = 0;
  void add(const Pt& a, const Pt& b) {
    ++n;
    if (a.loss != b.loss) { ++loss_mismatch; return; }
    if (a.loss != 0) return;
    auto rel = [](double u, double v, double floor_) { return fabs(u - v) / fmax(fabs(v), floor_); };
    x = fmax(x, rel(a.x, b.x, 1e-3));
    xp = fmax(xp, rel(a.xp, b.xp, 1e-3));
    y = fmax(y, rel(a.y, b.y, 1e-3));
    yp = fmax(yp, rel(a.yp, b.yp, 1e-3));
    phi = fmax(phi, rel(a.phi, b.phi, 1.0));
    w = fmax(w, rel(a.w, b.w, 1e-3));
  }
  double worst() const { return fmax(fmax(fmax(x, xp), fmax(y, yp)), fmax(phi, w)); }
};

// whole-range tracking (space charge off) of a beam file with either form of the maps
static bool read_beam(const char* path, std::vector<Pt>& out) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  char magic[4]; uint32_t ver; unsigned long long n;
  if (fread(magic, 1, 4, f) != 4 || fread(&ver, 4, 1, f) != 1 || fread(&n, 8, 1, f) != 1) return false;
  out.resize(n);
  std::vector<double> a(n);
  double Pt::*fld[6] = {&Pt::x, &Pt::xp, &Pt::y, &Pt::yp, &Pt::phi, &Pt::w};
  for (int k = 0; k < 6; ++k) {
    if (fread(a.data(), 8, n, f) != n) return false;
    for (size_t i = 0; i < n; ++i) out[i].*fld[k] = a[i];
  }
  std::vector<uint32_t> l(n);
  if (fread(l.data(), 4, n, f) != n) return false;
  for (size_t i = 0; i < n; ++i) out[i].loss = l[i];
  fclose(f);
  return true;
}

static int track_main(int argc, char** argv) {
  std::vector<hps_elem> lat;
  std::vector<Pt> beam0;
  if (argc < 6 || !read_lattice(argv[2], lat) || !read_beam(argv[5], beam0)) { fprintf(stderr, "track: bad arguments\n"); return 2; }
  const u32 first = (u32)atoi(argv[3]), last = (u32)atoi(argv[4]);
  const double mass = 939.294, charge = -1.0, inv_mass = 1.0 / mass;
  EngineState eng;
  BeamHost bh = {mass, charge, 0.015, 201.25, 0.75, (u32)beam0.size()};
  ScHost sc;
  Plan plan = build_plan(lat, first, last, eng, bh, sc);
  if (!plan.error.empty()) { fprintf(stderr, "plan: %s\n", plan.error.c_str()); return 2; }
  std::vector<OpC> opc(plan.ops.size());
  for (size_t k = 0; k < plan.ops.size(); ++k) { memset(&opc[k], 0, sizeof(OpC)); prepare_op(plan.ops[k], opc[k], mass, charge); }
  std::vector<Pt> A = beam0, B = beam0;
  for (int variant = 0; variant < 2; ++variant) {
    std::vector<Pt>& P = variant ? B : A;
    for (size_t i = 0; i < P.size(); ++i) {
      Pt p = P[i];
      for (size_t k = 0; k < plan.ops.size() && p.loss == 0; ++k) {
        const Op& op = plan.ops[k];
        const double* q = op.p; const double* c = opc[k].c;
        switch (op.type) {
          case OP_DRIFT: if (variant) op_drift_f(p, q, c, op.elem, inv_mass); else op_drift<MathLib>(p, q, c, op.elem, mass, inv_mass); break;
          case OP_QUAD_HALF: if (variant) op_quad_half_f(p, q, c, op.elem, inv_mass); else op_quad_half<MathLib>(p, q, c, op.elem, mass, inv_mass); break;
          case OP_GAP_FIRST: if (variant) op_gap_first_f(p, q, c, op.flag != 0, inv_mass); else op_gap_first<MathLib>(p, q, c, op.flag != 0, mass, inv_mass); break;
          case OP_GAP_SECOND: if (variant) op_gap_second_f(p, q, c, op.flag != 0, inv_mass); else op_gap_second<MathLib>(p, q, c, op.flag != 0, mass, inv_mass); break;
          case OP_DIPOLE_FIRST: op_dipole(p, q, false, mass); break;
          case OP_DIPOLE_SECOND: op_dipole(p, q, true, mass); break;
          case OP_BUNCHER: op_buncher(p, q, op.elem, mass, charge); break;
          case OP_ROTATION: {
            const double cosa = c[R_COS], sina = c[R_SIN];
            const double x = p.x, y = p.y, xp = p.xp, yp = p.yp;
            p.x = x * cosa + y * sina; p.y = y * cosa - x * sina; p.xp = xp * cosa + yp * sina; p.yp = yp * cosa - xp * sina;
            break;
          }
          case OP_STEERER: {
            const double gm1 = p.w / mass;
            const double pm = sqrt(gm1 * (gm1 + 2.0)) * mass / K_CLIGHT;
            if (q[0] > 1e-10) p.xp += q[0] * charge / pm;
            if (q[1] > 1e-10) p.yp += q[1] * charge / pm;
            break;
          }
          case OP_CAPERTURE: if (i > 0 && p.x * p.x + p.y * p.y > q[0] * q[0]) p.loss = op.elem; break;
          case OP_FREQ_CHANGE: p.phi *= q[0]; break;
          default: break;
        }
      }
      P[i] = p;
    }
  }
  // deviation of B (fast) from A (reference order), relative to the rms extent of each coordinate
  double Pt::*fld[6] = {&Pt::x, &Pt::xp, &Pt::y, &Pt::yp, &Pt::phi, &Pt::w};
  const char* nm[6] = {"x", "xp", "y", "yp", "phi", "w"};
  long alive = 0, mism = 0;
  for (size_t i = 0; i < A.size(); ++i) { if (A[i].loss != B[i].loss) ++mism; else if (A[i].loss == 0) ++alive; }
  printf("TRACK particles %zu alive %ld loss_mismatch %ld ops %zu\n", A.size(), alive, mism, plan.ops.size());
  std::vector<double> phis;
  for (auto& p : A) if (p.loss == 0) phis.push_back(p.phi);
  std::sort(phis.begin(), phis.end());
  const double phi_med = phis[phis.size() / 2], phi_mean = phi_med;
  for (int k = 0; k < 6; ++k) {
    double m = 0, m2 = 0;
    for (size_t i = 0; i < A.size(); ++i) if (A[i].loss == 0 && B[i].loss == 0) { m += A[i].*fld[k]; m2 += (A[i].*fld[k]) * (A[i].*fld[k]); }
    m /= alive; const double sd = sqrt(fmax(m2 / alive - m * m, 1e-300));
    std::vector<double> e;
    long big = 0, big_core = 0, core = 0;
    for (size_t i = 0; i < A.size(); ++i) if (A[i].loss == 0 && B[i].loss == 0) {
      const double d = fabs(A[i].*fld[k] - B[i].*fld[k]) / sd;
      e.push_back(d);
      const bool in_core = fabs(A[i].phi - phi_mean) < 1.0;
      core += in_core;
      if (d > 1e-9) { ++big; big_core += in_core; if (k == 0) printf("   outlier i=%zu dphi=%.3f w=%.6g (design %.4g) dev=%.2e\n", i, A[i].phi - phi_med, A[i].w, bh.design_w, d); }
    }
    std::sort(e.begin(), e.end());
    printf("%-3s sd %.3e median %.2e p99 %.2e max %.2e  >1e-9: %ld (in core |dphi|<1: %ld of %ld)\n", nm[k], sd, e[e.size() / 2], e[e.size() * 99 / 100], e.back(), big, big_core, core);
  }
  return 0;
}

// planner only: host_maps plan <lattice> <first> <last> <interval> -- counts of the program hps_simulate would launch
static int plan_main(int argc, char** argv) {
  std::vector<hps_elem> lat;
  if (argc < 6 || !read_lattice(argv[2], lat)) { fprintf(stderr, "plan: bad arguments\n"); return 2; }
  EngineState eng;
  BeamHost bh = {939.294, -1.0, 0.015, 201.25, 0.75, 1u << 16};
  ScHost sc;
  sc.on = true; sc.nr = 32; sc.nz = 128; sc.adj_bunch = 3; sc.interval = atof(argv[5]); sc.adj_bunch_w = 0.8; sc.remesh_val = 0.02;
  Plan plan = build_plan(lat, (u32)atoi(argv[3]), (u32)atoi(argv[4]), eng, bh, sc);
  if (!plan.error.empty()) { fprintf(stderr, "plan: %s\n", plan.error.c_str()); return 2; }
  size_t kick_segs = 0, max_ops = 0, fin_aper = 0;
  for (const Segment& s : plan.segs) {
    kick_segs += s.kick_follows;
    max_ops = s.op_count > max_ops ? s.op_count : max_ops;
    fin_aper += (s.fin_mode & 3) == FIN_APERTURE;
  }
  size_t nkick_ops = 0, ngap = 0;
  for (const Op& op : plan.ops) { nkick_ops += op.type == OP_SC_KICK; ngap += op.type == OP_GAP_FIRST; }
  if (getenv("HPS_PLAN_PATTERNS")) {  // histogram of the op-type sequences of the segments (which fused sweeps dominate)
    std::vector<std::pair<std::string, size_t>> hist;
    for (const Segment& s : plan.segs) {
      std::string key;
      for (u32 k = 0; k < s.op_count; ++k) {
        const Op& op = plan.ops[s.op_begin + k];
        static const char* nm[] = {"none", "drift", "quad", "gap1", "gap2", "dip1", "dip2", "bunch", "rot", "steer", "capert", "rapert", "KICK", "freq"};
        key += nm[op.type];
        if (op.type == OP_GAP_FIRST || op.type == OP_GAP_SECOND) key += op.flag ? "c" : "d";
        key += " ";
      }
      key += s.kick_follows ? "| kick" : ((s.fin_mode & 3) == FIN_APERTURE ? "| aperture" : "| -");
      bool found = false;
      for (auto& h : hist) if (h.first == key) { h.second++; found = true; break; }
      if (!found) hist.push_back({key, 1});
    }
    std::sort(hist.begin(), hist.end(), [](const std::pair<std::string, size_t>& a, const std::pair<std::string, size_t>& b) { return a.second > b.second; });
    for (auto& h : hist) printf("PATTERN %6zu  %s\n", h.second, h.first.c_str());
  }
  printf("PLAN ops %zu segments %zu kicks %u kick_segments %zu kick_ops %zu gaps %zu aperture_syncs %zu max_ops_per_segment %zu "
         "freq %.4f design_w %.6f\n", plan.ops.size(), plan.segs.size(), plan.n_kicks, kick_segs, nkick_ops, ngap, fin_aper, max_ops, bh.freq, bh.design_w);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && strcmp(argv[1], "track") == 0) return track_main(argc, argv);
  if (argc > 1 && strcmp(argv[1], "plan") == 0) return plan_main(argc, argv);
  if (argc < 6) { fprintf(stderr, "usage: host_maps lattice first last n seed\n"); return 2; }
  std::vector<hps_elem> lat;
  if (!read_lattice(argv[1], lat)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  const u32 first = (u32)atoi(argv[2]), last = (u32)atoi(argv[3]);
  const int npart = atoi(argv[4]);
  std::mt19937_64 rng((unsigned long long)atoll(argv[5]));
  std::uniform_real_distribution<double> U(-1.0, 1.0);

  const double mass = 939.294, charge = -1.0;
  EngineState eng;
  BeamHost beam = {mass, charge, 0.015, 201.25, 0.75, 1u << 20};
  ScHost sc;  // off: the plan holds the element maps only
  Plan plan = build_plan(lat, first, last, eng, beam, sc);
  if (!plan.error.empty()) { fprintf(stderr, "plan: %s\n", plan.error.c_str()); return 2; }

  Dev dev[16];
  long nops[16] = {0};
  double design_w = 0.75;
  const double inv_mass = 1.0 / mass;
  for (const Op& op : plan.ops) {
    OpC oc;
    memset(&oc, 0, sizeof(oc));
    prepare_op(op, oc, mass, charge);
    if (op.type == OP_GAP_FIRST || op.type == OP_GAP_SECOND) design_w = op.p[G_DESIGN_W];
    nops[op.type]++;
    for (int i = 0; i < npart; ++i) {
      Pt a;
      a.x = 4e-3 * U(rng); a.y = 4e-3 * U(rng); a.xp = 8e-3 * U(rng); a.yp = 8e-3 * U(rng);
      a.phi = K_PI * U(rng) + K_TWOPI * floor(2000.0 * U(rng));
      a.w = design_w * (1.0 + 0.01 * U(rng));
      a.loss = 0;
      if (i == 0) { a.x = a.y = a.xp = a.yp = 0.0; }  // the reference particle sits on the axis
      Pt b = a;
      switch (op.type) {
        case OP_DRIFT: op_drift<MathLib>(a, op.p, oc.c, op.elem, mass, inv_mass); op_drift_f(b, op.p, oc.c, op.elem, inv_mass); break;
        case OP_QUAD_HALF: op_quad_half<MathLib>(a, op.p, oc.c, op.elem, mass, inv_mass); op_quad_half_f(b, op.p, oc.c, op.elem, inv_mass); break;
        case OP_GAP_FIRST: op_gap_first<MathLib>(a, op.p, oc.c, op.flag != 0, mass, inv_mass); op_gap_first_f(b, op.p, oc.c, op.flag != 0, inv_mass); break;
        case OP_GAP_SECOND: op_gap_second<MathLib>(a, op.p, oc.c, op.flag != 0, mass, inv_mass); op_gap_second_f(b, op.p, oc.c, op.flag != 0, inv_mass); break;
        default: continue;
      }
      dev[op.type].add(b, a);
    }
    if (op.type == OP_GAP_SECOND) design_w = op.p[G_EOUT];
  }

  // space-charge kick: synthetic but representative block-uniform context, smooth field table
  {
    const u32 nr = 32, nz = 128;
    std::vector<double2> fld((nr + 1) * (nz + 1));
    for (u32 j = 0; j <= nz; ++j)
      for (u32 i = 0; i <= nr; ++i) fld[j * (nr + 1) + i] = make_double2(1e-3 * sin(0.2 * i) * cos(0.05 * j), 1e-3 * cos(0.1 * i) * sin(0.07 * j));
    for (int rep = 0; rep < 64; ++rep) {
      const double wavg = rep % 2 ? 0.75 : 40.0 + 700.0 * fabs(U(rng));
      KickCtx k;
      memset(&k, 0, sizeof(k));
      const double gm = wavg / mass + 1.0;
      k.beta = sqrt(1.0 - 1.0 / (gm * gm));
      k.phimc = 0.01 * U(rng); k.xbar = 1e-4 * U(rng); k.ybar = 1e-4 * U(rng); k.epsq = 1.0 + 0.5 * U(rng);
      k.phi_ref = K_TWOPI * floor(1000.0 * U(rng)) + 0.3 * U(rng);
      k.dr = 3.0 * 2e-3 / nr; k.dz = k.beta * K_CLIGHT / 201.25 / nz * (rep % 2 ? 1.0 : 0.2);
      k.ratio_r = 1.0 + 0.01 * U(rng); k.ratio_z = 1.0 + 0.01 * U(rng); k.ratio_q = 1.0; k.ratio_g = 1.0 + 0.001 * U(rng);
      k.adj_bunch = rep % 2 ? 3 : 0; k.num_on_mesh = 900000; k.nr = nr; k.nz = nz;
      const double q[3] = {0.025, 0.015, 201.25};
      kick_prepare(k, q, mass, 1u << 20);
      for (int i = 0; i < npart * 8; ++i) {
        Pt a;
        const double spread = (i % 5 == 0) ? 5.0 : 1.0;  // every 5th particle far out: the off-mesh branch
        a.x = spread * 2e-3 * U(rng); a.y = spread * 2e-3 * U(rng); a.xp = 8e-3 * U(rng); a.yp = 8e-3 * U(rng);
        a.phi = k.phi_ref + (rep % 2 ? K_PI : 0.3) * spread * U(rng);
        a.w = wavg * (1.0 + 0.01 * U(rng));
        a.loss = 0;
        Pt b = a;
        op_sc_kick<MathLib>(a, 0u, k, fld.data(), mass, inv_mass);
        op_sc_kick_f(b, 0u, k, fld.data(), inv_mass);
        dev[OP_SC_KICK].add(b, a);
      }
    }
    nops[OP_SC_KICK] = 64;
  }

  const char* names[16] = {"none", "drift", "quad_half", "gap_first", "gap_second", "dipole_first", "dipole_second", "buncher",
                           "rotation", "steerer", "caperture", "raperture", "sc_kick", "freq_change", "", ""};
  double worst = 0.0;
  long mism = 0;
  for (int t = 0; t < 16; ++t) {
    if (dev[t].n == 0) continue;
    printf("%-12s ops %6ld particles %8ld loss_mismatch %ld  x %.2e xp %.2e y %.2e yp %.2e phi %.2e w %.2e\n", names[t], nops[t],
           dev[t].n, dev[t].loss_mismatch, dev[t].x, dev[t].xp, dev[t].y, dev[t].yp, dev[t].phi, dev[t].w);
    worst = fmax(worst, dev[t].worst());
    mism += dev[t].loss_mismatch;
  }
  printf("WORST %.3e LOSS_MISMATCH %ld\n", worst, mism);
  return 0;
}
