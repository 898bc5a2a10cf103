// gpu_maps_check.cu -- TEST INFRASTRUCTURE (-m gpu).  Runs every op record of a lattice range through the
// reference-order maps twice on the GPU -- with nvcc's own division / sqrt (MathLib) and with the branch-free
// exact sequences of hps_math.cuh (MathNR) -- on pseudo-random particles around the design energy, and counts
// the outputs that are not bit-identical.
//   gpu_maps_check <lattice.lat> <first> <last> <particles per op>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../hpsim_b200/csrc/hps_kernels.cuh"
#include "../hpsim_b200/csrc/hps_plan.h"

using namespace hps;

__device__ __forceinline__ unsigned long long mix(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double uni(unsigned long long r) { return (double)(r >> 11) * (2.0 / 9007199254740992.0) - 1.0; }

__device__ bool same(const Pt& a, const Pt& b) {
  return __double_as_longlong(a.x) == __double_as_longlong(b.x) && __double_as_longlong(a.xp) == __double_as_longlong(b.xp) &&
         __double_as_longlong(a.y) == __double_as_longlong(b.y) && __double_as_longlong(a.yp) == __double_as_longlong(b.yp) &&
         __double_as_longlong(a.phi) == __double_as_longlong(b.phi) && __double_as_longlong(a.w) == __double_as_longlong(b.w) && a.loss == b.loss;
}

__global__ void check(const Op* ops, const OpC* opcs, const double* design_w, int nops, int npart, double mass, unsigned long long* bad,
                      unsigned long long* count) {
  const double inv_mass = 1.0 / mass;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)nops * npart; t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t / npart), i = (int)(t % npart);
    const Op& op = ops[k];
    const double* q = op.p; const double* c = opcs[k].c;
    const unsigned long long s = mix(t * 8 + 1);
    Pt a;
    a.x = 4e-3 * uni(mix(s)); a.y = 4e-3 * uni(mix(s + 1)); a.xp = 8e-3 * uni(mix(s + 2)); a.yp = 8e-3 * uni(mix(s + 3));
    a.phi = K_PI * uni(mix(s + 4)) + K_TWOPI * floor(2000.0 * uni(mix(s + 5)));
    a.w = design_w[k] * (1.0 + 0.01 * uni(mix(s + 6)));
    a.loss = 0;
    if (i == 0) { a.x = a.y = a.xp = a.yp = 0.0; }
    Pt b = a;
    switch (op.type) {
      case OP_DRIFT: op_drift<MathLib>(a, q, c, op.elem, mass, inv_mass); op_drift<MathNR>(b, q, c, op.elem, mass, inv_mass); break;
      case OP_QUAD_HALF: op_quad_half<MathLib>(a, q, c, op.elem, mass, inv_mass); op_quad_half<MathNR>(b, q, c, op.elem, mass, inv_mass); break;
      case OP_GAP_FIRST: op_gap_first<MathLib>(a, q, c, op.flag != 0, mass, inv_mass); op_gap_first<MathNR>(b, q, c, op.flag != 0, mass, inv_mass); break;
      case OP_GAP_SECOND: op_gap_second<MathLib>(a, q, c, op.flag != 0, mass, inv_mass); op_gap_second<MathNR>(b, q, c, op.flag != 0, mass, inv_mass); break;
      default: continue;
    }
    atomicAdd(&count[op.type], 1ull);
    if (!same(a, b)) atomicAdd(&bad[op.type], 1ull);
  }
}

__global__ void prep(const Op* ops, OpC* opcs, int nops, double mass, double charge) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nops) prepare_op(ops[i], opcs[i], mass, charge);
}

// the space-charge kick with a synthetic, representative context
__global__ void check_kick(const double2* fld, int n, double mass, unsigned long long* bad, unsigned long long* count) {
  const double inv_mass = 1.0 / mass;
  __shared__ KickCtx k;
  const int rep = blockIdx.x;
  if (threadIdx.x == 0) {
    const unsigned long long s = mix(rep * 977 + 5);
    const double wavg = rep % 2 ? 0.75 : 40.0 + 700.0 * fabs(uni(mix(s)));
    memset(&k, 0, sizeof(k));
    const double gm = wavg / mass + 1.0;
    k.beta = sqrt(1.0 - 1.0 / (gm * gm));
    k.phimc = 0.01 * uni(mix(s + 1)); k.xbar = 1e-4 * uni(mix(s + 2)); k.ybar = 1e-4 * uni(mix(s + 3)); k.epsq = 1.0 + 0.5 * uni(mix(s + 4));
    k.phi_ref = K_TWOPI * floor(1000.0 * uni(mix(s + 5))) + 0.3 * uni(mix(s + 6));
    k.dr = 3.0 * 2e-3 / 32; k.dz = k.beta * K_CLIGHT / 201.25 / 128 * (rep % 2 ? 1.0 : 0.2);
    k.ratio_r = 1.0 + 0.01 * uni(mix(s + 7)); k.ratio_z = 1.0 + 0.01 * uni(mix(s + 8)); k.ratio_q = 1.0; k.ratio_g = 1.0 + 0.001 * uni(mix(s + 9));
    k.adj_bunch = rep % 2 ? 3 : 0; k.num_on_mesh = 900000; k.nr = 32; k.nz = 128;
    k.wavg_dbg = wavg;
    const double q[3] = {0.025, 0.015, 201.25};
    kick_prepare(k, q, mass, 1u << 20);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long s = mix((unsigned long long)rep * 1000003ull + i);
    const double spread = (i % 5 == 0) ? 5.0 : 1.0;
    Pt a;
    a.x = spread * 2e-3 * uni(mix(s)); a.y = spread * 2e-3 * uni(mix(s + 1)); a.xp = 8e-3 * uni(mix(s + 2)); a.yp = 8e-3 * uni(mix(s + 3));
    a.phi = k.phi_ref + (rep % 2 ? K_PI : 0.3) * spread * uni(mix(s + 4));
    a.w = k.wavg_dbg * (1.0 + 0.01 * uni(mix(s + 5)));
    a.loss = 0;
    Pt b = a;
    op_sc_kick<MathLib>(a, 0u, k, fld, mass, inv_mass);
    op_sc_kick<MathNR>(b, 0u, k, fld, mass, inv_mass);
    atomicAdd(&count[OP_SC_KICK], 1ull);
    if (!same(a, b)) atomicAdd(&bad[OP_SC_KICK], 1ull);
  }
}

static bool read_lattice(const char* path, std::vector<hps_elem>& out) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  char magic[4];
  uint32_t ver = 0, n = 0;
  bool ok = fread(magic, 1, 4, f) == 4 && fread(&ver, 4, 1, f) == 1 && fread(&n, 4, 1, f) == 1 && memcmp(magic, "HPSL", 4) == 0;
  if (ok) { out.resize(n); ok = fread(out.data(), sizeof(hps_elem), n, f) == n; }
  fclose(f);
  return ok;
}

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: gpu_maps_check lattice first last n\n"); return 2; }
  std::vector<hps_elem> lat;
  if (!read_lattice(argv[1], lat)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  const double mass = 939.294, charge = -1.0;
  EngineState eng;
  BeamHost beam = {mass, charge, 0.015, 201.25, 0.75, 1u << 20};
  ScHost sc;
  Plan plan = build_plan(lat, (u32)atoi(argv[2]), (u32)atoi(argv[3]), eng, beam, sc);
  if (!plan.error.empty()) { fprintf(stderr, "plan: %s\n", plan.error.c_str()); return 2; }
  const int nops = (int)plan.ops.size(), npart = atoi(argv[4]);
  std::vector<double> dw(nops);
  double design_w = 0.75;
  for (int k = 0; k < nops; ++k) {
    const Op& op = plan.ops[k];
    if (op.type == OP_GAP_FIRST || op.type == OP_GAP_SECOND) design_w = op.p[G_DESIGN_W];
    dw[k] = design_w;
    if (op.type == OP_GAP_SECOND) design_w = op.p[G_EOUT];
  }
  Op* d_ops; OpC* d_opc; double* d_dw; unsigned long long *d_bad, *d_cnt; double2* d_fld;
  if (cudaMalloc(&d_ops, nops * sizeof(Op)) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 2; }
  cudaMalloc(&d_opc, nops * sizeof(OpC)); cudaMalloc(&d_dw, nops * sizeof(double));
  cudaMalloc(&d_bad, 16 * 8); cudaMalloc(&d_cnt, 16 * 8);
  cudaMemset(d_bad, 0, 16 * 8); cudaMemset(d_cnt, 0, 16 * 8);
  cudaMemcpy(d_ops, plan.ops.data(), nops * sizeof(Op), cudaMemcpyHostToDevice);
  cudaMemcpy(d_dw, dw.data(), nops * sizeof(double), cudaMemcpyHostToDevice);
  prep<<<(nops + 63) / 64, 64>>>(d_ops, d_opc, nops, mass, charge);
  check<<<148 * 4, 256>>>(d_ops, d_opc, d_dw, nops, npart, mass, d_bad, d_cnt);
  std::vector<double2> fld(33 * 129);
  for (int j = 0; j <= 128; ++j)
    for (int i = 0; i <= 32; ++i) fld[j * 33 + i] = make_double2(1e-3 * sin(0.2 * i) * cos(0.05 * j), 1e-3 * cos(0.1 * i) * sin(0.07 * j));
  cudaMalloc(&d_fld, fld.size() * sizeof(double2));
  cudaMemcpy(d_fld, fld.data(), fld.size() * sizeof(double2), cudaMemcpyHostToDevice);
  check_kick<<<256, 256>>>(d_fld, 8192, mass, d_bad, d_cnt);
  unsigned long long bad[16], cnt[16];
  if (cudaMemcpy(bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost) != cudaSuccess) {
    fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 2;
  }
  const char* names[16] = {"none", "drift", "quad_half", "gap_first", "gap_second", "", "", "", "", "", "", "", "sc_kick", "", "", ""};
  unsigned long long total = 0;
  for (int t = 0; t < 16; ++t) if (cnt[t]) { printf("%-10s evaluated %llu not bit-identical %llu\n", names[t], cnt[t], bad[t]); total += bad[t]; }
  printf("MAPSCHECK mismatches %llu\n", total);
  return 0;
}
