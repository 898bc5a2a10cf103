// hps_capi.cpp -- implementation of the C ABI in include/hpsim_b200.h on top of the sm_100a kernels.
// Host side of the hot path: owns the device state, turns simulate() into a planned program
// (hps_plan.cpp) and enqueues it on the context's stream without any host synchronisation.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <cstring>
#include <ctime>
#include <string>
#include <strings.h>
#include <vector>

#include "hps_comm.h"
#include "hps_launch.h"
#include "hps_plan.h"

using namespace hps;

namespace {
std::string g_err;  // error of the last failed hps_ctx_create

u32 next_pow2(u32 v) {  // util/src/utility.cpp:7-16
  --v; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return ++v;
}
}  // namespace

struct hps_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int num_sms = 0;
  int compat = 1;
  std::string err;
  EngineState eng;
  // scratch shared by every reducing kernel of this context
  double* partials = nullptr;
  u32* ticket = nullptr;
  unsigned max_grid = 0;    // cap for the light statistic sweeps
  unsigned seg_grid_cap = 0;  // resident blocks of the segment kernel variant in use
  unsigned seg_grid_cap_full = 0;
  int seg_variant = 20;  // HPS_MATH_REFERENCE, 2 particles per thread (seg_kernel() in hps_kernels.cu)
  // moments + deposit + field solve in ONE launch (k_scheff, two grid barriers): opt-in with HPSIM_FUSED=1.  Measured
  // on C2 it loses to the three PDL-chained kernels (48.2 vs 44.2 ms per pass): the 64-register cap of a fully
  // resident grid spills in the deposit loop, and a grid barrier costs as much as a programmatic kernel boundary.
  int fused = 0;
  // the whole moment chain of a kick in the epilogue of the sweep before it (FIN_KICK_FUSED); HPSIM_FUSED_MOMENTS=0 keeps
  // the stand-alone moment kernel of round 1 (k_moments re-reads 40 B/particle per kick)
  int fused_moments = 1;
  // cp.async staging of the next particle pair in the sweep: -1 = automatic (on when the rank's particle state does not
  // fit the L2 anyway), 0 / 1 = forced (HPSIM_STAGE)
  int stage = -1;
  int patterns = 2;  // HPSIM_PATTERNS=0: always the generic sweep; 1: pattern sweeps at 2 blocks per SM; 2: 3 blocks per SM for the drift patterns
  // program upload buffers
  Op* d_ops = nullptr;
  OpC* d_opc = nullptr;  // device-prepared constants, parallel to d_ops
  Op* h_ops = nullptr;  // pinned
  size_t ops_cap = 0;
  Comm comm;
  // cross-rank exchange state of a particle-sharded beam (hps_ctx_comm_init)
  XrView xr;
  XrBox* xr_box = nullptr;
  double* xr_sync = nullptr;  // one device double: operand of the host-skew rendezvous (xr_rendezvous)
  void* xr_mapped[HPS_XR_MAXR] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t launches = 0;
  int pdl = 1;       // HPSIM_PDL=0 turns programmatic dependent launch off
  // CUDA graphs: the launch sequence of a planned range is captured once and replayed for every later simulate() of
  // the same program on the same objects (scans, benchmarks, optimisation loops re-run one range thousands of times).
  struct GraphEntry {
    uint64_t key = 0, stamp = 0, sweeps = 0, extra = 0, launches = 0;
    std::vector<unsigned char> material;  // everything the captured launches depend on; compared in full on a hash hit
    const void* beam = nullptr;           // the objects whose device pointers the graph captured: destroying or
    const void* scheff = nullptr;         // re-allocating one of them drops the entry (drop_graphs_of)
    const void* monitor = nullptr;
    cudaGraphExec_t exec = nullptr;
    Op* d_ops = nullptr;
    OpC* d_opc = nullptr;
  };
  std::vector<GraphEntry> graphs;
  uint64_t graph_stamp = 0;
  int use_graphs = 1;  // HPSIM_GRAPH=0 turns them off
  // L2 residency (opt-in, HPSIM_L2_PERSIST=1): the particle slab of the beam being tracked becomes the stream's
  // access-policy window (persisting lines).  Measured on C2: no gain (48.7 vs 48.0 ms per pass) -- the sweeps
  // are bound by fp64 latency, not by where the 59 MB beam is read from -- so it is off by default.
  int l2_persist = 0;
  size_t l2_persist_max = 0, l2_window_max = 0;
  const void* l2_window_base = nullptr;
  size_t l2_window_bytes = 0;
  u32 mesh_seq = 0;  // launch counter of k_field: sequence number of the mesh hand-off (same on all ranks)
  // optional per-launch timing (hps_ctx_set_profile): one event between consecutive launches
  int profile = 0;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_cls;
  double prof_ms[HPS_PROF_NCLASS] = {0, 0, 0, 0, 0};
  uint64_t prof_n[HPS_PROF_NCLASS] = {0, 0, 0, 0, 0};
  // caller-visible timing events and the L2-flush scratch (bench.py)
  cudaEvent_t user_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  char* flush_buf = nullptr;
  size_t flush_bytes = 0;
  hps_monitor* monitor = nullptr;  // param->graphics_on / plot_data of InitEngine (not owned)
  int comm_failed = 0;             // an NCCL collective of a Beam::Update* call failed (reported by the next getter)
};

struct hps_beam {
  hps_ctx* ctx;
  uint64_t gen = 0;  // unique per object (graph-cache key: a recycled address is not the same beam)
  BeamHost h;           // mass charge current freq design_w n_global
  u32 n = 0, offset = 0, gstride = 1, n_alloc = 0;
  char* slab = nullptr;  // one allocation: 6 fp64 arrays + 2 u32 arrays
  double* c[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  u32 *loss = nullptr, *lloss = nullptr;
  DevStats* st = nullptr;
  DeferList* defer = nullptr;  // side list of the fused-moment epilogue (hps_internal.h)
  double* tmp = nullptr;  // n doubles of scratch (relative-phase download)
  struct Slot { char* slab = nullptr; bool valid = false; double design_w, freq, current; } slot[2];
  size_t slab_bytes = 0;
};

struct hps_monitor {  // PlotData (util/inc/plot_data.h:43-66) + monitor_num / loss_num of the engine
  hps_ctx* ctx;
  u32 cap = 0;
  double* data = nullptr;   // [HPS_MON_NFIELD][cap]
  u32* counters = nullptr;  // [0] monitor_num, [1] loss_num
};

struct hps_lattice {
  hps_ctx* ctx;
  std::vector<hps_elem> e;
};

struct hps_scheff {
  hps_ctx* ctx;
  uint64_t gen = 0;  // bumped by every (re-)allocation of the tables
  ScHost h;
  u32 nr_init, nz_init;
  double interval_init = 0.01, mesh_w = 0.0;
  ScDev* dev = nullptr;
  double* bins = nullptr;
  double2* field = nullptr;
  double2* field_part = nullptr;  // [nr+1][chunks][nz+1] partial fields of the split solve
  u32* field_tickets = nullptr;   // [nr+1]
  double bin_scale = 0.0;         // fixed-point scale of the last simulate() (depends on the beam size)
  unsigned fused_cap = 0;         // co-resident blocks of k_scheff for this mesh (0: fused chain not available)
  // energy-dependent mesh: SetInterval() overrides what the last kick derived until the next kick has run
  bool interval_user_set = false;
  uint64_t interval_valid_after = 0;
};

namespace {
uint64_t g_next_gen = 1;

// CUDA graphs that captured device pointers of `obj` (a beam, a SCHEFF or a monitor) must not outlive them
void drop_graphs_of(hps_ctx* c, const void* obj) {
  bool synced = false;
  for (size_t i = 0; i < c->graphs.size();) {
    auto& g = c->graphs[i];
    if (g.beam == obj || g.scheff == obj || g.monitor == obj) {
      if (!synced) { cudaStreamSynchronize(c->stream); synced = true; }
      if (g.exec) cudaGraphExecDestroy(g.exec);
      cudaFree(g.d_ops); cudaFree(g.d_opc);
      c->graphs.erase(c->graphs.begin() + (long)i);
    } else ++i;
  }
}

int fail(hps_ctx* c, int code, const std::string& m) {
  if (c) c->err = m; else g_err = m;
  return code;
}
#define CK(ctx, call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, HPS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

BeamView view(const hps_beam* b) {
  BeamView v;
  v.x = b->c[0]; v.xp = b->c[1]; v.y = b->c[2]; v.yp = b->c[3]; v.phi = b->c[4]; v.w = b->c[5];
  v.loss = b->loss; v.lloss = b->lloss;
  v.n = b->n; v.offset = b->offset; v.gstride = b->gstride; v.n_global = b->h.n_global;
  const u32 np2 = next_pow2(b->h.n_global);
  v.np2_quarter = np2 >= 4 ? np2 / 4 : 1;
  v.streaming = (b->ctx->stage >= 0 ? b->ctx->stage != 0 : b->slab_bytes > ((size_t)96 << 20)) ? 1u : 0u;  // (as LaunchCfg::stage)
  v.mass = b->h.mass; v.charge = b->h.charge;
  return v;
}

LaunchCfg cfg(hps_ctx* c, const hps_beam* b) {
  LaunchCfg l;
  l.stream = c->stream;
  const unsigned pairs = (b->n + 1) / 2;
  l.seg_variant = c->seg_variant;
  const unsigned seg_items = segment_pairs(c->seg_variant) ? pairs : b->n;
  l.grid_seg = std::max(1u, std::min((seg_items + HPS_BLOCK - 1) / HPS_BLOCK, c->seg_grid_cap));
  l.grid_seg1 = std::max(1u, std::min((b->n + HPS_BLOCK - 1) / HPS_BLOCK, c->seg_grid_cap_full));
  l.grid_pairs = std::max(1u, std::min((pairs + HPS_BLOCK - 1) / HPS_BLOCK, c->max_grid));
  l.grid_single = std::max(1u, std::min((b->n + HPS_BLOCK - 1) / HPS_BLOCK, c->max_grid));
  l.grid_deposit = std::max(1u, std::min((pairs + 511) / 512, (unsigned)c->num_sms * 2));
  l.partials = c->partials;
  l.ticket = c->ticket;
  l.xr = c->xr;
  l.pdl = c->pdl && !c->profile;  // the per-launch events of profile mode sit between the kernels anyway
  l.stage = c->stage >= 0 ? c->stage : (b->slab_bytes > ((size_t)96 << 20) ? 1 : 0);
  l.patterns = c->patterns;
  l.num_sms = c->num_sms;
  return l;
}

// XR_NCCL: all-gather the partial vector the reducing kernel left in xr.send, then finish on the device
int xr_after_reduce(hps_ctx* c, const LaunchCfg& l, int kind, int fin_mode, int do_policy, const StatReq* rq,
                    const KickParams* kp, const BeamView& v, DevStats* st, ScDev* sc, double* bins) {
  if (c->xr.mode != XR_NCCL) return HPS_OK;
  if (comm_allgather(c->comm, c->xr.send, c->xr.recv, HPS_XR_NS * sizeof(double), c->stream))
    return fail(c, HPS_ERR_COMM, "ncclAllGather of a reduction's partials failed");
  FinReq f;
  std::memset(&f, 0, sizeof(f));
  f.kind = kind; f.fin_mode = fin_mode; f.do_policy = do_policy;
  f.ns = reduce_slots(kind); f.max_mask = reduce_max_mask(kind);
  if (rq) f.rq = *rq;
  if (kp) f.kp = *kp;
  launch_xr_finalize(l, f, v, st, sc, bins);
  c->launches++;
  return HPS_OK;
}

// Sharded beam, P2P exchange: a peer that did not answer within HPS_XR_TIMEOUT_NS left the error word set (the kernels
// carry on with stale data rather than hang the GPU).  Called at every synchronising entry point, with the stream idle:
// reports the failure once and clears the word.
int xr_check(hps_ctx* c) {
  if (c->xr.mode != XR_P2P || !c->xr.err) return HPS_OK;
  u32 xerr = 0;
  CK(c, cudaMemcpy(&xerr, c->xr.err, sizeof(u32), cudaMemcpyDeviceToHost));
  if (!xerr) return HPS_OK;
  CK(c, cudaMemset(c->xr.err, 0, sizeof(u32)));
  return fail(c, HPS_ERR_COMM, "a peer rank did not answer a cross-rank exchange within 2 s: the results since the last synchronisation are invalid");
}

// Sharded beam, P2P exchange: the in-kernel exchanges poll peer memory with a 2 s dead-peer limit, which must only ever
// see DEVICE-side skew (microseconds).  Host-side skew -- one rank reaching an entry point seconds after the others
// because its process was busy -- is taken out here: a one-double NCCL all-reduce enqueued on the context's stream at every
// entry point that goes on to launch exchanging kernels.  It is stream-ordered, so no rank's first exchanging kernel
// starts before every rank has arrived; NCCL itself waits for late peers without a limit.
int xr_rendezvous(hps_ctx* c) {
  if (c->xr.mode != XR_P2P || c->xr.nranks <= 1 || !c->xr_sync) return HPS_OK;
  if (comm_allreduce_sum_f64(c->comm, c->xr_sync, 1, c->stream)) return fail(c, HPS_ERR_COMM, "rendezvous all-reduce failed");
  return HPS_OK;
}

void carve(hps_beam* b, char* slab, double* c[6], u32** loss, u32** lloss) {
  const size_t da = align_up((size_t)b->n_alloc * sizeof(double), 256), ua = align_up((size_t)b->n_alloc * sizeof(u32), 256);
  for (int i = 0; i < 6; ++i) c[i] = reinterpret_cast<double*>(slab + i * da);
  *loss = reinterpret_cast<u32*>(slab + 6 * da);
  *lloss = reinterpret_cast<u32*>(slab + 6 * da + ua);
}

// ---- Beam::Update* family on the device ------------------------------------------------------
StatReq req(int kind, int a = 0, int bvar = -1, int good = 0, int out0 = 0) {
  StatReq r;
  std::memset(&r, 0, sizeof(r));
  r.kind = kind; r.var_a = a; r.var_b = bvar; r.good_only = good; r.out[0] = out0;
  return r;
}
void stat(hps_beam* b, const StatReq& r) {
  if (xr_rendezvous(b->ctx) != HPS_OK) b->ctx->comm_failed = 1;
  const LaunchCfg l = cfg(b->ctx, b);
  const BeamView v = view(b);
  launch_stat(l, v, b->st, r);
  b->ctx->launches++;
  // (XR_NCCL: a failed all-gather leaves its message in the context; hps_beam_update / hps_beam_get report it)
  if (xr_after_reduce(b->ctx, l, FK_STAT, 0, 0, &r, nullptr, v, b->st, nullptr, nullptr) != HPS_OK) b->ctx->comm_failed = 1;
}
void UpdateLoss(hps_beam* b) { stat(b, req(KS_COUNT_LOSS)); }
void UpdateGoodParticleCount(hps_beam* b) { stat(b, req(KS_COUNT_GOOD, 0, -1, 1)); }
void UpdateAvg(hps_beam* b, int var, int out, bool good) { stat(b, req(KS_AVG, var, -1, good, out)); }
void UpdateSig(hps_beam* b, int var, int out, bool good) {
  if (b->h.n_global > 2) stat(b, req(KS_SIG, var, -1, good, out));
}
void UpdateLongitudinalLoss(hps_beam* b) {  // beam.cpp:776-782
  UpdateGoodParticleCount(b);
  UpdateAvg(b, HPS_PHI, HPS_STAT_PHI_AVG_GOOD, true);
  stat(b, req(KS_MARK_LLOSS));
}
int ref_phi_of(hps_beam* b, double* out) {
  // phi[0] lives on the rank that owns global particle 0
  double v = 0.0;
  if (b->offset == 0) CK(b->ctx, cudaMemcpyAsync(&v, b->c[4], sizeof(double), cudaMemcpyDeviceToHost, b->ctx->stream));
  CK(b->ctx, cudaStreamSynchronize(b->ctx->stream));
  if (b->ctx->comm.nranks > 1) {
    int rc = comm_bcast_double(b->ctx->comm, &v, 0, b->ctx->stream);
    if (rc) return fail(b->ctx, HPS_ERR_COMM, "broadcast of the reference phase failed");
  }
  *out = v;
  return HPS_OK;
}
int ref_w_of(hps_beam* b, double* out) {
  double v = 0.0;
  if (b->offset == 0) CK(b->ctx, cudaMemcpyAsync(&v, b->c[5], sizeof(double), cudaMemcpyDeviceToHost, b->ctx->stream));
  CK(b->ctx, cudaStreamSynchronize(b->ctx->stream));
  if (b->ctx->comm.nranks > 1) {
    int rc = comm_bcast_double(b->ctx->comm, &v, 0, b->ctx->stream);
    if (rc) return fail(b->ctx, HPS_ERR_COMM, "broadcast of the reference energy failed");
  }
  *out = v;
  return HPS_OK;
}
int UpdateRelativePhi(hps_beam* b, bool use_good) {  // beam.cu:1122-1130
  double phi0 = 0.0;
  if (!use_good) { int rc = ref_phi_of(b, &phi0); if (rc) return rc; }
  launch_set_phi_ref(cfg(b->ctx, b), view(b), b->st, use_good ? 1 : 0, phi0);
  b->ctx->launches++;
  return HPS_OK;
}
int UpdateEmittance(hps_beam* b) {  // beam.cpp:981-990
  if (!(b->h.n_global > 2)) return HPS_OK;
  double w0;
  int rc = ref_w_of(b, &w0);
  if (rc) return rc;
  StatReq r = req(KS_EMIT_T, HPS_X, HPS_XP);
  r.w_ref = w0;
  r.out[0] = HPS_STAT_X_AVG; r.out[1] = HPS_STAT_X_SIG; r.out[2] = HPS_STAT_XP_AVG; r.out[3] = HPS_STAT_XP_SIG; r.out[4] = HPS_STAT_X_EMIT;
  stat(b, r);
  r = req(KS_EMIT_T, HPS_Y, HPS_YP);
  r.w_ref = w0;
  r.out[0] = HPS_STAT_Y_AVG; r.out[1] = HPS_STAT_Y_SIG; r.out[2] = HPS_STAT_YP_AVG; r.out[3] = HPS_STAT_YP_SIG; r.out[4] = HPS_STAT_Y_EMIT;
  stat(b, r);
  rc = UpdateRelativePhi(b, false);
  if (rc) return rc;
  stat(b, req(KS_EMIT_Z, 6, HPS_W));
  return HPS_OK;
}
void InitPhiAvgGood(hps_beam* b) {  // beam.cpp:655-661
  UpdateLoss(b);
  UpdateAvg(b, HPS_PHI, HPS_STAT_PHI_AVG, false);
  launch_copy_stat(cfg(b->ctx, b), b->st, HPS_STAT_PHI_AVG_GOOD, HPS_STAT_PHI_AVG);
  b->ctx->launches++;
}
// beam.cpp:1003-1016 through the same three kernels the kick loop uses
void UpdateStatForSpaceCharge(hps_beam* b) {
  if (!(b->h.n_global > 2)) return;
  hps_ctx* c = b->ctx;
  if (xr_rendezvous(c) != HPS_OK) c->comm_failed = 1;
  LaunchCfg l = cfg(c, b);
  BeamView v = view(b);
  KickParams kp;
  std::memset(&kp, 0, sizeof(kp));
  kp.n_global = b->h.n_global; kp.mass = b->h.mass;
  launch_segment(l, v, nullptr, c->d_ops, c->d_opc, 0, FIN_KICK, 0, c->compat, b->st, nullptr, nullptr, kp, nullptr, nullptr);
  xr_after_reduce(c, l, FK_SEGMENT, FIN_KICK, 0, nullptr, nullptr, v, b->st, nullptr, nullptr);
  launch_lloss_pass(l, v, b->st, nullptr);
  xr_after_reduce(c, l, FK_LLOSS, 0, 0, nullptr, nullptr, v, b->st, nullptr, nullptr);
  launch_moments(l, v, b->st, nullptr, nullptr, kp, 0, 0, 0);
  xr_after_reduce(c, l, FK_MOMENTS, 0, 0, nullptr, &kp, v, b->st, nullptr, nullptr);
  c->launches += 3;
}

// profiling: mark() before each launch of class `cls`, finish() after the last one
void prof_mark(hps_ctx* c, int cls) {
  if (!c->profile) return;
  const size_t k = c->prof_cls.size();
  if (k >= c->prof_ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->prof_ev.push_back(e);
  }
  cudaEventRecord(c->prof_ev[k], c->stream);
  c->prof_cls.push_back(cls);
}
void prof_finish(hps_ctx* c) {
  if (!c->profile || c->prof_cls.empty()) return;
  prof_mark(c, -1);
  cudaStreamSynchronize(c->stream);
  for (size_t k = 0; k + 1 < c->prof_cls.size(); ++k) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->prof_ev[k], c->prof_ev[k + 1]);
    c->prof_ms[c->prof_cls[k]] += ms;
    c->prof_n[c->prof_cls[k]]++;
  }
  c->prof_cls.clear();
}

int math_variant(int base, int mode) {  // base: 0 = 2 particles/thread, 1 / 2 = 1 particle/thread (3 / 4 blocks per SM)
  switch (mode) {
    case HPS_MATH_FAST: return base;
    case HPS_MATH_REFERENCE: return base == 0 ? 20 : 21;
    default: return base == 0 ? 10 : 11;  // HPS_MATH_REFERENCE_NVCC
  }
}

int dirty_mask(const Op* ops, u32 n) {
  int d = 0;
  for (u32 i = 0; i < n; ++i) switch (ops[i].type) {
      case OP_DRIFT: d |= 1 | 4 | 16 | 64; break;
      case OP_QUAD_HALF: d |= 1 | 2 | 4 | 8 | 16 | 64; break;
      case OP_GAP_FIRST: d |= 1 | 4 | 16; break;
      case OP_GAP_SECOND: d |= 127; break;
      case OP_DIPOLE_FIRST: case OP_DIPOLE_SECOND: d |= 1 | 2 | 4 | 8 | 16; break;
      case OP_BUNCHER: d |= 2 | 8 | 32 | 64; break;
      case OP_ROTATION: d |= 1 | 2 | 4 | 8; break;
      case OP_STEERER: d |= 2 | 8; break;
      case OP_CAPERTURE: case OP_RAPERTURE: d |= 64; break;
      case OP_SC_KICK: d |= 2 | 8 | 32 | 64; break;
      case OP_FREQ_CHANGE: d |= 16; break;
      default: break;
    }
  return d;
}

}  // namespace

// =================================================================================================
extern "C" {

int hps_abi_version(void) { return HPS_ABI_VERSION; }

hps_ctx* hps_ctx_create(int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fail(nullptr, HPS_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    return nullptr;
  }
  if (device < 0 || device >= ndev) { fail(nullptr, HPS_ERR_ARG, "device ordinal out of range"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { fail(nullptr, HPS_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    fail(nullptr, HPS_ERR_CUDA, "hpsim_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
    return nullptr;
  }
  hps_ctx* c = new hps_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess;
  std::memset(&c->xr, 0, sizeof(c->xr));
  c->xr.nranks = 1;
  if (const char* v = getenv("HPSIM_SEG_VARIANT")) c->seg_variant = atoi(v);  // raw variant number (diagnostics)
  if (const char* v = getenv("HPSIM_MATH"))  // "reference" / "fast" / "nvcc": same switch as hps_ctx_set_math
    c->seg_variant = math_variant(c->seg_variant % 10, v[0] == 'r' ? HPS_MATH_REFERENCE : (v[0] == 'n' ? HPS_MATH_REFERENCE_NVCC : HPS_MATH_FAST));
  if (const char* v = getenv("HPSIM_PDL")) c->pdl = atoi(v);
  if (const char* v = getenv("HPSIM_FUSED")) c->fused = atoi(v);
  if (const char* v = getenv("HPSIM_FUSED_MOMENTS")) c->fused_moments = atoi(v);
  if (const char* v = getenv("HPSIM_STAGE")) c->stage = atoi(v);
  if (const char* v = getenv("HPSIM_PATTERNS")) c->patterns = atoi(v);
  if (const char* v = getenv("HPSIM_GRAPH")) c->use_graphs = atoi(v);
  if (const char* v = getenv("HPSIM_L2_PERSIST")) c->l2_persist = atoi(v);
  c->l2_persist_max = (size_t)std::max(0, prop.persistingL2CacheMaxSize);
  c->l2_window_max = (size_t)std::max(0, prop.accessPolicyMaxWindowSize);
  if (c->l2_persist && c->l2_persist_max > 0 && c->l2_window_max > 0) {
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->l2_persist_max) != cudaSuccess) { cudaGetLastError(); c->l2_persist = 0; }
  } else c->l2_persist = 0;
  int occ = ok ? segment_occupancy(c->seg_variant) : 0;
  if (occ < 1) occ = 1;
  c->seg_grid_cap = (unsigned)c->num_sms * (unsigned)occ;
  c->seg_grid_cap_full = (unsigned)c->num_sms * (unsigned)std::max(1, ok ? segment_occupancy(segment_full_variant(c->seg_variant)) : 1);
  // the pair sweeps (K2/K3) hold a grid barrier: every block must be resident
  c->max_grid = (unsigned)c->num_sms * (unsigned)std::max(1, std::min(4, ok ? moments_occupancy() : 1));
  ok = ok && cudaMalloc(&c->partials, (size_t)c->num_sms * 8 * HPS_NSLOT * sizeof(double)) == cudaSuccess &&
       cudaMalloc(&c->ticket, 256) == cudaSuccess && cudaMemset(c->ticket, 0, 256) == cudaSuccess;
  c->ops_cap = 4096;
  ok = ok && cudaMalloc(&c->d_ops, c->ops_cap * sizeof(Op)) == cudaSuccess &&
       cudaMalloc(&c->d_opc, c->ops_cap * sizeof(OpC)) == cudaSuccess &&
       cudaHostAlloc(&c->h_ops, c->ops_cap * sizeof(Op), cudaHostAllocDefault) == cudaSuccess;
  if (!ok) {
    fail(nullptr, HPS_ERR_CUDA, std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError()));
    delete c;
    return nullptr;
  }
  return c;
}
void hps_ctx_destroy(hps_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (void* p : c->xr_mapped) if (p) cudaIpcCloseMemHandle(p);
  if (c->xr.seq) cudaFree(c->xr.seq);
  cudaFree(c->xr_box); cudaFree(c->xr.bins_gather); cudaFree(c->xr_sync);
  comm_destroy(c->comm);
  cudaFree(c->partials); cudaFree(c->ticket); cudaFree(c->d_ops); cudaFree(c->d_opc); cudaFreeHost(c->h_ops);
  for (auto& g : c->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); cudaFree(g.d_ops); cudaFree(g.d_opc); }
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : c->user_ev) if (e) cudaEventDestroy(e);
  cudaFree(c->flush_buf);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaStreamDestroy(c->stream);
  delete c;
}
const char* hps_last_error(const hps_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }
int hps_ctx_set_compat(hps_ctx* c, int v) { c->compat = v; return HPS_OK; }
int hps_ctx_get_math(hps_ctx* c) { return c->seg_variant >= 20 ? HPS_MATH_REFERENCE : (c->seg_variant >= 10 ? HPS_MATH_REFERENCE_NVCC : HPS_MATH_FAST); }
int hps_ctx_set_math(hps_ctx* c, int mode) {
  if (mode != HPS_MATH_REFERENCE && mode != HPS_MATH_FAST && mode != HPS_MATH_REFERENCE_NVCC) return fail(c, HPS_ERR_ARG, "unknown math mode");
  c->seg_variant = math_variant(c->seg_variant % 10, mode);
  const int occ = std::max(1, segment_occupancy(c->seg_variant));
  c->seg_grid_cap = (unsigned)c->num_sms * (unsigned)occ;
  c->seg_grid_cap_full = (unsigned)c->num_sms * (unsigned)std::max(1, segment_occupancy(segment_full_variant(c->seg_variant)));
  return HPS_OK;
}
int hps_ctx_synchronize(hps_ctx* c) {
  CK(c, cudaStreamSynchronize(c->stream));
  return xr_check(c);
}
uint64_t hps_ctx_launch_count(hps_ctx* c) { return c->launches; }
int hps_ctx_set_profile(hps_ctx* c, int on) {
  c->profile = on;
  for (int k = 0; k < HPS_PROF_NCLASS; ++k) { c->prof_ms[k] = 0.0; c->prof_n[k] = 0; }
  return HPS_OK;
}
int hps_ctx_get_profile(hps_ctx* c, double* ms, uint64_t* n) {
  for (int k = 0; k < HPS_PROF_NCLASS; ++k) { ms[k] = c->prof_ms[k]; n[k] = c->prof_n[k]; }
  return HPS_OK;
}

int hps_ctx_event_record(hps_ctx* c, int slot) {
  if (slot < 0 || slot >= 8) return fail(c, HPS_ERR_ARG, "event slot out of range");
  cudaSetDevice(c->device);
  if (!c->user_ev[slot]) CK(c, cudaEventCreate(&c->user_ev[slot]));
  CK(c, cudaEventRecord(c->user_ev[slot], c->stream));
  return HPS_OK;
}
int hps_ctx_event_elapsed(hps_ctx* c, int a, int b, double* ms) {
  if (a < 0 || a >= 8 || b < 0 || b >= 8 || !c->user_ev[a] || !c->user_ev[b]) return fail(c, HPS_ERR_ARG, "event slot not recorded");
  CK(c, cudaEventSynchronize(c->user_ev[b]));
  float f = 0.f;
  CK(c, cudaEventElapsedTime(&f, c->user_ev[a], c->user_ev[b]));
  if (ms) *ms = f;
  return HPS_OK;
}
int hps_ctx_flush_l2(hps_ctx* c) {
  cudaSetDevice(c->device);
  if (!c->flush_buf) {
    c->flush_bytes = (size_t)384 << 20;  // 3x the 126 MB L2
    CK(c, cudaMalloc(&c->flush_buf, c->flush_bytes));
  }
  CK(c, cudaMemsetAsync(c->flush_buf, 0x5a, c->flush_bytes, c->stream));
  if (c->l2_persist) {  // persisting lines survive ordinary traffic: drop them too
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaCtxResetPersistingL2Cache());
  }
  return HPS_OK;
}

int hps_comm_unique_id(void* out128) { return comm_unique_id(out128) ? HPS_ERR_COMM : HPS_OK; }
int hps_ctx_comm_init(hps_ctx* c, int rank, int nranks, const void* id) {
  if (nranks <= 1) return HPS_OK;
  if (nranks > HPS_XR_MAXR) return fail(c, HPS_ERR_ARG, "at most 8 ranks (one NVSwitch domain)");
  if (c->comm.nranks > 1) return fail(c, HPS_ERR_STATE, "communicator already initialised");
  cudaSetDevice(c->device);
  std::string err;
  if (comm_init(c->comm, rank, nranks, id, &err)) return fail(c, HPS_ERR_COMM, err);
  // exchange buffers: this rank's box, the NCCL staging vectors, the error word
  XrView& x = c->xr;
  std::memset(&x, 0, sizeof(x));
  x.rank = rank; x.nranks = nranks;
  char* misc = nullptr;  // [seq 2 x u64][err u32 + pad][send NS][recv R*NS][stats 4 x u64]
  const size_t misc_bytes = 32 + (size_t)(1 + nranks) * HPS_XR_NS * sizeof(double) + 4 * sizeof(unsigned long long);
  CK(c, cudaMalloc(&c->xr_box, sizeof(XrBox)));
  CK(c, cudaMemset(c->xr_box, 0, sizeof(XrBox)));
  CK(c, cudaMalloc(&misc, misc_bytes));
  CK(c, cudaMemset(misc, 0, misc_bytes));
  CK(c, cudaMalloc(&x.bins_gather, (size_t)nranks * HPS_XR_BINS_CAP * sizeof(double)));
  CK(c, cudaMalloc(&c->xr_sync, sizeof(double)));
  CK(c, cudaMemset(c->xr_sync, 0, sizeof(double)));
  x.seq = reinterpret_cast<unsigned long long*>(misc);
  x.err = reinterpret_cast<u32*>(misc + 16);
  x.send = reinterpret_cast<double*>(misc + 32);
  x.recv = x.send + HPS_XR_NS;
  x.stats = reinterpret_cast<unsigned long long*>(x.recv + (size_t)nranks * HPS_XR_NS);
  CK(c, cudaDeviceSynchronize());
  // map every peer's box (CUDA IPC; NVLink / NVSwitch peer access).  HPSIM_XR=nccl keeps to NCCL.
  const char* want = getenv("HPSIM_XR");
  bool p2p = !(want && std::strcmp(want, "nccl") == 0);
  double ok = 1.0;
  std::vector<cudaIpcMemHandle_t> handles(nranks);
  if (p2p) {
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, c->xr_box) != cudaSuccess) { ok = 0.0; std::memset(&mine, 0, sizeof(mine)); cudaGetLastError(); }
    char* d_h = nullptr;
    CK(c, cudaMalloc(&d_h, (size_t)(nranks + 1) * sizeof(mine)));
    CK(c, cudaMemcpyAsync(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
    if (comm_allgather(c->comm, d_h, d_h + sizeof(mine), sizeof(mine), c->stream))
      return fail(c, HPS_ERR_COMM, "all-gather of the IPC handles failed");
    CK(c, cudaMemcpyAsync(handles.data(), d_h + sizeof(mine), (size_t)nranks * sizeof(mine), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_h);
    for (int r = 0; r < nranks && ok == 1.0; ++r) {
      if (r == rank) { x.peer[r] = c->xr_box; continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, handles[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0.0; cudaGetLastError(); break; }
      c->xr_mapped[r] = p;
      x.peer[r] = reinterpret_cast<XrBox*>(p);
    }
    // every rank must have mapped every box, or all fall back to NCCL together
    double* d_ok = x.send;
    CK(c, cudaMemcpyAsync(d_ok, &ok, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (comm_allreduce_sum_f64(c->comm, d_ok, 1, c->stream)) return fail(c, HPS_ERR_COMM, "all-reduce failed");
    double all_ok = 0.0;
    CK(c, cudaMemcpyAsync(&all_ok, d_ok, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaMemsetAsync(x.send, 0, sizeof(double), c->stream));
    p2p = all_ok == (double)nranks;
  }
  x.mode = p2p ? XR_P2P : XR_NCCL;
  if (!p2p) x.peer[rank] = c->xr_box;
  return HPS_OK;
}
int hps_ctx_comm_mode(hps_ctx* c) { return c->xr.mode; }
int hps_ctx_comm_stats(hps_ctx* c, uint64_t out[4]) {
  for (int k = 0; k < 4; ++k) out[k] = 0;
  if (c->xr.nranks <= 1 || !c->xr.stats) return HPS_OK;
  cudaSetDevice(c->device);
  unsigned long long h[4];
  CK(c, cudaMemcpyAsync(h, c->xr.stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 4; ++k) out[k] = h[k];
  return HPS_OK;
}

void* hps_host_alloc(size_t bytes) {
  void* p = nullptr;
  return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
void hps_host_free(void* p) { cudaFreeHost(p); }

// ---- beam -----------------------------------------------------------------------------------------
hps_beam* hps_beam_create(hps_ctx* c, uint64_t n_global, double mass, double charge, double current) {
  if (!c) return nullptr;
  if (n_global == 0 || n_global > 0xffffffffull) { fail(c, HPS_ERR_ARG, "bad particle count"); return nullptr; }
  cudaSetDevice(c->device);
  hps_beam* b = new hps_beam();
  b->ctx = c;
  b->gen = g_next_gen++;
  b->h.mass = mass; b->h.charge = charge; b->h.current = current; b->h.freq = 0.0; b->h.design_w = 0.0;
  b->h.n_global = (u32)n_global;
  // Interleaved ownership: rank r holds the global particles r, r + R, r + 2R, ...  The C1 beam recipe
  // orders particles by phase, so contiguous blocks would give every rank a different slice of the bunch
  // (different loss rates -> unequal sweep times -> every exchange waits for the slowest rank).
  const int nr = std::max(1, c->comm.nranks), rk = c->comm.rank;
  b->offset = (u32)rk;
  b->gstride = (u32)nr;
  b->n = (u32)(n_global > (uint64_t)rk ? (n_global - rk + nr - 1) / nr : 0);
  b->n_alloc = (u32)align_up(std::max<u32>(b->n, 4), 4);
  const size_t da = align_up((size_t)b->n_alloc * sizeof(double), 256), ua = align_up((size_t)b->n_alloc * sizeof(u32), 256);
  b->slab_bytes = 6 * da + 2 * ua;
  bool ok = cudaMalloc(&b->slab, b->slab_bytes) == cudaSuccess && cudaMemsetAsync(b->slab, 0, b->slab_bytes, c->stream) == cudaSuccess &&
            cudaMalloc(&b->st, sizeof(DevStats)) == cudaSuccess && cudaMemsetAsync(b->st, 0, sizeof(DevStats), c->stream) == cudaSuccess &&
            cudaMemsetAsync(reinterpret_cast<char*>(b->st) + offsetof(DevStats, diag_t0), 0xff, sizeof(unsigned long long), c->stream) == cudaSuccess &&
            cudaMalloc(&b->defer, sizeof(DeferList)) == cudaSuccess && cudaMemsetAsync(b->defer, 0, 8, c->stream) == cudaSuccess &&
            cudaMalloc(&b->tmp, (size_t)b->n_alloc * sizeof(double)) == cudaSuccess;
  if (!ok) {
    fail(c, HPS_ERR_CUDA, std::string("beam allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
    cudaFree(b->slab); cudaFree(b->st); cudaFree(b->tmp); cudaFree(b->defer);
    delete b;
    return nullptr;
  }
  carve(b, b->slab, b->c, &b->loss, &b->lloss);
  return b;
}
void hps_beam_destroy(hps_beam* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  drop_graphs_of(b->ctx, b);
  if (b->ctx->l2_window_base == b->slab && b->slab) {  // the stream's L2 window must not outlive the slab
    cudaStreamAttrValue av;
    std::memset(&av, 0, sizeof(av));
    cudaStreamSetAttribute(b->ctx->stream, cudaStreamAttributeAccessPolicyWindow, &av);
    cudaCtxResetPersistingL2Cache();
    b->ctx->l2_window_base = nullptr;
    b->ctx->l2_window_bytes = 0;
  }
  cudaFree(b->slab); cudaFree(b->st); cudaFree(b->tmp); cudaFree(b->defer);
  for (auto& s : b->slot) cudaFree(s.slab);
  delete b;
}
int hps_beam_local_range(const hps_beam* b, uint64_t* off, uint64_t* nl) {
  if (off) *off = b->offset;
  if (nl) *nl = b->n;
  return HPS_OK;
}
uint64_t hps_beam_local_stride(const hps_beam* b) { return b->gstride; }

int hps_beam_upload(hps_beam* b, const double* x, const double* xp, const double* y, const double* yp,
                    const double* phi, const double* w, const uint32_t* loss, const uint32_t* lloss) {
  hps_ctx* c = b->ctx;
  cudaSetDevice(c->device);
  const double* src[6] = {x, xp, y, yp, phi, w};
  const size_t nb = (size_t)b->n * sizeof(double);
  for (int i = 0; i < 6; ++i) {
    if (!src[i]) return fail(c, HPS_ERR_ARG, "coordinate array is NULL");
    CK(c, cudaMemcpyAsync(b->c[i], src[i], nb, cudaMemcpyHostToDevice, c->stream));
  }
  if (loss) CK(c, cudaMemcpyAsync(b->loss, loss, (size_t)b->n * 4, cudaMemcpyHostToDevice, c->stream));
  else CK(c, cudaMemsetAsync(b->loss, 0, (size_t)b->n * 4, c->stream));
  if (lloss) CK(c, cudaMemcpyAsync(b->lloss, lloss, (size_t)b->n * 4, cudaMemcpyHostToDevice, c->stream));
  else CK(c, cudaMemsetAsync(b->lloss, 0, (size_t)b->n * 4, c->stream));
  if (b->offset == 0 && b->n > 0) {
    // reference particle: on the axis, never lost (beam.cpp:344-359)
    for (int i = 0; i < 4; ++i) CK(c, cudaMemsetAsync(b->c[i], 0, sizeof(double), c->stream));
    CK(c, cudaMemsetAsync(b->loss, 0, 4, c->stream));
    CK(c, cudaMemsetAsync(b->lloss, 0, 4, c->stream));
  }
  InitPhiAvgGood(b);
  return HPS_OK;
}
int hps_beam_download(hps_beam* b, double* x, double* xp, double* y, double* yp, double* phi, double* w,
                      uint32_t* loss, uint32_t* lloss) {
  hps_ctx* c = b->ctx;
  cudaSetDevice(c->device);
  UpdateLoss(b);  // wrap_beam.cpp:456
  double* dst[6] = {x, xp, y, yp, phi, w};
  for (int i = 0; i < 6; ++i)
    if (dst[i]) CK(c, cudaMemcpyAsync(dst[i], b->c[i], (size_t)b->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (loss) CK(c, cudaMemcpyAsync(loss, b->loss, (size_t)b->n * 4, cudaMemcpyDeviceToHost, c->stream));
  if (lloss) CK(c, cudaMemcpyAsync(lloss, b->lloss, (size_t)b->n * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return xr_check(c);
}
int hps_beam_download_phi_relative(hps_beam* b, int use_good, double* out) {
  hps_ctx* c = b->ctx;
  cudaSetDevice(c->device);
  int rc = UpdateRelativePhi(b, use_good != 0);
  if (rc) return rc;
  launch_rel_phi_out(cfg(c, b), view(b), b->st, b->tmp);
  c->launches++;
  CK(c, cudaMemcpyAsync(out, b->tmp, (size_t)b->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return HPS_OK;
}

// Host generators.  Every rank draws the complete glibc rand() sequence and keeps its own block, so a
// sharded beam is bit-identical to the single-GPU one.
static int upload_generated(hps_beam* b, std::vector<double> v[6]) {
  if (b->gstride == 1)
    return hps_beam_upload(b, v[0].data(), v[1].data(), v[2].data(), v[3].data(), v[4].data(), v[5].data(), nullptr, nullptr);
  std::vector<double> loc[6];
  for (int k = 0; k < 6; ++k) {
    loc[k].resize(b->n);
    for (u32 i = 0; i < b->n; ++i) loc[k][i] = v[k][(size_t)b->offset + (size_t)i * b->gstride];
  }
  return hps_beam_upload(b, loc[0].data(), loc[1].data(), loc[2].data(), loc[3].data(), loc[4].data(), loc[5].data(), nullptr, nullptr);
}
static inline double urand() { return 2.0 * std::rand() / RAND_MAX - 1.0; }

int hps_beam_set_dc(hps_beam* b, double ax, double bx, double ex, double ay, double by, double ey, double dphi,
                    double sync_phi, double sync_w, uint32_t seed) {
  // Beam::InitDCBeam (beam.cpp:481-537): uniform 4-D ball by nested rejection, phase ramp, constant W
  const u32 n = b->h.n_global;
  std::vector<double> v[6];
  for (auto& a : v) a.resize(n);
  std::srand(seed == 0 ? (unsigned)std::time(NULL) : seed);
  const double gx = (1.0 + ax * ax) / bx, xmx = std::sqrt(ex / gx), xpmx = std::sqrt(ex * gx), dx = -ax / gx;
  const double gy = (1.0 + ay * ay) / by, ymx = std::sqrt(ey / gy), ypmx = std::sqrt(ey * gy), dy = -ay / gy;
  for (u32 i = 0; i < n; ++i) {
    double r[4];
    for (;;) {
      for (int pair = 0; pair < 2; ++pair)
        do { r[2 * pair] = urand(); r[2 * pair + 1] = urand(); } while (r[2 * pair] * r[2 * pair] + r[2 * pair + 1] * r[2 * pair + 1] > 1.0);
      if (!(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3] > 1.0)) break;
    }
    const double ramp = 2.0 * (double)(i + 1.0) / (double)n - 1.0;
    v[0][i] = 0.01 * (r[0] * xmx + dx * r[1] * xpmx);
    v[1][i] = r[1] * xpmx;
    v[2][i] = 0.01 * (r[2] * ymx + dy * r[3] * ypmx);
    v[3][i] = r[3] * ypmx;
    v[4][i] = ramp * dphi + sync_phi;
    v[5][i] = sync_w;
  }
  return upload_generated(b, v);
}
int hps_beam_set_waterbag(hps_beam* b, double ax, double bx, double ex, double ay, double by, double ey, double az,
                          double bz, double ez, double sync_phi, double sync_w, double freq, uint32_t seed) {
  // Beam::InitWaterbagBeam (beam.cpp:391-463): uniform 6-D ball; also sets the beam frequency (:421)
  const u32 n = b->h.n_global;
  std::vector<double> v[6];
  for (auto& a : v) a.resize(n);
  std::srand(seed == 0 ? (unsigned)std::time(NULL) : seed);
  const double gx = (1.0 + ax * ax) / bx, xmx = std::sqrt(ex / gx), xpmx = std::sqrt(ex * gx), dx = -ax / gx;
  const double gy = (1.0 + ay * ay) / by, ymx = std::sqrt(ey / gy), ypmx = std::sqrt(ey * gy), dy = -ay / gy;
  const double gz = (1.0 + az * az) / bz, zmx = std::sqrt(ez / gz), zpmx = ez / zmx, dz = -az / gz;
  const double wave_len = 29979.2458 / freq;  // cm (H15)
  b->h.freq = freq;
  const double gm = sync_w / b->h.mass + 1.0;
  const double beta = std::sqrt(1.0 - 1.0 / (gm * gm));
  for (u32 i = 0; i < n; ++i) {
    double r[6];
    for (;;) {
      for (int pair = 0; pair < 3; ++pair)
        do { r[2 * pair] = urand(); r[2 * pair + 1] = urand(); } while (r[2 * pair] * r[2 * pair] + r[2 * pair + 1] * r[2 * pair + 1] > 1.0);
      if (!(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3] + r[4] * r[4] + r[5] * r[5] > 1.0)) break;
    }
    v[0][i] = 0.01 * (r[0] * xmx + dx * r[1] * xpmx);
    v[1][i] = r[1] * xpmx;
    v[2][i] = 0.01 * (r[2] * ymx + dy * r[3] * ypmx);
    v[3][i] = r[3] * ypmx;
    v[4][i] = -(r[4] * zmx + dz * zpmx * r[5]) * 2.0 * M_PI / (beta * wave_len) + sync_phi;
    v[5][i] = 2.0 * sync_w * zpmx * r[5] + sync_w;
  }
  return upload_generated(b, v);
}

int hps_beam_save(hps_beam* b, int k) {  // beam.cpp:547-593 (device-resident copy)
  hps_ctx* c = b->ctx;
  if (k < 0 || k > 1) return fail(c, HPS_ERR_ARG, "slot must be 0 or 1");
  cudaSetDevice(c->device);
  auto& s = b->slot[k];
  if (!s.slab) CK(c, cudaMalloc(&s.slab, b->slab_bytes));
  CK(c, cudaMemcpyAsync(s.slab, b->slab, b->slab_bytes, cudaMemcpyDeviceToDevice, c->stream));
  s.design_w = b->h.design_w; s.freq = b->h.freq; s.current = b->h.current;
  s.valid = true;
  return HPS_OK;
}
int hps_beam_restore(hps_beam* b, int k) {  // beam.cpp:603-647
  hps_ctx* c = b->ctx;
  if (k < 0 || k > 1) return fail(c, HPS_ERR_ARG, "slot must be 0 or 1");
  auto& s = b->slot[k];
  if (!s.valid) return fail(c, HPS_ERR_STATE, "Beam can't be restored because no previous beam has been saved");
  cudaSetDevice(c->device);
  CK(c, cudaMemcpyAsync(b->slab, s.slab, b->slab_bytes, cudaMemcpyDeviceToDevice, c->stream));
  b->h.design_w = s.design_w; b->h.freq = s.freq; b->h.current = s.current;
  UpdateAvg(b, HPS_PHI, HPS_STAT_PHI_AVG_GOOD, true);  // H21: divides by the pre-restore num_good
  return UpdateRelativePhi(b, true);
}

int hps_beam_set_ref_w(hps_beam* b, double w) {  // beam.cpp:1051-1055
  hps_ctx* c = b->ctx;
  cudaSetDevice(c->device);
  if (b->offset == 0) CK(c, cudaMemcpyAsync(b->c[5], &w, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  b->h.design_w = w;
  return HPS_OK;
}
int hps_beam_set_ref_phi(hps_beam* b, double phi) {
  hps_ctx* c = b->ctx;
  cudaSetDevice(c->device);
  if (b->offset == 0) CK(c, cudaMemcpyAsync(b->c[4], &phi, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  return HPS_OK;
}
int hps_beam_set_frequency(hps_beam* b, double f) { b->h.freq = f; return HPS_OK; }
double hps_beam_get_ref_w(hps_beam* b) { double v = 0; cudaSetDevice(b->ctx->device); ref_w_of(b, &v); return v; }
double hps_beam_get_ref_phi(hps_beam* b) { double v = 0; cudaSetDevice(b->ctx->device); ref_phi_of(b, &v); return v; }
double hps_beam_get_frequency(const hps_beam* b) { return b->h.freq; }
double hps_beam_get_mass(const hps_beam* b) { return b->h.mass; }
double hps_beam_get_charge(const hps_beam* b) { return b->h.charge; }
double hps_beam_get_current(const hps_beam* b) { return b->h.current; }
uint64_t hps_beam_get_size(const hps_beam* b) { return b->h.n_global; }

int hps_beam_shift(hps_beam* b, int coord, double v) {  // beam.cpp:1595-1642
  if (coord < 0 || coord > 5) return fail(b->ctx, HPS_ERR_ARG, "bad coordinate");
  cudaSetDevice(b->ctx->device);
  if (b->n) launch_shift(cfg(b->ctx, b), b->c[coord], v, b->n, b->offset);
  b->ctx->launches++;
  if (coord == HPS_PHI) InitPhiAvgGood(b);
  return HPS_OK;
}
int hps_beam_cut(hps_beam* b, int coord, double lo, double hi) {  // beam.cpp:1574-1588
  if (!(coord == HPS_X || coord == HPS_Y || coord == HPS_W || coord == HPS_PHI))
    return fail(b->ctx, HPS_ERR_ARG, "cut supports x, y, w, phi");
  cudaSetDevice(b->ctx->device);
  if (b->n) launch_cut(cfg(b->ctx, b), b->c[coord], b->loss, lo, hi, b->n, b->offset);
  b->ctx->launches++;
  return HPS_OK;
}

int hps_beam_update(hps_beam* b, int id) {
  cudaSetDevice(b->ctx->device);
  switch (id) {
    case HPS_UPD_LOSS: UpdateLoss(b); break;
    case HPS_UPD_LLOSS: UpdateLongitudinalLoss(b); break;
    case HPS_UPD_GOOD_COUNT: UpdateGoodParticleCount(b); break;
    case HPS_UPD_AVG_X: UpdateAvg(b, HPS_X, HPS_STAT_X_AVG, false); break;
    case HPS_UPD_AVG_X_GOOD: UpdateAvg(b, HPS_X, HPS_STAT_X_AVG_GOOD, true); break;
    case HPS_UPD_AVG_XP: UpdateAvg(b, HPS_XP, HPS_STAT_XP_AVG, false); break;
    case HPS_UPD_AVG_Y: UpdateAvg(b, HPS_Y, HPS_STAT_Y_AVG, false); break;
    case HPS_UPD_AVG_Y_GOOD: UpdateAvg(b, HPS_Y, HPS_STAT_Y_AVG_GOOD, true); break;
    case HPS_UPD_AVG_YP: UpdateAvg(b, HPS_YP, HPS_STAT_YP_AVG, false); break;
    case HPS_UPD_AVG_PHI: UpdateAvg(b, HPS_PHI, HPS_STAT_PHI_AVG, false); break;
    case HPS_UPD_AVG_PHI_GOOD: UpdateAvg(b, HPS_PHI, HPS_STAT_PHI_AVG_GOOD, true); break;
    case HPS_UPD_AVG_PHI_R: UpdateAvg(b, 6, HPS_STAT_PHI_AVG_R, false); break;
    case HPS_UPD_AVG_PHI_R_GOOD: UpdateAvg(b, 6, HPS_STAT_PHI_AVG_R, true); break;
    case HPS_UPD_AVG_W: UpdateAvg(b, HPS_W, HPS_STAT_W_AVG, false); break;
    case HPS_UPD_AVG_W_GOOD: UpdateAvg(b, HPS_W, HPS_STAT_W_AVG_GOOD, true); break;
    case HPS_UPD_SIG_X: UpdateSig(b, HPS_X, HPS_STAT_X_SIG, false); break;
    case HPS_UPD_SIG_X_GOOD: UpdateSig(b, HPS_X, HPS_STAT_X_SIG_GOOD, true); break;
    case HPS_UPD_SIG_XP: UpdateSig(b, HPS_XP, HPS_STAT_XP_SIG, false); break;
    case HPS_UPD_SIG_Y: UpdateSig(b, HPS_Y, HPS_STAT_Y_SIG, false); break;
    case HPS_UPD_SIG_Y_GOOD: UpdateSig(b, HPS_Y, HPS_STAT_Y_SIG_GOOD, true); break;
    case HPS_UPD_SIG_YP: UpdateSig(b, HPS_YP, HPS_STAT_YP_SIG, false); break;
    case HPS_UPD_SIG_PHI: UpdateSig(b, HPS_PHI, HPS_STAT_PHI_SIG, false); break;
    case HPS_UPD_SIG_PHI_R: UpdateSig(b, 6, HPS_STAT_PHI_SIG_R, false); break;
    case HPS_UPD_SIG_PHI_R_GOOD: UpdateSig(b, 6, HPS_STAT_PHI_SIG_GOOD, true); break;
    case HPS_UPD_SIG_W: UpdateSig(b, HPS_W, HPS_STAT_W_SIG, false); break;
    case HPS_UPD_SIG_W_GOOD: UpdateSig(b, HPS_W, HPS_STAT_W_SIG_GOOD, true); break;
    case HPS_UPD_EMITTANCE: return UpdateEmittance(b);
    case HPS_UPD_STAT_FOR_SPCH: UpdateStatForSpaceCharge(b); break;
    case HPS_UPD_AVG_XY: stat(b, req(KS_AVG_XY, HPS_X, HPS_Y)); break;
    case HPS_UPD_MAX_R_PHI: stat(b, req(KS_MAX_R_PHI)); break;
    case HPS_UPD_REL_PHI: return UpdateRelativePhi(b, false);
    case HPS_UPD_REL_PHI_GOOD: return UpdateRelativePhi(b, true);
    default: return fail(b->ctx, HPS_ERR_ARG, "unknown update id");
  }
  return HPS_OK;
}
hps_monitor* hps_monitor_create(hps_ctx* c, uint32_t capacity) {
  if (!c || capacity == 0) return nullptr;
  cudaSetDevice(c->device);
  hps_monitor* m = new hps_monitor();
  m->ctx = c; m->cap = capacity;
  const size_t bytes = (size_t)HPS_MON_NFIELD * capacity * sizeof(double);
  if (cudaMalloc(&m->data, bytes) != cudaSuccess || cudaMalloc(&m->counters, 2 * sizeof(u32)) != cudaSuccess ||
      cudaMemsetAsync(m->data, 0, bytes, c->stream) != cudaSuccess || cudaMemsetAsync(m->counters, 0, 2 * sizeof(u32), c->stream) != cudaSuccess) {
    fail(c, HPS_ERR_CUDA, "monitor allocation failed");
    cudaFree(m->data); cudaFree(m->counters);
    delete m;
    return nullptr;
  }
  return m;
}
void hps_monitor_destroy(hps_monitor* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  drop_graphs_of(m->ctx, m);
  if (m->ctx->monitor == m) m->ctx->monitor = nullptr;
  cudaFree(m->data); cudaFree(m->counters);
  delete m;
}
int hps_ctx_set_monitor(hps_ctx* c, hps_monitor* m) {
  if (m && m->ctx != c) return fail(c, HPS_ERR_ARG, "monitor belongs to another context");
  c->monitor = m;
  return HPS_OK;
}
int hps_monitor_download(hps_monitor* m, double* out, uint32_t* count) {
  hps_ctx* c = m->ctx;
  cudaSetDevice(c->device);
  u32 h[2] = {0, 0};
  if (out) CK(c, cudaMemcpyAsync(out, m->data, (size_t)HPS_MON_NFIELD * m->cap * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaMemcpyAsync(h, m->counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  if (count) *count = h[0];
  return HPS_OK;
}

int hps_beam_moment_stats(hps_beam* b, uint64_t out[18], double* pivot_err_max, double* pivot_margin) {
  hps_ctx* c = b->ctx;
  cudaSetDevice(c->device);
  DevStats h;
  CK(c, cudaMemcpyAsync(&h, b->st, sizeof(DevStats), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  if (out) {
    out[0] = h.n_fused; out[1] = h.n_fallback; out[2] = h.n_deferred;
    for (int k = 0; k < 6; ++k) out[3 + k] = h.diag_ns[k];
    for (int k = 0; k < 3; ++k) out[9 + k] = h.n_fb[k];
    for (int k = 0; k < 6; ++k) out[12 + k] = h.fin_ns[k];
  }
  if (pivot_err_max) *pivot_err_max = h.pivot_err_max;
  if (pivot_margin) *pivot_margin = h.pivot_m;
  return HPS_OK;
}
int hps_beam_get(hps_beam* b, int id, double* out) {
  hps_ctx* c = b->ctx;
  if (id < 0 || id >= HPS_STAT_COUNT) return fail(c, HPS_ERR_ARG, "unknown stat id");
  if (c->comm_failed) { c->comm_failed = 0; return HPS_ERR_COMM; }
  cudaSetDevice(c->device);
  DevStats h;
  CK(c, cudaMemcpyAsync(&h, b->st, sizeof(DevStats), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  if (id == HPS_STAT_NUM_LOSS) *out = h.num_loss;
  else if (id == HPS_STAT_NUM_LLOSS) *out = h.num_lloss;
  else if (id == HPS_STAT_NUM_GOOD) *out = h.num_good;
  else *out = h.v[id];
  return HPS_OK;
}

// ---- lattice ----------------------------------------------------------------------------------------
hps_lattice* hps_lattice_create(hps_ctx* c, const hps_elem* e, uint32_t n) {
  if (!c || (!e && n)) return nullptr;
  hps_lattice* l = new hps_lattice();
  l->ctx = c;
  l->e.assign(e, e + n);
  return l;
}
void hps_lattice_destroy(hps_lattice* l) { delete l; }
uint32_t hps_lattice_size(const hps_lattice* l) { return (uint32_t)l->e.size(); }
int64_t hps_lattice_index(const hps_lattice* l, const char* name) {  // beamline.cpp:70-77
  for (size_t i = 0; i < l->e.size(); ++i)
    if (strncasecmp(l->e[i].name, name, HPS_NAME_LEN) == 0) return (int64_t)i;
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

// ---- SCHEFF -----------------------------------------------------------------------------------------
static int scheff_alloc(hps_scheff* s) {
  hps_ctx* c = s->ctx;
  drop_graphs_of(c, s);
  s->gen = g_next_gen++;
  cudaFree(s->bins); cudaFree(s->field); cudaFree(s->field_part); cudaFree(s->field_tickets);
  s->bins = nullptr; s->field = nullptr; s->field_part = nullptr; s->field_tickets = nullptr;
  const size_t nb = (size_t)s->h.nr * s->h.nz, nf = (size_t)(s->h.nr + 1) * (s->h.nz + 1);
  KickParams kp;
  std::memset(&kp, 0, sizeof(kp));
  kp.nr = s->h.nr; kp.nz = s->h.nz;
  CK(c, cudaMalloc(&s->bins, (nb + 2) * sizeof(double)));  // + on-mesh count, max |phi| (sharded hand-off)
  CK(c, cudaMalloc(&s->field, nf * sizeof(double2)));
  const size_t chunks = std::max<size_t>(field_chunks(kp), scheff_supported(kp) ? scheff_items(kp) / (s->h.nr + 1) : 0);
  CK(c, cudaMalloc(&s->field_part, nf * chunks * sizeof(double2)));
  s->fused_cap = scheff_supported(kp) ? (unsigned)std::max(0, scheff_occupancy(kp)) * (unsigned)c->num_sms : 0u;
  CK(c, prepare_chain_attributes(kp));
  CK(c, cudaMalloc(&s->field_tickets, (s->h.nr + 1) * sizeof(u32)));
  CK(c, cudaMemsetAsync(s->field_tickets, 0, (s->h.nr + 1) * sizeof(u32), c->stream));
  CK(c, cudaMemsetAsync(s->bins, 0, (nb + 2) * sizeof(double), c->stream));
  CK(c, cudaMemsetAsync(s->field, 0, nf * sizeof(double2), c->stream));
  return HPS_OK;
}
static int scheff_check_mesh(hps_ctx* c, u32 nr, u32 nz) {
  KickParams kp;
  std::memset(&kp, 0, sizeof(kp));
  kp.nr = nr; kp.nz = nz;
  const int th = field_threads(kp);
  if (nr == 0 || nz == 0 || th == 0 || field_smem(kp, th) > 200 * 1024 || deposit_smem(kp) > 200 * 1024)
    return fail(c, HPS_ERR_ARG, "SCHEFF mesh too large for the shared-memory deposit / field solve (need nz <= 1024 and 8*nr*nz <= 200 KB)");
  return HPS_OK;
}
hps_scheff* hps_scheff_create(hps_ctx* c, uint32_t nr, uint32_t nz, int adj_bunch) {
  if (!c) return nullptr;
  if (scheff_check_mesh(c, nr, nz)) return nullptr;
  cudaSetDevice(c->device);
  hps_scheff* s = new hps_scheff();
  s->ctx = c;
  s->h.on = true; s->h.nr = nr; s->h.nz = nz; s->h.adj_bunch = (u32)adj_bunch;
  s->nr_init = nr; s->nz_init = nz;
  bool ok = cudaMalloc(&s->dev, sizeof(ScDev)) == cudaSuccess &&
            cudaMemsetAsync(s->dev, 0, sizeof(ScDev), c->stream) == cudaSuccess && scheff_alloc(s) == HPS_OK;
  if (!ok) {
    if (c->err.empty()) fail(c, HPS_ERR_CUDA, "SCHEFF allocation failed");
    cudaFree(s->dev); cudaFree(s->bins); cudaFree(s->field); cudaFree(s->field_part); cudaFree(s->field_tickets);
    delete s;
    return nullptr;
  }
  return s;
}
void hps_scheff_destroy(hps_scheff* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  drop_graphs_of(s->ctx, s);
  cudaFree(s->dev); cudaFree(s->bins); cudaFree(s->field); cudaFree(s->field_part); cudaFree(s->field_tickets);
  delete s;
}
int hps_scheff_set(hps_scheff* s, int p, double v) {
  switch (p) {
    case HPS_SC_INTERVAL: s->h.interval = v; s->interval_init = v; s->interval_user_set = true; break;
    case HPS_SC_ADJ_BUNCH: s->h.adj_bunch = (u32)(int)v; break;
    case HPS_SC_ADJ_BUNCH_CUTOFF_W: s->h.adj_bunch_w = v; break;
    case HPS_SC_MESH_CUTOFF_W: s->mesh_w = v; break;
    case HPS_SC_REMESH_THRESHOLD: s->h.remesh_val = v; break;
    case HPS_SC_FRACTION: s->h.fraction = v; break;
    default: return fail(s->ctx, HPS_ERR_ARG, "parameter is read-only or unknown");
  }
  return HPS_OK;
}
int hps_scheff_get(hps_scheff* s, int p, double* v) {
  switch (p) {
    case HPS_SC_INTERVAL: *v = s->h.interval; break;
    case HPS_SC_ADJ_BUNCH: *v = s->h.adj_bunch; break;
    case HPS_SC_ADJ_BUNCH_CUTOFF_W: *v = s->h.adj_bunch_w; break;
    case HPS_SC_MESH_CUTOFF_W: *v = s->mesh_w; break;
    case HPS_SC_REMESH_THRESHOLD: *v = s->h.remesh_val; break;
    case HPS_SC_FRACTION: *v = s->h.fraction; break;
    case HPS_SC_NR:
    case HPS_SC_NZ: {  // GetNx / GetNz report the mesh in force: with the energy cutoff on it lives on the device
      *v = p == HPS_SC_NR ? s->h.nr : s->h.nz;
      if (s->mesh_w != 0.0) {
        ScDev h;
        CK(s->ctx, cudaMemcpyAsync(&h, s->dev, sizeof(ScDev), cudaMemcpyDeviceToHost, s->ctx->stream));
        CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
        if (h.nr != 0u) *v = p == HPS_SC_NR ? h.nr : h.nz;
      }
      break;
    }
    default: return fail(s->ctx, HPS_ERR_ARG, "unknown parameter");
  }
  return HPS_OK;
}
int hps_scheff_set_mesh_size(hps_scheff* s, uint32_t nr, uint32_t nz) {  // space_charge.cpp:62-74
  if (nr == s->h.nr && nz == s->h.nz) return HPS_OK;
  int rc = scheff_check_mesh(s->ctx, nr, nz);
  if (rc) return rc;
  cudaSetDevice(s->ctx->device);
  CK(s->ctx, cudaStreamSynchronize(s->ctx->stream));
  s->h.nr = nr; s->h.nz = nz;
  // the device copy of "mesh in force" (energy-dependent mesh) falls back to the configured size
  CK(s->ctx, cudaMemsetAsync(reinterpret_cast<char*>(s->dev) + offsetof(ScDev, nr), 0, 2 * sizeof(u32), s->ctx->stream));
  return scheff_alloc(s);
}
int hps_scheff_phase_cycles(hps_scheff* s, uint64_t out[8]) {
  hps_ctx* c = s->ctx;
  ScDev h;
  CK(c, cudaMemcpyAsync(&h, s->dev, sizeof(ScDev), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 8; ++i) out[i] = h.phase_clk[i];
  return HPS_OK;
}
int hps_scheff_download_tables(hps_scheff* s, double* bins, double* field, double* dr_dz) {
  hps_ctx* c = s->ctx;
  cudaSetDevice(c->device);
  if (bins) CK(c, cudaMemcpyAsync(bins, s->bins, (size_t)s->h.nr * s->h.nz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (field) CK(c, cudaMemcpyAsync(field, s->field, (size_t)(s->h.nr + 1) * (s->h.nz + 1) * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
  ScDev h;
  CK(c, cudaMemcpyAsync(&h, s->dev, sizeof(ScDev), cudaMemcpyDeviceToHost, c->stream));
  CK(c, cudaStreamSynchronize(c->stream));
  if (bins) {  // the device mesh is 64-bit fixed point (k_deposit)
    const size_t nb = (size_t)s->h.nr * s->h.nz;
    const double scale = s->bin_scale > 0.0 ? s->bin_scale : 1.0;
    for (size_t k = 0; k < nb; ++k) {
      long long q;
      std::memcpy(&q, &bins[k], sizeof(q));
      bins[k] = (double)q / scale;
    }
  }
  if (dr_dz) { dr_dz[0] = h.dr; dr_dz[1] = h.dz; }
  return HPS_OK;
}

// ---- simulate ---------------------------------------------------------------------------------------
int hps_engine_reset(hps_ctx* c, hps_beam* b) {  // simulation_engine.cu:101-111
  if (c->monitor) { cudaSetDevice(c->device); cudaMemsetAsync(c->monitor->counters, 0, 2 * sizeof(u32), c->stream); }
  c->eng.ccl_cell_num = 0;
  c->eng.design_w = b->h.design_w;
  c->eng.wavelen = 299.792458 / b->h.freq;
  return HPS_OK;
}

// Pre-filter of the graph cache (a hit is confirmed by comparing the whole key material): FNV-style, eight bytes per step --
// the material of a full-linac program is 7 MB, and the byte-wise form cost 14 ms of host time per simulate() call.
static uint64_t fnv(uint64_t h, const void* p, size_t n) {
  const unsigned char* q = static_cast<const unsigned char*>(p);
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w;
    std::memcpy(&w, q + i, 8);
    h = (h ^ w) * 1099511628211ull;
    h ^= h >> 29;
  }
  for (; i < n; ++i) { h ^= q[i]; h *= 1099511628211ull; }
  return h;
}

// The launch sequence of one plan on the context's stream (directly, or into a stream capture).
static int launch_plan(hps_ctx* c, hps_beam* b, hps_scheff* sc, const Plan& plan, const Op* d_ops, OpC* d_opc, uint64_t* sweeps_io,
                       uint64_t* extra_io) {
  const LaunchCfg l = cfg(c, b);
  const BeamView v = view(b);
  launch_prepare_ops(l, d_ops, d_opc, (int)plan.ops.size(), b->h.mass, b->h.charge);
  c->launches++;
  uint64_t sweeps = 0, extra = 0;
  // fused moments: the sweep in front of a kick finishes the kick's statistics and the Scheff::Start policy itself
  const bool fused_mom = c->fused_moments && sc && !c->fused && c->xr.mode != XR_NCCL;
  MonView mon;
  std::memset(&mon, 0, sizeof(mon));
  if (c->monitor) {
    mon.data = c->monitor->data; mon.counters = c->monitor->counters; mon.cap = c->monitor->cap;
    if (plan.engine_reset) cudaMemsetAsync(mon.counters, 0, 2 * sizeof(u32), c->stream);  // Reset(): monitor_num = loss_num = 0
  }
  for (const Segment& s : plan.segs) {
    if (s.monitor) {
      if (!c->monitor) continue;
      if (c->xr.mode == XR_NCCL) return fail(c, HPS_ERR_ARG, "monitors are not supported with the NCCL all-gather exchange (HPSIM_XR=nccl)");
      launch_monitor(l, v, b->st, mon);
      c->launches++;
      continue;
    }
    const int dirty = dirty_mask(plan.ops.data() + s.op_begin, s.op_count);
    // (a kick right behind a monitor has no ops of its own to carry the pilot through: stand-alone chain)
    const bool fm = fused_mom && s.kick_follows && (s.fin_mode & 3) == FIN_KICK && s.op_count > 0;
    const int fin_mode = fm ? ((s.fin_mode & ~3) | FIN_KICK_FUSED) : s.fin_mode;
    prof_mark(c, HPS_PROF_SEGMENT);
    launch_segment(l, v, plan.ops.data() + s.op_begin, d_ops + s.op_begin, d_opc + s.op_begin, (int)s.op_count, fin_mode, dirty, c->compat, b->st, sc ? sc->dev : nullptr,
                   sc ? sc->field : nullptr, s.kp, sc ? sc->bins : nullptr, b->defer);
    c->launches++;
    sweeps++;
    if (s.fin_mode != FIN_NONE && !fm) {
      int rc = xr_after_reduce(c, l, FK_SEGMENT, s.fin_mode, 0, nullptr, nullptr, v, b->st, sc ? sc->dev : nullptr, nullptr);
      if (rc) return rc;
    }
    if (!s.kick_follows) { extra++; continue; }
    int rc = HPS_OK;
    if (c->fused && c->xr.mode != XR_NCCL && sc->fused_cap > 0 && sc->mesh_w == 0.0) {
      // the whole chain in one resident grid: every block with particles, and enough blocks for the field items
      const unsigned blk = (unsigned)scheff_block();
      const unsigned want = std::max((((unsigned)b->n + 1) / 2 + blk - 1) / blk, scheff_items(s.kp));
      const unsigned grid = std::max(1u, std::min(want, sc->fused_cap));
      prof_mark(c, HPS_PROF_MOMENTS);
      cudaError_t e = launch_scheff(l, grid, v, b->st, sc->dev, sc->bins, sc->field, sc->field_part, sc->field_tickets, s.kp, ++c->mesh_seq);
      if (e != cudaSuccess) return fail(c, HPS_ERR_CUDA, std::string("kick launch: ") + cudaGetErrorString(e));
      c->launches++;
      continue;
    }
    const int merged = c->xr.mode != XR_NCCL;  // K2 rides inside K3 (grid barrier) unless its finaliser is a separate kernel
    if (!merged) {  // (XR_NCCL never runs the fused epilogue)
      prof_mark(c, HPS_PROF_LLOSS);
      launch_lloss_pass(l, v, b->st, sc->dev);
      c->launches++;
      rc = xr_after_reduce(c, l, FK_LLOSS, 0, 0, nullptr, nullptr, v, b->st, sc->dev, nullptr);
      if (rc) return rc;
    }
    prof_mark(c, HPS_PROF_MOMENTS);
    launch_moments(l, v, b->st, sc->dev, sc->bins, s.kp, 1, merged, fm ? 1 : 0);  // fm: does nothing unless the epilogue asked
    rc = xr_after_reduce(c, l, FK_MOMENTS, 0, 1, nullptr, &s.kp, v, b->st, sc->dev, sc->bins);
    if (rc) return rc;
    prof_mark(c, HPS_PROF_DEPOSIT);
    cudaError_t e = launch_deposit(l, v, b->st, sc->dev, sc->bins, s.kp);
    if (e == cudaSuccess && c->xr.mode == XR_NCCL) {
      // the rank-local mesh is all-gathered on every kick (the remesh decision lives on the device);
      // k_field folds the copies in rank order when it runs
      const size_t bytes = ((size_t)s.kp.nr * s.kp.nz + 2) * sizeof(double);
      if (comm_allgather(c->comm, sc->bins, c->xr.bins_gather, bytes, c->stream))
        return fail(c, HPS_ERR_COMM, "ncclAllGather of the charge mesh failed");
    }
    prof_mark(c, HPS_PROF_FIELD);
    if (e == cudaSuccess) e = launch_field(l, b->st, sc->dev, sc->bins, sc->field, sc->field_part, sc->field_tickets, s.kp, ++c->mesh_seq);
    if (e != cudaSuccess) return fail(c, HPS_ERR_CUDA, std::string("kick launch: ") + cudaGetErrorString(e));
    c->launches += 3;
  }
  *sweeps_io += sweeps;
  *extra_io += extra;
  return HPS_OK;
}

// One planned stretch of elements [first, last]: build the program, upload it, launch the chain (or replay the graph
// captured for the same program earlier).  Accumulates the sweep / launch counters of the enclosing hps_simulate call.
static int run_range(hps_ctx* c, hps_beam* b, hps_lattice* lat, hps_scheff* sc, uint32_t first, uint32_t last, uint64_t* sweeps_io,
                     uint64_t* extra_io) {
  ScHost off;
  ScHost& sh = sc ? sc->h : off;
  sh.on = sc != nullptr;
  if (sc) sh.mesh_w = sc->mesh_w;
  c->eng.monitor_on = c->monitor != nullptr;
  Plan plan = build_plan(lat->e, first, last, c->eng, b->h, sh);
  if (!plan.error.empty()) return fail(c, HPS_ERR_LATTICE, plan.error);
  const size_t nops = plan.ops.size();

  // ---- CUDA graph path: single GPU, no per-launch profiling, no drift-wise re-planning
  if (c->use_graphs && c->xr.nranks <= 1 && !c->profile && (!sc || sc->mesh_w == 0.0) && nops > 0) {
    // key material: everything the captured launches depend on -- the op records, the segment descriptors, the kick
    // parameters, the objects (generation ids, not just addresses) and the launch switches
    std::vector<unsigned char> mat;
    auto put = [&](const void* p, size_t n) { const unsigned char* q = static_cast<const unsigned char*>(p); mat.insert(mat.end(), q, q + n); };
    put(plan.ops.data(), nops * sizeof(Op));
    for (const Segment& s : plan.segs) {
      const uint64_t d[5] = {s.op_begin, s.op_count, (uint64_t)s.fin_mode, (uint64_t)s.kick_follows, (uint64_t)s.monitor};
      put(d, sizeof(d));
      if (s.kick_follows || (s.op_count && plan.ops[s.op_begin].type == OP_SC_KICK)) {
        const double kd[6] = {s.kp.freq, s.kp.mass, s.kp.current_frac, s.kp.remesh_val, s.kp.adj_bunch_w, s.kp.mesh_w};
        const uint64_t ku[4] = {s.kp.adj_bunch_cfg, s.kp.nr, s.kp.nz, s.kp.n_global};
        put(kd, sizeof(kd)); put(ku, sizeof(ku));
      }
    }
    const uint64_t env[16] = {b->gen, (uint64_t)(uintptr_t)b->slab, (uint64_t)(uintptr_t)b->st, b->n, b->offset, sc ? sc->gen : 0ull,
                              (uint64_t)(uintptr_t)(sc ? sc->dev : nullptr), (uint64_t)(uintptr_t)(sc ? sc->bins : nullptr),
                              (uint64_t)(uintptr_t)(sc ? sc->field : nullptr), (uint64_t)(uintptr_t)(sc ? sc->field_tickets : nullptr),
                              (uint64_t)c->seg_variant, (uint64_t)(c->fused | (c->fused_moments << 4) | ((c->stage & 3) << 8) | (c->patterns << 12)),
                              (uint64_t)c->compat, (uint64_t)c->pdl, (uint64_t)(uintptr_t)(c->monitor ? c->monitor->data : nullptr),
                              (uint64_t)plan.engine_reset};
    put(env, sizeof(env));
    const double bm[2] = {b->h.mass, b->h.charge};
    put(bm, sizeof(bm));
    const uint64_t key = fnv(1469598103934665603ull, mat.data(), mat.size());
    for (auto& g : c->graphs)
      if (g.key == key && g.exec && g.material == mat) {
        g.stamp = ++c->graph_stamp;
        CK(c, cudaGraphLaunch(g.exec, c->stream));
        c->launches += g.launches;
        *sweeps_io += g.sweeps;
        *extra_io += g.extra;
        return HPS_OK;
      }
    // new program: own copy of the op records, capture, instantiate
    if (c->graphs.size() >= 16) {  // evict the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < c->graphs.size(); ++i) if (c->graphs[i].stamp < c->graphs[lru].stamp) lru = i;
      CK(c, cudaStreamSynchronize(c->stream));
      if (c->graphs[lru].exec) cudaGraphExecDestroy(c->graphs[lru].exec);
      cudaFree(c->graphs[lru].d_ops); cudaFree(c->graphs[lru].d_opc);
      c->graphs.erase(c->graphs.begin() + (long)lru);
    }
    hps_ctx::GraphEntry g;
    g.key = key;
    g.material = std::move(mat);
    g.beam = b; g.scheff = sc; g.monitor = c->monitor;
    CK(c, cudaMalloc(&g.d_ops, nops * sizeof(Op)));
    CK(c, cudaMalloc(&g.d_opc, nops * sizeof(OpC)));
    CK(c, cudaMemcpy(g.d_ops, plan.ops.data(), nops * sizeof(Op), cudaMemcpyHostToDevice));
    const uint64_t l0 = c->launches;
    cudaGraph_t graph = nullptr;
    CK(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = launch_plan(c, b, sc, plan, g.d_ops, g.d_opc, &g.sweeps, &g.extra);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc == HPS_OK && e == cudaSuccess) e = cudaGraphInstantiate(&g.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc != HPS_OK || e != cudaSuccess) {
      cudaGetLastError();
      cudaFree(g.d_ops); cudaFree(g.d_opc);
      c->launches = l0;
      c->use_graphs = 0;  // fall through to direct launches from now on
    } else {
      g.launches = c->launches - l0;
      g.stamp = ++c->graph_stamp;
      c->graphs.push_back(g);
      CK(c, cudaGraphLaunch(g.exec, c->stream));
      *sweeps_io += g.sweeps;
      *extra_io += g.extra;
      return HPS_OK;
    }
  }

  // ---- direct launches
  if (nops > c->ops_cap) {
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_ops); cudaFree(c->d_opc); cudaFreeHost(c->h_ops);
    c->d_ops = nullptr; c->d_opc = nullptr; c->h_ops = nullptr;
    c->ops_cap = nops * 2;
    CK(c, cudaMalloc(&c->d_ops, c->ops_cap * sizeof(Op)));
    CK(c, cudaMalloc(&c->d_opc, c->ops_cap * sizeof(OpC)));
    CK(c, cudaHostAlloc(&c->h_ops, c->ops_cap * sizeof(Op), cudaHostAllocDefault));
  }
  CK(c, cudaStreamSynchronize(c->stream));  // h_ops may still feed the previous copy
  if (nops) {
    std::memcpy(c->h_ops, plan.ops.data(), nops * sizeof(Op));
    CK(c, cudaMemcpyAsync(c->d_ops, c->h_ops, nops * sizeof(Op), cudaMemcpyHostToDevice, c->stream));
  }
  return launch_plan(c, b, sc, plan, c->d_ops, c->d_opc, sweeps_io, extra_io);
}

int hps_simulate(hps_ctx* c, hps_beam* b, hps_lattice* lat, hps_scheff* sc, uint32_t first, uint32_t last,
                 hps_run_stats* out) {
  if (!c || !b || !lat) return fail(c, HPS_ERR_ARG, "null handle");
  if (last >= lat->e.size()) return fail(c, HPS_ERR_ARG, "element index out of range");
  if (sc && sc->mesh_w != 0.0 && c->xr.mode == XR_NCCL)
    return fail(c, HPS_ERR_ARG, "mesh_size_cutoff_w != 0 is not supported with the NCCL all-gather exchange (HPSIM_XR=nccl)");
  if (sc && !(b->h.n_global > 2)) return fail(c, HPS_ERR_ARG, "space charge needs more than 2 particles");
  if (!(b->h.freq > 0.0)) return fail(c, HPS_ERR_STATE, "beam frequency is not set");
  cudaSetDevice(c->device);

  ScDev before;
  std::memset(&before, 0, sizeof(before));
  if (sc) CK(c, cudaMemcpyAsync(&before, sc->dev, sizeof(ScDev), cudaMemcpyDeviceToHost, c->stream));

  if (c->l2_persist && b->slab && (c->l2_window_base != b->slab || c->l2_window_bytes != b->slab_bytes)) {
    cudaStreamAttrValue av;
    std::memset(&av, 0, sizeof(av));
    const size_t win = std::min(b->slab_bytes, c->l2_window_max);
    av.accessPolicyWindow.base_ptr = b->slab;
    av.accessPolicyWindow.num_bytes = win;
    av.accessPolicyWindow.hitRatio = win ? (float)std::min(1.0, (double)c->l2_persist_max / (double)win) : 0.f;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) { cudaGetLastError(); c->l2_persist = 0; }
    c->l2_window_base = b->slab;
    c->l2_window_bytes = b->slab_bytes;
  }
  if (sc) { int e = 0; while (e < 32 && (1ull << e) < (unsigned long long)b->h.n_global) ++e; sc->bin_scale = (double)(1ull << (62 - e)); }
  const uint64_t launches0 = c->launches;
  uint64_t sweeps = 0, extra = 0;
  { int rc = xr_rendezvous(c); if (rc) return rc; }
  CK(c, cudaEventRecord(c->ev0, c->stream));
  if (!sc || sc->mesh_w == 0.0) {
    int rc = run_range(c, b, lat, sc, first, last, &sweeps, &extra);
    if (rc) return rc;
  } else {
    // Energy-dependent mesh (space_charge.cpp:107-117): above the cutoff energy Scheff::Start also quadruples
    // interval_, which SimulateDrift reads when it STARTS the next drift (simulation_engine.cu:129-131) -- the number
    // of kicks of a drift depends on <W>_good of the kick before it.  The range is therefore planned in stretches
    // that begin at a drift; before each the state of the last kick is read back (one host synchronisation per drift
    // while the option is on; the mesh size itself switches on the device).
    uint32_t e0 = first;
    while (e0 <= last) {
      uint32_t e1 = e0;
      while (e1 < last && lat->e[e1 + 1].type != HPS_ELEM_DRIFT) ++e1;
      ScDev now;
      CK(c, cudaMemcpyAsync(&now, sc->dev, sizeof(ScDev), cudaMemcpyDeviceToHost, c->stream));
      CK(c, cudaStreamSynchronize(c->stream));
      if (sc->interval_user_set) { sc->interval_valid_after = now.n_kicks; sc->interval_user_set = false; }
      if (now.n_kicks > sc->interval_valid_after) sc->h.interval = now.mesh_hi ? sc->interval_init * 4.0 : sc->interval_init;
      int rc = run_range(c, b, lat, sc, e0, e1, &sweeps, &extra);
      if (rc) return rc;
      e0 = e1 + 1;
    }
  }
  if (c->monitor) {
    // online mode: Simulate() ends with Beam::UpdateStatForPlotting (simulation_engine.cpp:93-94, beam.cpp:1024-1029) --
    // the whole kick-statistics chain once more (it MARKS longitudinal losses, also with space charge off), sigma_W and
    // <phi_r> of the good particles
    if (c->xr.mode == XR_NCCL) return fail(c, HPS_ERR_ARG, "monitors are not supported with the NCCL all-gather exchange (HPSIM_XR=nccl)");
    UpdateStatForSpaceCharge(b);
    UpdateSig(b, HPS_W, HPS_STAT_W_SIG_GOOD, true);
    UpdateAvg(b, 6, HPS_STAT_PHI_AVG_R, true);
  }
  CK(c, cudaEventRecord(c->ev1, c->stream));
  CK(c, cudaGetLastError());
  prof_finish(c);
  if (sc && sc->mesh_w != 0.0) {  // interval_ as the reference's getter would report it after the run
    ScDev now;
    CK(c, cudaMemcpyAsync(&now, sc->dev, sizeof(ScDev), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (now.n_kicks > sc->interval_valid_after) sc->h.interval = now.mesh_hi ? sc->interval_init * 4.0 : sc->interval_init;
  }
  if (out) {
    std::memset(out, 0, sizeof(*out));
    ScDev after;
    std::memset(&after, 0, sizeof(after));
    if (sc) CK(c, cudaMemcpyAsync(&after, sc->dev, sizeof(ScDev), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    CK(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    out->particles = b->h.n_global;
    out->elements = last >= first ? (uint64_t)last - first + 1 : 0;
    out->kicks = after.n_kicks - before.n_kicks;
    out->remesh = after.n_remesh - before.n_remesh;
    out->stat_redo = after.n_lloss_pass - before.n_lloss_pass;
    out->sweeps = sweeps;
    // the sweep that opens the range and every sweep not followed by a kick are "extra" (P)
    out->extra_sweeps = extra;
    out->launches = c->launches - launches0;
    out->device_ms = ms;
    return xr_check(c);
  }
  return HPS_OK;
}

}  // extern "C"
