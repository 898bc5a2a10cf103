// hps_launch.h -- host-callable wrappers around the kernels in hps_kernels.cu.
#ifndef HPS_LAUNCH_H
#define HPS_LAUNCH_H

#include <cuda_runtime.h>
#include "hps_internal.h"

namespace hps {

struct LaunchCfg {
  cudaStream_t stream;
  int seg_variant;        // see seg_kernel() in hps_kernels.cu; 0 = default (2 particles/thread, fast maps)
  unsigned grid_seg;      // grid of the fused segment sweep (variant in use)
  unsigned grid_seg1;     // grid of the full-op-set variant (1 particle per thread)
  unsigned grid_pairs;    // grid of the 2-particles-per-thread statistic sweeps
  unsigned grid_single;   // grid of the 1-particle-per-thread API statistics kernel
  unsigned grid_deposit;
  double* partials;       // [max grid][HPS_NSLOT]
  u32* ticket;            // [0] reductions, [1] deposit hand-off (sharded beams)
  int pdl;                // launch the kick-chain kernels with programmatic stream serialization
  int stage;              // the sweep stages each thread's next particle pair in shared memory with cp.async
  int patterns;           // use the pattern sweeps (op sequence compiled in) where the launch matches one
  int num_sms;
  XrView xr;              // cross-rank exchange of a particle-sharded beam (mode XR_NONE otherwise)
};

enum StatKind {
  KS_COUNT_LOSS = 0, KS_COUNT_LLOSS, KS_COUNT_GOOD, KS_MARK_LLOSS, KS_AVG, KS_SIG, KS_AVG_XY, KS_EMIT_T, KS_EMIT_Z,
  KS_MAX_R_PHI
};
struct StatReq {
  int kind;
  int var_a, var_b;   // 0..5 = x xp y yp phi w, 6 = relative phase; var_b = -1 when unused
  int good_only;
  int out[5];
  double w_ref;       // w[0] for the emittance normalisation (beam_kernel.cu:1345)
};

// XR_NCCL: what k_xr_finalize has to finish after the all-gather of a reducing kernel's partials
enum FinKind { FK_SEGMENT = 0, FK_LLOSS, FK_MOMENTS, FK_STAT };
struct FinReq {
  int kind, fin_mode, do_policy, ns;
  unsigned max_mask;
  StatReq rq;
  KickParams kp;
};
void launch_xr_finalize(const LaunchCfg&, const FinReq&, const BeamView&, DevStats*, ScDev*, double* bins);
int reduce_slots(int kind);
unsigned reduce_max_mask(int kind);

void launch_prepare_ops(const LaunchCfg&, const Op* ops, OpC* opcs, int nops, double mass, double charge);
void launch_segment(const LaunchCfg&, const BeamView&, const Op* host_ops, const Op* ops, const OpC* opcs, int nops,
                    int fin_mode, int dirty, int compat, DevStats*, ScDev*, const double2* field, const KickParams&,
                    double* bins, DeferList* defer);
void launch_lloss_pass(const LaunchCfg&, const BeamView&, DevStats*, ScDev*);
void launch_moments(const LaunchCfg&, const BeamView&, DevStats*, ScDev*, double* bins, const KickParams&, int do_policy,
                    int with_lloss, int mode);  // mode 1: only when the fused epilogue of the sweep asked for it
int moments_occupancy();
cudaError_t prepare_chain_attributes(const KickParams&);
cudaError_t launch_deposit(const LaunchCfg&, const BeamView&, DevStats*, ScDev*, double* bins, const KickParams&);
cudaError_t launch_field(const LaunchCfg&, DevStats*, ScDev*, const double* bins, double2* field, double2* part,
                         u32* tickets, const KickParams&, u32 mesh_seq);
unsigned field_chunks(const KickParams&);
// fused moments + deposit + field solve (k_scheff); grid must not exceed scheff_occupancy() x SM count
int scheff_block();  // threads per block of k_scheff (1024; HPSIM_SCHEFF_BLOCK=512 for the 2-blocks-per-SM form)
bool scheff_supported(const KickParams&);
int scheff_occupancy(const KickParams&);
unsigned scheff_items(const KickParams&);
size_t scheff_smem(const KickParams&);
cudaError_t launch_scheff(const LaunchCfg&, unsigned grid, const BeamView&, DevStats*, ScDev*, double* bins, double2* field,
                          double2* part, u32* tickets, const KickParams&, u32 mesh_seq);
void launch_stat(const LaunchCfg&, const BeamView&, DevStats*, const StatReq&);
// Update2DPlotData on the device: loss count + the three emittance passes in one sweep, one row of the monitor table
struct MonView {
  double* data;    // [HPS_MON_NFIELD][cap]
  u32* counters;   // [0] monitor_num, [1] loss_num (simulation_engine.cu:32-33)
  u32 cap;
};
void launch_monitor(const LaunchCfg&, const BeamView&, DevStats*, const MonView&);
void launch_set_phi_ref(const LaunchCfg&, const BeamView&, DevStats*, int use_good, double phi0);
void launch_copy_stat(const LaunchCfg&, DevStats*, int dst, int src);
void launch_shift(const LaunchCfg&, double* a, double v, u32 n, u32 offset);
void launch_cut(const LaunchCfg&, const double* a, u32* loss, double lo, double hi, u32 n, u32 offset);
void launch_rel_phi_out(const LaunchCfg&, const BeamView&, const DevStats*, double* out);
int segment_occupancy(int variant);
int segment_full_variant(int variant);  // the full-op-set kernel that goes with a variant
bool segment_pairs(int variant);        // variant processes 2 particles per thread
size_t deposit_smem(const KickParams&);
size_t field_smem(const KickParams&, int threads);
int field_threads(const KickParams&);

}  // namespace hps
#endif
