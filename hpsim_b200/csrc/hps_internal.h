// hps_internal.h -- structures shared by the host planner, the C-ABI layer and the kernels.
//
// Data layout in HBM
//   particle state  : SoA, one 256-byte aligned fp64 array per coordinate (x, x', y, y', phi, W)
//                     plus two uint32 flag arrays (loss, lloss); 56 B/particle.  The relative phase
//                     array of the reference (beam/inc/beam.h:166) is NOT stored: it is recomputed
//                     from phi and the reference phase wherever the reference reads it.
//   DevStats        : the 33 one-element device scalars of the reference Beam
//                     (beam/inc/beam.h:200-262) as one struct.
//   ScDev           : SCHEFF device state: mesh cell sizes (reference `scheff_sz`,
//                     space_charge_kernel.cu:12), num_on_mesh, the remesh memory that the reference
//                     keeps on the host (space_charge.h:112-122) and the per-kick decision.
//   program         : the flat list of Op records one simulate() call executes; one "segment"
//                     (= one kernel launch = one sweep over the particles) is a contiguous slice.
#ifndef HPS_INTERNAL_H
#define HPS_INTERNAL_H

#include <stdint.h>
#include "../../include/hpsim_b200.h"

typedef uint32_t u32;

// ---- per-particle micro-ops executed by the fused segment kernel -------------------------------
enum OpType {
  OP_NONE = 0,
  OP_DRIFT,        // p: 0 half_len, 1 aperture, 2 wavelen
  OP_QUAD_HALF,    // p: 0 length, 1 aperture, 2 gradient, 3 wavelen
  OP_GAP_FIRST,    // p: 0..24 RFGapParameter, 25 design_w_in, 26 phi_in, 27 qlen1, 28 qlen2; flag = ccl
  OP_GAP_SECOND,   // p: same + 29 ccl_cell_num, 30 wavelen (d_wavelen in force); flag = ccl
  OP_DIPOLE_FIRST, // p: 0..8 DipoleParameter, 9 wavelen
  OP_DIPOLE_SECOND,
  OP_BUNCHER,      // p: 0 w_ref, 1 freq, 2 voltage, 3 phase, 4 aperture, 5 freq_ratio, 6 wavelen
  OP_ROTATION,     // p: 0 angle
  OP_STEERER,      // p: 0 blh, 1 blv
  OP_CAPERTURE,    // p: 0 aperture                     (uses stats x_avg / y_avg)
  OP_RAPERTURE,    // p: 0 xl, 1 xr, 2 yt, 3 yb
  OP_SC_KICK,      // p: 0 length, 1 current*fraction, 2 freq
  OP_FREQ_CHANGE   // p: 0 ratio  (applies to lost particles too)
};

#define HPS_OP_NPARAM 31
struct Op {
  int type;
  u32 elem;  // model index: the loss code written by aperture checks
  int flag;
  int pad;
  double p[HPS_OP_NPARAM];
};

// Particle-independent sub-expressions of an op, evaluated ONCE on the device (k_prepare_ops) with
// the same expression order the reference evaluates per particle, so the hoisting is bit-exact.
#define HPS_OPC_N 48
struct OpC {
  double c[HPS_OPC_N];
};

// what the segment kernel's last block does with its partial sums
enum FinMode {
  FIN_NONE = 0,
  FIN_KICK = 1,      // first stage of Beam::UpdateStatForSpaceCharge (good count, <phi>_good, lloss trigger)
  FIN_APERTURE = 2,  // Beam::UpdateLoss + UpdateAvgXY ahead of an "in" aperture
  FIN_KICK_FUSED = 3 // the WHOLE of Beam::UpdateStatForSpaceCharge + the Scheff::Start policy in the sweep's epilogue
                     // (moment sums about a predicted pivot phase, exact re-centring; hps_kernels.cu, "fused moments")
};
#define FIN_FREQ 4   // OR-ed in: segment contains OP_FREQ_CHANGE -> refresh phi_avg like Beam::ChangeFrequency

struct DevStats {
  double v[HPS_STAT_COUNT];  // doubles; the three counters below shadow v[NUM_*]
  u32 num_loss, num_lloss, num_good;
  u32 need_lloss_pass;       // set by FIN_KICK when some particle crosses <phi>_good + 3 pi
  double phi_r_ref;          // reference phase of the (virtual) relative-phase array
  // fused moments (FIN_KICK_FUSED): the epilogue sums the relative phase about a PREDICTED <phi>_good (a pilot particle
  // carried through the segment's maps) and re-centres exactly afterwards; particles within `pivot_m` of a decision
  // boundary (2 pi wrap of the relative phase, the <phi>_good + 3 pi longitudinal-loss threshold) are deferred to a
  // side list and resolved with the true <phi>_good.  When the prediction misses by more than pivot_m, or the list
  // overflows, moments_fallback asks the stand-alone moment kernel (k_moments) to redo the kick's statistics.
  u32 moments_fallback;
  u32 pad0;
  double pivot_m;            // wrap margin in force [rad]; adapted from the prediction errors seen
  double pivot_mc;           // marking-threshold margin in force [rad]
  unsigned long long n_fb[3];  // fallbacks by cause: list overflow, marking threshold outside its band, wrap margin exceeded
  double pivot_err_max;      // largest |<phi>_good - pivot| of the fused kicks so far (diagnostics)
  unsigned long long n_fused, n_fallback, n_deferred;  // kicks finished in the epilogue / by k_moments; deferred particles
  double mon_ref_w;          // k_monitor: energy of the reference particle, handed from its owner to the other blocks / ranks
  // timeline of the fused sweeps as the finalising block sees it (%globaltimer ns, summed over the launches; diagnostics):
  // 0 first block's start -> last block's loop end, 1 -> partials folded, 2 -> finaliser done, 3 start skew of the
  // finalising block, 4 pilot (block 0), 5 launches
  unsigned long long diag_ns[6];
  unsigned long long fin_ns[6];  // inside the finaliser: 0 pass 1 (+ block sum), 1 -> <phi>_good known (exchange 1), 2 pass 2 (marking, exchange 2),
                                 // 3 pass 3 (+ exchange 3), 4 finalisers + Scheff policy, 5 mesh zeroing
  unsigned long long diag_t0;  // earliest block start of the launch in flight (atomicMin; ~0 between launches)
};

// side list of the fused-moment epilogue: particles whose classification needs the true <phi>_good
#define HPS_DEFER_CAP 65536u
struct DeferEntry {
  double phi, x, y, w;
  unsigned long long tag;  // local particle index | flags << 32  (1 alive, 2 lloss already set, 4 lloss candidate)
};
struct DeferList {
  u32 count;   // entries appended by the sweep (may exceed HPS_DEFER_CAP: overflow)
  u32 pad;
  DeferEntry e[HPS_DEFER_CAP];
};

struct ScDev {
  double dr, dz;
  u32 num_on_mesh;      // particles this rank deposited on the mesh (k_deposit)
  u32 num_on_mesh_all;  // ... of the whole beam (k_field); what the off-mesh point-charge kick uses
  u32 good_cnt_old;
  double sigr_old, sigz_old, gm_old;
  // decision of the current kick
  int remesh;
  u32 adj_bunch;   // effective number of adjacent bunches for this kick
  double beta;
  double ratio_r, ratio_z, ratio_q, ratio_g;
  unsigned long long abs_phi_max_bits;
  unsigned long long n_kicks, n_remesh, n_lloss_pass;
  // energy-dependent mesh (SetMeshSizeCutoffW, space_charge.cpp:107-117); nr == 0 means "the configured size"
  u32 nr, nz;       // mesh size in force
  u32 mesh_hi;      // 1: <W>_good of the last kick was above the cutoff (half mesh, interval x 4)
  u32 mesh_resized; // the last kick changed the mesh size (the reference then holds fresh, zeroed tables)
  // SM-clock cycles block 0 of k_scheff spent in each phase, summed over the kicks (hps_scheff_phase_cycles):
  // 0 moments sweep, 1 reduction + finalisers + barrier, 2 deposit + flush, 3 Green slice, 4 barrier, 5 field solve
  unsigned long long phase_clk[8];
};

// static (host-known) parameters of one space-charge kick
struct KickParams {
  double freq;            // beam frequency in force [MHz]
  double mass;
  double current_frac;    // beam current * compensation fraction
  double remesh_val;
  double adj_bunch_w;
  u32 adj_bunch_cfg;
  u32 nr, nz;             // configured (= largest) mesh
  u32 n_global;
  double mesh_w;          // mesh-size cutoff energy [MeV], 0 = off
};

struct BeamView {
  double *x, *xp, *y, *yp, *phi, *w;
  u32 *loss, *lloss;
  u32 n;        // local particles
  u32 offset;   // global index of local particle 0
  u32 gstride;  // global index of local particle i = offset + i * gstride (sharded beams are interleaved)
  u32 n_global;
  u32 np2_quarter;  // NextPow2(n_global)/4: stride of the reference's kick launch (H2)
  u32 streaming;    // the beam is larger than the L2 keeps between sweeps: its traffic is marked evict-first (HPS_EVICT_FIRST)
  double mass, charge;
};

// ---- cross-rank exchange ("xr") for particle-sharded beams -------------------------
// Every rank owns one XrBox in its own HBM; all ranks map every box (CUDA IPC over NVLink/NVSwitch).
// A reduction's last block stores its rank-local partial vector into slot [rank] of EVERY peer's box,
// waits for the R slots of its own box and folds them in rank order -- so all ranks obtain bit-identical
// totals and every device-side decision stays rank-consistent.  The stores are "LL" words (the NCCL
// low-latency idea): every 8-byte word carries 32 bits of payload and the 32-bit sequence number of the
// exchange, and 8-byte stores are atomic on NVLink, so the receiver polls the data itself -- no separate
// flag, no system-scope fence, one one-way trip of latency.  Two parity buffers suffice: a rank can run
// at most one exchange ahead of its slowest peer.
#define HPS_XR_MAXR 8
#define HPS_XR_NS 24            // doubles per small-reduction slot (>= every kernel's NS)
#define HPS_XR_BINS_CAP 25608   // nr*nz + 2 (count, max|phi|) 64-bit values; 8*nr*nz <= 200 KB
struct XrBox {
  unsigned long long data[2][HPS_XR_MAXR][2 * HPS_XR_NS];       // LL words: payload | seq << 32
  unsigned long long bins[2][HPS_XR_MAXR][2 * HPS_XR_BINS_CAP]; // LL words of the charge mesh
};
enum XrMode { XR_NONE = 0, XR_P2P = 1, XR_NCCL = 2 };
struct XrView {
  int rank, nranks, mode;
  XrBox* peer[HPS_XR_MAXR];      // XR_P2P: peer[r] = rank r's box mapped here (peer[rank] = own box)
  unsigned long long* seq;       // device counter of small exchanges issued by this rank
  double* send;                  // XR_NCCL: [HPS_XR_NS] partial vector handed to ncclAllGather
  double* recv;                  //          [nranks][HPS_XR_NS] gathered vectors
  double* bins_gather;           // XR_NCCL: [nranks][nb+2] gathered charge meshes
  u32* err;                      // set non-zero when a peer did not answer in time
  unsigned long long* stats;     // [0] ns spent waiting in the small exchanges, [1] their count, [2] ns between a rank's own
                                 // mesh being pushed and the last peer's mesh having arrived (block 0 of k_field), [3] count
};

#ifndef HPS_BLOCK
#define HPS_BLOCK 256      // threads per block of the sweeps (a build-time experiment knob: -DHPS_BLOCK=320)
#endif
#define HPS_NSLOT 16          // partial-sum slots per block
#define HPS_FIELD_RS 4         // source rings per block of the field solve
#define HPS_MAX_SEG_OPS 24    // ops staged in shared memory per segment

#endif
