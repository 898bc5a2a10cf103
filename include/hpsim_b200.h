/*
 * hpsim_b200.h -- C ABI of the B200-native HPSim tracking + SCHEFF hot path.
 *
 * This is the drop-in boundary.  The reference has no true C ABI for this
 * path: its seams are the `extern "C"` kernel-call layer that takes C++ object
 * pointers (reference src/simulation/inc/simulation_engine_cu.h:4-25), the C++
 * launch layers (src/beam/inc/beam_cu.h:9-41, src/simulation/inc/space_charge_cu.h:6-26)
 * and, outermost, the CPython-2 extension types in src/pywrapper.  Every entry
 * point below names the reference interface it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ or torch types cross this boundary.
 *  - every function returns an int status (HPS_OK == 0) unless it returns a
 *    handle (NULL on failure) or a plain value; hps_last_error() gives the text.
 *  - units are the reference's C++ units: m, rad, MeV, MHz, A
 *    (reference src/beam/inc/beam.h:14-29).
 *  - one host thread per hps_ctx.  All device work of a context is issued on the
 *    context's own CUDA stream; calls that return host data synchronise it.
 *  - there is NO CPU fallback: every call fails with HPS_ERR_CUDA when no
 *    sm_100 device is usable.
 */
#ifndef HPSIM_B200_H
#define HPSIM_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPS_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------- */
enum {
  HPS_OK = 0,
  HPS_ERR_CUDA = 1,      /* CUDA runtime error / no device               */
  HPS_ERR_ARG = 2,       /* bad argument                                  */
  HPS_ERR_STATE = 3,     /* call order violated (e.g. restore w/o save)   */
  HPS_ERR_LATTICE = 4,   /* lattice not simulable (reference would exit)  */
  HPS_ERR_NOTFOUND = 5,  /* element name unknown (reference throws
                            BeamLineElementNotFoundException)             */
  HPS_ERR_COMM = 6       /* NCCL error                                    */
};

/* ---- lattice: flat element table ---------------------------------------- */
/* One record per beamline element, in model-index order.  Replaces the
 * std::vector<BeamLineElement*> built by GenerateBeamLine
 * (reference src/beamline/src/init.cpp:20-429) and the element classes of
 * src/beamline/inc/beamline_element.h.  Parameter slots per type: */
enum hps_elem_type {
  HPS_ELEM_DRIFT = 1,       /* p[0]=length p[1]=aperture                              */
  HPS_ELEM_QUAD = 2,        /* p[0]=length p[1]=aperture p[2]=gradient [T/m]          */
  HPS_ELEM_RFGAP_DTL = 3,   /* p[0..24] = RFGapParameter order (rfgap_parameter.h:4-32):
                               length frequency cell_length_over_beta_lambda amplitude
                               phase_ref phase_shift energy_out beta_center dg t tp sp
                               fit_beta_center fit_beta_min fit_t0..fit_t5 fit_s1..fit_s5 */
  HPS_ELEM_RFGAP_CCL = 4,   /* same slots as DTL                                      */
  HPS_ELEM_ROTATION = 5,    /* p[0]=angle                                             */
  HPS_ELEM_CAPERTURE = 6,   /* p[0]=aperture; HPS_FLAG_ON = "in"                      */
  HPS_ELEM_RAPERTURE = 7,   /* p[0]=x_left p[1]=x_right p[2]=y_top p[3]=y_bottom; ON="in" */
  HPS_ELEM_DIAGNOSTICS = 8, /* no parameters                                          */
  HPS_ELEM_BUNCHER = 9,     /* p[0]=voltage[MV] p[1]=frequency p[2]=phase p[3]=aperture; ON */
  HPS_ELEM_DIPOLE = 10,     /* p[0..8] = DipoleParameter order (dipole_parameter.h:4-15):
                               radius angle edge_angle_in edge_angle_out half_gap k1 k2
                               field_index kinetic_energy                             */
  HPS_ELEM_SPCH_COMP = 11,  /* p[0]=fraction                                          */
  HPS_ELEM_STEERER = 12     /* p[0]=bl_h p[1]=bl_v                                    */
};

#define HPS_FLAG_ON 1u       /* aperture "in" / buncher "on"  */
#define HPS_FLAG_MONITOR 2u  /* monitor column > 0            */
#define HPS_ELEM_NPARAM 26
#define HPS_NAME_LEN 32

typedef struct hps_elem {
  uint32_t type;  /* enum hps_elem_type */
  uint32_t flags; /* HPS_FLAG_*          */
  char name[HPS_NAME_LEN];
  double p[HPS_ELEM_NPARAM];
} hps_elem;

/* ---- opaque handles ------------------------------------------------------ */
typedef struct hps_ctx hps_ctx;
typedef struct hps_beam hps_beam;
typedef struct hps_lattice hps_lattice;
typedef struct hps_scheff hps_scheff;

/* ---- context ------------------------------------------------------------- */
/* Replaces HPSim.set_gpu / SetGPU (reference src/beamline/src/init.cpp:12-18).
 * `device` is the CUDA ordinal this context (and all objects made from it) uses. */
hps_ctx* hps_ctx_create(int device);
void hps_ctx_destroy(hps_ctx*);
const char* hps_last_error(const hps_ctx*); /* ctx may be NULL: last error of create */
int hps_abi_version(void);
/* 0 (default): bit-compatible launch-geometry quirks of the reference are kept
 * (H2 skipped kick indices, H7 integer charge ratio, ...).
 * There is deliberately no "fixed physics" mode in this round; the switch exists
 * so that tests can assert the quirks are the only difference. */
int hps_ctx_set_compat(hps_ctx*, int compat_reference);
/* Arithmetic of the element maps and of the space-charge kick in the fused segment kernel.
 * HPS_MATH_REFERENCE: the reference's own expression order, operation by operation
 *   (simulation_kernel.cu:16-871, space_charge_kernel.cu:573-706), every division and square root
 *   IEEE-correctly rounded -- computed with branch-free Newton sequences and, for divisions by per-op
 *   constants, one multiplication by a prepared reciprocal plus an exact-remainder correction
 *   (hpsim_b200/csrc/hps_math.cuh); bit-identical to HPS_MATH_REFERENCE_NVCC.
 * HPS_MATH_REFERENCE_NVCC: the same expressions with nvcc's own `/` and sqrt (kept as the cross-check).
 * HPS_MATH_FAST: divisions by per-op constants are plain multiplications by reciprocals and
 *   sqrt-then-divide pairs are one rsqrt; every map result is within a few ulp of the reference-order
 *   value (tests/host_maps.cu), loss decisions are identical on every fixture, but the dynamics amplify
 *   the ulp-level differences: particles that dropped out of the RF bucket can deviate by 1e-6.
 * No reference counterpart. */
enum { HPS_MATH_REFERENCE = 0, HPS_MATH_FAST = 1, HPS_MATH_REFERENCE_NVCC = 2 };
int hps_ctx_set_math(hps_ctx*, int mode);
int hps_ctx_get_math(hps_ctx*);

/* Multi-GPU particle sharding.  The reference is single-GPU; this
 * has no reference counterpart.  `nccl_unique_id` is the 128-byte ncclUniqueId
 * produced by hps_comm_unique_id() on rank 0 and distributed by the caller
 * (bench.py uses torch.distributed for that and nothing else). */
int hps_comm_unique_id(void* out128);
int hps_ctx_comm_init(hps_ctx*, int rank, int nranks, const void* nccl_unique_id);
/* 0 = single rank, 1 = in-kernel exchange over peer memory (CUDA IPC, NVLink), 2 = NCCL all-gather
 * between kernels (fallback when the peer mapping is unavailable, or HPSIM_XR=nccl). */
int hps_ctx_comm_mode(hps_ctx*);
/* Wait-time accounting of the in-kernel exchanges of a sharded beam (no reference counterpart): out[0] ns the reducing
 * kernels' last blocks spent between issuing their own stores and holding every peer's words, out[1] number of such
 * exchanges, out[2] ns block 0 of the field solve waited for the peers' charge meshes after its Green slice, out[3] count. */
int hps_ctx_comm_stats(hps_ctx*, uint64_t out[4]);

/* Pinned host staging buffers for hps_beam_upload / hps_beam_download (optional: any host pointer
 * works, pinned ones copy at full PCIe rate), and the number of kernels this context has launched
 * so far (bench.py reports the per-step delta as gpu_launches).  No reference counterpart. */
void* hps_host_alloc(size_t bytes);
void hps_host_free(void*);
uint64_t hps_ctx_launch_count(hps_ctx*);
/* Per-kernel-class device timing of hps_simulate (CUDA events between consecutive launches on the
 * context's stream).  While on, every hps_simulate call ends with a stream synchronisation.
 * ms[k] / n[k]: accumulated milliseconds and launch count of class k since set_profile(1). */
enum { HPS_PROF_SEGMENT = 0, HPS_PROF_LLOSS, HPS_PROF_MOMENTS, HPS_PROF_DEPOSIT, HPS_PROF_FIELD, HPS_PROF_NCLASS };
int hps_ctx_set_profile(hps_ctx*, int on);
int hps_ctx_get_profile(hps_ctx*, double* ms, uint64_t* n);
/* Device timing for callers (bench.py): record CUDA event `slot` (0..7) on the context's stream;
 * elapsed() synchronises on event b and returns the milliseconds between two recorded slots.
 * hps_ctx_flush_l2 overwrites a scratch buffer larger than the 126 MB L2 on the same stream. */
int hps_ctx_event_record(hps_ctx*, int slot);
int hps_ctx_event_elapsed(hps_ctx*, int slot_a, int slot_b, double* ms);
int hps_ctx_flush_l2(hps_ctx*);

/* ---- beam ---------------------------------------------------------------- */
/* Replaces Beam::Beam(num, mass, charge, current) + AllocateBeam
 * (reference src/beam/src/beam.cpp:70-147).  With sharding, `n_global` is the
 * whole beam and this rank owns n_local particles, interleaved: local particle j
 * is global particle offset + j * stride (offset = rank, stride = number of
 * ranks; 0 and 1 for an unsharded beam).  Upload / download move the local
 * particles in local order. */
hps_beam* hps_beam_create(hps_ctx*, uint64_t n_global, double mass, double charge, double current);
void hps_beam_destroy(hps_beam*);
int hps_beam_local_range(const hps_beam*, uint64_t* offset, uint64_t* n_local);
uint64_t hps_beam_local_stride(const hps_beam*);

/* Replaces UpdateBeamOnDevice + Beam::InitBeamFromDistribution
 * (reference src/beam/src/beam.cu:197-218, beam.cpp:317-369): host arrays of
 * n_local entries each; loss / lloss may be NULL (=> zero).  Particle 0 of the
 * global beam is the reference particle: its x,xp,y,yp are forced to 0 and its
 * loss flags cleared exactly as the reference does.  Finishes with
 * InitPhiAvgGood (beam.cpp:655-661). */
int hps_beam_upload(hps_beam*, const double* x, const double* xp, const double* y,
                    const double* yp, const double* phi, const double* w,
                    const uint32_t* loss, const uint32_t* lloss);
/* Replaces CopyBeamFromDevice / Beam::GetX... (beam.cu:228-245, beam.cpp:1124-1231).
 * Any pointer may be NULL.  Calls UpdateLoss first like the Python getters do
 * (reference src/pywrapper/src/wrap_beam.cpp:456). */
int hps_beam_download(hps_beam*, double* x, double* xp, double* y, double* yp,
                      double* phi, double* w, uint32_t* loss, uint32_t* lloss);
/* Beam::GetRelativePhi after UpdateRelativePhi(use_good) (beam.cpp:1071-1074). */
int hps_beam_download_phi_relative(hps_beam*, int use_good, double* phi_r);

/* Host generators, bit-identical to the reference's glibc rand() rejection
 * samplers: Beam::InitDCBeam (beam.cpp:481-537) and Beam::InitWaterbagBeam
 * (beam.cpp:391-463).  Only valid on an unsharded beam or on all ranks with the
 * same seed (every rank generates the full sequence and keeps its block). */
int hps_beam_set_dc(hps_beam*, double ax, double bx, double ex, double ay, double by,
                    double ey, double dphi, double sync_phi, double sync_w, uint32_t seed);
int hps_beam_set_waterbag(hps_beam*, double ax, double bx, double ex, double ay, double by,
                          double ey, double az, double bz, double ez, double sync_phi,
                          double sync_w, double freq, uint32_t seed);

/* Save/Restore{Initial,Intermediate}Beam (beam.cpp:547-647).  slot 0 = initial,
 * 1 = intermediate.  Copies live in device memory (the reference keeps host
 * copies); semantics, including the statistics refreshed on restore, match. */
int hps_beam_save(hps_beam*, int slot);
int hps_beam_restore(hps_beam*, int slot);

/* SetRefEnergy / SetRefPhase / freq member (beam.cpp:1051-1066; wrap_beam.cpp set_frequency). */
int hps_beam_set_ref_w(hps_beam*, double w);
int hps_beam_set_ref_phi(hps_beam*, double phi);
int hps_beam_set_frequency(hps_beam*, double mhz);
double hps_beam_get_ref_w(hps_beam*);
double hps_beam_get_ref_phi(hps_beam*);
double hps_beam_get_frequency(const hps_beam*);
double hps_beam_get_mass(const hps_beam*);
double hps_beam_get_charge(const hps_beam*);
double hps_beam_get_current(const hps_beam*);
uint64_t hps_beam_get_size(const hps_beam*);

/* Beam::Shift{X,Xp,Y,Yp,Phi,W} (beam.cpp:1595-1642) -- coord in "x","xp","y","yp","phi","w".
 * Beam::ApplyCut (beam.cpp:1574-1588) -- coord char 'x','y','w','p'. */
enum { HPS_X = 0, HPS_XP = 1, HPS_Y = 2, HPS_YP = 3, HPS_PHI = 4, HPS_W = 5 };
int hps_beam_shift(hps_beam*, int coord, double value);
int hps_beam_cut(hps_beam*, int coord, double lo, double hi);

/* Statistics.  hps_beam_update(which) performs the reference's Update* call,
 * hps_beam_get(which) its Get* (blocking scalar read).  The ids mirror the 33
 * device scalars of Beam (beam.h:200-262). */
enum hps_stat_id {
  HPS_STAT_X_AVG = 0, HPS_STAT_X_AVG_GOOD, HPS_STAT_XP_AVG, HPS_STAT_Y_AVG,
  HPS_STAT_Y_AVG_GOOD, HPS_STAT_YP_AVG, HPS_STAT_PHI_AVG, HPS_STAT_PHI_AVG_R,
  HPS_STAT_PHI_AVG_GOOD, HPS_STAT_W_AVG, HPS_STAT_W_AVG_GOOD, HPS_STAT_X_SIG,
  HPS_STAT_X_SIG_GOOD, HPS_STAT_XP_SIG, HPS_STAT_Y_SIG, HPS_STAT_Y_SIG_GOOD,
  HPS_STAT_YP_SIG, HPS_STAT_PHI_SIG, HPS_STAT_PHI_SIG_R, HPS_STAT_PHI_SIG_GOOD,
  HPS_STAT_W_SIG, HPS_STAT_W_SIG_GOOD, HPS_STAT_X_EMIT, HPS_STAT_Y_EMIT,
  HPS_STAT_Z_EMIT, HPS_STAT_R_MAX, HPS_STAT_ABS_PHI_MAX, HPS_STAT_NUM_LOSS,
  HPS_STAT_NUM_LLOSS, HPS_STAT_NUM_GOOD, HPS_STAT_COUNT
};
enum hps_update_id {
  HPS_UPD_LOSS = 0,          /* Beam::UpdateLoss                       beam.cpp:754  */
  HPS_UPD_LLOSS,             /* Beam::UpdateLongitudinalLoss           beam.cpp:776  */
  HPS_UPD_GOOD_COUNT,        /* Beam::UpdateGoodParticleCount          beam.cpp:788  */
  HPS_UPD_AVG_X, HPS_UPD_AVG_X_GOOD, HPS_UPD_AVG_XP, HPS_UPD_AVG_Y, HPS_UPD_AVG_Y_GOOD,
  HPS_UPD_AVG_YP, HPS_UPD_AVG_PHI, HPS_UPD_AVG_PHI_GOOD, HPS_UPD_AVG_PHI_R,
  HPS_UPD_AVG_PHI_R_GOOD, HPS_UPD_AVG_W, HPS_UPD_AVG_W_GOOD,
  HPS_UPD_SIG_X, HPS_UPD_SIG_X_GOOD, HPS_UPD_SIG_XP, HPS_UPD_SIG_Y, HPS_UPD_SIG_Y_GOOD,
  HPS_UPD_SIG_YP, HPS_UPD_SIG_PHI, HPS_UPD_SIG_PHI_R, HPS_UPD_SIG_PHI_R_GOOD,
  HPS_UPD_SIG_W, HPS_UPD_SIG_W_GOOD,
  HPS_UPD_EMITTANCE,         /* Beam::UpdateEmittance                  beam.cpp:981  */
  HPS_UPD_STAT_FOR_SPCH,     /* Beam::UpdateStatForSpaceCharge         beam.cpp:1003 */
  HPS_UPD_AVG_XY,            /* Beam::UpdateAvgXY                      beam.cpp:1036 */
  HPS_UPD_MAX_R_PHI,         /* Beam::UpdateMaxRPhi                    beam.cpp:1048 */
  HPS_UPD_REL_PHI,           /* Beam::UpdateRelativePhi(false)         beam.cpp:1071 */
  HPS_UPD_REL_PHI_GOOD,      /* Beam::UpdateRelativePhi(true)                        */
  HPS_UPD_COUNT
};
int hps_beam_update(hps_beam*, int update_id);
int hps_beam_get(hps_beam*, int stat_id, double* out);

/* ---- lattice ------------------------------------------------------------- */
/* Replaces BeamLine + AddElement (reference src/beamline/inc/beamline.h:11-34). */
hps_lattice* hps_lattice_create(hps_ctx*, const hps_elem* elems, uint32_t n);
void hps_lattice_destroy(hps_lattice*);
uint32_t hps_lattice_size(const hps_lattice*);
/* BeamLine::GetElementModelIndex (beamline.cpp:70-77): case-insensitive; -1 if absent. */
int64_t hps_lattice_index(const hps_lattice*, const char* name);
int hps_lattice_get_elem(const hps_lattice*, uint32_t idx, hps_elem* out);
/* Element setters reached through set_db_epics/set_db_model
 * (reference src/server/src/pv_observer.cpp:92-364). */
int hps_lattice_update_elem(hps_lattice*, uint32_t idx, const hps_elem* in);

/* ---- SCHEFF -------------------------------------------------------------- */
/* Replaces Scheff::Scheff(nr, nz, adj_bunch) and the SpaceCharge setters
 * (reference src/simulation/src/space_charge.cpp:32-41, inc/space_charge.h:13-125). */
hps_scheff* hps_scheff_create(hps_ctx*, uint32_t nr, uint32_t nz, int adj_bunch);
void hps_scheff_destroy(hps_scheff*);
enum hps_sc_param {
  HPS_SC_INTERVAL = 0,        /* SetInterval (also resets interval_init_)   */
  HPS_SC_ADJ_BUNCH,           /* SetAdjBunch                                 */
  HPS_SC_ADJ_BUNCH_CUTOFF_W,  /* SetAdjBunchCutoffW                          */
  HPS_SC_MESH_CUTOFF_W,       /* SetMeshSizeCutoffW                          */
  HPS_SC_REMESH_THRESHOLD,    /* SetRemeshThreshold                          */
  HPS_SC_FRACTION,            /* SetFraction                                 */
  HPS_SC_NR,                  /* SetMeshSize / GetNx                         */
  HPS_SC_NZ                   /* SetMeshSize / GetNz                         */
};
int hps_scheff_set(hps_scheff*, int param, double value);
int hps_scheff_get(hps_scheff*, int param, double* value);
int hps_scheff_set_mesh_size(hps_scheff*, uint32_t nr, uint32_t nz);
/* Debug/parity taps (the reference's -D_DEBUG table dumps, space_charge.cu:246-304).
 * bins: nr*nz doubles; field: (nr+1)*(nz+1)*2 doubles (Er,Ez interleaved). */
int hps_scheff_download_tables(hps_scheff*, double* bins, double* field, double* dr_dz);
/* SM-clock cycles block 0 of the fused space-charge chain (k_scheff) spent per phase, summed over all
 * kicks so far: 0 moment sweep, 1 reduction tail + policy + grid barrier, 2 deposit + flush, 3 Green
 * slice, 4 grid barrier (+ mesh push of a sharded beam), 5 field solve; 6 unused.
 * out[7] is a COUNT, not cycles: particles the charge deposit dropped because their axial bin index came out as nz + 1
 * (zz / dzg rounds up to nz exactly) -- there the reference writes one row past its table
 * (src/simulation/inc/space_charge_kernel.cu:199-216); the one documented deviation of the deposit. */
int hps_scheff_phase_cycles(hps_scheff*, uint64_t out[8]);

/* ---- online monitor outputs ------------------------------------------------
 * Replaces PlotData (reference src/util/inc/plot_data.h:43-66) and Update2DPlotData (src/simulation/src/
 * simulation_engine.cu:34-82), i.e. SimulationEngine::InitEngine(beam, bl, spch, graph_on = true, plot_data): with a
 * monitor attached, every quadrupole (between its halves, ahead of the kick; simulation_engine.cu:153-154) and every
 * diagnostics element (:331-335) whose HPS_FLAG_MONITOR is set refreshes the loss count and the emittances
 * (Beam::UpdateLoss, Beam::UpdateEmittance -- with their side effects on the stale counters the kick chain reads) and
 * appends one row of 17 values.  Rows restart at 0 whenever the engine is reset (monitor_num, simulation_engine.cu:105).
 * Like the reference's online mode, hps_simulate then also ends with Beam::UpdateStatForPlotting
 * (simulation_engine.cpp:93-94, beam.cpp:1024-1029), which marks longitudinal losses even with space charge off.
 * Everything happens on the device inside hps_simulate; hps_monitor_download is the only synchronising call. */
typedef struct hps_monitor hps_monitor;
enum hps_monitor_field {           /* members of PlotData, in this order */
  HPS_MON_XAVG = 0, HPS_MON_XSIG, HPS_MON_XPAVG, HPS_MON_XPSIG, HPS_MON_XEMIT,
  HPS_MON_YAVG, HPS_MON_YSIG, HPS_MON_YPAVG, HPS_MON_YPSIG, HPS_MON_YEMIT,
  HPS_MON_PHIAVG, HPS_MON_PHISIG, HPS_MON_WAVG /* <W> - W_ref */, HPS_MON_WSIG, HPS_MON_ZEMIT,
  HPS_MON_LOSS_RATIO, HPS_MON_LOSS_LOCAL /* log10(local loss fraction + 1e-6) */,
  HPS_MON_NFIELD
};
hps_monitor* hps_monitor_create(hps_ctx*, uint32_t capacity);      /* PlotData(int)                          */
void hps_monitor_destroy(hps_monitor*);
int hps_ctx_set_monitor(hps_ctx*, hps_monitor*);                    /* NULL: graphics_on = false              */
/* out: [HPS_MON_NFIELD][capacity] doubles (PlotDataMember::GetValue of every member); *count: rows written */
int hps_monitor_download(hps_monitor*, double* out, uint32_t* count);

/* Diagnostics of the fused moment chain (no reference counterpart): the reference's Beam::UpdateStatForSpaceCharge
 * (src/beam/src/beam.cpp:1003-1016) re-reads the particles for every link of its chain; here the sweep in front of a
 * kick sums the moments about a predicted <phi>_good and re-centres exactly.  out[0] kicks finished that way, out[1]
 * kicks handed back to the stand-alone moment kernel, out[2] particles that had to be resolved one by one with the
 * true <phi>_good; out[3..7]: timeline of those sweeps as their finalising block sees it, %globaltimer ns summed over
 * out[8] launches -- 3 first block's start -> own loop end, 4 -> block partials folded, 5 -> finaliser done (moments,
 * exchanges of a sharded beam, Scheff policy), 6 start skew of that block, 7 pilot particle of block 0; out[9..11]:
 * hand-backs by cause (side list overflowed, marking threshold left its band, wrap margin exceeded); out[12..17]: the
 * finaliser's own phases in ns (side-list pass 1, <phi>_good, marking pass, relative-phase pass, finalisers + policy, mesh zeroing);
 * *pivot_err_max: largest |<phi>_good - prediction| [rad]; *pivot_margin: wrap margin in force [rad]. */
int hps_beam_moment_stats(hps_beam*, uint64_t out[18], double* pivot_err_max, double* pivot_margin);

/* ---- simulate ------------------------------------------------------------ */
typedef struct hps_run_stats {
  uint64_t particles;        /* n_global                                             */
  uint64_t elements;         /* idx_last - idx_first + 1                             */
  uint64_t kicks;            /* K: space-charge kicks executed                       */
  uint64_t remesh;           /* R: kicks that re-meshed (deposit + field solve)      */
  uint64_t sweeps;           /* segment sweeps launched (K + P)                      */
  uint64_t extra_sweeps;     /* P: sweeps not attached to a kick                     */
  uint64_t stat_redo;        /* kicks whose moment pass had to be redone (new lloss) */
  uint64_t launches;         /* kernels launched inside the call                     */
  double device_ms;          /* CUDA-event time of the whole call on the ctx stream  */
} hps_run_stats;

/* Replaces SimulationEngine::InitEngine + Simulate(start,end)
 * (reference src/simulation/src/simulation_engine.cpp:34-95) and everything
 * below it.  `sc` may be NULL (space charge off).  first/last are model indices
 * (inclusive).  Engine state that the reference keeps in file statics
 * (simulation_engine.cu:19-33) lives in the ctx; hps_engine_reset is Reset()
 * (simulation_engine.cu:101-111) and is applied automatically when
 * first < previous last, like simulation_engine.cpp:81-82. */
int hps_simulate(hps_ctx*, hps_beam*, hps_lattice*, hps_scheff*, uint32_t first,
                 uint32_t last, hps_run_stats* out);
int hps_engine_reset(hps_ctx*, hps_beam*);
int hps_ctx_synchronize(hps_ctx*);

#ifdef __cplusplus
}
#endif
#endif /* HPSIM_B200_H */
