// hps_plan.h -- host planner: lattice range -> flat program of fused segments.
//
// Everything simulation/src/simulation_engine.cu decides on the host per element (sub-step
// counts, neighbour-gap phases, quad lengths, ccl_cell_num, the design_w chain, d_wavelen
// statefulness, frequency jumps, SpchComp fractions) depends only on the lattice and the
// space-charge settings, never on particle data, so it is evaluated once per simulate() call here
// and frozen into Op records.  The kernels then run without any host synchronisation.
#ifndef HPS_PLAN_H
#define HPS_PLAN_H

#include <string>
#include <vector>
#include "hps_internal.h"

namespace hps {

struct EngineState {  // the file statics of simulation_engine.cu:19-33 that outlive a Simulate() call
  u32 prev_end = 0xffffffffu;
  double design_w = 0.0;
  double wavelen = 0.0;   // d_wavelen (simulation_kernel.cu:13)
  u32 ccl_cell_num = 0;
  bool monitor_on = false;  // param->graphics_on: monitor elements call Update2DPlotData
};

struct BeamHost {  // host-side members of Beam that the engine reads / writes
  double mass, charge, current, freq, design_w;
  u32 n_global;
};

struct ScHost {  // SpaceCharge members (space_charge.h:88-125) that are known on the host
  bool on = false;
  u32 nr = 0, nz = 0, adj_bunch = 0;
  double fraction = 1.0, interval = 0.01, adj_bunch_w = 0.0, remesh_val = 0.05, mesh_w = 0.0;
};

struct Segment {
  u32 op_begin = 0, op_count = 0;
  int fin_mode = FIN_NONE;   // FinMode | FIN_FREQ
  bool kick_follows = false; // the stat / deposit / field kernels of one kick run after this sweep
  bool monitor = false;      // no sweep: Update2DPlotData (simulation_engine.cu:34-82) at this point of the program
  KickParams kp;             // valid when kick_follows, and for a leading OP_SC_KICK
};

struct Plan {
  std::vector<Op> ops;
  std::vector<Segment> segs;
  u32 n_kicks = 0;
  bool engine_reset = false;  // this call restarted upstream of the previous end: Reset() (monitor_num = loss_num = 0)
  std::string error;  // non-empty => lattice not simulable (the reference would exit(-1) / crash)
};

// Builds the program for Simulate(first, last) (simulation_engine.cpp:73-95) and advances
// `eng`, `beam` (freq/current/design_w) and `sc.fraction` exactly as the reference's host code does.
Plan build_plan(const std::vector<hps_elem>& lat, u32 first, u32 last, EngineState& eng, BeamHost& beam, ScHost& sc);

}  // namespace hps
#endif
