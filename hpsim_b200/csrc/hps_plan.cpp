// hps_plan.cpp -- see hps_plan.h.  Restates the host logic of
// simulation/src/simulation_engine.cu:101-373 and simulation_engine.cpp:73-95 as a planner.
#include "hps_plan.h"

#include <cmath>
#include <cstring>

namespace hps {
namespace {

const double CLIGHT = 299.792458;

struct Builder {
  Plan plan;
  Segment cur;
  BeamHost* beam;
  ScHost* sc;

  void begin() {
    cur = Segment();
    cur.op_begin = (u32)plan.ops.size();
  }
  KickParams kick_params() const {
    KickParams kp;
    kp.freq = beam->freq;
    kp.mass = beam->mass;
    kp.current_frac = beam->current * sc->fraction;
    kp.remesh_val = sc->remesh_val;
    kp.adj_bunch_w = sc->adj_bunch_w;
    kp.adj_bunch_cfg = sc->adj_bunch;
    kp.nr = sc->nr;
    kp.nz = sc->nz;
    kp.n_global = beam->n_global;
    kp.mesh_w = sc->mesh_w;
    return kp;
  }
  void close(int fin, bool kick) {
    cur.op_count = (u32)plan.ops.size() - cur.op_begin;
    cur.fin_mode |= fin;
    cur.kick_follows = kick;
    if (kick) cur.kp = kick_params();
    if (cur.op_count > 0 || kick || (fin & 3) != FIN_NONE) plan.segs.push_back(cur);
    begin();
  }
  Op& emit(int type, u32 elem, int flag = 0) {
    if (plan.ops.size() - cur.op_begin >= HPS_MAX_SEG_OPS) close(FIN_NONE, false);
    Op op;
    std::memset(&op, 0, sizeof(op));
    op.type = type; op.elem = elem; op.flag = flag;
    plan.ops.push_back(op);
    return plan.ops.back();
  }
  // Update2DPlotData(): everything before it must have been applied, and the kick statistics that follow must see the
  // loss count and centroid it refreshes -- so the sweep in front of it ends without a reduction of its own
  void monitor() {
    close(FIN_NONE, false);
    Segment m;
    m.op_begin = (u32)plan.ops.size();
    m.monitor = true;
    plan.segs.push_back(m);
    begin();
  }
  // Scheff::Start(beam, length) in the middle of an element: the sweep so far ends with the first
  // reduction of the moment chain, the next sweep starts with the kick itself.
  void kick(double length, u32 elem) {
    close(FIN_KICK, true);
    plan.n_kicks++;
    Op& op = emit(OP_SC_KICK, elem);
    op.p[0] = length;
    op.p[1] = beam->current * sc->fraction;
    op.p[2] = beam->freq;
    cur.kp = kick_params();
  }
};

bool is_type(const std::vector<hps_elem>& lat, long i, u32 t) {
  return i >= 0 && i < (long)lat.size() && lat[i].type == t;
}

}  // namespace

Plan build_plan(const std::vector<hps_elem>& lat, u32 first, u32 last, EngineState& eng, BeamHost& beam, ScHost& sc) {
  Builder b;
  b.beam = &beam;
  b.sc = &sc;
  b.begin();

  // simulation_engine.cpp:81-83 + Reset() (simulation_engine.cu:101-111)
  if (first < eng.prev_end) {
    b.plan.engine_reset = true;
    eng.ccl_cell_num = 0;
    eng.design_w = beam.design_w;
    eng.wavelen = CLIGHT / beam.freq;
  }
  eng.prev_end = last;

  const long n = (long)lat.size();
  for (u32 id = 0; id <= last && b.plan.error.empty(); ++id) {
    const hps_elem& e = lat[id];
    if (!(id >= first || e.type == HPS_ELEM_SPCH_COMP)) continue;  // H12
    const double* p = e.p;
    switch (e.type) {
      case HPS_ELEM_DRIFT: {  // SimulateDrift :125-145
        int nk = 1;
        const double length = p[0];
        if (sc.on) nk = (int)std::ceil(length / sc.interval);
        const double hl = 0.5 * length / nk;
        for (int k = 0; k < nk; ++k) {
          for (int half = 0; half < 2; ++half) {
            if (half == 1 && sc.on) b.kick(2.0 * hl, id);
            Op& op = b.emit(OP_DRIFT, id);
            op.p[0] = hl; op.p[1] = p[1]; op.p[2] = eng.wavelen;
          }
        }
        break;
      }
      case HPS_ELEM_QUAD: {  // SimulateQuad :147-161
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && eng.monitor_on && (e.flags & HPS_FLAG_MONITOR)) b.monitor();  // :153-154
          if (half == 1 && sc.on) b.kick(p[0], id);
          Op& op = b.emit(OP_QUAD_HALF, id);
          op.p[0] = p[0]; op.p[1] = p[1]; op.p[2] = p[2]; op.p[3] = eng.wavelen;
        }
        break;
      }
      case HPS_ELEM_RFGAP_DTL:
      case HPS_ELEM_RFGAP_CCL: {
        const bool ccl = e.type == HPS_ELEM_RFGAP_CCL;
        const double gfreq = p[1], ref_phase = p[4];
        // PrepSimulateRFGap :163-175 + Beam::ChangeFrequency (beam.cpp:1658-1669)
        if (beam.freq != gfreq) {
          eng.wavelen = CLIGHT / gfreq;
          Op& fc = b.emit(OP_FREQ_CHANGE, id);
          fc.p[0] = gfreq / beam.freq;
          b.cur.fin_mode |= FIN_FREQ;
          beam.current *= gfreq / beam.freq;
          beam.freq = gfreq;
        }
        double phi_in = 0.0, phi_out = 0.0, q1 = 0.0, q2 = 0.0;
        if (!ccl) {  // SimulateDTLRFGap :182-241
          if (id == 1 || (id > 1 && !is_type(lat, (long)id - 2, HPS_ELEM_RFGAP_DTL))) {
            if ((long)id + 2 < n && is_type(lat, (long)id + 2, HPS_ELEM_RFGAP_DTL))
              phi_in = ref_phase - 0.5 * (lat[id + 2].p[4] - ref_phase);
            else if ((long)id + 1 >= n) {
              b.plan.error = "BeamLine is not set up correctly: a quad must follow a DTL gap";
              break;
            } else phi_in = ref_phase;
          } else if (id > 1) {
            phi_in = 0.5 * (lat[id - 2].p[4] + ref_phase);
          } else {
            b.plan.error = "A quad is missing before the DTL gap";
            break;
          }
          if (!is_type(lat, (long)id - 1, HPS_ELEM_QUAD) || !is_type(lat, (long)id + 1, HPS_ELEM_QUAD)) {
            b.plan.error = "a DTL gap must sit between two quads";  // reference: null dynamic_cast
            break;
          }
          q1 = lat[id - 1].p[0];
          q2 = lat[id + 1].p[0];
          phi_out = phi_in;  // the DTL second half receives phi_in again (:237)
        } else {  // SimulateCCLRFGap :243-293
          if (id == 0 || !is_type(lat, (long)id - 1, HPS_ELEM_RFGAP_CCL)) eng.ccl_cell_num = 0;
          const bool next_ccl = is_type(lat, (long)id + 1, HPS_ELEM_RFGAP_CCL);
          if (eng.ccl_cell_num == 0) {
            if (!next_ccl) { b.plan.error = "single-cell CCL tank"; break; }  // reference: bad cast
            const double nxt = lat[id + 1].p[4];
            phi_in = ref_phase - 0.5 * (nxt - ref_phase);
            phi_out = 0.5 * (nxt + ref_phase);
          } else if (!next_ccl) {
            const double prv = lat[id - 1].p[4];
            phi_in = 0.5 * (ref_phase + prv);
            phi_out = ref_phase + 0.5 * (ref_phase - prv);
          } else {
            phi_in = 0.5 * (ref_phase + lat[id - 1].p[4]);
            phi_out = 0.5 * (ref_phase + lat[id + 1].p[4]);
          }
        }
        {
          Op& op = b.emit(OP_GAP_FIRST, id, ccl ? 1 : 0);
          std::memcpy(op.p, p, 25 * sizeof(double));
          op.p[25] = eng.design_w; op.p[26] = phi_in; op.p[27] = q1; op.p[28] = q2;
        }
        eng.wavelen = CLIGHT / gfreq;  // every thread of the first-half kernel writes d_wavelen (:252)
        if (sc.on) b.kick(p[0], id);
        {
          Op& op = b.emit(OP_GAP_SECOND, id, ccl ? 1 : 0);
          std::memcpy(op.p, p, 25 * sizeof(double));
          op.p[25] = eng.design_w; op.p[26] = phi_out; op.p[27] = q1; op.p[28] = q2;
          op.p[29] = ccl ? (double)eng.ccl_cell_num : 0.0;
          op.p[30] = eng.wavelen;
        }
        if (ccl) eng.ccl_cell_num++;
        eng.design_w = p[6];
        beam.design_w = eng.design_w;
        break;
      }
      case HPS_ELEM_ROTATION: {  // :295-299
        Op& op = b.emit(OP_ROTATION, id);
        op.p[0] = p[0];
        break;
      }
      case HPS_ELEM_SPCH_COMP:  // :301-305
        if (sc.on) sc.fraction = p[0];
        break;
      case HPS_ELEM_CAPERTURE:  // :307-316 -- needs UpdateLoss + UpdateAvgXY first: a grid-wide dependency
        if (e.flags & HPS_FLAG_ON) {
          b.close(FIN_APERTURE, false);
          Op& op = b.emit(OP_CAPERTURE, id);
          op.p[0] = p[0];
        }
        break;
      case HPS_ELEM_RAPERTURE:  // :318-329
        if (e.flags & HPS_FLAG_ON) {
          b.close(FIN_APERTURE, false);
          Op& op = b.emit(OP_RAPERTURE, id);
          op.p[0] = p[0]; op.p[1] = p[1]; op.p[2] = p[2]; op.p[3] = p[3];
        }
        break;
      case HPS_ELEM_DIAGNOSTICS:  // :331-335
        if (eng.monitor_on && (e.flags & HPS_FLAG_MONITOR)) b.monitor();
        break;
      case HPS_ELEM_STEERER:  // :337-344 (H10: sign-sensitive test)
        if (p[0] > 1e-10 || p[1] > 1e-10) {
          Op& op = b.emit(OP_STEERER, id);
          op.p[0] = p[0]; op.p[1] = p[1];
        }
        break;
      case HPS_ELEM_BUNCHER:  // :346-359
        if (e.flags & HPS_FLAG_ON) {
          const double bfreq = p[1];
          eng.wavelen = CLIGHT / bfreq;  // UpdateWaveLengthKernel(bfreq); never restored (H9)
          Op& op = b.emit(OP_BUNCHER, id);
          op.p[0] = eng.design_w; op.p[1] = bfreq; op.p[2] = p[0]; op.p[3] = p[2]; op.p[4] = p[3];
          op.p[5] = bfreq / beam.freq; op.p[6] = eng.wavelen;
        }
        break;
      case HPS_ELEM_DIPOLE: {  // :361-373
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && sc.on) b.kick(p[0] * p[1], id);
          Op& op = b.emit(half == 0 ? OP_DIPOLE_FIRST : OP_DIPOLE_SECOND, id);
          std::memcpy(op.p, p, 9 * sizeof(double));
          op.p[9] = eng.wavelen;
        }
        break;
      }
      default:
        b.plan.error = "unknown element type";
    }
  }
  b.close(FIN_NONE, false);
  return b.plan;
}

}  // namespace hps
