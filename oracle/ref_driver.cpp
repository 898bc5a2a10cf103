// ref_driver.cpp -- Oracle A: drives the UNMODIFIED reference CUDA implementation.
//
// TEST INFRASTRUCTURE ONLY.  This file is ours; everything it calls is the
// reference's own code, compiled where it lies under /root/reference by
// oracle/Makefile into oracle/_ref/ (never copied into this repository).  It
// follows the canonical C++ driver sequence of the reference's only main()
// (reference src/test/online_2d_graphics.cpp:265-296) minus sqlite / GL / EPICS:
//   BeamLine::AddElement + setters   <- flat lattice file (hpsim_b200/lattice.py)
//   Beam(num, mass, charge, current) + InitDCBeam / InitBeamFromDistribution
//   Scheff(nr, nz, adj_bunch) + setters, SimulationEngine::InitEngine, Simulate
// and dumps the eight particle arrays plus the beam's moment scalars.
//
// usage: ref_driver <config.txt>      (key value per line, see parse below)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <chrono>
#include <cuda_runtime_api.h>

#include "beam.h"
#include "beamline.h"
#include "beamline_element.h"
#include "space_charge.h"
#include "simulation_engine.h"
#include "plot_data.h"

namespace {

struct ElemRec {  // must match hps_elem in include/hpsim_b200.h
  uint32_t type, flags;
  char name[32];
  double p[26];
};

std::map<std::string, std::string> ReadConfig(const char* path) {
  std::map<std::string, std::string> cfg;
  std::ifstream in(path);
  if (!in) { std::cerr << "ref_driver: cannot open " << path << std::endl; exit(2); }
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string k, v;
    if (!(ss >> k)) continue;
    if (k[0] == '#') continue;
    std::getline(ss, v);
    size_t b = v.find_first_not_of(" \t");
    cfg[k] = (b == std::string::npos) ? "" : v.substr(b);
  }
  return cfg;
}

double D(const std::map<std::string, std::string>& c, const char* k, double dflt) {
  auto it = c.find(k);
  return it == c.end() ? dflt : atof(it->second.c_str());
}
std::string S(const std::map<std::string, std::string>& c, const char* k, const char* dflt) {
  auto it = c.find(k);
  return it == c.end() ? std::string(dflt) : it->second;
}

void BuildBeamLine(BeamLine& bl, const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { std::cerr << "ref_driver: cannot open lattice " << path << std::endl; exit(2); }
  char magic[4]; uint32_t ver, n;
  if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "HPSL", 4) || fread(&ver, 4, 1, f) != 1 ||
      fread(&n, 4, 1, f) != 1) { std::cerr << "bad lattice header" << std::endl; exit(2); }
  std::vector<ElemRec> recs(n);
  if (fread(recs.data(), sizeof(ElemRec), n, f) != n) { std::cerr << "short lattice" << std::endl; exit(2); }
  fclose(f);
  for (const ElemRec& r : recs) {
    std::string name(r.name);
    const double* p = r.p;
    switch (r.type) {
      case 1: { Drift* e = new Drift(name); e->SetLength(p[0]); e->SetAperture(p[1]); bl.AddElement(e); break; }
      case 2: { Quad* e = new Quad(name); e->SetGradient(p[2]); e->SetLength(p[0]); e->SetAperture(p[1]);
                if (r.flags & 2u) e->SetMonitorOn(); bl.AddElement(e); break; }
      case 3: case 4: {
        RFGap* e = new RFGap(name, "RFGap-DTL");
        if (r.type == 4) e->SetType("RFGap-CCL");
        e->SetFitBetaCenter(p[12]); e->SetFitBetaMin(p[13]);
        e->SetFitT0(p[14]); e->SetFitT1(p[15]); e->SetFitT2(p[16]); e->SetFitT3(p[17]);
        e->SetFitT4(p[18]); e->SetFitT5(p[19]);
        e->SetFitS1(p[20]); e->SetFitS2(p[21]); e->SetFitS3(p[22]); e->SetFitS4(p[23]); e->SetFitS5(p[24]);
        e->SetLength(p[0]); e->SetFrequency(p[1]); e->SetCellLengthOverBetaLambda(p[2]);
        e->SetRFAmplitude(p[3]); e->SetRefPhase(p[4]); e->SetPhaseShift(p[5]); e->SetEnergyOut(p[6]);
        e->SetBetaCenter(p[7]); e->SetDg(p[8]); e->SetT(p[9]); e->SetTp(p[10]); e->SetSp(p[11]);
        bl.AddElement(e); break; }
      case 5: { Rotation* e = new Rotation(name); e->SetAngle(p[0]); bl.AddElement(e); break; }
      case 6: { ApertureCircular* e = new ApertureCircular(name); e->SetAperture(p[0]);
                if (r.flags & 1u) e->SetIn(); bl.AddElement(e); break; }
      case 7: { ApertureRectangular* e = new ApertureRectangular(name);
                e->SetApertureXLeft(p[0]); e->SetApertureXRight(p[1]); e->SetApertureYTop(p[2]);
                e->SetApertureYBottom(p[3]); if (r.flags & 1u) e->SetIn(); bl.AddElement(e); break; }
      case 8: { Diagnostics* e = new Diagnostics(name); if (r.flags & 2u) e->SetMonitorOn(); bl.AddElement(e); break; }
      case 9: { Buncher* e = new Buncher(name); e->SetVoltage(p[0]); e->SetFrequency(p[1]); e->SetPhase(p[2]);
                e->SetAperture(p[3]); if (r.flags & 1u) e->TurnOn(); bl.AddElement(e); break; }
      case 10: { Dipole* e = new Dipole(name); e->SetRadius(p[0]); e->SetAngle(p[1]); e->SetEdgeAngleIn(p[2]);
                 e->SetEdgeAngleOut(p[3]); e->SetHalfGap(p[4]); e->SetK1(p[5]); e->SetK2(p[6]);
                 e->SetFieldIndex(p[7]); e->SetKineticEnergy(p[8]); bl.AddElement(e); break; }
      case 11: { SpaceChargeCompensation* e = new SpaceChargeCompensation(name); e->SetFraction(p[0]);
                 bl.AddElement(e); break; }
      case 12: { Steerer* e = new Steerer(name); e->SetIntegratedFieldHorizontal(p[0]);
                 e->SetIntegratedFieldVertical(p[1]); bl.AddElement(e); break; }
      default: std::cerr << "unknown element type " << r.type << std::endl; exit(2);
    }
  }
}

// beam file: "HPSB" u32 version(1) u64 n, then x,xp,y,yp,phi,w (f64[n] each), loss, lloss (u32[n] each)
bool ReadBeamFile(const std::string& path, std::vector<double> v[6], std::vector<uint>& loss,
                  std::vector<uint>& lloss) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char magic[4]; uint32_t ver; uint64_t n;
  if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "HPSB", 4) || fread(&ver, 4, 1, f) != 1 ||
      fread(&n, 8, 1, f) != 1) { fclose(f); return false; }
  for (int k = 0; k < 6; ++k) { v[k].resize(n); if (fread(v[k].data(), 8, n, f) != n) { fclose(f); return false; } }
  loss.resize(n); lloss.resize(n);
  bool ok = fread(loss.data(), 4, n, f) == n && fread(lloss.data(), 4, n, f) == n;
  fclose(f);
  return ok;
}

void WriteBeamFile(const std::string& path, Beam& beam, const std::vector<double>& stats) {
  const uint64_t n = beam.num_particle;
  std::vector<double> buf(n);
  std::vector<uint> ibuf(n);
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { std::cerr << "ref_driver: cannot write " << path << std::endl; exit(2); }
  uint32_t ver = 1;
  fwrite("HPSB", 1, 4, f); fwrite(&ver, 4, 1, f); fwrite(&n, 8, 1, f);
  beam.GetX(buf.data()); fwrite(buf.data(), 8, n, f);
  beam.GetXp(buf.data()); fwrite(buf.data(), 8, n, f);
  beam.GetY(buf.data()); fwrite(buf.data(), 8, n, f);
  beam.GetYp(buf.data()); fwrite(buf.data(), 8, n, f);
  beam.GetPhi(buf.data()); fwrite(buf.data(), 8, n, f);
  beam.GetW(buf.data()); fwrite(buf.data(), 8, n, f);
  beam.GetLoss(ibuf.data()); fwrite(ibuf.data(), 4, n, f);
  beam.GetLongitudinalLoss(ibuf.data()); fwrite(ibuf.data(), 4, n, f);
  uint32_t ns = stats.size();
  fwrite(&ns, 4, 1, f);
  fwrite(stats.data(), 8, ns, f);
  fclose(f);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: ref_driver <config.txt>" << std::endl; return 2; }
  auto cfg = ReadConfig(argv[1]);
  int dev = (int)D(cfg, "device", 0);
  cudaSetDeviceFlags(cudaDeviceMapHost);
  if (cudaSetDevice(dev) != cudaSuccess) { std::cerr << "ref_driver: no CUDA device" << std::endl; return 3; }

  BeamLine bl;
  BuildBeamLine(bl, S(cfg, "lattice", "lattice.lat"));

  const uint n = (uint)D(cfg, "n", 65536);
  Beam beam(n, D(cfg, "mass", 939.294), D(cfg, "charge", -1.0), D(cfg, "current", 0.015));
  std::string beam_in = S(cfg, "beam_in", "");
  if (!beam_in.empty()) {
    std::vector<double> v[6]; std::vector<uint> loss, lloss;
    if (!ReadBeamFile(beam_in, v, loss, lloss) || v[0].size() != n) {
      std::cerr << "ref_driver: bad beam file " << beam_in << std::endl; return 2; }
    beam.InitBeamFromDistribution(v[0], v[1], v[2], v[3], v[4], v[5], &loss, &lloss);
  } else if (S(cfg, "gen", "dc") == "dc") {
    beam.InitDCBeam(D(cfg, "ax", 0.095), D(cfg, "bx", 47.0), D(cfg, "ex", 0.00327),
                    D(cfg, "ay", -0.102), D(cfg, "by", 60.0), D(cfg, "ey", 0.002514),
                    D(cfg, "dphi", 3.14159265358979323846), D(cfg, "sync_phi", 0.0),
                    D(cfg, "sync_w", 0.7518), (uint)D(cfg, "seed", 20161201));
  } else {
    beam.InitWaterbagBeam(D(cfg, "ax", 0), D(cfg, "bx", 1), D(cfg, "ex", 1), D(cfg, "ay", 0),
                          D(cfg, "by", 1), D(cfg, "ey", 1), D(cfg, "az", 0), D(cfg, "bz", 1),
                          D(cfg, "ez", 1), D(cfg, "sync_phi", 0.0), D(cfg, "sync_w", 0.75),
                          D(cfg, "gen_freq", 201.25), (uint)D(cfg, "seed", 20161201));
  }
  // wrap_beam.cpp set_frequency / set_ref_w / set_ref_phi / translate("phi")
  if (cfg.count("freq")) beam.freq = D(cfg, "freq", 201.25);
  if (cfg.count("ref_w")) beam.SetRefEnergy(D(cfg, "ref_w", 0.75));
  if (cfg.count("ref_phi")) beam.SetRefPhase(D(cfg, "ref_phi", 0.0));
  if (cfg.count("shift_phi")) beam.ShiftPhi(D(cfg, "shift_phi", 0.0));
  if (cfg.count("design_w")) beam.design_w = D(cfg, "design_w", 0.75);
  beam.SaveInitialBeam();
  std::string dump_in = S(cfg, "beam_dump_initial", "");
  if (!dump_in.empty()) WriteBeamFile(dump_in, beam, {});

  Scheff* spch = NULL;
  bool sc_on = D(cfg, "sc", 1) != 0;
  if (sc_on) {
    spch = new Scheff((uint)D(cfg, "nr", 32), (uint)D(cfg, "nz", 128), (int)D(cfg, "adj_bunch", 3));
    spch->SetInterval(D(cfg, "interval", 0.025));
    spch->SetAdjBunchCutoffW(D(cfg, "adj_bunch_cutoff_w", 0.8));
    spch->SetRemeshThreshold(D(cfg, "remesh_threshold", 0.02));
    if (cfg.count("mesh_cutoff_w")) spch->SetMeshSizeCutoffW(D(cfg, "mesh_cutoff_w", 0.0));
  }
  SimulationEngine engine;
  // online mode (graph_on + PlotData): monitor elements call Update2DPlotData (simulation_engine.cu:34-82)
  const int mon_cap = (int)D(cfg, "monitor", 0);
  // (PlotData(int) leaves xpavg / xpsig / ypavg / ypsig unallocated, plot_data.cpp:7-14; Resize allocates all 17 members)
  PlotData* plot = mon_cap > 0 ? new PlotData() : NULL;
  if (plot) plot->Resize(mon_cap);
  engine.InitEngine(&beam, &bl, spch, plot != NULL, plot);

  std::string first = S(cfg, "first", ""), last = S(cfg, "last", "");
  int repeats = (int)D(cfg, "repeats", 1);
  // optional untimed lead-in (e.g. the LEBT in front of the timed DTL range); the repeats then
  // restart from the beam saved at its end
  const bool lead_in = cfg.count("pre_first") != 0;
  if (lead_in) {
    engine.Simulate(S(cfg, "pre_first", ""), S(cfg, "pre_last", ""));
    beam.SaveIntermediateBeam();
  }
  for (int r = 0; r < repeats; ++r) {
    if (r > 0) { if (lead_in) beam.RestoreIntermediateBeam(); else beam.RestoreInitialBeam(); }
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    engine.Simulate(first, last);  // prints "Whole simulation time: X [sec]" itself
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    std::cout << "ref_driver wall: " << std::chrono::duration<double>(t1 - t0).count()
              << " [sec] run " << r << std::endl;
  }
  // optional second leg (exercises Reset / resume semantics)
  if (cfg.count("first2")) engine.Simulate(S(cfg, "first2", ""), S(cfg, "last2", ""));

  // moments exactly as the Python getters would refresh them
  beam.UpdateLoss();
  beam.UpdateLongitudinalLoss();
  std::cout.precision(17);
  std::cout << "ref_driver: after UpdateLongitudinalLoss num_good=" << beam.GetGoodParticleNum() << " phi_avg_good=" << beam.GetAvgPhi(true)
            << " num_loss=" << beam.GetLossNum() << " num_lloss=" << beam.GetLongitudinalLossNum() << std::endl;
  beam.UpdateEmittance();
  beam.UpdateAvgW(); beam.UpdateSigW();
  beam.UpdateAvgPhi(); beam.UpdateSigPhi();
  std::vector<double> stats = {
      (double)beam.GetLossNum(), (double)beam.GetLongitudinalLossNum(), (double)beam.GetGoodParticleNum(),
      beam.GetAvgX(), beam.GetSigX(), beam.GetAvgY(), beam.GetSigY(),
      beam.GetAvgPhi(), beam.GetSigPhi(), beam.GetAvgW(), beam.GetSigW(),
      beam.GetEmittanceX(), beam.GetEmittanceY(), beam.GetEmittanceZ(),
      beam.GetAvgRelativePhi(), beam.GetSigRelativePhi(), beam.GetRefEnergy(), beam.GetRefPhase(),
      beam.freq, beam.current, beam.design_w};
  WriteBeamFile(S(cfg, "beam_out", "beam_out.bin"), beam, stats);
  if (plot) {  // 17 x mon_cap doubles in the order of hps_monitor_field (include/hpsim_b200.h)
    PlotDataMember<double>* mem[17] = {&plot->xavg, &plot->xsig, &plot->xpavg, &plot->xpsig, &plot->xemit, &plot->yavg, &plot->ysig,
                                       &plot->ypavg, &plot->ypsig, &plot->yemit, &plot->phiavg, &plot->phisig, &plot->wavg, &plot->wsig,
                                       &plot->zemit, &plot->loss_ratio, &plot->loss_local};
    FILE* f = fopen(S(cfg, "monitor_out", "monitor.bin").c_str(), "wb");
    if (!f) { std::cerr << "ref_driver: cannot write monitor_out" << std::endl; return 2; }
    for (int k = 0; k < 17; ++k) { std::vector<double> v = mem[k]->GetValue(); fwrite(v.data(), 8, v.size(), f); }
    fclose(f);
  }
  // optional: the reference's own text dump (Beam::PrintToFile, beam.cpp:670-707)
  if (cfg.count("print_to")) beam.PrintToFile(S(cfg, "print_to", "beam.txt"), S(cfg, "print_msg", "ref_driver"));
  std::cout << "ref_driver: loss=" << stats[0] << " lloss=" << stats[1] << " good=" << stats[2] << std::endl;
  return 0;
}
