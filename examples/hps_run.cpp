// hps_run.cpp -- a plain C++ host program on the C ABI (include/hpsim_b200.h), no Python involved: loads a lattice
// table (the binary dump hpsim_b200/lattice.py writes from the SQLite databases), generates the C1 DC beam with the
// reference's own generator, tracks a range with SCHEFF on and prints the run counters and the loss count.
// It is what a C++ maintainer's driver looks like when the engine behind SimulationEngine::Simulate
// (src/simulation/src/simulation_engine.cpp:73-95) is this library; see INTEGRATION.md section 4.
//
//   hps_run <lattice.lat> <first element> <last element> [particles=65536] [device=0]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hpsim_b200.h"

#define CHECK(ctx, call)                                                               \
  do {                                                                                 \
    int rc_ = (call);                                                                  \
    if (rc_ != HPS_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, hps_last_error(ctx)); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: hps_run lattice.lat first last [particles] [device]\n"); return 2; }
  const uint64_t n = argc > 4 ? strtoull(argv[4], nullptr, 10) : 65536;
  const int device = argc > 5 ? atoi(argv[5]) : 0;

  // lattice file: "HPSL", version, count, then `count` hps_elem records
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  char magic[4];
  uint32_t ver = 0, cnt = 0;
  if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "HPSL", 4) != 0 || fread(&ver, 4, 1, f) != 1 || fread(&cnt, 4, 1, f) != 1) { fprintf(stderr, "not a lattice file\n"); return 2; }
  std::vector<hps_elem> elems(cnt);
  if (fread(elems.data(), sizeof(hps_elem), cnt, f) != cnt) { fprintf(stderr, "short lattice file\n"); return 2; }
  fclose(f);

  hps_ctx* ctx = hps_ctx_create(device);
  if (!ctx) { fprintf(stderr, "hps_ctx_create: %s\n", hps_last_error(nullptr)); return 1; }  // no CPU fallback
  hps_lattice* lat = hps_lattice_create(ctx, elems.data(), cnt);
  const int64_t first = hps_lattice_index(lat, argv[2]), last = hps_lattice_index(lat, argv[3]);
  if (first < 0 || last < 0) { fprintf(stderr, "unknown element name\n"); return 2; }

  // the C1 beam recipe of the test-suite (tests/harness.py DC), seed fixed
  hps_beam* beam = hps_beam_create(ctx, n, 939.294, -1.0, 0.015);
  if (!beam) { fprintf(stderr, "hps_beam_create: %s\n", hps_last_error(ctx)); return 1; }
  CHECK(ctx, hps_beam_set_dc(beam, 0.095, 47.0, 0.00327, -0.102, 60.0, 0.002514, 3.14159265358979323846, 0.0, 0.7518, 20161201u));
  CHECK(ctx, hps_beam_set_frequency(beam, 201.25));
  CHECK(ctx, hps_beam_set_ref_w(beam, 0.750));
  const double phi_offset = -707.907480574;  // -L(TBDB02 -> BLZ) / (beta lambda) * 2 pi
  CHECK(ctx, hps_beam_set_ref_phi(beam, phi_offset));
  CHECK(ctx, hps_beam_shift(beam, HPS_PHI, phi_offset));
  CHECK(ctx, hps_beam_save(beam, 0));

  hps_scheff* sc = hps_scheff_create(ctx, 32, 128, 3);
  if (!sc) { fprintf(stderr, "hps_scheff_create: %s\n", hps_last_error(ctx)); return 1; }
  CHECK(ctx, hps_scheff_set(sc, HPS_SC_INTERVAL, 0.025));
  CHECK(ctx, hps_scheff_set(sc, HPS_SC_ADJ_BUNCH_CUTOFF_W, 0.8));
  CHECK(ctx, hps_scheff_set(sc, HPS_SC_REMESH_THRESHOLD, 0.02));

  CHECK(ctx, hps_engine_reset(ctx, beam));
  hps_run_stats rs;
  for (int rep = 0; rep < 2; ++rep) {  // the second pass replays the captured launch graph
    CHECK(ctx, hps_beam_restore(beam, 0));
    CHECK(ctx, hps_simulate(ctx, beam, lat, sc, (uint32_t)first, (uint32_t)last, &rs));
    printf("pass %d: %llu particles x %llu elements, %llu kicks (%llu re-mesh), %llu launches, %.3f ms on the device, %.3e particle-element steps/s\n",
           rep, (unsigned long long)rs.particles, (unsigned long long)rs.elements, (unsigned long long)rs.kicks, (unsigned long long)rs.remesh,
           (unsigned long long)rs.launches, rs.device_ms, (double)rs.particles * (double)rs.elements / (rs.device_ms * 1e-3));
  }
  CHECK(ctx, hps_beam_update(beam, HPS_UPD_LOSS));
  double nloss = 0.0;
  CHECK(ctx, hps_beam_get(beam, HPS_STAT_NUM_LOSS, &nloss));
  printf("lost %.0f of %llu\n", nloss, (unsigned long long)n);

  hps_scheff_destroy(sc);
  hps_beam_destroy(beam);
  hps_lattice_destroy(lat);
  hps_ctx_destroy(ctx);
  return 0;
}
