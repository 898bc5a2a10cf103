// hps_kernels.cuh -- sm_100a kernels of the HPSim tracking + SCHEFF hot path.
//
// One "segment" kernel applies every element map between two grid-wide dependency points
// (space-charge kick, centroid-centred aperture) to particles held in registers: 2 adjacent
// particles per thread, 128-bit coalesced loads/stores of each SoA array, no tensor cores
// (nothing here is a contraction).  The reference launches one kernel per half element
// (simulation/src/simulation_engine.cu:125-373); the maps themselves restate
// simulation/inc/simulation_kernel.cu:16-871 and space_charge_kernel.cu:573-706.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include "hps_internal.h"
#include "hps_math.cuh"

namespace hps {

// util/inc/constant.h:7-11
#define K_CLIGHT 299.792458
#define K_PI 3.14159265358979323846
#define K_TWOPI (3.14159265358979323846 * 2.0)

// the element maps are plain fp64 arithmetic: they also compile for the host, where tests/host_maps.cu runs
// the fast maps against the reference-order maps without a GPU
#define HPS_HD __host__ __device__ __forceinline__
HPS_HD double2 ld_field(const double2* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

struct Pt {
  double x, xp, y, yp, phi, w;
  u32 loss;
};

// ------------------------------------------------------------------------------------------------
// block / grid reduction with a deterministic final pass by the last block to finish
// ------------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void warp_reduce(double (&v)[NS], unsigned max_mask) {
#pragma unroll
  for (int s = 0; s < NS; ++s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double t = __shfl_xor_sync(0xffffffffu, v[s], o);
      v[s] = ((max_mask >> s) & 1u) ? fmax(v[s], t) : v[s] + t;
    }
  }
}

// ---- programmatic dependent launch ---------------------------------------------------------------
// The kick chain is a string of short kernels, each ending in a one-block reduction tail.  Launched with
// the programmatic-stream-serialization attribute, the next kernel's blocks become resident while that
// tail runs and wait here; pdl_wait() returns once the preceding grid has completed and flushed.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Loads the compiler may neither sink into a later branch nor reorder: the statistic sweeps test the loss flag of
// a particle before they touch its coordinates, and nvcc then moves the coordinate loads behind that branch --
// two or three dependent L2 round trips per iteration instead of one (the moment sweep took 15 us instead of 4).
__device__ __forceinline__ double2 ld_f64x2(const double* base, u32 i2) {
  double2 v;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(reinterpret_cast<const double2*>(base) + i2));
  return v;
}
// the same with the evict-first (cache-streaming) hint when `cs` is set: the statistic / deposit sweeps of a beam that
// streams through the L2 (BeamView::streaming) must not push the rest of the working set out of it
__device__ __forceinline__ double2 ld_f64x2(const double* base, u32 i2, u32 cs) {
  double2 v;
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p ld.global.cs.v2.f64 {%0, %1}, [%2]; @!p ld.global.v2.f64 {%0, %1}, [%2]; }"
               : "=d"(v.x), "=d"(v.y) : "l"(reinterpret_cast<const double2*>(base) + i2), "r"(cs));
  return v;
}
__device__ __forceinline__ uint2 ld_u32x2(const u32* base, u32 i2, u32 cs) {
  uint2 v;
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %3, 0; @p ld.global.cs.v2.u32 {%0, %1}, [%2]; @!p ld.global.v2.u32 {%0, %1}, [%2]; }"
               : "=r"(v.x), "=r"(v.y) : "l"(reinterpret_cast<const uint2*>(base) + i2), "r"(cs));
  return v;
}
__device__ __forceinline__ uint2 ld_u32x2(const u32* base, u32 i2) {
  uint2 v;
  asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(reinterpret_cast<const uint2*>(base) + i2));
  return v;
}

// ---- cross-rank exchange over peer memory (XrBox, hps_internal.h) ------------------------------
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define HPS_XR_TIMEOUT_NS 2000000000ull  // a peer that stays silent this long (2 s) is reported, not waited for

// LL word = 32-bit payload | (32-bit sequence number << 32); 8-byte stores are single transactions.
__device__ __forceinline__ void ll_send(unsigned long long* word, u32 payload, u32 seq32) {
  st_sys_u64(word, (unsigned long long)payload | ((unsigned long long)seq32 << 32));
}
// Polls a word of this rank's own box until a peer has written it for exchange `seq32`.
__device__ __forceinline__ u32 ll_recv(const unsigned long long* word, u32 seq32, u32* err) {
  unsigned long long w = ld_sys_u64(word);
  if ((u32)(w >> 32) != seq32) {
    const unsigned long long t0 = global_ns();
    do {
      w = ld_sys_u64(word);
      if (global_ns() - t0 > HPS_XR_TIMEOUT_NS) { *err = 1u; break; }
    } while ((u32)(w >> 32) != seq32);
  }
  return (u32)w;
}
__device__ __forceinline__ unsigned long long ll_recv64(const unsigned long long* pair, u32 seq32, u32* err) {
  const u32 lo = ll_recv(pair, seq32, err), hi = ll_recv(pair + 1, seq32, err);
  return (unsigned long long)lo | ((unsigned long long)hi << 32);
}

// All threads of ONE block (the rank's last block, >= 2 * nranks * ns threads) call this with the
// rank-local totals in shared memory; on return total[] holds the fold over all ranks in rank order
// (bit-identical everywhere).
// int_mask: slots that carry a signed 64-bit integer (bit pattern in the double) and are summed as integers -- the
// fixed-point sums of the fused-moment epilogue, exact and order-independent.
__device__ __forceinline__ void xr_exchange_p2p(double* total, int ns, unsigned max_mask, const XrView& xr, unsigned int_mask = 0u) {
  __shared__ unsigned long long s_seq;
  __shared__ u32 s_recv[HPS_XR_MAXR * 2 * HPS_XR_NS];
  if (threadIdx.x == 0) { s_seq = *xr.seq + 1ull; *xr.seq = s_seq; }
  __syncthreads();
  const u32 seq32 = (u32)s_seq;
  const int par = (int)(s_seq & 1ull);
  const int nw = 2 * ns;  // LL words per rank
  for (int t = threadIdx.x; t < xr.nranks * nw; t += blockDim.x) {
    const int p = t / nw, w = t - p * nw;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(total[w >> 1]);
    ll_send(&xr.peer[p]->data[par][xr.rank][w], (w & 1) ? (u32)(bits >> 32) : (u32)bits, seq32);
  }
  const XrBox* own = xr.peer[xr.rank];
  const unsigned long long t_sent = threadIdx.x == 0 ? global_ns() : 0ull;
  for (int t = threadIdx.x; t < xr.nranks * nw; t += blockDim.x) {
    const int r = t / nw, w = t - r * nw;
    s_recv[r * (2 * HPS_XR_NS) + w] = ll_recv(&own->data[par][r][w], seq32, xr.err);
  }
  __syncthreads();
  if (threadIdx.x == 0) { xr.stats[0] += global_ns() - t_sent; xr.stats[1] += 1ull; }  // own stores issued -> all peers' words here
  if ((int)threadIdx.x < ns) {
    const int s = threadIdx.x;
    const bool mx = (max_mask >> s) & 1u, as_int = (int_mask >> s) & 1u;
    double a = 0.0;
    long long ia = 0;
    for (int r = 0; r < xr.nranks; ++r) {
      const u32* q = &s_recv[r * (2 * HPS_XR_NS) + 2 * s];
      const long long bits = (long long)((unsigned long long)q[0] | ((unsigned long long)q[1] << 32));
      const double t = __longlong_as_double(bits);
      ia += bits;
      a = r == 0 ? t : (mx ? fmax(a, t) : a + t);
    }
    total[s] = as_int ? __longlong_as_double(ia) : a;
  }
  __syncthreads();
}

// Returns true in every thread of the block that finished last AND owns the final totals; `total`
// (shared) then holds the result of every slot over the whole (possibly sharded) beam.
// partials: [gridDim.x][NS].  XR_NCCL: the last block only hands its totals to xr.send and returns
// false; the host then enqueues an all-gather and k_xr_finalize.
// do_xr = false: the caller runs the cross-rank exchange itself (fused-moment epilogue); `total` then holds this
// rank's totals only.
template <int NS>
__device__ __forceinline__ bool grid_reduce(double (&v)[NS], unsigned max_mask, double* partials, u32* ticket,
                                            double* total /* shared, NS */, double* scratch /* shared, blockDim.x doubles */,
                                            const XrView& xr, bool do_xr = true) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  warp_reduce<NS>(v, max_mask);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) scratch[wid * NS + s] = v[s];
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    const int s = threadIdx.x;
    double a = scratch[s];
    for (int k = 1; k < nw; ++k) a = ((max_mask >> s) & 1u) ? fmax(a, scratch[k * NS + s]) : a + scratch[k * NS + s];
    partials[(size_t)blockIdx.x * NS + s] = a;
  }
  __shared__ u32 s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    // one fence by the thread that takes the ticket: fences are cumulative, so the partials the other threads wrote
    // before the barrier are ordered before the ticket too (256 fences per block each waited for the thread's own
    // streaming stores to drain)
    __threadfence();
    u32 t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
    __threadfence();
  }
  __syncthreads();
  if (!s_last) return false;
  // last block: thread t folds slot (t % NSP) of the blocks g, g+G, g+2G, ... (g = t / NSP; the loads of
  // one thread are independent, so they overlap); the G group sums are then folded in group order.
  // Every order is fixed, so the result is bit-reproducible for a given grid.
  {
    constexpr int NSP = NS <= 4 ? 4 : (NS <= 8 ? 8 : 16);
    const int G = blockDim.x / NSP;  // 256 threads -> 64 / 32 / 16 groups
    const int s = threadIdx.x % NSP, g = threadIdx.x / NSP;
    const bool mx = (max_mask >> s) & 1u;
    double a = mx ? -DBL_MAX : 0.0;
    if (s < NS) {
      // batches of 16 loads in flight before the (ordered) folds: the tail is L2-latency bound
      for (u32 k0 = g; k0 < gridDim.x; k0 += 16 * G) {
        double t[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const u32 k = k0 + j * G;
          t[j] = k < gridDim.x ? __ldcg(&partials[(size_t)k * NS + s]) : (mx ? -DBL_MAX : 0.0);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) a = mx ? fmax(a, t[j]) : a + t[j];
      }
    }
    __syncthreads();  // scratch (blockDim.x doubles) is free again
    scratch[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x < NS) {
      const bool m2 = (max_mask >> threadIdx.x) & 1u;
      double r = scratch[threadIdx.x];
      for (int q = 1; q < G; ++q) {
        const double t = scratch[q * NSP + threadIdx.x];
        r = m2 ? fmax(r, t) : r + t;
      }
      total[threadIdx.x] = r;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
  __syncthreads();
  if (!do_xr) return true;
  if (xr.mode == XR_P2P) xr_exchange_p2p(total, NS, max_mask, xr);
  else if (xr.mode == XR_NCCL) {
    if (threadIdx.x < NS) xr.send[threadIdx.x] = total[threadIdx.x];
    return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------
// element maps (one particle)
// ------------------------------------------------------------------------------------------------
// The reference-order maps are templates over the arithmetic back end M (hps_math.cuh): MathLib = nvcc's own
// `/` and sqrt (with their slow-path branches), MathNR = the same IEEE results from branch-free sequences, with
// divisions by per-op constants done as one multiplication + exact-remainder correction.  Same bits either way.
template <class M>
HPS_HD double beta_of(double w, double mass, double inv_mass) {
  double g1 = M::udiv_(w, mass, inv_mass);
  return M::div(M::sqrt_(g1 * (g1 + 2.0)), g1 + 1.0);
}

// ---- hoisted constants -------------------------------------------------------------------------
// slots of OpC::c per op type
enum { D_APER2 = 0, D_TWOPI_HL, D_PH_OVER_WL };
enum { Q_APER2 = 0, Q_QG, Q_ABSQG, Q_LEN4, Q_TWOPI_LEN_HALF, Q_KC, Q_PH_OVER_WL };
enum { R_COS = 0, R_SIN };
enum { GC_CELL_LEN = 0, GC_DD, GC_WAVE_LEN, GC_BETA_IN, GC_DPS, GC_PI_CLBL_DPS, GC_D_CELL_LEN, GC_PI_DCL, GC_TWOPI_HALF_QLEN,
       GC_BETAG_CCL, GC_TWOPI_DD,
       // second half only
       GC_INV_MASS, GC_BETAG, GC_GMG, GC_INV_GMG2, GC_AQ, GC_BETA_OUT, GC_EFL, GC_CONCOMM, GC_CONCOMM1, GC_CON1, GC_CON2NEX,
       GC_INV_GMGCL, GC_INV_FKIS, GC_FKR, GC_SINPS, GC_COSPS, GC_X5, GC_TTF_DEN, GC_PH_CONST, GC_TWOPI_DG, GC_AQ_EFL,
       // reciprocal / pre-divided forms used by the fast maps (slots 32..)
       GF_DD_OVER_WL = 32,   // 2 pi dd / wavelength in force (the amp == 0 drift of either half)
       GF_DCLQ_OVER_WL,      // (pi d_cell_len [0 when |d_cell_len| <= 1e-15] - 2 pi qlen/2) / wave_len
       GF_IBETA_IN,          // 1 / beta_in
       GF_IBETA_OUT,         // 1 / beta_out
       GF_INV_TTF_DEN, GF_INV_GMG, GF_DG_OVER_WL, GF_INV_FKR, GF_BETAG_OVER_BIN };

// indices into Op::p for the RF gap (rfgap_parameter.h:4-32)
enum { G_LEN = 0, G_FREQ, G_CLBL, G_AMP, G_PHREF, G_PHSHIFT, G_EOUT, G_BETAC, G_DG, G_T, G_TP, G_SP,
       G_FBC, G_FBMIN, G_T0, G_T1, G_T2, G_T3, G_T4, G_T5, G_S1, G_S2, G_S3, G_S4, G_S5,
       G_DESIGN_W = 25, G_PHI_IN = 26, G_QLEN1 = 27, G_QLEN2 = 28, G_CELLNUM = 29, G_WAVELEN = 30 };

// Evaluates the particle-independent sub-expressions of one op.  Every value is produced by the very
// expression (same operand order) the reference evaluates inside its per-particle loop.
__host__ __device__ inline void prepare_op(const Op& op, OpC& o, double mass, double charge) {
  const double* q = op.p;
  double* c = o.c;
  switch (op.type) {
    case OP_DRIFT:
      c[D_APER2] = q[1] * q[1];
      c[D_TWOPI_HL] = K_TWOPI * q[0];
      c[D_PH_OVER_WL] = c[D_TWOPI_HL] / q[2];
      break;
    case OP_QUAD_HALF: {
      double aper2 = q[1]; aper2 *= aper2;
      const double qg = q[2] * charge;
      c[Q_APER2] = aper2;
      c[Q_QG] = qg;
      c[Q_ABSQG] = qg > 0.0 ? qg : -qg;
      c[Q_LEN4] = q[0] * 0.25;
      c[Q_TWOPI_LEN_HALF] = K_TWOPI * q[0] * 0.5;
      c[Q_KC] = c[Q_ABSQG] * K_CLIGHT / mass;
      c[Q_PH_OVER_WL] = c[Q_TWOPI_LEN_HALF] / q[3];
      break;
    }
    case OP_ROTATION:
      c[R_COS] = cos(q[0]);
      c[R_SIN] = sin(q[0]);
      break;
    case OP_GAP_FIRST:
    case OP_GAP_SECOND: {
      const bool ccl = op.flag != 0, second = op.type == OP_GAP_SECOND;
      const double qlen1 = q[G_QLEN1], qlen2 = q[G_QLEN2];
      const double cell_len = q[G_LEN] + qlen1 * 0.5 + qlen2 * 0.5;
      const double wave_len = K_CLIGHT / q[G_FREQ];
      c[GC_CELL_LEN] = cell_len;
      c[GC_WAVE_LEN] = wave_len;
      const double dd = second ? 0.5 * cell_len + q[G_DG] - qlen2 * 0.5 : 0.5 * cell_len - q[G_DG] - 0.5 * qlen1;
      c[GC_DD] = dd;
      c[GC_TWOPI_DD] = K_TWOPI * dd;
      const double d_cell_len = cell_len - q[G_BETAC] * wave_len;
      c[GC_D_CELL_LEN] = d_cell_len;
      c[GC_PI_DCL] = K_PI * d_cell_len;
      c[GC_TWOPI_HALF_QLEN] = K_TWOPI * 0.5 * (second ? qlen2 : qlen1);
      c[GF_DD_OVER_WL] = c[GC_TWOPI_DD] / (second ? q[G_WAVELEN] : wave_len);
      c[GF_DCLQ_OVER_WL] = (((d_cell_len > 1e-15 || d_cell_len < -1e-15) ? c[GC_PI_DCL] : 0.0) - c[GC_TWOPI_HALF_QLEN]) / wave_len;
      if (!second) {
        double gam = q[G_DESIGN_W] / mass;
        c[GC_BETA_IN] = sqrt((gam + 2.0) * gam) / (gam + 1.0);
        const double dps = q[G_PHREF] - q[G_PHI_IN];
        c[GC_DPS] = dps;
        c[GC_PI_CLBL_DPS] = K_PI * q[G_CLBL] + dps;
        c[GC_BETAG_CCL] = cell_len / (q[G_CLBL] * wave_len);  // "betag" of the CCL first half
        c[GF_IBETA_IN] = 1.0 / c[GC_BETA_IN];
      } else {
        const double inv_mass = 1.0 / mass;
        c[GC_INV_MASS] = inv_mass;
        const double gm_in = q[G_DESIGN_W] * inv_mass + 1.0;
        const double beta_in = sqrt(1.0 - 1.0 / (gm_in * gm_in));
        c[GC_BETA_IN] = beta_in;
        const double cl = cell_len / q[G_CLBL];
        const double betag_ccl = cl / wave_len;
        c[GC_BETAG_CCL] = betag_ccl;
        const double betag = ccl ? q[G_BETAC] : betag_ccl;
        c[GC_BETAG] = betag;
        const double betag2 = betag * betag;
        const double gmg = 1.0 / sqrt(1.0 - betag2);
        c[GC_GMG] = gmg;
        const double inv_gmg2 = 1.0 / (gmg * gmg);
        c[GC_INV_GMG2] = inv_gmg2;
        const double aq = charge > 0.0 ? charge : -charge;
        c[GC_AQ] = aq;
        const double gm_out = q[G_EOUT] * inv_mass + 1.0;
        c[GC_BETA_OUT] = sqrt(1.0 - 1.0 / (gm_out * gm_out));
        const double efl = q[G_AMP] * cell_len;
        c[GC_EFL] = efl;
        c[GC_AQ_EFL] = aq * efl;
        const double concomm = aq * efl / (mass * betag2);
        c[GC_CONCOMM] = concomm;
        const double concomm1 = concomm * inv_gmg2 / gmg;
        c[GC_CONCOMM1] = concomm1;
        c[GC_CON1] = K_PI * concomm1;
        c[GC_CON2NEX] = concomm * inv_gmg2;
        double inv_gmgcl, inv_fkis;
        if (ccl) {
          inv_gmgcl = 1.0 / (gmg * betag * wave_len);
          inv_fkis = betag * wave_len / K_TWOPI;
          c[GC_X5] = betag_ccl * wave_len;
        } else {
          inv_gmgcl = 1.0 / (gmg * cl);
          inv_fkis = cl / K_TWOPI;
          c[GC_X5] = beta_in * wave_len;
        }
        c[GC_INV_GMGCL] = inv_gmgcl;
        c[GC_INV_FKIS] = inv_fkis;
        c[GC_FKR] = K_TWOPI * inv_gmgcl;
        double sinps, cosps;
        sincos(q[G_PHREF], &sinps, &cosps);
        c[GC_SINPS] = sinps;
        c[GC_COSPS] = cosps;
        c[GC_TTF_DEN] = K_TWOPI * q[G_FBC];
        c[GC_TWOPI_DG] = K_TWOPI * q[G_DG];
        const double dps = ccl ? q[G_PHI_IN] - q[G_PHREF] : q[G_PHREF] - q[G_PHI_IN];
        c[GC_DPS] = dps;
        c[GC_PI_CLBL_DPS] = K_PI * q[G_CLBL] + dps;
        c[GC_PH_CONST] = (u32)q[G_CELLNUM] * K_PI;
        c[GF_IBETA_IN] = 1.0 / beta_in;
        c[GF_IBETA_OUT] = 1.0 / c[GC_BETA_OUT];
        c[GF_INV_TTF_DEN] = 1.0 / c[GC_TTF_DEN];
        c[GF_INV_GMG] = 1.0 / gmg;
        c[GF_DG_OVER_WL] = c[GC_TWOPI_DG] / wave_len;
        c[GF_INV_FKR] = 1.0 / c[GC_FKR];
        c[GF_BETAG_OVER_BIN] = betag_ccl / beta_in;
      }
      break;
    }
    default: break;
  }
}

// simulation_kernel.cu:87-118 -- a lost particle keeps its pre-drift x,y (H11)
template <class M>
HPS_HD void op_drift(Pt& p, const double* q, const double* c, u32 elem, double mass, double inv_mass) {
  const double hl = q[0], aper2 = c[D_APER2], wavelen = q[2];
  double x = p.x, y = p.y;
  x += p.xp * hl;
  y += p.yp * hl;
  if (aper2 != 0.0 && x * x + y * y > aper2) p.loss = elem;
  else {
    p.x = x; p.y = y;
    p.phi += M::div(c[D_TWOPI_HL], wavelen * beta_of<M>(p.w, mass, inv_mass));
  }
}

// simulation_kernel.cu:142-218 -- two quarter-length thick-lens maps, then the aperture test
template <class M>
HPS_HD void op_quad_half(Pt& p, const double* q, const double* c, u32 elem, double mass, double inv_mass) {
  const double length = q[0], wavelen = q[3];
  const double aper2 = c[Q_APER2], qg = c[Q_QG];
  double tx = p.x, ty = p.y;  // what the aperture test sees
  if (qg == 0) {
    // the reference stores the drifted x,y but tests the aperture with the OLD values
    p.x = tx + p.xp * length * 0.5;
    p.y = ty + p.yp * length * 0.5;
  } else {
    double gamma1 = M::udiv_(p.w, mass, inv_mass);
    double bg = M::sqrt_(gamma1 * (gamma1 + 2.0) * (1.0 + p.xp * p.xp + p.yp * p.yp));
    double k = M::sqrt_(M::div(c[Q_ABSQG], bg * mass) * K_CLIGHT);
    double inv_k = M::rcp(k);
    double kds = c[Q_LEN4] * k;
    double s, cs, sh = sinh(kds), ch = cosh(kds);
    sincos(kds, &s, &cs);
    double sovk = s * inv_k, shovk = sh * inv_k, sk = s * k, shk = sh * k;
    double x = p.x, xp = p.xp, y = p.y, yp = p.yp, t;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int rep = 0; rep < 2; ++rep) {
      if (qg > 0.0) {
        t = cs * x + sovk * xp; xp = -sk * x + cs * xp; x = t;
        t = ch * y + shovk * yp; yp = shk * y + ch * yp; y = t;
      } else {
        t = ch * x + shovk * xp; xp = shk * x + ch * xp; x = t;
        t = cs * y + sovk * yp; yp = -sk * y + cs * yp; y = t;
      }
    }
    p.x = x; p.xp = xp; p.y = y; p.yp = yp;
    tx = x; ty = y;
  }
  if (aper2 != 0 && tx * tx + ty * ty > aper2) p.loss = elem;
  else p.phi += M::div(c[Q_TWOPI_LEN_HALF], wavelen * beta_of<M>(p.w, mass, inv_mass));
}

// simulation_kernel.cu:226-288
template <class M>
HPS_HD void op_gap_first(Pt& p, const double* q, const double* c, bool ccl, double mass, double inv_mass) {
  const double dd1 = c[GC_DD];
  const double phi = p.phi, w = p.w;
  p.x = p.x + dd1 * p.xp;
  p.y = p.y + dd1 * p.yp;
  const double wave_len = c[GC_WAVE_LEN];
  if (q[G_AMP] == 0.0) {
    p.phi += M::div(c[GC_TWOPI_DD], wave_len * beta_of<M>(w, mass, inv_mass));
  } else {
    const double beta_in = c[GC_BETA_IN];
    const double gam = M::udiv_(w, mass, inv_mass);
    const double beta = M::div(M::sqrt_((gam + 2.0) * gam), gam + 1.0);
    if (!ccl) {
      const double d_cell_len = c[GC_D_CELL_LEN];
      const double den = beta * wave_len, rden = M::rcp(den);  // two quotients share this denominator
      const double dphi_qlen1 = M::udiv_(c[GC_TWOPI_HALF_QLEN], den, rden);
      double dphi_clen = 0.0;
      if (d_cell_len > 1e-15 || d_cell_len < -1e-15) dphi_clen = M::udiv_(c[GC_PI_DCL], den, rden);
      p.phi = phi + (1.0 - M::div(beta - beta_in, beta)) * c[GC_PI_CLBL_DPS] + dphi_clen - dphi_qlen1;
    } else {
      p.phi = phi + (1.0 - M::div((beta - beta_in) * c[GC_BETAG_CCL], beta * beta_in)) * c[GC_PI_CLBL_DPS];
    }
  }
}

// simulation_kernel.cu:16-73 (polynomial modified Bessel functions)
template <class M>
HPS_HD double bessi0c(double rx) {
  double ans, y, ax = fabs(rx);
  if (ax < 3.75) {
    y = M::udiv_(rx, 3.75, 1.0 / 3.75); y = y * y;
    ans = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
  } else {
    y = M::div(3.75, ax);
    ans = (exp(ax) * rsqrt(ax)) * (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 +
          y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
  }
  return ans;
}
template <class M>
HPS_HD double bessi1c(double rx) {
  double y, ans, ax = fabs(rx);
  if (ax < 3.75) {
    y = M::udiv_(rx, 3.75, 1.0 / 3.75); y = y * y;
    ans = ax * (0.5 + y * (0.87890594 + y * (0.51498869 + y * (0.15084934 + y * (0.2658733e-1 + y * (0.301532e-2 + y * 0.32411e-3))))));
  } else {
    y = M::div(3.75, ax);
    ans = 0.2282967e-1 + y * (-0.2895312e-1 + y * (0.1787654e-1 - y * 0.420059e-2));
    ans = 0.39894228 + y * (-0.3988024e-1 + y * (-0.362018e-2 + y * (0.163801e-2 + y * (-0.1031555e-1 + y * ans))));
    ans = ans * (exp(ax) * rsqrt(ax));
  }
  return rx < 0.0 ? -ans : ans;
}

// simulation_kernel.cu:290-497
template <class M>
HPS_HD void op_gap_second(Pt& p, const double* q, const double* c, bool ccl, double mass, double inv_mass_rn) {
  const double dd2 = c[GC_DD];
  double x = p.x, y = p.y, xp = p.xp, yp = p.yp;
  const double inv_mass = c[GC_INV_MASS];
  const double phi = p.phi;
  double w = p.w;
  const double beta_in = c[GC_BETA_IN];
  double gam = w * inv_mass;
  const double bg1 = M::sqrt_(gam * (gam + 2.0));
  double beta = M::div(bg1, gam + 1.0);
  // transit-time factors from the 5th-order fits (GetTransitTimeFactors)
  double t_fac, tp_fac, sp_fac;
  {
    double bb = beta;
    if (beta > 1.1 * q[G_FBC]) bb = q[G_FBC];
    else if (beta < q[G_FBMIN]) bb = q[G_FBMIN];
    t_fac = q[G_T0] + bb * (q[G_T1] + bb * (q[G_T2] + bb * (q[G_T3] + bb * (q[G_T4] + bb * q[G_T5]))));
    double coef = M::udiv_(bb * bb, c[GC_TTF_DEN], c[GF_INV_TTF_DEN]);
    tp_fac = coef * (q[G_T1] + bb * (2 * q[G_T2] + bb * (3 * q[G_T3] + bb * (4 * q[G_T4] + 5 * bb * q[G_T5]))));
    sp_fac = -coef * (q[G_S1] + bb * (2 * q[G_S2] + bb * (3 * q[G_S3] + bb * (4 * q[G_S4] + 5 * bb * q[G_S5]))));
  }
  if (q[G_AMP] == 0.0 || (ccl && beta < q[G_FBMIN])) {
    p.x = x + dd2 * xp;
    p.y = y + dd2 * yp;
    p.phi += M::div(c[GC_TWOPI_DD], q[G_WAVELEN] * beta_of<M>(w, mass, inv_mass_rn));
    return;
  }
  const double wave_len = c[GC_WAVE_LEN];
  const double betag_ccl = c[GC_BETAG_CCL];
  const double gmg = c[GC_GMG], rgmg = c[GF_INV_GMG];
  const double inv_gmg2 = c[GC_INV_GMG2];
  const double efl = c[GC_EFL];
  const double concomm = c[GC_CONCOMM];
  const double con1 = c[GC_CON1];
  const double con2nex = c[GC_CON2NEX];
  const double inv_gmgcl = c[GC_INV_GMGCL], inv_fkis = c[GC_INV_FKIS];
  const double con5nex = concomm * (t_fac * inv_fkis - c[GC_X5] * tp_fac * inv_gmg2);
  double tw;
  if (ccl) tw = t_fac - M::udiv_(K_TWOPI * tp_fac * (M::div(beta_in, beta) - 1.0) * betag_ccl, beta_in, c[GF_IBETA_IN]);
  else tw = t_fac;
  const double con5nex1 = c[GC_CONCOMM1] * t_fac;
  const double con5nex2 = -con2nex * t_fac * inv_fkis;
  const double fkr = c[GC_FKR];
  const double sinps = c[GC_SINPS], cosps = c[GC_COSPS];
  double phi1, sinp1, cosp1;
  if (ccl) phi1 = phi + q[G_PHSHIFT] - M::div(c[GC_TWOPI_DG], beta * wave_len) + c[GC_PH_CONST];
  else phi1 = phi + q[G_PHSHIFT];
  sincos(phi1, &sinp1, &cosp1);
  const double prome1 = -con1 * (tp_fac * (sinp1 - sinps) + sp_fac * (cosp1 - cosps));
  sincos(phi1 + prome1, &sinp1, &cosp1);

  const double rs_p = x * x + y * y;
  const double rrp = x * xp + y * yp;
  const double ris = M::sqrt_(rs_p);
  const double rris = M::rcp(ris);  // b1c / ris and rrp / ris share it; unused (and NaN) when ris == 0
  const double fkrris = fkr * ris;
  const double b0c = bessi0c<M>(fkrris);
  const double b1c = bessi1c<M>(fkrris);
  const double b1p = (fkrris != 0.0) ? b0c - M::div(b1c, fkrris) : 0.5;
  double bes1covr, rpris = 0.0;
  if (ris != 0.0) bes1covr = M::udiv_(b1c, ris, rris); else bes1covr = fkr * 0.5;
  const double delte = c[GC_AQ_EFL] * (tw * b0c * cosp1 + (M::udiv_(tw - K_TWOPI * tp_fac, gmg, rgmg) * bes1covr + M::udiv_(tw, gmg, rgmg) * fkr * b1p -
                                                           tw * gmg * bes1covr) * rrp * sinp1);
  w += delte;
  gam = w * inv_mass;
  const double bg = M::sqrt_(gam * (gam + 2.0));
  beta = M::div(bg, gam + 1.0);
  const double bgr = M::div(bg1, bg);
  const double onema = 1.0 - (con5nex * bes1covr + con5nex1 * b1p + con5nex2 * bes1covr) * cosp1;
  const double dovbl = con2nex * bes1covr * tw * sinp1;
  const double bgrovonema = M::div(bgr, onema);
  xp = -dovbl * x + bgrovonema * xp;
  x = x * onema + dd2 * xp;
  yp = -dovbl * y + bgrovonema * yp;
  y = y * onema + dd2 * yp;
  if (rs_p > 0.0) rpris = M::udiv_(rrp, ris, rris);
  const double prome = 2.0 * con1 * (-tp_fac * b0c + ris * tw * b1c * inv_gmgcl) * (sinp1 - sinps) -
                       efl * inv_mass * tw * b1c * rpris * (cosp1 - cosps);
  const double beta_out = c[GC_BETA_OUT];
  double phi_new;
  if (ccl) {
    phi_new = phi + (1.0 - M::div((beta - beta_out) * betag_ccl, beta * beta_out)) * c[GC_PI_CLBL_DPS] + prome;
  } else {
    const double d_cell_len = c[GC_D_CELL_LEN];
    const double den = beta * wave_len, rden = M::rcp(den);
    const double dphi_qlen2 = M::udiv_(c[GC_TWOPI_HALF_QLEN], den, rden);
    double dphi_clen = 0.0;
    if (d_cell_len > 1e-15 || d_cell_len < -1e-15) dphi_clen = M::udiv_(c[GC_PI_DCL], den, rden);
    phi_new = phi + (1.0 - M::div(beta - beta_out, beta)) * c[GC_PI_CLBL_DPS] + prome + dphi_clen - dphi_qlen2;
  }
  if (w <= 0.0 || x != x || y != y || xp != xp || yp != yp || w != w || phi_new != phi_new) {
    p.loss = 33333333u;  // the reference also printf()s here; coordinates keep their pre-gap values
  } else {
    p.x = x; p.y = y; p.xp = xp; p.yp = yp; p.w = w; p.phi = phi_new;
  }
}

// simulation_kernel.cu:603-805.  q: 0 radius 1 angle 2 edge_in 3 edge_out 4 half_gap 5 k1 6 k2 7 field_index
// 8 kinetic_energy 9 wavelen
HPS_HD void op_dipole(Pt& p, const double* q, bool second, double mass) {
  double x = p.x, xp = p.xp, y = p.y, yp = p.yp;
  const double h = 1.0 / q[0];
  const double hf_g = q[4], k1 = q[5], k2 = q[6], findx = q[7], design_w = q[8];
  if (!second) {
    const double b1 = q[2];
    double c1 = 2.0 * hf_g * k1 * h;
    double sinb, inv_cosb;
    sincos(b1, &sinb, &inv_cosb);
    inv_cosb = 1.0 / inv_cosb;
    double psi = c1 * (1.0 + sinb * sinb) * inv_cosb * (1.0 - c1 * k2 * sinb * inv_cosb);
    xp += tan(b1) * h * x;
    yp -= tan(b1 - psi) * h * y;
  }
  const double h2 = h * h;
  const double dl = 0.5 * q[0] * q[1];
  const double yk2 = findx * h2;
  const double yk = sqrt(fabs(yk2));
  const double inv_yk = 1.0 / yk;
  const double ykl = yk * dl;
  const double xk2 = h2 - yk2;
  const double inv_xk2 = 1.0 / xk2;
  const double xk = sqrt(fabs(xk2));
  const double inv_xk = 1.0 / xk;
  const double xkl = xk * dl;
  double cxkl = 1.0, sxokx = dl;
  if (xk2 > 0.0) { cxkl = cos(xkl); sxokx = sin(xkl) * inv_xk; }
  else if (xk2 < 0.0) { cxkl = cosh(xkl); sxokx = sinh(xkl) * inv_xk; }
  double cykl = 1.0, syoky = dl;
  if (yk2 > 0.0) { cykl = cos(ykl); syoky = sin(ykl) * inv_yk; }
  else if (yk2 < 0.0) { cykl = cosh(ykl); syoky = sinh(ykl) * inv_yk; }
  const double r11 = cxkl, r12 = sxokx, r21 = -xk2 * sxokx, r22 = cxkl;
  const double r33 = cykl, r34 = syoky, r43 = -yk2 * syoky, r44 = cykl;
  const double r26 = h * sxokx, r16 = h * (1.0 - cxkl) * inv_xk2;
  const double r51 = -r26, r52 = -r16;
  const double gamma = design_w / mass + 1.0;
  const double beta2 = 1.0 - 1.0 / (gamma * gamma);
  const double beta = sqrt(beta2);
  const double r56 = -h2 * (dl * beta2 - sxokx) * inv_xk2 + dl * (1.0 - h2 * inv_xk2) / (gamma * gamma);
  const double dpop = gamma / (design_w * (gamma + 1.0)) * (p.w - design_w);
  p.phi += K_TWOPI * (-(r51 * x + r52 * xp + r56 * dpop) + dl) / (beta * q[9]);
  const double x_new = r11 * x + r12 * xp + r16 * dpop;
  const double xp_new = r21 * x + r22 * xp + r26 * dpop;
  const double y_new = r33 * y + r34 * yp;
  const double yp_new = r43 * y + r44 * yp;
  p.x = x_new; p.y = y_new;
  if (!second) {
    p.xp = xp_new; p.yp = yp_new;
  } else {
    const double b2 = q[3];
    double c1 = 2.0 * hf_g * k1 * h;
    double sinb = sin(b2), inv_cosb = 1.0 / cos(b2);
    double psi = c1 * (1.0 + sinb * sinb) * inv_cosb * (1.0 - c1 * k2 * sinb * inv_cosb);
    p.xp = xp_new + tan(b2) * h * x_new;
    p.yp = yp_new - tan(b2 - psi) * h * y_new;
  }
}

// simulation_kernel.cu:807-871.  q: 0 w_ref 1 freq 2 voltage 3 phase 4 aperture 5 freq_ratio 6 wavelen
HPS_HD void op_buncher(Pt& p, const double* q, u32 elem, double mass, double charge) {
  const double wavelen = q[6];
  const double x = p.x, y = p.y, w = p.w;
  const double phi = p.phi * q[5];
  const double r2 = x * x + y * y;
  const double aper2 = q[4] * q[4];
  if (aper2 != 0.0 && r2 > aper2) { p.loss = elem; return; }
  const double h = q[1] * wavelen / K_CLIGHT;
  const double gm1 = q[0] / mass;
  const double gm = gm1 + 1.0;
  const double btgm = sqrt(gm1 * (gm1 + 2.0));
  const double beta = btgm / gm;
  const double k = h * K_TWOPI / (btgm * wavelen);
  const double k2 = k * k;
  const double aq = charge > 0.0 ? charge : -charge;
  const double con1 = K_TWOPI * h * q[2] * aq / (beta * wavelen * mass);
  const double arg = k2 * r2 * 0.25;
  const double ph = h * phi + q[3];
  const double dw = aq * q[2] * cos(ph) * (1.0 + arg * (1.0 + arg * (0.25 + arg / 36.0)));
  const double wb_par = w + 0.5 * dw;
  const double wf = w + dw;
  if (wb_par <= 0.0 || wf <= 0.0) { p.loss = elem; return; }
  double gmtmp = wb_par / mass;
  const double bav = sqrt(gmtmp * (gmtmp + 2.0)) / (gmtmp + 1.0);
  gmtmp = w / mass;
  const double bg = sqrt(gmtmp * (gmtmp + 2.0));
  gmtmp = wf / mass;
  const double rbgf = rsqrt(gmtmp * (gmtmp + 2.0));
  const double del = -con1 * sin(ph) * (1.0 - bav * beta) * (0.5 + arg * (0.25 + arg / 24.0)) / bav;
  p.xp = (p.xp * bg + del * x) * rbgf;
  p.yp = (p.yp * bg + del * y) * rbgf;
  p.w = wf;
}

// ------------------------------------------------------------------------------------------------
// fast maps (MATH == 1): the same formulas with fewer fp64 divisions and square roots
// ------------------------------------------------------------------------------------------------
// The sweep is bound by the latency of dependent fp64 chains (profiles/r1_ncu_k_segment.txt: fp64 pipe 30 %,
// top stall "wait"); an IEEE division is ~12 and a square root ~16 dependent instructions.  Here every
// division by a block-uniform quantity is a multiplication by a reciprocal prepared once per op
// (prepare_op / kick_prepare), sqrt-then-divide pairs are one rsqrt, and quotients that share a
// denominator share one reciprocal.  Each rewrite moves a result by a few ulp (1e-16 relative); the parity
// budget is 1e-9.  The aperture / loss tests themselves are evaluated exactly as in the reference.
HPS_HD double inv_beta_of(double w, double inv_mass) {
  const double g1 = w * inv_mass;
  return (g1 + 1.0) * rsqrt(g1 * (g1 + 2.0));
}

HPS_HD void op_drift_f(Pt& p, const double* q, const double* c, u32 elem, double inv_mass) {
  const double hl = q[0], aper2 = c[D_APER2];
  double x = p.x, y = p.y;
  x += p.xp * hl;
  y += p.yp * hl;
  if (aper2 != 0.0 && x * x + y * y > aper2) p.loss = elem;
  else {
    p.x = x; p.y = y;
    p.phi += c[D_PH_OVER_WL] * inv_beta_of(p.w, inv_mass);
  }
}

HPS_HD void op_quad_half_f(Pt& p, const double* q, const double* c, u32 elem, double inv_mass) {
  const double length = q[0];
  const double aper2 = c[Q_APER2], qg = c[Q_QG];
  double tx = p.x, ty = p.y;
  const double g1 = p.w * inv_mass;
  const double t = g1 * (g1 + 2.0);
  if (qg == 0) {
    p.x = tx + p.xp * length * 0.5;
    p.y = ty + p.yp * length * 0.5;
  } else {
    const double rbg = rsqrt(t * (1.0 + p.xp * p.xp + p.yp * p.yp));  // 1 / (beta gamma)_perp
    const double k2 = c[Q_KC] * rbg;                                   // |qG| c / (m bg)
    const double inv_k = rsqrt(k2), k = k2 * inv_k;
    const double kds = c[Q_LEN4] * k;
    double s, cs, sh = sinh(kds), ch = cosh(kds);
    sincos(kds, &s, &cs);
    const double sovk = s * inv_k, shovk = sh * inv_k, sk = s * k, shk = sh * k;
    double x = p.x, xp = p.xp, y = p.y, yp = p.yp, u;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int rep = 0; rep < 2; ++rep) {
      if (qg > 0.0) {
        u = cs * x + sovk * xp; xp = -sk * x + cs * xp; x = u;
        u = ch * y + shovk * yp; yp = shk * y + ch * yp; y = u;
      } else {
        u = ch * x + shovk * xp; xp = shk * x + ch * xp; x = u;
        u = cs * y + sovk * yp; yp = -sk * y + cs * yp; y = u;
      }
    }
    p.x = x; p.xp = xp; p.y = y; p.yp = yp;
    tx = x; ty = y;
  }
  if (aper2 != 0 && tx * tx + ty * ty > aper2) p.loss = elem;
  else p.phi += c[Q_PH_OVER_WL] * ((g1 + 1.0) * rsqrt(t));
}

HPS_HD void op_gap_first_f(Pt& p, const double* q, const double* c, bool ccl, double inv_mass) {
  const double dd1 = c[GC_DD];
  const double phi = p.phi;
  p.x = p.x + dd1 * p.xp;
  p.y = p.y + dd1 * p.yp;
  const double ib = inv_beta_of(p.w, inv_mass);
  if (q[G_AMP] == 0.0) p.phi = phi + c[GF_DD_OVER_WL] * ib;
  else if (!ccl) p.phi = phi + (c[GC_BETA_IN] * ib) * c[GC_PI_CLBL_DPS] + c[GF_DCLQ_OVER_WL] * ib;
  else p.phi = phi + (1.0 - c[GC_BETAG_CCL] * (c[GF_IBETA_IN] - ib)) * c[GC_PI_CLBL_DPS];
}

// both modified Bessel functions from one argument reduction (simulation_kernel.cu:16-73)
HPS_HD void bessi01c(double rx, double& b0, double& b1) {
  const double ax = fabs(rx);
  if (ax < 3.75) {
    double y = rx * (1.0 / 3.75); y = y * y;
    b0 = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
    b1 = ax * (0.5 + y * (0.87890594 + y * (0.51498869 + y * (0.15084934 + y * (0.2658733e-1 + y * (0.301532e-2 + y * 0.32411e-3))))));
  } else {
    const double y = 3.75 / ax;
    const double e = exp(ax) * rsqrt(ax);
    b0 = e * (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 +
         y * (-0.2057706e-1 + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
    double a = 0.2282967e-1 + y * (-0.2895312e-1 + y * (0.1787654e-1 - y * 0.420059e-2));
    a = 0.39894228 + y * (-0.3988024e-1 + y * (-0.362018e-2 + y * (0.163801e-2 + y * (-0.1031555e-1 + y * a))));
    b1 = a * e;
  }
  if (rx < 0.0) b1 = -b1;
}

HPS_HD void op_gap_second_f(Pt& p, const double* q, const double* c, bool ccl, double inv_mass) {
  const double dd2 = c[GC_DD];
  double x = p.x, y = p.y, xp = p.xp, yp = p.yp;
  const double phi = p.phi;
  double w = p.w;
  double gam = w * inv_mass;
  const double t1 = gam * (gam + 2.0);
  const double rbg1 = rsqrt(t1);
  const double bg1 = t1 * rbg1;
  const double ib1 = (gam + 1.0) * rbg1;  // 1 / beta
  const double beta = bg1 / (gam + 1.0);
  double t_fac, tp_fac, sp_fac;
  {
    double bb = beta;
    if (beta > 1.1 * q[G_FBC]) bb = q[G_FBC];
    else if (beta < q[G_FBMIN]) bb = q[G_FBMIN];
    t_fac = q[G_T0] + bb * (q[G_T1] + bb * (q[G_T2] + bb * (q[G_T3] + bb * (q[G_T4] + bb * q[G_T5]))));
    const double coef = bb * bb * c[GF_INV_TTF_DEN];
    tp_fac = coef * (q[G_T1] + bb * (2 * q[G_T2] + bb * (3 * q[G_T3] + bb * (4 * q[G_T4] + 5 * bb * q[G_T5]))));
    sp_fac = -coef * (q[G_S1] + bb * (2 * q[G_S2] + bb * (3 * q[G_S3] + bb * (4 * q[G_S4] + 5 * bb * q[G_S5]))));
  }
  if (q[G_AMP] == 0.0 || (ccl && beta < q[G_FBMIN])) {
    p.x = x + dd2 * xp;
    p.y = y + dd2 * yp;
    p.phi = phi + c[GF_DD_OVER_WL] * ib1;
    return;
  }
  const double gmg = c[GC_GMG], inv_gmg = c[GF_INV_GMG];
  const double inv_gmg2 = c[GC_INV_GMG2];
  const double concomm = c[GC_CONCOMM];
  const double con1 = c[GC_CON1];
  const double con2nex = c[GC_CON2NEX];
  const double inv_gmgcl = c[GC_INV_GMGCL], inv_fkis = c[GC_INV_FKIS];
  const double con5nex = concomm * (t_fac * inv_fkis - c[GC_X5] * tp_fac * inv_gmg2);
  double tw = t_fac;
  if (ccl) tw = t_fac - K_TWOPI * tp_fac * (c[GC_BETA_IN] * ib1 - 1.0) * c[GF_BETAG_OVER_BIN];
  const double con5nex1 = c[GC_CONCOMM1] * t_fac;
  const double con5nex2 = -con2nex * t_fac * inv_fkis;
  const double fkr = c[GC_FKR];
  const double sinps = c[GC_SINPS], cosps = c[GC_COSPS];
  double phi1, sinp1, cosp1;
  if (ccl) phi1 = phi + q[G_PHSHIFT] - c[GF_DG_OVER_WL] * ib1 + c[GC_PH_CONST];
  else phi1 = phi + q[G_PHSHIFT];
  sincos(phi1, &sinp1, &cosp1);
  const double prome1 = -con1 * (tp_fac * (sinp1 - sinps) + sp_fac * (cosp1 - cosps));
  sincos(phi1 + prome1, &sinp1, &cosp1);

  const double rs_p = x * x + y * y;
  const double rrp = x * xp + y * yp;
  double iris = 0.0, ris = 0.0;
  if (rs_p > 0.0) { iris = rsqrt(rs_p); ris = rs_p * iris; }
  const double fkrris = fkr * ris;
  double b0c, b1c;
  bessi01c(fkrris, b0c, b1c);
  const double b1p = (fkrris != 0.0) ? b0c - b1c * (iris * c[GF_INV_FKR]) : 0.5;
  const double bes1covr = (ris != 0.0) ? b1c * iris : fkr * 0.5;
  const double delte = c[GC_AQ_EFL] * (tw * b0c * cosp1 + ((tw - K_TWOPI * tp_fac) * inv_gmg * bes1covr + tw * inv_gmg * fkr * b1p -
                                                           tw * gmg * bes1covr) * rrp * sinp1);
  w += delte;
  gam = w * inv_mass;
  const double t2 = gam * (gam + 2.0);
  const double rbg = rsqrt(t2);
  const double ib = (gam + 1.0) * rbg;  // 1 / beta after the energy gain
  const double bgr = bg1 * rbg;
  const double onema = 1.0 - (con5nex * bes1covr + con5nex1 * b1p + con5nex2 * bes1covr) * cosp1;
  const double dovbl = con2nex * bes1covr * tw * sinp1;
  const double bgrovonema = bgr / onema;
  xp = -dovbl * x + bgrovonema * xp;
  x = x * onema + dd2 * xp;
  yp = -dovbl * y + bgrovonema * yp;
  y = y * onema + dd2 * yp;
  const double rpris = rrp * iris;  // iris == 0 when rs_p == 0, as in the reference
  const double prome = 2.0 * con1 * (-tp_fac * b0c + ris * tw * b1c * inv_gmgcl) * (sinp1 - sinps) -
                       c[GC_EFL] * inv_mass * tw * b1c * rpris * (cosp1 - cosps);
  double phi_new;
  if (ccl) phi_new = phi + (1.0 - c[GC_BETAG_CCL] * (c[GF_IBETA_OUT] - ib)) * c[GC_PI_CLBL_DPS] + prome;
  else phi_new = phi + (c[GC_BETA_OUT] * ib) * c[GC_PI_CLBL_DPS] + prome + c[GF_DCLQ_OVER_WL] * ib;
  if (!(w > 0.0) || x != x || y != y || xp != xp || yp != yp || phi_new != phi_new) {
    p.loss = 33333333u;
  } else {
    p.x = x; p.y = y; p.xp = xp; p.yp = yp; p.w = w; p.phi = phi_new;
  }
}

// fmod(t, 2 pi) for 0 <= t: the exact remainder (fmod never rounds) from one division estimate and one fused
// multiply-add instead of libdevice's bit-serial loop (12 % of the instructions of the moment sweep: every warp that
// holds one particle outside the bucket took it).  With n = floor(t / 2pi) possibly off by one, t - n * 2pi is
// computed with a single rounding; a result outside [0, 2pi) reveals the wrong n and the retry with the right one is
// exact because the true remainder is representable.  Same bits as fmod().
HPS_HD double fmod_twopi(double t) {
#ifdef __CUDA_ARCH__
  if (!(t < 4.0e9)) return fmod(t, K_TWOPI);  // n must stay exactly representable with room to spare (also NaN / inf)
  const double n = floor(t * (1.0 / K_TWOPI));
  double r = __fma_rn(-n, K_TWOPI, t);
  if (r < 0.0) r = __fma_rn(-(n - 1.0), K_TWOPI, t);
  else if (r >= K_TWOPI) r = __fma_rn(-(n + 1.0), K_TWOPI, t);
  return r;
#else
  return fmod(t, K_TWOPI);
#endif
}

// relative phase (beam_kernel.cu:1506-1522)
HPS_HD double rel_phi(double phi, double ref) {
  double tmp = phi - ref + K_PI;
  double sign = 1.0;
  if (tmp < 0) { sign = -1.0; tmp = -tmp; }
  // fmod(t, 2pi) == t exactly for 0 <= t < 2pi (the case of every particle inside the bunch)
  const double m = tmp < K_TWOPI ? tmp : fmod_twopi(tmp);
  return sign * (m - K_PI);
}

struct KickCtx {  // block-uniform inputs of the space-charge kick, staged in shared memory
  double phimc, xbar, ybar, epsq, phi_ref;
  double dr, dz, beta, ratio_r, ratio_z, ratio_q, ratio_g;
  u32 adj_bunch, num_on_mesh, nr, nz;
  // derived once per block (kick_prepare) with the reference's expression order
  double gmg, dr_eff, dzg, hl, c2, c3, xfac, yfac, pl, cbgcoef, nr_dr, er_c3, ez_c3;
  double inv_epsq, inv_dr, inv_dzg;  // reciprocals for the fast kick
  double wavg_dbg;                   // (tests/gpu_maps_check.cu only)
};

// particle-independent part of space_charge_kernel.cu:597-616, 680-681
__host__ __device__ inline void kick_prepare(KickCtx& k, const double* q, double mass, u32 n_global) {
  const double length = q[0], current = q[1], freq = q[2];
  const double beta = k.beta;
  const double gmg = rsqrt(1.0 - beta * beta);
  k.gmg = gmg;
  k.dr_eff = k.dr * k.ratio_r;
  k.dzg = k.dz * gmg * k.ratio_z;
  const double wave_len = K_CLIGHT / freq;
  k.hl = k.nz * 0.5 * k.dzg;
  const double c1 = 5721.67 * current / (freq * 1e6 * (double)n_global * mass);
  k.c2 = beta * gmg * wave_len / K_TWOPI;
  const double c3 = length / (beta * gmg);
  k.c3 = c3;
  k.xfac = 2.0 / (k.epsq + 1.0);
  k.yfac = k.epsq * k.xfac;
  k.pl = beta * gmg * wave_len;
  k.cbgcoef = c1 * c3 * K_PI * 0.5 * k.num_on_mesh;
  k.nr_dr = k.nr * k.dr_eff;
  const double er_ratio = k.ratio_q / (k.ratio_r * k.ratio_z * k.ratio_g * k.ratio_g);
  const double ez_ratio = k.ratio_q / (k.ratio_z * k.ratio_z);
  k.er_c3 = er_ratio * c3;
  k.ez_c3 = ez_ratio * c3;
  k.inv_epsq = 1.0 / k.epsq;
  k.inv_dr = 1.0 / k.dr_eff;
  k.inv_dzg = 1.0 / k.dzg;
}

// space_charge_kernel.cu:573-706
template <class M>
HPS_HD void op_sc_kick(Pt& p, u32 lloss, const KickCtx& k, const double2* __restrict__ fld, double mass, double inv_mass) {
  if (!(p.loss == 0 && lloss == 0)) return;
  const double beta = k.beta;
  const double phi = rel_phi(p.phi, k.phi_ref);
  const double gmg = k.gmg;
  const double dr = k.dr_eff;
  const double dzg = k.dzg;
  const double hl = k.hl;
  const double gm = M::udiv_(p.w, mass, inv_mass) + 1.0;
  const double bgz = M::sqrt_((gm - 1.0) * (gm + 1.0));
  double bgx = bgz * p.xp, bgy = bgz * p.yp;
  double bgzstar = gmg * (bgz - beta * gm);
  const double xxbar = p.x - k.xbar, yybar = p.y - k.ybar;
  double r = M::sqrt_(M::udiv_(xxbar * xxbar, k.epsq, k.inv_epsq) + yybar * yybar * k.epsq);
  if (r == 0.0) r = 1.0e-8;
  const double rr = M::rcp(r);  // x/r and y/r share the denominator
  const double xovr = M::udiv_(xxbar * k.xfac, r, rr), yovr = M::udiv_(yybar * k.yfac, r, rr);
  double ddz = phi - k.phimc;
  while (ddz > K_PI || ddz < -K_PI) {
    if (ddz > 0.0) ddz -= K_TWOPI; else ddz += K_TWOPI;
  }
  double z = -k.c2 * ddz;
  double cbgr, cbgzs;
  if (r > k.nr_dr || fabs(z) > hl) {  // off the mesh: point-charge field of the on-mesh charge (H17)
    z = -k.c2 * (phi - k.phimc);
    double d = M::sqrt_(z * z + r * r);
    double invd3 = M::rcp(d * d * d);
    double rod3 = r * invd3, zod3 = z * invd3;
    for (int nb = 1; nb <= (int)k.adj_bunch; ++nb) {
      int xi = nb;
      for (int lr = 1; lr <= 2; ++lr) {
        double s = z + xi * k.pl;
        d = M::sqrt_(s * s + r * r);
        invd3 = M::rcp(d * d * d);
        rod3 += r * invd3;
        zod3 += s * invd3;
        xi = -xi;
      }
    }
    cbgr = k.cbgcoef * rod3;
    cbgzs = k.cbgcoef * zod3;
  } else {
    const double ri = M::udiv_(r, dr, k.inv_dr), zj = M::udiv_(z + hl, dzg, k.inv_dzg);
    const u32 i = 1 + (int)ri;
    const double a = ri - (i - 1.0), b = 1.0 - a;
    const u32 j = 1 + (int)zj;
    const double c = zj - (j - 1.0), d = 1.0 - c;
    const u32 l = (j - 1) * (k.nr + 1) + i;
    u32 m = l + (k.nr + 1);
    const u32 tbl_sz = (k.nr + 1) * (k.nz + 1);
    if (m >= tbl_sz) m = tbl_sz - 1;
    const double2 fl = ld_field(&fld[l]), fl1 = ld_field(&fld[l - 1]), fm = ld_field(&fld[m]), fm1 = ld_field(&fld[m - 1]);
    cbgr = k.er_c3 * (d * (a * fl.x + b * fl1.x) + c * (a * fm.x + b * fm1.x));
    cbgzs = k.ez_c3 * (d * (a * fl.y + b * fl1.y) + c * (a * fm.y + b * fm1.y));
  }
  bgx += cbgr * xovr;
  bgy += cbgr * yovr;
  bgzstar += cbgzs;
  const double gstar = 1.0 + 0.5 * bgzstar * bgzstar;
  const double bgzf = gmg * (bgzstar + beta * gstar);
  if (bgzf <= 0.0) p.loss = 9999u;
  else {
    const double rbgzf = M::rcp(bgzf);  // shared by the two transverse quotients
    p.xp = M::udiv_(bgx, bgzf, rbgzf);
    p.yp = M::udiv_(bgy, bgzf, rbgzf);
    p.w *= 1.0 + M::div((gm + 1.0) * (bgzf - bgz), gm * bgz);
  }
}

// fast form of op_sc_kick: 1 sqrt-free rsqrt pair, 1 reciprocal and 1 division instead of 9 divisions + 2 sqrt
HPS_HD void op_sc_kick_f(Pt& p, u32 lloss, const KickCtx& k, const double2* __restrict__ fld, double inv_mass) {
  if (!(p.loss == 0 && lloss == 0)) return;
  const double beta = k.beta;
  const double phi = rel_phi(p.phi, k.phi_ref);
  const double gmg = k.gmg;
  const double hl = k.hl;
  const double gm = p.w * inv_mass + 1.0;
  const double tb = (gm - 1.0) * (gm + 1.0);
  const double rbgz = rsqrt(tb);
  const double bgz = tb * rbgz;
  double bgx = bgz * p.xp, bgy = bgz * p.yp;
  double bgzstar = gmg * (bgz - beta * gm);
  const double xxbar = p.x - k.xbar, yybar = p.y - k.ybar;
  const double r2 = xxbar * xxbar * k.inv_epsq + yybar * yybar * k.epsq;
  double r = 1.0e-8, rinv = 1.0e8;
  if (r2 > 0.0) { rinv = rsqrt(r2); r = r2 * rinv; }
  const double xovr = xxbar * k.xfac * rinv, yovr = yybar * k.yfac * rinv;
  double ddz = phi - k.phimc;
  while (ddz > K_PI || ddz < -K_PI) {
    if (ddz > 0.0) ddz -= K_TWOPI; else ddz += K_TWOPI;
  }
  double z = -k.c2 * ddz;
  double cbgr, cbgzs;
  if (r > k.nr_dr || fabs(z) > hl) {  // off the mesh: point-charge field of the on-mesh charge (H17)
    z = -k.c2 * (phi - k.phimc);
    const double rr = r * r;
    double rd = rsqrt(z * z + rr);
    double invd3 = rd * rd * rd;
    double rod3 = r * invd3, zod3 = z * invd3;
    for (int nb = 1; nb <= (int)k.adj_bunch; ++nb) {
      int xi = nb;
      for (int lr = 1; lr <= 2; ++lr) {
        const double s = z + xi * k.pl;
        rd = rsqrt(s * s + rr);
        invd3 = rd * rd * rd;
        rod3 += r * invd3;
        zod3 += s * invd3;
        xi = -xi;
      }
    }
    cbgr = k.cbgcoef * rod3;
    cbgzs = k.cbgcoef * zod3;
  } else {
    const double ri = r * k.inv_dr, zj = (z + hl) * k.inv_dzg;
    const u32 i = 1 + (int)ri;
    const double a = ri - (i - 1.0), b = 1.0 - a;
    const u32 j = 1 + (int)zj;
    const double c = zj - (j - 1.0), d = 1.0 - c;
    const u32 l = (j - 1) * (k.nr + 1) + i;
    u32 m = l + (k.nr + 1);
    const u32 tbl_sz = (k.nr + 1) * (k.nz + 1);
    if (m >= tbl_sz) m = tbl_sz - 1;
    const double2 fl = ld_field(&fld[l]), fl1 = ld_field(&fld[l - 1]), fm = ld_field(&fld[m]), fm1 = ld_field(&fld[m - 1]);
    cbgr = k.er_c3 * (d * (a * fl.x + b * fl1.x) + c * (a * fm.x + b * fm1.x));
    cbgzs = k.ez_c3 * (d * (a * fl.y + b * fl1.y) + c * (a * fm.y + b * fm1.y));
  }
  bgx += cbgr * xovr;
  bgy += cbgr * yovr;
  bgzstar += cbgzs;
  const double gstar = 1.0 + 0.5 * bgzstar * bgzstar;
  const double bgzf = gmg * (bgzstar + beta * gstar);
  if (bgzf <= 0.0) p.loss = 9999u;
  else {
    const double ibgzf = 1.0 / bgzf;
    p.xp = bgx * ibgzf;
    p.yp = bgy * ibgzf;
    p.w *= 1.0 + (gm + 1.0) * (bgzf - bgz) * rbgz / gm;
  }
}

}  // namespace hps
