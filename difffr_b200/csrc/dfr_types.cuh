// Device-visible data structures of one DiffDFSPH context.  See DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>

#include "dfr_math.cuh"

namespace dfr {

// Uniform cell grid shared by all point sets (fluid, static boundary, dynamic boundary).
// Cell edge >= support radius / reach (reach = 2: half-size cells and a 5x5x5 stencil, 15.6 h^3 of candidates
// instead of 27 h^3, and a finer sort order); coordinates are clamped into the grid, which keeps the stencil
// exhaustive for particles that leave the initial bounding box.
struct GridGeom {
  double ox, oy, oz;
  double inv_cell;
  int nx, ny, nz;
  int ncells;
  int reach;
  int z_shift;  // slab decomposition: local z cell = global z cell - z_shift (global origin and arithmetic on every rank)
};

// Solver parameters that never change during a trajectory (constant for all kernels).
struct Params {
  GridGeom grid;
  double support_radius, r2;       // h_s = 4 r (Simulation.cpp:382-386)
  double inv_h;                    // 1 / h_s
  double k_cubic, l_cubic, W_zero; // CubicKernel::setRadius (SPHKernels.h:25-34)
  double coh_k, coh_c, adh_k;      // Cohesion / Adhesion kernels (SPHKernels.h:451-458, 520-525)
  double particle_radius, density0, volume, mass;
  double gx, gy, gz;
  double cfl_factor, cfl_min, cfl_max;
  int cfl_method;
  int min_iter, max_iter, max_iter_v;
  double max_error, max_error_v;
  int use_warm_p, use_warm_v;
  int visc_method, st_method;
  double viscosity, viscosity_b, surface_tension, surface_tension_b;
  int gradient_mode, rigid_body_mode, optimize_rotation, use_manager, use_contact;
  double target_time, uniform_acc_time;
  int release_mode;  // useReleaseRigidBodyMode (TimeStepDiffDFSPH.cpp:381-407)
  double time_step_size0;
  int n_bodies, n_dyn_bodies;
  int slab;                // slab-decomposed context: residual sums and body accumulators are all-reduced between kernels
  long long n_global;      // fluid particles of the whole scene (slab mode)
};

// Mutable per-step scalars; lives in device memory so a step never needs the host.
struct StepState {
  double h;       // TimeManager::getTimeStepSize() ("NEW" once the CFL update ran)
  double h_step;  // the "OLD" h captured at the top of step() (TimeStepDiffDFSPH.cpp:535)
  double time;
  unsigned long long cfl_max_bits;  // max |v + a h|^2 as ordered bits (all values > 0)
  int nf;                           // active fluid particles held by this context (slab mode: owned + ghosts)
  int own_begin, own_end;           // sorted range this context computes (everything unless slab-decomposed)
  int step_count;
  int finished;
  int div_active, div_iters;
  int prs_active, prs_iters;
  int last_iters, last_iters_v;
  unsigned int ticket;              // last-block-done counter for residual reductions
  int error_flags;                  // bit0: fluid list overflow, bit1: boundary list overflow, bit2: D list overflow
  long long total_iters, total_iters_v, total_particle_steps, total_neighbors;
  long long nbr_entries_f, nbr_entries_b;   // of the current step (sum of counts)
  unsigned int list_used_f, list_used_b, list_used_d;
  double last_residual;
  double res_sum;                   // slab mode: local residual sum of the iteration, all-reduced before the stopping rule
  int slab_ranges[8];               // slab mode: own_begin, own_end, end of the low boundary layer, begin of the high one, nf
  // CUDA-graph stepping (dfr_api.cu: StepGraph): the solver loops are WHILE nodes, so what the host used to decide
  // between speculated batches is decided here
  int spec_div, div_streak;         // divergence iterations of the previous step / consecutive steps that matched the prediction
  int spec_div_prev;                // ... and of the step before that
  int fuse_now;                     // the next divergence iteration's k_rho also evaluates the non-pressure accelerations
  int np_done;                      // ... and that pass was the last active iteration: k_apply_accel instead of k_nonpressure
};

// Accumulator row written by the boundary-side kernel, per block: the eight net Jacobian blocks of
// BoundaryModel_Akinci2012.h:52-61 followed by force and torque.
enum {
  ACC_FV = 0,   // dF/dv   3x3
  ACC_FX = 9,   // dF/dx   3x3
  ACC_FQ = 18,  // dF/dq   3x4
  ACC_FW = 30,  // dF/dw   3x3
  ACC_TV = 39,  // dT/dv
  ACC_TX = 48,  // dT/dx
  ACC_TQ = 57,  // dT/dq   3x4
  ACC_TW = 69,  // dT/dw
  ACC_F = 78,
  ACC_T = 81,
  ACC_N = 84
};

struct BodyDev {
  int dynamic, animated;
  int p_begin, p_count;    // slice of the boundary particle arrays
  int blk_begin, blk_count;  // slice of the accumulator rows
  double mass, inv_mass;
  m33 I0, I, Iinv;
  d3 pos, vel, omega, pos0;
  quat q, q0;
  d3 init_v, init_omega;
  d3 force, torque;            // accumulated since the last clear (BoundaryModel.h:24-27)
  d3 force_last, torque_last;  // backup for getForce()/getTorque()
  // net Jacobians of the step (BoundaryModel_Akinci2012.h:52-61)
  m33 net_f_v, net_f_x, net_f_w, net_t_v, net_t_x, net_t_w;
  m34 net_f_q, net_t_q;
  // sensitivities (BoundaryModel_Akinci2012.h:63-77)
  m33 x_v0, x_w0, v_v0, v_w0, w_v0, w_w0;
  m43 q_v0, q_w0, partial_q_w;
};

// Box emitter (Emitter.cpp:40-87, 89-227; type 0), state advanced on the device
struct EmitterDev {
  int width, height;
  d3 x;
  m33 rot;
  double velocity, emit_start, emit_end;
  double next_emit_time;
  int emit_counter;
};

// RigidBodyGradientManager block (R, RR)  (RigidBodyGradientManager.h:62-91)
struct MgrBlock {
  m33 xn_v0, xn_w0, vn_v0, vn_w0, wn_v0, wn_w0;
  m43 qn_v0, qn_w0;
  m33 f_vn, f_xn, f_wn, t_vn, t_xn, t_wn;
  m34 f_qn, t_qn;
  m33 f_v0, f_w0, t_v0, t_w0;
};

// Warp-interleaved ELL neighbour list in groups of four slots ("ELL-4"): slots 4g..4g+3 of sorted particle i are
// one int4 at ((int4 *)idx)[((i >> 5) * (cap / 4) + g) * 32 + (i & 31)].  A warp reads 512 contiguous bytes per
// group with one LDG.128 per lane, and every loop over neighbours naturally works on batches of four: the four
// record gathers of a batch are issued back to back before any of them is consumed (memory-level parallelism;
// profiles/r1b showed the gathers latency-bound with one or two loads in flight per warp).
struct NbrList {
  const int *cnt;
  const int *idx;
  int cap;  // slots per particle, multiple of 4
};
__device__ __forceinline__ size_t nbr_slot(int cap, int i, int k) {
  return ((((size_t)(i >> 5) * (size_t)(cap >> 2) + (size_t)(k >> 2)) * 32 + (size_t)(i & 31)) << 2) + (size_t)(k & 3);
}
// load(j) -> payload gathered for neighbour j; use(payload, j) accumulates it.  Neighbours are consumed in slot
// order, so sums have the same order as a plain sequential loop.  Slots past the row's count are replaced by the
// always-valid index `safe` for the (discarded) gather.
// One ELL-4 group of a row.  DFR_LIST_NOALLOC=1 (tuning build): the index stream does not allocate in L1, so that it
// cannot evict the gathered records the warps of the SM share.
#ifndef DFR_LIST_NOALLOC
#define DFR_LIST_NOALLOC 0
#endif
__device__ __forceinline__ int4 ld_list4(const int4 *p) {
#if DFR_LIST_NOALLOC
  int4 r;
  asm("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
#else
  return __ldg(p);
#endif
}
template <int U = 4, class Load, class Use>
__device__ __forceinline__ void for_neighbors4(const NbrList &l, int i, int safe, Load load, Use use) {
  static_assert(U == 1 || U == 2 || U == 4, "sub-batch of the int4 group");
  const int n = l.cnt[i];
  if (n <= 0) return;
  const int4 *row = reinterpret_cast<const int4 *>(l.idx) + ((size_t)(i >> 5) * (size_t)(l.cap >> 2)) * 32 + (i & 31);
  const int nb = (n + 3) >> 2;
  int4 jn = ld_list4(row);
  for (int b = 0; b < nb; b++) {
    const int4 j4 = jn;
    if (b + 1 < nb) jn = ld_list4(row + (size_t)(b + 1) * 32);
    const int k0 = b << 2;
    int j[4] = {j4.x, j4.y, j4.z, j4.w};
#pragma unroll
    for (int u = 1; u < 4; u++)
      if (k0 + u >= n) j[u] = safe;
#pragma unroll
    for (int s = 0; s < 4; s += U) {  // U gathers in flight per record array (register budget of 2-/3-record kernels)
      decltype(load(0)) p[U];
#pragma unroll
      for (int u = 0; u < U; u++) p[u] = load(j[s + u]);
#pragma unroll
      for (int u = 0; u < U; u++)
        if (k0 + s + u < n) use(p[u], j[s + u]);
    }
  }
}

}  // namespace dfr
