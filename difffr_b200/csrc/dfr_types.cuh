// Device-visible data structures of one DiffDFSPH context.  See DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>

#include "dfr_math.cuh"

namespace dfr {

// Uniform cell grid shared by all point sets (fluid, static boundary, dynamic boundary).
// Cell edge >= support radius; coordinates are clamped into the grid, which keeps the 27-cell
// stencil exhaustive for particles that leave the initial bounding box.
struct GridGeom {
  double ox, oy, oz;
  double inv_cell;
  int nx, ny, nz;
  int ncells;
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
  double time_step_size0;
  int n_bodies, n_dyn_bodies;
};

// Mutable per-step scalars; lives in device memory so a step never needs the host.
struct StepState {
  double h;       // TimeManager::getTimeStepSize() ("NEW" once the CFL update ran)
  double h_step;  // the "OLD" h captured at the top of step() (TimeStepDiffDFSPH.cpp:535)
  double time;
  unsigned long long cfl_max_bits;  // max |v + a h|^2 as ordered bits (all values > 0)
  int nf;                           // active fluid particles
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

// RigidBodyGradientManager block (R, RR)  (RigidBodyGradientManager.h:62-91)
struct MgrBlock {
  m33 xn_v0, xn_w0, vn_v0, vn_w0, wn_v0, wn_w0;
  m43 qn_v0, qn_w0;
  m33 f_vn, f_xn, f_wn, t_vn, t_xn, t_wn;
  m34 f_qn, t_qn;
  m33 f_v0, f_w0, t_v0, t_w0;
};

// Warp-interleaved ELL neighbour list: the k-th neighbour of sorted particle i sits at
// idx[((i >> 5) * cap + k) * 32 + (i & 31)], so that a warp reads 128 contiguous bytes per k.
struct NbrList {
  const int *cnt;
  const int *idx;
  int cap;
};
__device__ __forceinline__ const int *nbr_row(const NbrList &l, int i) { return l.idx + ((size_t)(i >> 5) * l.cap) * 32 + (i & 31); }

}  // namespace dfr
