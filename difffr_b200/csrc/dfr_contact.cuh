// Penalty rigid-rigid contact on the device (BASELINE.json configs[2], `useRigidContactSolver`):
// RigidContactSolver ctor (RigidContactSolver.cpp:23-263), beforePenaltyInitialize (:307-345),
// solveRigidContactPenalty (:419-555), update_rigid_body_gradient_manager (:1341-1370).
//
// Work split:
//   k_contact_rest      once: rest volume / rest density of every boundary particle (same-body neighbours)
//   k_contact_prepare   per step, one thread per dynamic boundary particle: particle velocity, contact flag, density
//   k_contact_force     per step, one thread per dynamic boundary particle in contact: penalty + friction force and the
//                       2 x 60 Jacobian entries of this particle, written to a 1 KB record
//   k_contact_apply     per step, one warp per dynamic body: walks the body's particles in the reference's storage order
//                       (every contacting particle calls addForce/addTorque, and addTorque re-applies the gyroscopic
//                       increment, Dynamic3dRigidBody.h:125-142 - the sequence is order dependent), lane 0 integrates the
//                       recurrence while all lanes sum the Jacobian records into shared memory in the same order
// Particles of static bodies are skipped: addForce on them is a no-op and their Jacobian rows are never read.
#pragma once
#include "dfr_kernels.cuh"

namespace dfr {

enum {
  CREC_F = 0,       // force (3)
  CREC_ACTIVE = 3,  // 1.0 if the particle applies a force this step
  CREC_RR = 4,      // the other body (as double)
  CREC_DIAG = 8,    // [R][R]:  f_x 9, t_x 9, f_q 12, t_q 12, f_v 9, f_w 9
  CREC_CROSS = 68,  // [R][RR]: same layout
  CREC_N = 128
};
enum { CG_FX = 0, CG_TX = 9, CG_FQ = 18, CG_TQ = 30, CG_FV = 42, CG_FW = 51, CG_N = 60 };

struct ContactParams {
  double inv_h, k_cubic;  // cubic kernel with the contact support radius (W_with_h, SPHKernels.h:57-73)
  double gamma, beta, mu;
};

__device__ __forceinline__ double contact_W(const ContactParams &C, double r2) {
  const double q = sqrt(r2) * C.inv_h;
  if (q > 1.0) return 0.0;
  if (q <= 0.5) {
    const double q2 = q * q;
    return C.k_cubic * (6.0 * q2 * q - 6.0 * q2 + 1.0);
  }
  const double f = 1.0 - q;
  return C.k_cubic * 2.0 * f * f * f;
}

// visits every boundary particle within the (fluid) support radius of x, except `self` (device index)
template <class F>
__device__ __forceinline__ void for_each_boundary_neighbor(const Params &P, const GridView &gs, const GridView &gd, int has_static, int has_dyn,
                                                           int n_static, double x, double y, double z, int self, F f) {
  if (has_static) for_each_in_range(P, gs, x, y, z, self < n_static ? self : -1, f);
  if (has_dyn) for_each_in_range(P, gd, x, y, z, self >= n_static ? self - n_static : -1, f);
}

// pass 0: vol0 = gamma / (W(0) + sum_same_body W);  pass 1: density0 = vol0 W(0) + sum_same_body vol0_k W   (:225-254)
__global__ void __launch_bounds__(128) k_contact_rest(const __grid_constant__ Params P, const __grid_constant__ ContactParams C, int pass,
                                                       const double4 *bpos, const int *bbody, int n_b, int n_static, GridView gs, GridView gd,
                                                       int has_static, int has_dyn, double *vol0, double *dens0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_b) return;
  const double4 p = bpos[b];
  const int body = bbody[b];
  const double W0 = C.k_cubic;
  double s = (pass == 0) ? W0 : vol0[b] * W0;
  for_each_boundary_neighbor(P, gs, gd, has_static, has_dyn, n_static, p.x, p.y, p.z, b, [&](int k) {
    if (bbody[k] != body) return;
    const double4 q = ldg4(bpos + k);
    const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
    const double w = contact_W(C, dx * dx + dy * dy + dz * dz);
    s += (pass == 0) ? w : vol0[k] * w;
  });
  if (pass == 0)
    vol0[b] = C.gamma / s;
  else
    dens0[b] = s;
}

// beforePenaltyInitialize (:307-345) for the particles of dynamic / animated bodies
__global__ void __launch_bounds__(128) k_contact_prepare(const __grid_constant__ Params P, const __grid_constant__ ContactParams C,
                                                          const BodyDev *bodies, const double4 *bpos, const int *bbody, int dyn_begin, int n_dyn,
                                                          int n_static, GridView gs, GridView gd, int has_static, const double *vol0,
                                                          double4 *cvel, double *cdens) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_dyn) return;
  const int j = dyn_begin + t;
  const double4 p = bpos[j];
  const int body = bbody[j];
  const BodyDev &B = bodies[body];
  const d3 v = B.vel + cross(B.omega, mk3(p.x, p.y, p.z) - B.pos);
  cvel[t] = make_double4(v.x, v.y, v.z, 0.0);
  bool has_contact = false;
  double dens = vol0[j] * C.k_cubic;
  for_each_boundary_neighbor(P, gs, gd, has_static, 1, n_static, p.x, p.y, p.z, j, [&](int k) {
    if (bbody[k] != body) has_contact = true;
    const double4 q = ldg4(bpos + k);
    const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
    dens += vol0[k] * contact_W(C, dx * dx + dy * dy + dz * dz);
  });
  cdens[t] = has_contact ? dens : -1.0;
}

__device__ __forceinline__ void crec_store33(double *rec, int off, const m33 &m) {
#pragma unroll
  for (int k = 0; k < 9; k++) rec[off + k] = m.a[k];
}
__device__ __forceinline__ void crec_store34(double *rec, int off, const m34 &m) {
#pragma unroll
  for (int k = 0; k < 12; k++) rec[off + k] = m.a[k];
}
// column vector (3) times row vector (4)
__device__ __forceinline__ m34 outer34(d3 a, const double b[4]) {
  m34 m;
  const double av[3] = {a.x, a.y, a.z};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int jj = 0; jj < 4; jj++) m.a[i * 4 + jj] = av[i] * b[jj];
  return m;
}

// solveRigidContactPenalty (:419-555), per particle; the sums over particles are done in order by k_contact_apply
__global__ void __launch_bounds__(128) k_contact_force(const __grid_constant__ Params P, const __grid_constant__ ContactParams C,
                                                        const BodyDev *bodies, const double4 *bpos, const double4 *bx0, const int *bbody,
                                                        int dyn_begin, int n_dyn, int n_static, GridView gs, GridView gd, int has_static,
                                                        const double *vol0, const double *dens0, const double4 *cvel, const double *cdens,
                                                        double *records) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_dyn) return;
  double *rec = records + (size_t)t * CREC_N;
  rec[CREC_ACTIVE] = 0.0;
  const double dens = cdens[t];
  if (dens < 0.0) return;  // no particle of another body in range
  const int j = dyn_begin + t;
  const double ratio = dens / dens0[j];
  if (!(ratio > 1.0)) return;
  const double4 p = bpos[j];
  const int body = bbody[j];
  const BodyDev &B = bodies[body];
  const d3 x_r = mk3(p.x, p.y, p.z);
  const d3 r_r = x_r - B.pos;
  const double4 x04 = bx0[j];
  const d3 r_r0 = mk3(x04.x, x04.y, x04.z);
  const m34 Qr = grad_Rqp_to_q(B.q, r_r0, -1.0);
  d3 sum_x = mk3(0, 0, 0), sum_vel_k = mk3(0, 0, 0), avg_r_k = mk3(0, 0, 0), gdx = mk3(0, 0, 0);
  double sum_w = 0.0;
  double gdq[4] = {0, 0, 0, 0}, gdq_k[4] = {0, 0, 0, 0};
  int RR = body;
  for_each_boundary_neighbor(P, gs, gd, has_static, 1, n_static, p.x, p.y, p.z, j, [&](int k) {
    const double4 q = ldg4(bpos + k);
    const d3 x_k = mk3(q.x, q.y, q.z);
    const d3 d = x_r - x_k;
    const double w = contact_W(C, dot(d, d));
    const int bk = bbody[k];
    if (bk == body) {
      sum_x += w * x_k;
      sum_w += w;
    } else {
      RR = max(RR == body ? -1 : RR, bk);  // the reference keeps the last (= highest) other point set that has neighbours
      const BodyDev &O = bodies[bk];
      const d3 r_k = x_k - O.pos;
      const double4 k04 = bx0[k];
      d3 vk = mk3(0, 0, 0);
      if (k >= dyn_begin) {
        const double4 v4k = cvel[k - dyn_begin];
        vk = mk3(v4k.x, v4k.y, v4k.z);
      }
      sum_vel_k += w * vk;
      avg_r_k += w * r_k;
      const d3 gW = cubic_gradW(P, d);  // sim->gradW: the FLUID support radius (:466)
      const double vk0 = vol0[k];
      gdx += vk0 * gW;
      const d3 a = vk0 * gW;
      const m34 Qk = grad_Rqp_to_q(O.q, mk3(k04.x, k04.y, k04.z), -1.0);
#pragma unroll
      for (int cc = 0; cc < 4; cc++) {
        gdq[cc] += a.x * Qr.a[cc] + a.y * Qr.a[4 + cc] + a.z * Qr.a[8 + cc];
        gdq_k[cc] += -(a.x * Qk.a[cc] + a.y * Qk.a[4 + cc] + a.z * Qk.a[8 + cc]);
      }
    }
  });
  const d3 normal_r = x_r - mk3(sum_x.x / sum_w, sum_x.y / sum_w, sum_x.z / sum_w);
  const double4 cv = cvel[t];
  const d3 vel_rel = mk3(cv.x, cv.y, cv.z) - mk3(sum_vel_k.x / sum_w, sum_vel_k.y / sum_w, sum_vel_k.z / sum_w);
  avg_r_k = mk3(avg_r_k.x / sum_w, avg_r_k.y / sum_w, avg_r_k.z / sum_w);
  const double pen = fmax(ratio - 1.0, 0.0);
  const d3 normal_force = (-C.beta * pen) * normal_r;
  const double vn = sqrt(dot(vel_rel, vel_rel));
  const d3 unit = (vn * vn > 0.0) ? mk3(vel_rel.x / vn, vel_rel.y / vn, vel_rel.z / vn) : vel_rel;  // Eigen normalized()
  const double nfn = sqrt(dot(normal_force, normal_force));
  const d3 friction = (-C.mu * nfn) * unit;
  const d3 f = normal_force + friction;
  rec[CREC_F + 0] = f.x;
  rec[CREC_F + 1] = f.y;
  rec[CREC_F + 2] = f.z;
  rec[CREC_ACTIVE] = 1.0;
  rec[CREC_RR] = (double)RR;
  // ---- Jacobians (:487-548) ----
  const d3 g_nf_dens = (-C.beta) * normal_r;
  m33 Mf = outer((-C.mu) * unit, normal_force);
#pragma unroll
  for (int k = 0; k < 9; k++) Mf.a[k] = Mf.a[k] / nfn;
  const d3 g_ff_dens = Mf * g_nf_dens;
  const d3 g_f_dens = g_nf_dens + g_ff_dens;  // d(normal + friction)/d density (both terms multiply the same row vectors)
  const m33 uu = outer(unit, unit);
  m33 g_ff_vel;
  const double cfv = -C.mu * nfn / vn;
#pragma unroll
  for (int k = 0; k < 9; k++) g_ff_vel.a[k] = cfv * ((k % 4 == 0 ? 1.0 : 0.0) - uu.a[k]);
  const m33 g_ff_omega = g_ff_vel * transpose(skew(r_r));
  const m33 g_ff_omega_k = ((-1.0) * g_ff_vel) * transpose(skew(avg_r_k));
  // grad_f_to_x = grad_normal_force_to_x + grad_friction_force_to_x, each a column times grad_density_to_x^T
  const m33 g_f_x = outer(g_nf_dens, gdx) + outer(g_ff_dens, gdx);
  const m34 g_f_q = outer34(g_nf_dens, gdq) + outer34(g_ff_dens, gdq);
  const m34 g_f_q_k = outer34(g_nf_dens, gdq_k) + outer34(g_ff_dens, gdq_k);
  (void)g_f_dens;
  const m33 Sr = skew(r_r);
  double *dg = rec + CREC_DIAG, *cg = rec + CREC_CROSS;
  crec_store33(dg, CG_FX, g_f_x);
  crec_store33(dg, CG_TX, Sr * g_f_x);
  crec_store34(dg, CG_FQ, g_f_q);
  crec_store34(dg, CG_TQ, Sr * g_f_q + transpose(skew(f)) * Qr);
  crec_store33(dg, CG_FV, g_ff_vel);
  crec_store33(dg, CG_FW, g_ff_omega);
  if (RR != body) {
    crec_store33(cg, CG_FX, (-1.0) * g_f_x);
    crec_store33(cg, CG_TX, (-1.0) * (Sr * g_f_x));
    crec_store34(cg, CG_FQ, g_f_q_k);
    crec_store34(cg, CG_TQ, Sr * g_f_q_k);
    crec_store33(cg, CG_FV, (-1.0) * g_ff_vel);
    crec_store33(cg, CG_FW, g_ff_omega_k);
  }
}

// One warp per body.  `order` holds, per dynamic body, its particles (relative index within the dynamic range) in the
// reference's storage order.  Dynamic shared memory: (1 + n_bodies) * CG_N doubles.
__global__ void __launch_bounds__(32) k_contact_apply(const __grid_constant__ Params P, const StepState *st, BodyDev *bodies, MgrBlock *M,
                                                       const double4 *bpos, int dyn_begin, const int *order, const double *records) {
  extern __shared__ double cacc[];
  const int R = blockIdx.x;
  BodyDev &B = bodies[R];
  if (!B.dynamic) return;
  const int lane = threadIdx.x;
  const int n = P.n_bodies;
  for (int k = lane; k < (1 + n) * CG_N; k += 32) cacc[k] = 0.0;
  __syncwarp();
  const double h = st->h;
  const int first = B.p_begin - dyn_begin;
  // lane 0 integrates v and omega; the inertia is constant during this phase
  d3 vel = B.vel, omega = B.omega;
  const m33 I = B.I, Iinv = B.Iinv;
  const d3 pos = B.pos;
  const double inv_mass = B.inv_mass;
  const bool gyro = (P.rigid_body_mode == 0);
  for (int s0 = 0; s0 < B.p_count; s0 += 32) {
    const int s = s0 + lane;
    int id = -1;
    bool active = false;
    if (s < B.p_count) {
      id = order[first + s];
      active = records[(size_t)id * CREC_N + CREC_ACTIVE] != 0.0;
    }
    unsigned int mask = __ballot_sync(DFR_FULL, active);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const int pid = __shfl_sync(DFR_FULL, id, src);
      const double *rec = records + (size_t)pid * CREC_N;
      const int RR = (int)rec[CREC_RR];
      for (int k = lane; k < CG_N; k += 32) {
        cacc[k] += rec[CREC_DIAG + k];
        if (RR != R) cacc[(1 + RR) * CG_N + k] += rec[CREC_CROSS + k];
      }
      if (lane == 0) {
        const d3 f = mk3(rec[CREC_F], rec[CREC_F + 1], rec[CREC_F + 2]);
        const double4 p = bpos[dyn_begin + pid];
        const d3 tq = cross(mk3(p.x, p.y, p.z) - pos, f);
        vel += (inv_mass * f) * h;  // Dynamic3dRigidBody::addForce (no isAnimated test on this path)
        if (gyro) {
          const d3 L = I * omega;
          omega += (Iinv * (cross(L, omega) + tq)) * h;
        } else
          omega += (Iinv * tq) * h;
      }
    }
  }
  __syncwarp();
  if (lane == 0) {
    B.vel = vel;
    B.omega = omega;
  }
  if (!P.use_manager) return;
  // update_rigid_body_gradient_manager (:1341-1370): overwrite the per-step blocks of every dynamic pair
  for (int RR = 0; RR < n; RR++) {
    if (!bodies[RR].dynamic) continue;
    const double *a = (RR == R) ? cacc : cacc + (1 + RR) * CG_N;
    MgrBlock &o = M[R * n + RR];
    for (int k = lane; k < CG_N; k += 32) {
      const double v = a[k];
      if (k < CG_TX) o.f_xn.a[k - CG_FX] = v;
      else if (k < CG_FQ) o.t_xn.a[k - CG_TX] = v;
      else if (k < CG_TQ) o.f_qn.a[k - CG_FQ] = v;
      else if (k < CG_FV) o.t_qn.a[k - CG_TQ] = v;
      else if (k < CG_FW) o.f_vn.a[k - CG_FV] = v;
      else o.f_wn.a[k - CG_FW] = v;
    }
    for (int k = lane; k < 9; k += 32) {  // never accumulated by the penalty solver
      o.t_vn.a[k] = 0.0;
      o.t_wn.a[k] = 0.0;
    }
  }
}

}  // namespace dfr
