// Per-body device code: begin/end of step bookkeeping, the sensitivity chain rule
// (BoundaryModel_Akinci2012::perform_chain_rule / RigidBodyGradientManager) and the rigid integrator.
// O(1) work per body; runs in one small kernel so that a step never returns to the host.
#pragma once
#include "dfr_types.cuh"

namespace dfr {

// TimeStepDiffDFSPH::beginStep (TimeStepDiffDFSPH.cpp:353-430)
__global__ void k_begin_step(const __grid_constant__ Params P, StepState *st, BodyDev *bodies, int fuse_mode) {
  const int b = threadIdx.x;
  if (b == 0) {
    st->step_count += 1;
    // graph stepping (fuse_mode 1): does the first divergence iteration carry the non-pressure pass?
    st->np_done = 0;
    st->fuse_now = (fuse_mode == 1 && min(st->spec_div, st->spec_div_prev) <= 1) ? 1 : 0;
    st->h_step = st->h;  // "const Real h = tm->getTimeStepSize()" at :535
    st->div_active = 1;
    st->div_iters = 0;
    st->ticket = 0;
    st->cfl_max_bits = 0ull;
    st->list_used_f = st->list_used_b = 0u;  // longest rows of THIS step's neighbour build (capacity watch, dfr_api.cu)
    if (!P.slab) {  // one context holds everything (emitters may have grown nf at the end of the last step)
      st->own_begin = 0;
      st->own_end = st->nf;
    }
  }
  if (b < P.n_bodies) {
    BodyDev &B = bodies[b];
    const double T = P.uniform_acc_time;
    if (B.dynamic && P.release_mode) {
      // :381-407: only boundary model 1 (the acting body) is touched - held at rest, then released; the reference sets its
      // velocities back to the initial ones at the start of EVERY later step, and so does this
      if (b == 1) {
        if (st->time <= T + st->h) {
          if (T > 1e-3) B.animated = 1;
          B.vel = mk3(0.0, 0.0, 0.0);
          B.omega = mk3(0.0, 0.0, 0.0);
        } else {
          B.vel = B.init_v;
          B.omega = B.init_omega;
          B.animated = 0;
        }
      }
    } else if (B.dynamic) {
      if (st->time <= T + st->h) {
        double factor = 1.0;
        if (T > 1e-3) {
          factor = (st->time / T) > 1.0 ? 1.0 : (st->time / T);
          B.animated = 1;
        }
        B.vel = factor * B.init_v;
        B.omega = factor * B.init_omega;
      } else
        B.animated = 0;
    }
  }
}

// BoundaryModel_Akinci2012::reset_gradient (BoundaryModel_Akinci2012.cpp:62-108), non-articulated branch
__global__ void k_reset_gradient(const __grid_constant__ Params P, BodyDev *bodies) {
  const int b = threadIdx.x;
  if (b >= P.n_bodies) return;
  BodyDev &B = bodies[b];
  if (!B.dynamic) return;
  B.net_f_v = B.net_f_x = B.net_f_w = B.net_t_v = B.net_t_x = B.net_t_w = m33::zero();
  B.net_f_q = B.net_t_q = m34::zero();
  B.v_v0 = m33::identity();
  B.v_w0 = m33::zero();
  B.w_w0 = m33::identity();
  B.w_v0 = m33::zero();
  B.q_w0 = B.q_v0 = B.partial_q_w = m43::zero();
  B.x_v0 = B.x_w0 = m33::zero();
}

// sums the accumulator rows of each body in block order (deterministic) and clears them
// (replaces accumulate_and_reset_gradient, BoundaryModel_Akinci2012.cpp:453-497, and the per-thread
// force slots of BoundaryModel.cpp:38-54)
#define BR_SLICES 8
// Sum (and clear) the accumulator rows of one body into tot[ACC_N] (shared memory), whole block.  Rows are dealt to
// BR_SLICES slices (row r -> slice r % BR_SLICES), each summed in row order, slices added in order: a fixed shape, so
// the sums are reproducible, with BR_SLICES independent chains per column instead of one serial walk over all rows.
__device__ __forceinline__ void body_rows_sum(const BodyDev &B, double *acc_rows, double (*part)[ACC_N], double *tot) {
  for (int t = threadIdx.x; t < ACC_N * BR_SLICES; t += blockDim.x) {
    const int k = t % ACC_N, sl = t / ACC_N;
    double s = 0.0;
    if (B.dynamic)
      for (int r = sl; r < B.blk_count; r += BR_SLICES) {
        double *p = acc_rows + (size_t)(B.blk_begin + r) * ACC_N + k;
        s += *p;
        *p = 0.0;
      }
    part[sl][k] = s;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < ACC_N; k += blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int sl = 0; sl < BR_SLICES; sl++) s += part[sl][k];
    tot[k] = s;
  }
  __syncthreads();
}
__global__ void k_body_reduce(BodyDev *bodies, double *acc_rows) {
  BodyDev &B = bodies[blockIdx.x];
  if (!B.dynamic) return;
  __shared__ double tot[ACC_N];
  __shared__ double part[BR_SLICES][ACC_N];
  body_rows_sum(B, acc_rows, part, tot);
  if (threadIdx.x == 0) {
    B.force += mk3(tot[ACC_F], tot[ACC_F + 1], tot[ACC_F + 2]);
    B.torque += mk3(tot[ACC_T], tot[ACC_T + 1], tot[ACC_T + 2]);
    if (!B.animated) {
      for (int k = 0; k < 9; k++) {
        B.net_f_v.a[k] = tot[ACC_FV + k];
        B.net_f_x.a[k] = tot[ACC_FX + k];
        B.net_f_w.a[k] = tot[ACC_FW + k];
        B.net_t_v.a[k] = tot[ACC_TV + k];
        B.net_t_x.a[k] = tot[ACC_TX + k];
        B.net_t_w.a[k] = tot[ACC_TW + k];
      }
      for (int k = 0; k < 12; k++) {
        B.net_f_q.a[k] = tot[ACC_FQ + k];
        B.net_t_q.a[k] = tot[ACC_TQ + k];
      }
    }
  }
}

// compute_grad_inertia_v_to_* (BoundaryModel_Akinci2012.cpp:500-541, RigidBodyGradientManager.cpp:241-288)
__device__ inline m33 grad_inertia_v(const BodyDev &B, const m43 &dq, d3 v) {
  const m33 R = qrot(B.q);
  const d3 u = B.I0 * (transpose(R) * v);
  const m33 a = grad_Rqp_to_q(B.q, u, -1.0) * dq;
  const m33 bb = (R * B.I0) * (grad_Rqp_to_q(B.q, v, +1.0) * dq);
  return a + bb;
}

struct OmegaChain {
  d3 temp_v;
};

// the omega-sensitivity update shared by all modes (:622-650, :689-701, :769-790; manager :318-347)
__device__ inline void omega_chain(const Params &P, const BodyDev &B, d3 torque, bool with_torque, const m33 &t_w0, const m33 &t_v0,
                                   const m43 &q_w0, const m43 &q_v0, double dt, m33 &w_w0, m33 &w_v0, d3 &temp_v) {
  const bool gyro = (P.rigid_body_mode == 0);
  const d3 omega = B.omega;
  m33 Tau_w0, Tau_v0;
  if (gyro) {
    const d3 L = B.I * omega;
    temp_v = cross(L, omega);
    if (with_torque) temp_v += torque;
    const m33 SL = skew(L), SwT = transpose(skew(omega));
    const m33 L_w0 = SL * w_w0 + SwT * (B.I * w_w0 + grad_inertia_v(B, q_w0, omega));
    const m33 L_v0 = SL * w_v0 + SwT * (B.I * w_v0 + grad_inertia_v(B, q_v0, omega));
    Tau_w0 = L_w0 + t_w0;
    Tau_v0 = L_v0 + t_v0;
  } else {
    temp_v = with_torque ? torque : mk3(0, 0, 0);
    Tau_w0 = t_w0;
    Tau_v0 = t_v0;
  }
  const d3 u = B.Iinv * temp_v;
  const m33 gi_w0 = B.Iinv * grad_inertia_v(B, q_w0, u);
  const m33 gi_v0 = B.Iinv * grad_inertia_v(B, q_v0, u);
  w_w0 = w_w0 + dt * (gi_w0 + B.Iinv * Tau_w0);
  w_v0 = w_v0 + dt * (gi_v0 + B.Iinv * Tau_v0);
}

// quaternion-integration Jacobian (:827-870; manager :395-453)
__device__ inline void rotation_chain(const BodyDev &B, d3 new_omega, double dt, const m33 &w_w0, const m33 &w_v0, m43 &q_w0, m43 &q_v0,
                                      m43 *partial) {
  const quat q = B.q;
  quat p;
  p.w = 0.0;
  p.x = new_omega.x;
  p.y = new_omega.y;
  p.z = new_omega.z;
  const quat pq = qmul(p, q);
  quat nq;
  nq.w = q.w + dt * 0.5 * pq.w;
  nq.x = q.x + dt * 0.5 * pq.x;
  nq.y = q.y + dt * 0.5 * pq.y;
  nq.z = q.z + dt * 0.5 * pq.z;
  const double nn = qnorm(nq);
  v4 qn;
  qn.a[0] = nq.w / nn;
  qn.a[1] = nq.x / nn;
  qn.a[2] = nq.y / nn;
  qn.a[3] = nq.z / nn;
  const m44 gpq = grad_pq_to_q(p);
  const m43 gpo = grad_omega_q_to_omega(q);
  m44 gn = m44::identity() - qn * transpose(qn);
  gn = (1.0 / nn) * gn;
  if (partial) *partial = gn * ((dt / 2.0) * gpo);
  q_w0 = gn * (q_w0 + (dt / 2.0) * (gpo * w_w0 + gpq * q_w0));
  q_v0 = gn * (q_v0 + (dt / 2.0) * (gpo * w_v0 + gpq * q_v0));
}

// BoundaryModel_Akinci2012::perform_chain_rule (BoundaryModel_Akinci2012.cpp:543-889)
__device__ inline void perform_chain_rule(const Params &P, BodyDev &B, double dt) {
  const double invMass = 1.0 / (B.mass + 1e-10);
  d3 temp_v = mk3(0, 0, 0);
  if (P.gradient_mode == 0) {  // Complete
    const m33 f_v0 = B.net_f_x * B.x_v0 + B.net_f_v * B.v_v0 + B.net_f_q * B.q_v0 + B.net_f_w * B.w_v0;
    const m33 f_w0 = B.net_f_x * B.x_w0 + B.net_f_v * B.v_w0 + B.net_f_q * B.q_w0 + B.net_f_w * B.w_w0;
    B.v_v0 = B.v_v0 + (dt * invMass) * f_v0;
    B.v_w0 = B.v_w0 + (dt * invMass) * f_w0;
    B.x_v0 = B.x_v0 + dt * B.v_v0;
    B.x_w0 = B.x_w0 + dt * B.v_w0;
    const m33 t_w0 = B.net_t_x * B.x_w0 + B.net_t_v * B.v_w0 + B.net_t_q * B.q_w0 + B.net_t_w * B.w_w0;
    const m33 t_v0 = B.net_t_x * B.x_v0 + B.net_t_v * B.v_v0 + B.net_t_q * B.q_v0 + B.net_t_w * B.w_v0;
    omega_chain(P, B, B.torque, true, t_w0, t_v0, B.q_w0, B.q_v0, dt, B.w_w0, B.w_v0, temp_v);
  } else if (P.gradient_mode == 2) {  // RigidGradOnly
    B.x_v0 = B.x_v0 + dt * B.v_v0;
    B.x_w0 = B.x_w0 + dt * B.v_w0;
    if (P.rigid_body_mode == 0) {
      const m33 z = m33::zero();
      omega_chain(P, B, B.torque, false, z, z, B.q_w0, B.q_v0, dt, B.w_w0, B.w_v0, temp_v);
    }
  } else {  // Incomplete
    const m33 f_v0 = B.net_f_v * B.v_v0 + B.net_f_w * B.w_v0;
    const m33 f_w0 = B.net_f_v * B.v_w0 + B.net_f_w * B.w_w0;
    B.v_v0 = B.v_v0 + (dt * invMass) * f_v0;
    B.v_w0 = B.v_w0 + (dt * invMass) * f_w0;
    B.x_v0 = B.x_v0 + dt * B.v_v0;
    B.x_w0 = B.x_w0 + dt * B.v_w0;
    const m33 t_w0 = B.net_t_v * B.v_w0 + B.net_t_w * B.w_w0;
    const m33 t_v0 = B.net_t_v * B.v_v0 + B.net_t_w * B.w_v0;
    omega_chain(P, B, B.torque, true, t_w0, t_v0, B.q_w0, B.q_v0, dt, B.w_w0, B.w_v0, temp_v);
  }
  if (P.optimize_rotation) {
    const d3 new_omega = B.omega + dt * (B.Iinv * temp_v);
    rotation_chain(B, new_omega, dt, B.w_w0, B.w_v0, B.q_w0, B.q_v0, &B.partial_q_w);
  }
}

// Dynamic3dRigidBody::addForce / addTorque / animate / updateInertia (Dynamic3dRigidBody.h:113-154, 215-219)
__device__ inline void rb_add_force(BodyDev &B, d3 f, double dt) { B.vel += (B.inv_mass * f) * dt; }
__device__ inline void rb_add_torque(const Params &P, BodyDev &B, d3 t, double dt) {
  if (P.rigid_body_mode == 0) {
    const d3 L = B.I * B.omega;
    B.omega += (B.Iinv * (cross(L, B.omega) + t)) * dt;
  } else
    B.omega += (B.Iinv * t) * dt;
}
__device__ inline void rb_update_inertia(BodyDev &B) {
  const m33 R = qrot(B.q);
  B.I = R * B.I0 * transpose(R);
  B.Iinv = inverse(B.I);
}
__device__ inline void rb_animate(BodyDev &B, double dt) {
  B.pos += B.vel * dt;
  quat w;
  w.w = 0.0;
  w.x = B.omega.x;
  w.y = B.omega.y;
  w.z = B.omega.z;
  const quat d = qmul(w, B.q);
  quat nq;
  nq.w = B.q.w + dt * 0.5 * d.w;
  nq.x = B.q.x + dt * 0.5 * d.x;
  nq.y = B.q.y + dt * 0.5 * d.y;
  nq.z = B.q.z + dt * 0.5 * d.z;
  const double n = qnorm(nq);
  B.q.w = nq.w / n;
  B.q.x = nq.x / n;
  B.q.y = nq.y / n;
  B.q.z = nq.z / n;
  rb_update_inertia(B);
}

// ---- RigidBodyGradientManager (RigidBodyGradientManager.cpp:96-475), run by one thread ----
__device__ inline void mgr_force_torque_chain(const Params &P, const BodyDev *bodies, MgrBlock *M, bool rigid) {
  const int n = P.n_bodies;
  for (int k = 0; k < n * n; k++) M[k].f_v0 = M[k].f_w0 = M[k].t_v0 = M[k].t_w0 = m33::zero();
  for (int R = 0; R < n; R++) {
    if (!bodies[R].dynamic) continue;
    for (int RR = 0; RR < n; RR++) {
      if (!bodies[RR].dynamic) continue;
      for (int Rk = 0; Rk < n; Rk++) {
        if (!bodies[Rk].dynamic) continue;
        const MgrBlock &a = M[R * n + Rk];
        const MgrBlock &b = M[Rk * n + RR];
        MgrBlock &o = M[R * n + RR];
        if (!rigid) {
          o.f_v0 += a.f_vn * b.vn_v0 + a.f_wn * b.wn_v0;
          o.t_v0 += a.t_vn * b.vn_v0 + a.t_wn * b.wn_v0;
          o.f_w0 += a.f_vn * b.vn_w0 + a.f_wn * b.wn_w0;
          o.t_w0 += a.t_vn * b.vn_w0 + a.t_wn * b.wn_w0;
        } else {
          // RigidBodyGradientManager.cpp:199-200, 204-205: a stray ';' drops the q-term from the v0 blocks only
          o.f_v0 += a.f_vn * b.vn_v0 + a.f_wn * b.wn_v0 + a.f_xn * b.xn_v0;
          o.t_v0 += a.t_vn * b.vn_v0 + a.t_wn * b.wn_v0 + a.t_xn * b.xn_v0;
          o.f_w0 += a.f_vn * b.vn_w0 + a.f_wn * b.wn_w0 + a.f_xn * b.xn_w0 + a.f_qn * b.qn_w0;
          o.t_w0 += a.t_vn * b.vn_w0 + a.t_wn * b.wn_w0 + a.t_xn * b.xn_w0 + a.t_qn * b.qn_w0;
        }
      }
    }
  }
}
__device__ inline void mgr_velocity_chain(const Params &P, BodyDev *bodies, MgrBlock *M, double dt) {
  const int n = P.n_bodies;
  for (int R = 0; R < n; R++) {
    BodyDev &B = bodies[R];
    if (!B.dynamic) continue;
    for (int RR = 0; RR < n; RR++) {
      if (!bodies[RR].dynamic) continue;
      MgrBlock &o = M[R * n + RR];
      o.vn_v0 = o.vn_v0 + (dt * B.inv_mass) * o.f_v0;
      o.vn_w0 = o.vn_w0 + (dt * B.inv_mass) * o.f_w0;
      d3 temp_v;
      omega_chain(P, B, B.torque, true, o.t_w0, o.t_v0, o.qn_w0, o.qn_v0, dt, o.wn_w0, o.wn_v0, temp_v);
    }
  }
}
__device__ inline void mgr_position_rotation_chain(const Params &P, BodyDev *bodies, MgrBlock *M, double dt) {
  const int n = P.n_bodies;
  for (int R = 0; R < n; R++) {
    BodyDev &B = bodies[R];
    if (!B.dynamic) continue;
    for (int RR = 0; RR < n; RR++) {
      if (!bodies[RR].dynamic) continue;
      MgrBlock &o = M[R * n + RR];
      o.xn_v0 += dt * o.vn_v0;
      o.xn_w0 += dt * o.vn_w0;
      rotation_chain(B, B.omega, dt, o.wn_w0, o.wn_v0, o.qn_w0, o.qn_v0, nullptr);  // omega already updated (:411)
    }
  }
}

// End of TimeStepDiffDFSPH::step (time advance :645, backwardPerStep :493-524, endStep :432-490) followed by
// the rest of SimulatorBase::timeStepNoGUI (:1159-1169): manager stages, BoundarySimulator::updateBoundaryForces
// (BoundarySimulator.cpp:10-36), RigidBody3dBoundarySimulator::velocityTimeStep / positionTimeStep (:278-344).
// phase: BODY_ALL (no contact solver) runs the whole sequence; with the penalty contact solver the sequence is cut in
// two around the contact kernels (dfr_contact.cuh): BODY_PRE ends after addGravity (RigidBody3dBoundarySimulator.cpp:
// 284-296, gravity first and without the isAnimated test), BODY_POST starts at the repeated updateBoundaryForces
// (RigidContactSolver.cpp:357-360).
enum { BODY_ALL = 0, BODY_PRE = 1, BODY_POST = 2 };

__global__ void k_body_update(const __grid_constant__ Params P, StepState *st, BodyDev *bodies, MgrBlock *M, int phase) {
  const double h = st->h;  // NEW h: backwardPerStep, the manager and the rigid integrator re-read the TimeManager
  const int b = threadIdx.x;
  const int n = P.n_bodies;
  if (b == 0 && phase != BODY_POST) {
    st->last_iters = st->prs_iters;
    st->total_iters += st->prs_iters;
    st->total_particle_steps += P.slab ? st->own_end - st->own_begin : st->nf;
    st->time += st->h_step;
    st->finished = (st->time >= P.target_time + P.uniform_acc_time) ? 1 : 0;
  }
  const d3 g = mk3(P.gx, P.gy, P.gz);
  if (phase == BODY_POST) {
    if (b != 0) return;
    // updateBoundaryForces() once per boundary model: the first call applies the fluid force, each later call adds a
    // zero torque, i.e. one more gyroscopic increment (addTorque), and overwrites the getForce()/getTorque() backup
    for (int rep = 0; rep < n; rep++)
      for (int R = 0; R < n; R++) {
        BodyDev &B = bodies[R];
        if (!B.dynamic) continue;
        if (!B.animated) {
          rb_add_force(B, B.force, h);
          rb_add_torque(P, B, B.torque, h);
        }
        B.force_last = B.force;
        B.torque_last = B.torque;
        B.force = mk3(0, 0, 0);
        B.torque = mk3(0, 0, 0);
      }
    if (P.use_manager) {  // after_Rigid_Rigid_coupling_step (RigidBodyGradientManager.cpp:460-468)
      mgr_force_torque_chain(P, bodies, M, true);
      mgr_velocity_chain(P, bodies, M, h);  // sees the cleared force slots
      mgr_position_rotation_chain(P, bodies, M, h);
    }
    for (int R = 0; R < n; R++)
      if (bodies[R].dynamic) rb_animate(bodies[R], h);
    return;
  }
  if (!P.use_manager) {
    if (b < n) {
      BodyDev &B = bodies[b];
      if (B.dynamic) {
        if (!B.animated) perform_chain_rule(P, B, h);
        if (phase == BODY_PRE) {
          rb_add_force(B, B.mass * g, h);
          return;
        }
        // velocityTimeStep (no contact solver)
        if (!B.animated) {
          rb_add_force(B, B.force, h);
          rb_add_torque(P, B, B.torque, h);
        }
        B.force_last = B.force;
        B.torque_last = B.torque;
        B.force = mk3(0, 0, 0);
        B.torque = mk3(0, 0, 0);
        if (!B.animated) rb_add_force(B, B.mass * g, h);
        rb_animate(B, h);
      }
    }
  } else if (b == 0) {
    for (int R = 0; R < n; R++) {
      BodyDev &B = bodies[R];
      if (B.dynamic && !B.animated) {  // update_rigid_body_gradient_manager (BoundaryModel_Akinci2012.cpp:952-967)
        MgrBlock &o = M[R * n + R];
        o.f_vn = B.net_f_v; o.f_xn = B.net_f_x; o.f_wn = B.net_f_w; o.f_qn = B.net_f_q;
        o.t_vn = B.net_t_v; o.t_xn = B.net_t_x; o.t_wn = B.net_t_w; o.t_qn = B.net_t_q;
      }
    }
    mgr_force_torque_chain(P, bodies, M, false);
    mgr_velocity_chain(P, bodies, M, h);
    if (phase == BODY_PRE) {
      for (int R = 0; R < n; R++)
        if (bodies[R].dynamic) rb_add_force(bodies[R], bodies[R].mass * g, h);
      return;
    }
    for (int R = 0; R < n; R++) {
      BodyDev &B = bodies[R];
      if (!B.dynamic) continue;
      if (!B.animated) {
        rb_add_force(B, B.force, h);
        rb_add_torque(P, B, B.torque, h);
      }
      B.force_last = B.force;
      B.torque_last = B.torque;
      B.force = mk3(0, 0, 0);
      B.torque = mk3(0, 0, 0);
    }
    for (int R = 0; R < n; R++) {
      BodyDev &B = bodies[R];
      if (B.dynamic && !B.animated) rb_add_force(B, B.mass * g, h);
    }
    mgr_position_rotation_chain(P, bodies, M, h);
    for (int R = 0; R < n; R++)
      if (bodies[R].dynamic) rb_animate(bodies[R], h);
  }
}

// SimulatorBase::updateBoundaryParticles (SimulatorBase.cpp:1827-1858)
__global__ void k_update_boundary_particles(const BodyDev *bodies, const int *bbody, const double4 *bx0, double4 *bpos, double4 *bvel,
                                            int begin, int count, int force_all) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int j = begin + t;
  const BodyDev &B = bodies[bbody[j]];
  if (!(B.dynamic || B.animated || force_all)) return;
  const m33 R = qrot(B.q);
  const double4 x0 = bx0[j];
  const d3 x = R * mk3(x0.x, x0.y, x0.z) + B.pos;
  double4 p = bpos[j];
  p.x = x.x;
  p.y = x.y;
  p.z = x.z;
  bpos[j] = p;
  d3 v = mk3(0, 0, 0);
  if (B.dynamic || B.animated) v = cross(B.omega, x - B.pos) + B.vel;
  bvel[j] = make_double4(v.x, v.y, v.z, 0.0);
}

}  // namespace dfr
