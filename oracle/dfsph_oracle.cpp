// ============================================================================
// TEST INFRASTRUCTURE ONLY.  CPU oracle for the differentiable DFSPH time step.
//
// This file is a CPU restatement (FP64, OpenMP) of the reference's algorithm for
// the one hot path this repository accelerates.  It exists so that tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can
// check and time the CUDA path against it.  Nothing in the product
// (difffr_b200/, include/) may import, link or call anything in oracle/.
//
// Parity pinning: the restatement is validated against the reference's own
// translation units compiled in this container (oracle/_ref, built by
// oracle/Makefile from the sources where they lie under /root/reference, against
// the shim headers in oracle/shims/) — see tests/golden/README.md and DESIGN.md.
// The neighbour search (CompactNSearch @ 3f11ece1, not vendored in the reference)
// is restated from its published algorithm: uniform grid, cell = support radius,
// strict d^2 < r^2, self excluded.
//
// Reference files followed (paths relative to the reference checkout):
//   SPlisHSPlasH/DiffDFSPH/TimeStepDiffDFSPH.cpp   :353-652, 654-2056
//   SPlisHSPlasH/TimeStep.cpp                      :66-81, 147-200
//   SPlisHSPlasH/SPHKernels.h                      :16-174, 437-560
//   SPlisHSPlasH/BoundaryModel.{h,cpp}             :46-58 / :38-85
//   SPlisHSPlasH/BoundaryModel_Akinci2012.cpp      :62-164, 257-284, 453-889, 952-967
//   SPlisHSPlasH/RigidBodyGradientManager.cpp      :96-524
//   SPlisHSPlasH/GradientUtils.cpp                 :4-53
//   SPlisHSPlasH/Simulation.cpp                    :382-386, 524-616, 831-902
//   SPlisHSPlasH/FluidModel.cpp                    :195-275
//   SPlisHSPlasH/Dynamic3dRigidBody.h              :113-219
//   Simulator/BoundarySimulator.cpp                :10-36
//   Simulator/RigidBody3dBoundarySimulator.cpp     :278-344
//   Simulator/SimulatorBase.cpp                    :887-934, 1142-1169, 1827-1858
//   SPlisHSPlasH/SurfaceTension/SurfaceTension_Akinci2013.cpp :25-151
//   SPlisHSPlasH/Viscosity/Viscosity_Standard.cpp  :233-334
//   SPlisHSPlasH/Emitter.cpp :89-227, EmitterSystem.cpp :54-83
//   SPlisHSPlasH/InterlinkedSPH/RigidContactSolver.cpp :23-263, 307-345, 419-555, 1341-1370
// ============================================================================
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/dfr.h"
#include "smallmat.h"

using namespace orc;

namespace {

inline Vec3 ld3(const double *p) { return vec3(p[0], p[1], p[2]); }
inline void st3(double *p, const Vec3 &v) {
  p[0] = X(v);
  p[1] = Y(v);
  p[2] = Z(v);
}

// ---------------------------------------------------------------------------
// SPH kernels (SPHKernels.h:16-174, 437-560)
// ---------------------------------------------------------------------------
struct Kernels {
  double radius, k, l, W_zero;  // CubicKernel::setRadius (SPHKernels.h:25-34)
  double coh_k, coh_c;          // CohesionKernel::setRadius (:451-458)
  double adh_k;                 // AdhesionKernel::setRadius (:520-525)
  void set_radius(double h) {
    radius = h;
    const double pi = M_PI;
    const double h3 = h * h * h;
    k = 8.0 / (pi * h3);
    l = 48.0 / (pi * h3);
    W_zero = W(0.0);
    coh_k = 32.0 / (pi * std::pow(h, 9.0));
    coh_c = std::pow(h, 6.0) / 64.0;
    adh_k = 0.007 / std::pow(h, 3.25);
  }
  // CubicKernel::W (SPHKernels.h:37-55)
  double W(double r) const {
    double res = 0.0;
    const double q = r / radius;
    if (q <= 1.0) {
      if (q <= 0.5) {
        const double q2 = q * q;
        const double q3 = q2 * q;
        res = k * (6.0 * q3 - 6.0 * q2 + 1.0);
      } else {
        res = k * (2.0 * std::pow(1.0 - q, 3.0));
      }
    }
    return res;
  }
  double W(const Vec3 &r) const { return W(norm(r)); }
  // CubicKernel::gradW (SPHKernels.h:80-102)
  Vec3 gradW(const Vec3 &r) const {
    const double rl = norm(r);
    const double q = rl / radius;
    if ((rl > 1.0e-5) && (q <= 1.0)) {
      const Vec3 gradq = r * (1.0 / (rl * radius));
      if (q <= 0.5) return (l * q * (3.0 * q - 2.0)) * gradq;
      const double factor = 1.0 - q;
      return (l * (-factor * factor)) * gradq;
    }
    return Vec3::zero();
  }
  // CubicKernel::gradGradW (SPHKernels.h:123-150)
  Mat3 gradGradW(const Vec3 &r) const {
    const double rl = norm(r);
    const double q = rl / radius;
    if ((rl > 1.0e-5) && (q <= 1.0)) {
      const Vec3 gradq = r * (1.0 / (rl * radius));
      const Mat3 gradGradq = (1.0 / (rl * radius)) * (Mat3::identity() - outer(r, r) * (1.0 / (rl * rl)));
      if (q <= 0.5)
        return (l * q * (3.0 * q - 2.0)) * gradGradq + (l * (6.0 * q - 2.0)) * outer(gradq, gradq);
      const double factor = 1.0 - q;
      return (l * (-factor * factor)) * gradGradq + (l * 2.0 * factor) * outer(gradq, gradq);
    }
    return Mat3::zero();
  }
  // CohesionKernel::W(Vector3r) (SPHKernels.h:483-499)
  double cohesionW(const Vec3 &r) const {
    double res = 0.0;
    const double r2 = sqnorm(r);
    if (r2 <= radius * radius) {
      const double r1 = std::sqrt(r2);
      const double r3 = r2 * r1;
      if (r1 > 0.5 * radius)
        res = coh_k * std::pow(radius - r1, 3.0) * r3;
      else
        res = coh_k * 2.0 * std::pow(radius - r1, 3.0) * r3 - coh_c;
    }
    return res;
  }
  // AdhesionKernel::W(Vector3r) (SPHKernels.h:544-556)
  double adhesionW(const Vec3 &r) const {
    double res = 0.0;
    const double r2 = sqnorm(r);
    if (r2 <= radius * radius) {
      const double rl = std::sqrt(r2);
      if (rl > 0.5 * radius) res = adh_k * std::pow(-4.0 * r2 / radius + 6.0 * rl - 2.0 * radius, 0.25);
    }
    return res;
  }
};

// ---------------------------------------------------------------------------
// GradientUtils.cpp:4-53
// ---------------------------------------------------------------------------
Mat3 skewMatrix(const Vec3 &v) {
  Mat3 m = Mat3::zero();
  m(0, 1) = -Z(v);
  m(0, 2) = Y(v);
  m(1, 0) = Z(v);
  m(1, 2) = -X(v);
  m(2, 0) = -Y(v);
  m(2, 1) = X(v);
  return m;
}
Mat34 assemble34(const Vec3 &c0, const Mat3 &rest) {
  Mat34 r;
  for (int i = 0; i < 3; i++) {
    r(i, 0) = c0(i, 0);
    for (int j = 0; j < 3; j++) r(i, j + 1) = rest(i, j);
  }
  return r;
}
// d(R(q) p)/dq, columns ordered (w, x, y, z)  (GradientUtils.cpp:9-19)
Mat34 get_grad_Rqp_to_q(const Quat &q, const Vec3 &p) {
  const Vec3 qv = vec3(q.x, q.y, q.z);
  const Vec3 tmp1 = 2.0 * (q.w * p - cross(p, qv));
  const Mat3 tmp2 = 2.0 * (dot(qv, p) * Mat3::identity() + outer(qv, p) - outer(p, qv) - q.w * skewMatrix(p));
  return assemble34(tmp1, tmp2);
}
// d(R(q)^T p)/dq  (GradientUtils.cpp:22-34)
Mat34 get_grad_RqTp_to_q(const Quat &q, const Vec3 &p) {
  const Vec3 qv = vec3(q.x, q.y, q.z);
  const Vec3 tmp1 = 2.0 * (q.w * p + cross(p, qv));
  const Mat3 tmp2 = 2.0 * (dot(qv, p) * Mat3::identity() + outer(qv, p) - outer(p, qv) + q.w * skewMatrix(p));
  return assemble34(tmp1, tmp2);
}
// GradientUtils.cpp:36-43
Mat4 get_grad_p_q_product_to_q(const Quat &p) {
  Mat4 m;
  const double r[4][4] = {{p.w, -p.x, -p.y, -p.z}, {p.x, p.w, -p.z, p.y}, {p.y, p.z, p.w, -p.x}, {p.z, -p.y, p.x, p.w}};
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) m(i, j) = r[i][j];
  return m;
}
// GradientUtils.cpp:45-53
Mat43 get_grad_omega_q_product_to_omega(const Quat &q) {
  Mat43 m;
  const double r[4][3] = {{-q.x, -q.y, -q.z}, {q.w, q.z, -q.y}, {-q.z, q.w, q.x}, {q.y, -q.x, q.w}};
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 3; j++) m(i, j) = r[i][j];
  return m;
}

// ---------------------------------------------------------------------------
// Neighbour search (restates CompactNSearch @ 3f11ece1: uniform grid with cell size = radius,
// query of the 27 surrounding cells, strict d^2 < r^2, a point is never its own neighbour).
// Implemented with sorted 63-bit cell keys so that it shares nothing with the CUDA cell table.
// ---------------------------------------------------------------------------
struct PointGrid {
  std::vector<uint64_t> keys;   // sorted cell keys
  std::vector<int32_t> order;   // point index per sorted slot
  std::vector<uint64_t> ukeys;  // unique keys
  std::vector<int32_t> ustart;  // start offset per unique key (+ sentinel)
  double inv_cell;
  static inline int64_t cell_coord(double x, double inv) { return (int64_t)std::floor(x * inv); }
  static inline uint64_t make_key(int64_t cx, int64_t cy, int64_t cz) {
    const int64_t off = (int64_t)1 << 20;
    return ((uint64_t)(cx + off) << 42) | ((uint64_t)(cy + off) << 21) | (uint64_t)(cz + off);
  }
  void build(const double *x, int64_t n, double cell) {
    inv_cell = 1.0 / cell;
    std::vector<std::pair<uint64_t, int32_t>> kv((size_t)n);
    for (int64_t i = 0; i < n; i++) {
      kv[i].first = make_key(cell_coord(x[3 * i], inv_cell), cell_coord(x[3 * i + 1], inv_cell), cell_coord(x[3 * i + 2], inv_cell));
      kv[i].second = (int32_t)i;
    }
    std::sort(kv.begin(), kv.end());
    keys.resize(n);
    order.resize(n);
    ukeys.clear();
    ustart.clear();
    for (int64_t i = 0; i < n; i++) {
      keys[i] = kv[i].first;
      order[i] = kv[i].second;
      if (i == 0 || keys[i] != keys[i - 1]) {
        ukeys.push_back(keys[i]);
        ustart.push_back((int32_t)i);
      }
    }
    ustart.push_back((int32_t)n);
  }
  // appends, in ascending index order, all points of this grid with |xq - x|^2 < r2 (excluding `self`)
  void query(const double *x, const double *xq, double r2, int32_t self, std::vector<int32_t> &out) const {
    const size_t begin = out.size();
    const int64_t cx = cell_coord(xq[0], inv_cell), cy = cell_coord(xq[1], inv_cell), cz = cell_coord(xq[2], inv_cell);
    for (int64_t dx = -1; dx <= 1; dx++)
      for (int64_t dy = -1; dy <= 1; dy++)
        for (int64_t dz = -1; dz <= 1; dz++) {
          const uint64_t key = make_key(cx + dx, cy + dy, cz + dz);
          auto it = std::lower_bound(ukeys.begin(), ukeys.end(), key);
          if (it == ukeys.end() || *it != key) continue;
          const size_t u = (size_t)(it - ukeys.begin());
          for (int32_t s = ustart[u]; s < ustart[u + 1]; s++) {
            const int32_t j = order[s];
            if (j == self) continue;
            // CompactNSearch distance test: l2 = dx*dx; l2 += dy*dy; l2 += dz*dz; l2 < r2  (no FMA: -ffp-contract=off)
            double tmp = xq[0] - x[3 * j];
            double l2 = tmp * tmp;
            tmp = xq[1] - x[3 * j + 1];
            l2 += tmp * tmp;
            tmp = xq[2] - x[3 * j + 2];
            l2 += tmp * tmp;
            if (l2 < r2) out.push_back(j);
          }
        }
    std::sort(out.begin() + begin, out.end());
  }
};

typedef std::vector<std::vector<int32_t>> NList;

// ---------------------------------------------------------------------------
// Scene data
// ---------------------------------------------------------------------------
struct Body {
  // BoundaryModel_Akinci2012.h:34-38
  int64_t n = 0;
  std::vector<double> x0, x, v;  // n*3
  std::vector<double> V;         // n
  // Dynamic3dRigidBody.h:20-40
  bool dynamic = false, animated = false;
  Vec3 pos0, pos, vel, omega;
  Quat q0, q;
  double density = 1000.0, mass = 0.0, invMass = 0.0;
  Mat3 inertia0, inertia, invInertia;
  // BoundaryModel.h:24-27 (per-thread slots)
  std::vector<Vec3> forcePerThread, torquePerThread, forceBackup, torqueBackup;
  // per-particle Jacobians (BoundaryModel_Akinci2012.h:42-50)
  std::vector<Mat3> g_force_v, g_force_x, g_force_omega, g_torque_omega, g_torque_x, g_torque_v;
  std::vector<Mat34> g_force_q, g_torque_q;
  // net (per step) Jacobians (BoundaryModel_Akinci2012.h:52-61)
  Mat3 net_force_vn, net_force_xn, net_force_omega_n, net_torque_omega_n, net_torque_vn, net_torque_xn;
  Mat34 net_force_qn, net_torque_qn;
  // sensitivities (BoundaryModel_Akinci2012.h:63-77)
  Mat3 grad_x_to_v0, grad_x_to_omega0, grad_v_to_v0, grad_v_to_omega0, grad_omega_to_v0, grad_omega_to_omega0;
  Mat43 grad_q_to_v0, grad_q_to_omega0, partial_grad_qn_to_omega_n;
  // SimulationDataDiffDFSPH per-body
  Vec3 init_v, init_omega;
  // RigidContactSolver per-particle state (RigidContactSolver.h:134-147), id order
  std::vector<double> c_vol0, c_density0, c_density, c_vol, c_vel;
  // storage order of the reference (slot -> particle id): CompactNSearch z_sort permutes dynamic point sets
  // (BoundaryModel_Akinci2012.cpp:388-416); only the penalty solver's per-contact addTorque sequence depends on it
  std::vector<int32_t> order;
  // neighbour lists
  NList nb_fluid;                 // body particle -> fluid
  std::vector<NList> nb_body;     // body particle -> other body particles (boundary volume / contact)
  PointGrid grid;

  void updateInertia() {  // Dynamic3dRigidBody.h:215-219
    const Mat3 R = qrot(q);
    inertia = R * inertia0 * transpose(R);
    invInertia = inverse(inertia);
  }
  void resetGradient() {  // BoundaryModel_Akinci2012.cpp:118-162
    for (int64_t j = 0; j < n; j++) {
      g_force_v[j] = g_force_x[j] = g_force_omega[j] = Mat3::zero();
      g_torque_omega[j] = g_torque_v[j] = g_torque_x[j] = Mat3::zero();
      g_force_q[j] = g_torque_q[j] = Mat34::zero();
    }
    net_force_vn = net_force_xn = net_force_omega_n = Mat3::zero();
    net_torque_omega_n = net_torque_vn = net_torque_xn = Mat3::zero();
    net_force_qn = net_torque_qn = Mat34::zero();
    grad_v_to_v0 = Mat3::identity();
    grad_v_to_omega0 = Mat3::zero();
    grad_omega_to_omega0 = Mat3::identity();
    grad_omega_to_v0 = Mat3::zero();
    grad_q_to_omega0 = grad_q_to_v0 = partial_grad_qn_to_omega_n = Mat43::zero();
    grad_x_to_v0 = grad_x_to_omega0 = Mat3::zero();
  }
  // BoundaryModel::addForce (BoundaryModel.h:46-58)
  inline void addForce(const Vec3 &p, const Vec3 &f, int tid) {
    if (dynamic) {
      forcePerThread[tid] += f;
      torquePerThread[tid] += cross(p - pos, f);
    }
  }
  void getForceAndTorque(Vec3 &force, Vec3 &torque) {  // BoundaryModel.cpp:38-54
    force = Vec3::zero();
    torque = Vec3::zero();
    for (size_t j = 0; j < forcePerThread.size(); j++) {
      force += forcePerThread[j];
      torque += torquePerThread[j];
      forceBackup[j] = forcePerThread[j];
      torqueBackup[j] = torquePerThread[j];
    }
  }
  void clearForceAndTorque() {  // BoundaryModel.cpp:75-85
    for (size_t j = 0; j < forcePerThread.size(); j++) {
      forceBackup[j] = forcePerThread[j];
      torqueBackup[j] = torquePerThread[j];
      forcePerThread[j] = Vec3::zero();
      torquePerThread[j] = Vec3::zero();
    }
  }
  Vec3 getForce() const {
    Vec3 f = Vec3::zero();
    for (size_t j = 0; j < forceBackup.size(); j++) f += forceBackup[j];
    return f;
  }
  Vec3 getTorque() const {
    Vec3 t = Vec3::zero();
    for (size_t j = 0; j < torqueBackup.size(); j++) t += torqueBackup[j];
    return t;
  }
};

struct Emitter {
  int width, height;
  Vec3 x;
  Mat3 rot;
  double velocity, emitStart, emitEnd;
  double nextEmitTime;
  int64_t emitCounter;
};

struct Manager {  // RigidBodyGradientManager.h:62-91
  int n = 0;
  std::vector<Mat3> xn_v0, xn_w0, vn_v0, vn_w0, wn_v0, wn_w0;
  std::vector<Mat43> qn_v0, qn_w0;
  std::vector<Mat3> f_vn, f_xn, f_wn, t_vn, t_xn, t_wn;
  std::vector<Mat34> f_qn, t_qn;
  std::vector<Mat3> f_v0, f_w0, t_v0, t_w0;
  int at(int R, int RR) const { return R * n + RR; }
  void init(int nb) {
    n = nb;
    const size_t s = (size_t)nb * nb;
    xn_v0.resize(s); xn_w0.resize(s); vn_v0.resize(s); vn_w0.resize(s); wn_v0.resize(s); wn_w0.resize(s);
    qn_v0.resize(s); qn_w0.resize(s);
    f_vn.resize(s); f_xn.resize(s); f_wn.resize(s); t_vn.resize(s); t_xn.resize(s); t_wn.resize(s);
    f_qn.resize(s); t_qn.resize(s);
    f_v0.resize(s); f_w0.resize(s); t_v0.resize(s); t_w0.resize(s);
    reset();
  }
  void reset() {  // RigidBodyGradientManager.cpp:477-524
    for (int R = 0; R < n; R++)
      for (int RR = 0; RR < n; RR++) {
        const int k = at(R, RR);
        xn_v0[k] = xn_w0[k] = vn_w0[k] = wn_v0[k] = Mat3::zero();
        qn_v0[k] = qn_w0[k] = Mat43::zero();
        vn_v0[k] = (R == RR) ? Mat3::identity() : Mat3::zero();
        wn_w0[k] = (R == RR) ? Mat3::identity() : Mat3::zero();
        f_vn[k] = f_xn[k] = f_wn[k] = t_vn[k] = t_xn[k] = t_wn[k] = Mat3::zero();
        f_qn[k] = t_qn[k] = Mat34::zero();
        f_v0[k] = f_w0[k] = t_v0[k] = t_w0[k] = Mat3::zero();
      }
  }
};

}  // namespace

struct dfr_context {
  dfr_config cfg;
  Kernels K;
  double supportRadius = 0.0;
  bool finalized = false;
  std::string err;
  int nthreads = 1;

  // fluid (FluidModel.cpp:277-291, SimulationDataDiffDFSPH.h)
  int64_t nf = 0, nfActive0 = 0, nfCapacity = 0;
  std::vector<double> x, v, a, x_init, v_init, kappa_init, kappaV_init;
  std::vector<double> density, pressure, factor, kappa, kappaV, densityAdv, sum_grad_p_k, normals;
  std::vector<int32_t> state;  // ParticleState (FluidModel.h:73)
  double Vf = 0.0, mass = 0.0;

  std::vector<Body> bodies;
  std::vector<Emitter> emitters;
  Manager mgr;

  NList nf_f;                 // fluid -> fluid
  std::vector<NList> nf_b;    // fluid -> body b
  PointGrid fluidGrid;

  // TimeManager + TimeStep counters
  double time = 0.0, h = 0.001;
  int iterations = 0, iterationsV = 0, step_count = 0;
  bool finished = false;
  int64_t totalIter = 0, totalIterV = 0, totalParticleSteps = 0, totalNeighbors = 0;
  double cpu_ms = 0.0;

  // contact solver state (RigidContactSolver), see oracle_contact.inc
  Kernels Kc;                                       // kernel with the contact support radius (W_with_h)
  std::vector<std::pair<int, int32_t>> in_contact;  // m_particle_indices_in_contact
  int64_t sortCounter = 0;                          // TimeStepDiffDFSPH::m_counter (z-sort every 500 steps)
  int64_t numContacts = 0;
};

namespace {

const double m_eps = 1.0e-5;  // TimeStepDiffDFSPH.h:26

// ---------------------------------------------------------------------------
// neighbourhood (Simulation.cpp:882-900: fluid -> all sets, every boundary set -> fluid)
// ---------------------------------------------------------------------------
void findNeighbors(dfr_context *c) {
  const double r2 = c->supportRadius * c->supportRadius;
  const int64_t nf = c->nf;
  c->fluidGrid.build(c->x.data(), nf, c->supportRadius);
  for (auto &b : c->bodies)
    if (b.dynamic || b.animated || b.grid.order.size() != (size_t)b.n) b.grid.build(b.x.data(), b.n, c->supportRadius);
  c->nf_f.resize(nf);
  c->nf_b.resize(c->bodies.size());
  for (auto &l : c->nf_b) l.resize(nf);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nf; i++) {
    c->nf_f[i].clear();
    c->fluidGrid.query(c->x.data(), &c->x[3 * i], r2, (int32_t)i, c->nf_f[i]);
    for (size_t b = 0; b < c->bodies.size(); b++) {
      c->nf_b[b][i].clear();
      c->bodies[b].grid.query(c->bodies[b].x.data(), &c->x[3 * i], r2, -1, c->nf_b[b][i]);
    }
  }
  for (auto &b : c->bodies) {
    if (!b.dynamic) {  // only consumed for dynamic bodies (TimeStepDiffDFSPH.cpp:1239)
      b.nb_fluid.clear();
      continue;
    }
    b.nb_fluid.resize(b.n);
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < b.n; j++) {
      b.nb_fluid[j].clear();
      c->fluidGrid.query(c->x.data(), &b.x[3 * j], r2, -1, b.nb_fluid[j]);
    }
  }
  int64_t tot = 0;
  for (int64_t i = 0; i < nf; i++) {
    tot += (int64_t)c->nf_f[i].size();
    for (size_t b = 0; b < c->bodies.size(); b++) tot += (int64_t)c->nf_b[b][i].size();
  }
  c->totalNeighbors += tot;
}

// Simulation::updateBoundaryVolume (Simulation.cpp:831-902) + computeBoundaryVolume
// (BoundaryModel_Akinci2012.cpp:257-284).  set_active(i, true, true) of CompactNSearch switches on
// row and column i of the activation table, so an activated body finds the particles of every
// boundary point set (activated or not) within the support radius.
void computeVolumeOf(dfr_context *c, size_t bi) {
  Body &b = c->bodies[bi];
  const double r2 = c->supportRadius * c->supportRadius;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < b.n; i++) {
    double delta = c->K.W_zero;
    std::vector<int32_t> nb;
    for (size_t pid = 0; pid < c->bodies.size(); pid++) {
      Body &o = c->bodies[pid];
      nb.clear();
      o.grid.query(o.x.data(), &b.x[3 * i], r2, (pid == bi) ? (int32_t)i : -1, nb);
      for (int32_t j : nb) delta += c->K.W(ld3(&b.x[3 * i]) - ld3(&o.x[3 * j]));
    }
    b.V[i] = 1.0 / delta;
  }
}
void updateBoundaryVolume(dfr_context *c) {
  for (auto &b : c->bodies) b.grid.build(b.x.data(), b.n, c->supportRadius);
  for (size_t b = 0; b < c->bodies.size(); b++)
    if (!c->bodies[b].dynamic && !c->bodies[b].animated) computeVolumeOf(c, b);
  for (size_t b = 0; b < c->bodies.size(); b++)
    if (c->bodies[b].dynamic || c->bodies[b].animated) computeVolumeOf(c, b);
}

// SimulatorBase::updateBoundaryParticles (SimulatorBase.cpp:1827-1858)
void updateBoundaryParticles(dfr_context *c, bool forceUpdate) {
  for (auto &b : c->bodies) {
    if (b.dynamic || b.animated || forceUpdate) {
      const Mat3 R = qrot(b.q);
#pragma omp parallel for schedule(static)
      for (int64_t j = 0; j < b.n; j++) {
        const Vec3 xj = R * ld3(&b.x0[3 * j]) + b.pos;
        st3(&b.x[3 * j], xj);
        if (b.dynamic || b.animated)
          st3(&b.v[3 * j], cross(b.omega, xj - b.pos) + b.vel);
        else
          st3(&b.v[3 * j], Vec3::zero());
      }
    }
  }
}

#ifdef _OPENMP
inline int tid() { return omp_get_thread_num(); }
#else
inline int tid() { return 0; }
#endif

// ---------------------------------------------------------------------------
// forward passes
// ---------------------------------------------------------------------------
// TimeStep::computeDensities (TimeStep.cpp:147-200)
void computeDensities(dfr_context *c) {
  const double density0 = c->cfg.density0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < c->nf; i++) {
    double density = c->Vf * c->K.W_zero;
    const Vec3 xi = ld3(&c->x[3 * i]);
    for (int32_t j : c->nf_f[i]) density += c->Vf * c->K.W(xi - ld3(&c->x[3 * j]));
    for (size_t b = 0; b < c->bodies.size(); b++) {
      const Body &bm = c->bodies[b];
      for (int32_t j : c->nf_b[b][i]) density += bm.V[j] * c->K.W(xi - ld3(&bm.x[3 * j]));
    }
    c->density[i] = density * density0;
  }
}

// computeDFSPHFactor (TimeStepDiffDFSPH.cpp:883-962)
void computeDFSPHFactor(dfr_context *c) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < c->nf; i++) {
    const Vec3 xi = ld3(&c->x[3 * i]);
    double sum_grad_p_k = 0.0;
    Vec3 grad_p_i = Vec3::zero();
    for (int32_t j : c->nf_f[i]) {
      const Vec3 grad_p_j = -c->Vf * c->K.gradW(xi - ld3(&c->x[3 * j]));
      sum_grad_p_k += sqnorm(grad_p_j);
      grad_p_i = grad_p_i - grad_p_j;
    }
    for (size_t b = 0; b < c->bodies.size(); b++) {
      const Body &bm = c->bodies[b];
      for (int32_t j : c->nf_b[b][i]) {
        const Vec3 grad_p_j = -bm.V[j] * c->K.gradW(xi - ld3(&bm.x[3 * j]));
        grad_p_i = grad_p_i - grad_p_j;
      }
    }
    st3(&c->sum_grad_p_k[3 * i], -grad_p_i);
    sum_grad_p_k += sqnorm(grad_p_i);
    c->factor[i] = (sum_grad_p_k > m_eps) ? -1.0 / sum_grad_p_k : 0.0;
  }
}

// shared by computeDensityAdv / computeDensityChange: sum_j V_j (v_i - v_j) . gradW(x_i - x_j)
inline double velocityDivergenceSum(dfr_context *c, int64_t i) {
  const Vec3 xi = ld3(&c->x[3 * i]);
  const Vec3 vi = ld3(&c->v[3 * i]);
  double delta = 0.0;
  for (int32_t j : c->nf_f[i]) delta += c->Vf * dot(vi - ld3(&c->v[3 * j]), c->K.gradW(xi - ld3(&c->x[3 * j])));
  for (size_t b = 0; b < c->bodies.size(); b++) {
    const Body &bm = c->bodies[b];
    for (int32_t j : c->nf_b[b][i]) delta += bm.V[j] * dot(vi - ld3(&bm.v[3 * j]), c->K.gradW(xi - ld3(&bm.x[3 * j])));
  }
  return delta;
}
// computeDensityAdv (TimeStepDiffDFSPH.cpp:1926-1975)
inline void computeDensityAdv(dfr_context *c, int64_t i, double h, double density0) {
  const double delta = velocityDivergenceSum(c, i);
  double densityAdv = c->density[i] / density0 + h * delta;
  c->densityAdv[i] = std::max(densityAdv, 1.0);
}
// computeDensityChange (TimeStepDiffDFSPH.cpp:1977-2041)
inline void computeDensityChange(dfr_context *c, int64_t i) {
  double densityAdv = velocityDivergenceSum(c, i);
  densityAdv = std::max(densityAdv, 0.0);
  size_t numNeighbors = c->nf_f[i].size();
  for (size_t b = 0; b < c->bodies.size(); b++) numNeighbors += c->nf_b[b][i].size();
  if (numNeighbors < 20) densityAdv = 0.0;
  c->densityAdv[i] = densityAdv;
}

// The velocity push shared by warm starts and Jacobi iterations
// (TimeStepDiffDFSPH.cpp:1016-1056, 1117-1196, 1745-1785, 1840-1901).
// kj(j) gives the neighbour stiffness.
template <class KJ>
inline void pushVelocity(dfr_context *c, int64_t i, double ki, double h, double invH, KJ kj) {
  Vec3 vel = ld3(&c->v[3 * i]);
  c->pressure[i] = ki * c->density[i];
  const Vec3 xi = ld3(&c->x[3 * i]);
  for (int32_t j : c->nf_f[i]) {
    const double kSum = ki + 1.0 * kj(j);  // density0_j / density0 = 1 (single fluid phase)
    if (std::fabs(kSum) > m_eps) {
      const Vec3 grad_p_j = -c->Vf * c->K.gradW(xi - ld3(&c->x[3 * j]));
      vel = vel - (h * kSum) * grad_p_j;
    }
  }
  if (std::fabs(ki) > m_eps) {
    for (size_t b = 0; b < c->bodies.size(); b++) {
      Body &bm = c->bodies[b];
      for (int32_t j : c->nf_b[b][i]) {
        const Vec3 xj = ld3(&bm.x[3 * j]);
        const Vec3 grad_p_j = -bm.V[j] * c->K.gradW(xi - xj);
        const Vec3 velChange = (-h * 1.0 * ki) * grad_p_j;
        vel = vel + velChange;
        const Vec3 force = (-c->mass * velChange) * invH;
        bm.addForce(xj, force, tid());
      }
    }
  }
  st3(&c->v[3 * i], vel);
}

// ---------------------------------------------------------------------------
// Force / torque Jacobians (TimeStepDiffDFSPH.cpp:1226-1694)
// ---------------------------------------------------------------------------
void computeGradient(dfr_context *c, Body &bm, int64_t i, int64_t rj, double b_i, double ki, const Vec3 &force, int mode) {
  const double dt = c->h;
  const double invH = 1.0 / dt;
  const double invH2 = 1.0 / dt / dt;
  const Vec3 xi = ld3(&c->x[3 * i]);
  const Vec3 xj = ld3(&bm.x[3 * rj]);
  const Vec3 vi = ld3(&c->v[3 * i]);
  const Vec3 vj = ld3(&bm.v[3 * rj]);
  const double vol = bm.V[rj];
  const Vec3 grad_p_j = -vol * c->K.gradW(xi - xj);
  const Mat3 grad2_p_j_to_xj = vol * c->K.gradGradW(xi - xj);
  const double factor = c->factor[i];
  const double densityAdv = c->densityAdv[i];
  const double density0 = c->cfg.density0;
  const Vec3 sum_grad = ld3(&c->sum_grad_p_k[3 * i]);

  // :1518-1577
  Vec3 grad_b_i_to_xj = Vec3::zero();
  if (mode == 0) {
    if (densityAdv > 1.0) {
      const Vec3 grad_density_to_x = -vol * c->K.gradW(xi - xj);
      grad_b_i_to_xj = grad_density_to_x * (1.0 / density0) - (dt * vol) * (c->K.gradGradW(xi - xj) * (vi - vj));
    }
  } else {
    if (densityAdv > 0.0) grad_b_i_to_xj = (-vol) * (c->K.gradGradW(xi - xj) * (vi - vj));
  }
  // :1485-1516
  Vec3 grad_factor_to_xj = Vec3::zero();
  if (!(factor >= 0.0)) grad_factor_to_xj = (2.0 * factor * factor) * (grad2_p_j_to_xj * sum_grad);

  const double coeff = (mode != 0) ? invH : invH2;
  const Vec3 grad_ki_to_xj = (grad_b_i_to_xj * factor + b_i * grad_factor_to_xj) * coeff;
  const Mat3 grad_velChange_to_xj = -1.0 * (outer(grad_p_j, grad_ki_to_xj) + ki * grad2_p_j_to_xj);

  // :1579-1628
  Vec3 grad_b_i_to_vj = Vec3::zero();
  if (mode == 0) {
    if (densityAdv > 1.0) grad_b_i_to_vj = (-dt * vol) * c->K.gradW(xi - xj);
  } else {
    if (densityAdv > 0.0) grad_b_i_to_vj = (-vol) * c->K.gradW(xi - xj);
  }
  const Vec3 grad_ki_to_vj = grad_b_i_to_vj * factor * coeff;
  const Mat3 grad_velChange_to_vj = -1.0 * outer(grad_p_j, grad_ki_to_vj);

  // :1632-1694 (re-walks all neighbours of i)
  Vec3 grad_b_i_to_vi = Vec3::zero();
  if ((mode == 0 && densityAdv > 1.0) || (mode != 0 && densityAdv > 0.0)) {
    Vec3 s = Vec3::zero();
    for (int32_t n : c->nf_f[i]) s += c->Vf * c->K.gradW(xi - ld3(&c->x[3 * n]));
    for (size_t b = 0; b < c->bodies.size(); b++) {
      const Body &o = c->bodies[b];
      for (int32_t n : c->nf_b[b][i]) s += o.V[n] * c->K.gradW(xi - ld3(&o.x[3 * n]));
    }
    grad_b_i_to_vi = (mode == 0) ? s * dt : s;
  }
  const Vec3 grad_ki_to_vi = grad_b_i_to_vi * factor * coeff;
  const Mat3 grad_velChange_to_vi = -outer(grad_p_j, grad_ki_to_vi);

  const Mat3 grad_force_to_xj = -c->mass * grad_velChange_to_xj;
  const Mat3 grad_force_to_vj = (inverse(Mat3::identity() - dt * grad_velChange_to_vi) * (-c->mass)) * grad_velChange_to_vj;

  bm.g_force_x[rj] += grad_force_to_xj;
  bm.g_force_v[rj] += grad_force_to_vj;

  if (c->cfg.optimize_rotation) {
    const Vec3 r = xj - bm.pos;
    const Vec3 r0 = ld3(&bm.x0[3 * rj]);
    const Quat q = bm.q;
    const Mat34 grad_rj_to_q = get_grad_Rqp_to_q(q, r0);
    const Vec3 omega = bm.omega;
    const Mat34 grad_force_to_quaternion = grad_force_to_xj * grad_rj_to_q + grad_force_to_vj * skewMatrix(omega) * grad_rj_to_q;
    const Mat3 grad_force_to_omega = grad_force_to_vj * transpose(skewMatrix(r));
    const Mat34 grad_torque_to_quaternion = skewMatrix(r) * grad_force_to_quaternion + transpose(skewMatrix(force)) * grad_rj_to_q;
    const Mat3 grad_torque_to_omega = skewMatrix(r) * grad_force_to_omega;
    bm.g_force_omega[rj] += grad_force_to_omega;
    bm.g_force_q[rj] += grad_force_to_quaternion;
    bm.g_torque_omega[rj] += grad_torque_to_omega;
    bm.g_torque_q[rj] += grad_torque_to_quaternion;
    bm.g_torque_v[rj] += skewMatrix(r) * grad_force_to_vj;
    bm.g_torque_x[rj] += skewMatrix(r) * grad_force_to_xj;
  }
  (void)invH;
}

// computeRigidBodyGradient (TimeStepDiffDFSPH.cpp:1226-1279)
void computeRigidBodyGradient(dfr_context *c, int mode) {
  const double dt = c->h;
  const double invH = 1.0 / dt;
  const double invH2 = 1.0 / dt / dt;
  for (auto &bm : c->bodies) {
    if (!(bm.dynamic && !bm.animated)) continue;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < bm.n; j++) {
      const Vec3 xj = ld3(&bm.x[3 * j]);
      for (int32_t i : bm.nb_fluid[j]) {
        const Vec3 xi = ld3(&c->x[3 * i]);
        double b_i = (mode == 0) ? c->densityAdv[i] - 1.0 : c->densityAdv[i];
        const unsigned int coeff = (unsigned int)((mode == 0) ? invH2 : invH);  // :1266 — truncating conversion, as coded
        const double ki = b_i * c->factor[i] * coeff;
        const Vec3 grad_p_j = -bm.V[j] * c->K.gradW(xi - xj);
        const Vec3 velChange = (-dt * 1.0 * ki) * grad_p_j;
        const Vec3 force = (-c->mass * velChange) * invH;
        computeGradient(c, bm, i, j, b_i, ki, force, mode);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// solvers
// ---------------------------------------------------------------------------
// warmstartDivergenceSolve (TimeStepDiffDFSPH.cpp:1698-1788)
void warmstartDivergenceSolve(dfr_context *c) {
  const double h = c->h, invH = 1.0 / h;
  const int64_t n = c->nf;
  if (n == 0) return;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    computeDensityChange(c, i);
    if (c->densityAdv[i] > 0.0)
      c->kappaV[i] = 0.5 * std::max(c->kappaV[i], -0.5) * invH;
    else
      c->kappaV[i] = 0.0;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    if (c->state[i] != 0) {
      c->kappaV[i] = 0.0;
      continue;
    }
    pushVelocity(c, i, c->kappaV[i], h, invH, [c](int32_t j) { return c->kappaV[j]; });
  }
}

// divergenceSolveIteration (TimeStepDiffDFSPH.cpp:1790-1922)
void divergenceSolveIteration(dfr_context *c, double &avg_density_err) {
  const double density0 = c->cfg.density0;
  const int64_t n = c->nf;
  if (n == 0) return;
  const double h = c->h, invH = 1.0 / h;
  computeRigidBodyGradient(c, 1);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    if (c->state[i] != 0) continue;
    const double b_i = c->densityAdv[i];
    const double ki = b_i * c->factor[i] * invH;
    if (c->cfg.use_divergence_warmstart) c->kappaV[i] += ki;
    pushVelocity(c, i, ki, h, invH, [c, invH](int32_t j) { return c->densityAdv[j] * c->factor[j] * invH; });
  }
  double density_error = 0.0;
  std::vector<double> part(c->nthreads, 0.0);
#pragma omp parallel
  {
    double local = 0.0;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; i++) {
      computeDensityChange(c, i);
      local += density0 * c->densityAdv[i];
    }
    part[tid()] = local;
  }
  for (double p : part) density_error += p;
  avg_density_err = density_error / (double)n;
}

// divergenceSolve (TimeStepDiffDFSPH.cpp:770-881)
void divergenceSolve(dfr_context *c) {
  const double h = c->h;
  const int maxIter = c->cfg.max_iterations_v;
  const double maxError = c->cfg.max_error_v;
  if (c->cfg.use_divergence_warmstart) warmstartDivergenceSolve(c);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < c->nf; i++) computeDensityChange(c, i);
  c->iterationsV = 0;
  double avg_density_err = 0.0;
  bool chk = false;
  while ((!chk || (c->iterationsV < 1)) && (c->iterationsV < maxIter)) {
    chk = true;
    const double density0 = c->cfg.density0;
    avg_density_err = 0.0;
    divergenceSolveIteration(c, avg_density_err);
    const double eta = (1.0 / h) * maxError * 0.01 * density0;
    chk = chk && (avg_density_err <= eta);
    c->iterationsV++;
  }
  c->totalIterV += c->iterationsV;
  if (c->cfg.use_divergence_warmstart)
    for (int64_t i = 0; i < c->nf; i++) c->kappaV[i] *= h;
}

// warmstartPressureSolve (TimeStepDiffDFSPH.cpp:964-1059)
void warmstartPressureSolve(dfr_context *c) {
  const double h = c->h, h2 = h * h, invH = 1.0 / h, invH2 = 1.0 / h2;
  const double density0 = c->cfg.density0;
  const int64_t n = c->nf;
  if (n == 0) return;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    computeDensityAdv(c, i, h, density0);
    if (c->densityAdv[i] > 1.0)
      c->kappa[i] = 0.5 * std::max(c->kappa[i], -0.00025) * invH2;
    else
      c->kappa[i] = 0.0;
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    if (c->state[i] != 0) {
      c->kappa[i] = 0.0;
      continue;
    }
    pushVelocity(c, i, c->kappa[i], h, invH, [c](int32_t j) { return c->kappa[j]; });
  }
}

// pressureSolveIteration (TimeStepDiffDFSPH.cpp:1061-1220)
void pressureSolveIteration(dfr_context *c, double &avg_density_err) {
  const double density0 = c->cfg.density0;
  const int64_t n = c->nf;
  if (n == 0) return;
  const double h = c->h, h2 = h * h, invH = 1.0 / h, invH2 = 1.0 / h2;
  computeRigidBodyGradient(c, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    if (c->state[i] != 0) continue;
    const double b_i = c->densityAdv[i] - 1.0;
    const double ki = b_i * c->factor[i] * invH2;
    if (c->cfg.use_pressure_warmstart) c->kappa[i] += ki;
    pushVelocity(c, i, ki, h, invH, [c, invH2](int32_t j) { return (c->densityAdv[j] - 1.0) * c->factor[j] * invH2; });
  }
  double density_error = 0.0;
  std::vector<double> part(c->nthreads, 0.0);
#pragma omp parallel
  {
    double local = 0.0;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; i++) {
      computeDensityAdv(c, i, h, density0);
      local += density0 * c->densityAdv[i] - density0;
    }
    part[tid()] = local;
  }
  for (double p : part) density_error += p;
  avg_density_err = density_error / (double)n;
}

// pressureSolve (TimeStepDiffDFSPH.cpp:654-768)
void pressureSolve(dfr_context *c) {
  const double h = c->h, h2 = h * h;
  const double density0 = c->cfg.density0;
  if (c->nf == 0) return;
  if (c->cfg.use_pressure_warmstart) warmstartPressureSolve(c);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < c->nf; i++) computeDensityAdv(c, i, h, density0);
  c->iterations = 0;
  double avg_density_err = 0.0;
  bool chk = false;
  while ((!chk || (c->iterations < c->cfg.min_iterations)) && (c->iterations < c->cfg.max_iterations)) {
    chk = true;
    avg_density_err = 0.0;
    pressureSolveIteration(c, avg_density_err);
    const double eta = c->cfg.max_error * 0.01 * density0;
    chk = chk && (avg_density_err <= eta);
    c->iterations++;
  }
  c->totalIter += c->iterations;
  if (c->cfg.use_pressure_warmstart)
    for (int64_t i = 0; i < c->nf; i++) c->kappa[i] *= h2;
}

// ---------------------------------------------------------------------------
// non-pressure forces
// ---------------------------------------------------------------------------
// SurfaceTension_Akinci2013 (SurfaceTension_Akinci2013.cpp:25-151)
void surfaceTensionAkinci2013(dfr_context *c) {
  const double density0 = c->cfg.density0;
  const double supportRadius = c->supportRadius;
  const double k = c->cfg.surface_tension, kb = c->cfg.surface_tension_boundary;
  const int64_t n = c->nf;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    const Vec3 xi = ld3(&c->x[3 * i]);
    Vec3 ni = Vec3::zero();
    for (int32_t j : c->nf_f[i]) ni += (c->mass / c->density[j]) * c->K.gradW(xi - ld3(&c->x[3 * j]));
    st3(&c->normals[3 * i], supportRadius * ni);
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    const Vec3 xi = ld3(&c->x[3 * i]);
    const Vec3 ni = ld3(&c->normals[3 * i]);
    const double rhoi = c->density[i];
    Vec3 ai = ld3(&c->a[3 * i]);
    for (int32_t j : c->nf_f[i]) {
      const Vec3 xj = ld3(&c->x[3 * j]);
      const double rhoj = c->density[j];
      const double K_ij = 2.0 * density0 / (rhoi + rhoj);
      Vec3 accel = Vec3::zero();
      Vec3 xixj = xi - xj;
      const double length2 = sqnorm(xixj);
      if (length2 > 1.0e-9) {
        xixj = (1.0 / std::sqrt(length2)) * xixj;
        accel = accel - (k * c->mass) * xixj * c->K.cohesionW(xi - xj);
      }
      const Vec3 nj = ld3(&c->normals[3 * j]);
      accel = accel - k * (ni - nj);
      ai += K_ij * accel;
    }
    for (size_t b = 0; b < c->bodies.size(); b++) {
      const Body &bm = c->bodies[b];
      for (int32_t j : c->nf_b[b][i]) {
        const Vec3 xj = ld3(&bm.x[3 * j]);
        Vec3 xixj = xi - xj;
        const double length2 = sqnorm(xixj);
        if (length2 > 1.0e-9) {
          xixj = (1.0 / std::sqrt(length2)) * xixj;
          ai = ai - (kb * density0 * bm.V[j]) * xixj * c->K.adhesionW(xi - xj);
        }
      }
    }
    st3(&c->a[3 * i], ai);
  }
}

// Viscosity_Standard::step (Viscosity_Standard.cpp:233-334), fluid part; boundary part only when mu_b != 0
void viscosityStandard(dfr_context *c) {
  const double h = c->supportRadius, h2 = h * h;
  const double density0 = c->cfg.density0;
  const double d = 10.0;
  const double mu = c->cfg.viscosity, mub = c->cfg.viscosity_boundary;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < c->nf; i++) {
    const Vec3 xi = ld3(&c->x[3 * i]);
    const Vec3 vi = ld3(&c->v[3 * i]);
    Vec3 ai = ld3(&c->a[3 * i]);
    const double density_i = c->density[i];
    for (int32_t j : c->nf_f[i]) {
      const Vec3 xj = ld3(&c->x[3 * j]);
      const Vec3 vj = ld3(&c->v[3 * j]);
      const double density_j = c->density[j];
      const Vec3 xixj = xi - xj;
      ai += (d * mu * (c->mass / density_j) * dot(vi - vj, xixj) / (sqnorm(xixj) + 0.01 * h2)) * c->K.gradW(xi - xj);
    }
    if (mub != 0.0) {
      for (size_t b = 0; b < c->bodies.size(); b++) {
        Body &bm = c->bodies[b];
        for (int32_t j : c->nf_b[b][i]) {
          const Vec3 xj = ld3(&bm.x[3 * j]);
          const Vec3 vj = ld3(&bm.v[3 * j]);
          const Vec3 xixj = xi - xj;
          const Vec3 acc = (d * mub * (density0 * bm.V[j] / density_i) * dot(vi - vj, xixj) / (sqnorm(xixj) + 0.01 * h2)) * c->K.gradW(xi - xj);
          ai += acc;
          bm.addForce(xj, -c->mass * acc, tid());
          // ===== backward (Viscosity_Standard.cpp:275-318; BACKWARD is defined at :12) =====
          if (bm.dynamic) {
            const double factor = d * mub * density0 * bm.V[j];
            const double n = sqnorm(xixj) + 0.01 * h2;
            const double tmp1 = dot(vi - vj, xixj) / n;
            const Vec3 gW = c->K.gradW(xi - xj);
            const Vec3 tmp2 = tmp1 * gW;
            const Mat3 grad1 = (-1.0 / density_i) * outer(tmp2, (-1.0) * gW);                  // grad density_i
            const Mat3 grad2 = (1.0 / n) * outer(gW, (-1.0) * (vi - vj));                      // grad xixj
            const Mat3 grad3 = (-dot(vi - vj, xixj) / (n * n) * 2.0) * outer(gW, (-1.0) * xixj);  // grad xixj.squaredNorm
            const Mat3 grad4 = tmp1 * ((-1.0) * c->K.gradGradW(xi - xj));
            const Mat3 grad_a_to_xj = (factor / density_i) * (grad1 + grad2 + grad3 + grad4);
            const Mat3 grad_a_to_vj = (-factor / density_i / n) * outer(gW, xixj);
            const double dt = c->h;  // TimeManager::getTimeStepSize() at this point of the step: the old step size
            const Mat3 grad_force_to_vj = (-c->mass) * (grad_a_to_vj + dt * grad_a_to_xj);
            // (the reference adds to the particle's array from inside its parallel loop over i; serialised here)
#pragma omp critical(visc_b_grad)
            bm.g_force_v[j] += grad_force_to_vj;
          }
        }
      }
    }
    st3(&c->a[3 * i], ai);
  }
}

// Simulation::updateTimeStepSizeCFL (Simulation.cpp:542-616)
void updateTimeStepSizeCFL(dfr_context *c) {
  const double radius = c->cfg.particle_radius;
  double h = c->h;
  double maxVel = 0.1;
  const double diameter = 2.0 * radius;
  for (int64_t i = 0; i < c->nf; i++) {
    const Vec3 vel = ld3(&c->v[3 * i]);
    const Vec3 accel = ld3(&c->a[3 * i]);
    const double velMag = sqnorm(vel + accel * h);
    if (velMag > maxVel) maxVel = velMag;
  }
  for (auto &bm : c->bodies) {
    if (bm.dynamic || bm.animated) {
      for (int64_t j = 0; j < bm.n; j++) {
        const double velMag = sqnorm(ld3(&bm.v[3 * j]));
        if (velMag > maxVel) maxVel = velMag;
      }
    }
  }
  h = c->cfg.cfl_factor * 0.4 * (diameter / (std::sqrt(maxVel)));
  h = std::min(h, c->cfg.cfl_max_time_step);
  h = std::max(h, c->cfg.cfl_min_time_step);
  c->h = h;
}
void updateTimeStepSize(dfr_context *c) {  // Simulation.cpp:524-540
  if (c->cfg.cfl_method == 1)
    updateTimeStepSizeCFL(c);
  else if (c->cfg.cfl_method == 2) {
    double h = c->h;
    updateTimeStepSizeCFL(c);
    if (c->iterations > 10)
      h *= 0.9;
    else if (c->iterations < 5)
      h *= 1.1;
    h = std::min(h, c->h);
    c->h = h;
  }
}

// ---------------------------------------------------------------------------
// Emitter (EmitterSystem.cpp:54-83, Emitter.cpp:40-87, 89-227) — box emitter (type 0), no particle reuse
// ---------------------------------------------------------------------------
void emitParticles(dfr_context *c) {
  if (c->emitters.empty()) return;
  // EmitterSystem::step (:64-76): particles animated in the previous step become Active again
  for (int64_t i = 0; i < c->nf; i++)
    if (c->state[i] == 1) c->state[i] = 0;
  const double t = c->time;
  const double timeStepSize = c->h;
  const double radius = c->cfg.particle_radius;
  const double diam = 2.0 * radius;
  for (auto &e : c->emitters) {
    const Vec3 emitDir = vec3(e.rot(0, 0), e.rot(1, 0), e.rot(2, 0));
    Vec3 emitVel = e.velocity * emitDir;
    if (t < e.emitStart || t > e.emitEnd) emitVel = emitDir * radius * 10 * (1.0 / 0.25);
    if (t >= e.emitStart - 0.25 && t <= e.emitEnd) {
      const double animationMarginAhead = c->supportRadius;
      // getSize, type 0, Akinci2012 (:52-58)
      const Vec3 size = vec3(2 * c->supportRadius, e.height * diam + 2 * diam, e.width * diam + 2 * diam);
      const Vec3 halfSize = 0.5 * size;
      const Vec3 pos = e.x + (0.5 * animationMarginAhead) * emitDir;
      const Mat3 rotT = transpose(e.rot);
      for (int64_t i = 0; i < c->nf; i++) {
        const Vec3 xi = ld3(&c->x[3 * i]);
        const Vec3 xl = rotT * (xi - pos);  // inBox (Emitter.h:35-41)
        if (std::fabs(X(xl)) < X(halfSize) && std::fabs(Y(xl)) < Y(halfSize) && std::fabs(Z(xl)) < Z(halfSize)) {
          st3(&c->v[3 * i], emitVel);
          st3(&c->x[3 * i], xi + timeStepSize * emitVel);
          c->state[i] = 1;
        }
      }
    }
    if (t < e.nextEmitTime || t > e.emitEnd) continue;
    const Vec3 axisHeight = vec3(e.rot(0, 1), e.rot(1, 1), e.rot(2, 1));
    const Vec3 axisWidth = vec3(e.rot(0, 2), e.rot(1, 2), e.rot(2, 2));
    const double startX = -0.5 * (e.width - 1) * diam;
    const double startZ = -0.5 * (e.height - 1) * diam;
    const double dt = t - e.nextEmitTime + timeStepSize;
    const Vec3 offset = e.x + dt * emitVel;
    if (c->nf < c->nfCapacity) {
      int64_t index = c->nf;
      int64_t numEmitted = 0;
      for (int i = 0; i < e.width; i++)
        for (int j = 0; j < e.height; j++) {
          if (index < c->nfCapacity) {
            st3(&c->x[3 * index], (i * diam + startX) * axisWidth + (j * diam + startZ) * axisHeight + offset);
            st3(&c->v[3 * index], emitVel);
            c->state[index] = 1;
            c->kappa[index] = 0.0;   // SimulationDataDiffDFSPH::emittedParticles (:173-183)
            c->kappaV[index] = 0.0;
            numEmitted++;
          }
          index++;
        }
      c->nf += numEmitted;
    }
    e.nextEmitTime += diam / e.velocity;
    e.emitCounter++;
  }
}

// ---------------------------------------------------------------------------
// per-body chain rule (BoundaryModel_Akinci2012.cpp:453-889)
// ---------------------------------------------------------------------------
void accumulate_and_reset_gradient(Body &b) {  // :453-497
  b.net_force_vn = b.net_force_xn = b.net_force_omega_n = Mat3::zero();
  b.net_force_qn = Mat34::zero();
  b.net_torque_omega_n = b.net_torque_vn = b.net_torque_xn = Mat3::zero();
  b.net_torque_qn = Mat34::zero();
  for (int64_t i = 0; i < b.n; i++) {
    b.net_force_vn += b.g_force_v[i];
    b.net_force_xn += b.g_force_x[i];
    b.net_force_omega_n += b.g_force_omega[i];
    b.net_force_qn += b.g_force_q[i];
    b.net_torque_omega_n += b.g_torque_omega[i];
    b.net_torque_qn += b.g_torque_q[i];
    b.net_torque_vn += b.g_torque_v[i];
    b.net_torque_xn += b.g_torque_x[i];
    b.g_force_v[i] = b.g_force_x[i] = b.g_force_omega[i] = Mat3::zero();
    b.g_force_q[i] = Mat34::zero();
    b.g_torque_omega[i] = b.g_torque_v[i] = b.g_torque_x[i] = Mat3::zero();
    b.g_torque_q[i] = Mat34::zero();
  }
}

// compute_grad_inertia_v_to_* (BoundaryModel_Akinci2012.cpp:500-541; manager :241-288)
Mat3 grad_inertia_v(const Body &b, const Mat43 &dq, const Vec3 &v) {
  const Mat3 R = qrot(b.q);
  const Mat3 a = get_grad_Rqp_to_q(b.q, b.inertia0 * transpose(R) * v) * dq;
  const Mat3 bb = R * b.inertia0 * (get_grad_RqTp_to_q(b.q, v) * dq);
  return a + bb;
}

// perform_chain_rule (BoundaryModel_Akinci2012.cpp:543-889)
void perform_chain_rule(dfr_context *c, Body &b, double dt, bool optimize_rotation) {
  const int gradient_mode = c->cfg.gradient_mode;
  const bool gyro = (c->cfg.rigid_body_mode == 0);
  const double invMass = 1.0 / (b.mass + 1e-10);
  const Mat3 inertia = b.inertia, inv_inertia = b.invInertia;
  const Vec3 omega = b.omega;
  Vec3 temp_v = Vec3::zero();

  if (gradient_mode == 0) {  // Complete (:569-672)
    const Mat3 f_v0 = b.net_force_xn * b.grad_x_to_v0 + b.net_force_vn * b.grad_v_to_v0 + b.net_force_qn * b.grad_q_to_v0 + b.net_force_omega_n * b.grad_omega_to_v0;
    const Mat3 f_w0 = b.net_force_xn * b.grad_x_to_omega0 + b.net_force_vn * b.grad_v_to_omega0 + b.net_force_qn * b.grad_q_to_omega0 + b.net_force_omega_n * b.grad_omega_to_omega0;
    b.grad_v_to_v0 = b.grad_v_to_v0 + (dt * invMass) * f_v0;
    b.grad_v_to_omega0 = b.grad_v_to_omega0 + (dt * invMass) * f_w0;
    b.grad_x_to_v0 = b.grad_x_to_v0 + dt * b.grad_v_to_v0;
    b.grad_x_to_omega0 = b.grad_x_to_omega0 + dt * b.grad_v_to_omega0;
    Vec3 force, torque;
    b.getForceAndTorque(force, torque);
    const Mat3 t_w0 = b.net_torque_xn * b.grad_x_to_omega0 + b.net_torque_vn * b.grad_v_to_omega0 + b.net_torque_qn * b.grad_q_to_omega0 + b.net_torque_omega_n * b.grad_omega_to_omega0;
    const Mat3 t_v0 = b.net_torque_xn * b.grad_x_to_v0 + b.net_torque_vn * b.grad_v_to_v0 + b.net_torque_qn * b.grad_q_to_v0 + b.net_torque_omega_n * b.grad_omega_to_v0;
    Mat3 Tau_w0, Tau_v0;
    if (gyro) {
      temp_v = cross(inertia * omega, omega) + torque;
      const Mat3 L_w0 = skewMatrix(inertia * omega) * b.grad_omega_to_omega0 + transpose(skewMatrix(omega)) * (inertia * b.grad_omega_to_omega0 + grad_inertia_v(b, b.grad_q_to_omega0, omega));
      const Mat3 L_v0 = skewMatrix(inertia * omega) * b.grad_omega_to_v0 + transpose(skewMatrix(omega)) * (inertia * b.grad_omega_to_v0 + grad_inertia_v(b, b.grad_q_to_v0, omega));
      Tau_w0 = L_w0 + t_w0;
      Tau_v0 = L_v0 + t_v0;
    } else {
      temp_v = torque;
      Tau_w0 = t_w0;
      Tau_v0 = t_v0;
    }
    const Mat3 gi_w0 = inv_inertia * grad_inertia_v(b, b.grad_q_to_omega0, inv_inertia * temp_v);
    const Mat3 gi_v0 = inv_inertia * grad_inertia_v(b, b.grad_q_to_v0, inv_inertia * temp_v);
    b.grad_omega_to_omega0 = b.grad_omega_to_omega0 + dt * (gi_w0 + inv_inertia * Tau_w0);
    b.grad_omega_to_v0 = b.grad_omega_to_v0 + dt * (gi_v0 + inv_inertia * Tau_v0);
  } else if (gradient_mode == 2) {  // RigidGradOnly (:674-709)
    b.grad_x_to_v0 = b.grad_x_to_v0 + dt * b.grad_v_to_v0;
    b.grad_x_to_omega0 = b.grad_x_to_omega0 + dt * b.grad_v_to_omega0;
    if (gyro) {
      temp_v = cross(inertia * omega, omega);
      const Mat3 gi_w0 = inv_inertia * grad_inertia_v(b, b.grad_q_to_omega0, inv_inertia * temp_v);
      const Mat3 gi_v0 = inv_inertia * grad_inertia_v(b, b.grad_q_to_v0, inv_inertia * temp_v);
      const Mat3 L_w0 = skewMatrix(inertia * omega) * b.grad_omega_to_omega0 + transpose(skewMatrix(omega)) * (inertia * b.grad_omega_to_omega0 + grad_inertia_v(b, b.grad_q_to_omega0, omega));
      const Mat3 L_v0 = skewMatrix(inertia * omega) * b.grad_omega_to_v0 + transpose(skewMatrix(omega)) * (inertia * b.grad_omega_to_v0 + grad_inertia_v(b, b.grad_q_to_v0, omega));
      b.grad_omega_to_omega0 = b.grad_omega_to_omega0 + dt * (gi_w0 + inv_inertia * L_w0);
      b.grad_omega_to_v0 = b.grad_omega_to_v0 + dt * (gi_v0 + inv_inertia * L_v0);
    }
  } else {  // Incomplete (:711-826)
    const Mat3 f_v0 = b.net_force_vn * b.grad_v_to_v0 + b.net_force_omega_n * b.grad_omega_to_v0;
    const Mat3 f_w0 = b.net_force_vn * b.grad_v_to_omega0 + b.net_force_omega_n * b.grad_omega_to_omega0;
    b.grad_v_to_v0 = b.grad_v_to_v0 + (dt * invMass) * f_v0;
    b.grad_v_to_omega0 = b.grad_v_to_omega0 + (dt * invMass) * f_w0;
    b.grad_x_to_v0 = b.grad_x_to_v0 + dt * b.grad_v_to_v0;
    b.grad_x_to_omega0 = b.grad_x_to_omega0 + dt * b.grad_v_to_omega0;
    const Mat3 t_w0 = b.net_torque_vn * b.grad_v_to_omega0 + b.net_torque_omega_n * b.grad_omega_to_omega0;
    const Mat3 t_v0 = b.net_torque_vn * b.grad_v_to_v0 + b.net_torque_omega_n * b.grad_omega_to_v0;
    Vec3 force, torque;
    b.getForceAndTorque(force, torque);
    Mat3 Tau_w0, Tau_v0;
    if (gyro) {
      temp_v = cross(inertia * omega, omega) + torque;
      const Mat3 L_w0 = skewMatrix(inertia * omega) * b.grad_omega_to_omega0 + transpose(skewMatrix(omega)) * (inertia * b.grad_omega_to_omega0 + grad_inertia_v(b, b.grad_q_to_omega0, omega));
      const Mat3 L_v0 = skewMatrix(inertia * omega) * b.grad_omega_to_v0 + transpose(skewMatrix(omega)) * (inertia * b.grad_omega_to_v0 + grad_inertia_v(b, b.grad_q_to_v0, omega));
      Tau_w0 = L_w0 + t_w0;
      Tau_v0 = L_v0 + t_v0;
    } else {
      temp_v = torque;
      Tau_w0 = t_w0;
      Tau_v0 = t_v0;
    }
    const Mat3 gi_w0 = inv_inertia * grad_inertia_v(b, b.grad_q_to_omega0, inv_inertia * temp_v);
    const Mat3 gi_v0 = inv_inertia * grad_inertia_v(b, b.grad_q_to_v0, inv_inertia * temp_v);
    b.grad_omega_to_omega0 = b.grad_omega_to_omega0 + dt * (gi_w0 + inv_inertia * Tau_w0);
    b.grad_omega_to_v0 = b.grad_omega_to_v0 + dt * (gi_v0 + inv_inertia * Tau_v0);
  }

  if (optimize_rotation) {  // :827-870
    const Quat q = b.q;
    const Vec3 new_omega = omega + dt * (inv_inertia * temp_v);
    const Quat p = {0.0, X(new_omega), Y(new_omega), Z(new_omega)};
    const Quat pq = qmul(p, q);
    Quat new_q = {q.w + dt * 0.5 * pq.w, q.x + dt * 0.5 * pq.x, q.y + dt * 0.5 * pq.y, q.z + dt * 0.5 * pq.z};
    const Quat new_qn = qnormalized(new_q);
    Vec4 new_qnv;
    new_qnv(0, 0) = new_qn.w;
    new_qnv(1, 0) = new_qn.x;
    new_qnv(2, 0) = new_qn.y;
    new_qnv(3, 0) = new_qn.z;
    const Mat4 grad_p_q_product_to_q = get_grad_p_q_product_to_q(p);
    const Mat43 grad_p_q_product_to_omega = get_grad_omega_q_product_to_omega(q);
    const Mat4 grad_normalized_q_to_q = (Mat4::identity() - new_qnv * transpose(new_qnv)) * (1.0 / qnorm(new_q));
    b.partial_grad_qn_to_omega_n = grad_normalized_q_to_q * ((dt / 2.) * grad_p_q_product_to_omega);
    b.grad_q_to_omega0 = grad_normalized_q_to_q * (b.grad_q_to_omega0 + (dt / 2.) * (grad_p_q_product_to_omega * b.grad_omega_to_omega0 + grad_p_q_product_to_q * b.grad_q_to_omega0));
    b.grad_q_to_v0 = grad_normalized_q_to_q * (b.grad_q_to_v0 + (dt / 2.) * (grad_p_q_product_to_omega * b.grad_omega_to_v0 + grad_p_q_product_to_q * b.grad_q_to_v0));
  }
}

// ---------------------------------------------------------------------------
// RigidBodyGradientManager (RigidBodyGradientManager.cpp:96-475)
// ---------------------------------------------------------------------------
void mgr_force_torque_chain(dfr_context *c, bool rigid) {  // :96-156 (fluid) / :158-239 (rigid)
  Manager &m = c->mgr;
  const int n = m.n;
  for (int k = 0; k < n * n; k++) m.f_v0[k] = m.f_w0[k] = m.t_v0[k] = m.t_w0[k] = Mat3::zero();
  for (int R = 0; R < n; R++) {
    if (!c->bodies[R].dynamic) continue;
    for (int RR = 0; RR < n; RR++) {
      if (!c->bodies[RR].dynamic) continue;
      for (int Rk = 0; Rk < n; Rk++) {
        if (!c->bodies[Rk].dynamic) continue;
        const int a = m.at(R, Rk), bq = m.at(Rk, RR), o = m.at(R, RR);
        if (!rigid) {
          m.f_v0[o] += m.f_vn[a] * m.vn_v0[bq] + m.f_wn[a] * m.wn_v0[bq];
          m.t_v0[o] += m.t_vn[a] * m.vn_v0[bq] + m.t_wn[a] * m.wn_v0[bq];
          m.f_w0[o] += m.f_vn[a] * m.vn_w0[bq] + m.f_wn[a] * m.wn_w0[bq];
          m.t_w0[o] += m.t_vn[a] * m.vn_w0[bq] + m.t_wn[a] * m.wn_w0[bq];
        } else {
          // :199-200 and :204-205 end the sum with ';' before the q-term, so grad_*_to_qn * grad_qn_to_v0
          // is a discarded expression in the v0 blocks; the omega0 blocks (:216-224) keep it.
          m.f_v0[o] += m.f_vn[a] * m.vn_v0[bq] + m.f_wn[a] * m.wn_v0[bq] + m.f_xn[a] * m.xn_v0[bq];
          m.t_v0[o] += m.t_vn[a] * m.vn_v0[bq] + m.t_wn[a] * m.wn_v0[bq] + m.t_xn[a] * m.xn_v0[bq];
          m.f_w0[o] += m.f_vn[a] * m.vn_w0[bq] + m.f_wn[a] * m.wn_w0[bq] + m.f_xn[a] * m.xn_w0[bq] + m.f_qn[a] * m.qn_w0[bq];
          m.t_w0[o] += m.t_vn[a] * m.vn_w0[bq] + m.t_wn[a] * m.wn_w0[bq] + m.t_xn[a] * m.xn_w0[bq] + m.t_qn[a] * m.qn_w0[bq];
        }
      }
    }
  }
}

void mgr_velocity_chain(dfr_context *c) {  // :292-393
  Manager &m = c->mgr;
  const int n = m.n;
  const bool gyro = (c->cfg.rigid_body_mode == 0);
  const double dt = c->h;
  for (int R = 0; R < n; R++) {
    Body &b = c->bodies[R];
    if (!b.dynamic) continue;
    const Mat3 inertia = b.inertia, inv_inertia = b.invInertia;
    const Vec3 omega = b.omega;
    Vec3 temp_v = Vec3::zero();
    Vec3 force, torque;
    b.getForceAndTorque(force, torque);
    for (int RR = 0; RR < n; RR++) {
      if (!c->bodies[RR].dynamic) continue;
      const int o = m.at(R, RR);
      m.vn_v0[o] = m.vn_v0[o] + (dt * b.invMass) * m.f_v0[o];
      m.vn_w0[o] = m.vn_w0[o] + (dt * b.invMass) * m.f_w0[o];
      Mat3 Tau_w0, Tau_v0;
      if (gyro) {
        temp_v = cross(inertia * omega, omega) + torque;
        const Mat3 L_w0 = skewMatrix(inertia * omega) * m.wn_w0[o] + transpose(skewMatrix(omega)) * (inertia * m.wn_w0[o] + grad_inertia_v(b, m.qn_w0[o], omega));
        const Mat3 L_v0 = skewMatrix(inertia * omega) * m.wn_v0[o] + transpose(skewMatrix(omega)) * (inertia * m.wn_v0[o] + grad_inertia_v(b, m.qn_v0[o], omega));
        Tau_w0 = L_w0 + m.t_w0[o];
        Tau_v0 = L_v0 + m.t_v0[o];
      } else {
        temp_v = torque;
        Tau_w0 = m.t_w0[o];
        Tau_v0 = m.t_v0[o];
      }
      const Mat3 gi_w0 = inv_inertia * grad_inertia_v(b, m.qn_w0[o], inv_inertia * temp_v);
      const Mat3 gi_v0 = inv_inertia * grad_inertia_v(b, m.qn_v0[o], inv_inertia * temp_v);
      m.wn_w0[o] = m.wn_w0[o] + dt * (gi_w0 + inv_inertia * Tau_w0);
      m.wn_v0[o] = m.wn_v0[o] + dt * (gi_v0 + inv_inertia * Tau_v0);
    }
  }
}

void mgr_position_rotation_chain(dfr_context *c) {  // :395-453
  Manager &m = c->mgr;
  const int n = m.n;
  const double dt = c->h;
  for (int R = 0; R < n; R++) {
    Body &b = c->bodies[R];
    if (!b.dynamic) continue;
    const Quat q = b.q;
    const Vec3 new_omega = b.omega;  // omega is already updated here (:411)
    const Quat p = {0.0, X(new_omega), Y(new_omega), Z(new_omega)};
    const Quat pq = qmul(p, q);
    Quat new_q = {q.w + dt * 0.5 * pq.w, q.x + dt * 0.5 * pq.x, q.y + dt * 0.5 * pq.y, q.z + dt * 0.5 * pq.z};
    const Quat new_qn = qnormalized(new_q);
    Vec4 new_qnv;
    new_qnv(0, 0) = new_qn.w;
    new_qnv(1, 0) = new_qn.x;
    new_qnv(2, 0) = new_qn.y;
    new_qnv(3, 0) = new_qn.z;
    const Mat4 gpq = get_grad_p_q_product_to_q(p);
    const Mat43 gpo = get_grad_omega_q_product_to_omega(q);
    const Mat4 gn = (Mat4::identity() - new_qnv * transpose(new_qnv)) * (1.0 / qnorm(new_q));
    for (int RR = 0; RR < n; RR++) {
      if (!c->bodies[RR].dynamic) continue;
      const int o = m.at(R, RR);
      m.xn_v0[o] += dt * m.vn_v0[o];
      m.xn_w0[o] += dt * m.vn_w0[o];
      m.qn_v0[o] = gn * (m.qn_v0[o] + (dt / 2.) * (gpo * m.wn_v0[o] + gpq * m.qn_v0[o]));
      m.qn_w0[o] = gn * (m.qn_w0[o] + (dt / 2.) * (gpo * m.wn_w0[o] + gpq * m.qn_w0[o]));
    }
  }
}

// update_rigid_body_gradient_manager (BoundaryModel_Akinci2012.cpp:952-967)
void update_rigid_body_gradient_manager(dfr_context *c, int R) {
  Manager &m = c->mgr;
  const Body &b = c->bodies[R];
  const int o = m.at(R, R);
  m.f_vn[o] = b.net_force_vn;
  m.f_xn[o] = b.net_force_xn;
  m.f_wn[o] = b.net_force_omega_n;
  m.f_qn[o] = b.net_force_qn;
  m.t_vn[o] = b.net_torque_vn;
  m.t_xn[o] = b.net_torque_xn;
  m.t_wn[o] = b.net_torque_omega_n;
  m.t_qn[o] = b.net_torque_qn;
}

// ---------------------------------------------------------------------------
// rigid bodies (Dynamic3dRigidBody.h:113-154, 230-235; BoundarySimulator.cpp:10-36;
// RigidBody3dBoundarySimulator.cpp:278-344)
// ---------------------------------------------------------------------------
inline void rb_addForce(dfr_context *c, Body &b, const Vec3 &f) { b.vel += (b.invMass * f) * c->h; }
inline void rb_addTorque(dfr_context *c, Body &b, const Vec3 &t) {
  const double dt = c->h;
  if (c->cfg.rigid_body_mode == 0) {
    const Vec3 L = b.inertia * b.omega;
    b.omega += (b.invInertia * (cross(L, b.omega) + t)) * dt;
  } else
    b.omega += (b.invInertia * t) * dt;
}
inline void rb_addGravity(dfr_context *c, Body &b) { rb_addForce(c, b, ld3(c->cfg.gravitation) * b.mass); }
inline void rb_animate(dfr_context *c, Body &b) {
  const double dt = c->h;
  b.pos += b.vel * dt;
  const Quat angVelQ = {0.0, X(b.omega), Y(b.omega), Z(b.omega)};
  const Quat d = qmul(angVelQ, b.q);
  Quat nq = {b.q.w + dt * 0.5 * d.w, b.q.x + dt * 0.5 * d.x, b.q.y + dt * 0.5 * d.y, b.q.z + dt * 0.5 * d.z};
  b.q = qnormalized(nq);
  b.updateInertia();
}

#include "oracle_contact.inc"

void velocityTimeStep(dfr_context *c) {  // RigidBody3dBoundarySimulator.cpp:278-320
  if (c->cfg.use_rigid_contact_solver) {
    for (auto &b : c->bodies)
      if (b.dynamic) rb_addGravity(c, b);
    solveRigidContact(c);
  } else {
    // BoundarySimulator::updateBoundaryForces (BoundarySimulator.cpp:10-36)
    for (auto &b : c->bodies) {
      if (b.dynamic) {
        if (!b.animated) {
          Vec3 force, torque;
          b.getForceAndTorque(force, torque);
          rb_addForce(c, b, force);
          rb_addTorque(c, b, torque);
        }
        b.clearForceAndTorque();
      }
    }
    for (auto &b : c->bodies)
      if (b.dynamic && !b.animated) rb_addGravity(c, b);
  }
}
void positionTimeStep(dfr_context *c) {  // RigidBody3dBoundarySimulator.cpp:322-344
  for (auto &b : c->bodies)
    if (b.dynamic) rb_animate(c, b);
  updateBoundaryParticles(c, false);
}

// ---------------------------------------------------------------------------
// TimeStepDiffDFSPH::beginStep / endStep / backwardPerStep / step (:353-652)
// ---------------------------------------------------------------------------
void beginStep(dfr_context *c) {
  c->step_count++;
  for (auto &b : c->bodies) {
    if (!b.dynamic) continue;
    const double T = c->cfg.uniform_acc_rb_time;
    if (c->cfg.use_release_rigid_body_mode) {  // :381-407: only boundary model 1 is touched
      if (&b != &c->bodies[1]) continue;
      if (c->time <= T + c->h) {
        if (T > 1e-3) b.animated = true;
        b.vel = Vec3::zero();
        b.omega = Vec3::zero();
      } else {
        b.vel = b.init_v;
        b.omega = b.init_omega;
        b.animated = false;
      }
      continue;
    }
    if (c->time <= T + c->h) {
      double factor = 1.;
      if (T > 1e-3) {
        factor = (c->time / T) > 1. ? 1. : (c->time / T);
        b.animated = true;
      }
      b.vel = factor * b.init_v;
      b.omega = factor * b.init_omega;
    } else
      b.animated = false;
  }
}
void backwardPerStep(dfr_context *c) {
  const double h = c->h;
  for (size_t R = 0; R < c->bodies.size(); R++) {
    Body &b = c->bodies[R];
    if (!b.dynamic) continue;
    if (!b.animated) {
      accumulate_and_reset_gradient(b);
      if (c->cfg.use_rigid_gradient_manager)
        update_rigid_body_gradient_manager(c, (int)R);
      else
        perform_chain_rule(c, b, h, c->cfg.optimize_rotation != 0);
    }
  }
}
void endStep(dfr_context *c) { c->finished = (c->time >= c->cfg.target_time + c->cfg.uniform_acc_rb_time); }

void fluidStep(dfr_context *c) {
  beginStep(c);
  const double h = c->h;  // OLD h (:535)
  if (c->cfg.use_rigid_contact_solver) {  // performNeighborhoodSearch (:2044-2056): z-sort every 500 steps
    if (c->sortCounter % 500 == 0) contactSort(c);
    c->sortCounter++;
  }
  findNeighbors(c);
  computeDensities(c);
  computeDFSPHFactor(c);
  if (c->cfg.enable_divergence_solver)
    divergenceSolve(c);
  else
    c->iterationsV = 0;
  // clearAccelerations (TimeStep.cpp:66-81)
  for (int64_t i = 0; i < c->nf; i++) st3(&c->a[3 * i], ld3(c->cfg.gravitation));
  if (c->cfg.surface_tension_method == 2) surfaceTensionAkinci2013(c);
  if (c->cfg.viscosity_method == 1) viscosityStandard(c);
  updateTimeStepSize(c);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < c->nf; i++)
    if (c->state[i] == 0) st3(&c->v[3 * i], ld3(&c->v[3 * i]) + h * ld3(&c->a[3 * i]));
  pressureSolve(c);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < c->nf; i++)
    if (c->state[i] == 0) st3(&c->x[3 * i], ld3(&c->x[3 * i]) + h * ld3(&c->v[3 * i]));
  c->totalParticleSteps += c->nf;
  emitParticles(c);
  c->time += h;
  backwardPerStep(c);
  endStep(c);
}

// SimulatorBase::timeStepNoGUI (SimulatorBase.cpp:1142-1169)
void timeStep(dfr_context *c) {
  fluidStep(c);
  if (c->cfg.use_rigid_gradient_manager && c->nf > 0) {
    mgr_force_torque_chain(c, false);
    mgr_velocity_chain(c);
  }
  velocityTimeStep(c);
  if (c->cfg.use_rigid_gradient_manager) {
    if (c->cfg.use_rigid_contact_solver) {
      mgr_force_torque_chain(c, true);
      mgr_velocity_chain(c);
    }
    mgr_position_rotation_chain(c);
  }
  positionTimeStep(c);
}

int fail(dfr_context *c, int code, const char *msg) {
  if (c) c->err = msg;
  return code;
}

void storeMat(double *out, const Mat3 &m) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) out[3 * i + j] = m(i, j);
}
void storeMat(double *out, const Mat43 &m) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 3; j++) out[3 * i + j] = m(i, j);
}
void storeMat(double *out, const Mat34 &m) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++) out[4 * i + j] = m(i, j);
}

}  // namespace

// ===========================================================================
// exported API (mirrors include/dfr.h with the prefix orc_)
// ===========================================================================
extern "C" {

void orc_default_config(dfr_config *cfg) {
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->particle_radius = 0.025;
  cfg->density0 = 1000.0;
  cfg->gravitation[0] = 0.0;
  cfg->gravitation[1] = -9.81;
  cfg->gravitation[2] = 0.0;
  cfg->cfl_method = 1;
  cfg->cfl_factor = 0.5;
  cfg->cfl_min_time_step = 0.0001;
  cfg->cfl_max_time_step = 0.005;
  cfg->time_step_size = 0.001;
  cfg->min_iterations = 2;
  cfg->max_iterations = 100;
  cfg->max_error = 0.01;
  cfg->max_iterations_v = 100;
  cfg->max_error_v = 0.1;
  cfg->enable_divergence_solver = 1;
  cfg->use_pressure_warmstart = 1;
  cfg->use_divergence_warmstart = 1;
  cfg->viscosity_method = 1;
  cfg->viscosity = 0.01;
  cfg->viscosity_boundary = 0.0;
  cfg->surface_tension_method = 0;
  cfg->surface_tension = 0.05;
  cfg->surface_tension_boundary = 0.0;
  cfg->gradient_mode = 1;
  cfg->rigid_body_mode = 0;
  cfg->optimize_rotation = 1;
  cfg->use_rigid_gradient_manager = 0;
  cfg->use_rigid_contact_solver = 0;
  cfg->rigid_contact_beta = 1.0;
  cfg->rigid_contact_gamma = 0.7;
  cfg->rigid_contact_friction = 0.0;
  cfg->rigid_contact_support_radius_factor = 4.0;
  cfg->target_time = 0.8;
  cfg->uniform_acc_rb_time = 0.0;
  cfg->max_emitted_particles = 0;
}

int orc_create(const dfr_config *cfg, int device, dfr_context **out) {
  (void)device;
  if (!cfg || !out) return DFR_ERR_INVALID;
  dfr_context *c = new dfr_context();
  c->cfg = *cfg;
  c->supportRadius = 4.0 * cfg->particle_radius;  // Simulation.cpp:382-386
  c->K.set_radius(c->supportRadius);
  const double diam = 2.0 * cfg->particle_radius;
  c->Vf = 0.8 * diam * diam * diam;  // FluidModel.cpp:255-275
  c->mass = c->Vf * cfg->density0;
  c->h = cfg->time_step_size;
#ifdef _OPENMP
  c->nthreads = omp_get_max_threads();
#endif
  *out = c;
  return DFR_OK;
}
void orc_destroy(dfr_context *c) { delete c; }
const char *orc_last_error(const dfr_context *c) { return c ? c->err.c_str() : "null context"; }

int orc_set_fluid(dfr_context *c, int64_t n, const double *x, const double *v) {
  if (!c || c->finalized || n < 0) return fail(c, DFR_ERR_STATE, "set_fluid after finalize");
  const int64_t cap = n + c->cfg.max_emitted_particles;
  c->nf = c->nfActive0 = n;
  c->nfCapacity = cap;
  c->x.assign(3 * cap, 0.0);
  c->v.assign(3 * cap, 0.0);
  if (n) std::memcpy(c->x.data(), x, sizeof(double) * 3 * n);
  if (n && v) std::memcpy(c->v.data(), v, sizeof(double) * 3 * n);
  c->a.assign(3 * cap, 0.0);
  c->density.assign(cap, 0.0);
  c->pressure.assign(cap, 0.0);
  c->factor.assign(cap, 0.0);
  c->kappa.assign(cap, 0.0);
  c->kappaV.assign(cap, 0.0);
  c->densityAdv.assign(cap, 0.0);
  c->sum_grad_p_k.assign(3 * cap, 0.0);
  c->normals.assign(3 * cap, 0.0);
  c->state.assign(cap, 0);
  return DFR_OK;
}

int orc_add_body(dfr_context *c, int64_t n, const double *x_local, int is_dynamic, double density, const double position[3], const double quat_wxyz[4]) {
  if (!c || c->finalized) return fail(c, DFR_ERR_STATE, "add_body after finalize");
  c->bodies.emplace_back();
  Body &b = c->bodies.back();
  b.n = n;
  b.dynamic = is_dynamic != 0;
  b.density = density;
  b.x0.assign(x_local, x_local + 3 * n);
  b.x = b.x0;
  b.v.assign(3 * n, 0.0);
  b.V.assign(n, 0.0);
  b.pos0 = b.pos = ld3(position);
  b.q0 = b.q = Quat{quat_wxyz[0], quat_wxyz[1], quat_wxyz[2], quat_wxyz[3]};
  b.vel = b.omega = b.init_v = b.init_omega = Vec3::zero();
  // determineMassProperties (Dynamic3dRigidBody.h:184-198)
  const double r = c->cfg.particle_radius;
  const double volume = (4.0 / 3.0 * M_PI) * r * r * r;
  const double deltaMass = volume * density;
  b.mass = 0.0;
  b.inertia0 = Mat3::zero();
  for (int64_t i = 0; i < n; i++) {
    b.mass += deltaMass;
    const Vec3 rr = ld3(&x_local[3 * i]);
    b.inertia0 += deltaMass * (dot(rr, rr) * Mat3::identity() - outer(rr, rr));
  }
  b.invMass = 1. / b.mass;
  b.updateInertia();
  const int nt = b.dynamic ? c->nthreads : 0;
  b.forcePerThread.assign(nt, Vec3::zero());
  b.torquePerThread.assign(nt, Vec3::zero());
  b.forceBackup.assign(nt, Vec3::zero());
  b.torqueBackup.assign(nt, Vec3::zero());
  b.g_force_v.resize(n); b.g_force_x.resize(n); b.g_force_omega.resize(n);
  b.g_torque_omega.resize(n); b.g_torque_x.resize(n); b.g_torque_v.resize(n);
  b.g_force_q.resize(n); b.g_torque_q.resize(n);
  b.resetGradient();
  return (int)c->bodies.size() - 1;
}

int orc_set_init_v_omega(dfr_context *c, int body, const double v0[3], const double omega0[3]) {
  if (!c || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  c->bodies[body].init_v = ld3(v0);
  c->bodies[body].init_omega = ld3(omega0);
  return DFR_OK;
}

static void snapshot(dfr_context *c) {
  c->x_init = c->x;
  c->v_init = c->v;
  c->kappa_init = c->kappa;
  c->kappaV_init = c->kappaV;
}

int orc_finalize(dfr_context *c) {
  if (!c || c->finalized) return fail(c, DFR_ERR_STATE, "already finalized");
  if (c->x.empty() && c->nfCapacity == 0) orc_set_fluid(c, 0, nullptr, nullptr);
  c->mgr.init((int)c->bodies.size());
  contactInit(c);                    // RigidContactSolver ctor + first z-sort: before the particles are moved to world space (:250-257)
  updateBoundaryParticles(c, true);  // RigidBody3dBoundarySimulator::deferredInit (:256-262)
  updateBoundaryVolume(c);
  snapshot(c);
  c->finalized = true;
  return DFR_OK;
}

int orc_reset(dfr_context *c);
int orc_load_fluid_state(dfr_context *c, const double *x, const double *v, const double *kappa, const double *kappa_v) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  const int64_t n = c->nfActive0;
  if (x) std::memcpy(c->x.data(), x, sizeof(double) * 3 * n);
  if (v) std::memcpy(c->v.data(), v, sizeof(double) * 3 * n);
  if (kappa) std::memcpy(c->kappa.data(), kappa, sizeof(double) * n);
  if (kappa_v) std::memcpy(c->kappaV.data(), kappa_v, sizeof(double) * n);
  snapshot(c);
  return orc_reset(c);  // checkLoadState is part of SimulatorBase::reset (SimulatorBase.cpp:887-934); the reference driver resets too
}

int orc_reset(dfr_context *c) {  // SimulatorBase::reset (SimulatorBase.cpp:887-934)
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  c->nf = c->nfActive0;
  c->x = c->x_init;
  c->v = c->v_init;
  c->kappa = c->kappa_init;
  c->kappaV = c->kappaV_init;
  std::fill(c->a.begin(), c->a.end(), 0.0);
  std::fill(c->density.begin(), c->density.end(), 0.0);
  std::fill(c->sum_grad_p_k.begin(), c->sum_grad_p_k.end(), 0.0);
  std::fill(c->state.begin(), c->state.end(), 0);
  for (auto &b : c->bodies) {
    // BoundaryModel::reset + BoundaryModel_Akinci2012::reset + Dynamic3dRigidBody::reset
    for (auto &f : b.forcePerThread) f = Vec3::zero();
    for (auto &t : b.torquePerThread) t = Vec3::zero();
    b.pos = b.pos0;
    b.q = b.q0;
    b.updateInertia();
    b.vel = Vec3::zero();
    b.omega = Vec3::zero();
    b.animated = false;
    b.resetGradient();
  }
  for (auto &e : c->emitters) {
    e.nextEmitTime = e.emitStart;
    e.emitCounter = 0;
  }
  contactReset(c);  // incl. the z-sort of RigidBody3dBoundarySimulator::reset (:358), which sees the stale particle positions
  updateBoundaryParticles(c, true);
  updateBoundaryVolume(c);
  c->mgr.reset();
  c->time = 0.0;
  c->h = c->cfg.time_step_size;
  c->iterations = c->iterationsV = c->step_count = 0;
  c->finished = false;
  c->totalIter = c->totalIterV = c->totalParticleSteps = c->totalNeighbors = 0;
  c->cpu_ms = 0.0;
  return DFR_OK;
}

int orc_reset_gradient(dfr_context *c) {  // TimeStepDiffDFSPH::reset_gradient (TimeStepDiffDFSPH.cpp:2234-2240)
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  for (auto &b : c->bodies)
    if (b.dynamic) b.resetGradient();  // BoundaryModel_Akinci2012::reset_gradient (:62-108) zeroes/identities the same members
  return DFR_OK;
}

int orc_set_gradient_mode(dfr_context *c, int mode) {  // Simulation::setGradientMode (Simulation.h:173)
  if (!c) return DFR_ERR_INVALID;
  if (mode < 0 || mode > 2) return fail(c, DFR_ERR_INVALID, "gradient mode must be 0, 1 or 2");
  c->cfg.gradient_mode = mode;
  return DFR_OK;
}

int orc_step(dfr_context *c, int n_steps) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  const auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < n_steps; s++) timeStep(c);
  c->cpu_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return DFR_OK;
}

int orc_run_trajectory(dfr_context *c, int max_steps, int *steps_done) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  int s = 0;
  const auto t0 = std::chrono::steady_clock::now();
  while (s < max_steps) {
    timeStep(c);
    s++;
    if (c->finished) break;
  }
  c->cpu_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (steps_done) *steps_done = s;
  return DFR_OK;
}

int orc_get_step_info(dfr_context *c, dfr_step_info *info) {
  if (!c || !info) return DFR_ERR_INVALID;
  info->time = c->time;
  info->time_step_size = c->h;
  info->iterations = c->iterations;
  info->iterations_v = c->iterationsV;
  info->step_count = c->step_count;
  info->trajectory_finished = c->finished ? 1 : 0;
  info->num_fluid_particles = c->nf;
  info->total_pressure_iterations = c->totalIter;
  info->total_divergence_iterations = c->totalIterV;
  info->total_particle_steps = c->totalParticleSteps;
  info->total_fluid_neighbors = c->totalNeighbors;
  return DFR_OK;
}

int orc_get_body_state(dfr_context *c, int body, double out[13]) {
  if (!c || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  const Body &b = c->bodies[body];
  st3(out, b.pos);
  out[3] = b.q.w; out[4] = b.q.x; out[5] = b.q.y; out[6] = b.q.z;
  st3(out + 7, b.vel);
  st3(out + 10, b.omega);
  return DFR_OK;
}
int orc_set_body_velocity(dfr_context *c, int body, const double v[3], const double omega[3]) {
  if (!c || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  if (v) c->bodies[body].vel = ld3(v);
  if (omega) c->bodies[body].omega = ld3(omega);
  return DFR_OK;
}
int orc_get_body_properties(dfr_context *c, int body, double out[17]) {
  if (!c || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  const Body &b = c->bodies[body];
  out[0] = b.mass;
  out[1] = b.invMass;
  storeMat(out + 2, b.inertia0);
  st3(out + 11, b.getForce());
  st3(out + 14, b.getTorque());
  return DFR_OK;
}

int orc_get_body_grad(dfr_context *c, int body, int which, double out[12]) {
  if (!c || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  const Body &b = c->bodies[body];
  std::memset(out, 0, sizeof(double) * 12);
  switch (which) {
    case 0: storeMat(out, b.grad_x_to_v0); break;
    case 1: storeMat(out, b.grad_x_to_omega0); break;
    case 2: storeMat(out, b.grad_q_to_v0); break;
    case 3: storeMat(out, b.grad_q_to_omega0); break;
    case 4: storeMat(out, b.grad_v_to_v0); break;
    case 5: storeMat(out, b.grad_v_to_omega0); break;
    case 6: storeMat(out, b.grad_omega_to_v0); break;
    case 7: storeMat(out, b.grad_omega_to_omega0); break;
    case 8: storeMat(out, b.net_force_vn); break;
    case 9: storeMat(out, b.net_force_xn); break;
    case 10: storeMat(out, b.net_force_qn); break;
    case 11: storeMat(out, b.net_force_omega_n); break;
    case 12: storeMat(out, b.net_torque_vn); break;
    case 13: storeMat(out, b.net_torque_xn); break;
    case 14: storeMat(out, b.net_torque_qn); break;
    case 15: storeMat(out, b.net_torque_omega_n); break;
    default: return fail(c, DFR_ERR_INVALID, "bad gradient selector");
  }
  return DFR_OK;
}

int orc_get_manager_grad(dfr_context *c, int R, int RR, int which, double out[12]) {
  if (!c || R < 0 || RR < 0 || R >= c->mgr.n || RR >= c->mgr.n) return fail(c, DFR_ERR_INVALID, "bad body index");
  const Manager &m = c->mgr;
  const int o = m.at(R, RR);
  std::memset(out, 0, sizeof(double) * 12);
  switch (which) {
    case 0: storeMat(out, m.xn_v0[o]); break;
    case 1: storeMat(out, m.xn_w0[o]); break;
    case 2: storeMat(out, m.qn_v0[o]); break;
    case 3: storeMat(out, m.qn_w0[o]); break;
    case 4: storeMat(out, m.vn_v0[o]); break;
    case 5: storeMat(out, m.vn_w0[o]); break;
    case 6: storeMat(out, m.wn_v0[o]); break;
    case 7: storeMat(out, m.wn_w0[o]); break;
    case 8: storeMat(out, m.f_vn[o]); break;
    case 9: storeMat(out, m.f_xn[o]); break;
    case 10: storeMat(out, m.f_qn[o]); break;
    case 11: storeMat(out, m.f_wn[o]); break;
    case 12: storeMat(out, m.t_vn[o]); break;
    case 13: storeMat(out, m.t_xn[o]); break;
    case 14: storeMat(out, m.t_qn[o]); break;
    case 15: storeMat(out, m.t_wn[o]); break;
    default: return fail(c, DFR_ERR_INVALID, "bad gradient selector");
  }
  return DFR_OK;
}

int orc_download_fluid(dfr_context *c, int field, double *out) {
  if (!c || !out) return DFR_ERR_INVALID;
  const int64_t n = c->nf;
  switch (field) {
    case 0: std::memcpy(out, c->x.data(), sizeof(double) * 3 * n); break;
    case 1: std::memcpy(out, c->v.data(), sizeof(double) * 3 * n); break;
    case 2: std::memcpy(out, c->density.data(), sizeof(double) * n); break;
    case 3: std::memcpy(out, c->factor.data(), sizeof(double) * n); break;
    case 4: std::memcpy(out, c->kappa.data(), sizeof(double) * n); break;
    case 5: std::memcpy(out, c->kappaV.data(), sizeof(double) * n); break;
    case 6: std::memcpy(out, c->densityAdv.data(), sizeof(double) * n); break;
    case 7: std::memcpy(out, c->a.data(), sizeof(double) * 3 * n); break;
    case 8: std::memcpy(out, c->sum_grad_p_k.data(), sizeof(double) * 3 * n); break;
    case 9: std::memcpy(out, c->normals.data(), sizeof(double) * 3 * n); break;
    default: return fail(c, DFR_ERR_INVALID, "bad field");
  }
  return DFR_OK;
}
int orc_download_body(dfr_context *c, int body, int field, double *out) {
  if (!c || body < 0 || body >= (int)c->bodies.size() || !out) return fail(c, DFR_ERR_INVALID, "bad body index");
  const Body &b = c->bodies[body];
  switch (field) {
    case 0: std::memcpy(out, b.x.data(), sizeof(double) * 3 * b.n); break;
    case 1: std::memcpy(out, b.v.data(), sizeof(double) * 3 * b.n); break;
    case 2: std::memcpy(out, b.V.data(), sizeof(double) * b.n); break;
    case 3: std::memcpy(out, b.x0.data(), sizeof(double) * 3 * b.n); break;
    default: return fail(c, DFR_ERR_INVALID, "bad field");
  }
  return DFR_OK;
}
int64_t orc_num_fluid(dfr_context *c) { return c ? c->nf : 0; }
int64_t orc_num_fluid_initial(dfr_context *c) { return c ? c->nfActive0 : 0; }
int64_t orc_num_body_particles(dfr_context *c, int body) { return (c && body >= 0 && body < (int)c->bodies.size()) ? c->bodies[body].n : 0; }
int orc_num_bodies(dfr_context *c) { return c ? (int)c->bodies.size() : 0; }

int orc_get_neighbors(dfr_context *c, int set_a, int set_b, int32_t *counts, int32_t *indices, int64_t capacity, int64_t *total) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  const int nb = (int)c->bodies.size();
  if (set_a < -1 || set_a >= nb || set_b < -1 || set_b >= nb) return fail(c, DFR_ERR_INVALID, "bad set index");
  const double r2 = c->supportRadius * c->supportRadius;
  const double *xa = (set_a < 0) ? c->x.data() : c->bodies[set_a].x.data();
  const int64_t na = (set_a < 0) ? c->nf : c->bodies[set_a].n;
  const double *xb = (set_b < 0) ? c->x.data() : c->bodies[set_b].x.data();
  const int64_t nbp = (set_b < 0) ? c->nf : c->bodies[set_b].n;
  PointGrid g;
  g.build(xb, nbp, c->supportRadius);
  int64_t tot = 0;
  std::vector<int32_t> nbv;
  for (int64_t i = 0; i < na; i++) {
    nbv.clear();
    g.query(xb, &xa[3 * i], r2, (set_a == set_b) ? (int32_t)i : -1, nbv);
    if (counts) counts[i] = (int32_t)nbv.size();
    if (indices) {
      if (tot + (int64_t)nbv.size() > capacity) return fail(c, DFR_ERR_CAPACITY, "neighbour buffer too small");
      std::memcpy(indices + tot, nbv.data(), sizeof(int32_t) * nbv.size());
    }
    tot += (int64_t)nbv.size();
  }
  if (total) *total = tot;
  return DFR_OK;
}

int orc_add_emitter(dfr_context *c, int width, int height, const double position[3], const double rot[9], double velocity, double emit_start, double emit_end) {
  if (!c || c->finalized) return fail(c, DFR_ERR_STATE, "add_emitter after finalize");
  Emitter e;
  e.width = width;
  e.height = height;
  e.x = ld3(position);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) e.rot(i, j) = rot[3 * i + j];
  e.velocity = velocity;
  e.emitStart = emit_start;
  e.emitEnd = emit_end;
  e.nextEmitTime = emit_start;
  e.emitCounter = 0;
  c->emitters.push_back(e);
  return DFR_OK;
}

int orc_get_device_time_ms(dfr_context *c, double *total_ms, int64_t *kernel_launches) {
  if (!c) return DFR_ERR_INVALID;
  if (total_ms) *total_ms = c->cpu_ms;
  if (kernel_launches) *kernel_launches = 0;
  return DFR_OK;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
