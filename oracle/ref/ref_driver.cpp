// ============================================================================
// TEST INFRASTRUCTURE ONLY.  Driver of oracle/_ref/libref.so: the reference's OWN hot-path translation
// units (compiled by oracle/ref/Makefile from /root/reference, unmodified, against the shim headers in
// oracle/ref/shims/) behind the same C ABI as include/dfr.h, prefix ref_.
//
// What comes from the reference itself (compiled where it lies, never copied):
//   SPlisHSPlasH/DiffDFSPH/TimeStepDiffDFSPH.cpp, SimulationDataDiffDFSPH.cpp, TimeStep.cpp, BoundaryModel.cpp,
//   BoundaryModel_Akinci2012.cpp, RigidBodyGradientManager.cpp, GradientUtils.cpp, SPHKernels.cpp, Simulation.cpp,
//   FluidModel.cpp, TimeManager.cpp, Emitter.cpp, EmitterSystem.cpp, SurfaceTension_Akinci2013.cpp,
//   Viscosity_Standard.cpp, InterlinkedSPH/RigidContactSolver.cpp, Dynamic3dRigidBody.h,
//   Simulator/BoundarySimulator.cpp, Simulator/RigidBody3dBoundarySimulator.cpp, ...
// What this file supplies instead of Simulator/SimulatorBase.cpp (which drags in the GUI, exporters, the scene
// parser, partio and Discregrid and is therefore not built):
//   * scene construction from arrays (mirrors SimulatorBase::initSimulation/buildModel/readParameters/deferredInit,
//     SimulatorBase.cpp:492-634, 699-732, 860-885, and RigidBody3dBoundarySimulator::initBoundaryData :188-214);
//   * the body of SimulatorBase::timeStepNoGUI (:1142-1169) and SimulatorBase::reset (:887-934);
//   * SimulatorBase::updateBoundaryParticles (:1827-1858) and Simulation::registerNonpressureForces
//     (NonPressureForceRegistration.cpp:31-62, reduced to the two methods the DiffFR scenes select).
// The SimulatorBase object itself is never constructed: the hot path only calls its inline getters
// getBoundarySimulator()/getRigidBodyGradientManager(), so a zero-filled block with those two members set is enough.
// Only one ref context can exist at a time (the reference keeps its state in process-wide singletons).
// ============================================================================
#include <algorithm>
#include <array>
#include <cassert>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <stack>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <any>
#include <optional>
#include <variant>
#include <valarray>
#include <forward_list>
#include <initializer_list>
#include <iterator>
#include <limits>
#include <tuple>
#include <type_traits>
#include <utility>
#include <Eigen/Dense>
#include <Eigen/Sparse>

// read access to a handful of private members (per-step net Jacobians, kappa arrays, the two SimulatorBase members)
#define private public
#define protected public
#include "SPlisHSPlasH/BoundaryModel_Akinci2012.h"
#include "SPlisHSPlasH/DiffDFSPH/TimeStepDiffDFSPH.h"
#include "SPlisHSPlasH/Dynamic3dRigidBody.h"
#include "SPlisHSPlasH/Emitter.h"
#include "SPlisHSPlasH/EmitterSystem.h"
#include "SPlisHSPlasH/RigidBodyGradientManager.h"
#include "SPlisHSPlasH/Simulation.h"
#include "SPlisHSPlasH/SurfaceTension/SurfaceTension_Akinci2013.h"
#include "SPlisHSPlasH/TimeManager.h"
#include "SPlisHSPlasH/Viscosity/Viscosity_Standard.h"
#include "Simulator/RigidBody3dBoundarySimulator.h"
#include "Simulator/SceneConfiguration.h"
#include "Simulator/SimulatorBase.h"
#include "Utilities/Counting.h"
#include "Utilities/Logger.h"
#include "Utilities/Timing.h"
#undef private
#undef protected

#include "../../include/dfr.h"

using namespace SPH;
using namespace Utilities;

INIT_LOGGING
INIT_TIMING
INIT_COUNTING

// ---- the few Simulator-layer functions the compiled reference code links against -------------------------------
// NonPressureForceRegistration.cpp:31-62, reduced: index 0 "None", surface tension index 2 = Akinci et al. 2013,
// viscosity index 1 = Standard (the ids the scene files use); the other slots stay empty.
void Simulation::registerNonpressureForces() {
  auto none = [](FluidModel *) -> NonPressureForceBase * { return nullptr; };
  addDragMethod("None", none);
  addElasticityMethod("None", none);
  addSurfaceTensionMethod("None", none);
  addSurfaceTensionMethod("Becker & Teschner 2007 (not built)", none);
  addSurfaceTensionMethod("Akinci et al. 2013", SurfaceTension_Akinci2013::creator);
  addViscosityMethod("None", none);
  addViscosityMethod("Standard", Viscosity_Standard::creator);
  addVorticityMethod("None", none);
}

// SimulatorBase.cpp:1827-1858
void SimulatorBase::updateBoundaryParticles(const bool forceUpdate) {
  Simulation *sim = Simulation::getCurrent();
  for (unsigned int i = 0; i < sim->numberOfBoundaryModels(); i++) {
    BoundaryModel_Akinci2012 *bm = static_cast<BoundaryModel_Akinci2012 *>(sim->getBoundaryModel(i));
    RigidBodyObject *rbo = bm->getRigidBodyObject();
    if (!(rbo->isDynamic() || rbo->isAnimated() || forceUpdate)) continue;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < (int)bm->numberOfParticles(); j++) {
      bm->getPosition(j) = rbo->getRotation() * bm->getPosition0(j) + rbo->getPosition();
      if (rbo->isDynamic() || rbo->isAnimated())
        bm->getVelocity(j) = rbo->getAngularVelocity().cross(bm->getPosition(j) - rbo->getPosition()) + rbo->getVelocity();
      else
        bm->getVelocity(j).setZero();
    }
  }
}

struct dfr_context {
  dfr_config cfg;
  std::string err;
  bool finalized = false;
  // scene description until finalize
  std::vector<Vector3r> fx, fv;
  struct BodyDesc {
    std::vector<Vector3r> x_local;
    bool dynamic;
    double density;
    Vector3r pos;
    Quaternionr q;
    Vector3r init_v = Vector3r::Zero(), init_w = Vector3r::Zero();
  };
  std::vector<BodyDesc> bodies;
  struct EmitterDesc {
    int width, height;
    Vector3r pos;
    Matrix3r rot;
    double velocity, start, end;
  };
  std::vector<EmitterDesc> emitters;
  // loaded state re-applied by every reset (SimulatorBase::checkLoadState)
  bool has_state = false;
  std::vector<Vector3r> sx, sv;
  std::vector<double> skappa, skappav;
  bool has_sv = false, has_sk = false, has_skv = false;
  // reference objects
  SimulatorBase *base = nullptr;  // raw block, see header comment
  RigidBody3dBoundarySimulator *bsim = nullptr;
  long long total_iter = 0, total_iter_v = 0, total_psteps = 0, total_nbrs = 0;
  double cpu_ms = 0.0;
};

static dfr_context *g_live = nullptr;

static int fail(dfr_context *c, int code, const char *msg) {
  if (c) c->err = msg;
  return code;
}
static TimeStepDiffDFSPH *ts() { return static_cast<TimeStepDiffDFSPH *>(Simulation::getCurrent()->getTimeStep()); }
static BoundaryModel_Akinci2012 *bm_of(int b) { return static_cast<BoundaryModel_Akinci2012 *>(Simulation::getCurrent()->getBoundaryModel(b)); }

static void destroy_singletons(dfr_context *c) {
  if (Simulation::hasCurrent()) {
    // rigid body objects are owned by nobody in the reference; leak them (test infrastructure)
    delete Simulation::getCurrent();
  }
  if (SceneConfiguration::hasCurrent()) {
    for (auto *b : SceneConfiguration::getCurrent()->getScene().boundaryModels) delete b;
    SceneConfiguration::getCurrent()->getScene().boundaryModels.clear();
    delete SceneConfiguration::getCurrent();
  }
  if (c && c->base) {
    c->base->m_rigidBodyGradientManager.reset();
    std::free(c->base);
    c->base = nullptr;
  }
  if (c) c->bsim = nullptr;
}

static void apply_state(dfr_context *c) {  // checkLoadState -> loadFluidParticlePositions[AndVelocities] (:2023-2058, 2576-2604)
  if (!c->has_state) return;
  FluidModel *model = Simulation::getCurrent()->getFluidModel(0);
  const unsigned int n = (unsigned int)c->sx.size();
  for (unsigned int i = 0; i < n; i++) {
    model->getPosition(i) = c->sx[i];
    if (c->has_sv) model->getVelocity(i) = c->sv[i];
    model->getParticleId(i) = i;
    if (c->has_sk) ts()->m_simulationData.getKappa(0, i) = c->skappa[i];
    if (c->has_skv) ts()->m_simulationData.getKappaV(0, i) = c->skappav[i];
  }
}

template <class M>
static void put(double *out, const M &m) {
  for (int a = 0; a < m.rows(); a++)
    for (int b = 0; b < m.cols(); b++) out[a * m.cols() + b] = m(a, b);
}

extern "C" {

int ref_reset(dfr_context *c);

void ref_default_config(dfr_config *cfg) {
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->particle_radius = 0.025;
  cfg->density0 = 1000.0;
  cfg->gravitation[1] = -9.81;
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
  cfg->surface_tension_method = 0;
  cfg->surface_tension = 0.05;
  cfg->gradient_mode = 1;
  cfg->rigid_body_mode = 0;
  cfg->optimize_rotation = 1;
  cfg->rigid_contact_beta = 1.0;
  cfg->rigid_contact_gamma = 0.7;
  cfg->rigid_contact_support_radius_factor = 4.0;
  cfg->target_time = 0.8;
}

int ref_create(const dfr_config *cfg, int, dfr_context **out) {
  if (!cfg || !out) return DFR_ERR_INVALID;
  if (g_live) return DFR_ERR_STATE;  // singletons: one context at a time
  dfr_context *c = new dfr_context();
  c->cfg = *cfg;
  g_live = c;
  *out = c;
  return DFR_OK;
}

void ref_destroy(dfr_context *c) {
  if (!c) return;
  destroy_singletons(c);
  if (g_live == c) g_live = nullptr;
  delete c;
}

const char *ref_last_error(const dfr_context *c) { return c ? c->err.c_str() : "null context (or another ref context is alive)"; }

int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int ref_set_fluid(dfr_context *c, int64_t n, const double *x, const double *v) {
  if (!c || c->finalized) return fail(c, DFR_ERR_STATE, "set_fluid after finalize");
  c->fx.resize(n);
  c->fv.assign(n, Vector3r::Zero());
  for (int64_t i = 0; i < n; i++) {
    c->fx[i] = Vector3r(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
    if (v) c->fv[i] = Vector3r(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
  }
  return DFR_OK;
}

int ref_add_body(dfr_context *c, int64_t n, const double *x_local, int is_dynamic, double density, const double position[3],
                 const double quat_wxyz[4]) {
  if (!c || c->finalized) return fail(c, DFR_ERR_STATE, "add_body after finalize");
  dfr_context::BodyDesc b;
  b.x_local.resize(n);
  for (int64_t i = 0; i < n; i++) b.x_local[i] = Vector3r(x_local[3 * i], x_local[3 * i + 1], x_local[3 * i + 2]);
  b.dynamic = is_dynamic != 0;
  b.density = density;
  b.pos = Vector3r(position[0], position[1], position[2]);
  b.q = Quaternionr(quat_wxyz[0], quat_wxyz[1], quat_wxyz[2], quat_wxyz[3]);
  c->bodies.push_back(b);
  return (int)c->bodies.size() - 1;
}

int ref_set_init_v_omega(dfr_context *c, int body, const double v0[3], const double omega0[3]) {
  if (!c || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  c->bodies[body].init_v = Vector3r(v0[0], v0[1], v0[2]);
  c->bodies[body].init_w = Vector3r(omega0[0], omega0[1], omega0[2]);
  if (c->finalized) {  // TimeStepDiffDFSPH::set_init_v_rb / set_init_omega_rb (DiffDFSPHModule.cpp:60-75)
    ts()->set_init_v_rb(body, c->bodies[body].init_v);
    ts()->set_init_omega_rb(body, c->bodies[body].init_w);
  }
  return DFR_OK;
}

int ref_add_emitter(dfr_context *c, int width, int height, const double position[3], const double rot[9], double velocity,
                    double emit_start, double emit_end) {
  if (!c || c->finalized) return fail(c, DFR_ERR_STATE, "add_emitter after finalize");
  dfr_context::EmitterDesc e;
  e.width = width;
  e.height = height;
  e.pos = Vector3r(position[0], position[1], position[2]);
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) e.rot(a, b) = rot[3 * a + b];
  e.velocity = velocity;
  e.start = emit_start;
  e.end = emit_end;
  c->emitters.push_back(e);
  return DFR_OK;
}

int ref_finalize(dfr_context *c) {
  if (!c || c->finalized) return fail(c, DFR_ERR_STATE, "already finalized");
  const dfr_config &cfg = c->cfg;
  // ---- scene description (what SceneLoader::readScene would have filled; SceneLoader.cpp:41-183) ----
  SceneLoader::Scene &scene = SceneConfiguration::getCurrent()->getScene();
  scene.particleRadius = cfg.particle_radius;
  scene.sim2D = false;
  scene.timeStepSize = cfg.time_step_size;
  scene.useRigidContactSolver = cfg.use_rigid_contact_solver != 0;
  scene.useRigidGradientManager = cfg.use_rigid_gradient_manager != 0;
  scene.useReleaseRigidBodyMode = cfg.use_release_rigid_body_mode != 0;
  scene.rigidContactGamma = cfg.rigid_contact_gamma;
  scene.rigidContactFrictionCoeff = cfg.rigid_contact_friction;
  scene.rigidContactBeta = cfg.rigid_contact_beta;
  scene.rigidContactSupportRadiusFactor = cfg.rigid_contact_support_radius_factor;
  scene.targetTime = cfg.target_time;
  scene.uniformAccelerateRBTime = cfg.uniform_acc_rb_time;
  scene.gradientMode = (unsigned int)cfg.gradient_mode;
  for (auto &b : c->bodies) {
    auto *bd = new SceneLoader::BoundaryData();
    bd->translation = b.pos;
    bd->rotation = b.q.toRotationMatrix();
    bd->scale = Vector3r::Ones();
    bd->density = b.density;
    bd->dynamic = b.dynamic;
    bd->isWall = false;
    bd->isAnimated = false;
    bd->target_x = Vector3r::Zero();
    bd->target_angle_in_degree = Vector3r::Zero();
    bd->init_velocity = b.init_v;
    bd->init_angular_velocity = b.init_w;
    scene.boundaryModels.push_back(bd);
  }
  // ---- SimulatorBase::initSimulation (:492-587) ----
  c->base = static_cast<SimulatorBase *>(std::calloc(1, sizeof(SimulatorBase)));
  new (&c->base->m_rigidBodyGradientManager) std::unique_ptr<RigidBodyGradientManager>(new RigidBodyGradientManager());
  c->bsim = new RigidBody3dBoundarySimulator(c->base);
  c->base->m_boundarySimulator = c->bsim;
  Simulation *sim = Simulation::getCurrentWithBase(c->base);
  sim->init(cfg.particle_radius, false);
  // buildModel (:860-885)
  TimeManager::getCurrent()->setTimeStepSize(cfg.time_step_size);
  {
    std::vector<unsigned int> objIds(c->fx.size(), 0u);
    static Vector3r dummy = Vector3r::Zero();
    sim->addFluidModel("Fluid", (unsigned int)c->fx.size(), c->fx.empty() ? &dummy : c->fx.data(), c->fv.empty() ? &dummy : c->fv.data(),
                       objIds.empty() ? nullptr : objIds.data(), (unsigned int)std::max(0, cfg.max_emitted_particles));
  }
  FluidModel *model = sim->getFluidModel(0);
  for (auto &e : c->emitters) {  // SimulatorBase::createEmitters (:1737-1789): box emitter, type 0
    model->getEmitterSystem()->addEmitter((unsigned int)e.width, (unsigned int)e.height, e.pos, e.rot, e.velocity, 0);
    Emitter *em = model->getEmitterSystem()->getEmitters().back();
    em->setEmitStartTime(e.start);
    em->setEmitEndTime(e.end);
  }
  // readParameters (:699-732): "Configuration" keys of the scene file
  sim->setBoundaryHandlingMethod(BoundaryHandlingMethods::Akinci2012);
  sim->setValue<int>(Simulation::SIMULATION_METHOD, (int)SimulationMethods::DiffDFSPH);
  Real g[3] = {cfg.gravitation[0], cfg.gravitation[1], cfg.gravitation[2]};
  sim->setVecValue<Real>(Simulation::GRAVITATION, g);
  sim->setValue<int>(Simulation::CFL_METHOD, cfg.cfl_method);
  sim->setValue<Real>(Simulation::CFL_FACTOR, cfg.cfl_factor);
  sim->setValue<Real>(Simulation::CFL_MIN_TIMESTEPSIZE, cfg.cfl_min_time_step);
  sim->setValue<Real>(Simulation::CFL_MAX_TIMESTEPSIZE, cfg.cfl_max_time_step);
  sim->setGradientMode(cfg.gradient_mode);
  sim->setRigidBodyMode(cfg.rigid_body_mode);
  TimeStepDiffDFSPH *t = ts();
  t->setValue<unsigned int>(TimeStep::MIN_ITERATIONS, (unsigned int)cfg.min_iterations);
  t->setValue<unsigned int>(TimeStep::MAX_ITERATIONS, (unsigned int)cfg.max_iterations);
  t->setValue<Real>(TimeStep::MAX_ERROR, cfg.max_error);
  t->setValue<unsigned int>(TimeStepDiffDFSPH::MAX_ITERATIONS_V, (unsigned int)cfg.max_iterations_v);
  t->setValue<Real>(TimeStepDiffDFSPH::MAX_ERROR_V, cfg.max_error_v);
  t->setValue<bool>(TimeStepDiffDFSPH::USE_DIVERGENCE_SOLVER, cfg.enable_divergence_solver != 0);
  t->setValue<bool>(TimeStepDiffDFSPH::USE_PRESSURE_WARMSTART, cfg.use_pressure_warmstart != 0);
  t->setValue<bool>(TimeStepDiffDFSPH::USE_DIV_WARMSTART, cfg.use_divergence_warmstart != 0);
  t->setValue<bool>(TimeStepDiffDFSPH::OPTIMIZE_ROTATION, cfg.optimize_rotation != 0);
  // "Materials"
  model->setValue<Real>(FluidModel::DENSITY0, cfg.density0);
  model->setValue<int>(FluidModel::SURFACE_TENSION_METHOD, cfg.surface_tension_method);
  model->setValue<int>(FluidModel::VISCOSITY_METHOD, cfg.viscosity_method);
  if (model->getSurfaceTensionBase()) {
    model->getSurfaceTensionBase()->setValue<Real>(SurfaceTensionBase::SURFACE_TENSION, cfg.surface_tension);
    model->getSurfaceTensionBase()->setValue<Real>(SurfaceTensionBase::SURFACE_TENSION_BOUNDARY, cfg.surface_tension_boundary);
  }
  if (model->getViscosityBase()) {
    model->getViscosityBase()->setValue<Real>(ViscosityBase::VISCOSITY_COEFFICIENT, cfg.viscosity);
    model->getViscosityBase()->setValue<Real>(Viscosity_Standard::VISCOSITY_COEFFICIENT_BOUNDARY, cfg.viscosity_boundary);
  }
  // ---- SimulatorBase::deferredInit (:590-647): initBoundaryData with the samples given instead of sampled ----
  for (auto &b : c->bodies) {  // RigidBody3dBoundarySimulator.cpp:98-101, 188-214
    Dynamic3dRigidBody *rb = new Dynamic3dRigidBody();
    rb->setIsAnimated(false);
    rb->setIsDynamic(b.dynamic);
    rb->setPosition0(b.pos);
    rb->setPosition(b.pos);
    rb->setRotation0(b.q);
    rb->setRotation(b.q);
    rb->determineMassProperties(b.density, cfg.particle_radius, b.x_local);
    BoundaryModel_Akinci2012 *bm = new BoundaryModel_Akinci2012();
    bm->initModel(rb, (unsigned int)b.x_local.size(), b.x_local.data());
    sim->addBoundaryModel(bm);
    rb->updateMeshTransformation();
  }
  sim->setSimulationInitialized(true);
  c->bsim->deferredInit();
  c->base->m_rigidBodyGradientManager->deferredInit();
  c->finalized = true;
  return DFR_OK;
}

int ref_load_fluid_state(dfr_context *c, const double *x, const double *v, const double *kappa, const double *kappa_v) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  const size_t n = c->fx.size();
  c->has_state = true;
  if (c->sx.empty()) c->sx = c->fx;
  if (x)
    for (size_t i = 0; i < n; i++) c->sx[i] = Vector3r(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
  if (v) {
    c->sv.resize(n);
    for (size_t i = 0; i < n; i++) c->sv[i] = Vector3r(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    c->has_sv = true;
  }
  if (kappa) {
    c->skappa.assign(kappa, kappa + n);
    c->has_sk = true;
  }
  if (kappa_v) {
    c->skappav.assign(kappa_v, kappa_v + n);
    c->has_skv = true;
  }
  return ref_reset(c);
}

int ref_reset(dfr_context *c) {  // SimulatorBase::reset (:887-934)
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  Utilities::Timing::reset();
  Utilities::Counting::reset();
  Simulation::getCurrent()->reset();
  c->bsim->reset();
  c->base->m_rigidBodyGradientManager->reset();
  apply_state(c);
  if (Simulation::getCurrent()->getValue<int>(Simulation::CFL_METHOD) != Simulation::ENUM_CFL_NONE)
    TimeManager::getCurrent()->setTimeStepSize(c->cfg.time_step_size);
  c->total_iter = c->total_iter_v = c->total_psteps = c->total_nbrs = 0;
  c->cpu_ms = 0.0;
  return DFR_OK;
}

static void one_step(dfr_context *c) {  // SimulatorBase::timeStepNoGUI (:1142-1169)
  Simulation *sim = Simulation::getCurrent();
  sim->getTimeStep()->step();
  if (sim->useRigidGradientManager() && sim->numberOfFluidModels() > 0) c->base->m_rigidBodyGradientManager->after_Fluid_Rigid_coupling_step();
  c->bsim->velocityTimeStep();
  if (sim->useRigidGradientManager()) c->base->m_rigidBodyGradientManager->after_Rigid_Rigid_coupling_step();
  c->bsim->positionTimeStep();
  c->total_iter += ts()->getValue<unsigned int>(TimeStep::SOLVER_ITERATIONS);
  c->total_iter_v += ts()->getValue<unsigned int>(TimeStepDiffDFSPH::SOLVER_ITERATIONS_V);
  c->total_psteps += sim->getFluidModel(0)->numActiveParticles();
}

int ref_reset_gradient(dfr_context *c) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  ts()->reset_gradient();  // TimeStepDiffDFSPH.cpp:2234-2240
  return DFR_OK;
}

int ref_set_gradient_mode(dfr_context *c, int mode) {
  if (!c) return DFR_ERR_INVALID;
  Simulation::getCurrent()->setGradientMode(mode);
  c->cfg.gradient_mode = mode;
  return DFR_OK;
}

int ref_step(dfr_context *c, int n_steps) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  const auto t0 = std::chrono::steady_clock::now();
  for (int s = 0; s < n_steps; s++) one_step(c);
  c->cpu_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return DFR_OK;
}

int ref_run_trajectory(dfr_context *c, int max_steps, int *steps_done) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  int s = 0;
  while (s < max_steps) {
    one_step(c);
    s++;
    if (ts()->is_trajectory_finish_callback()) break;
  }
  if (steps_done) *steps_done = s;
  return DFR_OK;
}

int ref_get_step_info(dfr_context *c, dfr_step_info *info) {
  if (!c || !info || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  info->time = TimeManager::getCurrent()->getTime();
  info->time_step_size = TimeManager::getCurrent()->getTimeStepSize();
  info->iterations = (int)ts()->getValue<unsigned int>(TimeStep::SOLVER_ITERATIONS);
  info->iterations_v = (int)ts()->getValue<unsigned int>(TimeStepDiffDFSPH::SOLVER_ITERATIONS_V);
  info->step_count = (int)ts()->get_step_count();
  info->trajectory_finished = ts()->is_trajectory_finish_callback() ? 1 : 0;
  info->num_fluid_particles = Simulation::getCurrent()->getFluidModel(0)->numActiveParticles();
  info->total_pressure_iterations = c->total_iter;
  info->total_divergence_iterations = c->total_iter_v;
  info->total_particle_steps = c->total_psteps;
  info->total_fluid_neighbors = c->total_nbrs;
  return DFR_OK;
}

int ref_get_body_state(dfr_context *c, int body, double out[13]) {
  if (!c || !c->finalized || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  BoundaryModel_Akinci2012 *bm = bm_of(body);
  const Vector3r x = bm->get_position_rb(), v = bm->get_velocity_rb(), w = bm->get_angular_velocity_rb();
  const Vector4r q = bm->get_quaternion_rb_vec4();
  for (int k = 0; k < 3; k++) {
    out[k] = x[k];
    out[7 + k] = v[k];
    out[10 + k] = w[k];
  }
  for (int k = 0; k < 4; k++) out[3 + k] = q[k];
  return DFR_OK;
}

int ref_set_body_velocity(dfr_context *c, int body, const double v[3], const double omega[3]) {
  if (!c || !c->finalized || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  if (v) bm_of(body)->set_velocity_rb(Vector3r(v[0], v[1], v[2]));
  if (omega) bm_of(body)->set_angular_velocity_rb(Vector3r(omega[0], omega[1], omega[2]));
  return DFR_OK;
}

int ref_get_body_properties(dfr_context *c, int body, double out[17]) {
  if (!c || !c->finalized || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  BoundaryModel_Akinci2012 *bm = bm_of(body);
  Dynamic3dRigidBody *rb = static_cast<Dynamic3dRigidBody *>(bm->getRigidBodyObject());
  out[0] = rb->getMass();
  out[1] = rb->getInvMass();
  const Matrix3r I0 = rb->getInertiaTensor0();
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) out[2 + 3 * a + b] = I0(a, b);
  Vector3r f = Vector3r::Zero(), t = Vector3r::Zero();
  if (rb->isDynamic()) {
    f = bm->getForce();
    t = bm->getTorque();
  }
  for (int k = 0; k < 3; k++) {
    out[11 + k] = f[k];
    out[14 + k] = t[k];
  }
  return DFR_OK;
}

int ref_get_body_grad(dfr_context *c, int body, int which, double out[12]) {
  if (!c || !c->finalized || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  BoundaryModel_Akinci2012 *bm = bm_of(body);
  std::memset(out, 0, 12 * sizeof(double));
  switch (which) {
    case 0: put(out, bm->get_grad_x_to_v0()); break;
    case 1: put(out, bm->get_grad_x_to_omega0()); break;
    case 2: put(out, bm->get_grad_quaternion_to_v0()); break;
    case 3: put(out, bm->get_grad_quaternion_to_omega0()); break;
    case 4: put(out, bm->get_grad_v_to_v0()); break;
    case 5: put(out, bm->get_grad_v_to_omega0()); break;
    case 6: put(out, bm->get_grad_omega_to_v0()); break;
    case 7: put(out, bm->get_grad_omega_to_omega0()); break;
    case 8: put(out, bm->grad_net_force_to_vn); break;
    case 9: put(out, bm->grad_net_force_to_xn); break;
    case 10: put(out, bm->grad_net_force_to_qn); break;
    case 11: put(out, bm->grad_net_force_to_omega_n); break;
    case 12: put(out, bm->grad_net_torque_to_vn); break;
    case 13: put(out, bm->grad_net_torque_to_xn); break;
    case 14: put(out, bm->grad_net_torque_to_qn); break;
    case 15: put(out, bm->grad_net_torque_to_omega_n); break;
    default: return fail(c, DFR_ERR_INVALID, "bad gradient selector");
  }
  return DFR_OK;
}

int ref_get_manager_grad(dfr_context *c, int R, int RR, int which, double out[12]) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  const int n = (int)c->bodies.size();
  if (R < 0 || RR < 0 || R >= n || RR >= n) return fail(c, DFR_ERR_INVALID, "bad body index");
  RigidBodyGradientManager *m = c->base->m_rigidBodyGradientManager.get();
  std::memset(out, 0, 12 * sizeof(double));
  switch (which) {
    case 0: put(out, m->get_grad_xn_to_v0(R, RR)); break;
    case 1: put(out, m->get_grad_xn_to_omega0(R, RR)); break;
    case 2: put(out, m->get_grad_qn_to_v0(R, RR)); break;
    case 3: put(out, m->get_grad_qn_to_omega0(R, RR)); break;
    case 4: put(out, m->get_grad_vn_to_v0(R, RR)); break;
    case 5: put(out, m->get_grad_vn_to_omega0(R, RR)); break;
    case 6: put(out, m->get_grad_omega_n_to_v0(R, RR)); break;
    case 7: put(out, m->get_grad_omega_n_to_omega0(R, RR)); break;
    case 8: put(out, m->get_grad_net_force_to_vn(R, RR)); break;
    case 9: put(out, m->get_grad_net_force_to_xn(R, RR)); break;
    case 10: put(out, m->get_grad_net_force_to_qn(R, RR)); break;
    case 11: put(out, m->get_grad_net_force_to_omega_n(R, RR)); break;
    case 12: put(out, m->get_grad_net_torque_to_vn(R, RR)); break;
    case 13: put(out, m->get_grad_net_torque_to_xn(R, RR)); break;
    case 14: put(out, m->get_grad_net_torque_to_qn(R, RR)); break;
    case 15: put(out, m->get_grad_net_torque_to_omega_n(R, RR)); break;
    default: return fail(c, DFR_ERR_INVALID, "bad gradient selector");
  }
  return DFR_OK;
}

int64_t ref_num_fluid(dfr_context *c) {
  if (!c) return 0;
  if (!c->finalized) return (int64_t)c->fx.size();
  return Simulation::getCurrent()->getFluidModel(0)->numActiveParticles();
}
int64_t ref_num_fluid_initial(dfr_context *c) { return c ? (int64_t)c->fx.size() : 0; }
int64_t ref_num_body_particles(dfr_context *c, int body) {
  return (c && body >= 0 && body < (int)c->bodies.size()) ? (int64_t)c->bodies[body].x_local.size() : 0;
}
int ref_num_bodies(dfr_context *c) { return c ? (int)c->bodies.size() : 0; }

int ref_download_fluid(dfr_context *c, int field, double *out) {
  if (!c || !c->finalized || !out) return fail(c, DFR_ERR_STATE, "not finalized");
  FluidModel *model = Simulation::getCurrent()->getFluidModel(0);
  auto &sd = ts()->m_simulationData;
  const unsigned int n = model->numActiveParticles();
  SurfaceTension_Akinci2013 *st = dynamic_cast<SurfaceTension_Akinci2013 *>(model->getSurfaceTensionBase());
  for (unsigned int i = 0; i < n; i++) {
    const unsigned int id = model->getParticleId(i);
    Vector3r v3 = Vector3r::Zero();
    double s = 0.0;
    bool vec = true;
    switch (field) {
      case 0: v3 = model->getPosition(i); break;
      case 1: v3 = model->getVelocity(i); break;
      case 2: s = model->getDensity(i); vec = false; break;
      case 3: s = sd.getFactor(0, i); vec = false; break;
      case 4: s = sd.getKappa(0, i); vec = false; break;
      case 5: s = sd.getKappaV(0, i); vec = false; break;
      case 6: s = sd.getDensityAdv(0, i); vec = false; break;
      case 7: v3 = model->getAcceleration(i); break;
      case 8: v3 = sd.get_sum_grad_p_k(0, i); break;
      case 9: if (st) v3 = st->getNormal(i); break;
      default: return fail(c, DFR_ERR_INVALID, "bad field");
    }
    if (vec) {
      out[3 * (size_t)id] = v3[0];
      out[3 * (size_t)id + 1] = v3[1];
      out[3 * (size_t)id + 2] = v3[2];
    } else
      out[id] = s;
  }
  return DFR_OK;
}

int ref_download_body(dfr_context *c, int body, int field, double *out) {
  if (!c || !c->finalized || !out || body < 0 || body >= (int)c->bodies.size()) return fail(c, DFR_ERR_INVALID, "bad body index");
  BoundaryModel_Akinci2012 *bm = bm_of(body);
  // boundary particles of dynamic bodies are z-sorted in place (BoundaryModel_Akinci2012.cpp:388-416); rows are matched
  // to the caller's sample order through the body-frame positions, which the sort permutes along
  const auto &xl = c->bodies[body].x_local;
  const unsigned int n = bm->numberOfParticles();
  std::vector<int> row_of(n, -1);
  {
    std::map<std::array<double, 3>, int> lut;
    for (unsigned int j = 0; j < n; j++) lut[{xl[j][0], xl[j][1], xl[j][2]}] = (int)j;
    for (unsigned int r = 0; r < n; r++) {
      const Vector3r &p0 = bm->getPosition0(r);
      auto it = lut.find({p0[0], p0[1], p0[2]});
      if (it == lut.end()) return fail(c, DFR_ERR_STATE, "boundary sample lookup failed");
      row_of[r] = it->second;
    }
  }
  for (unsigned int r = 0; r < n; r++) {
    const int j = row_of[r];
    Vector3r v3;
    switch (field) {
      case 0: v3 = bm->getPosition(r); break;
      case 1: v3 = bm->getVelocity(r); break;
      case 2: out[j] = bm->getVolume(r); continue;
      case 3: v3 = bm->getPosition0(r); break;
      default: return fail(c, DFR_ERR_INVALID, "bad field");
    }
    out[3 * j] = v3[0];
    out[3 * j + 1] = v3[1];
    out[3 * j + 2] = v3[2];
  }
  return DFR_OK;
}

int ref_get_neighbors(dfr_context *c, int set_a, int set_b, int32_t *counts, int32_t *indices, int64_t capacity, int64_t *total) {
  if (!c || !c->finalized) return fail(c, DFR_ERR_STATE, "not finalized");
  const int nb = (int)c->bodies.size();
  if (set_a < -1 || set_a >= nb || set_b < -1 || set_b >= nb) return fail(c, DFR_ERR_INVALID, "bad set index");
  Simulation *sim = Simulation::getCurrent();
  sim->performNeighborhoodSearch();  // Simulation.cpp:743-750 (the step's own search, on the current positions)
  FluidModel *model = sim->getFluidModel(0);
  auto pset = [&](int s) { return s < 0 ? model->getPointSetIndex() : bm_of(s)->getPointSetIndex(); };
  auto id_of = [&](int s, unsigned int row, std::vector<int> *lut) -> int { return s < 0 ? (int)model->getParticleId(row) : (*lut)[row]; };
  // body rows -> caller order
  auto body_lut = [&](int b) {
    std::vector<int> lut;
    if (b < 0) return lut;
    BoundaryModel_Akinci2012 *bm = bm_of(b);
    const auto &xl = c->bodies[b].x_local;
    std::map<std::array<double, 3>, int> m;
    for (size_t j = 0; j < xl.size(); j++) m[{xl[j][0], xl[j][1], xl[j][2]}] = (int)j;
    lut.resize(bm->numberOfParticles());
    for (unsigned int r = 0; r < bm->numberOfParticles(); r++) {
      const Vector3r &p0 = bm->getPosition0(r);
      lut[r] = m[{p0[0], p0[1], p0[2]}];
    }
    return lut;
  };
  std::vector<int> la = body_lut(set_a), lb = body_lut(set_b);
  const unsigned int na = set_a < 0 ? model->numActiveParticles() : bm_of(set_a)->numberOfParticles();
  std::vector<std::vector<int32_t>> rows(na);
  const unsigned int pa = pset(set_a), pb = pset(set_b);
  for (unsigned int r = 0; r < na; r++) {
    const int ia = id_of(set_a, r, &la);
    const unsigned int cnt = sim->numberOfNeighbors(pa, pb, r);
    for (unsigned int k = 0; k < cnt; k++) rows[ia].push_back(id_of(set_b, sim->getNeighbor(pa, pb, r, k), &lb));
    std::sort(rows[ia].begin(), rows[ia].end());
  }
  int64_t tot = 0;
  for (size_t i = 0; i < rows.size(); i++) {
    if (counts) counts[i] = (int32_t)rows[i].size();
    if (indices) {
      if (tot + (int64_t)rows[i].size() > capacity) return fail(c, DFR_ERR_CAPACITY, "neighbour buffer too small");
      std::memcpy(indices + tot, rows[i].data(), rows[i].size() * sizeof(int32_t));
    }
    tot += (int64_t)rows[i].size();
  }
  if (total) *total = tot;
  return DFR_OK;
}

int ref_get_device_time_ms(dfr_context *c, double *total_ms, int64_t *kernel_launches) {
  if (!c) return DFR_ERR_INVALID;
  if (total_ms) *total_ms = c->cpu_ms;
  if (kernel_launches) *kernel_launches = 0;
  return DFR_OK;
}

}  // extern "C"
