// Host side of the drop-in boundary: C++ mirrors of the reference classes the optimisation scripts drive
// (SimulatorBase, Simulation, TimeManager, TimeStepDiffDFSPH, BoundaryModel_Akinci2012, RigidBodyObject,
// RigidBodyGradientManager), implemented on top of the C ABI in include/dfr.h.  No CUDA and no Python here:
// bindings.cpp exposes these classes through pybind11 under the reference's names.
//
// Ownership mirrors the reference: SimulatorBase owns everything; Simulation::getCurrent() /
// TimeManager::getCurrent() hand out non-owning pointers to the objects of the most recently initialised
// simulator (Simulation.cpp:167-188, TimeManager.cpp:5-25).  Underneath, each SimulatorBase has its own
// dfr_context, so several simulators can coexist in one process (the reference cannot do that).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <sys/stat.h>

#include "scene_host.hpp"
#include "state_io.hpp"

namespace dfrhost {

class SimulatorBase;

inline void mkdir_p(const std::string &path) {
  std::string cur;
  for (size_t i = 0; i <= path.size(); i++) {
    if (i == path.size() || path[i] == '/') {
      if (!cur.empty()) ::mkdir(cur.c_str(), 0755);
    }
    if (i < path.size()) cur += path[i];
  }
}

struct DfrFailure : std::runtime_error {
  int code;
  DfrFailure(int c, const std::string &m) : std::runtime_error("dfr error " + std::to_string(c) + ": " + m), code(c) {}
};

// ---- RigidBodyObject (pySPlisHSPlasH/RigidBodyModule.cpp:18-45) ------------------------------------------------
class RigidBodyObject {
 public:
  RigidBodyObject(SimulatorBase *b, int i) : base(b), index(i) {}
  bool isDynamic() const;
  double getMass() const;
  Vec3 getPosition() const;
  Vec3 getVelocity() const;
  Vec3 getAngularVelocity() const;
  Quat getRotationXYZW() const;  // Eigen coeffs() order, as the reference binding returns it
  void setVelocity(const Vec3 &v);
  void setAngularVelocity(const Vec3 &w);

 private:
  SimulatorBase *base;
  int index;
};

// ---- BoundaryModel_Akinci2012 (pySPlisHSPlasH/BoundaryModelModule.cpp:23-91) ----------------------------------
class BoundaryModelAkinci2012 {
 public:
  BoundaryModelAkinci2012(SimulatorBase *b, int i) : base(b), index(i), rbo(b, i) {}
  unsigned int numberOfParticles() const;
  Vec3 get_position_rb() const;
  Quat get_quaternion_rb_vec4() const;  // (w, x, y, z)
  Vec3 get_velocity_rb() const;
  Vec3 get_angular_velocity_rb() const;
  void set_velocity_rb(const Vec3 &v);
  void set_angular_velocity_rb(const Vec3 &w);
  // which: numbering of dfr_get_body_grad
  std::array<double, 12> grad(int which) const;
  Vec3 getForce() const;
  Vec3 getTorque() const;
  Vec3 particle(int field, unsigned int i) const;  // 0 position, 1 velocity, 3 position0
  double getVolume(unsigned int i) const;
  RigidBodyObject *getRigidBodyObject() { return &rbo; }
  int bodyIndex() const { return index; }

 private:
  SimulatorBase *base;
  int index;
  RigidBodyObject rbo;
  mutable std::vector<double> cache[4];
  mutable int cache_step[4] = {-2, -2, -2, -2};
  const std::vector<double> &field(int f) const;
};

// ---- RigidBodyGradientManager (pySPlisHSPlasH/SimulationModule.cpp:406-414) -----------------------------------
class RigidBodyGradientManager {
 public:
  explicit RigidBodyGradientManager(SimulatorBase *b) : base(b) {}
  void reset();
  std::array<double, 12> grad(int R, int RR, int which) const;

 private:
  SimulatorBase *base;
};

// ---- TimeStepDiffDFSPH (pySPlisHSPlasH/DiffDFSPHModule.cpp:44-102) --------------------------------------------
class TimeStepDiffDFSPH {
 public:
  explicit TimeStepDiffDFSPH(SimulatorBase *b) : base(b) {}
  BoundaryModelAkinci2012 *get_boundary_model(unsigned int i);
  double loss = 0.0, loss_x = 0.0, loss_rotation = 0.0, lr = 0.0;
  void set_init_v_rb(unsigned int i, const Vec3 &v);
  void set_init_omega_rb(unsigned int i, const Vec3 &w);
  Vec3 get_init_v_rb(unsigned int i) const;
  Vec3 get_init_omega_rb(unsigned int i) const;
  Vec3 get_target_x(unsigned int i) const;
  void set_target_x(unsigned int i, const Vec3 &x);
  Vec3 get_target_angle_in_radian(unsigned int i) const;
  Quat get_target_quaternion_vec4(unsigned int i) const;
  bool is_trajectory_finish_callback() const { return trajectory_finished_cb; }
  void clear_all_callbacks() { trajectory_finished_cb = false; }
  bool is_in_new_trajectory() const { return in_new_trajectory; }
  void set_custom_log_message(const std::string &s) { custom_log += s; }
  std::string get_custom_log_message() const { return custom_log; }
  unsigned int get_step_count() const;
  void add_log(const std::string &s);
  void reset_gradient();
  unsigned int get_num_1ring_fluid_particle() const;
  unsigned int getIterations() const;   // TimeStep::SOLVER_ITERATIONS
  unsigned int getIterationsV() const;  // TimeStepDiffDFSPH::SOLVER_ITERATIONS_V

  bool trajectory_finished_cb = false, in_new_trajectory = false;
  std::string custom_log;

 private:
  SimulatorBase *base;
};

// ---- TimeManager (pySPlisHSPlasH/TimeModule.cpp:23-29) ----------------------------------------------------------
class TimeManager {
 public:
  explicit TimeManager(SimulatorBase *b) : base(b) {}
  double getTime() const;
  double getTimeStepSize() const;
  static TimeManager *current;

 private:
  SimulatorBase *base;
};

// ---- Simulation (pySPlisHSPlasH/SimulationModule.cpp:95-211) ----------------------------------------------------
class Simulation {
 public:
  explicit Simulation(SimulatorBase *b) : base(b) {}
  TimeStepDiffDFSPH *getTimeStep();
  BoundaryModelAkinci2012 *getBoundaryModel(unsigned int i);
  unsigned int numberOfBoundaryModels() const;
  unsigned int numberOfFluidModels() const;
  unsigned int numberOfFluidParticles() const;
  void setGradientMode(int m);
  int getGradientMode() const;
  bool useRigidGradientManager() const;
  bool useRigidContactSolver() const;
  double getParticleRadius() const;
  double getSupportRadius() const;
  SimulatorBase *simulatorBase() const { return base; }
  static Simulation *current;

 private:
  SimulatorBase *base;
};

// ---- SimulatorBase (pySPlisHSPlasH/SimulationModule.cpp:236-362) ------------------------------------------------
class SimulatorBase {
 public:
  // GenParam ids (the reference assigns them at run time in initParameters, SimulatorBase.cpp:132-256; the scripts
  // only ever pass them back into setValue*/getValue*)
  enum { PAUSE = 0, PAUSE_AT = 1, STOP_AT = 2, NUM_STEPS_PER_RENDER = 3, DATA_EXPORT_FPS = 4, STATE_EXPORT = 5, STATE_EXPORT_FPS = 6 };

  SimulatorBase() : simulation(this), time_manager(this), timestep(this), grad_manager(this) {}
  ~SimulatorBase() { cleanup(); }

  void init(const std::string &sceneFile, const std::string &programName, bool useCache, const std::string &stateFile,
            bool loadFluidPos, bool loadFluidPosAndVel, const std::string &outputDir, bool initialPause, bool useGui,
            double stopAt, const std::string &param) {
    (void)programName; (void)useCache; (void)initialPause;
    scene_file = sceneFile;
    state_file = stateFile;
    load_pos = loadFluidPos;
    load_pos_vel = loadFluidPosAndVel;
    use_gui = useGui;
    values_f[STOP_AT] = stopAt;
    param_str = param;
    output_path = outputDir.empty() ? dir_of(sceneFile) + "/output" : outputDir;
  }
  void setDevice(int d) { device = d; }  // not in the reference: which GPU this simulator's context lives on

  // SimulatorBase::initSimulation (SimulatorBase.cpp:494-588): scene -> Simulation, boundary models, time step
  void initSimulation() {
    if (ctx) throw std::runtime_error("initSimulation called twice");
    scene = load_scene(scene_file, param_str);
    build_context();
    Simulation::current = &simulation;
    TimeManager::current = &time_manager;
    mkdir_p(output_path + "/log");
    log_file = std::fopen((output_path + "/log/SPH_log.txt").c_str(), "a");
  }
  // scene assembled by the caller instead of a file (tests, synthetic scenes)
  void initSimulationFromScene(const Scene &sc) {
    if (ctx) throw std::runtime_error("initSimulation called twice");
    scene = sc;
    build_context();
    Simulation::current = &simulation;
    TimeManager::current = &time_manager;
  }
  // SimulatorBase::deferredInit (SimulatorBase.cpp:590-647): boundary volumes, first neighbourhood sort
  void deferredInit() {
    if (finalized) return;
    check(dfr_finalize(ctx));
    finalized = true;
  }
  void initSimulationWithDeferredInit() {  // SimulatorBase.cpp:1072-1091
    initSimulation();
    deferredInit();
    if (!state_file.empty()) checkLoadState();
  }
  void runSimulation() {  // SimulatorBase.cpp:649-683 (no-GUI branch; the GUI is out of scope, so useGui only relaxes stopAt)
    deferredInit();
    if (!state_file.empty()) checkLoadState();
    if (!use_gui && values_f[STOP_AT] < 0.0) throw std::runtime_error("StopAt parameter must be set when starting without GUI.");
    stop_requested = false;
    while (!stop_requested) {
      if (!timeStepNoGUI()) break;
    }
  }
  void runNewTrajectory() {  // SimulatorBase.cpp:1093-1111
    timestep.in_new_trajectory = true;
    reset();
    for (;;) {
      singleTimeStep();
      if (trajectory_finished()) break;
    }
    timestep.in_new_trajectory = false;
  }
  void forwardFixedSteps(unsigned int n) {
    for (unsigned int i = 0; i < n; i++) singleTimeStep();
  }
  void singleTimeStep() { step_once(true); }  // SimulatorBase::timeStep has the same body plus rendering (:969-1050)
  bool timeStepNoGUI() {                      // SimulatorBase.cpp:1142-1200
    const double stopAt = values_f[STOP_AT];
    if (stopAt > 0.0 && stopAt < time_manager.getTime()) return false;
    step_once(true);
    return true;
  }
  void reset() {  // SimulatorBase.cpp:887-934
    deferredInit();
    check(dfr_reset(ctx));  // the loaded state (checkLoadState) is part of the snapshot dfr_reset restores
    info_valid = false;
    timestep.trajectory_finished_cb = false;
    step_serial++;
    if (reset_cb) reset_cb();
  }
  void cleanup() {
    if (ctx) dfr_destroy(ctx);
    ctx = nullptr;
    finalized = false;
    if (Simulation::current == &simulation) Simulation::current = nullptr;
    if (TimeManager::current == &time_manager) TimeManager::current = nullptr;
    if (log_file) std::fclose(log_file);
    log_file = nullptr;
  }
  void stop() { stop_requested = true; }

  void setTimeStepCB(std::function<void()> f) { step_cb = std::move(f); }
  void setTimeStepCallBefore(std::function<void()> f) { step_cb_before = std::move(f); }
  void setResetCB(std::function<void()> f) { reset_cb = std::move(f); }

  // State files.  The reference writes state_<t>.bin + one partio .bgeo per fluid model (SimulatorBase.cpp:2060-2293);
  // here one little-endian file: "DFRSTAT1", n, time, then x, v (n*3), kappa, kappa_v (n) as doubles in id order.
  std::string saveState(const std::string &dir) {
    deferredInit();
    const std::string d = dir.empty() ? output_path + "/state" : dir;
    mkdir_p(d);
    char name[64];
    std::snprintf(name, sizeof(name), "/state_%.6f.dfrs", time_manager.getTime());
    const std::string path = d + name;
    const int64_t n = dfr_num_fluid(ctx);
    std::vector<double> x(3 * n), v(3 * n), k(n), kv(n);
    if (n) {
      check(dfr_download_fluid(ctx, 0, x.data()));
      check(dfr_download_fluid(ctx, 1, v.data()));
      check(dfr_download_fluid(ctx, 4, k.data()));
      check(dfr_download_fluid(ctx, 5, kv.data()));
    }
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    const double t = time_manager.getTime();
    std::fwrite("DFRSTAT1", 1, 8, f);
    std::fwrite(&n, sizeof(n), 1, f);
    std::fwrite(&t, sizeof(t), 1, f);
    std::fwrite(x.data(), sizeof(double), x.size(), f);
    std::fwrite(v.data(), sizeof(double), v.size(), f);
    std::fwrite(k.data(), sizeof(double), k.size(), f);
    std::fwrite(kv.data(), sizeof(double), kv.size(), f);
    std::fclose(f);
    return path;
  }
  // mode: 0 full state (loadState), 1 positions only (--load-fluid-pos), 2 positions and velocities
  void loadStateFile(const std::string &path, int mode) {
    deferredInit();
    const bool own = path.size() > 5 && path.substr(path.size() - 5) == ".dfrs";
    if (!own) {  // the reference's pair state_<n>.bin + state_<n>_particle_Fluid.bgeo (or a .bgeo given directly)
      const bool direct = path.size() > 5 && path.substr(path.size() - 5) == ".bgeo";
      const std::string bgeo = direct ? path : bgeo_of_state_file(path);
      std::ifstream probe(bgeo, std::ios::binary);
      if (!probe) {
        std::fprintf(stderr, "[warn] File %s does not exist; state unchanged\n", bgeo.c_str());
        return;
      }
      probe.close();
      FluidStateFile st = read_bgeo(bgeo);
      if (st.n != dfr_num_fluid(ctx)) throw std::runtime_error(bgeo + ": particle count differs from the scene's fluid");
      // readFluidParticlesState fills every field the file has (positions, velocities, kappa, kappa_v, ...);
      // --load-fluid-pos then clears the velocities (SimulatorBase.cpp:2043-2058).  Rigid body and parameter state of
      // the .bin file are not read.
      std::vector<double> zero;
      const double *v = st.v.empty() ? nullptr : st.v.data();
      if (mode == 1) {
        zero.assign(3 * (size_t)st.n, 0.0);
        v = zero.data();
      }
      check(dfr_load_fluid_state(ctx, st.x.data(), v, st.kappa.empty() ? nullptr : st.kappa.data(), st.kappa_v.empty() ? nullptr : st.kappa_v.data()));
      info_valid = false;
      step_serial++;
      return;
    }
    std::FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) {  // the reference logs a warning and carries on (SimulatorBase.cpp:2547-2558)
      std::fprintf(stderr, "[warn] state file %s not found; state unchanged\n", path.c_str());
      return;
    }
    char magic[8];
    int64_t n = 0;
    double t = 0.0;
    const bool ok = std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "DFRSTAT1", 8) == 0 && std::fread(&n, sizeof(n), 1, f) == 1 &&
                    std::fread(&t, sizeof(t), 1, f) == 1;
    if (!ok || n != dfr_num_fluid(ctx)) {
      std::fclose(f);
      throw std::runtime_error("state file " + path + ": not a DFRSTAT1 file of this scene's particle count");
    }
    std::vector<double> x(3 * n), v(3 * n), k(n), kv(n);
    const bool rd = std::fread(x.data(), sizeof(double), x.size(), f) == x.size() && std::fread(v.data(), sizeof(double), v.size(), f) == v.size() &&
                    std::fread(k.data(), sizeof(double), k.size(), f) == k.size() && std::fread(kv.data(), sizeof(double), kv.size(), f) == kv.size();
    std::fclose(f);
    if (!rd) throw std::runtime_error("state file " + path + " is truncated");
    // same semantics as the .bgeo branch: every field of the file is loaded, --load-fluid-pos then clears the velocities
    // (loadFluidParticlePositions -> clearVelocities, SimulatorBase.cpp:2043-2058)
    if (mode == 1) std::fill(v.begin(), v.end(), 0.0);
    check(dfr_load_fluid_state(ctx, x.data(), v.data(), k.data(), kv.data()));
    info_valid = false;
    step_serial++;
  }
  void loadState(const std::string &path) { loadStateFile(path, 0); }
  void checkLoadState() {  // SimulatorBase.cpp:1052-1070
    if (state_file.empty()) return;
    loadStateFile(state_file, load_pos ? 1 : (load_pos_vel ? 2 : 0));
  }
  void setStateExportPath(const std::string &p) { state_export_path = p; }
  std::string getOutputPath() const { return output_path; }
  std::string getStateFile() const { return state_file; }
  void setStateFile(const std::string &s) { state_file = s; }

  void setValueBool(int id, bool v) { values_f[id] = v ? 1.0 : 0.0; }
  void setValueInt(int id, int v) { values_f[id] = v; }
  void setValueFloat(int id, double v) { values_f[id] = v; }
  bool getValueBool(int id) const { return value(id) != 0.0; }
  int getValueInt(int id) const { return (int)value(id); }
  double getValueFloat(int id) const { return value(id); }

  RigidBodyGradientManager *getRigidBodyGradientManager() { return &grad_manager; }
  Simulation *getSimulation() { return &simulation; }
  TimeManager *getTimeManager() { return &time_manager; }

  // ---- plumbing used by the mirrors --------------------------------------------------------------------------
  dfr_context *context() const { return ctx; }
  const Scene &getScene() const { return scene; }
  Scene &mutableScene() { return scene; }
  void check(int rc) const {
    if (rc < 0) throw DfrFailure(rc, dfr_last_error(ctx));
  }
  const dfr_step_info &info() const {
    if (!info_valid) {
      if (finalized)
        check(dfr_get_step_info(ctx, &step_info));
      else {
        std::memset(&step_info, 0, sizeof(step_info));
        step_info.time_step_size = scene.cfg.time_step_size;
      }
      info_valid = finalized;
    }
    return step_info;
  }
  std::array<double, 13> body_state(int i) const {
    std::array<double, 13> s{};
    if (finalized)
      check(dfr_get_body_state(ctx, i, s.data()));
    else {  // before deferredInit: the scene pose, at rest
      const BodyDesc &b = scene.bodies.at(i);
      for (int k = 0; k < 3; k++) s[k] = b.translation[k];
      for (int k = 0; k < 4; k++) s[3 + k] = b.rotation[k];
    }
    return s;
  }
  BoundaryModelAkinci2012 *boundary_model(unsigned int i) {
    if (i >= models.size()) throw std::out_of_range("boundary model index " + std::to_string(i));  // the reference asserts (:2079-2081)
    return models[i].get();
  }
  bool trajectory_finished() const { return info().trajectory_finished != 0; }
  int serial() const { return step_serial; }
  bool isFinalized() const { return finalized; }
  void log(const std::string &s) {
    std::printf("%s\n", s.c_str());
    if (log_file) {
      std::fprintf(log_file, "%s\n", s.c_str());
      std::fflush(log_file);
    }
  }
  double wall_ms_steps = 0.0, device_ms_steps = 0.0;
  long long steps_taken = 0, launches_steps = 0;
  // Utilities::Timing::printAverageTimes / printTimeSums (opt-ng.py:178-179): the reference prints the START_TIMING
  // averages of its phases; here one line for the host wall time of a step and one for the device time
  std::string timing_report(bool sums) const {
    std::ostringstream os;
    const double n = (double)std::max<long long>(steps_taken, 1);
    const double dev_ms = device_ms_steps;
    const long long launches = launches_steps;
    os << "---------------------------------------------------------------------------\n";
    if (sums) {
      os << "Time sums:\n  timeStepNoGUI (host wall, read-back included): " << wall_ms_steps << " ms over " << steps_taken << " steps\n"
         << "  device (CUDA events around dfr_step): " << dev_ms << " ms, " << launches << " kernel launches\n";
    } else {
      os << "Average times:\n  timeStepNoGUI (host wall, read-back included): " << wall_ms_steps / n << " ms\n"
         << "  device (CUDA events around dfr_step): " << dev_ms / n << " ms, " << (double)launches / n << " kernel launches per step\n";
    }
    os << "---------------------------------------------------------------------------\n";
    return os.str();
  }

 private:
  void build_context() {
    check_create(dfr_create(&scene.cfg, device, &ctx));
    const int64_t n = (int64_t)scene.fluid_x.size() / 3;
    check(dfr_set_fluid(ctx, n, scene.fluid_x.data(), scene.fluid_v.data()));
    for (size_t i = 0; i < scene.bodies.size(); i++) {
      const BodyDesc &b = scene.bodies[i];
      const int idx = dfr_add_body(ctx, (int64_t)b.samples.size() / 3, b.samples.data(), b.dynamic ? 1 : 0, b.density, b.translation.data(),
                                   b.rotation.data());
      check(idx);
      if (b.dynamic) check(dfr_set_init_v_omega(ctx, idx, b.init_v.data(), b.init_omega.data()));
      models.emplace_back(new BoundaryModelAkinci2012(this, idx));
    }
    for (const EmitterDesc &e : scene.emitters) {
      double R[9];
      quat_to_matrix(e.rotation, R);
      check(dfr_add_emitter(ctx, e.width, e.height, e.x.data(), R, e.velocity, e.emit_start, e.emit_end));
    }
  }
  void check_create(int rc) {
    if (rc == DFR_ERR_NO_DEVICE) throw DfrFailure(rc, "no CUDA device: this module has no CPU fallback");
    if (rc < 0) throw DfrFailure(rc, "dfr_create failed");
  }
  void step_once(bool callbacks) {
    deferredInit();
    if (callbacks && step_cb_before) step_cb_before();
    const auto t0 = std::chrono::steady_clock::now();
    double dev0 = 0.0, dev1 = 0.0;
    int64_t l0 = 0, l1 = 0;
    dfr_get_device_time_ms(ctx, &dev0, &l0);  // cumulative since the last reset
    check(dfr_step(ctx, 1));
    dfr_get_device_time_ms(ctx, &dev1, &l1);
    device_ms_steps += dev1 - dev0;
    launches_steps += l1 - l0;
    wall_ms_steps += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    steps_taken++;
    info_valid = false;
    step_serial++;
    // TimeStepDiffDFSPH::endStep (:430-470): the callback flag follows the end-of-trajectory test of this step
    timestep.trajectory_finished_cb = trajectory_finished();
    if (timestep.trajectory_finished_cb) timestep.custom_log.clear();
    if (callbacks && step_cb) step_cb();
  }
  double value(int id) const {
    auto it = values_f.find(id);
    return it == values_f.end() ? 0.0 : it->second;
  }

  Simulation simulation;
  TimeManager time_manager;
  TimeStepDiffDFSPH timestep;
  RigidBodyGradientManager grad_manager;
  std::vector<std::unique_ptr<BoundaryModelAkinci2012>> models;
  friend class Simulation;

  Scene scene;
  dfr_context *ctx = nullptr;
  int device = 0;
  bool finalized = false, use_gui = false, load_pos = false, load_pos_vel = false, stop_requested = false;
  std::string scene_file, state_file, output_path, param_str, state_export_path;
  std::map<int, double> values_f{{STOP_AT, -1.0}};
  std::function<void()> step_cb, step_cb_before, reset_cb;
  mutable dfr_step_info step_info;
  mutable bool info_valid = false;
  int step_serial = 0;
  std::FILE *log_file = nullptr;
};

inline Simulation *Simulation::current = nullptr;
inline TimeManager *TimeManager::current = nullptr;

// ---- inline members ---------------------------------------------------------------------------------------------
inline bool RigidBodyObject::isDynamic() const { return base->getScene().bodies.at(index).dynamic; }
inline double RigidBodyObject::getMass() const {
  double p[17];
  base->check(dfr_get_body_properties(base->context(), index, p));
  return p[0];
}
inline Vec3 RigidBodyObject::getPosition() const { auto s = base->body_state(index); return {s[0], s[1], s[2]}; }
inline Vec3 RigidBodyObject::getVelocity() const { auto s = base->body_state(index); return {s[7], s[8], s[9]}; }
inline Vec3 RigidBodyObject::getAngularVelocity() const { auto s = base->body_state(index); return {s[10], s[11], s[12]}; }
inline Quat RigidBodyObject::getRotationXYZW() const { auto s = base->body_state(index); return {s[4], s[5], s[6], s[3]}; }
inline void RigidBodyObject::setVelocity(const Vec3 &v) { base->check(dfr_set_body_velocity(base->context(), index, v.data(), nullptr)); }
inline void RigidBodyObject::setAngularVelocity(const Vec3 &w) { base->check(dfr_set_body_velocity(base->context(), index, nullptr, w.data())); }

inline unsigned int BoundaryModelAkinci2012::numberOfParticles() const { return (unsigned int)(base->getScene().bodies.at(index).samples.size() / 3); }
inline Vec3 BoundaryModelAkinci2012::get_position_rb() const { return rbo.getPosition(); }
inline Quat BoundaryModelAkinci2012::get_quaternion_rb_vec4() const { auto s = base->body_state(index); return {s[3], s[4], s[5], s[6]}; }
inline Vec3 BoundaryModelAkinci2012::get_velocity_rb() const { return rbo.getVelocity(); }
inline Vec3 BoundaryModelAkinci2012::get_angular_velocity_rb() const { return rbo.getAngularVelocity(); }
inline void BoundaryModelAkinci2012::set_velocity_rb(const Vec3 &v) { const_cast<RigidBodyObject &>(rbo).setVelocity(v); }
inline void BoundaryModelAkinci2012::set_angular_velocity_rb(const Vec3 &w) { const_cast<RigidBodyObject &>(rbo).setAngularVelocity(w); }
inline std::array<double, 12> BoundaryModelAkinci2012::grad(int which) const {
  std::array<double, 12> out{};
  if (base->isFinalized())
    base->check(dfr_get_body_grad(base->context(), index, which, out.data()));
  else if (which == 4 || which == 7)  // BoundaryModel_Akinci2012::reset: d v/d v0 = d omega/d omega0 = I
    out[0] = out[4] = out[8] = 1.0;
  return out;
}
inline Vec3 BoundaryModelAkinci2012::getForce() const {
  double p[17];
  base->check(dfr_get_body_properties(base->context(), index, p));
  return {p[11], p[12], p[13]};
}
inline Vec3 BoundaryModelAkinci2012::getTorque() const {
  double p[17];
  base->check(dfr_get_body_properties(base->context(), index, p));
  return {p[14], p[15], p[16]};
}
inline const std::vector<double> &BoundaryModelAkinci2012::field(int f) const {
  if (cache_step[f] != base->serial()) {
    const size_t n = numberOfParticles();
    cache[f].assign(n * (f == 2 ? 1 : 3), 0.0);
    if (n) base->check(dfr_download_body(base->context(), index, f, cache[f].data()));
    cache_step[f] = base->serial();
  }
  return cache[f];
}
inline Vec3 BoundaryModelAkinci2012::particle(int f, unsigned int i) const {
  const auto &a = field(f);
  if (3 * (size_t)i + 2 >= a.size()) throw std::out_of_range("boundary particle index");
  return {a[3 * i], a[3 * i + 1], a[3 * i + 2]};
}
inline double BoundaryModelAkinci2012::getVolume(unsigned int i) const { return field(2).at(i); }

inline void RigidBodyGradientManager::reset() { /* part of SimulatorBase::reset here (dfr_reset) */ }
inline std::array<double, 12> RigidBodyGradientManager::grad(int R, int RR, int which) const {
  std::array<double, 12> out{};
  base->check(dfr_get_manager_grad(base->context(), R, RR, which, out.data()));
  return out;
}

inline BoundaryModelAkinci2012 *TimeStepDiffDFSPH::get_boundary_model(unsigned int i) { return base->boundary_model(i); }
inline void TimeStepDiffDFSPH::set_init_v_rb(unsigned int i, const Vec3 &v) {
  BodyDesc &b = base->mutableScene().bodies.at(i);
  b.init_v = v;
  base->check(dfr_set_init_v_omega(base->context(), (int)i, b.init_v.data(), b.init_omega.data()));
}
inline void TimeStepDiffDFSPH::set_init_omega_rb(unsigned int i, const Vec3 &w) {
  BodyDesc &b = base->mutableScene().bodies.at(i);
  b.init_omega = w;
  base->check(dfr_set_init_v_omega(base->context(), (int)i, b.init_v.data(), b.init_omega.data()));
}
inline Vec3 TimeStepDiffDFSPH::get_init_v_rb(unsigned int i) const { return base->getScene().bodies.at(i).init_v; }
inline Vec3 TimeStepDiffDFSPH::get_init_omega_rb(unsigned int i) const { return base->getScene().bodies.at(i).init_omega; }
inline Vec3 TimeStepDiffDFSPH::get_target_x(unsigned int i) const { return base->getScene().bodies.at(i).target_x; }
inline void TimeStepDiffDFSPH::set_target_x(unsigned int i, const Vec3 &x) { base->mutableScene().bodies.at(i).target_x = x; }
inline Vec3 TimeStepDiffDFSPH::get_target_angle_in_radian(unsigned int i) const {
  const Vec3 &d = base->getScene().bodies.at(i).target_angle_deg;
  return {d[0] / 180.0 * M_PI, d[1] / 180.0 * M_PI, d[2] / 180.0 * M_PI};
}
inline Quat TimeStepDiffDFSPH::get_target_quaternion_vec4(unsigned int i) const {
  return quat_from_euler_deg(base->getScene().bodies.at(i).target_angle_deg);  // SimulationDataDiffDFSPH.h:169-185
}
inline unsigned int TimeStepDiffDFSPH::get_step_count() const { return (unsigned int)base->info().step_count; }
// TimeStepDiffDFSPH::countNeighborDOF (TimeStepDiffDFSPH.cpp:283-350, getter :2242): fluid particles that have at least
// one neighbour among the particles of a dynamic (or animated) body, summed over those bodies.  The reference counts
// this inside step() when `enable count neighbor dof` is set; here it is evaluated on demand from the fluid -> body
// neighbour sets of the current positions (a diagnostic of experiments/others/python/grad-sensitivity-1ring.py).
inline unsigned int TimeStepDiffDFSPH::get_num_1ring_fluid_particle() const {
  dfr_context *ctx = base->context();
  if (!ctx) return 0;
  const int64_t n = dfr_num_fluid(ctx);
  std::vector<int32_t> counts((size_t)std::max<int64_t>(n, 1));
  unsigned int total = 0;
  const auto &bodies = base->getScene().bodies;
  for (size_t b = 0; b < bodies.size(); b++) {
    if (!bodies[b].dynamic) continue;
    int64_t entries = 0;
    base->check(dfr_get_neighbors(ctx, -1, (int)b, counts.data(), nullptr, 0, &entries));
    for (int64_t i = 0; i < n; i++) total += counts[(size_t)i] > 0 ? 1u : 0u;
  }
  return total;
}
inline unsigned int TimeStepDiffDFSPH::getIterations() const { return (unsigned int)base->info().iterations; }
inline unsigned int TimeStepDiffDFSPH::getIterationsV() const { return (unsigned int)base->info().iterations_v; }
inline void TimeStepDiffDFSPH::add_log(const std::string &s) { base->log(s); }
inline void TimeStepDiffDFSPH::reset_gradient() { base->check(dfr_reset_gradient(base->context())); }

inline double TimeManager::getTime() const { return base->info().time; }
inline double TimeManager::getTimeStepSize() const { return base->info().time_step_size; }

inline TimeStepDiffDFSPH *Simulation::getTimeStep() { return &base->timestep; }
inline BoundaryModelAkinci2012 *Simulation::getBoundaryModel(unsigned int i) { return base->boundary_model(i); }
inline unsigned int Simulation::numberOfBoundaryModels() const { return (unsigned int)base->getScene().bodies.size(); }
inline unsigned int Simulation::numberOfFluidModels() const { return base->getScene().fluid_x.empty() ? 0u : 1u; }
inline unsigned int Simulation::numberOfFluidParticles() const {
  return base->isFinalized() ? (unsigned int)base->info().num_fluid_particles : (unsigned int)(base->getScene().fluid_x.size() / 3);
}
inline void Simulation::setGradientMode(int m) {
  base->check(dfr_set_gradient_mode(base->context(), m));
  base->mutableScene().cfg.gradient_mode = m;
}
inline int Simulation::getGradientMode() const { return base->getScene().cfg.gradient_mode; }
inline bool Simulation::useRigidGradientManager() const { return base->getScene().cfg.use_rigid_gradient_manager != 0; }
inline bool Simulation::useRigidContactSolver() const { return base->getScene().cfg.use_rigid_contact_solver != 0; }
inline double Simulation::getParticleRadius() const { return base->getScene().cfg.particle_radius; }
inline double Simulation::getSupportRadius() const { return 4.0 * base->getScene().cfg.particle_radius; }

}  // namespace dfrhost
