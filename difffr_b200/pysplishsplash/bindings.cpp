// pybind11 module `pysplishsplash` — the reference's Python surface for the DiffDFSPH path
// (pySPlisHSPlasH/main.cpp:47-76 and the *Module.cpp files cited below), bound to the host classes of
// difffr_b200/host/simulator_host.hpp, which drive the CUDA path through the C ABI of include/dfr.h.
// Same class names, method names, argument meaning and return shapes as the reference bindings, so
// experiments/rigid_body_trajectory_optimization/python/*.py run against it unchanged.  Everything the
// reference module exposes that is not on this path (other solvers, exporters, GUI widgets) is absent.
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "../host/simulator_host.hpp"

namespace py = pybind11;
using namespace pybind11::literals;
using namespace dfrhost;

namespace {

py::array_t<double> np_vec3(const Vec3 &v) {
  py::array_t<double> a(3);
  std::copy(v.begin(), v.end(), a.mutable_data());
  return a;
}
py::array_t<double> np_vec4(const Quat &v) {
  py::array_t<double> a(4);
  std::copy(v.begin(), v.end(), a.mutable_data());
  return a;
}
py::array_t<double> np_mat(const std::array<double, 12> &m, int rows, int cols) {  // Eigen -> numpy copy (pybind11/eigen.h in the reference)
  py::array_t<double> a({rows, cols});
  std::copy(m.begin(), m.begin() + rows * cols, a.mutable_data());
  return a;
}
Vec3 to_vec3(const py::object &o) {
  auto a = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(o);
  if (!a || a.size() != 3) throw py::value_error("expected a vector of 3 numbers");
  return {a.data()[0], a.data()[1], a.data()[2]};
}

// stand-in for SPH::Simulator_GUI_imgui (GUIModule.cpp:52-53): the scripts construct it and hand it to setGui
struct GuiStub {
  explicit GuiStub(SimulatorBase *) {}
};
struct BoundarySimulatorStub {};

}  // namespace

PYBIND11_MODULE(_core, m) {
  m.doc() = "B200-native drop-in for the DiffDFSPH path of pysplishsplash";
  py::register_exception<DfrFailure>(m, "DfrError", PyExc_RuntimeError);

  // ---- RigidBodyModule.cpp:18-45 ------------------------------------------------------------------------------
  py::class_<RigidBodyObject>(m, "RigidBodyObject")
      .def("isDynamic", &RigidBodyObject::isDynamic)
      .def("getMass", &RigidBodyObject::getMass)
      .def("getPosition", [](const RigidBodyObject &o) { return np_vec3(o.getPosition()); })
      .def("getVelocity", [](const RigidBodyObject &o) { return np_vec3(o.getVelocity()); })
      .def("getAngularVelocity", [](const RigidBodyObject &o) { return np_vec3(o.getAngularVelocity()); })
      .def("getRotation", [](const RigidBodyObject &o) { return np_vec4(o.getRotationXYZW()); })
      .def("setVelocity", [](RigidBodyObject &o, const py::object &v) { o.setVelocity(to_vec3(v)); })
      .def("setAngularVelocity", [](RigidBodyObject &o, const py::object &v) { o.setAngularVelocity(to_vec3(v)); });

  // ---- BoundaryModelModule.cpp:23-91 --------------------------------------------------------------------------
  auto bm = py::class_<BoundaryModelAkinci2012>(m, "BoundaryModelAkinci2012");
  bm.def("numberOfParticles", &BoundaryModelAkinci2012::numberOfParticles)
      .def("getRigidBodyObject", &BoundaryModelAkinci2012::getRigidBodyObject, py::return_value_policy::reference_internal)
      .def("getPosition", [](const BoundaryModelAkinci2012 &b, unsigned int i) { return np_vec3(b.particle(0, i)); })
      .def("getVelocity", [](const BoundaryModelAkinci2012 &b, unsigned int i) { return np_vec3(b.particle(1, i)); })
      .def("getPosition0", [](const BoundaryModelAkinci2012 &b, unsigned int i) { return np_vec3(b.particle(3, i)); })
      .def("getVolume", &BoundaryModelAkinci2012::getVolume)
      .def("getForce", [](const BoundaryModelAkinci2012 &b) { return np_vec3(b.getForce()); })
      .def("getTorque", [](const BoundaryModelAkinci2012 &b) { return np_vec3(b.getTorque()); })
      .def("get_position_rb", [](const BoundaryModelAkinci2012 &b) { return np_vec3(b.get_position_rb()); })
      .def("get_quaternion_rb_vec4", [](const BoundaryModelAkinci2012 &b) { return np_vec4(b.get_quaternion_rb_vec4()); })
      .def("get_velocity_rb", [](const BoundaryModelAkinci2012 &b) { return np_vec3(b.get_velocity_rb()); })
      .def("get_angular_velocity_rb", [](const BoundaryModelAkinci2012 &b) { return np_vec3(b.get_angular_velocity_rb()); })
      .def("set_velocity_rb", [](BoundaryModelAkinci2012 &b, const py::object &v) { b.set_velocity_rb(to_vec3(v)); })
      .def("set_angular_velocity_rb", [](BoundaryModelAkinci2012 &b, const py::object &v) { b.set_angular_velocity_rb(to_vec3(v)); });
  {
    struct G { const char *name; int which, rows; };
    static const G grads[] = {{"get_grad_x_to_v0", 0, 3},          {"get_grad_x_to_omega0", 1, 3},          {"get_grad_quaternion_to_v0", 2, 4},
                              {"get_grad_quaternion_to_omega0", 3, 4}, {"get_grad_v_to_v0", 4, 3},          {"get_grad_v_to_omega0", 5, 3},
                              {"get_grad_omega_to_v0", 6, 3},      {"get_grad_omega_to_omega0", 7, 3}};
    for (const G &g : grads) {
      const int which = g.which, rows = g.rows;
      bm.def(g.name, [which, rows](const BoundaryModelAkinci2012 &b) { return np_mat(b.grad(which), rows, 3); });
    }
    // per-step net Jacobians (BoundaryModel_Akinci2012.h:52-61), not bound in the reference but handy for FD checks
    bm.def("get_grad_net", [](const BoundaryModelAkinci2012 &b, int which) {
      if (which < 8 || which > 15) throw py::value_error("which must be 8..15 (include/dfr.h)");
      const bool q = (which == 10 || which == 14);
      return np_mat(b.grad(which), 3, q ? 4 : 3);
    });
  }

  // ---- DiffDFSPHModule.cpp:44-102 -----------------------------------------------------------------------------
  py::class_<TimeStepDiffDFSPH>(m, "TimeStepDiffDFSPH")
      .def("get_boundary_model", &TimeStepDiffDFSPH::get_boundary_model, py::return_value_policy::reference_internal)
      .def("get_loss", [](const TimeStepDiffDFSPH &t) { return t.loss; })
      .def("set_loss", [](TimeStepDiffDFSPH &t, double v) { t.loss = v; })
      .def("get_loss_x", [](const TimeStepDiffDFSPH &t) { return t.loss_x; })
      .def("set_loss_x", [](TimeStepDiffDFSPH &t, double v) { t.loss_x = v; })
      .def("get_loss_rotation", [](const TimeStepDiffDFSPH &t) { return t.loss_rotation; })
      .def("set_loss_rotation", [](TimeStepDiffDFSPH &t, double v) { t.loss_rotation = v; })
      .def("get_lr", [](const TimeStepDiffDFSPH &t) { return t.lr; })
      .def("set_lr", [](TimeStepDiffDFSPH &t, double v) { t.lr = v; })
      .def("set_init_v_rb", [](TimeStepDiffDFSPH &t, unsigned int i, const py::object &v) { t.set_init_v_rb(i, to_vec3(v)); })
      .def("set_init_omega_rb", [](TimeStepDiffDFSPH &t, unsigned int i, const py::object &v) { t.set_init_omega_rb(i, to_vec3(v)); })
      // no articulated systems on this path: identical to set_init_omega_rb (TimeStepDiffDFSPH.cpp:2097-2116, system == nullptr)
      .def("set_init_omega_rb_to_joint", [](TimeStepDiffDFSPH &t, unsigned int i, const py::object &v) { t.set_init_omega_rb(i, to_vec3(v)); })
      .def("get_init_v_rb", [](const TimeStepDiffDFSPH &t, unsigned int i) { return np_vec3(t.get_init_v_rb(i)); })
      .def("get_init_omega_rb", [](const TimeStepDiffDFSPH &t, unsigned int i) { return np_vec3(t.get_init_omega_rb(i)); })
      .def("get_target_x", [](const TimeStepDiffDFSPH &t, unsigned int i) { return np_vec3(t.get_target_x(i)); })
      .def("set_target_x", [](TimeStepDiffDFSPH &t, unsigned int i, const py::object &v) { t.set_target_x(i, to_vec3(v)); })
      .def("get_target_angle_in_radian", [](const TimeStepDiffDFSPH &t, unsigned int i) { return np_vec3(t.get_target_angle_in_radian(i)); })
      .def("get_target_quaternion_vec4", [](const TimeStepDiffDFSPH &t, unsigned int i) { return np_vec4(t.get_target_quaternion_vec4(i)); })
      .def("is_trajectory_finish_callback", &TimeStepDiffDFSPH::is_trajectory_finish_callback)
      .def("clear_all_callbacks", &TimeStepDiffDFSPH::clear_all_callbacks)
      .def("is_in_new_trajectory", &TimeStepDiffDFSPH::is_in_new_trajectory)
      .def("set_custom_log_message", &TimeStepDiffDFSPH::set_custom_log_message)
      .def("get_custom_log_message", &TimeStepDiffDFSPH::get_custom_log_message)
      .def("get_step_count", &TimeStepDiffDFSPH::get_step_count)
      .def("add_log", &TimeStepDiffDFSPH::add_log)
      .def("reset_gradient", &TimeStepDiffDFSPH::reset_gradient)
      .def("get_num_1ring_fluid_particle", &TimeStepDiffDFSPH::get_num_1ring_fluid_particle)
      .def("getIterations", &TimeStepDiffDFSPH::getIterations)
      .def("getIterationsV", &TimeStepDiffDFSPH::getIterationsV);

  // ---- TimeModule.cpp:23-29 -----------------------------------------------------------------------------------
  py::class_<TimeManager>(m, "TimeManager")
      .def_static("getCurrent", []() {
        if (!TimeManager::current) throw std::runtime_error("no simulation initialised");
        return TimeManager::current;
      }, py::return_value_policy::reference)
      .def_static("hasCurrent", []() { return TimeManager::current != nullptr; })
      .def("getTime", &TimeManager::getTime)
      .def("getTimeStepSize", &TimeManager::getTimeStepSize);

  // ---- SimulationModule.cpp:95-211 ----------------------------------------------------------------------------
  py::class_<Simulation>(m, "Simulation")
      .def_static("getCurrent", []() {
        if (!Simulation::current) throw std::runtime_error("no simulation initialised");
        return Simulation::current;
      }, py::return_value_policy::reference)
      .def_static("hasCurrent", []() { return Simulation::current != nullptr; })
      .def("getTimeStep", &Simulation::getTimeStep, py::return_value_policy::reference_internal)
      .def("getBoundaryModel", &Simulation::getBoundaryModel, py::return_value_policy::reference_internal)
      .def("numberOfBoundaryModels", &Simulation::numberOfBoundaryModels)
      .def("numberOfFluidModels", &Simulation::numberOfFluidModels)
      .def("numberOfFluidParticles", &Simulation::numberOfFluidParticles)
      .def("setGradientMode", &Simulation::setGradientMode)
      .def("getGradientMode", &Simulation::getGradientMode)
      .def("useRigidGradientManager", &Simulation::useRigidGradientManager)
      .def("useRigidContactSolver", &Simulation::useRigidContactSolver)
      .def("getParticleRadius", &Simulation::getParticleRadius)
      .def("getSupportRadius", &Simulation::getSupportRadius)
      .def("is2DSimulation", [](const Simulation &) { return false; });

  // ---- Exec: SimulationModule.cpp:236-414 ---------------------------------------------------------------------
  py::module_ exec = m.def_submodule("Exec");
  py::class_<RigidBodyGradientManager>(exec, "RigidBodyGradientManager")
      .def("reset", &RigidBodyGradientManager::reset)
      .def("get_grad_x_to_v0", [](const RigidBodyGradientManager &g, int R, int RR) { return np_mat(g.grad(R, RR, 0), 3, 3); })
      .def("get_grad_x_to_omega0", [](const RigidBodyGradientManager &g, int R, int RR) { return np_mat(g.grad(R, RR, 1), 3, 3); })
      .def("get_grad_q_to_v0", [](const RigidBodyGradientManager &g, int R, int RR) { return np_mat(g.grad(R, RR, 2), 4, 3); })
      .def("get_grad_q_to_omega0", [](const RigidBodyGradientManager &g, int R, int RR) { return np_mat(g.grad(R, RR, 3), 4, 3); })
      .def("get_grad_net_force_to_vn", [](const RigidBodyGradientManager &g, int R, int RR) { return np_mat(g.grad(R, RR, 8), 3, 3); })
      .def("get_grad_net_torque_to_omega_n", [](const RigidBodyGradientManager &g, int R, int RR) { return np_mat(g.grad(R, RR, 15), 3, 3); });
  py::class_<BoundarySimulatorStub>(exec, "BoundarySimulator");

  py::class_<SimulatorBase>(exec, "SimulatorBase")
      .def(py::init<>())
      .def_property_readonly_static("PAUSE", [](py::object) { return (int)SimulatorBase::PAUSE; })
      .def_property_readonly_static("PAUSE_AT", [](py::object) { return (int)SimulatorBase::PAUSE_AT; })
      .def_property_readonly_static("STOP_AT", [](py::object) { return (int)SimulatorBase::STOP_AT; })
      .def_property_readonly_static("NUM_STEPS_PER_RENDER", [](py::object) { return (int)SimulatorBase::NUM_STEPS_PER_RENDER; })
      .def_property_readonly_static("DATA_EXPORT_FPS", [](py::object) { return (int)SimulatorBase::DATA_EXPORT_FPS; })
      .def_property_readonly_static("STATE_EXPORT", [](py::object) { return (int)SimulatorBase::STATE_EXPORT; })
      .def_property_readonly_static("STATE_EXPORT_FPS", [](py::object) { return (int)SimulatorBase::STATE_EXPORT_FPS; })
      .def("init", &SimulatorBase::init, "sceneFile"_a = "data/Scenes/DoubleDamBreak.json", "programName"_a = "pySPlisHSPlasH",
           "useCache"_a = true, "stateFile"_a = "", "loadFluidPos"_a = false, "loadFluidPosAndVel"_a = false, "outputDir"_a = "",
           "initialPause"_a = true, "useGui"_a = true, "stopAt"_a = -1.0, "param"_a = "")
      .def("setDevice", &SimulatorBase::setDevice, "CUDA device of this simulator's context (extension; default 0)")
      .def("setGui", [](SimulatorBase &, py::object) {})
      .def("initSimulation", &SimulatorBase::initSimulation)
      .def("initSimulationWithDeferredInit", &SimulatorBase::initSimulationWithDeferredInit)
      .def("runSimulation", &SimulatorBase::runSimulation)
      .def("runNewTrajectory", &SimulatorBase::runNewTrajectory)
      .def("forwardFixedSteps", &SimulatorBase::forwardFixedSteps)
      .def("runFixedTimeSteps", &SimulatorBase::forwardFixedSteps)
      .def("singleTimeStep", &SimulatorBase::singleTimeStep)
      .def("timeStep", &SimulatorBase::singleTimeStep)
      .def("timeStepNoGUI", &SimulatorBase::timeStepNoGUI)
      .def("reset", &SimulatorBase::reset)
      .def("cleanup", &SimulatorBase::cleanup)
      .def("stop", &SimulatorBase::stop, "leave runSimulation after the current step (the reference's GUI stop button)")
      .def("setTimeStepCB", &SimulatorBase::setTimeStepCB)
      .def("setTimeStepCallBefore", &SimulatorBase::setTimeStepCallBefore)
      .def("setResetCB", &SimulatorBase::setResetCB)
      .def("saveState", &SimulatorBase::saveState, "stateFile"_a = "")
      .def("loadState", &SimulatorBase::loadState)
      .def("loadStateWithRigidExisted", &SimulatorBase::loadState)
      .def("setStateExportPath", &SimulatorBase::setStateExportPath)
      .def("getOutputPath", &SimulatorBase::getOutputPath)
      .def("getStateFile", &SimulatorBase::getStateFile)
      .def("setStateFile", &SimulatorBase::setStateFile)
      .def("getBoundarySimulator", [](SimulatorBase &) { return BoundarySimulatorStub(); })
      .def("getRigidBodyGradientManager", &SimulatorBase::getRigidBodyGradientManager, py::return_value_policy::reference_internal)
      .def("setValueBool", &SimulatorBase::setValueBool)
      .def("setValueInt", &SimulatorBase::setValueInt)
      .def("setValueFloat", &SimulatorBase::setValueFloat)
      .def("getValueBool", &SimulatorBase::getValueBool)
      .def("getValueInt", &SimulatorBase::getValueInt)
      .def("getValueFloat", &SimulatorBase::getValueFloat)
      // scene assembled in Python instead of a JSON file (tests / synthetic scenes); not in the reference
      .def("initSimulationFromArrays",
           [](SimulatorBase &b, py::dict cfg, py::array_t<double, py::array::c_style | py::array::forcecast> fluid_x,
              py::list bodies) {
             Scene sc;
             dfr_default_config(&sc.cfg);
             auto setd = [&](const char *k, double &dst) { if (cfg.contains(k)) dst = cfg[k].cast<double>(); };
             auto seti = [&](const char *k, int32_t &dst) { if (cfg.contains(k)) dst = cfg[k].cast<int>(); };
             setd("particle_radius", sc.cfg.particle_radius); setd("time_step_size", sc.cfg.time_step_size);
             setd("cfl_max_time_step", sc.cfg.cfl_max_time_step); setd("max_error", sc.cfg.max_error); setd("max_error_v", sc.cfg.max_error_v);
             setd("target_time", sc.cfg.target_time); setd("uniform_acc_rb_time", sc.cfg.uniform_acc_rb_time);
             setd("surface_tension", sc.cfg.surface_tension); setd("viscosity", sc.cfg.viscosity);
             seti("cfl_method", sc.cfg.cfl_method); seti("gradient_mode", sc.cfg.gradient_mode);
             seti("surface_tension_method", sc.cfg.surface_tension_method); seti("viscosity_method", sc.cfg.viscosity_method);
             seti("use_rigid_gradient_manager", sc.cfg.use_rigid_gradient_manager); seti("rigid_body_mode", sc.cfg.rigid_body_mode);
             sc.fluid_x.assign(fluid_x.data(), fluid_x.data() + fluid_x.size());
             sc.fluid_v.assign(sc.fluid_x.size(), 0.0);
             for (py::handle h : bodies) {
               py::dict d = py::reinterpret_borrow<py::dict>(h);
               BodyDesc bd;
               auto xs = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(d["x_local"]);
               bd.samples.assign(xs.data(), xs.data() + xs.size());
               bd.dynamic = d["dynamic"].cast<bool>();
               bd.density = d["density"].cast<double>();
               bd.translation = to_vec3(d["position"]);
               auto q = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(d["quat"]);
               for (int k = 0; k < 4; k++) bd.rotation[k] = q.data()[k];
               if (d.contains("init_v")) bd.init_v = to_vec3(d["init_v"]);
               if (d.contains("init_omega")) bd.init_omega = to_vec3(d["init_omega"]);
               if (d.contains("target_x")) bd.target_x = to_vec3(d["target_x"]);
               if (d.contains("target_angle_deg")) bd.target_angle_deg = to_vec3(d["target_angle_deg"]);
               sc.bodies.push_back(std::move(bd));
             }
             b.initSimulationFromScene(sc);
           },
           "config"_a, "fluid_x"_a, "bodies"_a);

  // ---- GUIModule.cpp:52-53, UtilitiesModule.cpp (Timing) --------------------------------------------------------
  py::module_ gui = m.def_submodule("GUI");
  py::class_<GuiStub>(gui, "Simulator_GUI_imgui").def(py::init<SimulatorBase *>());
  py::module_ util = m.def_submodule("Utilities");
  struct TimingStub {};
  py::class_<TimingStub>(util, "Timing")
      .def_static("printAverageTimes", []() {
        if (Simulation::current) py::print(Simulation::current->simulatorBase()->timing_report(false));
      })
      .def_static("printTimeSums", []() {
        if (Simulation::current) py::print(Simulation::current->simulatorBase()->timing_report(true));
      })
      .def_static("reset", []() {});

  // partio .bgeo fluid state files (host/state_io.hpp), exposed for tests and for exporting settled states
  m.def("_read_bgeo", [](const std::string &path) {
    FluidStateFile st = read_bgeo(path);
    py::dict d;
    auto arr = [](const std::vector<double> &v, py::ssize_t cols) {
      if (cols > 1) {
        py::array_t<double> a({(py::ssize_t)(v.size() / cols), cols});
        std::copy(v.begin(), v.end(), a.mutable_data());
        return a;
      }
      py::array_t<double> a((py::ssize_t)v.size());
      std::copy(v.begin(), v.end(), a.mutable_data());
      return a;
    };
    d["n"] = st.n;
    d["x"] = arr(st.x, 3);
    if (!st.v.empty()) d["v"] = arr(st.v, 3);
    if (!st.kappa.empty()) d["kappa"] = arr(st.kappa, 1);
    if (!st.kappa_v.empty()) d["kappa_v"] = arr(st.kappa_v, 1);
    return d;
  });
  m.def("_write_bgeo", [](const std::string &path, py::array_t<double, py::array::c_style | py::array::forcecast> x,
                           py::array_t<double, py::array::c_style | py::array::forcecast> v,
                           py::array_t<double, py::array::c_style | py::array::forcecast> kappa,
                           py::array_t<double, py::array::c_style | py::array::forcecast> kappa_v) {
    const int64_t n = x.size() / 3;
    if (v.size() != 3 * n || kappa.size() != n || kappa_v.size() != n) throw py::value_error("array sizes");
    write_bgeo(path, n, x.data(), v.data(), kappa.data(), kappa_v.data());
  });
  m.def("_bgeo_of_state_file", &bgeo_of_state_file);

  // scene-side helpers exposed for tests (host logic runs without a GPU)
  m.def("_load_scene_summary", [](const std::string &file, const std::string &param) {
    Scene sc = load_scene(file, param);
    py::dict d;
    d["num_fluid"] = sc.fluid_x.size() / 3;
    d["particle_radius"] = sc.cfg.particle_radius;
    d["target_time"] = sc.cfg.target_time;
    d["surface_tension_method"] = sc.cfg.surface_tension_method;
    d["surface_tension"] = sc.cfg.surface_tension;
    d["gradient_mode"] = sc.cfg.gradient_mode;
    d["max_error"] = sc.cfg.max_error;
    d["use_rigid_contact_solver"] = sc.cfg.use_rigid_contact_solver;
    d["use_release_rigid_body_mode"] = sc.cfg.use_release_rigid_body_mode;
    py::list bodies;
    for (const BodyDesc &b : sc.bodies) {
      py::dict bd;
      bd["num_particles"] = b.samples.size() / 3;
      bd["dynamic"] = b.dynamic;
      bd["density"] = b.density;
      bd["init_v"] = np_vec3(b.init_v);
      bd["init_omega"] = np_vec3(b.init_omega);
      bd["target_x"] = np_vec3(b.target_x);
      bd["translation"] = np_vec3(b.translation);
      bodies.append(bd);
    }
    d["bodies"] = bodies;
    d["num_emitters"] = sc.emitters.size();
    return d;
  }, "scene_file"_a, "param"_a = "");
  // the complete scene as the loader hands it to the C ABI: dfr_config as raw bytes (difffr_b200.cabi.Config), fluid
  // particles, and per body the scaled body-frame samples with pose and initial velocities.  Used by
  // tests/golden/make_paper_golden.py to drive the reference build and the CUDA path with identical inputs.
  m.def("_load_scene_full", [](const std::string &file, const std::string &param) {
    Scene sc = load_scene(file, param);
    py::dict d;
    d["config"] = py::bytes(reinterpret_cast<const char *>(&sc.cfg), sizeof(sc.cfg));
    auto arr = [](const std::vector<double> &v) {
      py::array_t<double> a({(py::ssize_t)(v.size() / 3), (py::ssize_t)3});
      if (!v.empty()) std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(double));
      return a;
    };
    d["fluid_x"] = arr(sc.fluid_x);
    d["fluid_v"] = arr(sc.fluid_v);
    py::list bodies;
    for (const BodyDesc &b : sc.bodies) {
      py::dict bd;
      bd["samples"] = arr(b.samples);
      bd["dynamic"] = b.dynamic;
      bd["density"] = b.density;
      bd["translation"] = np_vec3(b.translation);
      py::array_t<double> q(4);
      for (int k = 0; k < 4; k++) q.mutable_data()[k] = b.rotation[k];
      bd["rotation"] = q;
      bd["init_v"] = np_vec3(b.init_v);
      bd["init_omega"] = np_vec3(b.init_omega);
      bodies.append(bd);
    }
    d["bodies"] = bodies;
    d["num_emitters"] = sc.emitters.size();
    py::list emitters;  // keyword arguments of difffr_b200.cabi.Context.add_emitter
    for (const EmitterDesc &e : sc.emitters) {
      py::dict ed;
      double R[9];
      quat_to_matrix(e.rotation, R);
      py::array_t<double> rot(9);
      std::memcpy(rot.mutable_data(), R, sizeof(R));
      ed["width"] = e.width;
      ed["height"] = e.height;
      ed["position"] = np_vec3(e.x);
      ed["rotation"] = rot;
      ed["velocity"] = e.velocity;
      ed["emit_start"] = e.emit_start;
      ed["emit_end"] = e.emit_end;
      emitters.append(ed);
    }
    d["emitters"] = emitters;
    return d;
  }, "scene_file"_a, "param"_a = "");
}
