// Host-side scene description for the DiffDFSPH path: what the reference assembles in
// Utilities::SceneLoader::readScene (SPlisHSPlasH/Utilities/SceneLoader.cpp:22-513),
// SimulatorBase::createFluidBlocks (Simulator/SimulatorBase.cpp:1638-1735) and
// RigidBody3dBoundarySimulator::initBoundaryData (Simulator/RigidBody3dBoundarySimulator.cpp:60-214),
// reduced to the inputs include/dfr.h consumes.  Plain C++17, no CUDA, no Python.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dfr.h"
#include "json_min.hpp"

namespace dfrhost {

using Vec3 = std::array<double, 3>;
using Quat = std::array<double, 4>;  // (w, x, y, z)

inline Quat quat_from_axis_angle(const double axis[3], double angle) {
  const double n = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  if (n == 0.0) return {1, 0, 0, 0};
  const double s = std::sin(0.5 * angle) / n;
  return {std::cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s};
}
inline void quat_to_matrix(const Quat &q, double R[9]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
}
// TimeStepDiffDFSPH::get_target_quaternion_vec4 (TimeStepDiffDFSPH.cpp:2118-2131): Euler angles in degrees,
// q = Rx(a0) * Ry(a1) * Rz(a2)
inline Quat quat_from_euler_deg(const Vec3 &deg) {
  const double k = M_PI / 180.0;
  auto mul = [](const Quat &a, const Quat &b) {
    return Quat{a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  };
  const double ax[3] = {1, 0, 0}, ay[3] = {0, 1, 0}, az[3] = {0, 0, 1};
  return mul(mul(quat_from_axis_angle(ax, deg[0] * k), quat_from_axis_angle(ay, deg[1] * k)), quat_from_axis_angle(az, deg[2] * k));
}

struct BodyDesc {  // SceneLoader::BoundaryData (SceneLoader.h:33-67), the fields this path reads
  std::string mesh_file, samples_file;
  Vec3 translation{0, 0, 0}, scale{1, 1, 1};
  Quat rotation{1, 0, 0, 0};
  double density = 1000.0;
  bool dynamic = false, is_wall = false;
  Vec3 init_v{0, 0, 0}, init_omega{0, 0, 0}, target_x{0, 0, 0}, target_angle_deg{0, 0, 0};
  std::vector<double> samples;  // n*3, body frame, scaled
};
struct EmitterDesc {  // SceneLoader::EmitterData (SceneLoader.h:121-134)
  int width = 5, height = 5, type = 0;
  Vec3 x{0, 0, 0};
  Quat rotation{1, 0, 0, 0};
  double velocity = 1.0, emit_start = 0.0, emit_end = 1e300;
};
struct Scene {
  dfr_config cfg;
  bool use_release_rigid_body_mode = false;
  std::vector<double> fluid_x, fluid_v;  // n*3
  std::vector<BodyDesc> bodies;
  std::vector<EmitterDesc> emitters;
  std::string scene_dir;
};

// ---- triangle meshes -------------------------------------------------------------------------
struct Mesh {
  std::vector<Vec3> v;
  std::vector<std::array<int, 3>> f;
};
// Wavefront OBJ: positions and faces only (SimulatorBase::loadObj, SimulatorBase.cpp:1253-1283); polygons are fanned
inline Mesh load_obj(const std::string &path, const Vec3 &scale) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("cannot open mesh file " + path);
  Mesh m;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string tag;
    ss >> tag;
    if (tag == "v") {
      Vec3 p;
      ss >> p[0] >> p[1] >> p[2];
      m.v.push_back({p[0] * scale[0], p[1] * scale[1], p[2] * scale[2]});
    } else if (tag == "f") {
      std::vector<int> idx;
      std::string tok;
      while (ss >> tok) {
        const int i = std::atoi(tok.substr(0, tok.find('/')).c_str());
        idx.push_back(i > 0 ? i - 1 : (int)m.v.size() + i);
      }
      for (size_t k = 1; k + 1 < idx.size(); k++) m.f.push_back({idx[0], idx[k], idx[k + 1]});
    }
  }
  return m;
}

// Sample spacing in particle radii.  The reference's Poisson-disk sampler is called with minimum distance r
// (RigidBody3dBoundarySimulator.cpp:143) and ends up with 159,496 / 810 samples on the stone-skipping tank / stone
// (SURVEY §6); a lattice spacing of 1.35 r reproduces those densities to within ~10 % (156,975 / 593 here).
constexpr double kSampleSpacingInRadii = 1.35;

// Deterministic surface sampling with a given spacing.  Stands in for the reference's Poisson-disk sampler
// (Utilities/PoissonDiskSampling.cpp, random: 159,496 vs 159,284 samples for the same scene on two hosts)
// and plays the role of RegularTriangleSampling (samplingMode 1): each triangle is covered by a barycentric lattice,
// then samples closer than 0.85 * spacing to an accepted one are dropped (hash grid, first come first kept).
// Interior lattice points are displaced inside the triangle's plane by a deterministic pseudo-random offset of up to
// 0.2 spacings (a hash of the point's running index; corners and edges stay put so that neighbouring triangles still
// de-duplicate): like the Poisson-disk samples it stands in for, the result has no exact symmetries.  That matters for the
// penalty contact solver, whose friction Jacobian divides by |x_r - centroid of r's own-body neighbours|
// (RigidContactSolver.cpp:483-505) - exactly zero for a particle in the middle of a regularly sampled flat face, NaN in
// the reference's own code as well, and never zero with its random samples.
inline std::vector<double> sample_mesh(const Mesh &m, double spacing) {
  std::vector<Vec3> cand;
  auto sub = [](const Vec3 &a, const Vec3 &b) { return Vec3{a[0] - b[0], a[1] - b[1], a[2] - b[2]}; };
  auto len = [](const Vec3 &a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };
  auto unit01 = [](uint64_t k) {  // splitmix64 -> [0, 1)
    k += 0x9e3779b97f4a7c15ull;
    k = (k ^ (k >> 30)) * 0xbf58476d1ce4e5b9ull;
    k = (k ^ (k >> 27)) * 0x94d049bb133111ebull;
    k ^= k >> 31;
    return (double)(k >> 11) * (1.0 / 9007199254740992.0);
  };
  uint64_t serial = 0;
  for (const auto &t : m.f) {
    const Vec3 &A = m.v[t[0]], &B = m.v[t[1]], &C = m.v[t[2]];
    const double lmax = std::max(len(sub(B, A)), std::max(len(sub(C, A)), len(sub(C, B))));
    const int n = std::max(1, (int)std::ceil(lmax / spacing));
    for (int i = 0; i <= n; i++)
      for (int j = 0; j <= n - i; j++) {
        double u = (double)i / n, w = (double)j / n;
        serial++;
        if (i > 0 && j > 0 && i + j < n) {  // interior point: jitter in barycentric coordinates (stays inside the triangle)
          const double amp = 0.2 / n;
          u += amp * (2.0 * unit01(2 * serial) - 1.0);
          w += amp * (2.0 * unit01(2 * serial + 1) - 1.0);
        }
        const double s = 1.0 - u - w;
        cand.push_back({s * A[0] + u * B[0] + w * C[0], s * A[1] + u * B[1] + w * C[1], s * A[2] + u * B[2] + w * C[2]});
      }
  }
  const double cell = spacing, dmin2 = 0.85 * 0.85 * spacing * spacing;
  std::unordered_map<uint64_t, std::vector<int>> grid;
  auto key = [&](int x, int y, int z) {
    return ((uint64_t)(uint32_t)(x + (1 << 20)) << 42) ^ ((uint64_t)(uint32_t)(y + (1 << 20)) << 21) ^ (uint64_t)(uint32_t)(z + (1 << 20));
  };
  std::vector<Vec3> kept;
  for (const Vec3 &p : cand) {
    const int cx = (int)std::floor(p[0] / cell), cy = (int)std::floor(p[1] / cell), cz = (int)std::floor(p[2] / cell);
    bool ok = true;
    for (int dz = -1; dz <= 1 && ok; dz++)
      for (int dy = -1; dy <= 1 && ok; dy++)
        for (int dx = -1; dx <= 1 && ok; dx++) {
          auto it = grid.find(key(cx + dx, cy + dy, cz + dz));
          if (it == grid.end()) continue;
          for (int k : it->second) {
            const Vec3 d = sub(kept[k], p);
            if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] < dmin2) { ok = false; break; }
          }
        }
    if (!ok) continue;
    grid[key(cx, cy, cz)].push_back((int)kept.size());
    kept.push_back(p);
  }
  std::vector<double> out;
  out.reserve(kept.size() * 3);
  for (const Vec3 &p : kept) { out.push_back(p[0]); out.push_back(p[1]); out.push_back(p[2]); }
  return out;
}

// Raw sample files: "<name>.xyz" text (x y z per line) or little-endian binary "<name>.f64" (n*3 doubles).  The
// reference reads partio .bgeo here (PartioReaderWriter::readParticles); freezing its samples into one of these two
// formats gives both sides identical boundary particles.
inline std::vector<double> load_samples(const std::string &path, const Vec3 &scale) {
  std::vector<double> out;
  if (path.size() > 4 && path.substr(path.size() - 4) == ".f64") {
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) throw std::runtime_error("cannot open samples file " + path);
    const size_t bytes = (size_t)in.tellg();
    in.seekg(0);
    out.resize(bytes / sizeof(double));
    in.read((char *)out.data(), (std::streamsize)(out.size() * sizeof(double)));
  } else {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open samples file " + path);
    double v;
    while (in >> v) out.push_back(v);
  }
  if (out.size() % 3) throw std::runtime_error("samples file " + path + ": length is not a multiple of 3");
  for (size_t i = 0; i < out.size(); i++) out[i] *= scale[i % 3];
  return out;
}

// createFluidBlocks, denseMode 0 (SimulatorBase.cpp:1638-1735): steps = round(extent / 2r) - 1 per axis,
// first particle at min + 2r, x outermost and z innermost.
inline void add_fluid_block(Scene &sc, const double lo[3], const double hi[3], const double vel[3]) {
  const double diam = 2.0 * sc.cfg.particle_radius;
  int steps[3];
  for (int k = 0; k < 3; k++) steps[k] = std::max(0, (int)std::round((hi[k] - lo[k]) / diam) - 1);
  for (int j = 0; j < steps[0]; j++)
    for (int k = 0; k < steps[1]; k++)
      for (int l = 0; l < steps[2]; l++) {
        sc.fluid_x.push_back(j * diam + (lo[0] + diam));
        sc.fluid_x.push_back(k * diam + (lo[1] + diam));
        sc.fluid_x.push_back(l * diam + (lo[2] + diam));
        for (int c = 0; c < 3; c++) sc.fluid_v.push_back(vel[c]);
      }
}

inline std::string dir_of(const std::string &path) {
  const size_t p = path.find_last_of('/');
  return p == std::string::npos ? std::string(".") : path.substr(0, p);
}

// SceneLoader::readScene + readParameterObject for the keys of this path.  `param` is the reference's
// "--param" syntax (SimulatorBase.cpp:734-839): "<name>:<value>" or "<fluid-id>:<name>:<value>", comma separated.
inline Scene load_scene(const std::string &scene_file, const std::string &param = "", bool sample_bodies = true) {
  std::ifstream in(scene_file);
  if (!in) throw std::runtime_error("cannot open scene file " + scene_file);
  std::stringstream buf;
  buf << in.rdbuf();
  const std::string text = buf.str();
  Json root = JsonParser(text).parse();
  Scene sc;
  sc.scene_dir = dir_of(scene_file);
  dfr_config &c = sc.cfg;
  dfr_default_config(&c);
  if (const Json *cfg = root.find("Configuration")) {
    cfg->read("particleRadius", c.particle_radius);
    cfg->read("timeStepSize", c.time_step_size);
    cfg->read_vec("gravitation", c.gravitation);
    cfg->read("cflMethod", c.cfl_method);
    cfg->read("cflFactor", c.cfl_factor);
    cfg->read("cflMinTimeStepSize", c.cfl_min_time_step);
    cfg->read("cflMaxTimeStepSize", c.cfl_max_time_step);
    cfg->read("maxIterations", c.max_iterations);
    cfg->read("maxError", c.max_error);
    cfg->read("maxIterationsV", c.max_iterations_v);
    cfg->read("maxErrorV", c.max_error_v);
    bool b;
    if (cfg->read("enableDivergenceSolver", b)) c.enable_divergence_solver = b;
    if (cfg->read("useRigidContactSolver", b)) c.use_rigid_contact_solver = b;
    if (cfg->read("useRigidGradientManager", b)) c.use_rigid_gradient_manager = b;
    cfg->read("useReleaseRigidBodyMode", sc.use_release_rigid_body_mode);
    c.use_release_rigid_body_mode = sc.use_release_rigid_body_mode ? 1 : 0;
    cfg->read("rigidContactGamma", c.rigid_contact_gamma);
    cfg->read("rigidContactBeta", c.rigid_contact_beta);
    cfg->read("rigidContactSupportRadiusFactor", c.rigid_contact_support_radius_factor);
    cfg->read("rigidContactFrictionCoeff", c.rigid_contact_friction);
    cfg->read("targetTime", c.target_time);
    cfg->read("uniformAccelerateRBTime", c.uniform_acc_rb_time);
    cfg->read("gradientMode", c.gradient_mode);
    int sim_method = 5;
    if (cfg->read("simulationMethod", sim_method) && sim_method != 5 && sim_method != 4)
      throw std::runtime_error("only the DiffDFSPH / DFSPH time step (simulationMethod 5) is provided by this module");
    int bh = 0;
    if (cfg->read("boundaryHandlingMethod", bh) && bh != 0)
      throw std::runtime_error("only Akinci2012 boundary handling (boundaryHandlingMethod 0) is provided by this module");
    // 2-D scenes use 2-D kernels, particle volumes and the 7-neighbour deficiency cut (Simulation.cpp:382-386,
    // TimeStepDiffDFSPH.cpp:2027-2040): not on this path, and running them as 3-D would be silently wrong
    bool sim2d = false;
    if (cfg->read("sim2D", sim2d) && sim2d) throw std::runtime_error("sim2D scenes are outside this path (3-D DiffDFSPH only)");
  }
  // articulated systems / joints (ArticulatedSystemSimulator.cpp) and particle-file fluids are not on this path
  if (const Json *as = root.find("ArticulatedSystems"))
    if (!as->arr.empty()) throw std::runtime_error("ArticulatedSystems (joints) are outside this path");
  if (const Json *js = root.find("Joints"))
    if (!js->arr.empty()) throw std::runtime_error("Joints are outside this path");
  if (const Json *fm = root.find("FluidModels"))
    if (!fm->arr.empty()) throw std::runtime_error("FluidModels (particle-file fluids) are outside this path; use FluidBlocks or a state file");
  if (const Json *mats = root.find("Materials"))
    for (const Json &m : mats->arr) {
      std::string id = "Fluid";
      m.read("id", id);
      if (id != "Fluid") continue;  // one fluid phase on this path
      m.read("density0", c.density0);
      m.read("viscosity", c.viscosity);
      m.read("viscosityMethod", c.viscosity_method);
      m.read("viscosityBoundary", c.viscosity_boundary);
      m.read("surfaceTension", c.surface_tension);
      m.read("surfaceTensionMethod", c.surface_tension_method);
      m.read("surfaceTensionBoundary", c.surface_tension_boundary);
      m.read("maxEmitterParticles", c.max_emitted_particles);
    }
  // --param overrides
  {
    std::stringstream ps(param);
    std::string item;
    while (std::getline(ps, item, ',')) {
      if (item.empty()) continue;
      std::vector<std::string> tok;
      std::stringstream is(item);
      std::string t;
      while (std::getline(is, t, ':')) tok.push_back(t);
      if (tok.size() < 2) continue;
      const std::string &name = tok[tok.size() - 2];
      const double v = std::atof(tok.back().c_str());
      if (name == "cflMethod") c.cfl_method = (int)v;
      else if (name == "cflFactor") c.cfl_factor = v;
      else if (name == "cflMaxTimeStepSize") c.cfl_max_time_step = v;
      else if (name == "cflMinTimeStepSize") c.cfl_min_time_step = v;
      else if (name == "maxIterations") c.max_iterations = (int)v;
      else if (name == "maxError") c.max_error = v;
      else if (name == "maxIterationsV") c.max_iterations_v = (int)v;
      else if (name == "maxErrorV") c.max_error_v = v;
      else if (name == "viscosity") c.viscosity = v;
      else if (name == "surfaceTension") c.surface_tension = v;
      else if (name == "gradientMode") c.gradient_mode = (int)v;
      else if (name == "targetTime") c.target_time = v;
      else throw std::runtime_error("--param: unknown parameter '" + name + "'");
    }
  }
  if (c.viscosity_method != 0 && c.viscosity_method != 1)
    throw std::runtime_error("viscosityMethod " + std::to_string(c.viscosity_method) + " is outside this path (0 none, 1 standard)");
  if (c.surface_tension_method != 0 && c.surface_tension_method != 2)
    throw std::runtime_error("surfaceTensionMethod " + std::to_string(c.surface_tension_method) + " is outside this path (0 none, 2 Akinci2013)");

  if (const Json *rbs = root.find("RigidBodies"))
    for (const Json &rb : rbs->arr) {
      BodyDesc b;
      const bool has_mesh = rb.read("geometryFile", b.mesh_file);
      const bool has_samples = rb.read("particleFile", b.samples_file);
      if (!has_mesh && !has_samples) continue;  // SceneLoader.cpp:105
      double tr[3] = {0, 0, 0}, ax[3] = {0, 0, 0}, s3[3] = {1, 1, 1}, ang = 0.0;
      rb.read_vec("translation", tr);
      b.translation = {tr[0], tr[1], tr[2]};
      if (rb.read_vec("rotationAxis", ax) && rb.read("rotationAngle", ang)) b.rotation = quat_from_axis_angle(ax, ang);
      rb.read("density", b.density);
      rb.read_vec("scale", s3);
      b.scale = {s3[0], s3[1], s3[2]};
      rb.read("isDynamic", b.dynamic);
      rb.read("isWall", b.is_wall);
      bool animated = false;
      if (rb.read("isAnimated", animated) && animated)
        throw std::runtime_error("animated rigid bodies (isAnimated) are outside this path");
      double v3[3];
      if (rb.read_vec("targetX", v3)) b.target_x = {v3[0], v3[1], v3[2]};
      if (rb.read_vec("targetAngleInDegree", v3)) b.target_angle_deg = {v3[0], v3[1], v3[2]};
      if (rb.read_vec("initVelocity", v3)) b.init_v = {v3[0], v3[1], v3[2]};
      if (rb.read_vec("initAngularVelocity", v3)) b.init_omega = {v3[0], v3[1], v3[2]};
      if (sample_bodies) {
        if (has_samples)
          b.samples = load_samples(sc.scene_dir + "/" + b.samples_file, b.scale);
        else
          b.samples = sample_mesh(load_obj(sc.scene_dir + "/" + b.mesh_file, b.scale), kSampleSpacingInRadii * c.particle_radius);
      }
      sc.bodies.push_back(std::move(b));
    }
  if (const Json *blocks = root.find("FluidBlocks"))
    for (const Json &fb : blocks->arr) {
      double tr[3] = {0, 0, 0}, s3[3] = {1, 1, 1}, lo[3], hi[3], vel[3] = {0, 0, 0};
      fb.read_vec("translation", tr);
      fb.read_vec("scale", s3);
      if (!(fb.read_vec("start", lo) && fb.read_vec("end", hi))) continue;
      int mode = 0;
      if (fb.read("denseMode", mode) && mode != 0) throw std::runtime_error("FluidBlocks denseMode 1/2 are outside this path");
      fb.read_vec("initialVelocity", vel);
      for (int k = 0; k < 3; k++) { lo[k] = s3[k] * lo[k] + tr[k]; hi[k] = s3[k] * hi[k] + tr[k]; }
      add_fluid_block(sc, lo, hi, vel);
    }
  if (const Json *ems = root.find("Emitters"))
    for (const Json &e : ems->arr) {
      EmitterDesc d;
      e.read("width", d.width);
      e.read("height", d.height);
      double tr[3] = {0, 0, 0}, ax[3] = {0, 0, 0}, ang = 0.0;
      if (e.read_vec("translation", tr)) d.x = {tr[0], tr[1], tr[2]};
      if (e.read_vec("rotationAxis", ax) && e.read("rotationAngle", ang)) d.rotation = quat_from_axis_angle(ax, ang);
      e.read("velocity", d.velocity);
      e.read("emitStartTime", d.emit_start);
      e.read("emitEndTime", d.emit_end);
      e.read("type", d.type);
      if (d.type != 0) throw std::runtime_error("only box emitters (type 0) are on this path");
      sc.emitters.push_back(d);
    }
  if (!sc.emitters.empty() && c.max_emitted_particles == 0) c.max_emitted_particles = 10000;  // FluidModel default (FluidModel.cpp:38)
  return sc;
}

}  // namespace dfrhost
