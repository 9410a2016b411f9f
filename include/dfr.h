/*
 * dfr.h — C ABI of the B200-native differentiable DFSPH time step.
 *
 * This is the drop-in boundary for the one hot path of zhehaoli1999/DiffFR that this
 * repository accelerates: TimeStepDiffDFSPH::step() + Akinci-2012 rigid coupling +
 * the per-step rigid sensitivity chain rule.  The reference reaches that path through
 * the pybind11 module `pysplishsplash`; each entry point below names the reference
 * interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; all floating point data is FP64 ("double"),
 *     vectors are xyz AoS on the host side (the device layout is SoA);
 *   - quaternions cross the boundary as (w, x, y, z), like
 *     BoundaryModel_Akinci2012::get_quaternion_rb_vec4 (BoundaryModel_Akinci2012.h:176-180);
 *   - small matrices cross row-major (3x3 = 9, 4x3 = 12, 3x4 = 12 doubles);
 *   - every function returns DFR_OK (0) or a negative error code; dfr_last_error()
 *     gives the message.  There is no CPU fallback: without a CUDA device
 *     dfr_create() fails with DFR_ERR_NO_DEVICE.
 *   - a context owns all of its device memory and one CUDA stream; it is not
 *     thread-safe, but distinct contexts are independent (the reference's
 *     process-wide singletons Simulation::current / TimeManager::current,
 *     Simulation.cpp:35,167-188, do not exist here) so many rollouts can share a GPU.
 *
 * The same entry points, prefixed orc_ instead of dfr_, are exported by the CPU
 * oracle (oracle/dfsph_oracle.cpp).  The oracle is test infrastructure only.
 */
#ifndef DFR_H
#define DFR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFR_OK 0
#define DFR_ERR_INVALID -1
#define DFR_ERR_NO_DEVICE -2
#define DFR_ERR_CUDA -3
#define DFR_ERR_STATE -4
#define DFR_ERR_CAPACITY -5

typedef struct dfr_context dfr_context;

/* Scene / solver parameters.  Names follow the reference's scene JSON keys
 * (SPlisHSPlasH/Utilities/SceneLoader.cpp:41-88) and GenericParameters names
 * (TimeStep.cpp:41-64, TimeStepDiffDFSPH.cpp:154-281, Simulation.cpp:227-379). */
typedef struct dfr_config {
  double particle_radius;          /* "particleRadius"; support radius = 4 r (Simulation.cpp:382-386) */
  double density0;                 /* "density0" */
  double gravitation[3];           /* "gravitation" */
  int32_t cfl_method;              /* "cflMethod": 0 none, 1 standard, 2 iteration-aware (Simulation.cpp:524-540) */
  double cfl_factor;               /* "cflFactor" (default 0.5) */
  double cfl_min_time_step;        /* "cflMinTimeStepSize" (default 1e-4) */
  double cfl_max_time_step;        /* "cflMaxTimeStepSize" (default 5e-3) */
  double time_step_size;           /* initial h; reset() restores it (SimulatorBase.cpp:908-910) */
  int32_t min_iterations;          /* TimeStep.cpp:27 (2) */
  int32_t max_iterations;          /* "maxIterations" */
  double max_error;                /* "maxError" in percent */
  int32_t max_iterations_v;        /* "maxIterationsV" */
  double max_error_v;              /* "maxErrorV" in percent */
  int32_t enable_divergence_solver;
  int32_t use_pressure_warmstart;  /* TimeStepDiffDFSPH.cpp:102 */
  int32_t use_divergence_warmstart;/* TimeStepDiffDFSPH.cpp:103 */
  int32_t viscosity_method;        /* 0 none, 1 standard (Viscosity_Standard.cpp:233-334) */
  double viscosity;                /* mu (default 0.01, ViscosityBase.cpp:12) */
  double viscosity_boundary;       /* mu_b (default 0) */
  int32_t surface_tension_method;  /* 0 none, 2 Akinci 2013 (SurfaceTension_Akinci2013.cpp) */
  double surface_tension;          /* k */
  double surface_tension_boundary; /* k_b */
  int32_t gradient_mode;           /* GradientMode: 0 Complete, 1 Incomplete, 2 RigidGradOnly (Simulation.h:173) */
  int32_t rigid_body_mode;         /* RigidBodyMode: 0 WithGyroscopic, 1 NoGyroscopic (Simulation.h:174) */
  int32_t optimize_rotation;       /* "optimize rotation" (TimeStepDiffDFSPH.cpp:108) */
  int32_t use_rigid_gradient_manager; /* "useRigidGradientManager" */
  int32_t use_rigid_contact_solver;   /* "useRigidContactSolver" (penalty branch, RigidContactSolver.cpp:419-555) */
  double rigid_contact_beta;          /* "rigidContactBeta" */
  double rigid_contact_gamma;         /* "rigidContactGamma" */
  double rigid_contact_friction;      /* "rigidContactFrictionCoeff" */
  double rigid_contact_support_radius_factor; /* "rigidContactSupportRadiusFactor" */
  double target_time;              /* "targetTime" */
  double uniform_acc_rb_time;      /* "uniformAccelerateRBTime" */
  int32_t max_emitted_particles;   /* capacity reserved for emitters (Emitter.cpp) */
  /* Device-side tuning (0 = default).  These have no counterpart in the reference; they never change results.
   * The three capacities are INITIAL values: rows that do not fit make the context grow them and rebuild the lists
   * before the step continues (the reference's lists are unbounded); DFR_ERR_CAPACITY is only left for a row that
   * more than doubles within one step outside the first steps after finalize / reset / load. */
  int32_t neighbor_capacity_fluid;    /* fluid neighbours stored per fluid particle (default 96) */
  int32_t neighbor_capacity_boundary; /* boundary neighbours stored per fluid particle (default 64) */
  int32_t body_neighbor_capacity;     /* mean fluid neighbours stored per dynamic boundary particle (default 96) */
  int32_t grid_reach;                 /* cell edge = support radius / grid_reach, stencil (2 reach + 1)^3 (default 2) */
  /* "useReleaseRigidBodyMode" (billiards-on-water scenes; TimeStepDiffDFSPH.cpp:381-407): the velocity ramp is off; body 1
   * is held at rest (animated) until uniformAccelerateRBTime has passed and from then on STARTS every step with its
   * initial velocities; all other dynamic bodies move freely from t = 0. */
  int32_t use_release_rigid_body_mode;
  int32_t reserved_i[2];
  double reserved_d[8];
} dfr_config;

/* Fill cfg with the reference's defaults (Simulation.cpp:100-130, TimeStep.cpp:22-29,
 * TimeStepDiffDFSPH.cpp:90-112, SceneLoader.cpp:41-88). */
void dfr_default_config(dfr_config *cfg);

/* Replaces Simulation::init + TimeStepDiffDFSPH ctor (Simulation.cpp:382-386, TimeStepDiffDFSPH.cpp:90-134). */
int dfr_create(const dfr_config *cfg, int device, dfr_context **out);
void dfr_destroy(dfr_context *ctx);
const char *dfr_last_error(const dfr_context *ctx);

/* Replaces Simulation::addFluidModel / FluidModel::initModel (Simulation.cpp:814-819).
 * x, v: n*3 doubles.  Particle ids are 0..n-1 in the order given. */
int dfr_set_fluid(dfr_context *ctx, int64_t n, const double *x, const double *v);

/* Replaces RigidBody3dBoundarySimulator::initBoundaryData -> BoundaryModel_Akinci2012::initModel
 * + Dynamic3dRigidBody::determineMassProperties (RigidBody3dBoundarySimulator.cpp:202-214,
 * BoundaryModel_Akinci2012.cpp:286-386, Dynamic3dRigidBody.h:184-219).
 * x_local: n*3 body-frame sample positions (already scaled). Returns the body index (>= 0). */
int dfr_add_body(dfr_context *ctx, int64_t n, const double *x_local, int is_dynamic,
                 double density, const double position[3], const double quat_wxyz[4]);

/* Per-body targets / initial velocities: SimulationDataDiffDFSPH get_init_v_rb / get_init_omega_rb
 * (DiffDFSPHModule.cpp:60-75; TimeStepDiffDFSPH.cpp:2087-2095).  After dfr_finalize the values are staged and reach the
 * device in stream order at the start of the next dfr_step / dfr_run_trajectory (or with dfr_reset): the call itself
 * never synchronises, so it can be made every step (controller experiments). */
int dfr_set_init_v_omega(dfr_context *ctx, int body, const double v0[3], const double omega0[3]);

/* Slab domain decomposition of one scene over the GPUs of a node (nothing comparable in the reference,
 * which runs one OpenMP process).  Every rank builds the SAME scene (same dfr_set_fluid / dfr_add_body calls), then
 * calls dfr_slab_configure before dfr_finalize: the z cell layers are cut into n_ranks ranges of equal particle count,
 * the context keeps its range plus one support radius of ghost particles, and dfr_step exchanges boundary layers,
 * residuals, the CFL maximum and the per-body force/torque/Jacobian rows over NVLink on the context's stream (peer
 * stores into cudaIpc-mapped neighbour memory; NCCL for setup, the steps after a reset, and as fallback).
 * id_bytes comes from dfr_slab_unique_id on rank 0 and is distributed by the caller (torch.distributed, MPI, a file).
 * Results equal the single-context run up to summation order.  Parity dumps (dfr_download_fluid) write only the ids a
 * rank owns; rigid bodies are replicated.  Not available with emitters or the rigid contact solver. */
#define DFR_SLAB_ID_BYTES 128
/* The cut dfr_finalize uses, callable on its own (host only, no device needed): z is the z coordinate of every fluid
 * particle, cell layer = floor((z - z_origin) * inv_cell) clamped to [0, nz).  planes gets n_ranks + 1 layer indices,
 * rank r owns layers [planes[r], planes[r+1]).  DFR_ERR_INVALID if a slab would be thinner than 2 * reach + 1 layers. */
int dfr_slab_plan(double z_origin, double inv_cell, int nz, int reach, int64_t n, const double *z, int n_ranks, int32_t *planes);
int dfr_slab_unique_id(char out[DFR_SLAB_ID_BYTES]);
int dfr_slab_configure(dfr_context *ctx, int rank, int n_ranks, const char id_bytes[DFR_SLAB_ID_BYTES]);
/* out = { fluid particles owned, ghost particles held, bytes exchanged over NVLink since reset, number of slabs -
 * negated when the ghost updates travel as stores from the producing kernels into cudaIpc-mapped neighbour memory
 * (default; DFR_SLAB_TRANSPORT=nccl or a failed mapping selects NCCL send/recv) } */
int dfr_slab_info(dfr_context *ctx, int64_t out[4]);

/* Ends scene construction: uploads everything, computes the Akinci boundary volumes
 * (Simulation::updateBoundaryVolume, Simulation.cpp:831-902) and snapshots the initial state in
 * HBM so that dfr_reset() is a device-to-device copy (replaces SimulatorBase::reset's re-parse,
 * SimulatorBase.cpp:887-934). */
int dfr_finalize(dfr_context *ctx);

/* Overwrite the fluid state after finalize (--load-fluid-pos[-and-vel] / state files,
 * SimulatorBase.cpp:2023-2058, 2576-2604).  Any pointer may be NULL (= keep). The new state also
 * becomes the snapshot dfr_reset() restores, as checkLoadState re-applies it on every reset. */
int dfr_load_fluid_state(dfr_context *ctx, const double *x, const double *v,
                         const double *kappa, const double *kappa_v);

/* The same for the rows THIS context holds (slab-decomposed contexts; on a single context identical to
 * dfr_load_fluid_state): dfr_slab_local_ids gives the particle ids (rows of the scene's arrays) the context held at t = 0,
 * n of them (returned; ids_out may be NULL to ask for n only), and dfr_load_fluid_state_local takes arrays of exactly those
 * rows in that order - a rank of a distributed job uploads its share instead of gathering from whole-scene arrays. */
int64_t dfr_slab_local_ids(dfr_context *ctx, int32_t *ids_out, int64_t capacity);
int dfr_load_fluid_state_local(dfr_context *ctx, int64_t n, const double *x, const double *v, const double *kappa,
                               const double *kappa_v);

/* SimulatorBase::reset (SimulatorBase.cpp:887-934). */
int dfr_reset(dfr_context *ctx);

/* TimeStepDiffDFSPH::reset_gradient -> BoundaryModel_Akinci2012::reset_gradient for every dynamic body
 * (TimeStepDiffDFSPH.cpp:2234-2240, BoundaryModel_Akinci2012.cpp:62-108): the sensitivities restart from
 * d(v,omega)/d(v0,omega0) = I, everything else zero, at the current state (short-horizon restarts,
 * cartpole-diff-controller.py:228).  As in the reference, the RigidBodyGradientManager blocks are left alone. */
int dfr_reset_gradient(dfr_context *ctx);

/* Simulation::setGradientMode (SimulationModule.cpp:206; Simulation.h:173): 0 Complete, 1 Incomplete,
 * 2 RigidGradOnly.  Legal at any time between steps. */
int dfr_set_gradient_mode(dfr_context *ctx, int mode);

/* n x SimulatorBase::timeStepNoGUI body (SimulatorBase.cpp:1142-1169): TimeStepDiffDFSPH::step,
 * gradient-manager stages, rigid velocity/position update.
 * Steady state (single context): every step is one replay of a recorded CUDA graph whose two Jacobi loops are
 * conditional WHILE nodes - the stopping rules of TimeStepDiffDFSPH.cpp:711-743 / 828-861 run on the device - so there
 * is no host round trip inside the n steps; the call ends with ONE synchronisation that brings back the status word and
 * the body records the getters read.  The stream path (speculated batches of iterations, one read-back per solve) is
 * used instead: in the first four steps after finalize / reset / load (the neighbour-list capacities are being
 * watched), with the rigid contact solver once every 500 steps (the reference's z-sort of the contact order runs on the
 * host), while per-kernel profiling is on, and with DFR_NO_GRAPH=1.
 * Slab-decomposed contexts (peer-memory transport) step the same way: one graph replay per step, nothing read back in
 * between.  The particle exchange at the head of the step stores the export layers straight into the neighbours'
 * receive areas and orders them with two flag passes; residual sums, the CFL maximum and the per-body rows are
 * all-reduced by small kernels over peer-mapped mailboxes (no collective call, no host read-back inside a step).
 * DFR_SLAB_HOST_EXCHANGE=1 keeps the NCCL particle exchange with its two read-backs at the head of every step,
 * DFR_SLAB_TRANSPORT=nccl the whole round-1 stream path (the variables must be set on all ranks alike). */
int dfr_step(dfr_context *ctx, int n_steps);

/* Runs steps until TimeStepDiffDFSPH::is_trajectory_finish_callback() would be true
 * (TimeStepDiffDFSPH.cpp:448) or max_steps is hit; writes the number of steps taken.
 * Replayed steps are enqueued in batches of 16 (DFR_TRAJECTORY_BATCH) with one state read-back per batch; steps of a
 * batch that follow the end of the trajectory are skipped on the device (an IF node around the step). */
int dfr_run_trajectory(dfr_context *ctx, int max_steps, int *steps_done);

/* Simulation time data: TimeManager::getTime/getTimeStepSize (TimeModule.cpp:23-29),
 * TimeStep iterations (TimeStep.cpp:41-47), get_step_count, is_trajectory_finish_callback. */
typedef struct dfr_step_info {
  double time;
  double time_step_size;
  int32_t iterations;      /* pressure solver, last step */
  int32_t iterations_v;    /* divergence solver, last step */
  int32_t step_count;
  int32_t trajectory_finished;
  int64_t num_fluid_particles;      /* active */
  int64_t total_pressure_iterations;  /* since reset */
  int64_t total_divergence_iterations;
  int64_t total_particle_steps;     /* sum over steps of active fluid particles */
  int64_t total_fluid_neighbors;    /* sum over steps of all stored neighbour entries */
} dfr_step_info;
int dfr_get_step_info(dfr_context *ctx, dfr_step_info *info);

/* Rigid body state: out = x[3], q(w,x,y,z)[4], v[3], omega[3]  (BoundaryModelModule.cpp:33-48). */
int dfr_get_body_state(dfr_context *ctx, int body, double out[13]);
int dfr_set_body_velocity(dfr_context *ctx, int body, const double v[3], const double omega[3]);
/* mass, inverse mass, inertia0 (9), force (3), torque (3) of the last step (getForce/getTorque). */
int dfr_get_body_properties(dfr_context *ctx, int body, double out[17]);

/* Per-body sensitivities (BoundaryModelModule.cpp:60-80):
 *   which = 0 grad_x_to_v0 (3x3)      1 grad_x_to_omega0 (3x3)
 *           2 grad_quaternion_to_v0 (4x3)  3 grad_quaternion_to_omega0 (4x3)
 *           4 grad_v_to_v0            5 grad_v_to_omega0
 *           6 grad_omega_to_v0        7 grad_omega_to_omega0
 * and the per-step net Jacobians (BoundaryModel_Akinci2012.h:52-61):
 *           8 grad_net_force_to_vn    9 grad_net_force_to_xn   10 grad_net_force_to_qn (3x4)
 *          11 grad_net_force_to_omega_n  12 grad_net_torque_to_vn 13 grad_net_torque_to_xn
 *          14 grad_net_torque_to_qn (3x4) 15 grad_net_torque_to_omega_n
 * out must hold 12 doubles. */
int dfr_get_body_grad(dfr_context *ctx, int body, int which, double out[12]);

/* RigidBodyGradientManager getters (SimulationModule.cpp:406-414): same `which` numbering,
 * for the block (R, RR). */
int dfr_get_manager_grad(dfr_context *ctx, int R, int RR, int which, double out[12]);

/* Parity dumps, keyed by particle id.
 *   field: 0 position (3) 1 velocity (3) 2 density 3 factor 4 kappa 5 kappa_v
 *          6 density_adv 7 acceleration (3) 8 sum_grad_p_k (3) 9 normal (3)  */
int dfr_download_fluid(dfr_context *ctx, int field, double *out);
/*   field: 0 position (3) 1 velocity (3) 2 volume 3 position0 (3); n = particles of that body */
int dfr_download_body(dfr_context *ctx, int body, int field, double *out);
int64_t dfr_num_fluid(dfr_context *ctx);
/* Fluid particles of the scene as given to dfr_set_fluid (FluidModel::numParticles() before any emission): the length
 * (in particles) dfr_load_fluid_state expects of its arrays. */
int64_t dfr_num_fluid_initial(dfr_context *ctx);
int64_t dfr_num_body_particles(dfr_context *ctx, int body);
int dfr_num_bodies(dfr_context *ctx);

/* Neighbour sets of the current positions (replaces CompactNSearch find_neighbors + the accessors
 * Simulation::numberOfNeighbors/getNeighbor, Simulation.h:520-543), in particle-id space:
 *   set_a = -1 fluid, or a body index; set_b likewise.
 * First call with indices == NULL to get counts (n_a entries) and the total; then with a buffer.
 * Neighbours of each point are returned in ascending id order. */
int dfr_get_neighbors(dfr_context *ctx, int set_a, int set_b, int32_t *counts,
                      int32_t *indices, int64_t indices_capacity, int64_t *total);

/* Box emitter (Emitter.cpp:89-227; scene keys SceneLoader.cpp:362-415). */
int dfr_add_emitter(dfr_context *ctx, int width, int height, const double position[3],
                    const double rot_matrix_rowmajor[9], double velocity,
                    double emit_start, double emit_end);

/* Timing of the device work of the steps run since the last reset, in milliseconds
 * (CUDA events on the context's stream). Replaces Utilities::Timing averages (Timing.h:22-41). */
int dfr_get_device_time_ms(dfr_context *ctx, double *total_ms, int64_t *kernel_launches);

/* Per-kernel device timing (replaces the START_TIMING/STOP_TIMING_AVG pairs around every phase of
 * TimeStepDiffDFSPH::step, TimeStepDiffDFSPH.cpp:528-651, and Timing::printAverageTimes, Timing.h:168-196).
 * While enabled every kernel launch is bracketed by CUDA events on the context's stream; rows are keyed by
 * kernel name.  Enabling or disabling clears the table.  dfr_get_kernel_profile returns DFR_ERR_INVALID
 * once index is past the last row. */
int dfr_set_profiling(dfr_context *ctx, int enable);
int dfr_get_kernel_profile(dfr_context *ctx, int index, char *name, int name_capacity,
                           double *total_ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* DFR_H */
