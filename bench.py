#!/usr/bin/env python
"""bench.py — forward + sensitivity ("adjoint") DiffDFSPH particle-steps/s on B200.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W            -> one JSON line, this repository's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  -> one JSON line, the reference's CPU algorithm
                                                              (oracle/_ref when built, else the oracle port)
A "step" is one SimulatorBase::timeStepNoGUI body (TimeStepDiffDFSPH::step + sensitivity chain rule + rigid
update) over the synthetic dam break with 4 dynamic rigid boxes and 2^20 fluid particles (BASELINE.json
configs[4], the configuration the north-star's throughput target is quoted on).  For N > 1 (weak scaling, 2^20 particles
per GPU) there are two shardings, `--mode`:
  slab      (default) ONE scene of N x 2^20 particles, slab-decomposed over the N GPUs: boundary-layer particles, ghost
            updates of every gathered array, residuals, the CFL maximum and the rigid force/torque/Jacobian rows travel
            over NVLink (NCCL) inside every step (BASELINE.json configs[4])
  rollouts  a population of 64 independent rollouts at the high-diving scene's size, sharded over the N GPUs with several
            contexts side by side per GPU (population sharding, configs[3]); no data-path collective; also valid at N = 1
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PARTICLES = 1 << 20
N_BOXES = 4
CFG = dict(surface_tension_method=2, surface_tension=0.2, target_time=1.0e9, max_error=0.05, max_error_v=0.1)


# ------------------------------------------------------------------------------------------------
def algorithmic_bytes_per_particle_step(nbar, D, P):
    """SURVEY.md §8d / DESIGN.md: FP64 state streamed once per pass + 4-byte neighbour indices."""
    return 1212.0 + 48.0 * nbar + D * (176.0 + 8.0 * nbar) + P * (184.0 + 8.0 * nbar)


# algorithmic bytes per fluid particle per launch of each kernel; nf/nb = mean fluid /
# boundary neighbours per particle of the run
KERNEL_BYTES = {
    "k_nbr_build": lambda nf, nb: 32 + 8 + 4 * (nf + nb),
    "k_density_factor": lambda nf, nb: 32 + 8 + 8 + 8 + 32 + 32 + 4 * (nf + nb),
    "k_rho": lambda nf, nb: 32 + 32 + 8 + 8 + 8 + 8 + 8 + 4 + 8 + 32 + 4 * (nf + nb),
    "k_push": lambda nf, nb: 32 + 32 + 32 + 8 + 8 + 4 + 8 + 4 * (nf + nb),
    "k_normals": lambda nf, nb: 32 + 4 + 32 + 4 * nf,
    "k_nonpressure": lambda nf, nb: 32 + 32 + 8 + 32 + 8 + 4 + 16 + 32 + 32 + 4 * nf,
    "k_permute_fluid": lambda nf, nb: 4 + 2 * (32 + 32 + 8 + 8 + 4 + 4),
    "k_advect_x": lambda nf, nb: 32 + 32 + 32 + 4 + 16,
}


def kernel_class(name):
    base = name.strip("()").split("<")[0].strip()
    return base


def kernel_bytes_per_particle(name, nbar):
    """Algorithmic bytes per fluid particle of ONE launch of `name` (the profile key, template arguments included)."""
    fn = KERNEL_BYTES.get(kernel_class(name))
    if fn is None:
        return None
    b = fn(nbar, 0.0)
    # passes fused into k_rho launches of the divergence solve (dfr_kernels.cuh: RhoExtra)
    if "X_DENSITY" in name:
        b += 8 + 8 + 32 + 32          # + density, factor, sum V gradW, (x, rho) written
    elif "X_NORMALS" in name:
        b += 32                        # + normal written
    elif "X_NONPRESSURE" in name:
        b += 32 + 32                   # + own normal read, acceleration written
    return b


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.window = None

    def mark_begin(self):
        """The sampler is started well before the timed region (nvidia-smi needs ~0.1 s to come up); only the samples taken
        between mark_begin() and stop() count."""
        import datetime

        self.window = datetime.datetime.now()
        self.window_pc = time.perf_counter()

    def start(self):
        # NVML in a thread (one sample every ~2 ms: a 20-step timed region is only ~60 ms long); nvidia-smi -lms as fallback
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml_samples = []
            self.nvml_stop = False
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001 - no NVML binding / no permission: use the command-line tool
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.nvml_stop:
            try:
                t = time.perf_counter()
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(reasons_fn(self.handle))
                try:
                    power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:  # noqa: BLE001
                    power = None
                self.nvml_samples.append((t, sm, mask, power))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def _stop_nvml(self):
        t_end = time.perf_counter()
        self.nvml_stop = True
        self.thread.join(timeout=1.0)
        t0 = getattr(self, "window_pc", None)
        inside = [x for x in self.nvml_samples if (t0 is None or x[0] >= t0) and x[0] <= t_end]
        note = None
        if not inside:
            inside, note = self.nvml_samples, "no sample fell inside the timed region; all samples of the run (warm-up included) are used"
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        reasons = sorted({nm for x in inside for bit, nm in names.items() if x[2] & bit})
        power = [x[3] for x in inside if x[3] is not None]
        out = {"sm_mhz": float(np.median([x[1] for x in inside])) if inside else None, "sm_max_mhz": self.sm_max,
               "power_w_max": max(power) if power else None, "samples": len(inside), "reasons": reasons, "source": "nvml"}
        if note:
            out["note"] = note
        return out

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if getattr(self, "nvml", None) is not None:
            return self._stop_nvml()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        import datetime

        t_end = datetime.datetime.now()
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        out = self._summarise(t_end, filtered=True)
        if not out["samples"]:  # the region was shorter than one sampling interval: fall back to every sample taken
            out = self._summarise(t_end, filtered=False)
            out["note"] = "no sample fell inside the timed region; all samples of the run (warm-up included) are used"
        return out

    def _summarise(self, t_end, filtered):
        import datetime

        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            if filtered and self.window is not None and len(f) >= 10:  # keep the samples of the timed region
                try:
                    ts = datetime.datetime.strptime(f[9], "%Y/%m/%d %H:%M:%S.%f")
                    if ts < self.window or ts > t_end:
                        continue
                except ValueError:
                    pass
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_scene(n_particles):
    from difffr_b200 import scenes

    return scenes.dam_break_scene(n_particles, n_boxes=N_BOXES)


# ------------------------------------------------------------------------------------------------
class quiet_stdout:
    """The reference's own code prints to stdout (e.g. Dynamic3dRigidBody::determineMassProperties); bench.py must
    print exactly one JSON line, so fd 1 is parked on /dev/null while the CPU arm runs."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def use_all_host_threads():
    """The CPU arm uses every host core it may run on.  torch.distributed.run injects OMP_NUM_THREADS=1 into every rank;
    that value is replaced before the OpenMP runtime of the CPU library reads it, and the runtime is told again after
    loading (it may already be mapped by another module)."""
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    return n


def cpu_arm(lib_path, prefix, scene, steps, warmup):
    """Time the CPU implementation (oracle port or oracle/_ref) on `scene`; returns (particle-steps/s, cores, ms/step, info)."""
    from difffr_b200 import scenes
    from difffr_b200.cabi import Context

    n_threads = use_all_host_threads()
    with quiet_stdout():
        lib = ctypes.CDLL(lib_path)
        try:
            gomp = ctypes.CDLL("libgomp.so.1")
            gomp.omp_set_dynamic(0)
            gomp.omp_set_num_threads(n_threads)
        except OSError:
            pass
        ctx = scenes.build_context(lambda **k: Context(lib=lib, prefix=prefix, **k), scene, **CFG)
        cores = 1
        if hasattr(lib, prefix + "num_threads"):
            f = getattr(lib, prefix + "num_threads")
            f.restype = ctypes.c_int
            cores = int(f())
        if warmup:
            ctx.step(warmup)
        p0 = ctx.step_info().total_particle_steps
        t0 = time.perf_counter()
        ctx.step(steps)
        dt = time.perf_counter() - t0
        info = ctx.step_info()
        psteps = info.total_particle_steps - p0
        ctx.close()
    return psteps / dt, cores, 1e3 * dt / max(steps, 1), info


def reference_arm(args, rank, world):
    if rank != 0:
        return
    ref_so = os.path.join(ROOT, "oracle", "_ref", "fast", "libref.so")
    fast = os.path.join(ROOT, "oracle", "liboracle_fast.so")
    if os.path.exists(ref_so):
        lib_path, prefix, kind = ref_so, "ref_", "reference"
    else:
        if not os.path.exists(fast):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle_fast.so"], check=True)
        lib_path, prefix, kind = fast, "orc_", "port"
    # bounded sample: same scene family, sized so that (K + W) steps take about two minutes on the host cores
    budget_particle_steps = 120.0 * 4.0e5
    n_ref = int(min(N_PARTICLES, max(20000, budget_particle_steps / max(args.steps + args.warmup, 1))))
    scene = make_scene(n_ref)
    value, cores, ms, info = cpu_arm(lib_path, prefix, scene, args.steps, args.warmup)
    nfl = len(scene['fluid'])
    if args.gpus > 1 and args.mode == "slab":
        rel = (f"bounded sample of the workload: the {args.gpus} x {N_PARTICLES}-particle scene of `config` is {args.gpus} x this one; "
               f"the CPU code's per-particle cost does not depend on the scene size, so particle-steps/s carry over")
    else:
        rel = "the full workload" if nfl >= N_PARTICLES * 0.99 else "same scene family, reduced size"
    sample = (f"{args.steps} steps (+{args.warmup} warm-up) of the dam break with {nfl} fluid particles ({rel}), FP64, "
              f"OpenMP on {cores} threads")
    line = {
        "impl": "reference", "metric": "fwd+adjoint particle-steps/s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, args.mode),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind, "sample": sample,
                         "sample_fluid_particles": nfl, "host_cores": host_cores()},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, mode="slab", per_gpu=None):
    N_PARTICLES = per_gpu or globals()["N_PARTICLES"]
    l2 = "working set per step (~0.5 GB of particle state and neighbour lists per GPU) exceeds the 126 MB L2; no flush needed"
    if n_gpus > 1 and mode == "slab":
        return {"workload": f"synthetic dam break + {N_BOXES} dynamic rigid boxes, ONE scene of {n_gpus} x {N_PARTICLES} fluid particles "
                            f"(BASELINE.json configs[4]); forward step + force/torque Jacobians + sensitivity chain rule",
                "particles_per_gpu": N_PARTICLES, "rollouts": 1,
                "parallelism": f"slab domain decomposition x{n_gpus} (NCCL/NVLink halo exchange per kernel pass)", "l2": l2}
    return {"workload": f"synthetic dam break + {N_BOXES} dynamic rigid boxes, {N_PARTICLES} fluid particles per rollout "
                        f"(BASELINE.json configs[4]); forward step + force/torque Jacobians + sensitivity chain rule",
            "particles_per_gpu": N_PARTICLES, "rollouts": n_gpus, "parallelism": f"independent rollouts x{n_gpus}", "l2": l2}


def stone_skipping_iteration(device, torch):
    """BASELINE.json's third figure on the scene it is defined on: the wall time of ONE gradient iteration of stone
    skipping (configs[0]; diff-stone-skipping.json as parsed and sampled by the host loader, with the regenerated settled
    fluid - the inputs of the whole-trajectory parity test, tests/golden/trajectory/).  Timed: dfr_reset + the trajectory
    to its end (dfr_run_trajectory: ~1,460 CFL-limited steps of 237,699 particles) + the final state and the eight
    sensitivity blocks.  The authors' log of the same iteration (float32 CPU build, their workstation):
    497 s (raw_record_and_plot/stone_skipping/2023-05-19-stone-skipping-ours/log/SPH_log.txt:29-44)."""
    from difffr_b200.cabi import Config, Context

    traj = os.path.join(ROOT, "tests", "golden", "trajectory")
    f_rec, f_x = os.path.join(traj, "traj_stone_skipping.npz"), os.path.join(traj, "stone_skipping_settled.npz")
    if not (os.path.exists(f_rec) and os.path.exists(f_x)):
        return None
    g = np.load(f_rec)
    x0 = np.load(f_x)["x"].astype(np.float64)
    ctx = Context(config=Config.from_buffer_copy(g["config_bytes"].tobytes()), device=device)
    ctx.set_fluid(x0, np.zeros_like(x0))
    nb = int(g["n_bodies"])
    for i in range(nb):
        ctx.add_body(g[f"body{i}_samples"], bool(g[f"body{i}_dynamic"]), float(g[f"body{i}_density"]), g[f"body{i}_translation"], g[f"body{i}_rotation"])
    dyn = [i for i in range(nb) if int(g[f"body{i}_dynamic"])]
    for i in dyn:
        ctx.set_init_v_omega(i, g[f"body{i}_init_v"], g[f"body{i}_init_omega"])
    ctx.finalize()
    ctx.load_fluid_state(x0, np.zeros_like(x0), None, None)
    ctx.run_trajectory(100000)  # warm-up iteration (graph capture, list capacities)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.reset()
    done = ctx.run_trajectory(100000)
    for b in dyn:
        ctx.body_state(b)
        for w in range(8):
            ctx.body_grad(b, w)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    info = ctx.step_info()
    out = {"seconds": dt, "steps": int(done), "fluid_particles": ctx.num_fluid, "ms_per_step": 1e3 * dt / max(int(done), 1),
           "mean_pressure_iterations": info.total_pressure_iterations / max(int(done), 1),
           "authors_log_seconds": 497.0,
           "what": "BASELINE.json configs[0], stone skipping: dfr_reset + dfr_run_trajectory to the end of the trajectory + final state and "
                   "sensitivity blocks on diff-stone-skipping.json with the regenerated settled fluid (tests/golden/trajectory/); "
                   "authors_log_seconds = the same iteration in the reference's own log (float32 CPU build on the authors' workstation)"}
    ctx.close()
    return out


def pysplishsplash_leg(scene, steps, device):
    """K steps through difffr_b200.pysplishsplash: SimulatorBase.runSimulation() with a per-step Python callback."""
    pkg = os.path.join(ROOT, "difffr_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)  # the mirror is imported under the reference's module name
    import pysplishsplash as sph

    base = sph.Exec.SimulatorBase()
    base.init(sceneFile="", useGui=False, initialPause=False, useCache=False, stopAt=1.0e9, stateFile="",
              outputDir=os.path.join(ROOT, "gpurun_out", "bench_pysph"))
    base.setDevice(device)
    cfg = dict(CFG, particle_radius=scene["radius"])
    bodies = [dict(x_local=b["x_local"], dynamic=bool(b["dynamic"]), density=float(b["density"]), position=np.asarray(b["position"], dtype=np.float64),
                   quat=np.asarray(b["quat"], dtype=np.float64)) for b in scene["bodies"]]
    base.initSimulationFromArrays(cfg, np.ascontiguousarray(scene["fluid"]), bodies)
    sim = sph.Simulation.getCurrent()
    ts = sim.getTimeStep()
    dyn = [i for i, b in enumerate(scene["bodies"]) if b["dynamic"]]
    state = {"n": 0, "bytes": 0, "t0": None, "t1": None, "warm": 3}

    def cb():
        state["n"] += 1
        if state["n"] == state["warm"]:
            state["t0"] = time.perf_counter()
            return
        if state["n"] < state["warm"]:
            return
        nbytes = 0
        for b in dyn:
            bm = ts.get_boundary_model(b)
            for a in (bm.get_position_rb(), bm.get_quaternion_rb_vec4(), bm.get_velocity_rb(), bm.get_angular_velocity_rb(),
                      bm.get_grad_x_to_v0(), bm.get_grad_x_to_omega0(), bm.get_grad_quaternion_to_v0(), bm.get_grad_quaternion_to_omega0(),
                      bm.get_grad_v_to_v0(), bm.get_grad_v_to_omega0(), bm.get_grad_omega_to_v0(), bm.get_grad_omega_to_omega0()):
                nbytes += a.nbytes
        state["bytes"] = nbytes
        if state["n"] == state["warm"] + steps:
            state["t1"] = time.perf_counter()
            base.stop()

    base.setTimeStepCB(cb)
    base.runSimulation()
    nf = sim.numberOfFluidParticles()
    base.cleanup()
    if state["t0"] is None or state["t1"] is None:
        return None
    dt = state["t1"] - state["t0"]
    return {"value": nf * steps / dt, "unit": "particle-steps/s", "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "d2h_bytes_per_step": state["bytes"],
            "what": "pysplishsplash.Exec.SimulatorBase.runSimulation() with a Python callback after every step that reads pose, "
                    "velocities and the eight sensitivity blocks of every dynamic body as numpy arrays"}


def rollouts_mode(args, rank, world, local_rank, dist, torch):
    """BASELINE.json configs[3]: a CMA-ES-sized population of independent rollouts sharded over the GPUs.

    64 candidate (v0, omega0) pairs drawn with numpy.random.default_rng(12) (opt-ng.py: seed 12; the reference evaluates
    its nevergrad population one candidate after the other) on a scene of the high-diving configuration's size (118,389
    fluid particles; synthetic dam break + one dynamic box, fixed time step, `--rollout-steps` steps per trajectory).
    Each rank evaluates its block with `--concurrency` contexts side by side on its GPU; no data-path collective, one
    small all_gather of the results.  Timed with the wall clock between device synchronisations (several streams per
    GPU), maximum over ranks."""
    from difffr_b200 import rollouts, scenes
    from difffr_b200.cabi import Context

    n_cand, n_part, steps = args.candidates, args.rollout_particles, args.rollout_steps
    rng = np.random.default_rng(12)
    real = args.rollout_scene == "high_diving"
    if real:
        # the scene itself: diff-high-diving-duck.json as the host loader parsed and sampled it (the inputs of the golden
        # tests/golden/paper_high_diving.npz: 118,389 fluid particles on the lattice, the duck + four static meshes with
        # 238 k samples, the box emitter), whole trajectories (targetTime 2.3 s); design variable = the duck's initial
        # angular velocity, as in opt-ng.py --taskType high-diving
        from difffr_b200.cabi import Config

        g = np.load(os.path.join(ROOT, "tests", "golden", "paper_high_diving.npz"))
        nb = int(g["n_bodies"])
        body = [i for i in range(nb) if int(g[f"body{i}_dynamic"])][0]
        cands = [(np.asarray(g[f"body{body}_init_v"], dtype=np.float64), rng.normal(size=3)) for _ in range(n_cand)]
        steps = 1 << 20

        def make():
            ctx = Context(config=Config.from_buffer_copy(g["config_bytes"].tobytes()), device=local_rank)
            ctx.set_fluid(g["fluid_x"], g["fluid_v"])
            for i in range(nb):
                ctx.add_body(g[f"body{i}_samples"].astype(np.float64), bool(g[f"body{i}_dynamic"]), float(g[f"body{i}_density"]),
                             g[f"body{i}_translation"], g[f"body{i}_rotation"])
            for k in range(int(g["n_emitters"])):
                wh, vse = g[f"emitter{k}_wh"], g[f"emitter{k}_vse"]
                ctx.add_emitter(width=int(wh[0]), height=int(wh[1]), position=g[f"emitter{k}_position"], rotation=g[f"emitter{k}_rotation"],
                                velocity=float(vse[0]), emit_start=float(vse[1]), emit_end=float(vse[2]))
            ctx.finalize()
            return ctx
    else:
        scene = scenes.dam_break_scene(n_part, n_boxes=1)
        cfg = dict(CFG)
        cfg.update(cfl_method=0, time_step_size=1.0e-3, uniform_acc_rb_time=0.02, target_time=steps * 1.0e-3 - 0.02 - 0.5e-3)
        cands = [(rng.normal(size=3), rng.normal(size=3)) for _ in range(n_cand)]
        body = [i for i, b in enumerate(scene["bodies"]) if b["dynamic"]][0]

        def make():
            return scenes.build_context(lambda **k: Context(device=local_rank, **k), scene, **cfg)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    gather = rollouts.torch_gather(world) if dist is not None else None
    conc = max(1, args.concurrency)
    ctxs = [make() for _ in range(conc)]
    nf = ctxs[0].num_fluid
    # warm-up: one rollout per context (graph capture, capacity checks)
    rollouts.run_population(None, cands[:conc * world], body, rank, world, max_steps=steps + 10, gather=gather, contexts=ctxs)
    results = {}
    legs = [("concurrent", ctxs)] + ([("serial", ctxs[:1])] if world == 1 and conc > 1 and not args.no_serial_leg else [])
    for label, cs in legs:
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0 and label == "concurrent":
            sampler.start()
        stats = {}
        t0 = time.perf_counter()
        res, nsteps = rollouts.run_population(None, cands, body, rank, world, max_steps=steps + 10, gather=gather, contexts=cs, stats=stats)
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if (rank == 0 and label == "concurrent") else None
        launches = stats.get("kernel_launches", 0)
        if dist is not None:
            t = torch.tensor([wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
            w = torch.tensor([launches], dtype=torch.float64, device="cuda")
            dist.all_reduce(w, op=dist.ReduceOp.SUM)
            launches = float(w.item())
        results[label] = dict(wall=wall, steps=int(nsteps.sum()), launches=launches, clocks=clocks, checksum=float(np.abs(res).sum()))
    if rank != 0:
        return
    c = results["concurrent"]
    value = nf * c["steps"] / c["wall"]  # (emitted particles not counted: a lower bound on the real scene)
    line = {
        "metric": "fwd+adjoint particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * c["wall"] / max(c["steps"], 1) * world * conc, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"population of {n_cand} independent rollouts (BASELINE.json configs[3]: CMA-ES-sized population, candidates from "
                                f"numpy default_rng(12)); " +
                                (f"diff-high-diving-duck.json itself ({nf} fluid particles + emitter, whole trajectories of targetTime 2.3 s, "
                                 f"{c['steps'] // max(n_cand, 1)} steps on average)" if real else
                                 f"synthetic dam break + 1 dynamic box at the high-diving scene's size ({nf} fluid particles), "
                                 f"{steps} steps per trajectory, fixed h = 1e-3") + "; forward step + Jacobians + sensitivity chain rule"),
                   "particles_per_rollout": nf, "rollouts": n_cand, "steps_per_rollout": c["steps"] // max(n_cand, 1),
                   "parallelism": f"candidates sharded over {world} GPU(s), {conc} contexts side by side per GPU; no data-path collective",
                   "l2": "several independent working sets per GPU, each larger than its share of the 126 MB L2; no flush needed"},
        "rollouts_per_s": n_cand / c["wall"], "seconds_per_population": c["wall"], "concurrency": conc,
        "clocks": c["clocks"], "gpu_launches": int(c["launches"]),
        "timing": "wall clock between device synchronisations (several streams per GPU), max over ranks",
    }
    if "serial" in results:
        sr = results["serial"]
        line["serial"] = {"seconds_per_population": sr["wall"], "value": nf * sr["steps"] / sr["wall"],
                          "speedup_of_concurrent": sr["wall"] / c["wall"], "same_results": sr["checksum"] == c["checksum"]}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("DFR_BENCH_MODE", "slab"), choices=["slab", "rollouts"],
                    help="sharding for --gpus > 1: one slab-decomposed scene (default) or independent rollouts")
    ap.add_argument("--particles", type=int, default=N_PARTICLES, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-iteration-leg", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-settled-leg", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-pysph-leg", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--preroll", type=int, default=1200, help=argparse.SUPPRESS)
    ap.add_argument("--candidates", type=int, default=64, help="--mode rollouts: population size")
    ap.add_argument("--concurrency", type=int, default=4, help="--mode rollouts: contexts side by side per GPU")
    ap.add_argument("--rollout-particles", type=int, default=118389, help=argparse.SUPPRESS)
    ap.add_argument("--rollout-steps", type=int, default=200, help=argparse.SUPPRESS)
    ap.add_argument("--no-serial-leg", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--rollout-scene", default="synthetic", choices=["synthetic", "high_diving"],
                    help="--mode rollouts: the synthetic scene of the high-diving size (default, 200 steps) or diff-high-diving-duck.json itself")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from difffr_b200 import scenes
    from difffr_b200.cabi import Context

    if args.mode == "rollouts":
        rollouts_mode(args, rank, world, local_rank, dist, torch)
        if dist is not None:
            dist.destroy_process_group()
        return

    slab = world > 1 and args.mode == "slab"
    n_particles = args.particles * (world if slab else 1)
    scene = make_scene(n_particles)

    def make_ctx(**k):
        ctx = Context(device=local_rank, **k)
        if slab:  # every rank builds the same scene and keeps one slab of it
            ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                ident = torch.frombuffer(bytearray(ctx.slab_unique_id()), dtype=torch.uint8).cuda()
            dist.broadcast(ident, 0)
            ctx.slab_configure(rank, world, ident.cpu().numpy().tobytes())
        return ctx

    ctx = scenes.build_context(make_ctx, scene, **CFG)
    nf = ctx.num_fluid
    dyn = [i for i, b in enumerate(scene["bodies"]) if b["dynamic"]]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.step(args.warmup)
    barrier()
    ms0, l0 = ctx.device_time_ms()
    i0 = ctx.step_info()
    barrier()
    sampler.mark_begin()
    t0 = time.perf_counter()
    ctx.step(args.steps)  # one dfr_step call: K steps enqueued back to back, events on the context's stream
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    ms1, l1 = ctx.device_time_ms()
    i1 = ctx.step_info()
    dev_ms = ms1 - ms0
    psteps = i1.total_particle_steps - i0.total_particle_steps
    D = (i1.total_divergence_iterations - i0.total_divergence_iterations) / args.steps
    P = (i1.total_pressure_iterations - i0.total_pressure_iterations) / args.steps
    nbar = (i1.total_fluid_neighbors - i0.total_fluid_neighbors) / max(psteps, 1)

    # ---- end to end through the C ABI with host buffers: one "gradient iteration" ----
    # H2D: the trajectory's initial fluid state (x, v, kappa, kappa_v: 64 B/particle, pinned) + per-step rigid control
    # input; D2H: per-step rigid state and the eight sensitivity blocks (what the optimisation scripts read per step).
    # A slab-decomposed rank holds (and uploads) its own rows of the scene's arrays: dfr_slab_local_ids /
    # dfr_load_fluid_state_local.  (Round 1 handed every rank the whole scene and let it pick its rows on the host.)
    ids = ctx.slab_local_ids() if slab else None
    x_src = scene["fluid"][ids] if slab else scene["fluid"]
    n_rows = x_src.shape[0]
    x_h = torch.from_numpy(np.ascontiguousarray(x_src)).pin_memory()
    v_h = torch.zeros_like(x_h).pin_memory()
    k_h = torch.zeros(n_rows, dtype=torch.float64).pin_memory()
    kv_h = torch.zeros(n_rows, dtype=torch.float64).pin_memory()
    e2e_steps = args.steps
    barrier()
    t0 = time.perf_counter()
    if slab:
        ctx.load_fluid_state_local(x_h.numpy(), v_h.numpy(), k_h.numpy(), kv_h.numpy())  # also resets the context
    else:
        ctx.load_fluid_state(x_h.numpy(), v_h.numpy(), k_h.numpy(), kv_h.numpy())
    t_load = time.perf_counter() - t0
    d2h = 0
    for k in range(e2e_steps):
        for b in dyn:
            # a control input that differs from the previous step's, so that the 48-byte H2D copy really happens every
            # step (dfr_set_init_v_omega skips the copy for unchanged values); it only acts during the velocity ramp
            ctx.set_init_v_omega(b, (0.0, 0.0, 0.0), (0.0, 0.0, 1.0e-12 * (k & 1)))
        ctx.step(1)
        for b in dyn:
            ctx.body_state(b)
            for w in range(8):
                ctx.body_grad(b, w)
    barrier()
    e2e_wall = time.perf_counter() - t0
    h2d_per_step = (64 * n_rows) / e2e_steps + 48 * len(dyn)  # per rank
    d2h_per_step = len(dyn) * (13 + 4 * 9 + 2 * 12 + 2 * 9) * 8 + 8 * 40
    e2e_psteps = nf * e2e_steps
    slab_info = ctx.slab_info() if slab else None
    if slab:  # nf is the whole scene; every rank uploads / steps its share
        e2e_psteps = (i1.total_particle_steps - i0.total_particle_steps) / args.steps * e2e_steps

    # ---- per-kernel launch durations, live, CUDA events on the context's stream ----
    ctx.set_profiling(True)
    ctx.step(max(3, min(10, args.steps)))
    prof = ctx.kernel_profile()
    ctx.set_profiling(False)
    info_p = ctx.step_info()
    nf_rank = info_p.num_fluid_particles  # fluid particles this rank computes (all of them unless slab-decomposed)

    # ---- settled-state leg (N = 1): the same K steps after a pre-roll that lets the lattice start break up ----
    # The 2r lattice with V = 0.8 (2r)^3 is under-dense: a flowing fluid has ~40 neighbours per particle instead of ~29,
    # and every list pass scales with that.  This is the operating point of a long rollout; reported next to the
    # lattice-start number above, not instead of it.
    settled = None
    if world == 1 and not args.no_settled_leg:
        ctx.step(args.preroll)
        ctx.step(3)  # untimed: a list capacity that the pre-roll outgrew is raised (and the step graphs re-recorded) here
        torch.cuda.synchronize()
        msa, _ = ctx.device_time_ms()
        ia = ctx.step_info()
        ctx.step(args.steps)
        torch.cuda.synchronize()
        msb, _ = ctx.device_time_ms()
        ib = ctx.step_info()
        ps = ib.total_particle_steps - ia.total_particle_steps
        Ds = (ib.total_divergence_iterations - ia.total_divergence_iterations) / args.steps
        Ps = (ib.total_pressure_iterations - ia.total_pressure_iterations) / args.steps
        nbs = (ib.total_fluid_neighbors - ia.total_fluid_neighbors) / max(ps, 1)
        vs = ps / ((msb - msa) * 1e-3)
        bs = algorithmic_bytes_per_particle_step(nbs, Ds, Ps)
        settled = {"value": vs, "unit": "particle-steps/s", "ms_per_step": (msb - msa) / args.steps, "preroll_steps": args.preroll,
                   "mean_neighbors": nbs, "D": Ds, "P": Ps, "h": ib.time_step_size,
                   "whole_step": {"bytes_per_particle_step": bs, "achieved": vs * bs / 1e9, "frac": vs * bs / 1e9 / load_peaks()[0]}}

    # ---- end to end through the pysplishsplash mirror with a per-step Python callback (N = 1) ----
    # what a user of the reference's scripts pays: SimulatorBase.runSimulation() calling a Python callback after every
    # step, the callback reading the rigid state and the sensitivity blocks as numpy arrays (gradient-based-optimize.py)
    e2e_py = None
    if world == 1 and not args.no_pysph_leg:
        e2e_py = pysplishsplash_leg(scene, args.steps, local_rank)

    # ---- aggregate over ranks: max time, sum of work ----
    if dist is not None:
        t = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, e2e_ms = [float(v) for v in t.tolist()]
        w = torch.tensor([psteps, e2e_psteps, l1 - l0], dtype=torch.float64, device="cuda")
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        psteps_all, e2e_psteps_all, launches_all = [float(v) for v in w.tolist()]
    else:
        wall_ms, e2e_ms = wall * 1e3, e2e_wall * 1e3
        psteps_all, e2e_psteps_all, launches_all = float(psteps), float(e2e_psteps), float(l1 - l0)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = psteps_all / (dev_ms * 1e-3)
    peak, peak_src = load_peaks()
    # dominant kernel class of the profiled steps
    classes = {}
    for name, (ms, n) in prof.items():
        if name.endswith("(idle)"):
            continue
        c = kernel_class(name)
        a = classes.setdefault(c, [0.0, 0])
        a[0] += ms
        a[1] += n
    total_ms = sum(v[0] for v in classes.values())
    top = max(classes.items(), key=lambda kv: kv[1][0])
    top_name, (top_ms, top_n) = top[0], top[1]
    # algorithmic bytes of the launches of that class in the profiled steps, variant by variant
    top_bytes_total, known = 0.0, True
    for name, (ms, n) in prof.items():
        if name.endswith("(idle)") or kernel_class(name) != top_name:
            continue
        b = kernel_bytes_per_particle(name, nbar)
        if b is None:
            known = False
            break
        top_bytes_total += b * nf_rank * n
    if known and top_n:
        top_bytes = top_bytes_total / top_n            # mean per launch
        achieved = top_bytes_total / (top_ms * 1e-3) / 1e9
    else:
        top_bytes, achieved = None, None
    step_bytes = algorithmic_bytes_per_particle_step(nbar, D, P)
    roofline = {
        "bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
        "kernel_share_of_step": top_ms / total_ms if total_ms else None,
        "kernel_avg_ms": top_ms / top_n, "algorithmic_bytes_per_launch": top_bytes,
        "whole_step": {"bytes_per_particle_step": step_bytes, "achieved": value * step_bytes / 1e9,
                       "frac": value * step_bytes / 1e9 / peak, "mean_neighbors": nbar, "D": D, "P": P},
        "kernels_ms_per_step": {k: v[0] / max(3, min(10, args.steps)) for k, v in sorted(classes.items(), key=lambda kv: -kv[1][0])[:16]},
    }
    ncu = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(ncu):
        try:
            prof_file = json.load(open(ncu))
            roofline["traffic"] = prof_file.get(top_name)
            # what actually bounds the list kernels on this part: the SM's L1 data stage (ncu, not measured in this run)
            busy = prof_file.get("l1_data_stage_busy", {}).get(top_name)
            if busy is not None:
                roofline["binding_resource"] = {"name": "L1 data-stage wavefronts (l1tex__data_pipe_lsu_wavefronts, share of peak)",
                                                "busy_frac": busy, "source": "profiles/traffic.json from the ncu --set full capture "
                                                "summarised in profiles/r2_ncu_full.md; see profiles/r2_gather_analysis.md"}
        except Exception:
            pass

    # ---- wall time of one gradient iteration at a paper-scene size (BASELINE.json's third figure) ----
    # the stone-skipping configuration's 237,699 fluid particles and ~1,450 steps, on the synthetic
    # scene family: reset + one trajectory run to its end by dfr_run_trajectory (forward + sensitivities, one host
    # read-back per step like the scripts' per-step callback) + the read of the final state and the eight blocks
    gradient_iteration = None
    if world == 1 and not args.no_iteration_leg:
        gradient_iteration = stone_skipping_iteration(local_rank, torch)
    if world == 1 and not args.no_iteration_leg and gradient_iteration is None:  # fixtures absent: synthetic stand-in of that size
        n_it, steps_it = 237699, 1450
        sc_it = make_scene(n_it)
        cfg_it = dict(CFG)
        cfg_it.update(cfl_method=0, time_step_size=1.0e-3, target_time=steps_it * 1.0e-3 - 0.5e-3)
        ctx_it = scenes.build_context(lambda **k: Context(device=local_rank, **k), sc_it, **cfg_it)
        ctx_it.run_trajectory(steps_it + 10)  # warm-up iteration (allocations, capacity checks)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx_it.reset()
        done = ctx_it.run_trajectory(steps_it + 10)
        for b in [i for i, bd in enumerate(sc_it["bodies"]) if bd["dynamic"]]:
            ctx_it.body_state(b)
            for w in range(8):
                ctx_it.body_grad(b, w)
        torch.cuda.synchronize()
        dt_it = time.perf_counter() - t0
        gradient_iteration = {"seconds": dt_it, "steps": int(done), "fluid_particles": ctx_it.num_fluid,
                              "ms_per_step": 1e3 * dt_it / max(int(done), 1),
                              "what": "dfr_reset + dfr_run_trajectory to the end of the trajectory + final state and sensitivity blocks; "
                                      "synthetic dam break at the stone-skipping scene's size (237,699 particles, ~1,450 steps, fixed h = 1e-3)"}
        ctx_it.close()

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        fast = os.path.join(ROOT, "oracle", "liboracle_fast.so")
        ref_so = os.path.join(ROOT, "oracle", "_ref", "fast", "libref.so")
        if os.path.exists(ref_so):
            lib_path, prefix, kind = ref_so, "ref_", "reference"
        else:
            lib_path, prefix, kind = fast, "orc_", "port"
        if not os.path.exists(lib_path):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle_fast.so"], check=True)
        k_cpu = 6
        cv, cores, cms, _ = cpu_arm(lib_path, prefix, scene, k_cpu, 1)
        cpu_baseline = {"value": cv, "unit": "particle-steps/s", "cores": cores, "kind": kind, "ms_per_step": cms,
                        "sample": f"steps 2..{k_cpu + 1} of the same {nf}-particle scene (1 warm-up step), FP64, OpenMP"}

    line = {
        "metric": "fwd+adjoint particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(world, args.mode, args.particles),
        "wall_ms_per_step": wall_ms / args.steps,
        "solver": {"mean_neighbors": nbar, "divergence_iters": D, "pressure_iters": P, "h": i1.time_step_size},
        "clocks": clocks,
        "e2e": {"value": e2e_psteps_all / (e2e_ms * 1e-3), "unit": "particle-steps/s", "h2d_bytes_per_step": h2d_per_step,
                "d2h_bytes_per_step": d2h_per_step, "ms_per_step": e2e_ms / e2e_steps, "load_state_ms_rank0": 1e3 * t_load,
                "what": "dfr_load_fluid_state (slab: dfr_load_fluid_state_local, each rank its own rows) from pinned host arrays + per step: "
                        "dfr_set_init_v_omega, dfr_step(1), dfr_get_body_state + 8x dfr_get_body_grad per dynamic body; byte counts per rank"},
        "gpu_launches": int(launches_all),
        "slab": ({"owned_rank0": slab_info["owned"], "ghosts_rank0": slab_info["ghosts"], "ghost_transport": slab_info["transport"],
                  "nvlink_bytes_per_step_rank0": slab_info["exchanged_bytes"] / max(e2e_steps, 1)} if slab else None),
        "roofline": roofline,
        "settled": settled,
        "e2e_pysplishsplash": e2e_py,
        "cpu_baseline": cpu_baseline,
        "gradient_iteration": gradient_iteration,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
