"""ctypes binding of the C ABI declared in include/dfr.h.

`Context` wraps one `dfr_context`.  By default it binds the CUDA library
`difffr_b200/csrc/libdfr.so` (symbols `dfr_*`).  There is no CPU fallback in the product: if
the CUDA library is missing, import fails loudly; if no CUDA device is present, `Context()`
raises `DfrError` (DFR_ERR_NO_DEVICE).

The tests bind the CPU oracle through the same class by passing `lib=` and `prefix="orc_"`
explicitly; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFR_LIBRARY") or os.path.join(_HERE, "csrc", "libdfr.so")  # DFR_LIBRARY: tuning builds


class DfrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dfr error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    """Mirror of `dfr_config` (include/dfr.h); field names follow the reference's scene keys."""

    _fields_ = [
        ("particle_radius", C.c_double),
        ("density0", C.c_double),
        ("gravitation", C.c_double * 3),
        ("cfl_method", C.c_int32),
        ("cfl_factor", C.c_double),
        ("cfl_min_time_step", C.c_double),
        ("cfl_max_time_step", C.c_double),
        ("time_step_size", C.c_double),
        ("min_iterations", C.c_int32),
        ("max_iterations", C.c_int32),
        ("max_error", C.c_double),
        ("max_iterations_v", C.c_int32),
        ("max_error_v", C.c_double),
        ("enable_divergence_solver", C.c_int32),
        ("use_pressure_warmstart", C.c_int32),
        ("use_divergence_warmstart", C.c_int32),
        ("viscosity_method", C.c_int32),
        ("viscosity", C.c_double),
        ("viscosity_boundary", C.c_double),
        ("surface_tension_method", C.c_int32),
        ("surface_tension", C.c_double),
        ("surface_tension_boundary", C.c_double),
        ("gradient_mode", C.c_int32),
        ("rigid_body_mode", C.c_int32),
        ("optimize_rotation", C.c_int32),
        ("use_rigid_gradient_manager", C.c_int32),
        ("use_rigid_contact_solver", C.c_int32),
        ("rigid_contact_beta", C.c_double),
        ("rigid_contact_gamma", C.c_double),
        ("rigid_contact_friction", C.c_double),
        ("rigid_contact_support_radius_factor", C.c_double),
        ("target_time", C.c_double),
        ("uniform_acc_rb_time", C.c_double),
        ("max_emitted_particles", C.c_int32),
        ("neighbor_capacity_fluid", C.c_int32),
        ("neighbor_capacity_boundary", C.c_int32),
        ("body_neighbor_capacity", C.c_int32),
        ("grid_reach", C.c_int32),
        ("use_release_rigid_body_mode", C.c_int32),
        ("reserved_i", C.c_int32 * 2),
        ("reserved_d", C.c_double * 8),
    ]


class StepInfo(C.Structure):
    _fields_ = [
        ("time", C.c_double),
        ("time_step_size", C.c_double),
        ("iterations", C.c_int32),
        ("iterations_v", C.c_int32),
        ("step_count", C.c_int32),
        ("trajectory_finished", C.c_int32),
        ("num_fluid_particles", C.c_int64),
        ("total_pressure_iterations", C.c_int64),
        ("total_divergence_iterations", C.c_int64),
        ("total_particle_steps", C.c_int64),
        ("total_fluid_neighbors", C.c_int64),
    ]


# every symbol include/dfr.h declares (without prefix); tests check the library exports all of them
SYMBOLS = [
    "default_config", "create", "destroy", "last_error", "set_fluid", "add_body",
    "set_init_v_omega", "finalize", "load_fluid_state", "reset", "step", "run_trajectory",
    "get_step_info", "get_body_state", "set_body_velocity", "get_body_properties",
    "get_body_grad", "get_manager_grad", "download_fluid", "download_body", "num_fluid", "num_fluid_initial",
    "num_body_particles", "num_bodies", "get_neighbors", "add_emitter", "get_device_time_ms",
    "set_profiling", "get_kernel_profile", "reset_gradient", "set_gradient_mode", "slab_plan", "slab_unique_id", "slab_configure", "slab_info", "slab_local_ids", "load_fluid_state_local",
]

GRAD_NAMES = [
    "grad_x_to_v0", "grad_x_to_omega0", "grad_quaternion_to_v0", "grad_quaternion_to_omega0",
    "grad_v_to_v0", "grad_v_to_omega0", "grad_omega_to_v0", "grad_omega_to_omega0",
    "grad_net_force_to_vn", "grad_net_force_to_xn", "grad_net_force_to_qn", "grad_net_force_to_omega_n",
    "grad_net_torque_to_vn", "grad_net_torque_to_xn", "grad_net_torque_to_qn", "grad_net_torque_to_omega_n",
]
_GRAD_SHAPE = {2: (4, 3), 3: (4, 3), 10: (3, 4), 14: (3, 4)}
FLUID_FIELDS = {
    "position": (0, 3), "velocity": (1, 3), "density": (2, 1), "factor": (3, 1), "kappa": (4, 1),
    "kappa_v": (5, 1), "density_adv": (6, 1), "acceleration": (7, 3), "sum_grad_p_k": (8, 3), "normal": (9, 3),
}
BODY_FIELDS = {"position": (0, 3), "velocity": (1, 3), "volume": (2, 1), "position0": (3, 3)}


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback."
        )
    return C.CDLL(path)


_default_lib = None


def default_library():
    global _default_lib
    if _default_lib is None:
        _default_lib = load_library()
    return _default_lib


def slab_plan(z, z_origin, cell, nz, reach, n_ranks, lib=None):
    """Cell-aligned cut of the z layers into `n_ranks` slabs of (nearly) equal particle count: `dfr_slab_plan`, the
    host-side planner `dfr_finalize` uses on a slab-decomposed context.  Returns the n_ranks + 1 plane indices."""
    lib = lib if lib is not None else default_library()
    z = np.ascontiguousarray(z, dtype=np.float64)
    planes = np.zeros(n_ranks + 1, dtype=np.int32)
    f = lib.dfr_slab_plan
    f.restype = C.c_int
    f.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int32)]
    rc = f(float(z_origin), 1.0 / float(cell), int(nz), int(reach), z.size, z.ctypes.data_as(C.POINTER(C.c_double)), int(n_ranks),
           planes.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise DfrError(rc, "slab plan: a slab would be thinner than two support radii")
    return planes


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Context:
    """One simulation context (one scene, one CUDA stream)."""

    def __init__(self, config: Config | None = None, device: int = 0, lib=None, prefix: str = "dfr_", **overrides):
        self._lib = lib if lib is not None else default_library()
        self._p = prefix
        self._setup_prototypes()
        if config is None:
            config = self.default_config()
        for k, v in overrides.items():
            if k == "gravitation":
                config.gravitation[:] = list(v)
            else:
                if not hasattr(config, k):
                    raise AttributeError(f"unknown config field {k}")
                setattr(config, k, v)
        self.config = config
        self._ctx = C.c_void_p()
        rc = self._fn("create")(C.byref(config), int(device), C.byref(self._ctx))
        if rc != 0:
            self._ctx = C.c_void_p()
            raise DfrError(rc, "dfr_create failed (no CUDA device / CUDA error)" if rc in (-2, -3) else "dfr_create failed")

    # -- plumbing -------------------------------------------------------------------------
    def _fn(self, name):
        return getattr(self._lib, self._p + name)

    def _setup_prototypes(self):
        L, p = self._lib, self._p
        if getattr(L, "_dfr_protos_" + p, False):
            return
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
        i64 = C.c_int64

        def proto(name, res, *args):
            f = getattr(L, p + name)
            f.restype = res
            f.argtypes = list(args)

        proto("default_config", None, C.POINTER(Config))
        proto("create", C.c_int, C.POINTER(Config), C.c_int, C.POINTER(vp))
        proto("destroy", None, vp)
        proto("last_error", C.c_char_p, vp)
        proto("set_fluid", C.c_int, vp, i64, dp, dp)
        proto("add_body", C.c_int, vp, i64, dp, C.c_int, C.c_double, dp, dp)
        proto("set_init_v_omega", C.c_int, vp, C.c_int, dp, dp)
        proto("finalize", C.c_int, vp)
        proto("load_fluid_state", C.c_int, vp, dp, dp, dp, dp)
        proto("reset", C.c_int, vp)
        proto("step", C.c_int, vp, C.c_int)
        proto("run_trajectory", C.c_int, vp, C.c_int, C.POINTER(C.c_int))
        proto("get_step_info", C.c_int, vp, C.POINTER(StepInfo))
        proto("get_body_state", C.c_int, vp, C.c_int, dp)
        proto("set_body_velocity", C.c_int, vp, C.c_int, dp, dp)
        proto("get_body_properties", C.c_int, vp, C.c_int, dp)
        proto("get_body_grad", C.c_int, vp, C.c_int, C.c_int, dp)
        proto("get_manager_grad", C.c_int, vp, C.c_int, C.c_int, C.c_int, dp)
        proto("download_fluid", C.c_int, vp, C.c_int, dp)
        proto("download_body", C.c_int, vp, C.c_int, C.c_int, dp)
        proto("num_fluid", i64, vp)
        proto("num_fluid_initial", i64, vp)
        proto("num_body_particles", i64, vp, C.c_int)
        proto("num_bodies", C.c_int, vp)
        proto("get_neighbors", C.c_int, vp, C.c_int, C.c_int, ip, ip, i64, C.POINTER(i64))
        proto("add_emitter", C.c_int, vp, C.c_int, C.c_int, dp, dp, C.c_double, C.c_double, C.c_double)
        proto("get_device_time_ms", C.c_int, vp, dp, C.POINTER(i64))
        if hasattr(L, p + "slab_configure"):  # multi-GPU entry points of the CUDA library only
            proto("slab_unique_id", C.c_int, C.c_char_p)
            proto("slab_configure", C.c_int, vp, C.c_int, C.c_int, C.c_char_p)
            proto("slab_info", C.c_int, vp, C.POINTER(i64))
            proto("slab_local_ids", i64, vp, ip, i64)
            proto("load_fluid_state_local", C.c_int, vp, i64, dp, dp, dp, dp)
        proto("reset_gradient", C.c_int, vp)
        proto("set_gradient_mode", C.c_int, vp, C.c_int)
        if hasattr(L, p + "set_profiling"):  # the CPU oracle has no kernels to profile
            proto("set_profiling", C.c_int, vp, C.c_int)
            proto("get_kernel_profile", C.c_int, vp, C.c_int, C.c_char_p, C.c_int, dp, C.POINTER(i64))
        setattr(L, "_dfr_protos_" + p, True)

    def default_config(self) -> Config:
        cfg = Config()
        self._fn("default_config")(C.byref(cfg))
        return cfg

    def _check(self, rc):
        if rc < 0:
            msg = self._fn("last_error")(self._ctx)
            raise DfrError(rc, msg.decode() if msg else "")
        return rc

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._fn("destroy")(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene construction ---------------------------------------------------------------
    def set_fluid(self, x, v=None):
        x = _f64(x, (-1, 3))
        v = _f64(v, (-1, 3)) if v is not None else np.zeros_like(x)
        self._check(self._fn("set_fluid")(self._ctx, x.shape[0], _dptr(x), _dptr(v)))

    def add_body(self, x_local, dynamic, density=1000.0, position=(0, 0, 0), quat_wxyz=(1, 0, 0, 0)) -> int:
        x = _f64(x_local, (-1, 3))
        pos = _f64(position, (3,))
        q = _f64(quat_wxyz, (4,))
        return self._check(self._fn("add_body")(self._ctx, x.shape[0], _dptr(x), int(bool(dynamic)), float(density), _dptr(pos), _dptr(q)))

    def add_emitter(self, width, height, position, rotation, velocity, emit_start=0.0, emit_end=1e300):
        pos = _f64(position, (3,))
        rot = _f64(rotation, (9,))
        self._check(self._fn("add_emitter")(self._ctx, int(width), int(height), _dptr(pos), _dptr(rot), float(velocity), float(emit_start), float(emit_end)))

    def set_init_v_omega(self, body, v0, omega0):
        v0 = _f64(v0, (3,))
        w0 = _f64(omega0, (3,))
        self._check(self._fn("set_init_v_omega")(self._ctx, int(body), _dptr(v0), _dptr(w0)))

    def slab_unique_id(self) -> bytes:
        """NCCL unique id for `slab_configure` (call on rank 0, broadcast the bytes to the other ranks)."""
        buf = C.create_string_buffer(128)
        rc = self._fn("slab_unique_id")(buf)
        if rc != 0:
            raise DfrError(rc, "dfr_slab_unique_id failed (NCCL not loadable?)")
        return buf.raw

    def slab_configure(self, rank, n_ranks, id_bytes):
        """Make this context slab `rank` of `n_ranks` of the scene (before finalize; every rank builds the same scene)."""
        self._check(self._fn("slab_configure")(self._ctx, int(rank), int(n_ranks), C.c_char_p(bytes(id_bytes))))

    def slab_info(self):
        out = (C.c_int64 * 4)()
        self._check(self._fn("slab_info")(self._ctx, out))
        return {"owned": out[0], "ghosts": out[1], "exchanged_bytes": out[2], "slabs": abs(out[3]),
                "transport": "peer stores (cudaIpc over NVLink)" if out[3] < 0 else "nccl send/recv"}

    def finalize(self):
        self._check(self._fn("finalize")(self._ctx))

    def slab_local_ids(self):
        """Particle ids (rows of the scene's arrays) this context held at t = 0: the rows load_fluid_state_local takes."""
        n = int(self._fn("slab_local_ids")(self._ctx, None, 0))
        ids = np.zeros(n, dtype=np.int32)
        self._fn("slab_local_ids")(self._ctx, ids.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return ids

    def load_fluid_state_local(self, x=None, v=None, kappa=None, kappa_v=None):
        x, v, k, kv = _f64(x), _f64(v), _f64(kappa), _f64(kappa_v)
        n = int(self._fn("slab_local_ids")(self._ctx, None, 0))
        for name, a, width in (("x", x, 3), ("v", v, 3), ("kappa", k, 1), ("kappa_v", kv, 1)):
            if a is not None and a.size != width * n:
                raise ValueError(f"load_fluid_state_local: {name} holds {a.size} values, this context's {n} rows need {width * n}")
        self._check(self._fn("load_fluid_state_local")(self._ctx, n, _dptr(x), _dptr(v), _dptr(k), _dptr(kv)))

    def load_fluid_state(self, x=None, v=None, kappa=None, kappa_v=None):
        x, v, k, kv = _f64(x), _f64(v), _f64(kappa), _f64(kappa_v)
        # dfr_load_fluid_state takes no count: it reads the scene's initial particle count from every array it is given
        n = int(self._fn("num_fluid_initial")(self._ctx))
        for name, a, width in (("x", x, 3), ("v", v, 3), ("kappa", k, 1), ("kappa_v", kv, 1)):
            if a is not None and a.size != width * n:
                raise ValueError(f"load_fluid_state: {name} holds {a.size} values, the scene's {n} fluid particles need {width * n}")
        self._check(self._fn("load_fluid_state")(self._ctx, _dptr(x), _dptr(v), _dptr(k), _dptr(kv)))

    # -- stepping -------------------------------------------------------------------------
    def reset(self):
        self._check(self._fn("reset")(self._ctx))

    def reset_gradient(self):
        self._check(self._fn("reset_gradient")(self._ctx))

    def set_gradient_mode(self, mode):
        self._check(self._fn("set_gradient_mode")(self._ctx, int(mode)))

    def step(self, n=1):
        self._check(self._fn("step")(self._ctx, int(n)))

    def run_trajectory(self, max_steps=1 << 30) -> int:
        done = C.c_int(0)
        self._check(self._fn("run_trajectory")(self._ctx, int(max_steps), C.byref(done)))
        return done.value

    def step_info(self) -> StepInfo:
        info = StepInfo()
        self._check(self._fn("get_step_info")(self._ctx, C.byref(info)))
        return info

    def device_time_ms(self):
        ms = C.c_double(0)
        n = C.c_int64(0)
        self._check(self._fn("get_device_time_ms")(self._ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def set_profiling(self, enable=True):
        self._check(self._fn("set_profiling")(self._ctx, int(bool(enable))))

    def kernel_profile(self):
        """{kernel name: (total ms, launches)} since profiling was enabled."""
        rows = {}
        buf = C.create_string_buffer(256)
        ms, n = C.c_double(0), C.c_int64(0)
        i = 0
        while self._fn("get_kernel_profile")(self._ctx, i, buf, 256, C.byref(ms), C.byref(n)) == 0:
            rows[buf.value.decode()] = (ms.value, n.value)
            i += 1
        return rows

    # -- state access ---------------------------------------------------------------------
    @property
    def num_fluid(self) -> int:
        return int(self._fn("num_fluid")(self._ctx))

    @property
    def num_bodies(self) -> int:
        return int(self._fn("num_bodies")(self._ctx))

    def num_body_particles(self, body) -> int:
        return int(self._fn("num_body_particles")(self._ctx, int(body)))

    def body_state(self, body):
        out = np.zeros(13)
        self._check(self._fn("get_body_state")(self._ctx, int(body), _dptr(out)))
        return {"x": out[0:3].copy(), "q": out[3:7].copy(), "v": out[7:10].copy(), "omega": out[10:13].copy()}

    def set_body_velocity(self, body, v=None, omega=None):
        v, w = _f64(v), _f64(omega)
        self._check(self._fn("set_body_velocity")(self._ctx, int(body), _dptr(v), _dptr(w)))

    def body_properties(self, body):
        out = np.zeros(17)
        self._check(self._fn("get_body_properties")(self._ctx, int(body), _dptr(out)))
        return {"mass": out[0], "inv_mass": out[1], "inertia0": out[2:11].reshape(3, 3).copy(), "force": out[11:14].copy(), "torque": out[14:17].copy()}

    def body_grad(self, body, which):
        if isinstance(which, str):
            which = GRAD_NAMES.index(which)
        out = np.zeros(12)
        self._check(self._fn("get_body_grad")(self._ctx, int(body), int(which), _dptr(out)))
        shape = _GRAD_SHAPE.get(which, (3, 3))
        return out[: shape[0] * shape[1]].reshape(shape).copy()

    def manager_grad(self, R, RR, which):
        if isinstance(which, str):
            which = GRAD_NAMES.index(which)
        out = np.zeros(12)
        self._check(self._fn("get_manager_grad")(self._ctx, int(R), int(RR), int(which), _dptr(out)))
        shape = _GRAD_SHAPE.get(which, (3, 3))
        return out[: shape[0] * shape[1]].reshape(shape).copy()

    def fluid(self, field):
        fid, w = FLUID_FIELDS[field]
        n = self.num_fluid
        out = np.zeros((n, w)) if w > 1 else np.zeros(n)
        if n:
            self._check(self._fn("download_fluid")(self._ctx, fid, _dptr(out)))
        return out

    def body_particles(self, body, field):
        fid, w = BODY_FIELDS[field]
        n = self.num_body_particles(body)
        out = np.zeros((n, w)) if w > 1 else np.zeros(n)
        if n:
            self._check(self._fn("download_body")(self._ctx, int(body), fid, _dptr(out)))
        return out

    def neighbors(self, set_a=-1, set_b=-1):
        """CSR neighbour sets (counts, indices) of set_a -> set_b in particle-id space; -1 = fluid."""
        na = self.num_fluid if set_a < 0 else self.num_body_particles(set_a)
        counts = np.zeros(max(na, 1), dtype=np.int32)
        total = C.c_int64(0)
        ip = C.POINTER(C.c_int32)
        self._check(self._fn("get_neighbors")(self._ctx, int(set_a), int(set_b), counts.ctypes.data_as(ip), None, 0, C.byref(total)))
        idx = np.zeros(max(total.value, 1), dtype=np.int32)
        self._check(self._fn("get_neighbors")(self._ctx, int(set_a), int(set_b), counts.ctypes.data_as(ip), idx.ctypes.data_as(ip), idx.size, C.byref(total)))
        return counts[:na], idx[: total.value]
