"""Synthetic scene construction (host side, numpy): fluid blocks, wall and box samplings.

These produce the inputs the hot path consumes.  The fluid lattice follows the reference's
`SimulatorBase::createFluidBlocks` rule (Simulator/SimulatorBase.cpp:1638-1735, denseMode 0) so
particle counts of the paper scenes reproduce exactly (e.g. 99x49x49 = 237,699 for
diff-stone-skipping.json).  Boundary samplings are regular (seed-free), standing in for the
reference's Poisson-disk sampler whose output is random (PoissonDiskSampling.cpp:125-135).
"""
from __future__ import annotations

import numpy as np


def fluid_block(start, end, radius):
    """createFluidBlocks, mode 0 (SimulatorBase.cpp:1644-1700)."""
    start = np.asarray(start, dtype=np.float64)
    end = np.asarray(end, dtype=np.float64)
    diam = 2.0 * radius
    diff = end - start
    steps = [int(round(diff[k] / diam)) - 1 for k in range(3)]
    steps = [max(s, 0) for s in steps]
    s = start + 2.0 * radius
    j, k, l = np.meshgrid(np.arange(steps[0]), np.arange(steps[1]), np.arange(steps[2]), indexing="ij")
    pts = np.stack([j.ravel() * diam, k.ravel() * diam, l.ravel() * diam], axis=1) + s
    return np.ascontiguousarray(pts)


def box_surface_samples(half_extent, spacing, inward=False):
    """Regular samples on the surface of an axis-aligned box centred at the origin (body frame)."""
    he = np.asarray(half_extent, dtype=np.float64)
    n = np.maximum(np.round(2.0 * he / spacing).astype(int), 1)
    axes = [np.linspace(-he[k], he[k], n[k] + 1) for k in range(3)]
    pts = []
    for ax in range(3):
        o = [k for k in range(3) if k != ax]
        a, b = np.meshgrid(axes[o[0]], axes[o[1]], indexing="ij")
        for sgn in (-1.0, 1.0):
            p = np.zeros((a.size, 3))
            p[:, ax] = sgn * he[ax]
            p[:, o[0]] = a.ravel()
            p[:, o[1]] = b.ravel()
            pts.append(p)
    pts = np.concatenate(pts, axis=0)
    # remove duplicates on edges/corners (deterministic: round to a fine lattice, keep first occurrence)
    key = np.round(pts / (spacing * 1e-3)).astype(np.int64)
    _, idx = np.unique(key, axis=0, return_index=True)
    return np.ascontiguousarray(pts[np.sort(idx)])


def sphere_surface_samples(r, spacing):
    """Fibonacci-lattice samples on a sphere of radius r (body frame), deterministic."""
    n = max(int(round(4.0 * np.pi * r * r / (spacing * spacing))), 8)
    i = np.arange(n) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    theta = np.pi * (1.0 + 5.0 ** 0.5) * i
    return np.ascontiguousarray(np.stack([r * np.cos(theta) * np.sin(phi), r * np.sin(theta) * np.sin(phi), r * np.cos(phi)], axis=1))


def quat_from_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    nrm = np.linalg.norm(axis)
    if nrm == 0.0:
        return np.array([1.0, 0, 0, 0])
    axis = axis / nrm
    s = np.sin(0.5 * angle)
    return np.array([np.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def dam_break_scene(n_target, n_boxes=4, radius=None, tank_aspect=(1.0, 0.5, 0.5), jitter=0.0, seed=0):
    """Synthetic dam break with dynamic rigid boxes.

    Tank L x 0.5L x 0.5L; the fluid column fills 40 % of the length up to 80 % of the height on the
    2r lattice; `n_boxes` dynamic boxes (density 500, edge 12 r) sit above the column.  If `radius`
    is None it is chosen so that the number of fluid particles is close to `n_target`.
    Returns a dict with `radius`, `fluid` (n,3), `bodies` (list of dicts).
    """
    ax, ay, az = tank_aspect

    def count_for(r, L):
        d = 2 * r
        fx = int(round(0.4 * L * ax / d)) - 1
        fy = int(round(0.8 * L * ay / d)) - 1
        fz = int(round((L * az - 2 * d) / d)) - 1
        return max(fx, 0) * max(fy, 0) * max(fz, 0)

    if radius is None:
        radius = 0.025
    # choose L for the target count at this radius
    lo, hi = 4 * radius, 4000 * radius
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if count_for(radius, mid) < n_target:
            lo = mid
        else:
            hi = mid
    L = hi
    d = 2 * radius
    tank_he = np.array([0.5 * L * ax, 0.5 * L * ay, 0.5 * L * az])
    tank_center = np.array([0.0, tank_he[1], 0.0])
    walls = box_surface_samples(tank_he, d)
    # fluid column in the -x end of the tank, one diameter off the walls
    f_start = np.array([-tank_he[0], 0.0, -tank_he[2] + d])
    f_end = np.array([-tank_he[0] + 0.4 * L * ax, 0.8 * L * ay, tank_he[2] - d])
    fluid = fluid_block(f_start, f_end, radius)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        fluid = fluid + rng.uniform(-jitter * radius, jitter * radius, size=fluid.shape)
    bodies = [dict(x_local=walls, dynamic=False, density=1000.0, position=tank_center, quat=np.array([1.0, 0, 0, 0]))]
    box_he = np.array([6 * radius] * 3)
    box_samples = box_surface_samples(box_he, d)
    col_top = f_end[1]
    for b in range(n_boxes):
        # boxes partly immersed in the top of the column so that coupling is active from step 1
        fx = (b % 2 + 0.5) / 2.0
        fz = (b // 2 % 2 + 0.5) / 2.0
        px = f_start[0] + (0.15 + 0.7 * fx) * (f_end[0] - f_start[0])
        pz = f_start[2] + (0.15 + 0.7 * fz) * (f_end[2] - f_start[2])
        py = col_top + (0.25 + 0.1 * (b // 4)) * box_he[1] + (b // 4) * 2.5 * box_he[1]
        bodies.append(dict(x_local=box_samples, dynamic=True, density=500.0, position=np.array([px, py, pz]),
                           quat=quat_from_axis_angle([0.3, 1.0, 0.2], 0.1 * (b + 1))))
    # fluid particles overlapping a box are removed (as the reference's scenes are authored non-overlapping)
    keep = np.ones(len(fluid), dtype=bool)
    for bd in bodies[1:]:
        rel = np.abs(fluid - bd["position"])
        keep &= ~np.all(rel < box_he + d, axis=1)
    fluid = np.ascontiguousarray(fluid[keep])
    return dict(radius=radius, fluid=fluid, bodies=bodies, tank_half_extent=tank_he, tank_center=tank_center)


def build_context(ctx_factory, scene, **cfg_overrides):
    """Create a context from a scene dict using `ctx_factory(**overrides)` (product or oracle binding)."""
    ctx = ctx_factory(particle_radius=scene["radius"], **cfg_overrides)
    ctx.set_fluid(scene["fluid"], scene.get("fluid_velocity"))
    for b in scene["bodies"]:
        ctx.add_body(b["x_local"], b["dynamic"], b["density"], b["position"], b["quat"])
    for e in scene.get("emitters", []):
        ctx.add_emitter(**e)
    for i, b in enumerate(scene["bodies"]):
        if b.get("init_v") is not None or b.get("init_omega") is not None:
            ctx.set_init_v_omega(i, b.get("init_v", (0, 0, 0)), b.get("init_omega", (0, 0, 0)))
    ctx.finalize()
    return ctx


def contact_scene(n_target=1500, radius=0.025, jitter=0.0, seed=0):
    """Two dynamic boxes within the contact range of the tank floor and of each other, in a shallow pool.

    Feeds the penalty rigid-rigid contact solver (BASELINE.json configs[2], `useRigidContactSolver`): a boundary
    particle is "in contact" as soon as a particle of another body lies within the support radius (4 r), and the
    penalty force acts while its artificial density exceeds the rest value (RigidContactSolver.cpp:307-345, 479-486).
    Box 1 sits 2 r above the floor, box 2 leans against box 1 (gap 2 r), both tilted so that many particles have
    distinct contact depths; both get initial linear and angular velocities so that the Coulomb friction term and
    the gyroscopic increment of every contacting particle are exercised.
    """
    sc = dam_break_scene(n_target, n_boxes=0, radius=radius, jitter=jitter, seed=seed)
    d = 2 * radius
    he = np.array([5 * radius] * 3)
    # irregular samples, like the reference's Poisson-disk ones: on a symmetric lattice the contact normal
    # x_r - sum(x_k w)/sum(w) of a face-centre particle is exactly zero and the reference's friction Jacobian divides by
    # |normal force| = 0 (RigidContactSolver.cpp:505) -> NaN on both sides
    rng_b = np.random.default_rng(1234 + seed)
    samples = box_surface_samples(he, d)
    samples = samples + rng_b.uniform(-0.3 * radius, 0.3 * radius, size=samples.shape)
    floor_y = 0.0
    x0 = -sc["tank_half_extent"][0] + 0.5 * (0.4 * 2 * sc["tank_half_extent"][0])
    pos1 = np.array([x0, floor_y + he[1] + 2.2 * radius, -1.5 * he[2]])
    pos2 = pos1 + np.array([0.0, 0.6 * radius, 2 * he[2] + 2.0 * radius])
    sc["bodies"].append(dict(x_local=samples, dynamic=True, density=1500.0, position=pos1, quat=quat_from_axis_angle([0.2, 1.0, 0.1], 0.05),
                             init_v=(0.4, -0.3, 0.2), init_omega=(1.5, -2.0, 0.7)))
    sc["bodies"].append(dict(x_local=samples, dynamic=True, density=800.0, position=pos2, quat=quat_from_axis_angle([1.0, 0.3, -0.2], 0.08),
                             init_v=(-0.2, -0.1, -0.5), init_omega=(-1.0, 0.5, 2.0)))
    keep = np.ones(len(sc["fluid"]), dtype=bool)
    for bd in sc["bodies"][1:]:
        keep &= ~np.all(np.abs(sc["fluid"] - bd["position"]) < he + d, axis=1)
    sc["fluid"] = np.ascontiguousarray(sc["fluid"][keep])
    return sc
