"""Slab decomposition check, run as one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/slab_check.py
Every rank builds the same dam-break scene and becomes one slab of it; rank 0 also runs the plain single-context path on
the same scene.  After every step the merged slab state must equal the single-context state up to summation order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from difffr_b200 import scenes  # noqa: E402
from difffr_b200.cabi import Context  # noqa: E402

FIELDS = ["position", "velocity", "density", "factor", "kappa", "kappa_v", "density_adv", "acceleration", "sum_grad_p_k"]


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def main():
    n_particles = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    manager = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    burst = int(sys.argv[4]) if len(sys.argv) > 4 else 0  # then `bursts` calls of dfr_step(burst)
    bursts = int(sys.argv[5]) if len(sys.argv) > 5 else (2 if burst else 0)
    row_capacity = int(sys.argv[6]) if len(sys.argv) > 6 else 0  # initial ELL row capacities (0: defaults); small values make the lists grow
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")  # control plane only (id broadcast, merging the parity dumps)
    sc = scenes.dam_break_scene(n_particles, n_boxes=2, jitter=0.2, seed=3)
    for b in (1, 2):
        sc["bodies"][b]["init_v"] = (0.3 * b, -0.2, 0.1)
        sc["bodies"][b]["init_omega"] = (0.5, 1.0 * b, -0.4)
    kw = dict(surface_tension_method=2, surface_tension=0.2, max_error=0.05, target_time=10.0, use_rigid_gradient_manager=manager)
    if row_capacity:
        kw.update(neighbor_capacity_fluid=row_capacity, neighbor_capacity_boundary=max(8, row_capacity // 2))

    def factory(**k):
        ctx = Context(device=local, **k)
        if rank == 0:
            ident = torch.frombuffer(bytearray(ctx.slab_unique_id()), dtype=torch.uint8).clone()
        else:
            ident = torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(ident, 0)
        ctx.slab_configure(rank, world, bytes(ident.numpy().tobytes()))
        return ctx

    slab = scenes.build_context(factory, sc, **kw)
    single = scenes.build_context(lambda **k: Context(device=local, **k), sc, **kw) if rank == 0 else None
    n = slab.num_fluid
    assert n == len(sc["fluid"])
    worst = 0.0
    def advance_and_compare(s, k):
        """k steps in one dfr_step call on both sides (k > 1: replayed steps back to back, nothing read back in between)"""
        nonlocal worst
        try:
            slab.step(k)
        except Exception as e:
            raise RuntimeError(f"rank {rank}: slab step {s} (x{k}) failed: {e}") from e
        info = slab.step_info()
        owned = torch.tensor([info.num_fluid_particles], dtype=torch.int64)
        dist.all_reduce(owned)
        assert int(owned) == n, (int(owned), n)  # every particle is owned by exactly one slab
        merged = {}
        for f in FIELDS:
            a = torch.from_numpy(slab.fluid(f))  # zero except for the ids this rank owns
            dist.all_reduce(a)
            merged[f] = a.numpy()
        if rank == 0:
            single.step(k)
            i1 = single.step_info()
            assert (info.iterations, info.iterations_v) == (i1.iterations, i1.iterations_v), (s, info.iterations, i1.iterations, info.iterations_v, i1.iterations_v)
            assert abs(info.time_step_size - i1.time_step_size) <= 1e-12 * i1.time_step_size
            for f in FIELDS:
                e = rel(merged[f], single.fluid(f))
                worst = max(worst, e)
                assert e < 1e-9, (s, f, e)
            for b in (1, 2):
                sa, sb = slab.body_state(b), single.body_state(b)
                for kk in sa:
                    e = rel(sa[kk], sb[kk])
                    worst = max(worst, e)
                    assert e < 1e-9, (s, b, kk, e)
                for w in range(16):
                    e = rel(slab.body_grad(b, w), single.body_grad(b, w))
                    assert e < 1e-7, (s, b, w, e)
                    if manager:
                        e = rel(slab.manager_grad(b, b, w), single.manager_grad(b, b, w))
                        assert e < 1e-7, (s, b, w, e)
        # rigid bodies are replicated: identical bits on every rank
        st = torch.from_numpy(np.concatenate([slab.body_state(1)[kk] for kk in ("x", "q", "v", "omega")]))
        lo, hi = st.clone(), st.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi)

    for s in range(steps):
        if s == steps // 2:  # reset in the middle: initial distribution is restored
            slab.reset()
            if single:
                single.reset()
        if s == steps // 2 + 1:  # load a state (id order, whole scene on every rank): each slab keeps its own rows
            rng = np.random.default_rng(17)
            st_x = sc["fluid"] + rng.uniform(-0.1, 0.1, size=sc["fluid"].shape) * sc["radius"]
            st_v = rng.normal(scale=0.1, size=sc["fluid"].shape)
            st_k = -1e-6 * rng.random(n)
            st_kv = -1e-3 * rng.random(n)
            if steps % 2 == 0:
                slab.load_fluid_state(st_x, st_v, st_k, st_kv)
            else:  # the per-rank form: only the rows this slab held at t = 0, in the order dfr_slab_local_ids gives
                ids = slab.slab_local_ids()
                slab.load_fluid_state_local(st_x[ids], st_v[ids], st_k[ids], st_kv[ids])
            if single:
                single.load_fluid_state(st_x, st_v, st_k, st_kv)
        advance_and_compare(s, 1)
    for r in range(bursts):  # several steps per call
        advance_and_compare(steps + r, burst)
    si = slab.slab_info()
    print(f"rank {rank}: owned {si['owned']} ghosts {si['ghosts']} exchanged {si['exchanged_bytes'] / 1e6:.1f} MB ({si['transport']}) over {steps - steps // 2 + burst * bursts} steps; worst rel diff {worst:.2e}", flush=True)
    dist.barrier()
    if rank == 0:
        print("SLAB_CHECK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:  # a failing rank must take the job down at once: the others would wait in a collective
        import traceback

        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
