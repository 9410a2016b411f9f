"""Independent-rollout sharding over GPUs.

The reference evaluates CMA-ES / (1+1)-ES populations and batched gradient evaluations serially in
one process because of its process-wide singletons (experiments/rigid_body_trajectory_optimization/
python/opt-ng.py:174-211; Simulation.cpp:35,167-188).  Here every rollout is its own `Context`, so a
population is split over the ranks of a `torch.distributed` job (one process per GPU) and nothing is
exchanged during a trajectory; only the per-rollout results (final rigid state and the eight
sensitivity blocks, < 1 KB) are gathered at the end.
"""
from __future__ import annotations

import numpy as np

RESULT_WIDTH = 13 + 4 * 9 + 2 * 12 + 2 * 9  # body state, d{x,v,omega}/d{v0,omega0} (6 x 3x3), dq/d{v0,omega0} (2 x 4x3)


def shard_indices(n_items: int, rank: int, world_size: int):
    """Contiguous block partition: the first `n_items % world_size` ranks get one extra item."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def pack_result(ctx, body: int) -> np.ndarray:
    """Final state + sensitivities of one rollout as a flat vector (what the optimiser consumes,
    gradient-based-optimize.py:170-208)."""
    s = ctx.body_state(body)
    parts = [s["x"], s["q"], s["v"], s["omega"]]
    for which in range(8):
        parts.append(ctx.body_grad(body, which).ravel())
    out = np.concatenate(parts)
    assert out.size == RESULT_WIDTH
    return out


def unpack_result(vec: np.ndarray):
    vec = np.asarray(vec, dtype=np.float64)
    o = 0

    def take(n, shape=None):
        nonlocal o
        a = vec[o:o + n]
        o += n
        return a.reshape(shape) if shape else a

    res = {"x": take(3), "q": take(4), "v": take(3), "omega": take(3)}
    for name, shape in (("grad_x_to_v0", (3, 3)), ("grad_x_to_omega0", (3, 3)), ("grad_quaternion_to_v0", (4, 3)),
                        ("grad_quaternion_to_omega0", (4, 3)), ("grad_v_to_v0", (3, 3)), ("grad_v_to_omega0", (3, 3)),
                        ("grad_omega_to_v0", (3, 3)), ("grad_omega_to_omega0", (3, 3))):
        res[name] = take(shape[0] * shape[1], shape)
    return res


def run_population(make_context, candidates, body: int, rank: int = 0, world_size: int = 1, max_steps: int = 1 << 30,
                   gather=None, concurrency: int = 1, contexts=None, stats=None):
    """Evaluate `candidates` (sequence of (v0, omega0) pairs for rigid body `body`), sharded over ranks.

    make_context() -> a finalized Context; `concurrency` of them are created per rank and reset between rollouts
    (dfr_reset is a device-to-device restore of the initial state held in HBM).  Contexts are independent (own device
    memory, own CUDA stream, no globals), so with concurrency > 1 the rank's candidates are evaluated by that many
    host threads, each driving its own context: the kernels of a paper-scale scene (~10^5 particles) fill only part of a
    B200, and rollouts running side by side on different streams fill the rest.  The result of a candidate does not
    depend on which context ran it or on what ran beside it (deterministic kernels, no shared state).
    `contexts`: reuse these finalized contexts instead of calling make_context (len(contexts) overrides concurrency).
    gather(local: np.ndarray[n_local, W]) -> np.ndarray[n_total, W] concatenates rank blocks in rank
    order; None means single process.  Returns (results [n, RESULT_WIDTH], steps [n]) on every rank.
    stats: optional dict; "kernel_launches" is incremented by the kernels this rank's rollouts launched.
    """
    mine = shard_indices(len(candidates), rank, world_size)
    local = np.zeros((len(mine), RESULT_WIDTH + 1))
    launches = np.zeros(len(mine), dtype=np.int64)
    if contexts is not None:
        concurrency = len(contexts)
    concurrency = max(1, min(int(concurrency), max(len(mine), 1)))

    def evaluate(ctx, k):
        v0, w0 = candidates[mine[k]]
        ctx.set_init_v_omega(body, v0, w0)
        ctx.reset()
        steps = ctx.run_trajectory(max_steps)
        local[k, :RESULT_WIDTH] = pack_result(ctx, body)
        local[k, RESULT_WIDTH] = steps
        launches[k] = ctx.device_time_ms()[1]  # counted since the reset that started this rollout

    if mine:
        ctxs = list(contexts[:concurrency]) if contexts is not None else [make_context() for _ in range(concurrency)]
        if concurrency == 1:
            for k in range(len(mine)):
                evaluate(ctxs[0], k)
        else:
            import queue
            import threading

            work = queue.SimpleQueue()
            for k in range(len(mine)):
                work.put(k)
            errors = []

            def worker(ctx):
                # the C-ABI calls release the GIL (ctypes), so the threads only serialise on the tiny Python parts
                try:
                    while True:
                        try:
                            k = work.get_nowait()
                        except queue.Empty:
                            return
                        evaluate(ctx, k)
                except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
                    errors.append(e)

            threads = [threading.Thread(target=worker, args=(c,)) for c in ctxs]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            if errors:
                raise errors[0]
    if stats is not None:
        stats["kernel_launches"] = stats.get("kernel_launches", 0) + int(launches.sum())
    allr = gather(local) if gather is not None else local
    return allr[:, :RESULT_WIDTH], allr[:, RESULT_WIDTH].astype(np.int64)


def torch_gather(world_size: int):
    """all_gather of variable-length row blocks with torch.distributed (NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    def gather(local: np.ndarray) -> np.ndarray:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
        counts = [torch.zeros_like(n) for _ in range(world_size)]
        dist.all_gather(counts, n)
        counts = [int(c.item()) for c in counts]
        width = local.shape[1]
        pad = max(max(counts), 1)
        buf = torch.zeros((pad, width), dtype=torch.float64, device=dev)
        if local.shape[0]:
            buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
        blocks = [torch.zeros_like(buf) for _ in range(world_size)]
        dist.all_gather(blocks, buf)
        return np.concatenate([b[:c].cpu().numpy() for b, c in zip(blocks, counts)], axis=0)

    return gather
