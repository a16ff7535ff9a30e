"""Batched independent worlds across the GPUs of one box (BASELINE.json config 5).

The reference has no such thing: a dbox user who wants many worlds makes many `b2World`s and steps them one after the
other on host cores (`b2World` shares nothing between instances, dynamics/b2world.d:34-40).  Here the worlds of one rank
live as replicas inside ONE device world (`dbx_world_replicate`), so every kernel of the step runs over all of them at
once, and the batch is partitioned across ranks with no data-path collective: the only communication is the reduction
of the final statistics (`reduce_stats`, NCCL on GPUs, gloo in the CPU tests).
"""
import ctypes as C


def partition(n_worlds, rank, world_size):
    """(first, count) of the contiguous block of worlds rank `rank` owns; blocks differ in size by at most one"""
    if world_size < 1 or not (0 <= rank < world_size) or n_worlds < 0:
        raise ValueError("bad partition request")
    base, extra = divmod(n_worlds, world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def reduce_stats(local, maxima=("ms", "seconds")):
    """sum every entry of `local` over the ranks, except the keys in `maxima`, which take the maximum (times are the
    slowest rank's).  No-op without an initialised process group.  Returns a new dict, identical on all ranks."""
    import torch
    import torch.distributed as dist
    keys = sorted(local)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: float(local[k]) for k in keys}
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    sums = torch.tensor([0.0 if k in maxima else float(local[k]) for k in keys], dtype=torch.float64, device=dev)
    maxs = torch.tensor([float(local[k]) if k in maxima else 0.0 for k in keys], dtype=torch.float64, device=dev)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    return {k: float(maxs[i] if k in maxima else sums[i]) for i, k in enumerate(keys)}


def run_pipelined(api, w, n_bodies, dt, velocity_iterations, position_iterations, n, force_ptrs, out_ptrs):
    """the act / step / observe loop on the pipelined calls of the ABI (include/dbox_b200.h "pipelined stepping and bulk I/O")"""
    def ck(rc):
        if rc < 0:
            raise RuntimeError("ABI call failed with %d %s" % (rc, api.last_error().decode()))
        return rc
    prev = 0
    for k in range(n):
        ck(api.world_apply_forces_async(w, force_ptrs[k & 1], n_bodies))
        ck(api.world_step_async(w, dt, velocity_iterations, position_iterations))
        ticket = ck(api.world_read_transforms_async(w, out_ptrs[k & 1], n_bodies))
        if prev:
            ck(api.world_io_wait(w, prev))          # observation k - 1 is on the host; its buffer is reused at k + 1
        prev = ticket
    if prev:
        ck(api.world_io_wait(w, prev))
    ck(api.world_sync(w))


class WorldBatch:
    """`n_worlds` copies of the world `build(api=..., caps=..., device=...)` returns, this rank's share on its GPU.

    build      callable returning a dbox_b200.world.b2World (or a tuple whose first element is one)
    body r * bodies_per_world + b of this rank is body b of its r-th world (global world index first + r)."""

    def __init__(self, build, n_worlds, rank=0, world_size=1, device=0, api=None, contacts_per_world=None):
        from . import _abi as A
        from . import lib
        self.api = api if api is not None else lib.api()
        self.n_worlds, self.rank, self.world_size = n_worlds, rank, world_size
        self.first, self.count = partition(n_worlds, rank, world_size)
        if self.count < 1:
            raise ValueError("rank %d of %d has no world to step (n_worlds=%d)" % (rank, world_size, n_worlds))
        caps = A.Caps()
        if contacts_per_world:
            caps.maxContacts = int(self.count * contacts_per_world)
        made = build(api=self.api, caps=caps, device=device)
        self.world = made[0] if isinstance(made, tuple) else made
        self.bodies_per_world = self.world.counts().bodies
        self.world.Replicate(self.count)
        self.n_bodies = self.bodies_per_world * self.count

    def step(self, dt, velocity_iterations, position_iterations, n=1):
        self.world.StepN(dt, velocity_iterations, position_iterations, n)

    def set_states(self, pose=None, vel=None, ids=None):
        self.world.SetBodyStates(ids, pose, vel)

    def apply_forces(self, host_ptr):
        n = self.api.world_apply_forces(self.world._w, host_ptr, self.n_bodies)
        if n != self.n_bodies:
            raise RuntimeError(self.api.last_error())

    def read_transforms(self, host_ptr):
        n = self.api.world_read_transforms(self.world._w, host_ptr, self.n_bodies)
        if n != self.n_bodies:
            raise RuntimeError(self.api.last_error())

    def run_pipelined(self, dt, velocity_iterations, position_iterations, n, force_ptrs, out_ptrs):
        """n iterations of act -> step -> observe with the copies on the copy streams beside the steps
        (dbx_world_apply_forces_async / step_async / read_transforms_async): iteration k uploads force_ptrs[k % 2], steps, and
        snapshots + downloads the transforms into out_ptrs[k % 2]; the host waits for observation k - 1 while step k runs.
        Both pointer pairs are pinned host buffers of n_bodies * 16 bytes."""
        run_pipelined(self.api, self.world._w, self.n_bodies, dt, velocity_iterations, position_iterations, n, force_ptrs, out_ptrs)

    def time_steps(self, dt, velocity_iterations, position_iterations, n, flush_l2=True):
        """n steps timed with CUDA events on the world's stream; returns (total ms, 9 per-stage ms)"""
        tot = C.c_float()
        stage = (C.c_float * 9)()
        rc = self.api.world_time_steps(self.world._w, dt, velocity_iterations, position_iterations, n, 1 if flush_l2 else 0, C.byref(tot), stage)
        if rc < 0:
            raise RuntimeError(self.api.last_error())
        return float(tot.value), [float(x) for x in stage]

    def stats(self):
        c = self.world.counts()
        return {"worlds": self.count, "bodies": c.bodies, "contacts": c.contacts, "touching": c.touching, "awake_bodies": c.awakeBodies}

    def close(self):
        self.world.close()
