#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on B200: body-steps/s of the 100k-body pile at 8 velocity / 3 position
iterations (config 4), one world per GPU.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N ...            # the reference algorithm's CPU path (oracle port) on host cores

A "step" is one b2World.Step of the settled pile.  `value` is timed with CUDA events on the world's own stream with
all state resident in HBM; `e2e` drives the same steps through the public API with host buffers (per-step H2D of a
force array from pinned memory and D2H of all body transforms).  A single large world does not shard:
with N > 1 every rank steps its own replica of the pile ("replicas only", weak scaling, no data-path collective; NCCL
only reduces the timing).  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
UNIT = "body-steps/s"
DT, VEL_ITERS, POS_ITERS = 1.0 / 60.0, 8, 3


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def algorithmic_bytes_solve(ct, nb, n_rev, n_dist, iv, ip):
    """compulsory traffic of what the persistent solve kernel does in one launch: contact warm start
    (128), Iv velocity iterations (212 each), store impulses (32), Ip position iterations (136 each) per solver contact;
    integrate positions (48), write-back + transform (64) and sleep bookkeeping (24) per awake body; joints: prep + Iv*vel +
    Ip*pos (revolute 200/164/120, distance 150/116/112)."""
    contact = (128 + 212 * iv + 32 + 136 * ip) * ct
    body = (48 + 64 + 24) * nb
    joints = n_rev * (200 + 164 * iv + 120 * ip) + n_dist * (150 + 116 * iv + 112 * ip)
    return contact + body + joints


def algorithmic_bytes_step(c, n_rev, n_dist, iv, ip, proxies, pairs, rebuild_period=8):
    """SURVEY.md 8(d), the whole step with the run's own counts: bytes_step = 232 Nb + (88 | 192 with LBVH rebuild) Np + 16 pairs
    + 232 Cc + (568 + 212 Iv + 136 Ip) Ct + joints (prep + Iv vel + Ip pos)."""
    joints = n_rev * (200 + 164 * iv + 120 * ip) + n_dist * (150 + 116 * iv + 112 * ip)
    return (232 * c.awakeBodies + (88 + 104.0 / rebuild_period) * proxies + 16 * pairs + 232 * c.contacts
            + (568 + 212 * iv + 136 * ip) * c.touching + joints)


def time_gpu_steps(world, steps, vi, pi, flush=0):
    tot = C.c_float()
    rc = world._api.world_time_steps(world._w, DT, vi, pi, steps, flush, C.byref(tot), None)
    assert rc >= 0, world._api.last_error()
    return tot.value / steps


def bench_small_configs(api, oracle_api, budget_s=90.0):
    """BASELINE.md section 3's other rows, bounded in time: C1 hello_world and C2 Pyramid (latency-bound on a GPU: microseconds
    per step, informational), C3 Tumbler at growing populations -- each with the CPU port beside it on the same state
    (transplanted from the device world, so the CPU times the same scene without simulating its way there)."""
    from dbox_b200 import scenes, state
    out = {}
    t_start = time.time()

    def cpu_ms(world, steps, vi, pi):
        arr = (C.c_void_p * 1)(world._w)
        return 1e3 * oracle_api.batch_step(arr, 1, DT, vi, pi, steps, 1) / steps
    # C1
    wg, _ = scenes.hello_world(api=api); wo, _ = scenes.hello_world(api=oracle_api)
    wg.StepN(DT, 6, 2, 3)
    out["C1_hello_world"] = {"gpu_us_per_step": 1e3 * time_gpu_steps(wg, 60, 6, 2), "cpu_us_per_step": 1e3 * cpu_ms(wo, 60, 6, 2), "iters": "6v/2p", "steps": 60}
    wg.close(); wo.close()
    # C2: awake (first 150 steps) and after the pyramid has gone to sleep
    wg, _ = scenes.pyramid(api=api); wo, _ = scenes.pyramid(api=oracle_api)
    wg.StepN(DT, 8, 3, 20); wo.StepN(DT, 8, 3, 20)
    g_awake, c_awake = time_gpu_steps(wg, 130, 8, 3), cpu_ms(wo, 130, 8, 3)
    wg.StepN(DT, 8, 3, 250); wo.StepN(DT, 8, 3, 250)
    out["C2_pyramid"] = {"gpu_us_per_step_awake": 1e3 * g_awake, "cpu_us_per_step_awake": 1e3 * c_awake,
                         "gpu_us_per_step_asleep": 1e3 * time_gpu_steps(wg, 100, 8, 3), "cpu_us_per_step_asleep": 1e3 * cpu_ms(wo, 100, 8, 3),
                         "awake_bodies_at_end": [wg.counts().awakeBodies, wo.counts().awakeBodies], "iters": "8v/3p"}
    wg.close(); wo.close()
    # C3: Tumbler (container x5, mixed boxes / circles), populations reached by spawning 8 bodies per step
    rows = []
    tg = scenes.Tumbler(api=api, count=20000, scale=5.0)
    for target in (1000, 5000, 10000, 20000):
        if time.time() - t_start > budget_s:
            break
        while tg.m_count < target:
            tg.Step(DT, 8, 3, spawn_per_step=8)
        for _ in range(60):
            tg.world.Step(DT, 8, 3)
        ms = time_gpu_steps(tg.world, 40, 8, 3)
        row = {"bodies": tg.m_count, "gpu_steps_per_s": 1e3 / ms, "contacts": tg.world.counts().contacts}
        cpu_steps = 6 if target <= 5000 else 3
        to = scenes.Tumbler(api=oracle_api, count=20000, scale=5.0)
        to.spawn(tg.m_count)
        state.transplant(tg.world, to.world)
        row["cpu_steps_per_s"] = 1e3 / cpu_ms(to.world, cpu_steps, 8, 3)
        row["cpu_sample_steps"] = cpu_steps
        to.world.close()
        rows.append(row)
    tg.world.close()
    out["C3_tumbler"] = rows
    out["seconds"] = time.time() - t_start
    return out


def run_cpu_port(n_bodies, columns, settle, steps, warmup, threads=1, replicas=1):
    """the reference algorithm's CPU path (oracle/liborc.so, a C++ restatement of dbox): one world per thread"""
    from oracle import orc
    from dbox_b200 import scenes
    api = orc.api()
    worlds = []
    for r in range(replicas):
        w, _, _ = scenes.pile(api=api, n=n_bodies, columns=columns, seed=12345 + r)
        w.SetAllowSleeping(False)
        worlds.append(w)
    arr = (C.c_void_p * replicas)(*[w._w for w in worlds])
    api.batch_step(arr, replicas, DT, VEL_ITERS, POS_ITERS, settle + warmup, threads)
    secs = api.batch_step(arr, replicas, DT, VEL_ITERS, POS_ITERS, steps, threads)
    c = worlds[0].counts()
    # the reference's own b2Profile split (b2timestep.d:37-47) for the last timed step: where the CPU time goes
    p = worlds[0].GetProfile()
    c.profile_ms = {k: round(float(getattr(p, k)), 4) for k in ("step", "collide", "solve", "solveInit", "solveVelocity", "solvePosition", "broadphase", "solveTOI")}
    return replicas * n_bodies * steps / secs, secs, c


def run_cpu_pile_from_snapshot(n_bodies, columns, snap, steps, warmup, replicas=1):
    """The reference algorithm's CPU path on the REAL headline configuration: the oracle builds the same n_bodies pile, takes
    over the settled state the device produced (dbox_b200.state.apply -> orc_world_write_*: bodies, fat AABBs, contact cache with
    impulses, joint impulses) and steps it single-threaded.  No CPU settle (600 oracle steps of this world take ~10 minutes)."""
    from oracle import orc
    from dbox_b200 import scenes, state
    api = orc.api()
    worlds = []
    for _ in range(replicas):             # N > 1: one world per host thread, like our arm's one world per GPU
        w, _, _ = scenes.pile(api=api, n=n_bodies, columns=columns, seed=12345)
        w.SetAllowSleeping(False)
        state.apply(w, snap)
        worlds.append(w)
    arr = (C.c_void_p * replicas)(*[w._w for w in worlds])
    if warmup > 0:
        api.batch_step(arr, replicas, DT, VEL_ITERS, POS_ITERS, warmup, replicas)
    secs = api.batch_step(arr, replicas, DT, VEL_ITERS, POS_ITERS, steps, replicas)
    c = worlds[0].counts()
    p = worlds[0].GetProfile()
    c.profile_ms = {k: round(float(getattr(p, k)), 4) for k in ("step", "collide", "solve", "solveInit", "solveVelocity", "solvePosition", "broadphase", "solveTOI")}
    for w in worlds:
        w.close()
    return replicas * n_bodies * steps / secs, secs, c


def settled_pile_snapshot(args, device=0):
    """settle the headline pile on the device and return its state as ABI records (the reference arm's starting point)"""
    from dbox_b200 import lib, scenes, state
    api = lib.api()
    if api.device_count() < 1:
        return None
    world, _, _ = scenes.pile(api=api, n=args.bodies, columns=args.columns, seed=12345, device=device)
    world.SetAllowSleeping(False)
    world.StepN(DT, VEL_ITERS, POS_ITERS, args.settle)
    snap = state.capture(world)
    world.close()
    return snap


def probe_d_toolchain():
    """is there a D compiler on this box?  (None in the build image; recorded with every reference line.)"""
    import shutil
    found = {t: shutil.which(t) for t in ("dmd", "ldc2", "gdc", "dub")}
    return {k: v for k, v in found.items() if v} or None


def run_cpu_pyramids(worlds, settle, steps, threads):
    """C5 on the host: `worlds` independent Pyramid worlds, one world per thread at a time (oracle port of the reference)"""
    from oracle import orc
    from dbox_b200 import scenes
    api = orc.api()
    ws = []
    for r in range(worlds):
        w, _ = scenes.pyramid(api=api)
        w.SetAllowSleeping(False)
        ws.append(w)
    arr = (C.c_void_p * worlds)(*[w._w for w in ws])
    api.batch_step(arr, worlds, DT, VEL_ITERS, POS_ITERS, settle, threads)
    secs = api.batch_step(arr, worlds, DT, VEL_ITERS, POS_ITERS, steps, threads)
    return worlds * steps / secs, secs


def bench_batched(api, args, rank, world_size, local_rank, barrier, torch):
    """C5: args.worlds independent Pyramid worlds (strong scaling: the batch is partitioned across the ranks, no data-path
    collective; only the final statistics are reduced), each rank's share as replicas inside one device world.  Every
    world gets its own random initial velocities (the RL reset), so the worlds do not evolve in lockstep."""
    import numpy as np
    from dbox_b200 import _abi as A
    from dbox_b200 import scenes
    from dbox_b200.batch import WorldBatch, reduce_stats
    batch = WorldBatch(scenes.pyramid, args.worlds, rank, world_size, device=local_rank, api=api, contacts_per_world=700)
    batch.world.SetAllowSleeping(False)
    nb = batch.n_bodies
    rng = np.random.RandomState(777 + batch.first)
    vel = np.zeros((nb, 4), np.float32)
    vel[:, 0] = rng.uniform(-0.5, 0.5, nb); vel[:, 1] = rng.uniform(-0.5, 0.5, nb); vel[:, 2] = rng.uniform(-0.5, 0.5, nb)
    batch.set_states(vel=vel)
    del vel
    batch.step(DT, VEL_ITERS, POS_ITERS, args.batch_settle)
    flush = not args.no_l2_flush
    batch.time_steps(DT, VEL_ITERS, POS_ITERS, 3, flush)
    K = args.batch_steps
    l0 = api.world_launch_count(batch.world._w)
    barrier()
    ms, stage_ms = batch.time_steps(DT, VEL_ITERS, POS_ITERS, K, flush)
    barrier()
    launches = api.world_launch_count(batch.world._w) - l0
    # e2e: RL-style loop, per step H2D of one force/torque record per body and D2H of every body transform
    # (12-byte records, DBX_IO_COMPACT: forces (fx, fy, torque) in, poses (x, y, angle) out)
    assert api.world_set_io_format(batch.world._w, A.IO_COMPACT) == 0
    forces = torch.zeros((nb, 3), dtype=torch.float32).pin_memory()
    xf_out = torch.empty((nb, 3), dtype=torch.float32).pin_memory()
    Ke = min(K, 30)
    for _ in range(2):
        batch.apply_forces(forces.data_ptr()); batch.step(DT, VEL_ITERS, POS_ITERS); batch.read_transforms(xf_out.data_ptr())
    barrier()
    t0 = time.time()
    for _ in range(Ke):
        batch.apply_forces(forces.data_ptr())
        batch.step(DT, VEL_ITERS, POS_ITERS)
        batch.read_transforms(xf_out.data_ptr())
    barrier()
    sync_seconds = time.time() - t0
    # the same loop on the pipelined calls: the copies of step k ride on the copy streams beside steps k and k + 1
    forces2 = torch.zeros((nb, 3), dtype=torch.float32).pin_memory()
    xf_out2 = torch.empty((nb, 3), dtype=torch.float32).pin_memory()
    fp, op = (forces.data_ptr(), forces2.data_ptr()), (xf_out.data_ptr(), xf_out2.data_ptr())
    batch.run_pipelined(DT, VEL_ITERS, POS_ITERS, 2, fp, op)
    barrier()
    t0 = time.time()
    batch.run_pipelined(DT, VEL_ITERS, POS_ITERS, Ke, fp, op)
    barrier()
    e2e_seconds = time.time() - t0
    local = batch.stats()
    sleeping_on = None
    if not args.skip_extras:
        # second figure: the same batch with sleeping on -- the work collapses once the pyramids rest
        batch.world.SetAllowSleeping(True)
        batch.step(DT, VEL_ITERS, POS_ITERS, 300)
        ms_on, _ = batch.time_steps(DT, VEL_ITERS, POS_ITERS, 20, flush)
        sleeping_on = {"ms_per_step_rank0": ms_on / 20, "value_rank0_share": batch.count * 20 / (ms_on / 1e3), "unit": "world-steps/s",
                       "awake_bodies_rank0": int(batch.stats()["awake_bodies"]), "steps_after_enabling": 300}
    local.update(ms=ms, seconds=e2e_seconds, launches=launches, sync_seconds=sync_seconds)
    tot = reduce_stats(local, maxima=("ms", "seconds", "sync_seconds"))          # the only collective of the batched path: final statistics (NCCL)
    out = {"workload": "C5: %d independent Pyramid worlds (20-row, %d bodies each, per-world random initial velocities), 60 Hz, %dv/%dp, "
                       "sleeping off, partitioned over %d GPU(s)" % (args.worlds, batch.bodies_per_world, VEL_ITERS, POS_ITERS, world_size),
           "value": args.worlds * K / (tot["ms"] / 1e3), "unit": "world-steps/s", "scaling": "strong", "worlds": args.worlds,
           "worlds_rank0": batch.count, "steps": K, "settle_steps": args.batch_settle, "ms_per_step": tot["ms"] / K,
           "body_steps_per_s": tot["bodies"] * K / (tot["ms"] / 1e3),
           "counts": {"bodies": int(tot["bodies"]), "contacts": int(tot["contacts"]), "touching": int(tot["touching"])},
           "stage_ms_rank0": dict(zip(["collide", "islands", "colour_sort", "prepare", "solve", "sync_fixtures", "find_new_contacts", "toi", "clear_forces"], stage_ms)),
           "e2e": {"value": args.worlds * Ke / tot["seconds"], "unit": "world-steps/s", "h2d_bytes_per_step": 12 * int(tot["bodies"]),
                   "d2h_bytes_per_step": 12 * int(tot["bodies"]), "steps": Ke, "record": "12 B per body each way (DBX_IO_COMPACT: forces fx fy torque in, poses x y angle out)",
                   "mode": "pipelined act/step/observe (dbx_world_apply_forces_async / step_async / read_transforms_async): every step's H2D and D2H are inside the timed region, on copy streams beside the steps",
                   "synchronous_value": args.worlds * Ke / tot["sync_seconds"]},
           "gpu_launches": int(tot["launches"])}
    if sleeping_on is not None:
        out["sleeping_on"] = sleeping_on
    # k_solve_worlds (one CTA per replica, rows and bodies stay on chip between passes): compulsory HBM traffic is the rows
    # once (216 B), the impulses (16 + 32 B) and the sort entries (8 B) per solver contact, and 156 B per awake body
    # (velocity, position, transform, flags, sleep time in and out); DESIGN.md 7.2
    alg = 272 * local["touching"] + 156 * local["awake_bodies"]
    peak, peak_src = peaks()
    ach = alg / (stage_ms[4] * 1e-3) / 1e9 if stage_ms[4] > 0 else 0.0
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_solve_worlds_traffic.json")))
        traffic = int(t["dram_bytes_per_launch"] * local["touching"] / t["touching"])      # scaled from the profiled batch size
    except Exception:
        pass
    out["roofline"] = {"bound": "hbm", "kernel": "k_solve_worlds (rank 0)", "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                       "frac": ach / peak, "algorithmic_bytes_per_launch": alg, "kernel_ms": stage_ms[4], "traffic": traffic,
                       "limiter": "instruction issue: 58 % of peak issue slots, 15 % of DRAM throughput (profiles/r01e_ncu_batched8192_summary.txt)"}
    batch.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bodies", type=int, default=100000)
    ap.add_argument("--columns", type=int, default=1000)
    ap.add_argument("--settle", type=int, default=600, help="untimed steps that let the pile settle before warm-up (SURVEY.md 8(d))")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--cpu-bodies", type=int, default=10000, help="bounded CPU sample: same generator and pile depth, fewer columns")
    ap.add_argument("--cpu-steps", type=int, default=8, help="timed CPU steps of the full pile (each takes seconds: SolveTOI rescans the contact list per event)")
    ap.add_argument("--cpu-warmup", type=int, default=1)
    ap.add_argument("--cpu-settle", type=int, default=240)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="skip the sleeping-on second figures and the C1 / C2 / C3 rows")
    ap.add_argument("--worlds", type=int, default=65536, help="C5: independent Pyramid worlds in the batched leg (0 = skip the leg)")
    ap.add_argument("--batch-steps", type=int, default=100)
    ap.add_argument("--batch-settle", type=int, default=60)
    ap.add_argument("--cpu-batch-steps", type=int, default=400)
    ap.add_argument("--cpu-worlds-per-thread", type=int, default=32)
    ap.add_argument("--save-state", default=None, help="write the settled world state here (profiling runs reload it instead of settling)")
    ap.add_argument("--load-state", default=None)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world_size)
    W = max(args.warmup, 3)
    K = args.steps
    rows = (args.bodies + args.columns - 1) // args.columns
    cpu_cols = max(1, args.cpu_bodies // rows)
    cpu_bodies = cpu_cols * rows
    cores = os.cpu_count() or 1
    config = {"workload": "C4: %d-body box/circle pile with revolute + distance joint chains, 60 Hz, %dv/%dp, sleeping off, one world per GPU"
                          % (args.bodies, VEL_ITERS, POS_ITERS),
              "bodies": args.bodies, "columns": args.columns, "settle_steps": args.settle,
              "parallelism": "replicas only (a single world does not shard); %d independent world(s)" % n_gpus,
              "l2": "256 MiB buffer overwritten between timed steps (outside the event brackets)" if not args.no_l2_flush else "no flush"}

    # ------------------------------------------------------------------ reference arm: the CPU path on host cores
    if args.impl == "reference":
        if rank != 0:
            return 0
        t0 = time.time()
        steps = max(1, min(K, args.cpu_steps))
        snap = None
        try:
            snap = settled_pile_snapshot(args, local_rank)
        except Exception as e:      # no usable device: fall back to a CPU-settled sample and say so
            print("reference arm: device settle unavailable (%s)" % e, file=sys.stderr)
        if snap is not None:
            # same config as our arm: the 100,000-body pile in the state 600 settle steps left it in (settled on the device,
            # transplanted through orc_world_write_*), then timed on one host core -- the reference is single-threaded per world
            threads, same_config = max(1, min(n_gpus, cores)), True
            value, secs, c = run_cpu_pile_from_snapshot(args.bodies, args.columns, snap, steps, args.cpu_warmup, replicas=threads)
            sample = ("%d replica(s) of the full %d-body pile, settled %d steps on the device and transplanted into the CPU port (orc_world_write_*), "
                      "%d warm-up + %d timed steps, one world per host thread (the reference is single-threaded per world); ours steps one such world per GPU"
                      % (threads, args.bodies, args.settle, args.cpu_warmup, steps))
        else:
            threads, same_config = 1, False
            steps = max(1, min(K, 60))
            value, secs, c = run_cpu_port(cpu_bodies, cpu_cols, args.cpu_settle, steps, W, threads=1, replicas=1)
            config = dict(config, bodies=cpu_bodies, columns=cpu_cols, settle_steps=args.cpu_settle,
                          workload=config["workload"].replace("%d-body" % args.bodies, "%d-body (REDUCED: no device to settle the full pile)" % cpu_bodies))
            sample = ("%d-body pile (same generator, same %d-row depth, %d columns), %d settle + %d warm-up + %d timed steps on the CPU, single thread"
                      % (cpu_bodies, rows, cpu_cols, args.cpu_settle, W, steps))
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": args.cpu_warmup,
                "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "same_config": same_config,
                "counts": {"contacts": c.contacts, "touching": c.touching, "awake_bodies": c.awakeBodies, "joints": c.joints},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "b2Profile_ms_last_step": c.profile_ms},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "d_toolchain": probe_d_toolchain(), "wall_s": time.time() - t0}
        if args.worlds > 0:
            nw = cores * args.cpu_worlds_per_thread
            bsteps = args.cpu_batch_steps
            v, secs = run_cpu_pyramids(nw, args.batch_settle, bsteps, cores)
            line["batched"] = {"workload": "C5 sample: %d Pyramid worlds, one world per host thread at a time" % nw, "value": v, "unit": "world-steps/s",
                               "cores": cores, "kind": "port", "steps": bsteps, "settle_steps": args.batch_settle, "seconds": secs}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm: CUDA path through the C ABI
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: dbox_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's version / debug lines must not share stdout with the JSON line
        dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", local_rank))
    from dbox_b200 import _abi as A
    from dbox_b200 import lib, scenes
    api = lib.api()

    world, bodies, n_joints = scenes.pile(api=api, n=args.bodies, columns=args.columns, seed=12345 + rank, device=local_rank)
    world.SetAllowSleeping(False)
    if args.load_state:
        from dbox_b200 import state
        state.load(world, args.load_state)
    world.StepN(DT, VEL_ITERS, POS_ITERS, args.settle)
    if args.save_state and rank == 0:
        from dbox_b200 import state
        state.save(world, args.save_state)
    cpu_snap = None
    if rank == 0 and n_gpus == 1 and not args.skip_cpu_baseline:
        from dbox_b200 import state
        cpu_snap = state.capture(world)      # the settled pile: the CPU baseline below steps exactly this world

    tot = C.c_float()
    stage = (C.c_float * 9)()
    flush = 0 if args.no_l2_flush else 1

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (untimed)
    rc = api.world_time_steps(world._w, DT, VEL_ITERS, POS_ITERS, W, flush, C.byref(tot), stage)
    assert rc >= 0, api.last_error()
    launches0 = api.world_launch_count(world._w)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall = time.time()
    rc = api.world_time_steps(world._w, DT, VEL_ITERS, POS_ITERS, K, flush, C.byref(tot), stage)
    assert rc >= 0, api.last_error()
    barrier()
    wall = time.time() - t_wall
    clocks = sampler.stop()
    launches = api.world_launch_count(world._w) - launches0
    dev_ms = torch.tensor([tot.value], dtype=torch.float64, device="cuda")
    if world_size > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    total_ms = float(dev_ms.item())
    counts = world.counts()
    stage_ms = [float(x) for x in stage]

    # ---- e2e: same steps through the public API with host buffers (pinned), H2D + D2H inside the timed region
    n = args.bodies + 1
    assert api.world_set_io_format(world._w, A.IO_COMPACT) == 0      # 12-byte records: forces (fx, fy, torque) in, poses (x, y, angle) out
    forces = torch.zeros((n, 3), dtype=torch.float32).pin_memory()
    xf_out = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    Ke = min(K, 100)
    for _ in range(3):
        api.world_apply_forces(world._w, forces.data_ptr(), n); world.Step(DT, VEL_ITERS, POS_ITERS); api.world_read_transforms(world._w, xf_out.data_ptr(), n)
    barrier()
    t0 = time.time()
    for _ in range(Ke):
        assert api.world_apply_forces(world._w, forces.data_ptr(), n) == n
        world.Step(DT, VEL_ITERS, POS_ITERS)
        assert api.world_read_transforms(world._w, xf_out.data_ptr(), n) == n
    barrier()
    e2e_sync_s = torch.tensor([time.time() - t0], dtype=torch.float64, device="cuda")
    # the same loop on the pipelined calls (copies on the copy streams beside the steps); this is the e2e headline
    from dbox_b200.batch import run_pipelined
    forces2 = torch.zeros((n, 3), dtype=torch.float32).pin_memory()
    xf_out2 = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    fp, op = (forces.data_ptr(), forces2.data_ptr()), (xf_out.data_ptr(), xf_out2.data_ptr())
    run_pipelined(api, world._w, n, DT, VEL_ITERS, POS_ITERS, 3, fp, op)
    barrier()
    t0 = time.time()
    run_pipelined(api, world._w, n, DT, VEL_ITERS, POS_ITERS, Ke, fp, op)
    barrier()
    e2e_s = torch.tensor([time.time() - t0], dtype=torch.float64, device="cuda")
    if world_size > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_sync_s, op=dist.ReduceOp.MAX)
    e2e_value = n_gpus * args.bodies * Ke / float(e2e_s.item())
    e2e_sync_value = n_gpus * args.bodies * Ke / float(e2e_sync_s.item())

    # which island solver ran (tile solver: info[3] = tiles), and the second figure of BASELINE.md section 3: the same pile with sleeping ON
    info = (C.c_int32 * 4)()
    api.world_debug_read_solve_order(world._w, None, 0, None, 0, info)
    tiles = int(info[3])
    proxies, pairs = counts.proxies, counts.pairs
    sleeping_on = None
    if not args.skip_extras and rank == 0:
        world.SetAllowSleeping(True)
        world.StepN(DT, VEL_ITERS, POS_ITERS, 120)
        ms_on = time_gpu_steps(world, 50, VEL_ITERS, POS_ITERS, flush)
        c_on = world.counts()
        sleeping_on = {"ms_per_step": ms_on, "value": args.bodies / (ms_on / 1e3), "unit": UNIT, "awake_bodies": c_on.awakeBodies,
                       "note": "sleeping enabled 170 steps before the end of this window; a pile this deep keeps creeping, so only part of it rests"}
        world.SetAllowSleeping(False)
    js, nj = world.read_joints()
    n_rev = sum(1 for i in range(nj) if js[i].type == A.JOINT_REVOLUTE)
    n_dist = sum(1 for i in range(nj) if js[i].type == A.JOINT_DISTANCE)
    world.close()      # free the pile's device buffers before the batched leg
    batched = None
    if args.worlds > 0:
        batched = bench_batched(api, args, rank, world_size, local_rank, barrier, torch)

    if rank == 0:
        value = n_gpus * args.bodies * K / (total_ms / 1e3)
        peak, peak_src = peaks()
        solve_ms = stage_ms[4]
        alg = algorithmic_bytes_solve(counts.touching, counts.awakeBodies, n_rev, n_dist, VEL_ITERS, POS_ITERS)
        achieved = alg / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_solve_traffic.json" if tiles else "r01_solve_traffic.json")))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = tj.get("source", "profiles/r01d_ncu_pile100k_summary.txt (k_solve, round 1 build)")
        except Exception:
            pass
        alg_step = algorithmic_bytes_step(counts, n_rev, n_dist, VEL_ITERS, POS_ITERS, proxies, pairs)
        kernel = ("k_solve_tiles (tile solver, %d tiles: a tile's bodies, local rows (TMA bulk copies), boundary rows and joints in shared memory; " % tiles if tiles else "k_solve (persistent coloured Gauss-Seidel: ") + \
                 "warm start + %d velocity + %d position iterations + write-back)" % (VEL_ITERS, POS_ITERS)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": W,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "steps_per_s_per_world": K / (total_ms / 1e3),
                "counts": {"contacts": counts.contacts, "touching": counts.touching, "awake_bodies": counts.awakeBodies, "joints": counts.joints,
                           "colours": counts.colours, "islands": counts.islands},
                "stage_ms": dict(zip(["collide", "islands", "colour_sort", "prepare", "solve", "sync_fixtures", "find_new_contacts", "toi", "clear_forces"], stage_ms)),
                "roofline": {"bound": "hbm", "kernel": kernel,
                             "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                             "algorithmic_bytes_per_launch": alg, "kernel_ms": solve_ms, "traffic": traffic, "traffic_source": traffic_src},
                "roofline_step": {"bound": "hbm", "what": "the whole step: SURVEY.md 8(d) bytes_step with this run's counts over ms_per_step",
                                  "algorithmic_bytes_per_step": alg_step, "achieved": alg_step / (total_ms / K * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": alg_step / (total_ms / K * 1e-3) / 1e9 / peak},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 12 * n, "d2h_bytes_per_step": 12 * n, "steps": Ke,
                        "record": "12 B per body each way (DBX_IO_COMPACT: forces fx fy torque in, poses x y angle out)",
                        "mode": "pipelined act/step/observe (dbx_world_apply_forces_async / step_async / read_transforms_async): every step's H2D and D2H are inside the timed region, on copy streams beside the steps",
                        "synchronous_value": e2e_sync_value},
                "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": wall}
        if batched is not None:
            line["batched"] = batched
        if sleeping_on is not None:
            line["sleeping_on"] = sleeping_on
        if n_gpus == 1 and not args.skip_extras:
            from oracle import orc
            line["other_configs"] = bench_small_configs(api, orc.api())
        if n_gpus == 1 and not args.skip_cpu_baseline:
            if batched is not None:
                nw = cores * args.cpu_worlds_per_thread
                v, secs = run_cpu_pyramids(nw, args.batch_settle, args.cpu_batch_steps, cores)
                batched["cpu_baseline"] = {"value": v, "unit": "world-steps/s", "cores": cores, "kind": "port",
                                           "sample": "%d Pyramid worlds, %d settle + %d timed steps, one world per host thread at a time, %.1f s"
                                                     % (nw, args.batch_settle, args.cpu_batch_steps, secs)}
            t0 = time.time()
            csteps = max(1, min(args.cpu_steps, 3))
            v, secs, c = run_cpu_pile_from_snapshot(args.bodies, args.columns, cpu_snap, csteps, args.cpu_warmup)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "the full %d-body pile in the settled state this run timed (transplanted into the CPU port through orc_world_write_*), "
                                              "%d warm-up + %d timed steps, single thread, %.1f s" % (args.bodies, args.cpu_warmup, csteps, time.time() - t0),
                                    "b2Profile_ms_last_step": c.profile_ms}
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
