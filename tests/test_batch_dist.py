"""N > 1 host logic on CPU: the batch of independent worlds is partitioned across ranks with no data-path collective and
only the final statistics are reduced (dbox_b200/batch.py).  Two gloo ranks step their share of a small batch with the
oracle (test infrastructure standing in for the device step, which needs a GPU) and the reduced statistics must equal
the single-process run of the whole batch."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dbox_b200.batch import partition  # noqa: E402

DT = 1.0 / 60.0


def test_partition_covers_every_world_once():
    for n in (0, 1, 7, 8, 65536, 65537):
        for ws in (1, 2, 3, 4, 8):
            seen = 0
            sizes = []
            for r in range(ws):
                first, count = partition(n, r, ws)
                assert first == seen
                seen += count
                sizes.append(count)
            assert seen == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition(4, 2, 2)


def _world_stats(first, count, steps):
    """step worlds [first, first+count) of the batch with the oracle; world g is a 6-row pyramid pushed sideways by g"""
    from oracle import orc
    from dbox_b200 import scenes
    api = orc.api()
    out = {"worlds": 0.0, "bodies": 0.0, "contacts": 0.0, "touching": 0.0, "checksum": 0.0, "seconds": 0.0}
    for g in range(first, first + count):
        w, bodies = scenes.pyramid(api=api, count=6)
        bodies[-1].SetLinearVelocity((0.25 * (g + 1), 0.0))
        for _ in range(steps):
            w.Step(DT, 8, 3)
        c = w.counts()
        out["worlds"] += 1; out["bodies"] += c.bodies; out["contacts"] += c.contacts; out["touching"] += c.touching
        p = bodies[-1].GetPosition()
        out["checksum"] += float(p.x) * (g + 1)
        out["seconds"] = max(out["seconds"], float(g))      # stands in for a per-rank time: reduced with MAX
    return out


def _rank_main(rank, world_size, port, n_worlds, steps, q):
    import torch.distributed as dist
    from dbox_b200.batch import partition, reduce_stats
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        first, count = partition(n_worlds, rank, world_size)
        local = _world_stats(first, count, steps)
        total = reduce_stats(local)
        q.put((rank, first, count, total))
    finally:
        dist.destroy_process_group()


def test_two_rank_batch_matches_single_process():
    import torch.multiprocessing as mp
    n_worlds, steps = 5, 40
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, n_worlds, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    assert [(g[1], g[2]) for g in got] == [(0, 3), (3, 2)]
    assert got[0][3] == got[1][3]                       # identical on all ranks
    whole = _world_stats(0, n_worlds, steps)
    red = got[0][3]
    for k in ("worlds", "bodies", "contacts", "touching"):
        assert red[k] == whole[k], k
    assert abs(red["checksum"] - whole["checksum"]) < 1e-9 * max(1.0, abs(whole["checksum"]))
    assert red["seconds"] == float(n_worlds - 1)        # MAX, not SUM
