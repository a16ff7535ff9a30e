import sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np
from dbox_b200 import scenes, lib
from dbox_b200.batch import WorldBatch
worlds = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
api = lib.api()
b = WorldBatch(scenes.pyramid, worlds, api=api, contacts_per_world=700)
b.world.SetAllowSleeping(False)
rng = np.random.RandomState(1)
vel = np.zeros((b.n_bodies, 4), np.float32); vel[:, :3] = rng.uniform(-0.5, 0.5, (b.n_bodies, 3))
b.set_states(vel=vel)
b.step(1 / 60., 8, 3, 60)
buf = (C.c_uint64 * 4096)()
api.world_debug_phase_times(b.world._w, buf, 4096)
hb = (C.c_int32 * 64)()
for rep in range(3):
    api.world_debug_header(b.world._w, hb, 256); ev0 = hb[14]
    b.step(1 / 60., 8, 3, 1)
    n = api.world_debug_phase_times(b.world._w, buf, 4096)
    api.world_debug_header(b.world._w, hb, 256)
    tt = [buf[3000 + i] for i in range(64) if buf[3000 + i]]
    d = [(tt[i + 1] - tt[i]) / 1000. for i in range(len(tt) - 1)]
    print("events this step", hb[14] - ev0, "marks", len(d), "total us %.0f" % sum(d))
    print(" ".join("%.0f" % x for x in d))
