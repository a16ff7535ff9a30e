"""ncu target: W pyramid worlds (randomised), settle, then profile `steps` steps between cudaProfilerStart/Stop.
   ncu --profile-from-start off --set full ... python tools/prof_batch.py 8192 2"""
import sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np
import torch
from dbox_b200 import scenes, lib
from dbox_b200.batch import WorldBatch
worlds = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
api = lib.api()
b = WorldBatch(scenes.pyramid, worlds, api=api, contacts_per_world=700)
b.world.SetAllowSleeping(False)
rng = np.random.RandomState(1)
vel = np.zeros((b.n_bodies, 4), np.float32); vel[:, :3] = rng.uniform(-0.5, 0.5, (b.n_bodies, 3))
b.set_states(vel=vel)
b.step(1 / 60., 8, 3, 100)
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
b.step(1 / 60., 8, 3, steps)
torch.cuda.synchronize()
rt.cudaProfilerStop()
print(b.stats())
