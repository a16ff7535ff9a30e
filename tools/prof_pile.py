"""ncu target: the bench's 100k-body pile, settled, then `steps` steps between cudaProfilerStart/Stop.
   ncu --profile-from-start off ... python tools/prof_pile.py 2"""
import sys
sys.path.insert(0, '.')
import torch
from dbox_b200 import scenes, lib
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
api = lib.api()
w, b, nj = scenes.pile(api=api, n=100000, columns=1000)
w.SetAllowSleeping(False)
w.StepN(1 / 60., 8, 3, 600)
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
w.StepN(1 / 60., 8, 3, steps)
torch.cuda.synchronize()
rt.cudaProfilerStop()
c = w.counts()
print("contacts", c.contacts, "touching", c.touching, "colours", c.colours)
