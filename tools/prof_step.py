"""a few steps of the settled 100k pile for ncu launch lists (state saved by an earlier run is reused when present)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbox_b200 import scenes, lib, state
ga = lib.api()
w, b, nj = scenes.pile(api=ga, n=100000, columns=1000)
w.SetAllowSleeping(False)
p = "/tmp/pile100k_settled.pkl"      # (scratch on the GPU box: gpurun_out/ is copied back and capped at 64 MiB)
if os.path.exists(p):
    state.load(w, p)
    w.StepN(1 / 60., 8, 3, 20)
else:
    w.StepN(1 / 60., 8, 3, 600)
    state.save(w, p)
w.StepN(1 / 60., 8, 3, int(os.environ.get("STEPS", "4")))
print("done", w.counts().touching)
