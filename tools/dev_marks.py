import sys, ctypes as C
sys.path.insert(0, '.')
from dbox_b200 import scenes, lib
ga = lib.api()
w, b, nj = scenes.pile(api=ga, n=100000, columns=1000)
w.SetAllowSleeping(False)
w.StepN(1 / 60., 8, 3, 600)
tot = C.c_float(); stage = (C.c_float * 9)()
for rep in range(3):
    ga.world_time_steps(w._w, 1 / 60., 8, 3, 100, 1, C.byref(tot), stage); a = tot.value / 100
    ga.world_time_steps(w._w, 1 / 60., 8, 3, 100, 1, C.byref(tot), None); b_ = tot.value / 100
    print("with stage marks %.4f ms   without %.4f ms" % (a, b_))
