"""dev: replicas of a jointed scene (tests' crawler): world-local solver with joints vs the global solver (DBX_DEBUG=512)"""
import ctypes as C, os, sys
sys.path.insert(0, ".")
from dbox_b200 import _abi as A, lib
import tests.test_gpu_next_rows as T
api = lib.api()
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
units = int(sys.argv[2]) if len(sys.argv) > 2 else 1
caps = A.Caps(); caps.maxContacts = copies * 64 * units
w, bodies, joints = T._crawler_scene(api, units=units, caps=caps)
w.SetAllowSleeping(False)
w.Replicate(copies)
w.StepN(1 / 60., 8, 3, 60)
tot = C.c_float(); stage = (C.c_float * 9)()
api.world_time_steps(w._w, 1 / 60., 8, 3, 50, 1, C.byref(tot), stage)
c = w.counts()
print("units=%d " % units, end="")
print("DBX_DEBUG=%s replicas=%d bodies=%d joints=%d contacts=%d: %.3f ms/step = %.2f M world-steps/s; solve %.3f ms, stages %s" % (
    os.environ.get("DBX_DEBUG"), copies, c.bodies, c.joints, c.contacts, tot.value / 50, copies / (tot.value / 50) / 1e3, stage[4], [round(x, 3) for x in stage]))
