"""dev: is a run reproducible from a snapshot?  A = continue, B = import + run, C = import + run again."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
from dbox_b200 import _abi as A, lib, scenes

api = lib.api()
DT = 1.0 / 60.0


def arr(w):
    buf, nb = w.read_bodies()
    return np.frombuffer(buf, dtype=np.float32, count=nb * 29).reshape(nb, 29).copy()


def run(n, columns, joints, circles, steps=12, settle=200, cont=True, collide_first=False):
    w, bodies, nj = scenes.pile(api=api, n=n, columns=columns, joints=joints, circles=circles)
    w.SetAllowSleeping(False)
    if not cont:
        w.SetContinuousPhysics(False)
    w.StepN(DT, 8, 3, settle)
    if collide_first:
        api.world_stage_collide(w._w)
    need = api.world_export_state(w._w, None, 0)
    blob = (C.c_char * need)()
    api.world_export_state(w._w, blob, need)
    w.StepN(DT, 8, 3, steps); a = arr(w); ca = w.counts()
    api.world_import_state(w._w, blob, need)
    w.StepN(DT, 8, 3, steps); b = arr(w); cb = w.counts()
    api.world_import_state(w._w, blob, need)
    w.StepN(DT, 8, 3, steps); c = arr(w); cc = w.counts()

    def d(x, y):
        m = (x.view(np.uint32) != y.view(np.uint32)).any(1)
        return int(m.sum()), float(np.abs(x[:, 2:4] - y[:, 2:4]).max())
    print("collide_first=%s " % collide_first, end="")
    print("n=%d joints=%s circles=%s toi=%s steps=%d: A-B %s  B-C %s  contacts %d/%d/%d toiEvents?" % (n, joints, circles, cont, steps, d(a, b), d(b, c), ca.contacts, cb.contacts, cc.contacts), flush=True)
    w.close()


for cf in (False, True):
    run(100000, 1000, True, True, 12, collide_first=cf)
    run(100000, 1000, True, True, 3, collide_first=cf)
    run(20000, 200, True, True, 12, collide_first=cf)
    run(20000, 200, False, False, 12, collide_first=cf)
