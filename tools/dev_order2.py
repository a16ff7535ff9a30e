"""dev: detailed dump for the device-order -> oracle hand-off on a tiny pile"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from dbox_b200 import lib, scenes, state, _abi as A
from oracle import orc
from tests.parity import hand_device_order_to_oracle, contact_key
from tests.test_gpu_full_size import _recs, _rel
ga, oa = lib.api(), orc.api()
DT = 1 / 60


def run(n, cols, settle, vi, pi, mode):
    wg, _, nj = scenes.pile(api=ga, n=n, columns=cols, joints=False, circles=False)
    wo, _, _ = scenes.pile(api=oa, n=n, columns=cols, joints=False, circles=False)
    for w in (wg, wo):
        w.SetAllowSleeping(False); w.SetContinuousPhysics(False)
    wg.StepN(DT, 8, 3, settle)
    snap = state.capture(wg)
    if mode != "noreapply":
        state.apply(wg, snap)
    state.apply(wo, snap)
    # pre-step comparison of what each side holds
    cg, ng = wg.read_contacts(); co, no = wo.read_contacts()
    dg = {contact_key(cg[i]): cg[i] for i in range(ng)}; do = {contact_key(co[i]): co[i] for i in range(no)}
    assert dg.keys() == do.keys()
    pre = max(abs(dg[k].manifold.points[j].normalImpulse - do[k].manifold.points[j].normalImpulse) for k in dg for j in range(2))
    wg.Step(DT, vi, pi)
    found, info = hand_device_order_to_oracle(oa, wg, wo)
    wo.Step(DT, vi, pi)
    bg, nb = wg.read_bodies(); bo, _ = wo.read_bodies()
    G, O = _recs(bg, nb, A.BodyState), _recs(bo, nb, A.BodyState)
    ev = np.maximum(np.maximum(_rel(G["v"]["x"], O["v"]["x"], 1.0), _rel(G["v"]["y"], O["v"]["y"], 1.0)), _rel(G["w"], O["w"], 1.0))
    w = int(ev.argmax())
    print("n=%d vi=%d pi=%d mode=%s: pre-step impulse diff %.3g, found=%d info=%s vel max %.3g bad %d worst %d  g=(%.6f %.6f %.6f) o=(%.6f %.6f %.6f)"
          % (n, vi, pi, mode, pre, found, info, ev.max(), int((ev > 1e-4).sum()), w, G["v"]["x"][w], G["v"]["y"][w], G["w"][w], O["v"]["x"][w], O["v"]["y"][w], O["w"][w]), flush=True)
    if ev.max() > 1e-4 and n <= 60:
        cg, ng = wg.read_contacts(); co, no = wo.read_contacts()
        cc = (C.c_int32 * ng)(); jc = (C.c_int32 * 1)(); inf = (C.c_int32 * 3)()
        ga.world_debug_read_solve_order(wg._w, cc, ng, jc, 0, inf)
        do = {contact_key(co[i]): co[i] for i in range(no)}
        for i in range(ng):
            r = cg[i]; o = do.get(contact_key(r))
            print("  contact", contact_key(r), "colour", cc[i], "pc", r.manifold.pointCount, "g imp", [round(r.manifold.points[j].normalImpulse, 6) for j in range(2)],
                  "o imp", [round(o.manifold.points[j].normalImpulse, 6) for j in range(2)] if o else None)
    wg.close(); wo.close()


run(30, 10, 200, 0, 0, "reapply")
run(30, 10, 200, 1, 0, "reapply")
run(30, 10, 200, 1, 0, "noreapply")
run(30, 10, 200, 8, 3, "reapply")
run(3000, 100, 300, 0, 0, "reapply")
run(3000, 100, 300, 1, 0, "reapply")
run(3000, 100, 300, 1, 0, "noreapply")
