"""dev: bisect the device-order -> oracle hand-off (tests/parity.hand_device_order_to_oracle) on small piles"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dbox_b200 import lib, scenes, state, _abi as A
from oracle import orc
from tests.parity import hand_device_order_to_oracle
from tests.test_gpu_full_size import _recs, _rel
ga, oa = lib.api(), orc.api()
DT = 1 / 60


def run(n, cols, settle, joints, circles, vi, pi, cont=False):
    wg, _, nj = scenes.pile(api=ga, n=n, columns=cols, joints=joints, circles=circles)
    wo, _, _ = scenes.pile(api=oa, n=n, columns=cols, joints=joints, circles=circles)
    for w in (wg, wo):
        w.SetAllowSleeping(False); w.SetContinuousPhysics(cont)
    wg.StepN(DT, 8, 3, settle)
    snap = state.capture(wg)
    state.apply(wg, snap); state.apply(wo, snap)
    wg.Step(DT, vi, pi)
    found, info = hand_device_order_to_oracle(oa, wg, wo)
    wo.Step(DT, vi, pi)
    bg, nb = wg.read_bodies(); bo, _ = wo.read_bodies()
    G, O = _recs(bg, nb, A.BodyState), _recs(bo, nb, A.BodyState)
    ev = np.maximum(np.maximum(_rel(G["v"]["x"], O["v"]["x"], 1.0), _rel(G["v"]["y"], O["v"]["y"], 1.0)), _rel(G["w"], O["w"], 1.0))
    ep = np.maximum(np.maximum(_rel(G["c"]["x"], O["c"]["x"], 1.0), _rel(G["c"]["y"], O["c"]["y"], 1.0)), _rel(G["a"], O["a"], 1.0))
    bad = int((ev > 1e-4).sum())
    print("n=%d joints=%s circles=%s vi=%d pi=%d cont=%s: found=%d info=%s  vel max %.3g (bodies > 1e-4: %d)  pos max %.3g  worst body %d"
          % (n, joints, circles, vi, pi, cont, found, info, ev.max(), bad, ep.max(), int(ev.argmax())), flush=True)
    wg.close(); wo.close()


for args in ((3000, 100, 300, False, False, 8, 3), (3000, 100, 300, False, True, 8, 3), (3000, 100, 300, True, False, 8, 3),
             (3000, 100, 300, True, True, 8, 3), (3000, 100, 300, True, True, 1, 0), (3000, 100, 300, True, True, 8, 0),
             (3000, 100, 300, True, True, 0, 3), (600, 30, 200, True, True, 8, 3), (3000, 100, 300, False, True, 8, 3, True)):
    run(*args)
