import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from dbox_b200 import lib
from oracle import orc
import test_gpu_features as T
from dbox_b200.world import *
from parity import contacts_by_key
ga, oa = lib.api(), orc.api()
def build(api):
    w = b2World((0.0, -10.0), api=api)
    T._ground(w, api)
    out = [T._box_body(w, api, -6.0 + 3.0 * k, 0.52) for k in range(5)]
    tops = [T._box_body(w, api, -6.0 + 3.0 * k, 1.55) for k in range(5)]
    return w, out + tops
wg, bg = build(ga); wo, bo = build(oa)
for k in range(75):
    for w, bs in ((wg, bg), (wo, bo)):
        if k == 40: bs[0].SetType(b2_staticBody); bs[6].SetActive(False)
        if k == 70:
            bs[0].SetType(b2_dynamicBody); bs[6].SetTransform((-3.0, 4.0), 0.3); bs[6].SetActive(True)
            bs[2].SetType(b2_kinematicBody); bs[2].SetLinearVelocity((0.5, 0.0))
    wg.Step(1/60., 8, 3); wo.Step(1/60., 8, 3)
    if k >= 69:
        kg, _, _ = contacts_by_key(wg); ko, _, _ = contacts_by_key(wo)
        print(k, "gpu-only", sorted(set(kg) - set(ko)), "oracle-only", sorted(set(ko) - set(kg)), wg.counts().proxies, wo.counts().proxies)
