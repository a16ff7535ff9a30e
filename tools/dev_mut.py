import sys; sys.path.insert(0, '.')
from dbox_b200 import lib
from oracle import orc
import tests.test_gpu_features as T
from dbox_b200.world import *
from tests.parity import contacts_by_key
ga, oa = lib.api(), orc.api()
def build(api):
    w = b2World((0.0, -10.0), api=api)
    T._ground(w, api)
    bs = [T._box_body(w, api, -8.0 + 4.0 * k, 0.52) for k in range(5)]
    tops = [T._box_body(w, api, -8.0 + 4.0 * k, 1.55) for k in range(5)]
    return w, bs + tops
def edit(k, bs):
    if k == 30:
        bs[5].fixtures[0].SetFilterData(0x0002, 0x0000, 0); bs[1].fixtures[0].SetSensor(True)
        bs[7].SetGravityScale(-0.5); bs[7].SetLinearDamping(0.4); bs[8].SetMassData(3.0, (0.2, 0.0), 2.0)
        bs[9].SetFixedRotation(True); bs[9].SetAngularDamping(0.2); bs[4].fixtures[0].SetFriction(0.0); bs[4].fixtures[0].SetRestitution(0.5)
    if k == 60:
        bs[5].SetTransform((-8.0, 4.0), 0.0); bs[5].SetLinearVelocity((0.0, 0.0)); bs[5].fixtures[0].SetFilterData()
        bs[3].fixtures[0].SetDensity(4.0); bs[3].ResetMassData()
        bs[3].ApplyLinearImpulse((6.0, 0.0), (bs[3].GetPosition().x, bs[3].GetPosition().y + 0.4))
wg, bg = build(ga); wo, bo = build(oa)
for k in range(70):
    edit(k, bg); edit(k, bo)
    wg.Step(1/60., 8, 3); wo.Step(1/60., 8, 3)
    kg, _, _ = contacts_by_key(wg); ko, _, _ = contacts_by_key(wo)
    if set(kg) != set(ko): print(k, "gpu-only", sorted(set(kg) - set(ko)), "oracle-only", sorted(set(ko) - set(kg)))
print("fixture ids: ground 0, bottoms 1..5, tops 6..10")
