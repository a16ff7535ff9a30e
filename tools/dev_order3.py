"""dev: which part of a vi=0, pi=0 step differs between the device and the oracle after a device -> oracle transplant"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from dbox_b200 import lib, scenes, state, _abi as A
from oracle import orc
from tests.parity import contact_key
from tests.test_gpu_full_size import _recs, _rel
ga, oa = lib.api(), orc.api()
DT = 1 / 60
n, cols, settle = 3000, 100, 300
wg, gb, nj = scenes.pile(api=ga, n=n, columns=cols, joints=False, circles=False)
wo, ob, _ = scenes.pile(api=oa, n=n, columns=cols, joints=False, circles=False)
for w in (wg, wo):
    w.SetAllowSleeping(False); w.SetContinuousPhysics(False)
wg.StepN(DT, 8, 3, settle)
snap = state.capture(wg)
state.apply(wo, snap)
fix_body = {}
for b in gb:
    for f in b.fixtures:
        fix_body[f.id if hasattr(f, "id") else f] = b
pre_g = {contact_key(r): (r.flags, r.manifold.pointCount, r.manifold.type, [(r.manifold.points[j].key, r.manifold.points[j].normalImpulse, r.manifold.points[j].tangentImpulse) for j in range(2)]) for r in [snap["contacts"][i] for i in range(snap["nc"])]}
co, no = wo.read_contacts()
pre_o = {contact_key(r): (r.flags, r.manifold.pointCount, r.manifold.type, [(r.manifold.points[j].key, r.manifold.points[j].normalImpulse, r.manifold.points[j].tangentImpulse) for j in range(2)]) for r in [co[i] for i in range(no)]}
print("pre-step records equal:", pre_g == pre_o)
wg.Step(DT, 0, 0); wo.Step(DT, 0, 0)
bg, nb = wg.read_bodies(); bo, _ = wo.read_bodies()
G, O = _recs(bg, nb, A.BodyState), _recs(bo, nb, A.BodyState)
ev = np.maximum(np.maximum(_rel(G["v"]["x"], O["v"]["x"], 1.0), _rel(G["v"]["y"], O["v"]["y"], 1.0)), _rel(G["w"], O["w"], 1.0))
bad = np.nonzero(ev > 1e-4)[0]
print("bad bodies", bad[:40])
cg, ng = wg.read_contacts(); co, no = wo.read_contacts()
dg = {contact_key(cg[i]): cg[i] for i in range(ng)}; do = {contact_key(co[i]): co[i] for i in range(no)}
P0 = _recs(snap["bodies"], snap["nb"], A.BodyState)
shown = 0
for k in sorted(dg):
    g, o = dg[k], do.get(k)
    if o is None:
        print("missing in oracle", k); continue
    gi = [(g.manifold.points[j].key, g.manifold.points[j].normalImpulse, g.manifold.points[j].tangentImpulse) for j in range(g.manifold.pointCount)]
    oi = [(o.manifold.points[j].key, o.manifold.points[j].normalImpulse, o.manifold.points[j].tangentImpulse) for j in range(o.manifold.pointCount)]
    involved = (k[0] - 2 in bad[:3]) or (k[2] - 2 in bad[:3])
    if gi != oi or (g.flags & 0x3E) != (o.flags & 0x3E) or involved:
        if shown < 40:
            print("contact", k, "flags g/o %x %x" % (g.flags, o.flags), "type", g.manifold.type, o.manifold.type, "ln", (g.manifold.localNormal.x, g.manifold.localNormal.y), (o.manifold.localNormal.x, o.manifold.localNormal.y), "fric", g.friction, o.friction, "\n   pre ", pre_g.get(k), "\n   g   ", gi, "\n   o   ", oi)
        shown += 1
print("contacts whose post-Collide impulses / flags differ:", shown, "of", ng)
for b in bad[:3]:
    print("body", b, "pre v", P0["v"][b], P0["w"][b], "a", P0["a"][b], "qs qc", P0["qs"][b], P0["qc"][b], "sin cos(a)", np.sin(np.float32(P0["a"][b])), np.cos(np.float32(P0["a"][b])),
          "\n   post g v", G["v"][b], G["w"][b], " o v", O["v"][b], O["w"][b])
