"""Regenerates tests/golden/oracle_golden.json from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

PARITY UNPINNED: the reference ships no expected outputs and cannot be compiled in this image (no D toolchain), so these
vectors pin the ORACLE (the C++ restatement of dbox) against regressions — on the reference's own fixed inputs
(examples/demo/tests/polycollision.d:40-57, distancetest.d:43-51, timeofimpact.d:40-74, hello_world.d:31-103,
pyramid.d:39-78) — not the oracle against the reference.  Toolchain used: g++ -O2 -ffp-contract=off -fno-fast-math, glibc libm.
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dbox_b200 import _abi as A  # noqa: E402
from dbox_b200 import scenes  # noqa: E402
from oracle import orc  # noqa: E402


def hexf(x):
    return C.c_float(x).value.hex()


def query_scene(api):
    """a small fixed scene for the world-query goldens: a chain ground, a vertical edge, nine bodies of three shape kinds"""
    from dbox_b200.world import b2BodyDef, b2ChainShape, b2CircleShape, b2EdgeShape, b2PolygonShape, b2World, b2_dynamicBody
    w = b2World((0.0, -10.0), api=api)
    g = w.CreateBody(b2BodyDef())
    ch = b2ChainShape(api); ch.CreateChain([(-12.0, 0.0), (-4.0, 0.5), (4.0, 0.0), (12.0, 1.0)])
    g.CreateFixture(ch, 0.0)
    e = b2EdgeShape(api); e.Set((-12.0, 0.0), (-12.0, 8.0)); g.CreateFixture(e, 0.0)
    bodies = []
    for k in range(9):
        bd = b2BodyDef(); bd.type = b2_dynamicBody
        bd.position.Set(-8.0 + 2.0 * k, 1.5 + 0.4 * (k % 3)); bd.angle = 0.3 * k
        b = w.CreateBody(bd)
        if k % 3 == 0:
            s = b2CircleShape(api); s.m_radius = 0.5
        elif k % 3 == 1:
            s = b2PolygonShape(api); s.SetAsBox(0.6, 0.4)
        else:
            s = b2PolygonShape(api); s.Set([(-0.7, -0.5), (0.8, -0.4), (0.5, 0.6), (-0.3, 0.9), (-0.8, 0.2)])
        b.CreateFixture(s, 1.0)
        bodies.append(b)
    return w, bodies


QUERY_RAYS = [((-13.0, 2.0), (13.0, 2.0)), ((-13.0, 1.2), (13.0, 2.6)), ((0.0, 9.0), (0.0, -1.0)), ((-8.0, 9.0), (-7.6, -1.0)),
              ((5.0, 5.0), (-5.0, 0.2)), ((-13.0, 6.0), (-11.0, 6.0)), ((2.0, 2.0), (2.0, 2.0))]


def queries_golden(api):
    """b2World.RayCast (closest / all), QueryAABB, b2Fixture.TestPoint and b2Contact.GetWorldManifold of the oracle on the fixed
    scene above, before the first step and after 60 steps"""
    w, bodies = query_scene(api)
    out = {}
    for tag, steps in (("initial", 0), ("after60", 60)):
        for _ in range(steps):
            w.Step(1.0 / 60.0, 8, 3)
        closest = [[h[0], h[1], hexf(h[2]), hexf(h[4][0]), hexf(h[4][1])] for h in w.RayCastClosest(QUERY_RAYS)]
        allhits = [[[h[0], h[1], hexf(h[2])] for h in hits] for hits in w.RayCastAll(QUERY_RAYS, cap=32)]
        boxes = w.QueryAABB([((-9.0, 0.0), (-3.0, 3.0)), ((3.0, 0.0), (9.0, 4.0))], cap=32)
        pts = [(f, (p.x + dx, p.y + dy)) for b in bodies for f in b.fixtures for p in [b.GetPosition()] for dx in (-0.55, 0.0, 0.45) for dy in (-0.45, 0.0, 0.55)]
        inside = "".join("1" if v else "0" for v in w.TestPoints(pts))
        wm = [[m[0], hexf(m[1][0]), hexf(m[1][1])] + [hexf(m[2][k][c]) for k in range(m[0]) for c in (0, 1)] + [hexf(m[3][k]) for k in range(m[0])]
              for m in w.GetWorldManifolds() if m[0] > 0]
        out[tag] = {"closest": closest, "all": allhits, "boxes": [[list(x) for x in b] for b in boxes], "inside": inside, "world_manifolds": wm}
    return out


def main():
    api = orc.api()
    out = {}
    # polycollision.d:40-57
    a, b = A.Shape(), A.Shape()
    api.shape_set_box(C.byref(a), 0.2, 0.4)
    api.shape_set_box(C.byref(b), 0.5, 0.5)
    m = A.Manifold()
    n = api.collide(C.byref(a), 0.0, 0.0, 0.0, 0, C.byref(b), 19.345284, 1.5632932, 1.9160721, 0, C.byref(m))
    out["polycollision"] = {"pointCount": n}
    # the demo's interesting case: the same boxes brought into contact
    n = api.collide(C.byref(a), 0.0, 0.0, 0.0, 0, C.byref(b), 0.55, 0.3, 1.9160721, 0, C.byref(m))
    out["polycollision_touching"] = {"pointCount": n, "type": m.type, "localNormal": [hexf(m.localNormal.x), hexf(m.localNormal.y)],
                                     "localPoint": [hexf(m.localPoint.x), hexf(m.localPoint.y)],
                                     "points": [[hexf(m.points[i].localPoint.x), hexf(m.points[i].localPoint.y), m.points[i].key] for i in range(n)]}
    # distancetest.d:43-51
    api.shape_set_box(C.byref(a), 10.0, 0.2)
    api.shape_set_box(C.byref(b), 2.0, 0.1)
    pa, pb, it = A.Vec2(), A.Vec2(), C.c_int32()
    d = api.distance(C.byref(a), 0.0, -0.2, 0.0, 0, C.byref(b), 12.017401, 0.13678508, -0.0109265, 0, 1, C.byref(pa), C.byref(pb), C.byref(it))
    out["distancetest"] = {"distance": hexf(d), "iterations": it.value, "pointA": [hexf(pa.x), hexf(pa.y)], "pointB": [hexf(pb.x), hexf(pb.y)]}
    # timeofimpact.d:40-74
    api.shape_set_box(C.byref(a), 25.0, 5.0)
    api.shape_set_box(C.byref(b), 2.5, 2.5)
    sa = (C.c_float * 9)(0, 0, 24.0, -60.0, 24.0, -60.0, 2.95, 2.95, 0)
    sb = (C.c_float * 9)(0, 0, 53.474274, -50.252514, 54.595478, -51.083473, 513.36676, 513.62781, 0)
    t = C.c_float()
    st = api.time_of_impact(C.byref(a), sa, 0, C.byref(b), sb, 0, 1.0, C.byref(t))
    out["timeofimpact"] = {"state": st, "t": hexf(t.value)}
    # hello_world.d: x y angle for 60 steps
    w, body = scenes.hello_world(api=api)
    traj = []
    for _ in range(60):
        w.Step(1.0 / 60.0, 6, 2)
        p = body.GetPosition()
        traj.append([hexf(p.x), hexf(p.y), hexf(body.GetAngle())])
    out["hello_world"] = traj
    # pyramid.d: contact counts + sleep step + final top-box pose over 400 steps
    w, bodies = scenes.pyramid(api=api)
    sleep = None
    hist = []
    for i in range(400):
        w.Step(1.0 / 60.0, 8, 3)
        c = w.counts()
        if i in (0, 10, 20, 50, 100, 399):
            hist.append([i, c.contacts, c.touching, c.awakeBodies, c.islands])
        if sleep is None and c.awakeBodies == 0:
            sleep = i
    p = bodies[-1].GetPosition()
    out["pyramid"] = {"history": hist, "sleep_step": sleep, "top": [hexf(p.x), hexf(p.y), hexf(bodies[-1].GetAngle())]}
    out["queries"] = queries_golden(api)
    # constants whose D compile-time folding may differ in the last ulp
    out["constants"] = {"angularSlop": hexf(2.0 / 180.0 * 3.14159265359), "maxAngularCorrection": hexf(8.0 / 180.0 * 3.14159265359)}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote oracle_golden.json")


if __name__ == "__main__":
    main()
