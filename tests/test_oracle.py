"""The CPU oracle against (a) its committed golden vectors on the reference's fixed inputs, (b) the reference's own
in-scene self-checks, (c) physical invariants.  PARITY UNPINNED: see tests/golden/make_golden.py."""
import ctypes as C
import json
import os
import random

from dbox_b200 import _abi as A
from dbox_b200 import scenes

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")))
DT = 1.0 / 60.0


def hexf(x):
    return C.c_float(x).value.hex()


def test_hello_world_golden_and_rest(oracle_api):
    w, body = scenes.hello_world(api=oracle_api)
    for i in range(60):
        w.Step(DT, 6, 2)
        p = body.GetPosition()
        assert [hexf(p.x), hexf(p.y), hexf(body.GetAngle())] == GOLD["hello_world"][i], i
    # rests one polygon skin above the ground's top face at y = 0 (hello_world.d:41,52,60,65)
    assert abs(body.GetPosition().y - 1.015) < 2e-3 and abs(body.GetAngle()) < 1e-4
    assert w.counts().colours == 1   # exactly one TOI event (the landing)


def test_pyramid_golden_stands_and_sleeps(oracle_api):
    w, bodies = scenes.pyramid(api=oracle_api)
    hist, sleep = [], None
    for i in range(400):
        w.Step(DT, 8, 3)
        c = w.counts()
        if i in (0, 10, 20, 50, 100, 399):
            hist.append([i, c.contacts, c.touching, c.awakeBodies, c.islands])
        if sleep is None and c.awakeBodies == 0:
            sleep = i
        if i % 50 == 0:
            assert oracle_api.world_tree_validate(w._w) == 1     # b2dynamictree.d:334-352
    assert hist == GOLD["pyramid"]["history"] and sleep == GOLD["pyramid"]["sleep_step"]
    p = bodies[-1].GetPosition()
    assert [hexf(p.x), hexf(p.y), hexf(bodies[-1].GetAngle())] == GOLD["pyramid"]["top"]
    assert 19.5 < p.y < 19.9                                       # 20 rows high, still standing
    # max penetration <= b2_linearSlop at rest: adjacent rows are 1 + 2*polygonRadius - penetration apart
    rows, k = [], 0
    for r in range(20):
        rows.append([bodies[k + j].GetPosition().y for j in range(20 - r)])
        k += 20 - r
    means = [sum(r) / len(r) for r in rows]
    # resting gap = 2 * b2_polygonRadius minus the penetration.  The reference algorithm itself goes to sleep 0.008-0.0125
    # deep here (b2_linearSlop plus Baumgarte lag at 3 position iterations), so "max penetration <= b2_linearSlop" is not
    # a property of the reference; what holds is that no row sinks through the 0.02 skin.
    assert all(1.0 + 0.02 - 0.015 < b - a < 1.0 + 0.0205 for a, b in zip(means, means[1:])), means


def test_fixed_input_fixtures(oracle_api):
    api = oracle_api
    a, b, m = A.Shape(), A.Shape(), A.Manifold()
    api.shape_set_box(C.byref(a), 0.2, 0.4)
    api.shape_set_box(C.byref(b), 0.5, 0.5)
    assert api.collide(C.byref(a), 0.0, 0.0, 0.0, 0, C.byref(b), 19.345284, 1.5632932, 1.9160721, 0, C.byref(m)) == GOLD["polycollision"]["pointCount"]
    n = api.collide(C.byref(a), 0.0, 0.0, 0.0, 0, C.byref(b), 0.55, 0.3, 1.9160721, 0, C.byref(m))
    g = GOLD["polycollision_touching"]
    assert n == g["pointCount"] and m.type == g["type"]
    assert [hexf(m.localNormal.x), hexf(m.localNormal.y)] == g["localNormal"]
    assert [[hexf(m.points[i].localPoint.x), hexf(m.points[i].localPoint.y), m.points[i].key] for i in range(n)] == g["points"]
    api.shape_set_box(C.byref(a), 10.0, 0.2)
    api.shape_set_box(C.byref(b), 2.0, 0.1)
    pa, pb, it = A.Vec2(), A.Vec2(), C.c_int32()
    d = api.distance(C.byref(a), 0.0, -0.2, 0.0, 0, C.byref(b), 12.017401, 0.13678508, -0.0109265, 0, 1, C.byref(pa), C.byref(pb), C.byref(it))
    assert hexf(d) == GOLD["distancetest"]["distance"] and it.value == GOLD["distancetest"]["iterations"]
    api.shape_set_box(C.byref(a), 25.0, 5.0)
    api.shape_set_box(C.byref(b), 2.5, 2.5)
    sa = (C.c_float * 9)(0, 0, 24.0, -60.0, 24.0, -60.0, 2.95, 2.95, 0)
    sb = (C.c_float * 9)(0, 0, 53.474274, -50.252514, 54.595478, -51.083473, 513.36676, 513.62781, 0)
    t = C.c_float()
    assert api.time_of_impact(C.byref(a), sa, 0, C.byref(b), sb, 0, 1.0, C.byref(t)) == GOLD["timeofimpact"]["state"]
    assert hexf(t.value) == GOLD["timeofimpact"]["t"]


def test_tree_query_equals_brute_force(oracle_api):
    """port of the reference's in-scene self-check (examples/demo/tests/dynamictreetest.d:320-335): the pair set found
    through the dynamic tree equals brute-force b2TestOverlap over all fat AABBs."""
    w, bodies, _ = scenes.pile(api=oracle_api, n=300, columns=15, joints=False)
    oracle_api.world_stage_find_new_contacts(w._w)
    pairs = set(w.read_pairs())
    prox, n = w.read_proxies()
    ps = sorted([prox[i] for i in range(n)], key=lambda p: p.proxyId)
    brute = set()
    for i in range(n):
        for j in range(i + 1, n):
            a, b = ps[i].fat, ps[j].fat
            if b.lo.x - a.hi.x > 0 or b.lo.y - a.hi.y > 0 or a.lo.x - b.hi.x > 0 or a.lo.y - b.hi.y > 0:
                continue
            brute.add((ps[i].fixture, ps[i].child, ps[j].fixture, ps[j].child))
    assert pairs == brute and len(pairs) > 300


def test_momentum_and_energy_sanity(oracle_api):
    """zero gravity, two circles colliding head-on: linear momentum is conserved through the contact solver"""
    from dbox_b200.world import b2BodyDef, b2CircleShape, b2World, b2_dynamicBody
    w = b2World((0.0, 0.0), api=oracle_api)
    bs = []
    for x, vx in ((-2.0, 3.0), (2.0, -1.0)):
        bd = b2BodyDef(); bd.type = b2_dynamicBody; bd.position.Set(x, 0.0); bd.linearVelocity.Set(vx, 0.0)
        b = w.CreateBody(bd)
        s = b2CircleShape(oracle_api); s.m_radius = 0.5
        b.CreateFixture(s, 1.0)
        bs.append(b)
    p0 = sum(b.GetMass() * b.GetLinearVelocity().x for b in bs)
    for _ in range(120):
        w.Step(DT, 8, 3)
    p1 = sum(b.GetMass() * b.GetLinearVelocity().x for b in bs)
    assert abs(p0 - p1) < 1e-4 * abs(p0)
    assert bs[0].GetLinearVelocity().x < 3.0 - 0.5       # they did collide


def test_joints_hold_in_pile(oracle_api):
    """config-4-style pile at small scale: revolute chains keep their anchors together, distance chains their length"""
    w, bodies, nj = scenes.pile(api=oracle_api, n=600, columns=30, seed=3)
    assert nj > 20
    for _ in range(240):
        w.Step(DT, 8, 3)
    st, n = w.read_bodies()
    assert all(abs(st[i].c.x) < 40 and -0.1 < st[i].c.y < 40 for i in range(n) if st[i].type == A.DYNAMIC_BODY)
    c = w.counts()
    assert c.touching > 600


def test_tumbler_motor_turns_container(oracle_api):
    t = scenes.Tumbler(api=oracle_api, count=60)
    for _ in range(200):
        t.Step()
    assert t.container.GetAngle() > 0.4          # 0.05*pi rad/s for 200/60 s
    assert t.world.counts().bodies == 3 + 60


def test_oracle_contact_event_log_balances(oracle_api):
    """the oracle's BeginContact/EndContact call log (b2contact.d:338-346, b2contactmanager.d:60-63): begins minus ends is
    the number of touching contacts, and destroying a resting body ends its contacts in the API phase"""
    from dbox_b200 import scenes
    w, bodies = scenes.pyramid(api=oracle_api, count=6)
    w.EnableContactEvents(1 << 14)
    balance = 0
    for _ in range(120):
        w.Step(1.0 / 60.0, 8, 3)
        for e in w.PollContactEvents():
            assert e[2] == 0 and e[1] in (1, 2)
            balance += 1 if e[0] == 1 else -1
    assert balance == w.counts().touching > 0
    w.DestroyBody(bodies[-1])
    ev = w.PollContactEvents()
    assert ev and all(e[0] == 2 and e[1] == 3 for e in ev)


def test_shift_origin_is_a_translation(oracle_api):
    """b2World.ShiftOrigin (b2world.d:758-780) restated: a pyramid shifted by (64, -32) -- exactly representable, and small
    enough that the coordinates keep their low bits -- evolves like the unshifted one, translated; the tree stays valid"""
    a, ba = scenes.pyramid(api=oracle_api, count=8)
    b, bb = scenes.pyramid(api=oracle_api, count=8)
    for _ in range(30):
        a.Step(DT, 8, 3); b.Step(DT, 8, 3)
    b.ShiftOrigin((64.0, -32.0))
    assert oracle_api.world_tree_validate(b._w) == 1
    for _ in range(60):
        a.Step(DT, 8, 3); b.Step(DT, 8, 3)
    ca, cb = a.counts(), b.counts()
    assert (ca.contacts, ca.touching) == (cb.contacts, cb.touching)
    for x, y in zip(ba, bb):
        px, py = x.GetPosition(), y.GetPosition()
        assert abs(px.x - 64.0 - py.x) < 2e-3 and abs(px.y + 32.0 - py.y) < 2e-3


def test_post_solve_log_carries_the_weight(oracle_api):
    """b2Island.Report (b2island.d:438-462) restated: at rest, the normal impulses reported for the contacts with the ground
    add up to the weight of the pyramid times the time step"""
    w, bodies = scenes.pyramid(api=oracle_api, count=6)
    w.SetAllowSleeping(False)
    w.EnablePostSolve(1 << 12)
    for _ in range(240):
        w.Step(DT, 8, 3)
    recs = w.ReadPostSolve()
    assert recs and all(r[0] == 1 for r in recs)
    ground = [r for r in recs if 0 in (w._fixtures[r[1]].body.id, w._fixtures[r[3]].body.id) or 1 in (w._fixtures[r[1]].body.id, w._fixtures[r[3]].body.id)]
    total = sum(sum(r[6][:r[5]]) for r in ground)
    weight = 21 * 5.0 * 1.0 * 10.0 * DT            # 21 boxes, density 5, area 1, g = 10
    assert abs(total - weight) < 0.02 * weight, (total, weight)


def test_test_point_and_world_manifold_fixed_inputs(oracle_api):
    """b2Shape.TestPoint (b2polygonshape.d:265-279, b2circleshape.d:60-65) and b2WorldManifold.Initialize (b2collision.d:123-191)
    on hand-checkable inputs: a unit box on the ground"""
    w, body = scenes.hello_world(api=oracle_api)
    for _ in range(90):
        w.Step(DT, 6, 2)
    f = body.fixtures[0]
    y = body.GetPosition().y
    assert f.TestPoint((0.0, y)) and f.TestPoint((0.99, y + 0.99)) and not f.TestPoint((1.01, y)) and not f.TestPoint((0.0, y + 1.01))
    wm = w.GetWorldManifolds()
    assert len(wm) == 1 and wm[0][0] == 2
    n, pts, sep = wm[0][1], wm[0][2], wm[0][3]
    assert abs(abs(n[1]) - 1.0) < 1e-6 and abs(n[0]) < 1e-6            # vertical normal, A -> B
    assert sorted(round(p[0], 3) for p in pts) == [-1.0, 1.0]            # the box's two bottom corners
    assert all(-0.011 < s < 0.0 for s in sep)                            # resting inside the 2 * b2_polygonRadius skin


def test_world_query_goldens(oracle_api):
    """ray casts (closest hit and every hit), AABB queries, TestPoint and world manifolds of the oracle on a fixed scene, before the
    first step and after 60: frozen in tests/golden/oracle_golden.json so that the checker itself cannot drift"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    got = mg.queries_golden(oracle_api)
    assert json.loads(json.dumps(got)) == GOLD["queries"]
    assert "1" in GOLD["queries"]["initial"]["inside"] and "0" in GOLD["queries"]["initial"]["inside"]
    assert len(GOLD["queries"]["after60"]["world_manifolds"]) >= 9


def test_state_import_round_trip_and_ordered_step(oracle_api):
    """orc_world_write_* (the mirror of dbx_world_write_*) and the solve-order hook: a world's state copied into a freshly
    built twin reads back record for record, the tree stays valid, and the twin -- told to walk its islands in the order the
    original did -- steps to the same bits, TOI sub-steps included (jointed pile, still settling)."""
    import ctypes as C
    from dbox_b200 import _abi as A
    from dbox_b200 import scenes, state
    api = oracle_api
    make = lambda: scenes.pile(api=api, n=600, columns=30)[0]
    wa, wb = make(), make()
    wa.StepN(1.0 / 60.0, 8, 3, 120)
    moved = state.transplant(wa, wb)
    assert moved["bodies"] == 601 and moved["joints"] > 50 and moved["contacts"] > 1000
    for rd, typ in (("read_bodies", A.BodyState), ("read_proxies", A.ProxyRec), ("read_joints", A.JointState)):
        a, na = getattr(wa, rd)(); b, nb = getattr(wb, rd)()
        assert na == nb and bytes(a)[:na * C.sizeof(typ)] == bytes(b)[:nb * C.sizeof(typ)], rd
    ca, na = wa.read_contacts(); cb, nb = wb.read_contacts()
    assert sorted(bytes(ca[i]) for i in range(na)) == sorted(bytes(cb[i]) for i in range(nb))
    assert api.world_tree_validate(wb._w) == 1
    wa.Step(1.0 / 60.0, 8, 3)
    n = api.world_read_solve_order(wa._w, None, 0)
    keys = (C.c_int32 * (4 * n))(); api.world_read_solve_order(wa._w, keys, n)
    nj = wa.counts().joints
    rank = (C.c_int32 * n)(*range(nj, nj + n))      # one rank space: the joints (ranks 0 .. nj-1) before every contact
    m = api.world_read_joint_solve_order(wa._w, None, 0)
    jo = (C.c_int32 * m)(); api.world_read_joint_solve_order(wa._w, jo, m)
    jr = (C.c_int32 * nj)(*([0x7fffffff] * nj))
    for k in range(m):
        jr[jo[k]] = k
    assert api.world_debug_set_solve_order(wb._w, keys, rank, n, jr, nj, 2) == n    # 2: position passes contacts, then joints (b2island.d:206-216)
    wb.Step(1.0 / 60.0, 8, 3)
    a, na = wa.read_bodies(); b, nb = wb.read_bodies()
    assert bytes(a)[:na * C.sizeof(A.BodyState)] == bytes(b)[:nb * C.sizeof(A.BodyState)]
    # the hook is one-shot: the next step is back on DFS order (and a different order gives a different unconverged step)
    wa.Step(1.0 / 60.0, 8, 3); wb.Step(1.0 / 60.0, 8, 3)
    a, na = wa.read_bodies(); b, nb = wb.read_bodies()
    assert bytes(a)[:na * C.sizeof(A.BodyState)] != bytes(b)[:nb * C.sizeof(A.BodyState)]


def test_pile_generator_draws_from_std_mt19937():
    """SURVEY.md 8(d) names std::mt19937(12345) for the pile's jitter.  scenes.Mt19937 against the generator's known answers (the
    C++ standard's check value: the 10,000th output of a default-constructed mt19937 is 4123659995) and numpy's legacy stream
    (init_genrand seeding, the same as std::mt19937(seed)); the pile built from it is the same scene every time."""
    import numpy as np
    r = scenes.Mt19937(5489)
    first = r.draw()
    for _ in range(9998):
        r.draw()
    assert first == 3499211612 and r.draw() == 4123659995
    r = scenes.Mt19937(12345)
    ref = np.random.RandomState(12345).randint(0, 2 ** 32, size=2000, dtype=np.uint64)
    assert [r.draw() for _ in range(2000)] == [int(x) for x in ref]
    r = scenes.Mt19937(12345)
    us = [r.uniform(-0.01, 0.01) for _ in range(5000)]
    assert min(us) >= -0.01 and max(us) < 0.01 and abs(sum(us) / len(us)) < 5e-4
    # (x, y, shape) per body, in body order: body 0 of the pile sits at the first two draws
    r = scenes.Mt19937(12345)
    jx, jy = r.uniform(-0.01, 0.01), r.uniform(-0.01, 0.01)
    from oracle import orc
    w, bodies, _ = scenes.pile(api=orc.api(), n=40, columns=10, joints=False)
    p = bodies[0].GetPosition()
    assert p.x == scenes.f32(scenes.f32(-1.05 * 10 * 0.5 + 0.525) + jx) and p.y == scenes.f32(0.55 + jy)
    w.close()


def test_one_way_platform_with_continuous_physics_on_the_oracle(oracle_api):
    """The PreSolve decision of a step stands through the TOI loop's re-evaluations, and the listener can be asked ahead of a first
    touch that happens there (include/dbox_b200.h, "PreSolve"): the oracle's side of tests/test_gpu_features.py::
    test_pre_solve_split_step[True] -- a box shot upwards passes a one-way platform with continuous physics on, and is stopped
    by it when nobody is asked ahead."""
    from dbox_b200.world import b2BodyDef, b2EdgeShape, b2PolygonShape, b2World, b2_dynamicBody

    def run(lookahead):
        w = b2World((0.0, -10.0), api=oracle_api)
        g = w.CreateBody(b2BodyDef()); e = b2EdgeShape(oracle_api); e.Set((-40.0, 0.0), (40.0, 0.0)); g.CreateFixture(e, 0.0)
        plat = w.CreateBody(b2BodyDef()); ps = b2PolygonShape(oracle_api); ps.SetAsBox(3.0, 0.25, (0.0, 6.0), 0.0)
        pf = plat.CreateFixture(ps, 0.0).id
        bd = b2BodyDef(); bd.type = b2_dynamicBody; bd.position.Set(0.0, 3.0)
        up = w.CreateBody(bd); s = b2PolygonShape(oracle_api); s.SetAsBox(0.5, 0.5); up.CreateFixture(s, 1.0)
        up.SetLinearVelocity((0.0, 14.0))

        def pre_solve(c):
            if pf in (c.fixtureA, c.fixtureB):
                return {"enabled": up.GetLinearVelocity().y <= 0.0}
            return None
        for _ in range(150):
            w.StepWithPreSolve(DT, 8, 3, pre_solve, toi_lookahead=lookahead)
        y = up.GetPosition().y
        w.close()
        return y
    assert 6.7 < run(True) < 6.85          # through the platform on the way up, resting on it afterwards
    assert run(False) < 1.0                 # first touch inside the TOI loop, nobody asked: the platform is solid from below


def test_oracle_matches_reference_golden():
    """tests/golden/reference_golden.json = output of oracle/dref/harness.d linked against the UNMODIFIED dbox (oracle/dref/build.sh,
    needs a D compiler).  When the file is there, the oracle's own goldens must equal it bit for bit; while it is not, the oracle is
    PARITY UNPINNED and this test says so instead of passing silently."""
    import json
    import os
    import struct
    import pytest
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    ref_path = os.path.join(here, "reference_golden.json")
    if not os.path.exists(ref_path):
        pytest.skip("PARITY UNPINNED: no reference_golden.json (no D compiler in the build image or on the GPU box; recipe: oracle/dref/build.sh)")
    ref = json.load(open(ref_path))
    orc_g = json.load(open(os.path.join(here, "oracle_golden.json")))

    def norm(x):
        """oracle goldens carry floats as float.hex() strings, the D harness as 8 hex digits of the IEEE bits"""
        if isinstance(x, str):
            if x.startswith(("0x", "-0x")):
                return struct.unpack("<I", struct.pack("<f", float.fromhex(x)))[0]
            return int(x, 16)
        if isinstance(x, list):
            return [norm(v) for v in x]
        if isinstance(x, dict):
            return {k: norm(v) for k, v in x.items()}
        return x
    for key in ("constants", "polycollision", "polycollision_touching", "distancetest", "timeofimpact", "hello_world"):
        want, got = norm(ref[key]), norm(orc_g[key])
        if isinstance(got, dict):
            want = {k: want[k] for k in got}        # the harness prints whole manifolds; compare what the oracle golden holds
        assert want == got, key
    rp, op = norm(ref["pyramid"]), norm(orc_g["pyramid"])
    assert rp["sleep_step"] == op["sleep_step"] and rp["top"] == op["top"]
    assert rp["history"] == [h[:4] for h in op["history"]]      # the reference has no island counter to read
