"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): pair set + contact feature keys bit-exact; manifold points/normals <= 1e-5 relative;
single-step velocities <= 1e-4 relative from identical pre-step state; long-horizon scenes by invariants."""
import ctypes as C
import math
import random

import pytest

from dbox_b200 import _abi as A
from dbox_b200 import scenes
from tests import parity as P

pytestmark = pytest.mark.gpu

DT = 1.0 / 60.0


def _maps(world):
    fixture_body = {fid: f.body.id for fid, f in world._fixtures.items()}
    bodies, n = world.read_bodies()
    body_dynamic = {i: bodies[i].type == A.DYNAMIC_BODY for i in range(n)}
    return fixture_body, body_dynamic


def test_hello_world_trajectory(gpu_api, oracle_api):
    """config 1: examples/hello_world/hello_world.d — 60 steps at 6/2 iterations, trajectory vs oracle (one TOI event)."""
    wg, bg = scenes.hello_world(api=gpu_api)
    wo, bo = scenes.hello_world(api=oracle_api)
    worst = 0.0
    for _ in range(60):
        wg.Step(DT, 6, 2)
        wo.Step(DT, 6, 2)
        pg, po = bg.GetPosition(), bo.GetPosition()
        worst = max(worst, abs(pg.x - po.x), abs(pg.y - po.y), abs(bg.GetAngle() - bo.GetAngle()))
    assert worst < 2e-5, worst
    assert abs(bg.GetPosition().y - 1.015) < 2e-3          # rests one skin above the ground top (hello_world.d:41,52,60,65)
    assert wo.counts().colours >= 1                         # the oracle saw a TOI event, so the GPU path had to handle one


@pytest.mark.parametrize("presteps", [15, 40, 90, 150])
def test_pyramid_single_step_sequential_order(gpu_api, oracle_api, presteps):
    """Identical pre-step state, the reference's own Gauss-Seidel order injected as level schedule: the coloured solver
    must then reproduce the sequential solve up to libm-vs-CUDA sin/cos rounding."""
    wo, _ = scenes.pyramid(api=oracle_api)
    wg, _ = scenes.pyramid(api=gpu_api)
    for _ in range(presteps):
        wo.Step(DT, 8, 3)
    P.transplant(wo, wg)
    wo.Step(DT, 8, 3)
    fixture_body, body_dynamic = _maps(wg)
    levels, m, nlev = P.sequential_levels(oracle_api, wo, wg, fixture_body, body_dynamic)
    assert gpu_api.world_debug_set_contact_levels(wg._w, levels, m) == 0, gpu_api.last_error()
    wg.Step(DT, 8, 3)
    so, n = wo.read_bodies()
    sg, _ = wg.read_bodies()
    ep, ev = P.body_state_errors(so, sg, n)
    assert ep < 1e-5 and ev < 1e-4, (presteps, nlev, ep, ev)
    # contact set, touching flags and feature keys after the step: bit-exact
    co, _, no = P.contacts_by_key(wo)
    cg, _, ng = P.contacts_by_key(wg)
    assert set(co) == set(cg)
    for k, ro in co.items():
        rg = cg[k]
        assert (ro.flags & A.CONTACT_TOUCHING) == (rg.flags & A.CONTACT_TOUCHING), k
        if ro.flags & A.CONTACT_TOUCHING:
            assert ro.manifold.pointCount == rg.manifold.pointCount and ro.manifold.type == rg.manifold.type, k
            for j in range(ro.manifold.pointCount):
                assert ro.manifold.points[j].key == rg.manifold.points[j].key, k


@pytest.mark.parametrize("presteps", [30, 120])
def test_pyramid_single_step_colour_order(gpu_api, oracle_api, presteps):
    """Same pre-step state, the GPU's own colouring: Gauss-Seidel order differs from the reference's, so only the
    converged (warm-started) regime is comparable: velocities within 1e-4 of the speed scale, positions 1e-5."""
    wo, _ = scenes.pyramid(api=oracle_api)
    wg, _ = scenes.pyramid(api=gpu_api)
    for _ in range(presteps):
        wo.Step(DT, 8, 3)
    P.transplant(wo, wg)
    wo.Step(DT, 8, 3)
    wg.Step(DT, 8, 3)
    so, n = wo.read_bodies()
    sg, _ = wg.read_bodies()
    worst_v = max(max(abs(so[i].v.x - sg[i].v.x), abs(so[i].v.y - sg[i].v.y), abs(so[i].w - sg[i].w)) for i in range(n))
    worst_p = max(max(abs(so[i].c.x - sg[i].c.x), abs(so[i].c.y - sg[i].c.y)) for i in range(n))
    # Gauss-Seidel in colour order is a different (equally valid) sweep of an unconverged system: bounded difference only.
    # The tight 1e-5 / 1e-4 bar is test_pyramid_single_step_sequential_order, which injects the reference's order.
    assert worst_p < 2e-2 and worst_v < 0.5, (worst_p, worst_v)    # step 30 is mid-impact (bodies meeting at ~5 m/s)
    assert wg.counts().colours <= 12
    assert gpu_api.world_debug_colour_conflicts(wg._w) == 0


def test_collide_stage_bit_exact(gpu_api, oracle_api):
    """b2ContactManager.Collide from identical state: manifolds are a pure function of transforms -> bit-exact."""
    wo, _ = scenes.pyramid(api=oracle_api)
    wg, _ = scenes.pyramid(api=gpu_api)
    for _ in range(60):
        wo.Step(DT, 8, 3)
    P.transplant(wo, wg)
    oracle_api.world_stage_collide(wo._w)
    assert gpu_api.world_stage_collide(wg._w) == 0
    co, _, _ = P.contacts_by_key(wo)
    cg, _, _ = P.contacts_by_key(wg)
    assert set(co) == set(cg)
    touching = 0
    for k, ro in co.items():
        rg = cg[k]
        assert (ro.flags & 0x6) == (rg.flags & 0x6), k
        mo, mg = ro.manifold, rg.manifold
        assert mo.pointCount == mg.pointCount
        if mo.pointCount:
            touching += 1
            assert mo.type == mg.type
            assert (mo.localNormal.x, mo.localNormal.y, mo.localPoint.x, mo.localPoint.y) == (mg.localNormal.x, mg.localNormal.y, mg.localPoint.x, mg.localPoint.y)
            for j in range(mo.pointCount):
                assert mo.points[j].key == mg.points[j].key
                assert (mo.points[j].localPoint.x, mo.points[j].localPoint.y) == (mg.points[j].localPoint.x, mg.points[j].localPoint.y)
                assert mo.points[j].normalImpulse == mg.points[j].normalImpulse and mo.points[j].tangentImpulse == mg.points[j].tangentImpulse
    assert touching > 100


def test_broadphase_pairs_bit_exact(gpu_api, oracle_api):
    """b2BroadPhase.UpdatePairs: the LBVH must report exactly the pair set of the dynamic tree for the same fat AABBs
    and move buffer, in the reference's (proxyIdA, proxyIdB) order, and AddPair must create the same contacts."""
    for build, kw in ((scenes.pyramid, {}), (scenes.pile, {"n": 400, "columns": 20})):
        r = build(api=oracle_api, **kw)
        wo = r[0]
        r = build(api=gpu_api, **kw)
        wg = r[0]
        # creation-time move buffer: every proxy queries
        assert wo.read_moves() == wg.read_moves()
        oracle_api.world_stage_find_new_contacts(wo._w)
        assert gpu_api.world_stage_find_new_contacts(wg._w) == 0, gpu_api.last_error()
        po, pg = wo.read_pairs(), wg.read_pairs()
        assert po == pg and len(po) > 50
        co, _, _ = P.contacts_by_key(wo)
        cg, _, _ = P.contacts_by_key(wg)
        assert set(co) == set(cg)       # same contacts, same fixture A/B order (proxy-id order + type-registry swap)
        # and again mid-simulation, with a partial move buffer and persistent fat boxes
        for _ in range(25):
            wo.Step(DT, 8, 3)
        P.transplant(wo, wg)
        wo.Step(DT, 8, 3)
        wg.Step(DT, 8, 3)
        fo = {(p.fixture, p.child): p for p in wo.read_proxies()[0][:wo.read_proxies()[1]]}
        fg = {(p.fixture, p.child): p for p in wg.read_proxies()[0][:wg.read_proxies()[1]]}
        assert set(fo) == set(fg)
        assert [p.proxyId for p in fo.values()] == [fg[k].proxyId for k in fo]
        co, _, _ = P.contacts_by_key(wo)
        cg, _, _ = P.contacts_by_key(wg)
        assert set(co) == set(cg)


def _random_shape(api, rng, kind):
    s = A.Shape()
    if kind == "circle":
        api.shape_set_circle(C.byref(s), rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(0.1, 1.0))
    elif kind == "box":
        api.shape_set_box(C.byref(s), rng.uniform(0.1, 1.5), rng.uniform(0.1, 1.5))
    elif kind == "poly":
        n = rng.randint(3, 8)
        pts = (A.Vec2 * n)(*[A.Vec2(rng.uniform(-1, 1), rng.uniform(-1, 1)) for _ in range(n)])
        api.shape_set_polygon(C.byref(s), pts, n)
    else:  # edge, sometimes with ghost vertices (chain child)
        api.shape_set_edge(C.byref(s), A.Vec2(-rng.uniform(0.5, 2), rng.uniform(-0.2, 0.2)), A.Vec2(rng.uniform(0.5, 2), rng.uniform(-0.2, 0.2)))
        if rng.random() < 0.5:
            s.v0 = A.Vec2(s.v1.x - rng.uniform(0.5, 2), s.v1.y + rng.uniform(-1, 1)); s.hasV0 = 1
        if rng.random() < 0.5:
            s.v3 = A.Vec2(s.v2.x + rng.uniform(0.5, 2), s.v2.y + rng.uniform(-1, 1)); s.hasV3 = 1
    return s


@pytest.mark.parametrize("kinds", [("box", "box"), ("poly", "poly"), ("circle", "circle"), ("poly", "circle"), ("edge", "circle"), ("edge", "poly")])
def test_narrowphase_random_bit_exact(gpu_api, oracle_api, kinds):
    """K5a-e: 4000 random shape pairs per contact class near touching; manifold type, counts, feature keys and local
    geometry must equal the oracle's bit for bit (rotations passed as (sin, cos) so libm does not enter)."""
    rng = random.Random(hash(kinds) & 0xFFFF)
    n = 4000
    sa = (A.Shape * n)(); sb = (A.Shape * n)()
    xa = (C.c_float * (4 * n))(); xb = (C.c_float * (4 * n))()
    for i in range(n):
        sa[i] = _random_shape(oracle_api, rng, kinds[0])
        sb[i] = _random_shape(oracle_api, rng, kinds[1])
        aa, ab = rng.uniform(-3.2, 3.2), rng.uniform(-3.2, 3.2)
        d = rng.uniform(0.0, 2.5); th = rng.uniform(0, 2 * math.pi)
        xa[4 * i:4 * i + 4] = [rng.uniform(-5, 5), rng.uniform(-5, 5), P.f32(math.sin(aa)), P.f32(math.cos(aa))]
        xb[4 * i:4 * i + 4] = [xa[4 * i] + d * math.cos(th), xa[4 * i + 1] + d * math.sin(th), P.f32(math.sin(ab)), P.f32(math.cos(ab))]
    out = (A.Manifold * n)()
    assert gpu_api.debug_collide(0, n, sa, xa, sb, xb, out) == n, gpu_api.last_error()
    hits = 0
    for i in range(n):
        mo = A.Manifold()
        oracle_api.collide_xf(C.byref(sa[i]), C.cast(C.byref(xa, 16 * i), C.POINTER(C.c_float)), 0, C.byref(sb[i]),
                              C.cast(C.byref(xb, 16 * i), C.POINTER(C.c_float)), 0, C.byref(mo))
        mg = out[i]
        assert mo.pointCount == mg.pointCount, (i, kinds)
        if mo.pointCount:
            hits += 1
            assert mo.type == mg.type
            assert (mo.localNormal.x, mo.localNormal.y, mo.localPoint.x, mo.localPoint.y) == (mg.localNormal.x, mg.localNormal.y, mg.localPoint.x, mg.localPoint.y), (i, kinds)
            for j in range(mo.pointCount):
                assert mo.points[j].key == mg.points[j].key, (i, kinds)
                assert (mo.points[j].localPoint.x, mo.points[j].localPoint.y) == (mg.points[j].localPoint.x, mg.points[j].localPoint.y), (i, kinds)
    assert hits > n // 10


def test_pyramid_long_horizon_invariants(gpu_api, oracle_api):
    """config 2: 1000 steps at 8/3.  Judged by invariants: the pyramid stands, rests within b2_linearSlop of the oracle's
    heights, everything falls asleep, and it does so within a few steps of the oracle."""
    wg, bg = scenes.pyramid(api=gpu_api)
    wo, bo = scenes.pyramid(api=oracle_api)
    sleep_g = sleep_o = None
    for i in range(1000):
        wg.Step(DT, 8, 3)
        wo.Step(DT, 8, 3)
        if sleep_g is None and i % 4 == 3 and wg.counts().awakeBodies == 0:
            sleep_g = i
        if sleep_o is None and wo.counts().awakeBodies == 0:
            sleep_o = i
    sg, n = wg.read_bodies()
    so, _ = wo.read_bodies()
    cg, co = wg.counts(), wo.counts()
    assert cg.awakeBodies == 0 and co.awakeBodies == 0
    assert sleep_g is not None and abs(sleep_g - sleep_o) < 60, (sleep_g, sleep_o)
    assert cg.touching == co.touching == 400 and cg.contacts == co.contacts == 590
    top_g, top_o = bg[-1].GetPosition(), bo[-1].GetPosition()
    assert abs(top_g.y - top_o.y) < 0.005 and abs(top_g.x - top_o.x) < 0.05, (top_g, top_o)
    # every box within 2 * b2_linearSlop of the oracle's resting height (the oracle itself sleeps 8-12 mm deep) and upright
    for i in range(n):
        if so[i].type != A.DYNAMIC_BODY:
            continue
        assert abs(sg[i].c.y - so[i].c.y) < 0.01, (i, sg[i].c.y, so[i].c.y)   # 2 * b2_linearSlop: stacked resting depths add up
        assert abs(sg[i].a) < 0.05 and abs(sg[i].a - so[i].a) < 0.05, (i, sg[i].a, so[i].a)


def test_replicated_worlds_match_single_world(gpu_api, oracle_api):
    """config 5 (batched independent worlds): 16 replicas of the Pyramid inside one device world evolve exactly like the
    single world (replica-local colouring priorities => bit-identical), never interact, and all fall asleep."""
    single, _ = scenes.pyramid(api=gpu_api)
    batch, _ = scenes.pyramid(api=gpu_api)
    copies = 16
    batch.Replicate(copies)
    assert gpu_api.world_replica_count(batch._w) == copies
    nb = single.counts().bodies
    assert batch.counts().bodies == nb * copies
    for k in range(6):
        single.StepN(DT, 8, 3, 50)
        batch.StepN(DT, 8, 3, 50)
        cs, cb = single.counts(), batch.counts()
        assert cb.contacts == cs.contacts * copies and cb.touching == cs.touching * copies, (k, cs.contacts, cb.contacts)
        assert cb.awakeBodies == cs.awakeBodies * copies
        assert gpu_api.world_debug_colour_conflicts(batch._w) == 0
        ss, _ = single.read_bodies()
        sb, n = batch.read_bodies()
        assert n == nb * copies
        for r in (0, 5, copies - 1):
            for i in range(nb):
                a, b = ss[i], sb[r * nb + i]
                assert (a.c.x, a.c.y, a.a, a.v.x, a.v.y, a.w) == (b.c.x, b.c.y, b.a, b.v.x, b.v.y, b.w), (k, r, i)
    assert batch.counts().awakeBodies == 0


def _perturbation(nb, seed):
    import numpy as np
    rng = np.random.RandomState(seed)
    pose = np.zeros((nb, 4), np.float32)
    vel = np.zeros((nb, 4), np.float32)
    vel[:, 0] = rng.uniform(-2, 2, nb); vel[:, 1] = rng.uniform(-1, 1, nb); vel[:, 2] = rng.uniform(-1, 1, nb)
    return rng, pose, vel


def test_bulk_set_body_states_equals_per_body_calls(gpu_api):
    """dbx_world_set_body_states == b2Body.SetTransform + SetLinearVelocity + SetAngularVelocity per body (b2body.d:261-326)"""
    import numpy as np
    a, bodies_a = scenes.pyramid(api=gpu_api)
    b, bodies_b = scenes.pyramid(api=gpu_api)
    a.StepN(DT, 8, 3, 5); b.StepN(DT, 8, 3, 5)
    sa, n = a.read_bodies()
    rng = np.random.RandomState(3)
    ids = np.array(sorted(rng.choice(np.arange(2, n), 40, replace=False)), np.int32)
    pose = np.zeros((len(ids), 4), np.float32); vel = np.zeros((len(ids), 4), np.float32)
    for k, i in enumerate(ids):
        pose[k, :3] = (sa[i].p.x + rng.uniform(-0.3, 0.3), sa[i].p.y + rng.uniform(0, 0.5), rng.uniform(-0.2, 0.2))
        vel[k, :3] = (rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-2, 2))
    a.SetBodyStates(ids, pose, vel)
    by_id = {bd.id: bd for bd in bodies_b}
    for k, i in enumerate(ids):
        bd = by_id[int(i)]
        bd.SetTransform((float(pose[k, 0]), float(pose[k, 1])), float(pose[k, 2]))
        bd.SetLinearVelocity((float(vel[k, 0]), float(vel[k, 1])))
        bd.SetAngularVelocity(float(vel[k, 2]))
    for step in range(30):
        a.Step(DT, 8, 3); b.Step(DT, 8, 3)
        ca, cb = a.counts(), b.counts()
        assert (ca.contacts, ca.touching, ca.awakeBodies) == (cb.contacts, cb.touching, cb.awakeBodies), step
    sa, _ = a.read_bodies(); sb, _ = b.read_bodies()
    for i in range(n):
        assert (sa[i].c.x, sa[i].c.y, sa[i].a, sa[i].v.x, sa[i].v.y, sa[i].w) == (sb[i].c.x, sb[i].c.y, sb[i].a, sb[i].v.x, sb[i].v.y, sb[i].w), i


def test_divergent_replicas_match_separately_built_worlds(gpu_api):
    """batched worlds that are randomised per replica (the RL reset) evolve exactly like the same world built alone"""
    import numpy as np
    copies = 6
    batch, _ = scenes.pyramid(api=gpu_api)
    batch.Replicate(copies)
    nb = batch.counts().bodies // copies
    vel_all = np.zeros((nb * copies, 4), np.float32)
    singles = []
    for r in range(copies):
        rng, pose, vel = _perturbation(nb, 100 + r)
        vel[:2] = 0                                  # the two static bodies
        vel_all[r * nb:(r + 1) * nb] = vel
        w, _ = scenes.pyramid(api=gpu_api)
        w.SetBodyStates(None, None, vel)
        singles.append(w)
    batch.SetBodyStates(None, None, vel_all)
    for k in range(4):
        batch.StepN(DT, 8, 3, 40)
        sb, n = batch.read_bodies()
        for r, w in enumerate(singles):
            w.StepN(DT, 8, 3, 40)
            ss, _ = w.read_bodies()
            for i in range(nb):
                a, b = ss[i], sb[r * nb + i]
                assert (a.c.x, a.c.y, a.a, a.v.x, a.v.y, a.w) == (b.c.x, b.c.y, b.a, b.v.x, b.v.y, b.w), (k, r, i)
    assert gpu_api.world_debug_colour_conflicts(batch._w) == 0


def _event_key(ev):
    # (type, phase, stepsAgo, fixtureA, fixtureB, childA, childB, bodyA, bodyB) -> order-free identity of one callback
    return (ev[0], ev[1], ev[3], ev[4], ev[5], ev[6], ev[7], ev[8])


def test_contact_events_match_reference_callbacks(gpu_api, oracle_api):
    """b2ContactListener.BeginContact / EndContact (b2contact.d:338-346, b2contactmanager.d:60-63): per step, the deferred
    device events are exactly the callbacks the reference makes (same fixtures in the same A/B order, Collide vs TOI)."""
    g, gb = scenes.pyramid(api=gpu_api, count=8)
    o, ob = scenes.pyramid(api=oracle_api, count=8)
    g.EnableContactEvents(4096); o.EnableContactEvents(4096)
    total = 0
    for step in range(23):      # free fall and the first impacts: later a marginal contact may flicker on one side only
        g.Step(DT, 8, 3); o.Step(DT, 8, 3)
        eg, eo = g.PollContactEvents(), o.PollContactEvents()
        assert sorted(map(_event_key, eg)) == sorted(map(_event_key, eo)), step
        assert all(e[2] == 0 for e in eg)
        keys = [(e[1], ) for e in eg]
        assert keys == sorted(keys)                       # delivered in (phase, pair key, type) order
        total += len(eg)
    assert total > 30
    # the balance of begins and ends is the number of touching contacts
    g2, _ = scenes.pyramid(api=gpu_api, count=8)
    g2.EnableContactEvents(1 << 14)
    balance = 0
    for step in range(150):
        g2.Step(DT, 8, 3)
        for e in g2.PollContactEvents():
            balance += 1 if e[0] == 1 else -1
    assert balance == g2.counts().touching


def test_contact_events_on_destroy_and_listener_delivery(gpu_api):
    """DestroyBody ends its touching contacts (b2world.d:128-137 -> b2contactmanager.d:60-63); the Python mirror of
    b2ContactListener receives the deferred calls right after Step"""
    from dbox_b200.world import b2ContactListener

    class Log(b2ContactListener):
        def __init__(self):
            self.begin, self.end = [], []

        def BeginContact(self, c):
            self.begin.append((c.GetFixtureA().id, c.GetFixtureB().id))

        def EndContact(self, c):
            self.end.append((c.GetFixtureA().id, c.GetFixtureB().id))

    w, box = scenes.hello_world(api=gpu_api)
    log = Log()
    w.SetContactListener(log)
    for _ in range(90):
        w.Step(DT, 6, 2)
    assert len(log.begin) == 1 and log.end == []
    w.DestroyBody(box)
    ev = w.PollContactEvents()
    assert len(ev) == 1 and ev[0][0] == 2 and ev[0][1] == 3 and (ev[0][3], ev[0][4]) == log.begin[0]
    # overflow is reported, not silent
    p, _ = scenes.pyramid(api=gpu_api, count=10)
    p.EnableContactEvents(4)
    for _ in range(40):
        p.Step(DT, 8, 3)
    with pytest.raises(Exception):
        p.PollContactEvents()


def test_tumbler_invariants(gpu_api, oracle_api):
    """config 3 (tumbler.d:40-97): a motor-driven container (one dynamic body touched by hundreds of contacts -> overflow
    colour lanes), one new body per step (host mirror pushes while stepping, contact pool grows on its own), boxes and
    circles.  Chaotic, so judged by invariants against the oracle: nothing leaks out, the container follows the motor,
    contact populations agree."""
    n, steps = 300, 500
    tg = scenes.Tumbler(api=gpu_api, count=n)
    to = scenes.Tumbler(api=oracle_api, count=n)
    for _ in range(steps):
        tg.Step(); to.Step()
    assert tg.m_count == to.m_count == n
    for t in (tg, to):
        c = t.container._state()
        for b in t.bodies:
            p = b.GetPosition()
            dx, dy = p.x - c.p.x, p.y - c.p.y
            lx, ly = c.qc * dx + c.qs * dy, -c.qs * dx + c.qc * dy          # container frame
            assert abs(lx) < 10.6 and abs(ly) < 10.6, (b.id, lx, ly)
    assert abs(tg.container.GetAngle() - to.container.GetAngle()) < 1e-3
    assert abs(tg.container.GetAngularVelocity() - to.container.GetAngularVelocity()) < 1e-3
    cg, co = tg.world.counts(), to.world.counts()
    assert cg.bodies == co.bodies and cg.joints == co.joints == 1
    assert abs(cg.contacts - co.contacts) < 0.15 * co.contacts, (cg.contacts, co.contacts)
    assert abs(cg.touching - co.touching) < 0.15 * co.touching, (cg.touching, co.touching)
    assert gpu_api.world_debug_colour_conflicts(tg.world._w) == 0
    mg = sum(b.GetPosition().y for b in tg.bodies) / n
    mo = sum(b.GetPosition().y for b in to.bodies) / n
    assert abs(mg - mo) < 0.5, (mg, mo)


def test_world_batch_api(gpu_api):
    """dbox_b200.batch.WorldBatch (what bench.py's batched leg drives): partitioned share, bulk reset / act / observe calls"""
    import numpy as np
    import ctypes as C
    from dbox_b200.batch import WorldBatch, partition
    total = 10
    first, count = partition(total, 1, 3)
    b = WorldBatch(lambda **kw: scenes.pyramid(count=6, **kw), total, rank=1, world_size=3, api=gpu_api, contacts_per_world=200)
    assert (b.first, b.count) == (first, count) == (4, 3) and b.n_bodies == b.bodies_per_world * 3
    vel = np.zeros((b.n_bodies, 4), np.float32); vel[:, 0] = 0.1
    b.set_states(vel=vel)
    forces = np.zeros((b.n_bodies, 4), np.float32)
    xf = np.zeros((b.n_bodies, 4), np.float32)
    for _ in range(20):
        b.apply_forces(forces.ctypes.data)
        b.step(DT, 8, 3)
        b.read_transforms(xf.ctypes.data)
    ms, stages = b.time_steps(DT, 8, 3, 5, flush_l2=False)
    assert ms > 0 and len(stages) == 9
    st = b.stats()
    assert st["worlds"] == 3 and st["bodies"] == b.n_bodies and st["contacts"] > 0
    per = b.bodies_per_world
    assert np.allclose(xf[:per], xf[per:2 * per]) and np.allclose(xf[:per], xf[2 * per:])      # same reset -> same worlds
    b.close()


def test_export_import_state_continues_bit_for_bit(gpu_api):
    """dbx_world_export_state / import_state: a world restored from a snapshot steps exactly like the one it was taken from"""
    a, _ = scenes.pyramid(api=gpu_api, count=10)
    a.StepN(DT, 8, 3, 45)
    need = gpu_api.world_export_state(a._w, None, 0)
    assert need > 0
    buf = (C.c_char * need)()
    assert gpu_api.world_export_state(a._w, buf, need) == need
    b, _ = scenes.pyramid(api=gpu_api, count=10)
    assert gpu_api.world_import_state(b._w, buf, need) == 0
    assert gpu_api.world_import_state(b._w, buf, 10) < 0            # truncated blob
    assert gpu_api.world_import_state(b._w, buf, need - 4) == A.DBX_E_INVALID      # length must match the header exactly
    # a hostile header: negative counts, or the counts of another scene, are refused before anything is written
    import struct
    for off, val in ((8, -1), (12, -5), (16, -1), (24, -2), (8, 3)):
        bad = (C.c_char * need).from_buffer_copy(buf)
        struct.pack_into("<i", bad, off, val)
        assert gpu_api.world_import_state(b._w, bad, need) == A.DBX_E_INVALID, (off, val)
    c, _ = scenes.pyramid(api=gpu_api, count=6)
    assert gpu_api.world_import_state(c._w, buf, need) == A.DBX_E_INVALID          # snapshot of a different scene
    c.close()
    sb0, n0 = b.read_bodies(); sa0, _ = a.read_bodies()
    assert all((sa0[i].c.x, sa0[i].c.y) == (sb0[i].c.x, sb0[i].c.y) for i in range(n0))    # the refused imports changed nothing
    for k in range(5):
        a.StepN(DT, 8, 3, 20); b.StepN(DT, 8, 3, 20)
        sa, n = a.read_bodies(); sb, _ = b.read_bodies()
        ca, cb = a.counts(), b.counts()
        assert (ca.contacts, ca.touching, ca.awakeBodies) == (cb.contacts, cb.touching, cb.awakeBodies)
        for i in range(n):
            assert (sa[i].c.x, sa[i].c.y, sa[i].a, sa[i].v.x, sa[i].v.y, sa[i].w) == (sb[i].c.x, sb[i].c.y, sb[i].a, sb[i].v.x, sb[i].v.y, sb[i].w), (k, i)
