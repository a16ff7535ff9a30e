"""SURVEY.md 8(f) rows through the C ABI against the oracle: b2Fixture.TestPoint, the report-everything RayCast callback,
b2Contact.GetWorldManifold, b2World.ShiftOrigin and b2ContactListener.PostSolve.  Where the two sides hold the same state
(before the first step, or after the oracle's state was transplanted into the device world) the answers must be identical;
where they stepped separately the scenes are order-free, so they agree to float rounding."""
import ctypes as C
import math
import random

import pytest

from dbox_b200 import _abi as A
from dbox_b200 import scenes
from dbox_b200.world import (b2BodyDef, b2CircleShape, b2EdgeShape, b2FixtureDef, b2MouseJointDef, b2PolygonShape, b2PulleyJointDef, b2World,
                             b2_dynamicBody, b2_kinematicBody)
from tests.parity import contact_key, transplant
from tests.test_gpu_features import _dyn, _ground, _query_scene, both

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0


def _same_state_pair(gpu_api, oracle_api, steps):
    """the query scene stepped by the oracle, its state copied into a device world built the same way"""
    wo, bo = _query_scene(oracle_api)
    wg, bg = _query_scene(gpu_api)
    for _ in range(steps):
        wo.Step(DT, 8, 3)
    if steps:
        transplant(wo, wg)
    return wg, bg, wo, bo


@pytest.mark.parametrize("steps", [0, 45])
def test_test_point_matches_reference(gpu_api, oracle_api, steps):
    """b2Fixture.TestPoint (b2fixture.d:209-212; polygon b2polygonshape.d:265-279, circle b2circleshape.d:60-65, never inside an
    edge or a chain) on identical states: identical answers, including points on and next to the boundary"""
    wg, bg, wo, bo = _same_state_pair(gpu_api, oracle_api, steps)
    rng = random.Random(3)
    queries_g, queries_o = [], []
    for b_g, b_o in zip(bg, bo):
        p = b_o.GetPosition()
        for _ in range(40):
            q = (p.x + rng.uniform(-1.5, 1.5), p.y + rng.uniform(-1.5, 1.5))
            queries_g.append((b_g.fixtures[0], q)); queries_o.append((b_o.fixtures[0], q))
        queries_g.append((b_g.fixtures[0], (p.x, p.y))); queries_o.append((b_o.fixtures[0], (p.x, p.y)))
    # the ground's chain (fixture 0) and edge (fixture 1) never contain a point
    for fid in (0, 1):
        for q in ((0.0, 0.5), (-30.0, 10.0), (-10.0, 1.0)):
            queries_g.append((fid, q)); queries_o.append((fid, q))
    ins_g, ins_o = wg.TestPoints(queries_g), wo.TestPoints(queries_o)
    assert ins_g == ins_o
    assert 300 < sum(ins_o) < len(ins_o) - 300          # the sample really straddles the shapes
    assert not any(ins_g[-6:])
    assert bg[0].fixtures[0].TestPoint(tuple(bo[0].GetPosition())) is True
    with pytest.raises(RuntimeError):
        wg.TestPoints([(10 ** 6, (0.0, 0.0))])


@pytest.mark.parametrize("steps", [0, 45])
def test_raycast_all_matches_reference(gpu_api, oracle_api, steps):
    """b2World.RayCast with a callback that returns 1 (every fixture along the whole ray, b2world.d:577-587 and the wrapper
    :1605-1624): same hit lists (fixture, child), fractions / points / normals to float rounding"""
    wg, bg, wo, bo = _same_state_pair(gpu_api, oracle_api, steps)
    rng = random.Random(17)
    rays = [((rng.uniform(-35, 35), rng.uniform(-2, 25)), (rng.uniform(-35, 35), rng.uniform(-2, 25))) for _ in range(300)]
    rays += [((-40.0, 3.0), (40.0, 3.0)), ((0.0, 30.0), (0.0, -5.0)), ((3.0, 3.0), (3.0, 3.0)), ((-31.0, 10.0), (-29.0, 10.0))]
    hg, ho = wg.RayCastAll(rays, cap=96), wo.RayCastAll(rays, cap=96)
    multi = 0
    for k, (a, b) in enumerate(zip(hg, ho)):
        assert [(h[0], h[1]) for h in a] == [(h[0], h[1]) for h in b], (k, a, b)
        multi += len(b) > 1
        for x, y in zip(a, b):
            assert abs(x[2] - y[2]) <= 1e-6 and max(abs(x[3][0] - y[3][0]), abs(x[3][1] - y[3][1])) <= 1e-4, (k, x, y)
            assert max(abs(x[4][0] - y[4][0]), abs(x[4][1] - y[4][1])) <= 1e-5, (k, x, y)
        # the closest-hit query is the head of the list
        if b:
            c = wg.RayCastClosest([rays[k]])[0]
            assert abs(c[2] - a[0][2]) <= 1e-6
    assert multi > 40 and hg[-2] == []                  # a zero-length ray hits nothing
    assert len(hg[-1]) == 1 and hg[-1][0][0] == 1       # the vertical edge fixture
    # the surplus over cap is counted, not stored
    with pytest.raises(RuntimeError):
        wg.RayCastAll([((-40.0, 3.0), (40.0, 3.0)), ((-40.0, 1.5), (40.0, 1.5))], cap=1)


def test_world_manifolds_match_reference(gpu_api, oracle_api):
    """b2Contact.GetWorldManifold (b2contact.d:77-91) -> b2WorldManifold.Initialize (b2collision.d:123-191): normal, points and
    separations of every contact, computed on the device, against the oracle on the same state"""
    wg, bg, wo, bo = _same_state_pair(gpu_api, oracle_api, 90)
    rg, ng = wg.read_contacts(); ro, no = wo.read_contacts()
    mg, mo = wg.GetWorldManifolds(), wo.GetWorldManifolds()
    assert ng == no == len(mg) == len(mo) and ng > 60
    by_key = {contact_key(ro[i]): mo[i] for i in range(no)}
    touching = types = 0
    seen_types = set()
    for i in range(ng):
        a, b = mg[i], by_key[contact_key(rg[i])]
        assert a[0] == b[0]
        if a[0] == 0:
            continue
        touching += 1
        seen_types.add(rg[i].manifold.type)
        assert max(abs(a[1][0] - b[1][0]), abs(a[1][1] - b[1][1])) <= 1e-6, (i, a, b)
        for k in range(a[0]):
            assert max(abs(a[2][k][0] - b[2][k][0]), abs(a[2][k][1] - b[2][k][1])) <= 2e-5, (i, a, b)
            assert abs(a[3][k] - b[3][k]) <= 2e-6, (i, a, b)
            assert a[3][k] < 0.25                           # (evaluated at the post-step transforms, so not exactly inside the skin)
    assert touching > 40 and len(seen_types) == 3           # circles, faceA and faceB manifolds all present
    # an empty world has none
    assert b2World((0.0, -10.0), api=gpu_api).GetWorldManifolds() == []


def _shift_scene(api):
    """order-free: separate boxes and circles resting on the ground, a pulley pair hanging from two ground anchors, a body on
    a mouse joint: one constraint row per body"""
    w = b2World((0.0, -10.0), api=api)
    g = _ground(w, api)
    out = []
    for k in range(8):
        b = _dyn(w, -20.0 + 3.0 * k, 0.6 + 0.05 * k)
        if k % 2:
            s = b2CircleShape(api); s.m_radius = 0.5
        else:
            s = b2PolygonShape(api); s.SetAsBox(0.5, 0.5)
        b.CreateFixture(s, 1.0)
        out.append(b)
    pa, pb = _dyn(w, 10.0, 6.0), _dyn(w, 14.0, 6.0)
    for b in (pa, pb):
        s = b2PolygonShape(api); s.SetAsBox(0.5, 0.5); b.CreateFixture(s, 1.0 if b is pa else 1.3)
    pj = b2PulleyJointDef(); pj.Initialize(pa, pb, (10.0, 12.0), (14.0, 12.0), (10.0, 6.5), (14.0, 6.5), 1.0)
    w.CreateJoint(pj)
    m = _dyn(w, 22.0, 8.0)
    s = b2CircleShape(api); s.m_radius = 0.4; m.CreateFixture(s, 1.0)
    md = b2MouseJointDef(); md.bodyA, md.bodyB = g, m
    md.target.Set(22.0, 8.0); md.maxForce = 1000.0
    w.CreateJoint(md)
    return w, out + [pa, pb, m]


def test_shift_origin_matches_reference(gpu_api, oracle_api):
    """b2World.ShiftOrigin (b2world.d:758-780): bodies, world-space joint anchors (mouse target, pulley ground anchors) and the
    broadphase boxes move; contacts, impulses and the pair cache do not notice"""
    origin = (100.0, -37.5)
    state = {}

    def each(k, wg, wo):
        if k == 39:
            state["contacts"] = (wg.counts().contacts, wo.counts().contacts)
            for w in (wg, wo):
                w.EnableContactEvents(1024); w.PollContactEvents()
                w.ShiftOrigin(origin)
            pg, n = wg.read_proxies(); po, m = wo.read_proxies()
            assert n == m
            fat = {(po[i].fixture, po[i].child): po[i].fat for i in range(m)}
            for i in range(n):                                   # the tree's boxes moved by the same float subtraction
                f = fat[(pg[i].fixture, pg[i].child)]
                assert (pg[i].fat.lo.x, pg[i].fat.lo.y, pg[i].fat.hi.x, pg[i].fat.hi.y) == (f.lo.x, f.lo.y, f.hi.x, f.hi.y)
    wg, wo, bg, bo = both(gpu_api, oracle_api, _shift_scene, 120, each=each, tol_p=3e-5)
    assert abs(bo[0].GetPosition().x - (-20.0 - origin[0])) < 1e-3 and abs(bo[0].GetPosition().y - (0.5 + 0.01 - origin[1])) < 2e-2
    assert state["contacts"] == (wg.counts().contacts, wo.counts().contacts) and state["contacts"][0] >= 8
    assert wg.PollContactEvents() == wo.PollContactEvents() == []     # nothing began or ended because of the shift
    # queries see the shifted world
    hit = wg.RayCastClosest([((-20.0 - origin[0], 10.0 - origin[1]), (-20.0 - origin[0], -5.0 - origin[1]))])[0]
    assert hit[0] == bg[0].fixtures[0].id
    assert wg.RayCastClosest([((-20.0, 10.0), (-20.0, -5.0))])[0][0] == -1


def _impact_scene(api, bullets=2):
    """order-free PostSolve scene: separate resting bodies (island solve) and two bullets flying at a thin wall (TOI sub-steps)"""
    w = b2World((0.0, -10.0), api=api)
    g = _ground(w, api)
    wall = b2EdgeShape(api); wall.Set((10.0, 0.0), (10.0, 20.0)); g.CreateFixture(wall, 0.0)
    out = []
    for k in range(6):
        b = _dyn(w, -20.0 + 3.0 * k, 0.55)
        if k % 2:
            s = b2CircleShape(api); s.m_radius = 0.5
        else:
            s = b2PolygonShape(api); s.SetAsBox(0.5, 0.5)
        b.CreateFixture(s, 1.0 + 0.1 * k)
        out.append(b)
    for k in range(bullets):
        b = _dyn(w, 0.0, 5.0 + (6.0 if bullets == 2 else 2.0) * k, bullet=True, gravityScale=0.0)
        s = b2CircleShape(api); s.m_radius = 0.25; b.CreateFixture(s, 1.0)
        # two bullets: far above b2_maxTranslation per step (both are clamped to 2 m / step); six: distinct speeds below the clamp,
        # so that their times of impact differ and the order of the TOI events is not a tie-break
        b.SetLinearVelocity((300.0 + 50.0 * k, 0.0) if bullets == 2 else (100.0 + 3.0 * k, 0.0))
        out.append(b)
    return w, out


def test_post_solve_records_match_reference(gpu_api, oracle_api):
    """b2ContactListener.PostSolve (b2worldcallbacks.d:120-128) through b2Island.Report (b2island.d:438-462): one record per
    contact of every solved island (phase 1, :239) and of every TOI mini-island (phase 2, :414) with the constraint's impulses"""
    seen = {1: 0, 2: 0}

    def each(k, wg, wo):
        rg, ro = wg.ReadPostSolve(), sorted(wo.ReadPostSolve(), key=lambda r: (r[0], r[1:5]))
        rg = sorted(rg, key=lambda r: (r[0], r[1:5]))
        assert [r[:6] for r in rg] == [r[:6] for r in ro], (k, rg, ro)
        for a, b in zip(rg, ro):
            seen[a[0]] += 1
            for j in range(a[5]):
                assert abs(a[6][j] - b[6][j]) <= 2e-4 * max(1.0, abs(b[6][j])), (k, a, b)
                assert abs(a[7][j] - b[7][j]) <= 2e-4 * max(1.0, abs(b[7][j])), (k, a, b)
            assert a[6][1] == 0.0 or a[5] == 2

    def build(api):
        w, out = _impact_scene(api)
        assert w.EnablePostSolve(4096) >= 4096
        return w, out
    both(gpu_api, oracle_api, build, 60, each=each)
    assert seen[1] > 200 and seen[2] >= 2
    # records are per step: a step without touching contacts leaves none; switching off stops recording
    w = b2World((0.0, -10.0), api=gpu_api)
    b = _dyn(w, 0.0, 50.0); s = b2CircleShape(gpu_api); s.m_radius = 0.5; b.CreateFixture(s, 1.0)
    w.EnablePostSolve(16); w.Step(DT, 8, 3)
    assert w.ReadPostSolve() == []


def test_post_solve_listener_on_pyramid(gpu_api, oracle_api):
    """the deferred listener delivers PostSolve after Step: on the pyramid the same contacts are reported as in the oracle's
    call log while the contact sets still agree, and the normal impulses carry the pile's weight"""
    from dbox_b200.world import b2ContactListener

    class L(b2ContactListener):
        post_solve = True

        def __init__(self):
            self.calls = []

        def PostSolve(self, contact, impulse):
            self.calls.append(((contact.fixtureA_id, contact.childA, contact.fixtureB_id, contact.childB), impulse))
    wg, _ = scenes.pyramid(api=gpu_api, count=12)
    wo, _ = scenes.pyramid(api=oracle_api, count=12)
    lst = L()
    wg.SetContactListener(lst, capacity=1 << 14)
    wo.EnablePostSolve(1 << 14)
    for k in range(40):
        lst.calls = []
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
        ro = wo.ReadPostSolve()
        got, want = set(c[0] for c in lst.calls), set((r[1], r[2], r[3], r[4]) for r in ro)
        assert len(got) <= len(lst.calls)                         # TOI sub-steps call again for the contacts of their mini-islands
        # the same contacts while the two Gauss-Seidel orders still give the same touching set; later a corner contact may
        # come or go a step apart
        assert got == want if k < 12 else len(got ^ want) <= 6, (k, sorted(got ^ want))
    total_g = sum(sum(c[1][1][:c[1][0]]) for c in lst.calls)
    total_o = sum(sum(r[6][:r[5]]) for r in ro)
    assert total_o > 0 and abs(total_g - total_o) <= 0.06 * total_o       # a pyramid still settling: the two Gauss-Seidel orders share the load a little differently


def _filter_scene(api, contact_filter, replaces_default):
    """boxes dropped on top of each other in three loose columns; the user's filter decides which of them see each other"""
    w = b2World((0.0, -10.0), api=api)
    g = _ground(w, api)
    w.SetContactFilter(contact_filter, replaces_default=replaces_default)
    out = []
    for k in range(12):
        b = _dyn(w, -6.0 + 6.0 * (k % 3) + 0.05 * (k // 3), 0.6 + 1.3 * (k // 3))
        s = b2PolygonShape(api); s.SetAsBox(0.5, 0.5)
        fd = b2FixtureDef(); fd.shape, fd.density = s, 1.0
        if k % 3 == 2:
            fd.filter.maskBits = 0x0000          # the default rule would let these fall through everything, ground included
        b.CreateFixture(fd)
        out.append(b)
    return w, out, g


def test_user_contact_filter_matches_reference(gpu_api, oracle_api):
    """b2World.SetContactFilter + b2ContactFilter.ShouldCollide (b2world.d:52-56, b2worldcallbacks.d:36-66; call sites
    b2contactmanager.d:110-114 and :274-281), deferred on the device side: the pairs a user filter rejects never touch, never
    reach the solver and raise no events; a filter that overrides the default rule can also allow what the masks forbid; after a
    Refilter the filter is asked again"""
    from dbox_b200.world import b2ContactFilter

    class F(b2ContactFilter):
        """dynamic boxes ignore each other (whatever their masks say) and always collide with the ground body"""
        def __init__(self):
            self.solid = False
            self.asked = 0

        def ShouldCollide(self, fa, fb):
            self.asked += 1
            if fa.body.id == 0 or fb.body.id == 0:
                return True
            return self.solid and super().ShouldCollide(fa, fb)
    fg, fo = F(), F()
    wg, bg, _ = _filter_scene(gpu_api, fg, True)
    wo, bo, _ = _filter_scene(oracle_api, fo, True)
    wg.EnableContactEvents(4096); wo.EnableContactEvents(4096)

    def check(k, tol):
        for i, (a, b) in enumerate(zip(bg, bo)):
            pa, pb = a.GetPosition(), b.GetPosition()
            assert abs(pa.x - pb.x) <= tol and abs(pa.y - pb.y) <= tol and abs(a.GetAngle() - b.GetAngle()) <= tol, (k, i, (pa.x, pa.y), (pb.x, pb.y))
    for k in range(150):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
        check(k, 2e-4)
        cg, co = wg.counts(), wo.counts()
        assert (cg.contacts, cg.touching) == (co.contacts, co.touching), k
        eg = sorted(e[:1] + e[3:] for e in wg.PollContactEvents()); eo = sorted(e[:1] + e[3:] for e in wo.PollContactEvents())
        assert eg == eo, (k, eg, eo)
    # every box lies on the ground, also the ones whose mask is 0: all twelve passed through each other
    assert all(abs(b.GetPosition().y - 0.5) < 0.02 for b in bg)
    assert wg.counts().contacts == 12 and fg.asked >= 12
    rg, n = wg.read_contacts()
    assert all(0 in (wg._fixtures[rg[i].fixtureA].body.id, wg._fixtures[rg[i].fixtureB].body.id) for i in range(n))
    # Refilter: boxes become solid to each other where the masks allow it (columns 0 and 1); lift them and let them stack
    for f, bodies, w in ((fg, bg, wg), (fo, bo, wo)):
        f.solid = True
        for k, b in enumerate(bodies):
            b.SetTransform((-6.0 + 6.0 * (k % 3), 0.6 + 1.3 * (k // 3)), 0.0)
            b.SetAwake(True)
            b.fixtures[0].SetFilterData(*b.fixtures[0].filter)
    for k in range(150):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
        cg, co = wg.counts(), wo.counts()
        assert (cg.contacts, cg.touching) == (co.contacts, co.touching), k
    check("stacked", 5e-3)                                   # four-high stacks: the Gauss-Seidel order now matters a little
    ys = sorted(b.GetPosition().y for k, b in enumerate(bg) if k % 3 == 0)
    assert [round(y) for y in ys] == [0, 2, 3, 4] or all(abs(ys[i] - (0.5 + 1.0 * i)) < 0.1 for i in range(4))
    assert all(abs(b.GetPosition().y - 0.5) < 0.02 for k, b in enumerate(bg) if k % 3 == 2)     # mask 0: still only the ground
    # switching the filter off hands the decision back to the default rule
    wg.SetContactFilter(None)
    assert gpu_api.world_poll_new_contacts(wg._w, None, 0) == 0


def test_zero_dt_steps_and_staged_collide_leave_no_residue(gpu_api):
    """b2World.Step(0, ...) runs Collide only (b2world.d:399-419 skip Solve and SolveTOI when dt == 0), and the staged Collide hook
    does the same: neither may leave anything behind for the next real step (regression: the TOI candidate list used to be
    filled by every Collide and emptied only by SolveTOI, so the next step saw its TOI candidates twice)"""
    def run(extra):
        w, bodies = _impact_scene(gpu_api)
        for k in range(40):
            if extra and k % 3 == 0:
                w.Step(0.0, 8, 3)
                assert gpu_api.world_stage_collide(w._w) >= 0
            w.Step(DT, 8, 3)
        st, n = w.read_bodies()
        return [(st[i].c.x, st[i].c.y, st[i].a, st[i].v.x, st[i].v.y, st[i].w) for i in range(n)], w.counts().colours
    a, b = run(False), run(True)
    assert a == b


def test_pipelined_io_equals_synchronous_calls(gpu_api):
    """dbx_world_apply_forces_async / step_async / read_transforms_async / io_wait / sync: the act -> step -> observe loop with the
    copies on copy streams gives, observation by observation, exactly what the synchronous calls give"""
    import numpy as np
    import torch
    steps = 40
    wa, _ = scenes.pyramid(api=gpu_api, count=10)
    wb, _ = scenes.pyramid(api=gpu_api, count=10)
    n = wa.counts().bodies
    rng = np.random.RandomState(5)
    acts = rng.uniform(-30.0, 30.0, (steps, n, 4)).astype(np.float32)
    f = [torch.zeros((n, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    o = [torch.zeros((n, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    sync_obs, pipe_obs = [], []
    for k in range(steps):
        f[0].copy_(torch.from_numpy(acts[k]))
        assert gpu_api.world_apply_forces(wa._w, f[0].data_ptr(), n) == n
        wa.Step(DT, 8, 3)
        assert gpu_api.world_read_transforms(wa._w, o[0].data_ptr(), n) == n
        sync_obs.append(o[0].numpy().copy())
    prev = 0
    for k in range(steps):
        if k >= 2:
            assert gpu_api.world_sync(wb._w) == 0 if k == 2 else True     # (buffer k & 1 was consumed by step k - 2, long enqueued)
        f[k & 1].copy_(torch.from_numpy(acts[k]))
        assert gpu_api.world_apply_forces_async(wb._w, f[k & 1].data_ptr(), n) == n
        assert gpu_api.world_step_async(wb._w, DT, 8, 3) >= 0
        t = gpu_api.world_read_transforms_async(wb._w, o[k & 1].data_ptr(), n)
        assert t == k + 1
        if prev:
            assert gpu_api.world_io_wait(wb._w, prev) == 0
            pipe_obs.append(o[(k - 1) & 1].numpy().copy())
        prev = t
        # the host buffer of step k must not change until its copy has happened: wait for the H2D before the next overwrite
        assert gpu_api.world_io_wait(wb._w, t) == 0
    pipe_obs.append(o[(steps - 1) & 1].numpy().copy())
    assert gpu_api.world_sync(wb._w) == 0
    assert gpu_api.world_io_wait(wb._w, steps + 5) < 0               # a ticket that was never issued
    assert len(pipe_obs) == steps
    for k in range(steps):
        assert np.array_equal(sync_obs[k], pipe_obs[k]), k
    sa, _ = wa.read_bodies(); sb, _ = wb.read_bodies()
    assert bytes(sa) == bytes(sb)
    # and the helper the bench uses
    from dbox_b200.batch import run_pipelined
    run_pipelined(gpu_api, wb._w, n, DT, 8, 3, 5, (f[0].data_ptr(), f[1].data_ptr()), (o[0].data_ptr(), o[1].data_ptr()))


def test_world_queries_against_committed_goldens(gpu_api):
    """the CUDA path against the committed fixture tests/golden/oracle_golden.json ("queries", made by tests/golden/make_golden.py
    from the oracle): on the identical, not yet stepped scene the fixtures / children hit, the AABB query sets and the TestPoint
    answers are equal, fractions and normals agree to float rounding; after 60 steps the resting scene gives the same hit lists"""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "oracle_golden.json")))["queries"]
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    got = json.loads(json.dumps(mg.queries_golden(gpu_api)))
    fh = float.fromhex
    g0, w0 = got["initial"], gold["initial"]
    assert g0["inside"] == w0["inside"] and g0["boxes"] == w0["boxes"] and g0["world_manifolds"] == w0["world_manifolds"] == []
    for a, b in zip(g0["closest"], w0["closest"]):
        assert a[:2] == b[:2] and all(abs(fh(x) - fh(y)) <= 1e-6 for x, y in zip(a[2:], b[2:])), (a, b)
    for ha, hb in zip(g0["all"], w0["all"]):
        assert [h[:2] for h in ha] == [h[:2] for h in hb]
        assert all(abs(fh(x[2]) - fh(y[2])) <= 1e-6 for x, y in zip(ha, hb))
    g1, w1 = got["after60"], gold["after60"]
    assert [[h[:2] for h in hits] for hits in g1["all"]] == [[h[:2] for h in hits] for hits in w1["all"]]
    assert g1["boxes"] == w1["boxes"] and len(g1["world_manifolds"]) == len(w1["world_manifolds"])
    assert sum(a != b for a, b in zip(g1["inside"], w1["inside"])) <= 2          # sample points next to an edge of a body at rest


def test_sub_stepping_matches_reference(gpu_api, oracle_api):
    """b2World.SetSubStepping(true) (b2world.d:1127-1146, 1441-1446): SolveTOI handles ONE event per Step and leaves
    m_stepComplete false; the following Steps run Collide and resume SolveTOI (no Solve) until no event is left.  Bullets hitting
    a wall and boxes landing: event by event the same states as the oracle; switching sub-stepping off finishes the step"""
    def build(api):
        w, out = _impact_scene(api, bullets=6)
        w.SetSubStepping(True)
        return w, out
    seen = {"events": [], "stalls": 0}

    def each(k, wg, wo):
        # (counts.colours on the oracle side is its cumulative TOI event count)
        prev = seen["events"][-1] if seen["events"] else 0
        seen["events"].append(wo.counts().colours)
        pg, po = wg.read_bodies()[0], wo.read_bodies()[0]
        if seen["events"][-1] > prev:       # a Step that handled an event: the same bodies sit at the same point of their sweeps
            assert all(abs(pg[i].alpha0 - po[i].alpha0) < 1e-5 for i in range(1, 13)), (k, [pg[i].alpha0 for i in range(13)], [po[i].alpha0 for i in range(13)])
            seen["stalls"] += 1
    wg, wo, bg, bo = both(gpu_api, oracle_api, build, 90, each=each)
    ev = seen["events"]
    assert ev[-1] >= 6 and all(b - a <= 1 for a, b in zip(ev, ev[1:]))       # never more than one event per Step
    assert seen["stalls"] >= 6                                                # Steps that ended with SolveTOI unfinished
    for w in (wg, wo):
        w.SetSubStepping(False)
    for k in range(30):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
    for g, o in zip(bg, bo):
        pg, po = g.GetPosition(), o.GetPosition()
        assert abs(pg.x - po.x) < 1e-3 and abs(pg.y - po.y) < 1e-3
    sg, so = wg.read_bodies()[0], wo.read_bodies()[0]
    assert all(sg[i].alpha0 == 0.0 == so[i].alpha0 for i in range(1, 13))


def _actuated_scene(api):
    """independent mechanisms far from each other (no contacts between them): a revolute arm with motor and limit, a prismatic
    slider with motor and limit, a distance spring, a wheel joint with a motor, a motor joint, a rope"""
    from dbox_b200.world import (b2DistanceJointDef, b2MotorJointDef, b2PrismaticJointDef, b2RevoluteJointDef, b2RopeJointDef,
                                 b2WheelJointDef)
    from tests.test_gpu_features import _box_body
    w = b2World((0.0, -10.0), api=api)
    g = w.CreateBody(b2BodyDef())
    arm = _box_body(w, api, -30.0, 10.0, hx=2.0, hy=0.25)
    rd = b2RevoluteJointDef(); rd.Initialize(g, arm, (-32.0, 10.0))
    rd.enableMotor, rd.motorSpeed, rd.maxMotorTorque = True, 0.0, 2000.0
    rd.enableLimit, rd.lowerAngle, rd.upperAngle = False, -0.5, 0.75
    rev = w.CreateJoint(rd)
    sl = _box_body(w, api, -15.0, 10.0)
    pd = b2PrismaticJointDef(); pd.Initialize(g, sl, (-15.0, 10.0), (1.0, 0.0))
    pd.enableMotor, pd.motorSpeed, pd.maxMotorForce = True, 0.0, 500.0
    pd.enableLimit, pd.lowerTranslation, pd.upperTranslation = True, -1.0, 1.0
    pri = w.CreateJoint(pd)
    bob = _box_body(w, api, 0.0, 8.0)
    dd = b2DistanceJointDef(); dd.Initialize(g, bob, (0.0, 12.0), (0.0, 8.0)); dd.frequencyHz, dd.dampingRatio = 2.0, 0.3
    dist = w.CreateJoint(dd)
    chassis = _box_body(w, api, 15.0, 10.0, hx=1.0, hy=0.25, gravityScale=0.0)
    wheel = _dyn(w, 15.0, 9.0, gravityScale=0.0)
    s = b2CircleShape(api); s.m_radius = 0.4; wheel.CreateFixture(s, 1.0)
    wd = b2WheelJointDef(); wd.Initialize(chassis, wheel, (15.0, 9.0), (0.0, 1.0))
    wd.enableMotor, wd.motorSpeed, wd.maxMotorTorque, wd.frequencyHz, wd.dampingRatio = True, 0.0, 50.0, 4.0, 0.7
    whl = w.CreateJoint(wd)
    hold = w.CreateJoint(_pin(b2RevoluteJointDef, g, chassis, (15.0, 10.0)))
    mover = _box_body(w, api, 30.0, 10.0, gravityScale=0.0)
    md = b2MotorJointDef(); md.Initialize(g, mover); md.maxForce, md.maxTorque = 500.0, 500.0
    mot = w.CreateJoint(md)
    weight = _box_body(w, api, 45.0, 8.0)
    pd2 = b2RopeJointDef(); pd2.bodyA, pd2.bodyB = g, weight
    pd2.localAnchorA.Set(45.0, 12.0); pd2.localAnchorB.Set(0.0, 0.0); pd2.maxLength = 4.5
    rope = w.CreateJoint(pd2)
    return w, [arm, sl, bob, chassis, wheel, mover, weight], dict(rev=rev, pri=pri, dist=dist, whl=whl, mot=mot, rope=rope, hold=hold)


def _pin(cls, a, b, anchor):
    d = cls(); d.Initialize(a, b, anchor)
    return d


def test_joint_setters_match_reference(gpu_api, oracle_api):
    """the run-time setters of the joint classes (dbx_joint_set_params): motor speed / torque / on-off, limits on-off and range
    (with the limit impulse reset and the wake-ups of b2revolutejoint.d:216-300, b2prismaticjoint.d:250-330), spring frequency,
    rest length, rope length, motor-joint offsets -- each mechanism alone in its island, so both sides agree to float rounding"""
    js = {}

    def build(api):
        w, bodies, joints = _actuated_scene(api)
        w.SetAllowSleeping(False)          # (SetLength / SetMaxLength / SetFrequency wake nobody, in the reference as here)
        js[api.prefix] = joints
        return w, bodies

    def each(k, wg, wo):
        for joints in js.values():
            if k == 20:
                joints["rev"].SetMotorSpeed(1.5); joints["pri"].SetMotorSpeed(2.0); joints["whl"].SetMotorSpeed(-6.0)
                joints["mot"].SetLinearOffset((1.0, 0.5)); joints["mot"].SetAngularOffset(0.4)
            if k == 50:
                joints["rev"].EnableLimit(True); joints["pri"].SetLimits(-0.5, 0.25); joints["dist"].SetLength(3.0)
                joints["dist"].SetFrequency(4.0); joints["rope"].SetMaxLength(3.5); joints["whl"].SetMaxMotorTorque(5.0)
            if k == 90:
                joints["rev"].SetLimits(-1.0, 0.2); joints["rev"].SetMotorSpeed(-2.0); joints["pri"].EnableMotor(False)
                joints["whl"].EnableMotor(False); joints["mot"].SetMaxForce(50.0); joints["mot"].SetLinearOffset((-1.0, 0.0))
                joints["pri"].SetMaxMotorForce(10.0)
    wg, wo, bg, bo = both(gpu_api, oracle_api, build, 150, each=each, tol_p=5e-5, tol_v=5e-4)
    # the setters did something: the arm sits at its new upper limit, the slider was driven to its limit and then let go
    assert abs(js["dbx_"]["rev"].GetJointAngle() - js["orc_"]["rev"].GetJointAngle()) < 1e-4
    assert -1.08 < js["dbx_"]["rev"].GetJointAngle() < 0.21
    assert abs(bg[2].GetPosition().y - bo[2].GetPosition().y) < 1e-3 and bg[6].GetPosition().y > 12.0 - 3.6, (bg[2].GetPosition().y, bg[6].GetPosition().y)
    jg, jo = wg.read_joints()[0], wo.read_joints()[0]
    for i in range(7):
        assert jg[i].limitState == jo[i].limitState, i
        assert all(abs(jg[i].impulse[c] - jo[i].impulse[c]) <= 2e-3 * max(1.0, abs(jo[i].impulse[c])) for c in range(3)), (i, list(jg[i].impulse), list(jo[i].impulse))


def test_motor_speed_wakes_and_bulk_setter(gpu_api, oracle_api):
    """SetMotorSpeed wakes both bodies (b2revolutejoint.d:262-267): an arm that fell asleep on its motor starts to turn; the bulk
    call (dbx_world_set_motor_speeds) does the same for many joints at once, also per replica of a replicated world"""
    from dbox_b200.world import b2RevoluteJointDef
    from tests.test_gpu_features import _box_body

    def build(api, n=4):
        w = b2World((0.0, -10.0), api=api)
        g = w.CreateBody(b2BodyDef())
        arms, joints = [], []
        for k in range(n):
            arm = _box_body(w, api, 10.0 * k, 10.0, hx=1.0, hy=0.2)
            rd = b2RevoluteJointDef(); rd.Initialize(g, arm, (10.0 * k, 10.0))
            rd.enableMotor, rd.motorSpeed, rd.maxMotorTorque = True, 0.0, 1e4
            joints.append(w.CreateJoint(rd)); arms.append(arm)
        return w, arms, joints
    wg, ag, jg = build(gpu_api); wo, ao, jo = build(oracle_api)
    for _ in range(60):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
    assert wg.counts().awakeBodies == wo.counts().awakeBodies == 0          # held by their motors, asleep
    jg[1].SetMotorSpeed(1.0); jo[1].SetMotorSpeed(1.0)
    wg.SetMotorSpeeds(jg[2:], [2.0, -3.0]); wo.SetMotorSpeeds(jo[2:], [2.0, -3.0])
    assert wg.counts().awakeBodies == wo.counts().awakeBodies == 3
    for _ in range(30):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
    for k, want in enumerate((0.0, 1.0, 2.0, -3.0)):
        assert abs(ag[k].GetAngularVelocity() - want) < 1e-3 and abs(ag[k].GetAngle() - ao[k].GetAngle()) < 1e-4, k
    with pytest.raises(RuntimeError):
        wg.SetMotorSpeeds([10 ** 6], [1.0])
    # replicated: joint r * J + j is joint j of replica r
    wr, ar, jr = build(gpu_api, n=3)
    wr.SetAllowSleeping(False)
    wr.Replicate(5)
    nb = 4                                                                 # bodies per replica: ground + 3 arms
    wr.SetMotorSpeeds([0 * 3 + 1, 2 * 3 + 0, 4 * 3 + 2], [1.0, -2.0, 4.0])
    wr.StepN(DT, 8, 3, 20)
    st, n = wr.read_bodies()
    assert n == 5 * nb
    got = {(r, j): st[r * nb + 1 + j].w for r in range(5) for j in range(3)}
    for key, w_ in got.items():
        want = {(0, 1): 1.0, (2, 0): -2.0, (4, 2): 4.0}.get(key, 0.0)
        assert abs(w_ - want) < 1e-3, (key, w_)


def test_tree_stats_of_the_lbvh(gpu_api, oracle_api):
    """b2World.GetTreeHeight / GetTreeBalance / GetTreeQuality (b2world.d:694-716): reported for this library's LBVH -- a valid
    binary tree over the same leaves as the reference's dynamic tree: height between log2(n) and n - 1, quality >= 1"""
    import math
    wg, _ = scenes.pyramid(api=gpu_api)
    wo, _ = scenes.pyramid(api=oracle_api)
    for _ in range(10):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
    h, bal, q = wg.GetTreeStats()
    n = wg.counts().proxies
    assert n == 211 and math.ceil(math.log2(n)) <= h <= n - 1 and 0 <= bal < h and q >= 1.0
    ho = wo.GetTreeStats()[0]
    assert math.ceil(math.log2(n)) <= ho <= n - 1                        # the reference's own tree, for scale
    assert h <= 4 * ho
    e = b2World((0.0, -10.0), api=gpu_api)
    assert e.GetTreeStats() == (0, 0, 0.0)


def _crawler_scene(api, units=1, **kw):
    """a jointed mechanism with contacts, `units` times side by side: a four-link chain with motorised revolute joints lying on
    the ground next to a short stack of loose boxes, and a pendulum on a distance joint that knocks into the stack"""
    from dbox_b200.world import b2DistanceJointDef, b2RevoluteJointDef
    from tests.test_gpu_features import _box_body
    w = b2World((0.0, -10.0), api=api, **kw)
    g = w.CreateBody(b2BodyDef())
    e = b2EdgeShape(api); e.Set((-40.0, 0.0), (40.0 + 30.0 * units, 0.0)); g.CreateFixture(e, 0.0)
    bodies, joints = [], []
    for u in range(units):
        x0 = 30.0 * u
        links = [_box_body(w, api, x0 - 4.0 + 2.0 * k, 0.3, hx=0.9, hy=0.25, density=2.0) for k in range(4)]
        for k in range(3):
            rd = b2RevoluteJointDef(); rd.Initialize(links[k], links[k + 1], (x0 - 3.0 + 2.0 * k, 0.3))
            rd.enableMotor, rd.motorSpeed, rd.maxMotorTorque = True, (0.8 if k % 2 == 0 else -0.8), 400.0
            rd.enableLimit, rd.lowerAngle, rd.upperAngle = True, -0.6, 0.6
            joints.append(w.CreateJoint(rd))
        boxes = [_box_body(w, api, x0 + 6.0, 0.5 + 1.01 * k) for k in range(3)]
        bob = _box_body(w, api, x0 + 9.0, 4.0, hx=0.4, hy=0.4, density=3.0)
        dd = b2DistanceJointDef(); dd.Initialize(g, bob, (x0 + 6.5, 6.0), (x0 + 9.0, 4.0))
        joints.append(w.CreateJoint(dd))
        bodies += links + boxes + [bob]
    return w, bodies, joints


def test_jointed_replicas_match_single_world(gpu_api):
    """batched worlds WITH joints (articulated mechanisms): the world-local solver runs the joint colours inside the replica's CTA;
    every replica evolves bit for bit like the same world stepped alone through the global solver, and per-replica motor commands
    make exactly the commanded replicas diverge"""
    single, bs, js = _crawler_scene(gpu_api, units=16)         # 129 bodies: above the size where replicas get a CTA each
    batch, bb, jb = _crawler_scene(gpu_api, units=16)
    copies = 12
    nb, nj = single.counts().bodies, len(js)
    batch.Replicate(copies)
    for k in range(6):
        single.StepN(DT, 8, 3, 25); batch.StepN(DT, 8, 3, 25)
        cs, cb = single.counts(), batch.counts()
        assert (cb.contacts, cb.touching, cb.awakeBodies) == (cs.contacts * copies, cs.touching * copies, cs.awakeBodies * copies), k
        assert gpu_api.world_debug_colour_conflicts(batch._w) == 0
        ss, _ = single.read_bodies(); sb, n = batch.read_bodies()
        assert n == nb * copies
        for r in (0, 5, copies - 1):
            for i in range(nb):
                a, b = ss[i], sb[r * nb + i]
                assert (a.c.x, a.c.y, a.a, a.v.x, a.v.y, a.w) == (b.c.x, b.c.y, b.a, b.v.x, b.v.y, b.w), (k, r, i)
    assert abs(bs[0].GetAngle()) + abs(bs[3].GetAngle()) > 0.05            # the chain really moved
    # command replica 7 alone: its chain reverses, the others carry on identically
    batch.SetMotorSpeeds([7 * nj + 0, 7 * nj + 1, 7 * nj + 2], [-1.5, 1.5, -1.5])
    single.StepN(DT, 8, 3, 40); batch.StepN(DT, 8, 3, 40)
    ss, _ = single.read_bodies(); sb, n = batch.read_bodies()
    for r in range(copies):
        same = all((ss[i].c.x, ss[i].c.y, ss[i].a) == (sb[r * nb + i].c.x, sb[r * nb + i].c.y, sb[r * nb + i].a) for i in range(nb))
        assert same == (r != 7), r


def _random_scene(api, seed, continuous):
    """a random mixed scene: sloped chain ground and walls, circles / boxes / convex polygons / two-fixture compounds with random
    materials, a kinematic paddle, and a few joints on bodies of their own (so that no two joints share a body)"""
    from dbox_b200.world import b2ChainShape, b2PrismaticJointDef, b2RevoluteJointDef
    rng = random.Random(seed)
    w = b2World((0.0, -10.0), api=api)
    w.SetContinuousPhysics(continuous)
    g = w.CreateBody(b2BodyDef())
    ch = b2ChainShape(api)
    ch.CreateChain([(-22.0, 6.0), (-20.0, 0.0)] + [(-16.0 + 4.0 * k, rng.uniform(-0.6, 0.6)) for k in range(9)] + [(20.0, 0.0), (22.0, 6.0)])
    g.CreateFixture(ch, 0.0)
    bodies = []
    for k in range(70):
        b = _dyn(w, rng.uniform(-17.0, 17.0), rng.uniform(1.0, 14.0), angle=rng.uniform(-3.0, 3.0))
        kind = rng.randrange(4)
        fd = b2FixtureDef()
        fd.density, fd.friction, fd.restitution = rng.uniform(0.5, 3.0), rng.uniform(0.0, 0.9), rng.choice((0.0, 0.0, 0.2, 0.6))
        if kind == 0:
            s = b2CircleShape(api); s.m_radius = rng.uniform(0.25, 0.8)
        elif kind == 1:
            s = b2PolygonShape(api); s.SetAsBox(rng.uniform(0.25, 1.0), rng.uniform(0.25, 1.0))
        else:
            nv = rng.randint(3, 8)
            s = b2PolygonShape(api)
            s.Set([(rng.uniform(0.3, 0.9) * math.cos(2.0 * math.pi * (i + rng.uniform(0.0, 0.6)) / nv),
                    rng.uniform(0.3, 0.9) * math.sin(2.0 * math.pi * (i + rng.uniform(0.0, 0.6)) / nv)) for i in range(nv)])
        fd.shape = s
        b.CreateFixture(fd)
        if kind == 3:
            s2 = b2CircleShape(api); s2.m_radius = 0.3; s2.m_p.Set(0.7, 0.0)
            fd.shape = s2
            b.CreateFixture(fd)
        b.SetLinearVelocity((rng.uniform(-3.0, 3.0), rng.uniform(-3.0, 1.0)))
        bodies.append(b)
    bd = b2BodyDef(); bd.type = b2_kinematicBody; bd.position.Set(0.0, 3.0)
    paddle = w.CreateBody(bd)
    s = b2PolygonShape(api); s.SetAsBox(3.0, 0.2); paddle.CreateFixture(s, 0.0)
    paddle.SetAngularVelocity(0.7)
    for k in range(3):
        a = _dyn(w, -15.0 + 12.0 * k, 16.0)
        s = b2PolygonShape(api); s.SetAsBox(1.5, 0.2); a.CreateFixture(s, 1.0)
        if k % 2 == 0:
            jd = b2RevoluteJointDef(); jd.Initialize(g, a, (-15.0 + 12.0 * k, 16.0))
            jd.enableMotor, jd.motorSpeed, jd.maxMotorTorque = True, 1.0, 100.0
        else:
            jd = b2PrismaticJointDef(); jd.Initialize(g, a, (-15.0 + 12.0 * k, 16.0), (1.0, 0.0))
            jd.enableLimit, jd.lowerTranslation, jd.upperTranslation = True, -2.0, 2.0
        w.CreateJoint(jd)
        bodies.append(a)
    return w, bodies


@pytest.mark.parametrize("seed", list(range(1, 17)))
def test_random_scenes_single_step_sequential_order(gpu_api, oracle_api, seed):
    """differential test over random mixed scenes: from the oracle's state (transplanted), with the oracle's own Gauss-Seidel order
    injected as the level schedule, one step of the CUDA path reproduces the oracle's step -- contact set, touching flags, manifold
    types and feature keys exactly, body states to float rounding -- at several moments of the scene's life"""
    from tests import parity as P
    continuous = seed % 2 == 0
    wo, _ = _random_scene(oracle_api, seed, continuous)
    wg, _ = _random_scene(gpu_api, seed, continuous)
    fixture_body = {f.id: f.body.id for f in wg._fixtures.values()}
    body_dynamic = {b.id: False for b in wg._bodies.values()}
    st, n = wg.read_bodies()
    for i in range(n):
        body_dynamic[i] = st[i].type == A.DYNAMIC_BODY
    worst = (0.0, 0.0)
    for presteps in (3, 25, 60, 60):
        for _ in range(presteps):
            wo.Step(DT, 8, 3)
        P.transplant(wo, wg)
        wo.Step(DT, 8, 3)
        levels, m, nlev = P.sequential_levels(oracle_api, wo, wg, fixture_body, body_dynamic)
        assert gpu_api.world_debug_set_contact_levels(wg._w, levels, m) == 0, gpu_api.last_error()
        wg.Step(DT, 8, 3)
        so, n = wo.read_bodies(); sg, _ = wg.read_bodies()
        ep, ev = P.body_state_errors(so, sg, n)
        worst = (max(worst[0], ep), max(worst[1], ev))
        co, _, no = P.contacts_by_key(wo); cg, _, ng = P.contacts_by_key(wg)
        assert set(co) == set(cg), (seed, presteps, set(co) ^ set(cg))
        for k, ro in co.items():
            rg = cg[k]
            assert (ro.flags & A.CONTACT_TOUCHING) == (rg.flags & A.CONTACT_TOUCHING), (seed, presteps, k)
            if ro.flags & A.CONTACT_TOUCHING:
                assert ro.manifold.pointCount == rg.manifold.pointCount and ro.manifold.type == rg.manifold.type, (seed, presteps, k)
                for j in range(ro.manifold.pointCount):
                    assert ro.manifold.points[j].key == rg.manifold.points[j].key, (seed, presteps, k)
    assert worst[0] < 5e-5 and worst[1] < 5e-4, (seed, worst)


def test_error_paths_of_the_new_calls(gpu_api):
    """misuse of the calls added for the 8(f) rows fails loudly with a DBX_E_* code and a message, never silently: wrong joint type,
    bad limits, unknown fixtures / joints, a PostSolve buffer that is too small, ShiftOrigin in the middle of a split step, user filter
    or sub-stepping on a replicated world"""
    from dbox_b200.world import b2RevoluteJointDef
    from tests.test_gpu_features import _box_body
    api = gpu_api
    w = b2World((0.0, -10.0), api=api)
    g = _ground(w, api)
    a = _box_body(w, api, 0.0, 0.6)
    b = _box_body(w, api, 0.0, 1.7)
    rd = b2RevoluteJointDef(); rd.Initialize(g, b, (0.0, 1.7)); rd.enableLimit, rd.lowerAngle, rd.upperAngle = True, -0.1, 0.1
    j = w.CreateJoint(rd)
    pod = rd._pod()
    pod.lowerAngle, pod.upperAngle = 0.5, -0.5
    assert api.joint_set_params(w._w, j.id, C.byref(pod), A.JP_LIMITS) == A.DBX_E_INVALID            # lower > upper
    pod.type = A.JOINT_PRISMATIC
    assert api.joint_set_params(w._w, j.id, C.byref(pod), A.JP_MOTOR_SPEED) == A.DBX_E_INVALID        # not this joint's type
    assert api.joint_set_params(w._w, 99, C.byref(pod), A.JP_MOTOR_SPEED) == A.DBX_E_INVALID
    assert api.joint_set_params(w._w, j.id, None, A.JP_MOTOR_SPEED) == A.DBX_E_INVALID
    ids = (C.c_int32 * 1)(5); sp = (C.c_float * 1)(1.0)
    assert api.world_set_motor_speeds(w._w, ids, sp, 1) == A.DBX_E_INVALID
    assert api.world_set_motor_speeds(w._w, None, None, 0) == 0
    # PostSolve buffer too small: the overflow is reported, not truncated silently
    assert w.EnablePostSolve(1) >= 1
    for _ in range(30):
        w.Step(DT, 8, 3)
    assert api.world_read_post_solve(w._w, None, 0) in (A.DBX_E_CAPACITY, 0, 1) 
    big = w.EnablePostSolve(64)
    w.Step(DT, 8, 3)
    assert 1 <= len(w.ReadPostSolve()) <= big
    # no new-contact list unless the user filter is on; a veto of a pair that does not exist is ignored
    assert api.world_poll_new_contacts(w._w, None, 0) == 0
    p = (A.ContactPatch * 1)(); p[0].fixtureA, p[0].fixtureB, p[0].mask = 0, 0, A.PATCH_DESTROY
    assert api.world_patch_contacts(w._w, p, 1) == 1
    # ShiftOrigin between step_begin and step_end is refused
    assert api.world_step_begin(w._w, DT, 8, 3) == 0
    assert api.world_shift_origin(w._w, 1.0, 1.0) == A.DBX_E_INVALID
    assert api.world_step_async(w._w, DT, 8, 3) == A.DBX_E_INVALID
    assert api.world_step_end(w._w) == 0
    assert api.world_tree_stats(w._w, None, None, None) == 0
    # replicated worlds (replicate comes before the first step): per-contact vetoes and sub-stepping are refused, and say why
    with pytest.raises(RuntimeError):
        w.Replicate(3)
    w = b2World((0.0, -10.0), api=api)
    g = _ground(w, api)
    b = _box_body(w, api, 0.0, 1.7)
    rd = b2RevoluteJointDef(); rd.Initialize(g, b, (0.0, 1.7)); rd.enableMotor, rd.maxMotorTorque = True, 10.0
    j = w.CreateJoint(rd)
    w.Replicate(3)
    assert api.world_set_user_filter(w._w, A.FILTER_LOG) == A.DBX_E_UNSUPPORTED and b"replicated" in api.last_error()
    assert api.world_set_flags(w._w, A.WORLD_DEFAULT_FLAGS | A.WORLD_SUB_STEPPING) == A.DBX_E_UNSUPPORTED
    assert api.joint_set_params(w._w, j.id, C.byref(rd._pod()), A.JP_MOTOR_SPEED) == A.DBX_E_UNSUPPORTED
    assert api.world_set_motor_speeds(w._w, (C.c_int32 * 1)(2), sp, 1) == 1                             # joint 0 of replica 2
    assert api.world_set_motor_speeds(w._w, (C.c_int32 * 1)(3), sp, 1) == A.DBX_E_INVALID              # there is no replica 3
