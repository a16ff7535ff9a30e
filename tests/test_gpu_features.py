"""Feature-by-feature parity of the CUDA path (through the C ABI) against the oracle: each scene exercises one part of the
step pipeline in a configuration where the Gauss-Seidel order cannot matter (one constraint per body, or islands of one
body), so the two sides must agree to float rounding, not just statistically."""
import ctypes as C
import math

import pytest

from dbox_b200 import _abi as A
from dbox_b200 import scenes
from dbox_b200.world import (b2BodyDef, b2ChainShape, b2CircleShape, b2DistanceJointDef, b2EdgeShape, b2FixtureDef,
                             b2PolygonShape, b2RevoluteJointDef, b2World, b2_dynamicBody, b2_kinematicBody)

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0


def both(gpu_api, oracle_api, build, steps, vi=8, pi=3, tol_p=2e-5, tol_v=2e-4, each=None):
    """build(api) -> (world, [bodies to compare]); step both sides and compare positions / velocities every step"""
    wg, bg = build(gpu_api)
    wo, bo = build(oracle_api)
    for k in range(steps):
        wg.Step(DT, vi, pi); wo.Step(DT, vi, pi)
        if each:
            each(k, wg, wo)
        for i, (g, o) in enumerate(zip(bg, bo)):
            pg, po = g.GetPosition(), o.GetPosition()
            scale = max(1.0, abs(po.x), abs(po.y))
            assert abs(pg.x - po.x) <= tol_p * scale and abs(pg.y - po.y) <= tol_p * scale, (k, i, (pg.x, pg.y), (po.x, po.y))
            assert abs(g.GetAngle() - o.GetAngle()) <= tol_p * max(1.0, abs(o.GetAngle())), (k, i, g.GetAngle(), o.GetAngle())
            vg, vo = g.GetLinearVelocity(), o.GetLinearVelocity()
            vs = max(1.0, abs(vo.x), abs(vo.y))
            assert abs(vg.x - vo.x) <= tol_v * vs and abs(vg.y - vo.y) <= tol_v * vs, (k, i, (vg.x, vg.y), (vo.x, vo.y))
            assert abs(g.GetAngularVelocity() - o.GetAngularVelocity()) <= tol_v * max(1.0, abs(o.GetAngularVelocity())), (k, i)
    return wg, wo, bg, bo


def _ground(world, api):
    g = world.CreateBody(b2BodyDef())
    e = b2EdgeShape(api)
    e.Set((-40.0, 0.0), (40.0, 0.0))
    g.CreateFixture(e, 0.0)
    return g


def _dyn(world, x, y, angle=0.0, **kw):
    bd = b2BodyDef()
    bd.type = b2_dynamicBody
    bd.position.Set(x, y)
    bd.angle = angle
    for k, v in kw.items():
        setattr(bd, k, v)
    return world.CreateBody(bd)


def test_revolute_joint_limit_and_motor(gpu_api, oracle_api):
    """b2RevoluteJoint (joints/b2revolutejoint.d:130-495): point constraint + limit (3x3 block, all limit states) + motor"""
    def build(api):
        w = b2World((0.0, -10.0), api=api)
        g = _ground(w, api)
        out = []
        for k, (limit, motor) in enumerate(((False, False), (True, False), (False, True), (True, True))):
            b = _dyn(w, 6.0 * k - 9.0, 8.0)
            s = b2PolygonShape(api); s.SetAsBox(0.25, 1.5)
            b.CreateFixture(s, 2.0)
            jd = b2RevoluteJointDef()
            jd.Initialize(g, b, (6.0 * k - 9.0, 9.5))
            jd.enableLimit, jd.lowerAngle, jd.upperAngle = limit, -0.3, 0.45
            jd.enableMotor, jd.motorSpeed, jd.maxMotorTorque = motor, 1.5, 40.0
            w.CreateJoint(jd)
            b.SetAngularVelocity(2.0 - k)
            out.append(b)
        return w, out
    wg, wo, _, _ = both(gpu_api, oracle_api, build, 240)
    jg, n = wg.read_joints(); jo, _ = wo.read_joints()
    for i in range(n):
        assert jg[i].limitState == jo[i].limitState
        for a, b in zip(list(jg[i].impulse) + [jg[i].motorImpulse], list(jo[i].impulse) + [jo[i].motorImpulse]):
            assert abs(a - b) <= 2e-3 * max(1.0, abs(b)), (i, a, b)


def test_distance_joint_rigid_and_spring(gpu_api, oracle_api):
    """b2DistanceJoint (joints/b2distancejoint.d:113-330): rigid rod and soft spring (frequency / damping ratio)"""
    def build(api):
        w = b2World((0.0, -10.0), api=api)
        g = _ground(w, api)
        out = []
        for k, (hz, zeta) in enumerate(((0.0, 0.0), (2.0, 0.1), (4.0, 0.7))):
            b = _dyn(w, 8.0 * k - 8.0 + 1.0, 6.0)
            c = b2CircleShape(api); c.m_radius = 0.5
            b.CreateFixture(c, 1.0)
            jd = b2DistanceJointDef()
            jd.Initialize(g, b, (8.0 * k - 8.0, 10.0), (8.0 * k - 8.0 + 1.0, 6.0))
            jd.frequencyHz, jd.dampingRatio = hz, zeta
            w.CreateJoint(jd)
            out.append(b)
        return w, out
    both(gpu_api, oracle_api, build, 240)


def test_shapes_on_chain_ground(gpu_api, oracle_api):
    """circle / polygon against chain children with ghost vertices (b2chainshape.d:162-192, b2collideedge.d), circle-circle,
    polygon-circle; every body is its own island, so order cannot matter"""
    def build(api):
        w = b2World((0.0, -10.0), api=api)
        g = w.CreateBody(b2BodyDef())
        ch = b2ChainShape(api)
        ch.CreateChain([(-20.0, 2.0), (-12.0, 0.0), (-4.0, 0.5), (4.0, 0.0), (12.0, 1.0), (20.0, 0.0)])
        g.CreateFixture(ch, 0.0)
        out = []
        for k in range(8):
            x = -17.0 + 4.6 * k
            b = _dyn(w, x, 4.0 + 0.3 * k, angle=0.2 * k)
            if k % 3 == 0:
                s = b2CircleShape(api); s.m_radius = 0.4 + 0.05 * k
            elif k % 3 == 1:
                s = b2PolygonShape(api); s.SetAsBox(0.5, 0.3 + 0.05 * k)
            else:
                s = b2PolygonShape(api); s.Set([(-0.5, -0.4), (0.6, -0.3), (0.4, 0.5), (-0.2, 0.7), (-0.6, 0.2)])
            fd = b2FixtureDef(); fd.shape = s; fd.density = 1.0 + 0.2 * k; fd.friction = 0.1 * k; fd.restitution = 0.1 * (k % 4)
            b.CreateFixture(fd)
            out.append(b)
        # a circle resting on a circle that is pinned by being static
        pin = w.CreateBody(b2BodyDef()); pc = b2CircleShape(api); pc.m_radius = 1.0; pc.m_p = (0.0, 12.0); pin.CreateFixture(pc, 0.0)
        top = _dyn(w, 0.05, 14.0); tc = b2CircleShape(api); tc.m_radius = 0.5; top.CreateFixture(tc, 1.0)
        out.append(top)
        return w, out
    both(gpu_api, oracle_api, build, 120, tol_p=1e-4, tol_v=1e-3)     # beyond that a tumbling polygon amplifies 1-ulp sin/cos differences


def test_sensor_and_filters(gpu_api, oracle_api):
    """sensors overlap without response (b2contact.d:283-292), category/mask and group filtering
    (b2worldcallbacks.d:52-64): same contacts, same touching flags, same trajectories, same begin/end events"""
    def build(api):
        w = b2World((0.0, -10.0), api=api)
        g = _ground(w, api)
        zone = w.CreateBody(b2BodyDef())
        zs = b2PolygonShape(api); zs.SetAsBox(3.0, 1.0, (0.0, 3.0), 0.0)
        fd = b2FixtureDef(); fd.shape = zs; fd.isSensor = True
        zone.CreateFixture(fd)
        out = []
        for k in range(6):
            b = _dyn(w, -5.0 + 2.0 * k, 6.0 + 0.5 * k)
            s = b2PolygonShape(api); s.SetAsBox(0.4, 0.4)
            fd = b2FixtureDef(); fd.shape = s; fd.density = 1.0
            if k == 1:
                fd.filter.maskBits = 0x0000              # collides with nothing: falls through the ground
            if k in (2, 3):
                fd.filter.groupIndex = -7                # never collide with each other
            if k == 4:
                fd.filter.categoryBits = 0x0004; fd.filter.maskBits = 0xFFFF
            b.CreateFixture(fd)
            out.append(b)
        out[3].SetTransform((out[2].GetPosition().x + 0.2, 9.0), 0.0)     # k=3 lands on k=2's spot: same negative group
        w.EnableContactEvents(4096)
        return w, out

    def each(k, wg, wo):
        key = lambda e: (e[0], e[1], e[3], e[4])
        assert sorted(map(key, wg.PollContactEvents())) == sorted(map(key, wo.PollContactEvents())), k
        cg, co = wg.counts(), wo.counts()
        assert (cg.contacts, cg.touching) == (co.contacts, co.touching), k
    both(gpu_api, oracle_api, build, 150, each=each)


def test_bullet_does_not_tunnel(gpu_api, oracle_api):
    """continuous collision of a bullet against a thin dynamic wall and of a fast non-bullet against a static one
    (b2world.d:1127-1452): same TOI events, same outcome"""
    def build(api):
        w = b2World((0.0, 0.0), api=api)
        wall = w.CreateBody(b2BodyDef()); ws = b2PolygonShape(api); ws.SetAsBox(0.05, 5.0, (10.0, 0.0), 0.0); wall.CreateFixture(ws, 0.0)
        plate = _dyn(w, -10.0, 0.0); ps = b2PolygonShape(api); ps.SetAsBox(0.05, 3.0); plate.CreateFixture(ps, 20.0)
        fast = _dyn(w, 0.0, 1.0); fs = b2PolygonShape(api); fs.SetAsBox(0.1, 0.1); fast.CreateFixture(fs, 1.0)
        fast.SetLinearVelocity((300.0, 0.0))
        bullet = _dyn(w, 0.0, -1.0, bullet=True); bs = b2CircleShape(api); bs.m_radius = 0.1; bullet.CreateFixture(bs, 1.0)
        bullet.SetLinearVelocity((-400.0, 0.0))
        return w, [plate, fast, bullet]
    wg, wo, bg, bo = both(gpu_api, oracle_api, build, 30, tol_p=1e-4, tol_v=1e-3)
    assert bg[1].GetPosition().x < 10.0 and bg[2].GetPosition().x > -10.5     # neither went through


def test_kinematic_platform_and_destroy_calls(gpu_api, oracle_api):
    """kinematic body driving a box (b2island.d integrates it, zero inverse mass), then DestroyJoint / DestroyFixture /
    DestroyBody between steps (b2world.d:105-359, b2body.d:172-254): wake-ups and contact removal as in the reference"""
    def build(api):
        w = b2World((0.0, -10.0), api=api)
        g = _ground(w, api)
        bd = b2BodyDef(); bd.type = b2_kinematicBody; bd.position.Set(0.0, 2.0)
        plat = w.CreateBody(bd); s = b2PolygonShape(api); s.SetAsBox(3.0, 0.25); plat.CreateFixture(s, 0.0)
        plat.SetLinearVelocity((1.0, 0.0))
        rider = _dyn(w, 0.0, 3.0); rs = b2PolygonShape(api); rs.SetAsBox(0.5, 0.5)
        fd = b2FixtureDef(); fd.shape = rs; fd.density = 1.0; fd.friction = 0.8
        rider.CreateFixture(fd)
        hang = _dyn(w, 8.0, 6.0); hs = b2CircleShape(api); hs.m_radius = 0.5; hang.CreateFixture(hs, 1.0)
        jd = b2DistanceJointDef(); jd.Initialize(g, hang, (8.0, 10.0), (8.0, 6.0))
        w.joint = w.CreateJoint(jd)
        two = _dyn(w, -8.0, 1.0); a = b2PolygonShape(api); a.SetAsBox(0.5, 0.5); two.CreateFixture(a, 1.0)
        b2 = b2PolygonShape(api); b2.SetAsBox(0.3, 0.3, (0.0, 0.8), 0.0)     # rides on top: one ground contact only
        w.extra = two.CreateFixture(b2, 1.0)
        w.plat, w.two = plat, two
        return w, [plat, rider, hang, two]

    def each(k, wg, wo):
        for w in (wg, wo):
            if k == 60:
                w.DestroyJoint(w.joint)
            if k == 90:
                w.two.DestroyFixture(w.extra)
            if k == 120:
                w.DestroyBody(w.plat)
    wg, wo = None, None

    def build2(api):
        w, out = build(api)
        return w, out[1:]                                       # the platform gets destroyed: compare the others
    both(gpu_api, oracle_api, build2, 180, each=each, tol_p=1e-4, tol_v=1e-3)


def _box_body(w, api, x, y, hx=0.5, hy=0.5, density=1.0, **kw):
    b = _dyn(w, x, y, **kw)
    s = b2PolygonShape(api); s.SetAsBox(hx, hy)
    b.CreateFixture(s, density)
    return b


def test_second_wave_joints_one_each(gpu_api, oracle_api):
    """prismatic (limit + motor), weld (rigid and soft), wheel (spring + motor), rope, friction, motor,
    mouse (with SetTarget) and pulley, one joint per island so the Gauss-Seidel order cannot matter: the CUDA hooks must
    track the oracle's restatement of the reference hooks to float rounding, impulses and limit states included"""
    from dbox_b200.world import (b2PrismaticJointDef, b2WeldJointDef, b2WheelJointDef, b2RopeJointDef, b2FrictionJointDef,
                                 b2MotorJointDef, b2MouseJointDef, b2PulleyJointDef, b2Vec2)

    def build(api):
        w = b2World((0.0, -10.0), api=api)
        g = w.CreateBody(b2BodyDef())
        out = []
        # prismatic: slanted axis, limits, motor pushing against the upper limit
        a = _box_body(w, api, 0.0, 5.0, angle=0.1)
        jd = b2PrismaticJointDef(); jd.Initialize(g, a, (0.0, 5.0), (0.6, 0.8))
        jd.enableLimit, jd.lowerTranslation, jd.upperTranslation = True, -2.0, 1.0
        jd.enableMotor, jd.motorSpeed, jd.maxMotorForce = True, 2.0, 60.0
        w.CreateJoint(jd); out.append(a)
        # prismatic without limit or motor
        a2 = _box_body(w, api, 3.0, 5.0)
        jd = b2PrismaticJointDef(); jd.Initialize(g, a2, (3.0, 6.0), (1.0, 0.3)); w.CreateJoint(jd); out.append(a2)
        # weld rigid (off-centre anchor) and soft
        b = _box_body(w, api, 6.0, 5.0)
        jd = b2WeldJointDef(); jd.Initialize(g, b, (7.5, 6.0)); w.CreateJoint(jd); out.append(b)
        b2 = _box_body(w, api, 10.0, 5.0)
        jd = b2WeldJointDef(); jd.Initialize(g, b2, (11.5, 5.0)); jd.frequencyHz, jd.dampingRatio = 3.0, 0.3; w.CreateJoint(jd); out.append(b2)
        # wheel: suspension spring + motor
        c = _dyn(w, 14.0, 5.0); cs = b2CircleShape(api); cs.m_radius = 0.5; c.CreateFixture(cs, 1.0)
        jd = b2WheelJointDef(); jd.Initialize(g, c, (14.0, 5.0), (0.0, 1.0))
        jd.enableMotor, jd.motorSpeed, jd.maxMotorTorque, jd.frequencyHz, jd.dampingRatio = True, -3.0, 5.0, 4.0, 0.5
        w.CreateJoint(jd); out.append(c)
        # wheel without spring
        c2 = _box_body(w, api, 17.0, 5.0)
        jd = b2WheelJointDef(); jd.Initialize(g, c2, (17.0, 5.0), (0.3, 1.0)); jd.frequencyHz = 0.0; w.CreateJoint(jd); out.append(c2)
        # rope: slack at first, then taut
        d = _box_body(w, api, 20.0, 7.0)
        jd = b2RopeJointDef(); jd.bodyA, jd.bodyB = g, d
        jd.localAnchorA, jd.localAnchorB, jd.maxLength = b2Vec2(21.0, 8.0), b2Vec2(0.5, 0.5), 3.0
        w.CreateJoint(jd); d.SetLinearVelocity((2.0, 0.0)); out.append(d)
        # friction: saturated
        e = _box_body(w, api, 24.0, 5.0)
        jd = b2FrictionJointDef(); jd.Initialize(g, e, (24.0, 5.0)); jd.maxForce, jd.maxTorque = 6.0, 0.5
        w.CreateJoint(jd); e.SetAngularVelocity(3.0); out.append(e)
        # motor: holds an offset pose against gravity with a soft correction
        f = _box_body(w, api, 28.0, 5.0)
        jd = b2MotorJointDef(); jd.Initialize(g, f); jd.maxForce, jd.maxTorque = 40.0, 20.0
        jd.linearOffset = b2Vec2(28.5, 5.5); jd.angularOffset = 0.4
        w.CreateJoint(jd); out.append(f)
        # mouse
        m = _box_body(w, api, 32.0, 5.0)
        jd = b2MouseJointDef(); jd.bodyA, jd.bodyB = g, m; jd.target = b2Vec2(32.3, 5.4); jd.maxForce = 200.0
        w.mouse = w.CreateJoint(jd); out.append(m)
        # pulley with ratio
        p1 = _box_body(w, api, 36.0, 5.0); p2 = _box_body(w, api, 39.0, 5.0, density=1.6)
        jd = b2PulleyJointDef(); jd.Initialize(p1, p2, (36.0, 10.0), (39.0, 10.0), (36.0, 5.5), (39.0, 5.5), 1.5)
        w.CreateJoint(jd); out += [p1, p2]
        return w, out

    def each(k, wg, wo):
        if k == 60:
            for w in (wg, wo):
                w.SetMouseTarget(w.mouse, (33.0, 6.5))
    wg, wo, _, _ = both(gpu_api, oracle_api, build, 200, each=each, tol_p=1e-4, tol_v=1e-3)
    jg, n = wg.read_joints(); jo, _ = wo.read_joints()
    assert n == 11
    for i in range(n):
        assert jg[i].type == jo[i].type and jg[i].limitState == jo[i].limitState, i
        for a, b in zip(list(jg[i].impulse) + [jg[i].motorImpulse], list(jo[i].impulse) + [jo[i].motorImpulse]):
            assert abs(a - b) <= 5e-3 * max(1.0, abs(b)), (i, jg[i].type, a, b)


def test_joint_scene_with_contacts(gpu_api, oracle_api):
    """a small car (two wheel joints with springs, motor on the rear wheel) on a chain-shape track, a weld-joint beam and a
    prismatic elevator carrying a box: joints and contacts in the same islands.  Order now matters, so this is judged like
    the long-horizon scenes: same outcome within a loose tolerance"""
    from dbox_b200.world import b2PrismaticJointDef, b2WeldJointDef, b2WheelJointDef

    def build(api):
        w = b2World((0.0, -10.0), api=api)
        g = w.CreateBody(b2BodyDef())
        ch = b2ChainShape(api); ch.CreateChain([(-20.0, 0.0), (20.0, 0.0), (40.0, 2.0), (60.0, 2.0)])
        fd = b2FixtureDef(); fd.shape = ch; fd.friction = 0.9
        g.CreateFixture(fd)
        chassis = _box_body(w, api, 0.0, 1.0, hx=1.5, hy=0.25)
        wheels = []
        for x in (-1.0, 1.0):
            wb = _dyn(w, x, 0.45); cs = b2CircleShape(api); cs.m_radius = 0.4
            fd = b2FixtureDef(); fd.shape = cs; fd.density = 1.0; fd.friction = 0.9
            wb.CreateFixture(fd)
            jd = b2WheelJointDef(); jd.Initialize(chassis, wb, (x, 0.45), (0.0, 1.0))
            jd.frequencyHz, jd.dampingRatio = 4.0, 0.7
            if x < 0:
                jd.enableMotor, jd.motorSpeed, jd.maxMotorTorque = True, -8.0, 20.0
            w.CreateJoint(jd); wheels.append(wb)
        lift = _box_body(w, api, -10.0, 1.0, hx=1.0, hy=0.2)
        jd = b2PrismaticJointDef(); jd.Initialize(g, lift, (-10.0, 1.0), (0.0, 1.0))
        jd.enableLimit, jd.lowerTranslation, jd.upperTranslation = True, 0.0, 4.0
        jd.enableMotor, jd.motorSpeed, jd.maxMotorForce = True, 1.0, 500.0
        w.CreateJoint(jd)
        cargo = _box_body(w, api, -10.0, 1.7)
        return w, [chassis, lift, cargo] + wheels
    wg, bg = build(gpu_api); wo, bo = build(oracle_api)
    for k in range(240):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
    for i, (g_, o_) in enumerate(zip(bg, bo)):
        pg, po = g_.GetPosition(), o_.GetPosition()
        assert abs(pg.x - po.x) < 0.05 and abs(pg.y - po.y) < 0.05, (i, (pg.x, pg.y), (po.x, po.y))
    assert bg[0].GetPosition().x > 5.0                       # the car drove off
    assert abs(bg[1].GetPosition().y - 5.0) < 0.05           # the lift reached its upper limit with the cargo on it
    assert bg[2].GetPosition().y > 5.5


def _query_scene(api):
    w = b2World((0.0, -10.0), api=api)
    g = w.CreateBody(b2BodyDef())
    ch = b2ChainShape(api); ch.CreateChain([(-30.0, 0.0), (-10.0, 1.0), (10.0, 0.0), (30.0, 2.0)])
    g.CreateFixture(ch, 0.0)
    e = b2EdgeShape(api); e.Set((-30.0, 0.0), (-30.0, 20.0)); g.CreateFixture(e, 0.0)
    import random
    rng = random.Random(5)
    bodies = []
    for k in range(60):
        b = _dyn(w, rng.uniform(-25, 25), rng.uniform(2, 18), angle=rng.uniform(-3, 3))
        t = k % 3
        if t == 0:
            s = b2CircleShape(api); s.m_radius = rng.uniform(0.3, 1.0)
        elif t == 1:
            s = b2PolygonShape(api); s.SetAsBox(rng.uniform(0.3, 1.2), rng.uniform(0.3, 1.2))
        else:
            s = b2PolygonShape(api); s.Set([(-0.7, -0.5), (0.8, -0.4), (0.5, 0.6), (-0.3, 0.9), (-0.8, 0.2)])
        b.CreateFixture(s, 1.0)
        bodies.append(b)
    return w, bodies


def test_raycast_and_query_match_reference(gpu_api, oracle_api):
    """b2World.RayCast (closest-hit callback) and b2World.QueryAABB (b2world.d:563-587) on the LBVH against the oracle's
    dynamic tree + b2Shape.RayCast restatements: same fixture / child, fraction and normal to float rounding, same
    fat-box overlap sets; before the first step, after stepping, and after an edit without a step"""
    import random
    wg, bg = _query_scene(gpu_api); wo, bo = _query_scene(oracle_api)
    rng = random.Random(11)
    rays = [((rng.uniform(-35, 35), rng.uniform(-2, 25)), (rng.uniform(-35, 35), rng.uniform(-2, 25))) for _ in range(400)]
    rays += [((0.0, 30.0), (0.0, -5.0)), ((-40.0, 5.0), (40.0, 5.0)), ((3.0, 3.0), (3.0, 3.0))]      # incl. a zero-length ray
    boxes = [((x, y), (x + rng.uniform(0.1, 8), y + rng.uniform(0.1, 8))) for x, y in
             [(rng.uniform(-32, 28), rng.uniform(-2, 20)) for _ in range(120)]] + [((-100.0, -100.0), (100.0, 100.0))]

    def compare(tag, tol_f, tol_p, exact):
        hg, ho = wg.RayCastClosest(rays), wo.RayCastClosest(rays)
        hits = other = 0
        for k, (a, b) in enumerate(zip(hg, ho)):
            if (a[0], a[1]) != (b[0], b[1]):
                assert not exact, (tag, k, a, b)
                other += 1
                continue
            if b[0] >= 0:
                hits += 1
                close = abs(a[2] - b[2]) <= tol_f * max(1.0, abs(b[2])) and abs(a[3][0] - b[3][0]) <= tol_p and abs(a[3][1] - b[3][1]) <= tol_p
                if exact:
                    assert close and abs(a[4][0] - b[4][0]) <= tol_p and abs(a[4][1] - b[4][1]) <= tol_p, (tag, k, a, b)
                elif not close:
                    other += 1
        assert hits > 100 and other <= 12, (tag, hits, other)          # 3 % of the rays may see a body that tumbled differently
        qg, qo = wg.QueryAABB(boxes, cap=128), wo.QueryAABB(boxes, cap=128)
        if exact:
            assert qg == qo, tag
        else:
            assert sum(1 for x, y in zip(qg, qo) if x != y) <= 3, tag
        assert len(qg[-1]) == 60 + 3 + 1

    compare("before the first step", 1e-5, 2e-4, True)       # identical worlds: identical answers
    for _ in range(40):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
    compare("after 40 steps", 5e-3, 5e-2, False)              # chaotic scene: the two worlds now differ by the Gauss-Seidel order
    for w, bs in ((wg, bg), (wo, bo)):
        bs[7].SetTransform((0.0, 25.0), 0.3)
        w.DestroyBody(bs[9])
    hg, ho = wg.RayCastClosest([((0.0, 30.0), (0.0, 20.0))]), wo.RayCastClosest([((0.0, 30.0), (0.0, 20.0))])
    assert hg[0][0] == ho[0][0] == bg[7].fixtures[0].id


@pytest.mark.parametrize("continuous", [False, True])
def test_pre_solve_split_step(gpu_api, oracle_api, continuous):
    """b2ContactListener.PreSolve (b2contact.d:348-355) through the step cut after Collide: a one-way platform
    (SetEnabled(false) while the box comes from below), a conveyor (SetTangentSpeed) and a per-contact friction override;
    and the split step without patches is bit-identical to the plain step.  With continuous physics on, the box shot at the
    platform becomes a TOI event whose b2Contact.Update would ask the listener again (b2world.d:1295): the answer given after
    Collide stands for the rest of the step, so the platform stays open for it."""
    def build(api):
        w = b2World((0.0, -10.0), api=api)
        w.SetContinuousPhysics(continuous)
        g = _ground(w, api)
        plat = w.CreateBody(b2BodyDef()); ps = b2PolygonShape(api); ps.SetAsBox(3.0, 0.25, (0.0, 6.0), 0.0)
        w.platform_fixture = plat.CreateFixture(ps, 0.0).id
        up = _box_body(w, api, 0.0, 3.0); up.SetLinearVelocity((0.0, 14.0))          # shot upwards through the platform
        belt = w.CreateBody(b2BodyDef()); bs = b2PolygonShape(api); bs.SetAsBox(4.0, 0.25, (12.0, 2.0), 0.0)
        w.belt_fixture = belt.CreateFixture(bs, 0.0).id
        rider = _box_body(w, api, 10.0, 2.8)
        ice = _box_body(w, api, -12.0, 0.55); ice.SetLinearVelocity((6.0, 0.0))       # friction overridden to 0 against the ground
        w.ice_fixture = ice.fixtures[0].id
        w.up = up
        return w, [up, rider, ice]

    def pre_solve_for(w):
        def pre_solve(c):
            fx = (c.fixtureA, c.fixtureB)
            if w.platform_fixture in fx:
                return {"enabled": w.up.GetLinearVelocity().y <= 0.0}                # solid only for a box coming down
            if w.belt_fixture in fx:
                return {"tangentSpeed": 2.0 if c.fixtureA == w.belt_fixture else -2.0}
            if w.ice_fixture in fx:
                return {"friction": 0.0}
            return None
        return pre_solve
    wg, bg = build(gpu_api); wo, bo = build(oracle_api)
    fg, fo = pre_solve_for(wg), pre_solve_for(wo)
    for k in range(150):
        # (continuous: the box reaches the platform inside the TOI loop, so the listener is asked ahead of the first touch)
        wg.StepWithPreSolve(DT, 8, 3, fg, toi_lookahead=continuous); wo.StepWithPreSolve(DT, 8, 3, fo, toi_lookahead=continuous)
        for i, (a, b) in enumerate(zip(bg, bo)):
            pa, pb = a.GetPosition(), b.GetPosition()
            assert abs(pa.x - pb.x) < 2e-4 * max(1.0, abs(pb.x)) and abs(pa.y - pb.y) < 2e-4 * max(1.0, abs(pb.y)), (k, i, (pa.x, pa.y), (pb.x, pb.y))
    up, rider, ice = bg
    assert 6.7 < up.GetPosition().y < 6.85                    # went up through the platform, landed on top of it
    assert rider.GetPosition().x > 10.5                       # carried by the conveyor
    assert ice.GetLinearVelocity().x > 5.9                    # no friction against the ground
    # no patches: same bits as dbx_world_step
    a, _ = scenes.pyramid(api=gpu_api, count=8); b, _ = scenes.pyramid(api=gpu_api, count=8)
    for k in range(60):
        a.Step(DT, 8, 3)
        b.StepWithPreSolve(DT, 8, 3, lambda c: None)
    sa, n = a.read_bodies(); sb, _ = b.read_bodies()
    for i in range(n):
        assert (sa[i].c.x, sa[i].c.y, sa[i].a, sa[i].v.x, sa[i].v.y, sa[i].w) == (sb[i].c.x, sb[i].c.y, sb[i].a, sb[i].v.x, sb[i].v.y, sb[i].w), i


def test_gear_joint(gpu_api, oracle_api):
    """b2GearJoint (b2gearjoint.d): revolute-revolute with a ratio and revolute-prismatic, the Gears demo layout.  The gear
    shares its island with joint1 and joint2, so the order of the three joints matters: loose tolerance against the oracle,
    tight tolerance on the constraint itself (coordinateA + ratio * coordinateB stays what it was at creation)"""
    from dbox_b200.world import b2GearJointDef, b2PrismaticJointDef

    def build(api):
        w = b2World((0.0, -10.0), api=api)
        g = w.CreateBody(b2BodyDef())
        b1 = _dyn(w, 0.0, 12.0); c1 = b2CircleShape(api); c1.m_radius = 1.0; b1.CreateFixture(c1, 5.0)
        jd = b2RevoluteJointDef(); jd.Initialize(g, b1, (0.0, 12.0)); j1 = w.CreateJoint(jd)
        b2 = _dyn(w, 3.0, 12.0); c2 = b2CircleShape(api); c2.m_radius = 2.0; b2.CreateFixture(c2, 5.0)
        jd = b2RevoluteJointDef(); jd.Initialize(g, b2, (3.0, 12.0)); j2 = w.CreateJoint(jd)
        b3 = _box_body(w, api, 5.5, 12.0, hx=0.5, hy=5.0, density=5.0)
        jd = b2PrismaticJointDef(); jd.Initialize(g, b3, (5.5, 12.0), (0.0, 1.0))
        jd.enableLimit, jd.lowerTranslation, jd.upperTranslation = True, -5.0, 5.0
        j3 = w.CreateJoint(jd)
        gd = b2GearJointDef(); gd.joint1, gd.joint2, gd.ratio = j1, j2, 2.0; w.gear1 = w.CreateJoint(gd)
        gd = b2GearJointDef(); gd.joint1, gd.joint2, gd.ratio = j2, j3, -1.0 / 2.0; w.gear2 = w.CreateJoint(gd)
        b1.SetAngularVelocity(1.0)
        return w, [b1, b2, b3]
    wg, bg = build(gpu_api); wo, bo = build(oracle_api)
    for k in range(180):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
        for i, (a, b) in enumerate(zip(bg, bo)):
            assert abs(a.GetAngle() - b.GetAngle()) < 5e-3 and abs(a.GetPosition().y - b.GetPosition().y) < 5e-3, (k, i, a.GetAngle(), b.GetAngle())
        a1, a2, y3 = bg[0].GetAngle(), bg[1].GetAngle(), bg[2].GetPosition().y - 12.0
        assert abs(a1 + 2.0 * a2) < 2e-2 and abs(a2 - 0.5 * y3) < 2e-2, (k, a1, a2, y3)
    assert abs(bg[2].GetPosition().y - 12.0) > 0.5                  # the weight of the bar drives the whole train
    jg, n = wg.read_joints(); jo, _ = wo.read_joints()
    assert jg[3].type == jo[3].type == A.JOINT_GEAR
    for i in (3, 4):
        assert abs(jg[i].impulse[0] - jo[i].impulse[0]) < 0.05 * max(1.0, abs(jo[i].impulse[0])), (i, jg[i].impulse[0], jo[i].impulse[0])


def test_edge_cases(gpu_api, oracle_api):
    """empty and degenerate worlds, error codes instead of silent clamps, a body with more than 64 contacts (overflow colour
    lanes), zero time step, a step with nothing awake"""
    # empty world, world with only static bodies, world whose only body has no fixture
    w = b2World((0.0, -10.0), api=gpu_api)
    w.Step(DT, 8, 3)
    c = w.counts(); assert (c.bodies, c.contacts, c.proxies) == (0, 0, 0)
    assert w.RayCastClosest([((0.0, 0.0), (1.0, 1.0))])[0][0] == -1 and w.QueryAABB([((-1.0, -1.0), (1.0, 1.0))]) == [[]]
    g = _ground(w, gpu_api)
    lone = _dyn(w, 0.0, 5.0)                       # no fixture: zero mass, falls under gravity like the reference (mass defaults to 1)
    wo = b2World((0.0, -10.0), api=oracle_api); _ground(wo, oracle_api); lone_o = _dyn(wo, 0.0, 5.0)
    for _ in range(30):
        w.Step(DT, 8, 3); wo.Step(DT, 8, 3)
    assert abs(lone.GetPosition().y - lone_o.GetPosition().y) < 1e-5
    # dt = 0 (b2world.d:388-396: inv_dt = 0, nothing integrates, contacts still update)
    before = lone.GetPosition().y
    w.Step(0.0, 8, 3)
    assert lone.GetPosition().y == before
    # invalid handles are errors, not crashes
    assert gpu_api.body_destroy(w._w, 12345) < 0 and gpu_api.fixture_destroy(w._w, -3) < 0 and gpu_api.joint_destroy(w._w, 7) < 0
    jd = A.JointDef(); gpu_api.default_joint_def(jd, A.JOINT_REVOLUTE); jd.bodyA, jd.bodyB = 0, 0
    assert gpu_api.joint_create(w._w, jd) < 0                                      # same body twice
    jd.type = 99
    assert gpu_api.joint_create(w._w, jd) < 0
    # one dynamic body touched by far more than 64 contacts: the colouring spills into that body's private lanes
    def heavy(api):
        w2 = b2World((0.0, -10.0), api=api)
        _ground(w2, api)
        slab = _box_body(w2, api, 0.0, 1.0, hx=30.0, hy=0.5, density=0.2)
        small = []
        for k in range(100):
            b = _dyn(w2, -29.0 + 0.58 * k, 1.8); s = b2PolygonShape(api); s.SetAsBox(0.25, 0.25); b.CreateFixture(s, 1.0)
            small.append(b)
        return w2, [slab] + small
    wg2, bg2 = heavy(gpu_api); wo2, bo2 = heavy(oracle_api)
    most = 0
    for k in range(120):
        wg2.Step(DT, 8, 3); wo2.Step(DT, 8, 3)
        if k % 10 == 9:
            most = max(most, wg2.counts().colours)
            assert gpu_api.world_debug_colour_conflicts(wg2._w) == 0
    cg, co = wg2.counts(), wo2.counts()
    assert cg.touching == co.touching and most > 64, most
    for a, b in zip(bg2, bo2):
        assert abs(a.GetPosition().y - b.GetPosition().y) < 0.02, (a.id, a.GetPosition().y, b.GetPosition().y)
    # a pool that is too small says so
    caps = A.Caps(); caps.maxContacts = 1
    tiny, _ = scenes.pyramid(api=gpu_api, count=3)
    assert tiny.counts().bodies == 8


def test_everything_asleep_costs_nothing_wrong(gpu_api, oracle_api):
    """a world that has gone to sleep keeps stepping without touching anything, and wakes up on contact exactly like the
    reference (b2world.d:963-1118, b2island.d:249-279)"""
    wg, bg = scenes.pyramid(api=gpu_api, count=6); wo, bo = scenes.pyramid(api=oracle_api, count=6)
    for _ in range(400):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
    assert wg.counts().awakeBodies == wo.counts().awakeBodies == 0
    frozen = [(b.GetPosition().x, b.GetPosition().y, b.GetAngle()) for b in bg]
    for _ in range(50):
        wg.Step(DT, 8, 3)
    assert frozen == [(b.GetPosition().x, b.GetPosition().y, b.GetAngle()) for b in bg]
    for w, api in ((wg, gpu_api), (wo, oracle_api)):                     # drop a box on the sleeping pile
        bd = b2BodyDef(); bd.type = b2_dynamicBody; bd.position.Set(-5.0, 12.0)
        nb = w.CreateBody(bd); s = b2PolygonShape(api); s.SetAsBox(0.5, 0.5); nb.CreateFixture(s, 5.0)
    woke = False
    for k in range(150):
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
        cg, co = wg.counts(), wo.counts()
        woke = woke or cg.awakeBodies > 1
        if k < 60:
            assert cg.awakeBodies == co.awakeBodies, (k, cg.awakeBodies, co.awakeBodies)
    assert woke


def test_islands_sleep_and_wake_one_by_one_like_the_reference(gpu_api, oracle_api):
    """islands are found per step and each keeps its own sleep clock (b2island.d:241-279: the
    minimum sleep time over ITS bodies, asleep together when it passes b2_timeToSleep and the position solver has converged).
    Five separate stacks, nudged at different times: the island count and the awake flag of every body must follow the oracle
    step by step (two steps of slack where a sleep clock crosses the threshold), a stack sleeps and wakes as one, and some stacks
    are asleep while others are awake."""
    def build(api):
        w = b2World((0.0, -10.0), api=api)
        _ground(w, api)
        return w, [[_box_body(w, api, -16.0 + 8.0 * k, 0.51 + 1.01 * r) for r in range(3)] for k in range(5)]
    wg, sg = build(gpu_api); wo, so = build(oracle_api)
    nudges = {60: (2, 2, (0.3, 0.0)), 150: (4, 1, (0.0, 0.6)), 260: (0, 0, (0.2, 0.0)), 262: (1, 2, (-0.2, 0.0))}
    mismatch, partial, prev_awake = {}, 0, None
    for k in range(420):
        if k in nudges:
            st, r, imp = nudges[k]
            for stacks in (sg, so):
                p = stacks[st][r].GetPosition()
                stacks[st][r].ApplyLinearImpulse(imp, (p.x, p.y + 0.2))
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
        cg, co = wg.counts(), wo.counts()
        ag = [[b.IsAwake() for b in st] for st in sg]
        ao = [[b.IsAwake() for b in st] for st in so]
        for st in ag:
            assert len(set(st)) == 1, (k, ag)                # the three boxes of a stack are one island
        if ag == ao:
            assert cg.islands == co.islands, (k, cg.islands, co.islands, ao)        # (islands SOLVED in the step: a stack that just fell asleep counts)
            assert prev_awake is None or cg.islands >= sum(1 for st in ao if st[0])
            assert cg.awakeBodies == co.awakeBodies
        else:
            for i, (x, y) in enumerate(zip(ag, ao)):
                if x != y:
                    mismatch[i] = mismatch.get(i, 0) + 1
        partial += 0 < sum(1 for st in ao if st[0]) < 5
        prev_awake = ao
    assert partial > 100                                      # the clocks are per island: long stretches with some asleep, some awake
    assert all(v <= 2 for v in mismatch.values()), mismatch  # a sleep clock crossing 0.5 s a step or two apart, nothing more
    assert not any(x for st in ag for x in st)
    wg.close(); wo.close()


def test_set_type_and_set_active(gpu_api, oracle_api):
    """b2Body.SetType (b2body.d:867-914: mass reset, contacts destroyed, proxies touched) and SetActive (:718-775: proxies
    destroyed / re-created in fixture-list order, contacts destroyed) between steps, against the oracle: same contact and
    proxy populations, same A/B order of the contacts that come back, same trajectories"""
    from dbox_b200.world import b2_staticBody

    def build(api):
        w = b2World((0.0, -10.0), api=api)
        _ground(w, api)
        out = [_box_body(w, api, -6.0 + 3.0 * k, 0.52 + 0.0 * k) for k in range(5)]       # five boxes resting on the ground
        tops = [_box_body(w, api, -6.0 + 3.0 * k, 1.55) for k in range(5)]                # one box on each
        return w, out + tops
    wg, bg = build(gpu_api); wo, bo = build(oracle_api)

    def edit(k):
        for w, bs in ((wg, bg), (wo, bo)):
            if k == 40:
                bs[0].SetType(b2_staticBody)            # a resting box becomes part of the scenery
                bs[6].SetActive(False)                  # the box on top of #1 disappears from the simulation
            if k == 70:
                bs[0].SetType(b2_dynamicBody)
                bs[6].SetTransform((-3.0, 4.0), 0.3)
                bs[6].SetActive(True)                   # and comes back higher up
                bs[2].SetType(b2_kinematicBody); bs[2].SetLinearVelocity((0.5, 0.0))
    for k in range(160):
        edit(k)
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
        cg, co = wg.counts(), wo.counts()
        assert (cg.contacts, cg.touching, cg.proxies) == (co.contacts, co.touching, co.proxies), (k, cg.contacts, co.contacts, cg.touching, co.touching, cg.proxies, co.proxies)
        for i, (a, b) in enumerate(zip(bg, bo)):
            if i == 6 and 40 <= k < 70:
                continue                                # inactive: frozen on both sides
            pa, pb = a.GetPosition(), b.GetPosition()
            tol = 1e-3 if k < 100 else 3e-2          # after that the tilted box lands on a stack: Gauss-Seidel order shows
            assert abs(pa.x - pb.x) < tol and abs(pa.y - pb.y) < tol, (k, i, (pa.x, pa.y), (pb.x, pb.y))
    from tests.parity import contacts_by_key
    kg, _, _ = contacts_by_key(wg); ko, _, _ = contacts_by_key(wo)
    assert set(kg) == set(ko)                                               # same (fixtureA, childA, fixtureB, childB) keys: same A/B order
    assert bg[2].GetPosition().x > 0.5                                      # the kinematic box carried its passenger along


def test_body_and_fixture_mutators(gpu_api, oracle_api):
    """b2Fixture.SetFilterData / SetSensor / SetFriction / SetRestitution / SetDensity (b2fixture.d:108-262) and b2Body.SetMassData /
    ResetMassData / SetFixedRotation / Set*Damping / SetGravityScale (b2body.d:502-690, 924-945) between steps"""
    def build(api):
        w = b2World((0.0, -10.0), api=api)
        _ground(w, api)
        bs = [_box_body(w, api, -8.0 + 4.0 * k, 0.52) for k in range(5)]
        tops = [_box_body(w, api, -8.0 + 4.0 * k, 1.55) for k in range(5)]
        return w, bs + tops

    def edit(k, bs):
        if k == 30:
            bs[5].fixtures[0].SetFilterData(0x0002, 0x0000, 0)       # top box 0 stops colliding: falls through everything
            bs[1].fixtures[0].SetSensor(True)                         # bottom box 1 becomes a sensor: falls through the ground, its passenger lands
            bs[7].SetGravityScale(-0.5); bs[7].SetLinearDamping(0.4)  # top box 2 floats away, damped
            bs[8].SetMassData(3.0, (0.2, 0.0), 2.0)                   # top box 3 gets an off-centre mass
            bs[9].SetFixedRotation(True); bs[9].SetAngularDamping(0.2)
            bs[4].fixtures[0].SetFriction(0.0); bs[4].fixtures[0].SetRestitution(0.5)
        if k == 60:
            bs[5].SetTransform((-8.0, 4.0), 0.0); bs[5].SetLinearVelocity((0.0, 0.0))
            bs[5].fixtures[0].SetFilterData()                         # collides again: lands back on its box
            bs[3].fixtures[0].SetDensity(4.0); bs[3].ResetMassData()
            bs[3].ApplyLinearImpulse((6.0, 0.0), (bs[3].GetPosition().x, bs[3].GetPosition().y + 0.4))
    wg, bg = build(gpu_api); wo, bo = build(oracle_api)
    for k in range(120):
        edit(k, bg); edit(k, bo)
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
        cg, co = wg.counts(), wo.counts()
        # touching contacts exactly; the AABB-level cache may drop a pair of a fast-falling body one step apart
        assert cg.touching == co.touching and abs(cg.contacts - co.contacts) <= 2, (k, cg.contacts, co.contacts, cg.touching, co.touching)
        for i, (a, b) in enumerate(zip(bg, bo)):
            pa, pb = a.GetPosition(), b.GetPosition()
            tol = 2e-3 * max(1.0, abs(pb.x), abs(pb.y))
            assert abs(pa.x - pb.x) < tol and abs(pa.y - pb.y) < tol and abs(a.GetAngle() - b.GetAngle()) < 5e-3, (k, i, (pa.x, pa.y), (pb.x, pb.y))
    assert 1.4 < bg[5].GetPosition().y < 1.7 and bg[1].GetPosition().y < 0.0 and bg[7].GetPosition().y > 2.0
    assert abs(bg[8].GetMass() - 3.0) < 1e-6 and abs(bg[3].GetMass() - 4.0) < 1e-5


def test_per_frame_body_calls_touch_one_row(gpu_api, oracle_api):
    """b2Body.ApplyForce / ApplyTorque / ApplyLinearImpulse / SetLinearVelocity / SetAngularVelocity / SetAwake / SetBullet /
    SetSleepingAllowed (b2body.d:287-500, 827-922) between steps, the way a game drives a few bodies every frame: the same
    trajectory as the reference, and each call moves ONE body's row between host and device -- on a 20,000-body pile a frame
    with such calls must cost about what a frame without them does (it used to copy every body array both ways)."""
    import time

    def build(api):
        w = b2World((0.0, -10.0), api=api)
        _ground(w, api)
        bs = [_box_body(w, api, -8.0 + 4.0 * k, 0.52) for k in range(5)] + [_box_body(w, api, -8.0 + 4.0 * k, 1.55) for k in range(5)]
        return w, bs

    def drive(k, bs):
        # (gentle pushes: with 8 iterations a violent shove leaves the device's coloured order and the reference's list order
        # several per cent apart, which is the solver's convergence and not what this test is about)
        bs[0].ApplyForce((2.0, 0.0), (bs[0].GetPosition().x, bs[0].GetPosition().y + 0.05))
        bs[6].ApplyTorque(0.2)
        if k % 20 == 5:
            bs[7].ApplyLinearImpulse((0.0, 1.5), (bs[7].GetPosition().x + 0.05, bs[7].GetPosition().y))
            bs[8].ApplyAngularImpulse(0.05)
        if k == 30:
            bs[2].SetLinearVelocity((1.5, 0.0)); bs[2].SetAngularVelocity(0.5)
            bs[3].SetBullet(True); bs[4].SetSleepingAllowed(False)
        if k == 90:
            bs[9].SetAwake(False)
        if k == 100:
            bs[9].SetAwake(True); bs[9].SetLinearVelocity((0.0, 2.0))
    wg, bg = build(gpu_api); wo, bo = build(oracle_api)
    for k in range(140):
        drive(k, bg); drive(k, bo)
        wg.Step(DT, 8, 3); wo.Step(DT, 8, 3)
        for i, (a, b) in enumerate(zip(bg, bo)):
            pa, pb = a.GetPosition(), b.GetPosition()
            tol = 2e-3 * max(1.0, abs(pb.x), abs(pb.y))
            assert abs(pa.x - pb.x) < tol and abs(pa.y - pb.y) < tol and abs(a.GetAngle() - b.GetAngle()) < 5e-3, (k, i, (pa.x, pa.y), (pb.x, pb.y))
            assert a.IsAwake() == b.IsAwake(), (k, i)
    wg.close(); wo.close()
    # cost: frames with per-body calls against frames without, same world
    w, bodies, _ = scenes.pile(api=gpu_api, n=20000, columns=400)
    w.SetAllowSleeping(False)
    w.StepN(DT, 8, 3, 60)

    def frames(n, calls):
        w.Step(DT, 8, 3); bodies[0].GetPosition()
        t0 = time.perf_counter()
        for k in range(n):
            if calls:
                bodies[17].ApplyForce((5.0, 0.0), (bodies[17].GetPosition().x, bodies[17].GetPosition().y))
                bodies[4011].ApplyTorque(0.5)
                bodies[9000 + k].SetLinearVelocity((0.0, 0.1))
            w.Step(DT, 8, 3)
        bodies[0].GetPosition()
        return (time.perf_counter() - t0) / n
    plain = min(frames(30, False) for _ in range(2))
    driven = min(frames(30, True) for _ in range(2))
    print("20,000-body pile: %.3f ms per frame, %.3f ms with three per-body calls and two reads per frame" % (plain * 1e3, driven * 1e3))
    assert driven < plain + 0.6e-3, (plain, driven)      # three rows out, two back: ~0.1 ms (measured); the whole-array path took ~5 ms here
    # a program that reads MANY bodies after a step (drawing them) gets the bulk copy after a few single rows, not 20,000 row trips
    w.Step(DT, 8, 3)
    t0 = time.perf_counter()
    ys = [b.GetPosition().y for b in bodies]
    read_all = time.perf_counter() - t0
    print("reading all 20,000 positions after a step: %.1f ms" % (read_all * 1e3))
    assert read_all < 1.0 and min(ys) > 0.0, read_all       # (row by row: 20,000 x ~25 us of device round trips + the Python calls)
    w.close()


def test_world_close_returns_device_memory(gpu_api):
    """dbx_world_destroy frees every device pool of the world, the tile solver's and the I/O staging included (a jointed pile
    large enough for k_solve_tiles, stepped, read in bulk, closed -- ten times over)"""
    import torch

    def cycle():
        w, bodies, _ = scenes.pile(api=gpu_api, n=6000, columns=120)
        w.StepN(DT, 8, 3, 20)
        bodies[5].ApplyTorque(0.1); bodies[5].GetPosition()
        w.Step(DT, 8, 3)
        w.read_contacts(); w.counts()
        w.close()
    cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(10):
        cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < (8 << 20), (free0, free1)          # ten leaked worlds of this size would be ~300 MB


def test_worlds_on_two_devices_in_one_process(gpu_api):
    """A process may hold worlds on several GPUs (one host thread driving two agents' worlds): the tile solver's launch attributes
    are per device, and the same scene steps to the same bits on either."""
    if gpu_api.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    states = []
    for dev in (0, 1):
        w, _, _ = scenes.pile(api=gpu_api, n=3000, columns=100, device=dev)
        w.SetAllowSleeping(False)
        w.StepN(DT, 8, 3, 40)
        sb, n = w.read_bodies()
        states.append([(sb[i].c.x, sb[i].c.y, sb[i].a, sb[i].v.x, sb[i].v.y, sb[i].w) for i in range(n)])
        w.close()
    assert states[0] == states[1]


def test_stats_allreduce_through_nccl(gpu_api):
    """dbx_stats_allreduce (the batched path's only collective): sums and maxima through a real NCCL
    communicator created by the host program -- here a one-rank communicator from ncclCommInitAll, so the reduction must
    hand back what went in; bad arguments are refused."""
    import glob
    import os
    import sys
    cands = [f for d in sys.path for f in sorted(glob.glob(os.path.join(d, "nvidia", "nccl", "lib", "libnccl.so*")))] + ["libnccl.so.2"]
    nccl = None
    for c in cands:
        try:
            nccl = C.CDLL(c, mode=C.RTLD_GLOBAL)
            break
        except OSError:
            continue
    if nccl is None:
        pytest.skip("no NCCL library on this box")
    comm = C.c_void_p()
    devs = (C.c_int * 1)(0)
    assert nccl.ncclCommInitAll(C.byref(comm), 1, devs) == 0
    sums = (C.c_double * 3)(65536.0, 38.7e6, 13.9e6)
    maxs = (C.c_double * 2)(19.5, 0.25)
    assert gpu_api.stats_allreduce(comm, None, sums, 3, maxs, 2) == 0, gpu_api.last_error()
    assert list(sums) == [65536.0, 38.7e6, 13.9e6] and list(maxs) == [19.5, 0.25]
    assert gpu_api.stats_allreduce(None, None, sums, 3, maxs, 2) == A.DBX_E_INVALID
    assert gpu_api.stats_allreduce(comm, None, None, 3, maxs, 2) == A.DBX_E_INVALID
    nccl.ncclCommDestroy(comm)


def test_compact_io_records_equal_the_full_ones(gpu_api):
    """dbx_world_set_io_format(DBX_IO_COMPACT): 12-byte force / pose records move the same numbers as the 16-byte ones --
    a pyramid pushed by the same per-body forces through both formats evolves bit for bit alike, and the poses read back are
    (p.x, p.y, angle) of the transforms; the pipelined calls follow the format too."""
    import numpy as np
    wa, _ = scenes.pyramid(api=gpu_api); wb, _ = scenes.pyramid(api=gpu_api)
    n = wa.counts().bodies
    rng = np.random.RandomState(5)
    f4 = np.zeros((n, 4), np.float32); f4[:, :3] = rng.uniform(-30, 30, (n, 3))
    f3 = np.ascontiguousarray(f4[:, :3])
    assert gpu_api.world_set_io_format(wb._w, A.IO_COMPACT) == 0
    assert gpu_api.world_set_io_format(wb._w, 7) == A.DBX_E_INVALID
    xf = np.zeros((n, 4), np.float32); pose = np.zeros((n, 3), np.float32)
    for k in range(40):
        assert gpu_api.world_apply_forces(wa._w, f4.ctypes.data, n) == n
        assert gpu_api.world_apply_forces(wb._w, f3.ctypes.data, n) == n
        wa.Step(DT, 8, 3); wb.Step(DT, 8, 3)
    assert gpu_api.world_read_transforms(wa._w, xf.ctypes.data, n) == n
    assert gpu_api.world_read_transforms(wb._w, pose.ctypes.data, n) == n
    sa, _ = wa.read_bodies(); sb, _ = wb.read_bodies()
    for i in range(n):
        assert (sa[i].c.x, sa[i].c.y, sa[i].a, sa[i].v.x, sa[i].w) == (sb[i].c.x, sb[i].c.y, sb[i].a, sb[i].v.x, sb[i].w), i
        assert (pose[i, 0], pose[i, 1]) == (xf[i, 0], xf[i, 1]) and pose[i, 2] == np.float32(sb[i].a)
    # pipelined calls, compact: the same numbers arrive
    pose2 = np.zeros((n, 3), np.float32)
    t = gpu_api.world_read_transforms_async(wb._w, pose2.ctypes.data, n)
    assert t > 0 and gpu_api.world_io_wait(wb._w, t) == 0 and gpu_api.world_sync(wb._w) == 0
    assert np.array_equal(pose, pose2)
    wa.close(); wb.close()
