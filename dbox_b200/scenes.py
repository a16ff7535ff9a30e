"""Scene generators for BASELINE.json's configs, written against the b2World mirror API (dbox_b200.world) exactly
as the reference programs are written against dbox, so one function builds the same scene on the GPU library and
(in the tests) on the CPU oracle.

 config 1  hello_world   examples/hello_world/hello_world.d:31-103
 config 2  pyramid       examples/demo/tests/pyramid.d:39-78 (+ the demo base class's empty body, framework/test.d:160-161)
 config 3  tumbler       examples/demo/tests/tumbler.d:40-97 (mixed boxes/circles, scalable container)
 config 4  pile          SURVEY.md 8(d): jittered box/circle grid on a chain floor with revolute + distance chains
"""
import ctypes as C

from .world import (b2BodyDef, b2ChainShape, b2CircleShape, b2DistanceJointDef, b2EdgeShape, b2FixtureDef,
                    b2PolygonShape, b2RevoluteJointDef, b2Vec2, b2World, b2_dynamicBody, b2_pi)


def f32(x):
    return C.c_float(x).value


class Mt19937:
    """std::mt19937 (SURVEY.md 8(d) names `std::mt19937(12345)` for the pile's jitter): MT19937 seeded with init_genrand(seed),
    one 32-bit draw per call, and uniform(a, b) the way libstdc++'s uniform_real_distribution<float> maps a draw
    (generate_canonical<float, 24>: float(draw) / 2^32, below 1; then r * (b - a) + a in float arithmetic)."""

    def __init__(self, seed):
        mt = [0] * 624
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt, self.idx = mt, 624

    def _twist(self):
        mt = self.mt
        for k in range(624):
            y = (mt[k] & 0x80000000) | (mt[(k + 1) % 624] & 0x7FFFFFFF)
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.idx = 0

    def draw(self):
        if self.idx >= 624:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def canonical(self):
        r = f32(f32(self.draw()) / 4294967296.0)
        return r if r < 1.0 else f32(1.0 - 2.0 ** -24)

    def uniform(self, a, b):
        return f32(f32(self.canonical() * f32(b - a)) + f32(a))


def hello_world(api=None, **kw):
    world = b2World((0.0, -10.0), api=api, **kw)
    groundBodyDef = b2BodyDef()
    groundBodyDef.position.Set(0.0, -10.0)
    groundBody = world.CreateBody(groundBodyDef)
    groundBox = b2PolygonShape(world._api)
    groundBox.SetAsBox(50.0, 10.0)
    groundBody.CreateFixture(groundBox, 0.0)
    bodyDef = b2BodyDef()
    bodyDef.type = b2_dynamicBody
    bodyDef.position.Set(0.0, 4.0)
    body = world.CreateBody(bodyDef)
    dynamicBox = b2PolygonShape(world._api)
    dynamicBox.SetAsBox(1.0, 1.0)
    fixtureDef = b2FixtureDef()
    fixtureDef.shape = dynamicBox
    fixtureDef.density = 1.0
    fixtureDef.friction = 0.3
    body.CreateFixture(fixtureDef)
    return world, body


def pyramid(api=None, count=20, demo_ground_body=True, world=None, offset=(0.0, 0.0), **kw):
    if world is None:
        world = b2World((0.0, -10.0), api=api, **kw)
    ox, oy = offset
    if demo_ground_body:
        world.CreateBody(b2BodyDef())  # framework/test.d:160-161: fixture-less static body created by every demo
    bd = b2BodyDef()
    bd.position.Set(ox, oy)
    ground = world.CreateBody(bd)
    shape = b2EdgeShape(world._api)
    shape.Set((-40.0, 0.0), (40.0, 0.0))
    ground.CreateFixture(shape, 0.0)
    a = 0.5
    box = b2PolygonShape(world._api)
    box.SetAsBox(a, a)
    x = [f32(-7.0), f32(0.75)]
    deltaX = (0.5625, 1.25)
    deltaY = (1.125, 0.0)
    bodies = []
    for i in range(count):
        y = list(x)
        for j in range(i, count):
            bd = b2BodyDef()
            bd.type = b2_dynamicBody
            bd.position.Set(f32(y[0] + ox), f32(y[1] + oy))
            body = world.CreateBody(bd)
            body.CreateFixture(box, 5.0)
            bodies.append(body)
            y = [f32(y[0] + deltaY[0]), f32(y[1] + deltaY[1])]
        x = [f32(x[0] + deltaX[0]), f32(x[1] + deltaX[1])]
    return world, bodies


class Tumbler:
    """tumbler.d:40-97.  `scale` grows the container (SURVEY.md 8(d) uses 5 for the 20k-body case);
    `mixed` alternates boxes and circles."""

    def __init__(self, api=None, count=800, scale=1.0, mixed=True, **kw):
        self.world = world = b2World((0.0, -10.0), api=api, **kw)
        self.count, self.m_count, self.mixed = count, 0, mixed
        s = scale
        world.CreateBody(b2BodyDef())  # demo base-class body
        ground = world.CreateBody(b2BodyDef())
        bd = b2BodyDef()
        bd.type = b2_dynamicBody
        bd.allowSleep = False
        bd.position.Set(0.0, 10.0 * s)
        body = world.CreateBody(bd)
        shape = b2PolygonShape(world._api)
        for (hx, hy, c) in ((0.5 * s, 10.0 * s, (10.0 * s, 0.0)), (0.5 * s, 10.0 * s, (-10.0 * s, 0.0)),
                            (10.0 * s, 0.5 * s, (0.0, 10.0 * s)), (10.0 * s, 0.5 * s, (0.0, -10.0 * s))):
            shape.SetAsBox(hx, hy, c, 0.0)
            body.CreateFixture(shape, 5.0)
        jd = b2RevoluteJointDef()
        jd.bodyA, jd.bodyB = ground, body
        jd.localAnchorA.Set(0.0, 10.0 * s)
        jd.localAnchorB.Set(0.0, 0.0)
        jd.referenceAngle = 0.0
        jd.motorSpeed = f32(f32(0.05) * f32(b2_pi))
        jd.maxMotorTorque = 1e8
        jd.enableMotor = True
        self.joint = world.CreateJoint(jd)
        self.container = body
        self.scale = s
        self.bodies = []

    def Step(self, dt=1.0 / 60.0, vi=8, pi=3, spawn_per_step=1):
        """tumbler.d:78-97: one new body per step at the container's centre; spawn_per_step > 1 (grow a big scene in fewer
        steps) lays the extra bodies out side by side, half a unit apart, instead of on top of each other"""
        self.world.Step(dt, vi, pi)
        self.spawn(spawn_per_step)

    def spawn(self, n=1):
        """the body creation half of tumbler.d:84-96, `n` bodies side by side (also used to build a twin of a grown scene
        without stepping it: same bodies in the same order, state transplanted afterwards)"""
        for k in range(n):
            if self.m_count < self.count:
                bd = b2BodyDef()
                bd.type = b2_dynamicBody
                bd.position.Set(f32(0.5 * (k - 0.5 * (n - 1))), 10.0 * self.scale)
                body = self.world.CreateBody(bd)
                if self.mixed and (self.m_count & 1):
                    shape = b2CircleShape(self.world._api)
                    shape.m_radius = 0.125
                else:
                    shape = b2PolygonShape(self.world._api)
                    shape.SetAsBox(0.125, 0.125)
                body.CreateFixture(shape, 1.0)
                self.bodies.append(body)
                self.m_count += 1


def islands_of_boxes(api=None, clusters=12, side=16, gap=40.0, seed=7, **kw):
    """`clusters` separate side x side grids of boxes on one long floor, `gap` metres apart: enough dynamic bodies for the tile
    solver, cut so that every tile is one cluster -- no boundary and no global constraint anywhere"""
    world = b2World((0.0, -10.0), api=api, **kw)
    rng = Mt19937(seed)
    ground = world.CreateBody(b2BodyDef())
    floor = b2EdgeShape(world._api)
    half = f32(0.5 * clusters * gap + 20.0)
    floor.Set((-half, 0.0), (half, 0.0))
    ground.CreateFixture(floor, 0.0)
    box = b2PolygonShape(world._api)
    box.SetAsBox(0.5, 0.5)
    fd = b2FixtureDef()
    fd.shape, fd.density, fd.friction = box, 1.0, 0.3
    bodies = []
    for c in range(clusters):
        cx = f32(-0.5 * (clusters - 1) * gap + c * gap)
        for i in range(side * side):
            col, row = i % side, i // side
            bd = b2BodyDef()
            bd.type = b2_dynamicBody
            bd.position.Set(f32(cx + 1.05 * (col - 0.5 * (side - 1)) + rng.uniform(-0.01, 0.01)), f32(0.55 + 1.05 * row + rng.uniform(-0.01, 0.01)))
            b = world.CreateBody(bd)
            b.CreateFixture(fd)
            bodies.append(b)
    return world, bodies


def pile(api=None, n=100000, columns=1000, joints=True, circles=True, seed=12345, world=None, long_links=0, **kw):
    """Config 4: `n` dynamic bodies in a jittered grid `columns` wide above a static chain floor with
    two edge walls, the jitter U(-0.01, 0.01) and the shape choice drawn from std::mt19937(seed) (x, y, shape per body, in body
    order); 70 % boxes (half-extent 0.5, density 1, friction 0.3), 30 % circles r=0.5; every 10th column is
    linked upward into 25-body revolute chains and every 10th row sideways into 25-body rigid distance chains."""
    if world is None:
        world = b2World((0.0, -10.0), api=api, **kw)
    rng = Mt19937(seed)
    half_w = f32(1.05 * columns * 0.5 + 5.0)
    rows = (n + columns - 1) // columns
    ground = world.CreateBody(b2BodyDef())
    floor = b2ChainShape(world._api)
    segs = max(2, columns // 50)
    floor.CreateChain([(f32(-half_w + 2.0 * half_w * k / segs), 0.0) for k in range(segs + 1)])
    ground.CreateFixture(floor, 0.0)
    wall = b2EdgeShape(world._api)
    top = f32(1.05 * rows + 20.0)
    wall.Set((-half_w, 0.0), (-half_w, top))
    ground.CreateFixture(wall, 0.0)
    wall.Set((half_w, 0.0), (half_w, top))
    ground.CreateFixture(wall, 0.0)

    box = b2PolygonShape(world._api)
    box.SetAsBox(0.5, 0.5)
    circle = b2CircleShape(world._api)
    circle.m_radius = 0.5
    fd = b2FixtureDef()
    fd.density, fd.friction = 1.0, 0.3
    bodies = []
    x0 = f32(-1.05 * columns * 0.5 + 0.525)
    for i in range(n):
        col, row = i % columns, i // columns
        bd = b2BodyDef()
        bd.type = b2_dynamicBody
        bd.position.Set(f32(x0 + 1.05 * col + rng.uniform(-0.01, 0.01)), f32(0.55 + 1.05 * row + rng.uniform(-0.01, 0.01)))
        body = world.CreateBody(bd)
        fd.shape = circle if (circles and rng.canonical() < 0.3) else box
        body.CreateFixture(fd)
        bodies.append(body)
    njoints = 0
    if joints:
        # revolute chains: column c (c % 10 == 0), rows r..r+24 linked pairwise at the midpoint, collideConnected=false
        for col in range(0, columns, 10):
            for row in range(rows - 1):
                if row % 25 == 24:
                    continue
                i, j = row * columns + col, (row + 1) * columns + col
                if j >= n:
                    continue
                jd = b2RevoluteJointDef()
                pa, pb = bodies[i].GetPosition(), bodies[j].GetPosition()
                jd.Initialize(bodies[i], bodies[j], (f32(0.5 * (pa.x + pb.x)), f32(0.5 * (pa.y + pb.y))))
                world.CreateJoint(jd)
                njoints += 1
        # rigid distance chains: row r (r % 10 == 5), columns linked pairwise centre to centre
        for row in range(5, rows, 10):
            for col in range(columns - 1):
                if col % 25 == 24 or col % 10 == 0 or (col + 1) % 10 == 0:
                    continue
                i, j = row * columns + col, row * columns + col + 1
                if j >= n:
                    continue
                jd = b2DistanceJointDef()
                pa, pb = bodies[i].GetPosition(), bodies[j].GetPosition()
                jd.Initialize(bodies[i], bodies[j], (pa.x, pa.y), (pb.x, pb.y))
                jd.collideConnected = True
                world.CreateJoint(jd)
                njoints += 1
    # `long_links` rigid distance joints between bodies half the pile apart (top rows, at their current distance): constraints
    # that no spatial partition of the bodies can keep local (the tile solver's "global" class)
    for t in range(long_links):
        row, col = rows - 1 - (t % 3), (7 * t) % max(1, columns // 2)
        i, j = row * columns + col, row * columns + col + columns // 2
        if i < 0 or j >= n:
            continue
        jd = b2DistanceJointDef()
        pa, pb = bodies[i].GetPosition(), bodies[j].GetPosition()
        jd.Initialize(bodies[i], bodies[j], (pa.x, pa.y), (pb.x, pb.y))
        jd.collideConnected = True
        world.CreateJoint(jd)
        njoints += 1
    return world, bodies, njoints
