"""Host-side mirror of the reference's public API for the hot path, over the C ABI.

Same names and argument meaning as dbox's D API (dynamics/b2world.d, b2body.d, b2fixture.d,
collision/shapes/*.d, dynamics/joints/b2revolutejoint.d, b2distancejoint.d) so that scene code and the parity
tests read like the reference's own programs (examples/hello_world/hello_world.d:31-103).  Every method is a
thin forward to one `dbx_*` entry point of include/dbox_b200.h — no physics runs in Python, and there is no CPU
fallback: `b2World()` raises if the CUDA library cannot create a device world.

`b2World(gravity, api=...)` takes the ABI binding to drive; the default is the product library
(dbox_b200.lib.api()).  The tests pass the CPU oracle's binding to build the *same* scene on both sides.
"""
import ctypes as C
import math

from . import _abi as A


class b2Vec2:
    __slots__ = ("x", "y")

    def __init__(self, x=0.0, y=0.0):
        self.x, self.y = float(x), float(y)

    def Set(self, x, y):
        self.x, self.y = float(x), float(y)

    def __iter__(self):
        yield self.x
        yield self.y

    def __repr__(self):
        return "b2Vec2(%r, %r)" % (self.x, self.y)


def _v(v):
    x, y = v
    return A.Vec2(x, y)


b2_staticBody, b2_kinematicBody, b2_dynamicBody = A.STATIC_BODY, A.KINEMATIC_BODY, A.DYNAMIC_BODY
b2_pi = 3.14159265359
b2_linearSlop = 0.005
b2_polygonRadius = 2.0 * b2_linearSlop


class b2BodyDef:
    """dynamics/b2body.d:51-104"""

    def __init__(self):
        self.type = b2_staticBody
        self.position = b2Vec2(0, 0)
        self.angle = 0.0
        self.linearVelocity = b2Vec2(0, 0)
        self.angularVelocity = 0.0
        self.linearDamping = 0.0
        self.angularDamping = 0.0
        self.allowSleep = True
        self.awake = True
        self.fixedRotation = False
        self.bullet = False
        self.active = True
        self.userData = 0
        self.gravityScale = 1.0

    def _pod(self):
        d = A.BodyDef()
        d.type = self.type
        d.position = _v(self.position)
        d.angle = self.angle
        d.linearVelocity = _v(self.linearVelocity)
        d.angularVelocity = self.angularVelocity
        d.linearDamping, d.angularDamping = self.linearDamping, self.angularDamping
        d.allowSleep, d.awake, d.fixedRotation = int(self.allowSleep), int(self.awake), int(self.fixedRotation)
        d.bullet, d.active, d.gravityScale, d.userData = int(self.bullet), int(self.active), self.gravityScale, self.userData
        return d


class b2Filter:
    """dynamics/b2fixture.d:32-45"""

    def __init__(self):
        self.categoryBits, self.maskBits, self.groupIndex = 0x0001, 0xFFFF, 0


class b2FixtureDef:
    """dynamics/b2fixture.d:49-73"""

    def __init__(self):
        self.shape = None
        self.userData = 0
        self.friction = 0.2
        self.restitution = 0.0
        self.density = 0.0
        self.isSensor = False
        self.filter = b2Filter()

    def _pod(self):
        d = A.FixtureDef()
        d.friction, d.restitution, d.density, d.isSensor = self.friction, self.restitution, self.density, int(self.isSensor)
        d.categoryBits, d.maskBits, d.groupIndex = self.filter.categoryBits, self.filter.maskBits, self.filter.groupIndex
        d.userData = self.userData
        return d


class b2Shape:
    """collision/shapes/b2shape.d:43-100.  Setup-time geometry helpers run in the library (host side)."""
    e_circle, e_edge, e_polygon, e_chain = 0, 1, 2, 3

    def __init__(self, api=None):
        if api is None:
            from . import lib
            api = lib.api()
        self._api = api
        self._pod = A.Shape()
        self._keep = None

    def GetType(self):
        return self._pod.type

    @property
    def m_radius(self):
        return self._pod.radius

    @m_radius.setter
    def m_radius(self, r):
        self._pod.radius = r


class b2CircleShape(b2Shape):
    def __init__(self, api=None):
        super().__init__(api)
        self._api.shape_set_circle(C.byref(self._pod), 0.0, 0.0, 0.0)

    @property
    def m_p(self):
        return b2Vec2(self._pod.p.x, self._pod.p.y)

    @m_p.setter
    def m_p(self, v):
        self._pod.p = _v(v)


class b2EdgeShape(b2Shape):
    def __init__(self, api=None):
        super().__init__(api)
        self._api.shape_set_edge(C.byref(self._pod), A.Vec2(0, 0), A.Vec2(0, 0))

    def Set(self, v1, v2):
        self._api.shape_set_edge(C.byref(self._pod), _v(v1), _v(v2))


class b2PolygonShape(b2Shape):
    def __init__(self, api=None):
        super().__init__(api)
        self._pod.type = A.SHAPE_POLYGON
        self._pod.radius = b2_polygonRadius

    def SetAsBox(self, hx, hy, center=None, angle=0.0):
        if center is None:
            self._api.shape_set_box(C.byref(self._pod), hx, hy)
        else:
            self._api.shape_set_box_at(C.byref(self._pod), hx, hy, _v(center), angle)

    def Set(self, vertices):
        arr = (A.Vec2 * len(vertices))(*[_v(p) for p in vertices])
        self._api.shape_set_polygon(C.byref(self._pod), arr, len(vertices))

    @property
    def m_count(self):
        return self._pod.count


class b2ChainShape(b2Shape):
    def __init__(self, api=None):
        super().__init__(api)
        self._pod.type = A.SHAPE_CHAIN
        self._pod.radius = b2_polygonRadius

    def CreateChain(self, vertices):
        self._keep = (A.Vec2 * len(vertices))(*[_v(p) for p in vertices])
        self._api.shape_set_chain(C.byref(self._pod), self._keep, len(vertices), 0)

    def CreateLoop(self, vertices):
        vs = list(vertices) + [vertices[0]]
        self._keep = (A.Vec2 * len(vs))(*[_v(p) for p in vs])
        self._api.shape_set_chain(C.byref(self._pod), self._keep, len(vs), 1)


class b2JointDef:
    def __init__(self):
        self.userData, self.bodyA, self.bodyB, self.collideConnected = 0, None, None, False


class b2RevoluteJointDef(b2JointDef):
    """dynamics/joints/b2revolutejoint.d:39-107"""
    type = A.JOINT_REVOLUTE

    def __init__(self):
        super().__init__()
        self.localAnchorA, self.localAnchorB = b2Vec2(), b2Vec2()
        self.referenceAngle = self.lowerAngle = self.upperAngle = self.maxMotorTorque = self.motorSpeed = 0.0
        self.enableLimit = self.enableMotor = False

    def Initialize(self, bA, bB, anchor):
        self.bodyA, self.bodyB = bA, bB
        self.localAnchorA = bA.GetLocalPoint(anchor)
        self.localAnchorB = bB.GetLocalPoint(anchor)
        self.referenceAngle = _f32(bB.GetAngle() - bA.GetAngle())

    def _pod(self):
        d = A.JointDef()
        d.type, d.bodyA, d.bodyB, d.collideConnected = self.type, self.bodyA.id, self.bodyB.id, int(self.collideConnected)
        d.localAnchorA, d.localAnchorB = _v(self.localAnchorA), _v(self.localAnchorB)
        d.referenceAngle, d.enableLimit, d.lowerAngle, d.upperAngle = self.referenceAngle, int(self.enableLimit), self.lowerAngle, self.upperAngle
        d.enableMotor, d.motorSpeed, d.maxMotorTorque, d.userData = int(self.enableMotor), self.motorSpeed, self.maxMotorTorque, self.userData
        return d


class b2DistanceJointDef(b2JointDef):
    """dynamics/joints/b2distancejoint.d:36-90"""
    type = A.JOINT_DISTANCE

    def __init__(self):
        super().__init__()
        self.localAnchorA, self.localAnchorB = b2Vec2(), b2Vec2()
        self.length, self.frequencyHz, self.dampingRatio = 1.0, 0.0, 0.0

    def Initialize(self, b1, b2, anchor1, anchor2):
        self.bodyA, self.bodyB = b1, b2
        self.localAnchorA = b1.GetLocalPoint(anchor1)
        self.localAnchorB = b2.GetLocalPoint(anchor2)
        a1, a2 = _v(anchor1), _v(anchor2)
        dx, dy = _f32(a2.x - a1.x), _f32(a2.y - a1.y)
        self.length = _f32(math.sqrt(_f32(_f32(dx * dx) + _f32(dy * dy))))

    def _pod(self):
        d = A.JointDef()
        d.type, d.bodyA, d.bodyB, d.collideConnected = self.type, self.bodyA.id, self.bodyB.id, int(self.collideConnected)
        d.localAnchorA, d.localAnchorB = _v(self.localAnchorA), _v(self.localAnchorB)
        d.length, d.frequencyHz, d.dampingRatio, d.userData = self.length, self.frequencyHz, self.dampingRatio, self.userData
        return d


class _AnchoredDef(b2JointDef):
    """shared by the second-wave joint defs: head of dbx_joint_def + localAnchorA/B"""

    def __init__(self):
        super().__init__()
        self.localAnchorA, self.localAnchorB = b2Vec2(), b2Vec2()

    def _head(self):
        d = A.JointDef()
        d.type, d.bodyA, d.bodyB, d.collideConnected = self.type, self.bodyA.id, self.bodyB.id, int(self.collideConnected)
        d.localAnchorA, d.localAnchorB, d.userData = _v(self.localAnchorA), _v(self.localAnchorB), self.userData
        return d

    def _anchor(self, bA, bB, anchor):
        self.bodyA, self.bodyB = bA, bB
        self.localAnchorA, self.localAnchorB = bA.GetLocalPoint(anchor), bB.GetLocalPoint(anchor)


class b2PrismaticJointDef(_AnchoredDef):
    """dynamics/joints/b2prismaticjoint.d:39-113"""
    type = A.JOINT_PRISMATIC

    def __init__(self):
        super().__init__()
        self.localAxisA = b2Vec2(1.0, 0.0)
        self.referenceAngle = self.lowerTranslation = self.upperTranslation = self.maxMotorForce = self.motorSpeed = 0.0
        self.enableLimit = self.enableMotor = False

    def Initialize(self, bA, bB, anchor, axis):
        self._anchor(bA, bB, anchor)
        self.localAxisA = bA.GetLocalVector(axis)
        self.referenceAngle = _f32(bB.GetAngle() - bA.GetAngle())

    def _pod(self):
        d = self._head()
        d.localAxisA, d.referenceAngle, d.enableLimit, d.enableMotor = _v(self.localAxisA), self.referenceAngle, int(self.enableLimit), int(self.enableMotor)
        d.lowerTranslation, d.upperTranslation, d.maxMotorForce, d.motorSpeed = self.lowerTranslation, self.upperTranslation, self.maxMotorForce, self.motorSpeed
        return d


class b2WeldJointDef(_AnchoredDef):
    """dynamics/joints/b2weldjoint.d:38-80"""
    type = A.JOINT_WELD

    def __init__(self):
        super().__init__()
        self.referenceAngle = self.frequencyHz = self.dampingRatio = 0.0

    def Initialize(self, bA, bB, anchor):
        self._anchor(bA, bB, anchor)
        self.referenceAngle = _f32(bB.GetAngle() - bA.GetAngle())

    def _pod(self):
        d = self._head()
        d.referenceAngle, d.frequencyHz, d.dampingRatio = self.referenceAngle, self.frequencyHz, self.dampingRatio
        return d


class b2WheelJointDef(_AnchoredDef):
    """dynamics/joints/b2wheeljoint.d:39-92"""
    type = A.JOINT_WHEEL

    def __init__(self):
        super().__init__()
        self.localAxisA = b2Vec2(1.0, 0.0)
        self.enableMotor = False
        self.maxMotorTorque = self.motorSpeed = 0.0
        self.frequencyHz, self.dampingRatio = 2.0, 0.7

    def Initialize(self, bA, bB, anchor, axis):
        self._anchor(bA, bB, anchor)
        self.localAxisA = bA.GetLocalVector(axis)

    def _pod(self):
        d = self._head()
        d.localAxisA, d.enableMotor, d.maxMotorTorque, d.motorSpeed = _v(self.localAxisA), int(self.enableMotor), self.maxMotorTorque, self.motorSpeed
        d.frequencyHz, d.dampingRatio = self.frequencyHz, self.dampingRatio
        return d


class b2RopeJointDef(_AnchoredDef):
    """dynamics/joints/b2ropejoint.d:40-66"""
    type = A.JOINT_ROPE

    def __init__(self):
        super().__init__()
        self.localAnchorA, self.localAnchorB, self.maxLength = b2Vec2(-1.0, 0.0), b2Vec2(1.0, 0.0), 0.0

    def _pod(self):
        d = self._head()
        d.maxLength = self.maxLength
        return d


class b2FrictionJointDef(_AnchoredDef):
    """dynamics/joints/b2frictionjoint.d:36-72"""
    type = A.JOINT_FRICTION

    def __init__(self):
        super().__init__()
        self.maxForce = self.maxTorque = 0.0

    def Initialize(self, bA, bB, anchor):
        self._anchor(bA, bB, anchor)

    def _pod(self):
        d = self._head()
        d.maxForce, d.maxTorque = self.maxForce, self.maxTorque
        return d


class b2MotorJointDef(_AnchoredDef):
    """dynamics/joints/b2motorjoint.d:36-80"""
    type = A.JOINT_MOTOR

    def __init__(self):
        super().__init__()
        self.linearOffset, self.angularOffset, self.maxForce, self.maxTorque, self.correctionFactor = b2Vec2(), 0.0, 1.0, 1.0, 0.3

    def Initialize(self, bA, bB):
        self.bodyA, self.bodyB = bA, bB
        self.linearOffset = bA.GetLocalPoint(bB.GetPosition())
        self.angularOffset = _f32(bB.GetAngle() - bA.GetAngle())

    def _pod(self):
        d = self._head()
        d.linearOffset, d.angularOffset, d.maxForce, d.maxTorque, d.correctionFactor = _v(self.linearOffset), self.angularOffset, self.maxForce, self.maxTorque, self.correctionFactor
        return d


class b2MouseJointDef(_AnchoredDef):
    """dynamics/joints/b2mousejoint.d:36-66"""
    type = A.JOINT_MOUSE

    def __init__(self):
        super().__init__()
        self.target, self.maxForce, self.frequencyHz, self.dampingRatio = b2Vec2(), 0.0, 5.0, 0.7

    def _pod(self):
        d = self._head()
        d.target, d.maxForce, d.frequencyHz, d.dampingRatio = _v(self.target), self.maxForce, self.frequencyHz, self.dampingRatio
        return d


class b2PulleyJointDef(_AnchoredDef):
    """dynamics/joints/b2pulleyjoint.d:41-100"""
    type = A.JOINT_PULLEY

    def __init__(self):
        super().__init__()
        self.groundAnchorA, self.groundAnchorB = b2Vec2(-1.0, 1.0), b2Vec2(1.0, 1.0)
        self.localAnchorA, self.localAnchorB = b2Vec2(-1.0, 0.0), b2Vec2(1.0, 0.0)
        self.lengthA = self.lengthB = 0.0
        self.ratio = 1.0
        self.collideConnected = True

    def Initialize(self, bA, bB, groundA, groundB, anchorA, anchorB, ratio):
        self.bodyA, self.bodyB = bA, bB
        self.groundAnchorA, self.groundAnchorB = _pt(groundA), _pt(groundB)
        self.localAnchorA, self.localAnchorB = bA.GetLocalPoint(anchorA), bB.GetLocalPoint(anchorB)
        a, ga, b, gb = _v(anchorA), _v(groundA), _v(anchorB), _v(groundB)
        self.lengthA = _len32(_f32(a.x - ga.x), _f32(a.y - ga.y))
        self.lengthB = _len32(_f32(b.x - gb.x), _f32(b.y - gb.y))
        self.ratio = ratio

    def _pod(self):
        d = self._head()
        d.groundAnchorA, d.groundAnchorB, d.lengthA, d.lengthB, d.ratio = _v(self.groundAnchorA), _v(self.groundAnchorB), self.lengthA, self.lengthB, self.ratio
        return d


class b2GearJointDef(b2JointDef):
    """dynamics/joints/b2gearjoint.d:36-58: joint1 / joint2 are revolute or prismatic b2Joint handles of this world"""
    type = A.JOINT_GEAR

    def __init__(self):
        super().__init__()
        self.joint1 = self.joint2 = None
        self.ratio = 1.0

    def _pod(self):
        d = A.JointDef()
        d.type, d.collideConnected, d.userData = self.type, int(self.collideConnected), self.userData
        d.bodyA, d.bodyB = self.joint1.bodyB.id, self.joint2.bodyB.id
        d.joint1, d.joint2, d.ratio = self.joint1.id, self.joint2.id, self.ratio
        return d


def _pt(p):
    v = _v(p)
    return b2Vec2(v.x, v.y)


def _len32(dx, dy):
    return _f32(math.sqrt(_f32(_f32(dx * dx) + _f32(dy * dy))))


def _f32(x):
    return C.c_float(x).value


class b2Fixture:
    def __init__(self, body, fid):
        self.body, self.id = body, fid

    def GetBody(self):
        return self.body

    # b2fixture.d:108-262
    def SetFilterData(self, categoryBits=0x0001, maskBits=0xFFFF, groupIndex=0):
        w = self.body.world
        self.filter = (categoryBits, maskBits, groupIndex)
        w._ck(w._api.fixture_set_filter(w._w, self.id, categoryBits, maskBits, groupIndex))
        w._refilter_user(self)

    def SetSensor(self, flag):
        w = self.body.world
        w._ck(w._api.fixture_set_sensor(w._w, self.id, int(flag)))
        self.sensor = bool(flag)

    def IsSensor(self):
        return getattr(self, "sensor", False)

    def SetFriction(self, v):
        w = self.body.world
        w._ck(w._api.fixture_set_friction(w._w, self.id, v))

    def SetRestitution(self, v):
        w = self.body.world
        w._ck(w._api.fixture_set_restitution(w._w, self.id, v))

    def SetDensity(self, v):
        w = self.body.world
        w._ck(w._api.fixture_set_density(w._w, self.id, v))

    def TestPoint(self, p):
        """b2fixture.d:209-212"""
        return self.body.world.TestPoints([(self, p)])[0]


class b2Joint:
    """handle of a joint; the run-time setters of the reference's joint classes forward to dbx_joint_set_params (the joint keeps
    the POD it was created from, so a setter only changes its own fields)"""

    def __init__(self, world, jid, bodyA, bodyB):
        self.world, self.id, self.bodyA, self.bodyB = world, jid, bodyA, bodyB
        self._pod = None

    def _set(self, mask, **fields):
        for k, v in fields.items():
            setattr(self._pod, k, v)
        self.world._ck(self.world._api.joint_set_params(self.world._w, self.id, C.byref(self._pod), mask))

    # b2revolutejoint.d:216-300, b2prismaticjoint.d:250-330, b2wheeljoint.d:170-230
    def SetMotorSpeed(self, speed):
        self._set(A.JP_MOTOR_SPEED, motorSpeed=speed)

    def EnableMotor(self, flag):
        self._set(A.JP_ENABLE_MOTOR, enableMotor=int(flag))

    def SetMaxMotorTorque(self, torque):
        self._set(A.JP_MAX_MOTOR, maxMotorTorque=torque)

    def SetMaxMotorForce(self, force):
        self._set(A.JP_MAX_MOTOR, maxMotorForce=force)

    def EnableLimit(self, flag):
        self._set(A.JP_ENABLE_LIMIT, enableLimit=int(flag))

    def SetLimits(self, lower, upper):
        if self._pod.type == A.JOINT_PRISMATIC:
            self._set(A.JP_LIMITS, lowerTranslation=lower, upperTranslation=upper)
        else:
            self._set(A.JP_LIMITS, lowerAngle=lower, upperAngle=upper)

    # springs and lengths: b2distancejoint.d, b2weldjoint.d, b2wheeljoint.d, b2mousejoint.d, b2ropejoint.d
    def SetFrequency(self, hz):
        self._set(A.JP_SPRING, frequencyHz=hz)

    def SetDampingRatio(self, ratio):
        self._set(A.JP_SPRING, dampingRatio=ratio)

    def SetLength(self, length):
        self._set(A.JP_LENGTH, length=length)

    def SetMaxLength(self, length):
        self._set(A.JP_LENGTH, maxLength=length)

    # b2frictionjoint.d, b2motorjoint.d, b2mousejoint.d
    def SetMaxForce(self, force):
        self._set(A.JP_MAX_FORCE, maxForce=force)

    def SetMaxTorque(self, torque):
        self._set(A.JP_MAX_FORCE, maxTorque=torque)

    def SetLinearOffset(self, offset):
        self._set(A.JP_OFFSETS, linearOffset=_v(offset))

    def SetAngularOffset(self, angle):
        self._set(A.JP_OFFSETS, angularOffset=angle)

    def SetCorrectionFactor(self, factor):
        self._set(A.JP_CORRECTION, correctionFactor=factor)

    # accessors computed in the shim from body states and the joint's accumulated impulses (b2revolutejoint.d:140-214)
    def _state(self):
        js, n = self.world.read_joints()
        return js[sorted(self.world._joints).index(self.id)]

    def GetJointAngle(self):
        return C.c_float(C.c_float(self.bodyB.GetAngle() - self.bodyA.GetAngle()).value - self._pod.referenceAngle).value

    def GetJointSpeed(self):
        return self.bodyB.GetAngularVelocity() - self.bodyA.GetAngularVelocity()

    def GetMotorTorque(self, inv_dt):
        return inv_dt * self._state().motorImpulse

    def GetReactionForce(self, inv_dt):
        st = self._state()
        return (inv_dt * st.impulse[0], inv_dt * st.impulse[1])


class b2Body:
    """dynamics/b2body.d:107-1219 (the accessors and mutators the hot path's callers use)"""

    def __init__(self, world, bid):
        self.world, self.id = world, bid
        self.fixtures = []

    def _state(self):
        s = A.BodyState()
        self.world._ck(self.world._api.body_get_state(self.world._w, self.id, C.byref(s)))
        return s

    def CreateFixture(self, shape_or_def, density=None):
        """b2body.d:116-170: CreateFixture(&fixtureDef) or CreateFixture(shape, density)."""
        if isinstance(shape_or_def, b2FixtureDef):
            fd, shape = shape_or_def, shape_or_def.shape
        else:
            fd, shape = b2FixtureDef(), shape_or_def
            fd.density = 0.0 if density is None else density
        pod = fd._pod()
        fid = self.world._ck(self.world._api.fixture_create(self.world._w, self.id, C.byref(pod), C.byref(shape._pod)))
        f = b2Fixture(self, fid)
        f.filter = (fd.filter.categoryBits, fd.filter.maskBits, fd.filter.groupIndex)
        f.sensor = bool(fd.isSensor)
        self.fixtures.append(f)
        self.world._fixtures[fid] = f
        return f

    def DestroyFixture(self, fixture):
        """b2body.d:172-254"""
        self.world._ck(self.world._api.fixture_destroy(self.world._w, fixture.id))
        self.fixtures.remove(fixture)
        self.world._fixtures.pop(fixture.id, None)

    def GetPosition(self):
        s = self._state()
        return b2Vec2(s.p.x, s.p.y)

    def GetAngle(self):
        return self._state().a

    def GetWorldCenter(self):
        s = self._state()
        return b2Vec2(s.c.x, s.c.y)

    def GetLinearVelocity(self):
        s = self._state()
        return b2Vec2(s.v.x, s.v.y)

    def GetAngularVelocity(self):
        return self._state().w

    def GetMass(self):
        return self._state().mass

    def IsAwake(self):
        return bool(self._state().flags & A.BODY_AWAKE)

    def GetLocalPoint(self, worldPoint):
        """b2body.d GetLocalPoint = b2MulT(m_xf, p) (common/b2math.d:738-746), evaluated in fp32."""
        s = self._state()
        wx, wy = worldPoint
        px, py = _f32(_f32(wx) - s.p.x), _f32(_f32(wy) - s.p.y)
        x = _f32(_f32(s.qc * px) + _f32(s.qs * py))
        y = _f32(_f32(-s.qs * px) + _f32(s.qc * py))
        return b2Vec2(x, y)

    def SetMassData(self, mass, center, I):
        """b2body.d:502-540"""
        cx, cy = center
        self.world._ck(self.world._api.body_set_mass_data(self.world._w, self.id, mass, cx, cy, I))

    def ResetMassData(self):
        self.world._ck(self.world._api.body_reset_mass_data(self.world._w, self.id))

    def SetFixedRotation(self, flag):
        self.world._ck(self.world._api.body_set_fixed_rotation(self.world._w, self.id, int(flag)))

    def SetLinearDamping(self, d):
        self.world._ck(self.world._api.body_set_linear_damping(self.world._w, self.id, d))

    def SetAngularDamping(self, d):
        self.world._ck(self.world._api.body_set_angular_damping(self.world._w, self.id, d))

    def SetGravityScale(self, s):
        self.world._ck(self.world._api.body_set_gravity_scale(self.world._w, self.id, s))

    def SetType(self, type):
        """b2body.d:867-914"""
        self.world._ck(self.world._api.body_set_type(self.world._w, self.id, type))

    def SetActive(self, flag):
        """b2body.d:718-775"""
        self.world._ck(self.world._api.body_set_active(self.world._w, self.id, int(flag)))

    def GetLocalVector(self, worldVector):
        """b2body.d GetLocalVector = b2MulT(m_xf.q, v) (common/b2math.d:640-643), evaluated in fp32."""
        s = self._state()
        vx, vy = worldVector
        vx, vy = _f32(vx), _f32(vy)
        return b2Vec2(_f32(_f32(s.qc * vx) + _f32(s.qs * vy)), _f32(_f32(-s.qs * vx) + _f32(s.qc * vy)))

    def SetTransform(self, position, angle):
        x, y = position
        self.world._ck(self.world._api.body_set_transform(self.world._w, self.id, x, y, angle))

    def SetLinearVelocity(self, v):
        x, y = v
        self.world._ck(self.world._api.body_set_linear_velocity(self.world._w, self.id, x, y))

    def SetAngularVelocity(self, w):
        self.world._ck(self.world._api.body_set_angular_velocity(self.world._w, self.id, w))

    def ApplyForce(self, force, point, wake=True):
        (fx, fy), (px, py) = force, point
        self.world._ck(self.world._api.body_apply_force(self.world._w, self.id, fx, fy, px, py, int(wake)))

    def ApplyTorque(self, torque, wake=True):
        self.world._ck(self.world._api.body_apply_torque(self.world._w, self.id, torque, int(wake)))

    def ApplyLinearImpulse(self, impulse, point, wake=True):
        (ix, iy), (px, py) = impulse, point
        self.world._ck(self.world._api.body_apply_linear_impulse(self.world._w, self.id, ix, iy, px, py, int(wake)))

    def ApplyAngularImpulse(self, impulse, wake=True):
        self.world._ck(self.world._api.body_apply_angular_impulse(self.world._w, self.id, impulse, int(wake)))

    def SetAwake(self, flag):
        self.world._ck(self.world._api.body_set_awake(self.world._w, self.id, int(flag)))

    def SetBullet(self, flag):
        self.world._ck(self.world._api.body_set_bullet(self.world._w, self.id, int(flag)))

    def SetSleepingAllowed(self, flag):
        self.world._ck(self.world._api.body_set_sleeping_allowed(self.world._w, self.id, int(flag)))


class b2ContactListener:
    """b2worldcallbacks.d:87-128.  BeginContact / EndContact / PostSolve are delivered right after b2World.Step returns
    (include/dbox_b200.h, "contact listener, deferred"); PreSolve goes through b2World.StepWithPreSolve."""

    def BeginContact(self, contact):
        pass

    def EndContact(self, contact):
        pass

    post_solve = False      # set True (or override PostSolve and set it) to have the step record b2ContactImpulse per contact

    def PostSolve(self, contact, impulse):
        """impulse = (count, normalImpulses, tangentImpulses) as b2ContactImpulse (b2worldcallbacks.d:73-79)"""


class b2ContactFilter:
    """b2worldcallbacks.d:36-66.  Subclass and override ShouldCollide; the base implementation is the default category / mask /
    group rule, so an override can call super().ShouldCollide(a, b) like the reference's subclasses do."""

    def ShouldCollide(self, fixtureA, fixtureB):
        catA, maskA, groupA = fixtureA.filter
        catB, maskB, groupB = fixtureB.filter
        if groupA == groupB and groupA != 0:
            return groupA > 0
        return (maskA & catB) != 0 and (catA & maskB) != 0


class b2ContactView:
    """what a deferred listener call sees of a b2Contact: its fixtures and child indices (b2contact.d:108-135)"""

    def __init__(self, world, ev):
        self.world, self.event = world, ev
        self.fixtureA_id, self.fixtureB_id, self.childA, self.childB, self.bodyA_id, self.bodyB_id = ev[3:9]

    def GetFixtureA(self):
        return self.world._fixtures.get(self.fixtureA_id)

    def GetFixtureB(self):
        return self.world._fixtures.get(self.fixtureB_id)

    def GetChildIndexA(self):
        return self.childA

    def GetChildIndexB(self):
        return self.childB


class b2World:
    """dynamics/b2world.d:34-1591 — construction, factories, Step, toggles and bulk state access."""

    def __init__(self, gravity, api=None, device=0, caps=None):
        if api is None:
            from . import lib
            api = lib.api()
        self._api = api
        gx, gy = gravity
        if api.prefix == "dbx_":
            self._w = api.world_create(gx, gy, device, C.byref(caps) if caps is not None else None)
            if not self._w:
                raise RuntimeError("dbx_world_create failed: %s" % api.last_error().decode())
        else:
            self._w = api.world_create(gx, gy)
        self._bodies, self._fixtures, self._joints = {}, {}, {}
        self._flags = A.WORLD_DEFAULT_FLAGS

    def close(self):
        if self._w:
            self._api.world_destroy(self._w)
            self._w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            msg = self._api.last_error().decode() if hasattr(self._api, "last_error") else ""
            raise RuntimeError("ABI call failed with %d %s" % (rc, msg))
        return rc

    # factories ---------------------------------------------------------------------------------------------
    def CreateBody(self, bodyDef):
        pod = bodyDef._pod()
        bid = self._ck(self._api.body_create(self._w, C.byref(pod)))
        b = b2Body(self, bid)
        self._bodies[bid] = b
        return b

    def SetDestructionListener(self, listener):
        """b2world.d:44-48; listener.SayGoodbye(joint | fixture) (b2worldcallbacks.d:34-49)"""
        self._destruction = listener

    def DestroyBody(self, body):
        """b2world.d:105-191: the joints attached to the body (newest first, like its joint list) and its fixtures go with it,
        each announced to the destruction listener first.  Host-side bookkeeping of the shim: the library destroys the same
        objects inside dbx_body_destroy."""
        lst = getattr(self, "_destruction", None)
        attached = [j for j in self._joints.values() if body in (j.bodyA, j.bodyB) or body in getattr(j, "extra_bodies", ())]
        for j in sorted(attached, key=lambda j: -j.id):
            if lst is not None:
                lst.SayGoodbye(j)
            self._joints.pop(j.id, None)
        for f in reversed(body.fixtures):
            if lst is not None:
                lst.SayGoodbye(f)
            self._fixtures.pop(f.id, None)
        self._ck(self._api.body_destroy(self._w, body.id))
        self._bodies.pop(body.id, None)

    def CreateJoint(self, jointDef):
        pod = jointDef._pod()
        jid = self._ck(self._api.joint_create(self._w, C.byref(pod)))
        if jointDef.bodyA is None and getattr(jointDef, "joint1", None) is not None:     # gear: bodies come from the two joints
            jointDef.bodyA, jointDef.bodyB = jointDef.joint1.bodyB, jointDef.joint2.bodyB
        j = b2Joint(self, jid, jointDef.bodyA, jointDef.bodyB)
        j.type, j.collideConnected = jointDef.type, bool(jointDef.collideConnected)
        j._pod = pod
        self._joints[jid] = j
        return j

    def SetMotorSpeeds(self, joints, speeds):
        """bulk SetMotorSpeed (dbx_world_set_motor_speeds): joints = b2Joint handles or (replicated worlds) global joint indices"""
        ids = [j if isinstance(j, int) else j.id for j in joints]
        n = len(ids)
        a = (C.c_int32 * max(n, 1))(*ids)
        v = (C.c_float * max(n, 1))(*speeds)
        self._ck(self._api.world_set_motor_speeds(self._w, a, v, n))

    def SetMouseTarget(self, joint, target):
        """b2MouseJoint.SetTarget (b2mousejoint.d:112-120)"""
        x, y = target
        self._ck(self._api.joint_set_target(self._w, joint.id, x, y))

    def DestroyJoint(self, joint):
        self._ck(self._api.joint_destroy(self._w, joint.id))
        self._joints.pop(joint.id, None)

    # user contact filter (b2world.d:52-56), deferred: see include/dbox_b200.h "user contact filter" ---------------
    def SetContactFilter(self, contact_filter, replaces_default=True):
        """contact_filter: a b2ContactFilter (None = back to the default rule).  replaces_default: the device leaves the
        category / mask / group test to the filter object (whose base class implements it)."""
        self._filter = contact_filter
        if self._api.prefix == "dbx_":
            mode = 0 if contact_filter is None else (A.FILTER_LOG | (A.FILTER_REPLACES_DEFAULT if replaces_default else 0))
            self._ck(self._api.world_set_user_filter(self._w, mode))
        else:       # the oracle calls the filter where the reference does
            if contact_filter is None:
                self._filter_cb = None
                self._ck(self._api.world_set_contact_filter(self._w, None))
            else:
                proto = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int)
                self._filter_cb = proto(lambda fa, fb, default: int(bool(self._filter.ShouldCollide(self._fixtures[fa], self._fixtures[fb]))))
                self._ck(self._api.world_set_contact_filter(self._w, C.cast(self._filter_cb, C.c_void_p)))

    def _veto(self, pairs):
        if not pairs:
            return
        arr = (A.ContactPatch * len(pairs))()
        for k, (fa, ca, fb, cb) in enumerate(pairs):
            arr[k].fixtureA, arr[k].childA, arr[k].fixtureB, arr[k].childB, arr[k].mask = fa, ca, fb, cb, A.PATCH_DESTROY
        self._ck(self._api.world_patch_contacts(self._w, arr, len(pairs)))

    def _apply_user_filter(self):
        """ask the user's filter about every contact the broadphase created since the last poll; destroy what it rejects"""
        if getattr(self, "_filter", None) is None or self._api.prefix != "dbx_":
            return
        n = self._ck(self._api.world_poll_new_contacts(self._w, None, 0))
        if n == 0:
            return
        buf = (C.c_int32 * (4 * n))()
        n = min(n, self._ck(self._api.world_poll_new_contacts(self._w, buf, n)))
        self._veto([tuple(buf[4 * k:4 * k + 4]) for k in range(n)
                    if not self._filter.ShouldCollide(self._fixtures[buf[4 * k]], self._fixtures[buf[4 * k + 2]])])

    def _refilter_user(self, fixture):
        """b2Fixture.Refilter (b2fixture.d:140-178) flags the fixture's contacts; the reference's next Collide asks the filter
        about each (b2contactmanager.d:264-284).  Deferred form: ask now, destroy what is rejected."""
        if getattr(self, "_filter", None) is None or self._api.prefix != "dbx_":
            return
        recs, n = self.read_contacts()
        self._veto([(recs[i].fixtureA, recs[i].childA, recs[i].fixtureB, recs[i].childB) for i in range(n)
                    if fixture.id in (recs[i].fixtureA, recs[i].fixtureB)
                    and not self._filter.ShouldCollide(self._fixtures[recs[i].fixtureA], self._fixtures[recs[i].fixtureB])])

    # stepping ----------------------------------------------------------------------------------------------
    def Step(self, dt, velocityIterations, positionIterations):
        self._apply_user_filter()          # pairs of fixtures created since the last step (b2world.d:372-376 runs the broadphase first)
        self._ck(self._api.world_step(self._w, dt, velocityIterations, positionIterations))
        self._apply_user_filter()          # pairs the step's FindNewContacts found
        if getattr(self, "_listener", None) is not None:
            self._deliver_contact_events()

    # PreSolve: the step cut after Collide (include/dbox_b200.h "PreSolve") ---------------------------------
    def StepWithPreSolve(self, dt, velocityIterations, positionIterations, pre_solve, toi_lookahead=False):
        """pre_solve(contact_rec) is called for every touching non-sensor contact after Collide and may return a dict with any
        of enabled / friction / restitution / tangentSpeed (what a b2ContactListener.PreSolve would set on the contact).
        SetEnabled(false) holds for the rest of the step, the TOI loop's re-evaluations of the contact included (where the
        reference would call PreSolve again).  toi_lookahead=True also asks about contacts that are not touching yet (empty
        manifold): a fast body can first touch INSIDE the TOI loop (b2world.d:1295), where nobody can be asked any more, and the
        answer given here is the one that loop applies."""
        self._ck(self._api.world_step_begin(self._w, dt, velocityIterations, positionIterations))
        recs, n = self.read_contacts()
        patches = []
        for i in range(n):
            r = recs[i]
            touching = bool(r.flags & A.CONTACT_TOUCHING) and r.manifold.pointCount > 0      # PreSolve skips sensors (b2contact.d:352)
            if not touching:
                fa, fb = self._fixtures.get(r.fixtureA), self._fixtures.get(r.fixtureB)
                if not toi_lookahead or (fa is not None and fa.IsSensor()) or (fb is not None and fb.IsSensor()):
                    continue
            ch = pre_solve(r)
            if not ch:
                continue
            p = A.ContactPatch()
            p.fixtureA, p.childA, p.fixtureB, p.childB = r.fixtureA, r.childA, r.fixtureB, r.childB
            for name, bit in (("enabled", A.PATCH_ENABLED), ("friction", A.PATCH_FRICTION), ("restitution", A.PATCH_RESTITUTION), ("tangentSpeed", A.PATCH_TANGENT_SPEED)):
                if name in ch:
                    p.mask |= bit
                    setattr(p, name, int(ch[name]) if name == "enabled" else float(ch[name]))
            patches.append(p)
        if patches:
            arr = (A.ContactPatch * len(patches))(*patches)
            self._ck(self._api.world_patch_contacts(self._w, arr, len(patches)))
        self._ck(self._api.world_step_end(self._w))
        if getattr(self, "_listener", None) is not None:
            self._deliver_contact_events()

    # world queries, batched (b2world.d:563-587; include/dbox_b200.h "world queries") --------------------------
    def RayCastClosest(self, rays):
        """rays: iterable of ((x1, y1), (x2, y2)); returns [(fixture_id | -1, child, fraction, (px, py), (nx, ny))]"""
        rays = list(rays)
        n = len(rays)
        buf = (A.Ray * max(n, 1))()
        for k, (a, b) in enumerate(rays):
            buf[k].p1, buf[k].p2 = _v(a), _v(b)
        out = (A.RayHit * max(n, 1))()
        self._ck(self._api.world_raycast_closest(self._w, buf, n, out))
        return [(out[k].fixture, out[k].child, out[k].fraction, (out[k].point.x, out[k].point.y), (out[k].normal.x, out[k].normal.y)) for k in range(n)]

    def QueryAABB(self, boxes, cap=64):
        """boxes: iterable of ((lox, loy), (hix, hiy)); returns one sorted list of (fixture_id, child) per box"""
        boxes = list(boxes)
        n = len(boxes)
        buf = (A.AABB * max(n, 1))()
        for k, (lo, hi) in enumerate(boxes):
            buf[k].lo, buf[k].hi = _v(lo), _v(hi)
        counts = (C.c_int32 * max(n, 1))()
        pairs = (C.c_int32 * (2 * max(n, 1) * cap))()
        self._ck(self._api.world_query_aabb(self._w, buf, n, cap, counts, pairs))
        out = []
        for k in range(n):
            if counts[k] > cap:
                raise RuntimeError("QueryAABB: %d hits, cap %d" % (counts[k], cap))
            out.append([(pairs[2 * (k * cap + i)], pairs[2 * (k * cap + i) + 1]) for i in range(counts[k])])
        return out

    def RayCastAll(self, rays, cap=64):
        """the RayCast callback that returns 1: per ray every (fixture_id, child, fraction, (px, py), (nx, ny)) hit, nearest first"""
        rays = list(rays)
        n = len(rays)
        buf = (A.Ray * max(n, 1))()
        for k, (a, b) in enumerate(rays):
            buf[k].p1, buf[k].p2 = _v(a), _v(b)
        counts = (C.c_int32 * max(n, 1))()
        hits = (A.RayHit * (max(n, 1) * cap))()
        self._ck(self._api.world_raycast_all(self._w, buf, n, cap, counts, hits))
        out = []
        for k in range(n):
            if counts[k] > cap:
                raise RuntimeError("RayCastAll: %d hits, cap %d" % (counts[k], cap))
            out.append([(h.fixture, h.child, h.fraction, (h.point.x, h.point.y), (h.normal.x, h.normal.y)) for h in hits[k * cap:k * cap + counts[k]]])
        return out

    def TestPoints(self, fixture_points):
        """b2Fixture.TestPoint for a batch of (fixture | fixture id, (x, y)) pairs -> [bool]"""
        fp = list(fixture_points)
        n = len(fp)
        ids = (C.c_int32 * max(n, 1))(*[f if isinstance(f, int) else f.id for f, _ in fp])
        pts = (A.Vec2 * max(n, 1))(*[_v(p) for _, p in fp])
        out = (C.c_int32 * max(n, 1))()
        self._ck(self._api.world_test_points(self._w, ids, pts, n, out))
        return [bool(out[k]) for k in range(n)]

    def ShiftOrigin(self, newOrigin):
        """b2world.d:758-780"""
        x, y = newOrigin
        self._ck(self._api.world_shift_origin(self._w, x, y))

    def GetWorldManifolds(self):
        """b2Contact.GetWorldManifold for every contact, in read_contacts order: [(pointCount, normal, points, separations)]"""
        n = self._ck(self._api.world_read_world_manifolds(self._w, None, 0))
        buf = (A.WorldManifold * max(n, 1))()
        n = self._ck(self._api.world_read_world_manifolds(self._w, buf, n))
        return [(m.pointCount, (m.normal.x, m.normal.y), [(m.points[i].x, m.points[i].y) for i in range(2)], [m.separations[i] for i in range(2)]) for m in buf[:n]]

    # contact listener (b2world.d:62-66, b2worldcallbacks.d:87-128), deferred: see include/dbox_b200.h ------------
    def SetContactListener(self, listener, capacity=1 << 16):
        self._listener = listener
        self._ck(self._api.world_enable_contact_events(self._w, capacity if listener is not None else 0))
        self._ck(self._api.world_enable_post_solve(self._w, capacity if (listener is not None and listener.post_solve) else 0))

    def EnablePostSolve(self, capacity=1 << 16):
        return self._ck(self._api.world_enable_post_solve(self._w, capacity))

    def ReadPostSolve(self):
        """PostSolve records of the last step: (phase, fixtureA, childA, fixtureB, childB, count, normalImpulses, tangentImpulses)"""
        n = self._ck(self._api.world_read_post_solve(self._w, None, 0))
        if n == 0:
            return []
        buf = (A.PostSolve * n)()
        n = self._ck(self._api.world_read_post_solve(self._w, buf, n))
        return [(r.phase, r.fixtureA, r.childA, r.fixtureB, r.childB, r.count, tuple(r.normalImpulses), tuple(r.tangentImpulses)) for r in buf[:n]]

    def EnableContactEvents(self, capacity=1 << 16):
        return self._ck(self._api.world_enable_contact_events(self._w, capacity))

    def PollContactEvents(self):
        """events since the last poll as (type, phase, stepsAgo, fixtureA, fixtureB, childA, childB, bodyA, bodyB) tuples"""
        n = self._ck(self._api.world_poll_contact_events(self._w, None, 0))
        if n == 0:
            return []
        buf = (A.ContactEvent * n)()
        n = self._ck(self._api.world_poll_contact_events(self._w, buf, n))
        return [tuple(getattr(buf[i], f) for f, _ in A.ContactEvent._fields_) for i in range(n)]

    def _deliver_contact_events(self):
        for ev in self.PollContactEvents():
            c = b2ContactView(self, ev)
            if ev[0] == A.CONTACT_BEGIN:
                self._listener.BeginContact(c)
            else:
                self._listener.EndContact(c)
        if self._listener.post_solve:
            for r in self.ReadPostSolve():
                c = b2ContactView(self, (0, r[0], 0, r[1], r[3], r[2], r[4], -1, -1))
                self._listener.PostSolve(c, (r[5], r[6], r[7]))

    def Replicate(self, copies):
        """dbx_world_replicate: this world's content becomes replica 0 of `copies` independent replicas stepped together
        (BASELINE config 5, batched RL-style worlds).  Body id of replica r = r * bodies_per_replica + local id."""
        self._ck(self._api.world_replicate(self._w, copies))

    def SetBodyStates(self, ids=None, pose=None, vel=None):
        """bulk b2Body.SetTransform / SetLinearVelocity / SetAngularVelocity (b2body.d:261-326) from numpy arrays:
        pose[n,4] = (x, y, angle, -), vel[n,4] = (vx, vy, w, -), ids[n] int32 (None = bodies 0..n-1); also valid on a
        replicated world, where body r * bodies_per_replica + b is body b of replica r"""
        import numpy as np
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (pose, vel)]
        n = next(a.shape[0] for a in arrs if a is not None)
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32); n = ids.shape[0]; idp = ids.ctypes.data
        for a in arrs:
            assert a is None or a.shape == (n, 4)
        self._ck(self._api.world_set_body_states(self._w, idp, None if arrs[0] is None else arrs[0].ctypes.data,
                                                 None if arrs[1] is None else arrs[1].ctypes.data, n))

    def StepN(self, dt, velocityIterations, positionIterations, n):
        self._ck(self._api.world_step_n(self._w, dt, velocityIterations, positionIterations, n))

    def _set_flag(self, bit, on):
        self._flags = (self._flags | bit) if on else (self._flags & ~bit)
        self._ck(self._api.world_set_flags(self._w, self._flags))

    def SetAllowSleeping(self, flag):
        self._set_flag(A.WORLD_ALLOW_SLEEP, flag)

    def SetWarmStarting(self, flag):
        self._set_flag(A.WORLD_WARM_STARTING, flag)

    def SetContinuousPhysics(self, flag):
        self._set_flag(A.WORLD_CONTINUOUS, flag)

    def SetSubStepping(self, flag):
        self._set_flag(A.WORLD_SUB_STEPPING, flag)

    def SetAutoClearForces(self, flag):
        self._set_flag(A.WORLD_AUTO_CLEAR_FORCES, flag)

    def SetGravity(self, g):
        gx, gy = g
        self._ck(self._api.world_set_gravity(self._w, gx, gy))

    # counts / profile (b2world.d:677-716, 789-792) ------------------------------------------------------------
    def counts(self):
        c = A.Counts()
        self._ck(self._api.world_counts(self._w, C.byref(c)))
        return c

    def GetTreeStats(self):
        """(GetTreeHeight, GetTreeBalance, GetTreeQuality) of the broadphase tree (b2world.d:694-716); the CUDA library reports its LBVH"""
        if self._api.prefix != "dbx_":
            return (self._api.world_tree_height(self._w), None, None)
        h, b, q = C.c_int32(), C.c_int32(), C.c_float()
        self._ck(self._api.world_tree_stats(self._w, C.byref(h), C.byref(b), C.byref(q)))
        return (h.value, b.value, q.value)

    def GetBodyCount(self):
        return self.counts().bodies

    def GetContactCount(self):
        return self.counts().contacts

    def GetJointCount(self):
        return self.counts().joints

    def GetProxyCount(self):
        return self.counts().proxies

    def GetProfile(self):
        p = A.Profile()
        self._ck(self._api.world_profile(self._w, C.byref(p)))
        return p

    # bulk state ----------------------------------------------------------------------------------------------
    def _read(self, fn, typ, n_hint=0):
        n = fn(self._w, None, 0) if n_hint <= 0 else n_hint
        self._ck(n)
        buf = (typ * max(n, 1))()
        n2 = self._ck(fn(self._w, buf, n))
        return buf, min(n, n2)

    def read_bodies(self):
        return self._read(self._api.world_read_bodies, A.BodyState)

    def read_contacts(self):
        return self._read(self._api.world_read_contacts, A.ContactRec)

    def read_proxies(self):
        return self._read(self._api.world_read_proxies, A.ProxyRec)

    def read_joints(self):
        return self._read(self._api.world_read_joints, A.JointState)

    def read_moves(self):
        n = self._ck(self._api.world_read_moves(self._w, None, 0))
        buf = (C.c_int32 * max(2 * n, 2))()
        n = self._ck(self._api.world_read_moves(self._w, buf, n))
        return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]

    def read_pairs(self):
        n = self._ck(self._api.world_read_pairs(self._w, None, 0))
        buf = (C.c_int32 * max(4 * n, 4))()
        n = self._ck(self._api.world_read_pairs(self._w, buf, n))
        return [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]

    def get_inv_dt0(self):
        f = C.c_float()
        self._ck(self._api.world_get_inv_dt0(self._w, C.byref(f)))
        return f.value
