"""ctypes view of include/dbox_b200.h (the C-ABI drop-in boundary).

The struct layouts here must match the header byte for byte; tests/test_abi.py checks sizes and that the
shared library exports every declared symbol.  `bind(lib, prefix)` attaches prototypes for a library that
exports the ABI under `prefix` ("dbx_" for the product; the CPU oracle used by the tests exports the same
shape under "orc_").
"""
import ctypes as C

c_i32, c_u32, c_f32, c_u64 = C.c_int32, C.c_uint32, C.c_float, C.c_uint64

DBX_OK, DBX_E_INVALID, DBX_E_LOCKED, DBX_E_NO_DEVICE, DBX_E_CUDA, DBX_E_CAPACITY, DBX_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
STATIC_BODY, KINEMATIC_BODY, DYNAMIC_BODY = 0, 1, 2
BODY_ISLAND, BODY_AWAKE, BODY_AUTOSLEEP, BODY_BULLET, BODY_FIXED_ROTATION, BODY_ACTIVE, BODY_TOI = 1, 2, 4, 8, 0x10, 0x20, 0x40
CONTACT_ISLAND, CONTACT_TOUCHING, CONTACT_ENABLED, CONTACT_FILTER, CONTACT_BULLET_HIT, CONTACT_TOI = 1, 2, 4, 8, 0x10, 0x20
SHAPE_CIRCLE, SHAPE_EDGE, SHAPE_POLYGON, SHAPE_CHAIN = 0, 1, 2, 3
MANIFOLD_CIRCLES, MANIFOLD_FACE_A, MANIFOLD_FACE_B = 0, 1, 2
JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_DISTANCE, JOINT_PULLEY, JOINT_MOUSE, JOINT_GEAR = 1, 2, 3, 4, 5, 6
JOINT_WHEEL, JOINT_WELD, JOINT_FRICTION, JOINT_ROPE, JOINT_MOTOR = 7, 8, 9, 10, 11
WORLD_ALLOW_SLEEP, WORLD_WARM_STARTING, WORLD_CONTINUOUS, WORLD_SUB_STEPPING, WORLD_AUTO_CLEAR_FORCES = 1, 2, 4, 8, 0x10
IO_FULL, IO_COMPACT = 0, 1      # dbx_world_set_io_format: 16-byte or 12-byte bulk I/O records
WORLD_DEFAULT_FLAGS = 0x17


class Vec2(C.Structure):
    _fields_ = [("x", c_f32), ("y", c_f32)]

    def __init__(self, x=0.0, y=0.0):
        super().__init__(x, y)

    def t(self):
        return (self.x, self.y)


class AABB(C.Structure):
    _fields_ = [("lo", Vec2), ("hi", Vec2)]


class BodyDef(C.Structure):
    _fields_ = [("type", c_i32), ("position", Vec2), ("angle", c_f32), ("linearVelocity", Vec2), ("angularVelocity", c_f32),
                ("linearDamping", c_f32), ("angularDamping", c_f32), ("allowSleep", c_i32), ("awake", c_i32),
                ("fixedRotation", c_i32), ("bullet", c_i32), ("active", c_i32), ("gravityScale", c_f32), ("userData", c_u64)]


class Shape(C.Structure):
    _fields_ = [("type", c_i32), ("radius", c_f32), ("p", Vec2), ("v0", Vec2), ("v1", Vec2), ("v2", Vec2), ("v3", Vec2),
                ("hasV0", c_i32), ("hasV3", c_i32), ("centroid", Vec2), ("vertices", Vec2 * 8), ("normals", Vec2 * 8),
                ("count", c_i32), ("chainVertices", C.POINTER(Vec2)), ("chainCount", c_i32), ("prevVertex", Vec2),
                ("nextVertex", Vec2), ("hasPrev", c_i32), ("hasNext", c_i32)]


class FixtureDef(C.Structure):
    _fields_ = [("friction", c_f32), ("restitution", c_f32), ("density", c_f32), ("isSensor", c_i32),
                ("categoryBits", C.c_uint16), ("maskBits", C.c_uint16), ("groupIndex", C.c_int16), ("_pad", C.c_uint16),
                ("userData", c_u64)]


class JointDef(C.Structure):
    _fields_ = [("type", c_i32), ("bodyA", c_i32), ("bodyB", c_i32), ("collideConnected", c_i32),
                ("localAnchorA", Vec2), ("localAnchorB", Vec2), ("referenceAngle", c_f32), ("enableLimit", c_i32),
                ("lowerAngle", c_f32), ("upperAngle", c_f32), ("enableMotor", c_i32), ("motorSpeed", c_f32),
                ("maxMotorTorque", c_f32), ("length", c_f32), ("frequencyHz", c_f32), ("dampingRatio", c_f32),
                ("userData", c_u64),
                ("localAxisA", Vec2), ("lowerTranslation", c_f32), ("upperTranslation", c_f32), ("maxMotorForce", c_f32),
                ("maxLength", c_f32), ("maxForce", c_f32), ("maxTorque", c_f32), ("linearOffset", Vec2),
                ("angularOffset", c_f32), ("correctionFactor", c_f32), ("target", Vec2), ("groundAnchorA", Vec2),
                ("groundAnchorB", Vec2), ("lengthA", c_f32), ("lengthB", c_f32), ("ratio", c_f32), ("joint1", c_i32),
                ("joint2", c_i32), ("_pad", c_i32)]


class BodyState(C.Structure):
    _fields_ = [("type", c_i32), ("flags", c_u32), ("p", Vec2), ("qs", c_f32), ("qc", c_f32), ("localCenter", Vec2),
                ("c0", Vec2), ("c", Vec2), ("a0", c_f32), ("a", c_f32), ("alpha0", c_f32), ("v", Vec2), ("w", c_f32),
                ("force", Vec2), ("torque", c_f32), ("mass", c_f32), ("invMass", c_f32), ("I", c_f32), ("invI", c_f32),
                ("linearDamping", c_f32), ("angularDamping", c_f32), ("gravityScale", c_f32), ("sleepTime", c_f32)]


class ManifoldPoint(C.Structure):
    _fields_ = [("localPoint", Vec2), ("normalImpulse", c_f32), ("tangentImpulse", c_f32), ("key", c_u32)]


class Manifold(C.Structure):
    _fields_ = [("points", ManifoldPoint * 2), ("localNormal", Vec2), ("localPoint", Vec2), ("type", c_i32), ("pointCount", c_i32)]


class ContactRec(C.Structure):
    _fields_ = [("fixtureA", c_i32), ("fixtureB", c_i32), ("childA", c_i32), ("childB", c_i32), ("flags", c_u32),
                ("manifold", Manifold), ("friction", c_f32), ("restitution", c_f32), ("tangentSpeed", c_f32),
                ("toiCount", c_i32), ("toi", c_f32)]


class ProxyRec(C.Structure):
    _fields_ = [("fixture", c_i32), ("child", c_i32), ("proxyId", c_i32), ("aabb", AABB), ("fat", AABB)]


class ContactPatch(C.Structure):
    _fields_ = [("fixtureA", c_i32), ("childA", c_i32), ("fixtureB", c_i32), ("childB", c_i32), ("mask", c_i32), ("enabled", c_i32),
                ("friction", c_f32), ("restitution", c_f32), ("tangentSpeed", c_f32)]


PATCH_ENABLED, PATCH_FRICTION, PATCH_RESTITUTION, PATCH_TANGENT_SPEED, PATCH_DESTROY = 1, 2, 4, 8, 16
FILTER_LOG, FILTER_REPLACES_DEFAULT = 1, 2
JP_MOTOR_SPEED, JP_MAX_MOTOR, JP_ENABLE_MOTOR, JP_ENABLE_LIMIT, JP_LIMITS, JP_SPRING, JP_LENGTH, JP_MAX_FORCE, JP_OFFSETS, JP_CORRECTION = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512


class Ray(C.Structure):
    _fields_ = [("p1", Vec2), ("p2", Vec2)]


class RayHit(C.Structure):
    _fields_ = [("fixture", c_i32), ("child", c_i32), ("fraction", c_f32), ("point", Vec2), ("normal", Vec2)]


class ContactEvent(C.Structure):
    _fields_ = [(n, c_i32) for n in ("type", "phase", "stepsAgo", "fixtureA", "fixtureB", "childA", "childB", "bodyA", "bodyB")]


CONTACT_BEGIN, CONTACT_END = 1, 2


class WorldManifold(C.Structure):
    _fields_ = [("normal", Vec2), ("points", Vec2 * 2), ("separations", c_f32 * 2), ("pointCount", c_i32), ("_pad", c_i32)]


class PostSolve(C.Structure):
    _fields_ = [("fixtureA", c_i32), ("fixtureB", c_i32), ("childA", c_i32), ("childB", c_i32), ("phase", c_i32), ("count", c_i32),
                ("normalImpulses", c_f32 * 2), ("tangentImpulses", c_f32 * 2)]


class JointState(C.Structure):
    _fields_ = [("type", c_i32), ("impulse", c_f32 * 3), ("motorImpulse", c_f32), ("limitState", c_i32)]


class Counts(C.Structure):
    _fields_ = [(n, c_i32) for n in ("bodies", "fixtures", "proxies", "contacts", "touching", "joints", "awakeBodies",
                                     "colours", "islands", "moves", "pairs")]


class Profile(C.Structure):
    _fields_ = [(n, c_f32) for n in ("step", "collide", "solve", "solveInit", "solveVelocity", "solvePosition", "broadphase", "solveTOI")]


class Caps(C.Structure):
    _fields_ = [(n, c_i32) for n in ("maxBodies", "maxProxies", "maxContacts", "maxJoints", "maxPairs")]


# sizes the header implies (checked by tests/test_abi.py and by the library's own static_asserts)
EXPECTED_SIZES = {"Vec2": 8, "AABB": 16, "BodyDef": 72, "Shape": 240, "FixtureDef": 32, "JointDef": 176, "BodyState": 116,
                  "ManifoldPoint": 20, "Manifold": 64, "ContactRec": 104, "ProxyRec": 44, "JointState": 24, "Counts": 44,
                  "Profile": 32, "Caps": 20, "ContactEvent": 36, "Ray": 16, "RayHit": 28, "ContactPatch": 36,
                  "WorldManifold": 40, "PostSolve": 40}

P = C.POINTER
W = C.c_void_p

# name -> (restype, argtypes); every function include/dbox_b200.h declares.  `W` first arg = world handle.
PROTOTYPES = {
    "abi_version": (c_i32, []),
    "last_error": (C.c_char_p, []),
    "device_count": (c_i32, []),
    "default_body_def": (None, [P(BodyDef)]),
    "default_fixture_def": (None, [P(FixtureDef)]),
    "default_joint_def": (None, [P(JointDef), c_i32]),
    "shape_set_circle": (None, [P(Shape), c_f32, c_f32, c_f32]),
    "shape_set_edge": (None, [P(Shape), Vec2, Vec2]),
    "shape_set_box": (None, [P(Shape), c_f32, c_f32]),
    "shape_set_box_at": (None, [P(Shape), c_f32, c_f32, Vec2, c_f32]),
    "shape_set_polygon": (c_i32, [P(Shape), P(Vec2), c_i32]),
    "shape_set_chain": (None, [P(Shape), P(Vec2), c_i32, c_i32]),
    "world_create": (W, [c_f32, c_f32, c_i32, P(Caps)]),
    "world_destroy": (None, [W]),
    "world_set_flags": (c_i32, [W, c_u32]),
    "world_get_flags": (c_u32, [W]),
    "world_set_gravity": (c_i32, [W, c_f32, c_f32]),
    "body_create": (c_i32, [W, P(BodyDef)]),
    "body_destroy": (c_i32, [W, c_i32]),
    "fixture_create": (c_i32, [W, c_i32, P(FixtureDef), P(Shape)]),
    "fixture_destroy": (c_i32, [W, c_i32]),
    "joint_create": (c_i32, [W, P(JointDef)]),
    "joint_destroy": (c_i32, [W, c_i32]),
    "joint_set_target": (c_i32, [W, c_i32, c_f32, c_f32]),
    "world_step": (c_i32, [W, c_f32, c_i32, c_i32]),
    "world_step_n": (c_i32, [W, c_f32, c_i32, c_i32, c_i32]),
    "world_clear_forces": (c_i32, [W]),
    "world_time_steps": (c_i32, [W, c_f32, c_i32, c_i32, c_i32, c_i32, P(c_f32), P(c_f32)]),
    "world_apply_forces": (c_i32, [W, C.c_void_p, c_i32]),
    "world_set_body_states": (c_i32, [W, C.c_void_p, C.c_void_p, C.c_void_p, c_i32]),
    "world_read_transforms": (c_i32, [W, C.c_void_p, c_i32]),
    "world_launch_count": (C.c_int64, [W]),
    "body_get_state": (c_i32, [W, c_i32, P(BodyState)]),
    "body_set_transform": (c_i32, [W, c_i32, c_f32, c_f32, c_f32]),
    "body_set_linear_velocity": (c_i32, [W, c_i32, c_f32, c_f32]),
    "body_set_angular_velocity": (c_i32, [W, c_i32, c_f32]),
    "body_apply_force": (c_i32, [W, c_i32, c_f32, c_f32, c_f32, c_f32, c_i32]),
    "body_apply_torque": (c_i32, [W, c_i32, c_f32, c_i32]),
    "body_apply_linear_impulse": (c_i32, [W, c_i32, c_f32, c_f32, c_f32, c_f32, c_i32]),
    "body_apply_angular_impulse": (c_i32, [W, c_i32, c_f32, c_i32]),
    "body_set_awake": (c_i32, [W, c_i32, c_i32]),
    "body_set_bullet": (c_i32, [W, c_i32, c_i32]),
    "body_set_sleeping_allowed": (c_i32, [W, c_i32, c_i32]),
    "body_set_mass_data": (c_i32, [W, c_i32, c_f32, c_f32, c_f32, c_f32]),
    "body_reset_mass_data": (c_i32, [W, c_i32]),
    "body_set_fixed_rotation": (c_i32, [W, c_i32, c_i32]),
    "body_set_linear_damping": (c_i32, [W, c_i32, c_f32]),
    "body_set_angular_damping": (c_i32, [W, c_i32, c_f32]),
    "body_set_gravity_scale": (c_i32, [W, c_i32, c_f32]),
    "fixture_set_filter": (c_i32, [W, c_i32, c_i32, c_i32, c_i32]),
    "fixture_set_sensor": (c_i32, [W, c_i32, c_i32]),
    "fixture_set_friction": (c_i32, [W, c_i32, c_f32]),
    "fixture_set_restitution": (c_i32, [W, c_i32, c_f32]),
    "fixture_set_density": (c_i32, [W, c_i32, c_f32]),
    "body_set_type": (c_i32, [W, c_i32, c_i32]),
    "body_set_active": (c_i32, [W, c_i32, c_i32]),
    "world_counts": (c_i32, [W, P(Counts)]),
    "world_profile": (c_i32, [W, P(Profile)]),
    "world_read_bodies": (c_i32, [W, P(BodyState), c_i32]),
    "world_write_bodies": (c_i32, [W, P(BodyState), c_i32]),
    "world_read_contacts": (c_i32, [W, P(ContactRec), c_i32]),
    "world_write_contacts": (c_i32, [W, P(ContactRec), c_i32]),
    "world_read_proxies": (c_i32, [W, P(ProxyRec), c_i32]),
    "world_write_proxies": (c_i32, [W, P(ProxyRec), c_i32]),
    "world_read_joints": (c_i32, [W, P(JointState), c_i32]),
    "world_write_joints": (c_i32, [W, P(JointState), c_i32]),
    "world_read_moves": (c_i32, [W, P(c_i32), c_i32]),
    "world_write_moves": (c_i32, [W, P(c_i32), c_i32]),
    "world_get_inv_dt0": (c_i32, [W, P(c_f32)]),
    "world_set_inv_dt0": (c_i32, [W, c_f32]),
    "world_stage_find_new_contacts": (c_i32, [W]),
    "world_stage_collide": (c_i32, [W]),
    "world_read_pairs": (c_i32, [W, P(c_i32), c_i32]),
    "world_debug_set_contact_levels": (c_i32, [W, P(c_i32), c_i32]),
    "world_debug_read_solve_order": (c_i32, [W, P(c_i32), c_i32, P(c_i32), c_i32, P(c_i32)]),
    "world_debug_colour_conflicts": (c_i32, [W]),
    "debug_collide": (c_i32, [c_i32, c_i32, P(Shape), P(c_f32), P(Shape), P(c_f32), P(Manifold)]),
    "debug_distance": (c_i32, [c_i32, c_i32, P(Shape), P(c_f32), P(Shape), P(c_f32), c_i32, P(c_f32), P(Vec2), P(Vec2), P(c_i32)]),
    "debug_time_of_impact": (c_i32, [c_i32, c_i32, P(Shape), P(c_f32), P(Shape), P(c_f32), c_f32, P(c_i32), P(c_f32)]),
    "world_debug_phase_times": (c_i32, [W, P(C.c_uint64), c_i32]),
    "world_debug_header": (c_i32, [W, C.c_void_p, c_i32]),
    "debug_barrier_us": (c_f32, [c_i32, c_i32, c_i32, c_i32]),
    "stats_allreduce": (c_i32, [C.c_void_p, C.c_void_p, P(C.c_double), c_i32, P(C.c_double), c_i32]),
    "world_replicate": (c_i32, [W, c_i32]),
    "world_replica_count": (c_i32, [W]),
    "world_export_state": (C.c_int64, [W, C.c_void_p, C.c_int64]),
    "world_import_state": (c_i32, [W, C.c_void_p, C.c_int64]),
    "world_step_begin": (c_i32, [W, c_f32, c_i32, c_i32]),
    "world_step_end": (c_i32, [W]),
    "world_patch_contacts": (c_i32, [W, P(ContactPatch), c_i32]),
    "world_raycast_closest": (c_i32, [W, P(Ray), c_i32, P(RayHit)]),
    "world_query_aabb": (c_i32, [W, P(AABB), c_i32, c_i32, P(c_i32), P(c_i32)]),
    "world_enable_contact_events": (c_i32, [W, c_i32]),
    "world_poll_contact_events": (c_i32, [W, P(ContactEvent), c_i32]),
    "world_test_points": (c_i32, [W, P(c_i32), P(Vec2), c_i32, P(c_i32)]),
    "world_raycast_all": (c_i32, [W, P(Ray), c_i32, c_i32, P(c_i32), P(RayHit)]),
    "world_shift_origin": (c_i32, [W, c_f32, c_f32]),
    "world_read_world_manifolds": (c_i32, [W, P(WorldManifold), c_i32]),
    "world_enable_post_solve": (c_i32, [W, c_i32]),
    "world_read_post_solve": (c_i32, [W, P(PostSolve), c_i32]),
    "world_set_user_filter": (c_i32, [W, c_i32]),
    "world_tree_stats": (c_i32, [W, P(c_i32), P(c_i32), P(c_f32)]),
    "joint_set_params": (c_i32, [W, c_i32, P(JointDef), c_u32]),
    "world_set_motor_speeds": (c_i32, [W, P(c_i32), P(c_f32), c_i32]),
    "world_step_async": (c_i32, [W, c_f32, c_i32, c_i32]),
    "world_apply_forces_async": (c_i32, [W, C.c_void_p, c_i32]),
    "world_read_transforms_async": (c_i32, [W, C.c_void_p, c_i32]),
    "world_io_wait": (c_i32, [W, c_i32]),
    "world_set_io_format": (c_i32, [W, c_i32]),
    "world_sync": (c_i32, [W]),
    "world_poll_new_contacts": (c_i32, [W, P(c_i32), c_i32]),
}


class Api:
    """Prototype-checked access to one shared library exporting the ABI under `prefix`."""

    def __init__(self, lib, prefix, required=None):
        self.lib, self.prefix = lib, prefix
        self.missing = []
        for name, (res, args) in PROTOTYPES.items():
            try:
                fn = getattr(lib, prefix + name)
            except AttributeError:
                self.missing.append(name)
                continue
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)
        need = PROTOTYPES.keys() if required is None else required
        lacking = [n for n in need if n in self.missing]
        if lacking:
            raise RuntimeError("library lacks ABI symbols: " + ", ".join(prefix + n for n in lacking))
