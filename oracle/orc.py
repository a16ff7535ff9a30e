"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liborc.so (the CPU restatement of dbox).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
PARITY UNPINNED: the reference has no golden vectors and cannot be compiled here (no D toolchain); see DESIGN.md.
"""
import ctypes as C
import os
import subprocess

from dbox_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborc.so")

# the oracle implements the subset of the ABI that has a CPU meaning
_REQUIRED = [
    "shape_set_circle", "shape_set_edge", "shape_set_box", "shape_set_box_at", "shape_set_polygon", "shape_set_chain",
    "world_destroy", "world_set_flags", "world_set_gravity", "body_create", "body_destroy", "fixture_create",
    "fixture_destroy", "joint_create", "joint_destroy", "world_step", "world_step_n", "body_get_state",
    "body_set_transform", "body_set_linear_velocity", "body_set_angular_velocity", "body_apply_force", "body_apply_torque",
    "body_apply_linear_impulse", "body_apply_angular_impulse", "body_set_awake", "body_set_bullet",
    "body_set_sleeping_allowed", "world_counts", "world_profile", "world_read_bodies", "world_read_contacts",
    "world_read_proxies", "world_read_joints", "world_read_moves", "world_get_inv_dt0", "world_stage_find_new_contacts",
    "world_stage_collide", "world_read_pairs",
    # state import: the mirror of dbx_world_write_* (transplant a device-resident world into the oracle)
    "world_write_bodies", "world_write_proxies", "world_write_contacts", "world_write_joints", "world_write_moves", "world_set_inv_dt0",
]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    srcs.append(os.path.join(_HERE, "..", "include", "dbox_b200.h"))
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liborc.so"])
    return _LIB


_api = None


def api():
    global _api
    if _api is None:
        lib = C.CDLL(build())
        a = A.Api(lib, "orc_", required=_REQUIRED)
        P, W, i32, f32 = C.POINTER, C.c_void_p, C.c_int32, C.c_float
        lib.orc_world_create.restype, lib.orc_world_create.argtypes = W, [f32, f32]
        a.world_create = lib.orc_world_create
        extra = {
            "shape_mass": (None, [P(A.Shape), f32, P(f32), P(A.Vec2), P(f32)]),
            "shape_aabb": (None, [P(A.Shape), f32, f32, f32, i32, P(A.AABB)]),
            "world_read_solve_order": (i32, [W, P(i32), i32]),
            "world_read_joint_solve_order": (i32, [W, P(i32), i32]),
            "world_tree_height": (i32, [W]),
            "world_tree_validate": (i32, [W]),
            "collide": (i32, [P(A.Shape), f32, f32, f32, i32, P(A.Shape), f32, f32, f32, i32, P(A.Manifold)]),
            "collide_xf": (i32, [P(A.Shape), P(f32), i32, P(A.Shape), P(f32), i32, P(A.Manifold)]),
            "distance": (f32, [P(A.Shape), f32, f32, f32, i32, P(A.Shape), f32, f32, f32, i32, i32, P(A.Vec2), P(A.Vec2), P(i32)]),
            "time_of_impact": (i32, [P(A.Shape), P(f32), i32, P(A.Shape), P(f32), i32, f32, P(f32)]),
            "batch_step": (C.c_double, [P(W), i32, f32, i32, i32, i32, i32]),
            "world_step_begin": (i32, [W, f32, i32, i32]),
            "world_step_end": (i32, [W]),
            "world_patch_contacts": (i32, [W, P(A.ContactPatch), i32]),
            "world_raycast_closest": (i32, [W, P(A.Ray), i32, P(A.RayHit)]),
            "world_query_aabb": (i32, [W, P(A.AABB), i32, i32, P(i32), P(i32)]),
            "joint_set_target": (i32, [W, i32, f32, f32]),
            "body_set_type": (i32, [W, i32, i32]),
            "body_set_mass_data": (i32, [W, i32, f32, f32, f32, f32]), "body_reset_mass_data": (i32, [W, i32]),
            "body_set_fixed_rotation": (i32, [W, i32, i32]), "body_set_linear_damping": (i32, [W, i32, f32]),
            "body_set_angular_damping": (i32, [W, i32, f32]), "body_set_gravity_scale": (i32, [W, i32, f32]),
            "fixture_set_filter": (i32, [W, i32, i32, i32, i32]), "fixture_set_sensor": (i32, [W, i32, i32]),
            "fixture_set_friction": (i32, [W, i32, f32]), "fixture_set_restitution": (i32, [W, i32, f32]), "fixture_set_density": (i32, [W, i32, f32]),
            "body_set_active": (i32, [W, i32, i32]),
            "world_enable_contact_events": (i32, [W, i32]),
            "world_poll_contact_events": (i32, [W, P(A.ContactEvent), i32]),
            "world_test_points": (i32, [W, P(i32), P(A.Vec2), i32, P(i32)]),
            "world_raycast_all": (i32, [W, P(A.Ray), i32, i32, P(i32), P(A.RayHit)]),
            "world_shift_origin": (i32, [W, f32, f32]),
            "world_read_world_manifolds": (i32, [W, P(A.WorldManifold), i32]),
            "world_enable_post_solve": (i32, [W, i32]),
            "world_read_post_solve": (i32, [W, P(A.PostSolve), i32]),
            "world_set_contact_filter": (i32, [W, C.c_void_p]),
            "joint_set_params": (i32, [W, i32, P(A.JointDef), C.c_uint32]),
            "world_set_motor_speeds": (i32, [W, P(i32), P(f32), i32]),
            "world_debug_set_solve_order": (i32, [W, P(i32), P(i32), i32, P(i32), i32, i32]),
        }
        for name, (res, args) in extra.items():
            fn = getattr(lib, "orc_" + name)
            fn.restype, fn.argtypes = res, args
            setattr(a, name, fn)
        _api = a
    return _api
