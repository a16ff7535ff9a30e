"""Helpers shared by the parity tests: move a world state from the CPU oracle into the GPU library, derive the
sequential-order level schedule, and compare states."""
import ctypes as C
import math

from dbox_b200 import _abi as A


def transplant(wo, wg):
    """Copy bodies, proxies (tight + fat AABBs), contacts (manifolds + impulses + flags), joints, the pending move
    buffer and inv_dt0 from world `wo` into world `wg` (same scene built on both); either direction between the oracle
    and the CUDA library."""
    from dbox_b200 import state
    return state.transplant(wo, wg)


def contact_key(r):
    return (r.fixtureA, r.childA, r.fixtureB, r.childB)


def contacts_by_key(world):
    recs, n = world.read_contacts()
    return {contact_key(recs[i]): recs[i] for i in range(n)}, recs, n


def sequential_levels(oracle_api, wo, wg, fixture_body, body_dynamic):
    """Levels for the GPU's next step such that running level by level is a topological re-ordering of the order in
    which the oracle solved its contacts in ITS last step (read after the oracle stepped)."""
    n = oracle_api.world_read_solve_order(wo._w, None, 0)
    buf = (C.c_int32 * max(4 * n, 4))()
    n = oracle_api.world_read_solve_order(wo._w, buf, n)
    order = [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]
    last = {}
    level_of = {}
    for k in order:
        bA, bB = fixture_body[k[0]], fixture_body[k[2]]
        lv = -1
        if body_dynamic[bA]:
            lv = max(lv, last.get(bA, -1))
        if body_dynamic[bB]:
            lv = max(lv, last.get(bB, -1))
        lv += 1
        level_of[k] = lv
        if body_dynamic[bA]:
            last[bA] = lv
        if body_dynamic[bB]:
            last[bB] = lv
    recs, m = wg.read_contacts()
    levels = (C.c_int32 * max(m, 1))()
    for i in range(m):
        levels[i] = level_of.get(contact_key(recs[i]), 0)
    return levels, m, (max(level_of.values()) + 1 if level_of else 0)


def rel_err(a, b, floor=1e-3):
    return abs(a - b) / max(abs(a), abs(b), floor)


def body_state_errors(so, sg, n, skip_static=True):
    """max relative error of positions / angles / velocities between two body-state arrays"""
    ep = ev = 0.0
    for i in range(n):
        if skip_static and so[i].type == A.STATIC_BODY:
            continue
        ep = max(ep, rel_err(so[i].c.x, sg[i].c.x, 1.0), rel_err(so[i].c.y, sg[i].c.y, 1.0), rel_err(so[i].a, sg[i].a, 1.0))
        ev = max(ev, rel_err(so[i].v.x, sg[i].v.x, 1.0), rel_err(so[i].v.y, sg[i].v.y, 1.0), rel_err(so[i].w, sg[i].w, 1.0))
    return ep, ev


def f32(x):
    return C.c_float(x).value


def hand_device_order_to_oracle(oracle_api, wg, wo):
    """Give the oracle the Gauss-Seidel schedule the device's LAST step ran (dbx_world_debug_read_solve_order): joints and
    contacts in one rank space (phase by phase; the tile solver's local / boundary / global classes above the colours), position
    iterations walking it backwards.  Within a rank no two constraints share a dynamic body, so the oracle's sequential sweep
    in that order is arithmetically the sweep the device ran.  Call between the device's step and the oracle's; returns
    (contacts the oracle found, info = [position passes backwards, contact colours, joint colours, tiles])."""
    recs, n = wg.read_contacts()
    nj = wg.read_joints()[1]
    cr = (C.c_int32 * max(n, 1))()
    jr = (C.c_int32 * max(nj, 1))()
    info = (C.c_int32 * 4)()
    rc = wg._api.world_debug_read_solve_order(wg._w, cr, n, jr, nj, info)
    assert rc == n, (rc, n, wg._api.last_error())
    keys = (C.c_int32 * max(4 * n, 4))()
    for i in range(n):
        r = recs[i]
        keys[4 * i], keys[4 * i + 1], keys[4 * i + 2], keys[4 * i + 3] = r.fixtureA, r.childA, r.fixtureB, r.childB
        if cr[i] < 0:
            cr[i] = 0x7fffffff
    for j in range(nj):
        if jr[j] < 0:
            jr[j] = 0x7fffffff
    found = oracle_api.world_debug_set_solve_order(wo._w, keys, cr, n, jr, nj, 1 if info[0] else 0)
    return found, list(info)
