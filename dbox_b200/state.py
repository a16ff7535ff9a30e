"""Snapshot / restore of a world's dynamic state through the bulk ABI calls (dbx_world_read_* / dbx_world_write_*).
The scene (bodies, fixtures, joints) must be rebuilt identically before `load`; the snapshot carries what a step
changes: body states, proxy AABBs (tight + fat), the contact cache with manifolds and impulses, joint impulses,
the pending move buffer and inv_dt0 (the checkpoint/resume row of SURVEY.md section 5)."""
import ctypes as C
import pickle

from . import _abi as A


def save(world, path):
    bodies, nb = world.read_bodies()
    prox, npx = world.read_proxies()
    cons, nc = world.read_contacts()
    joints, nj = world.read_joints()
    blob = {"nb": nb, "bodies": bytes(bodies)[:nb * C.sizeof(A.BodyState)], "np": npx, "proxies": bytes(prox)[:npx * C.sizeof(A.ProxyRec)],
            "nc": nc, "contacts": bytes(cons)[:nc * C.sizeof(A.ContactRec)], "nj": nj, "joints": bytes(joints)[:nj * C.sizeof(A.JointState)],
            "moves": world.read_moves(), "inv_dt0": world.get_inv_dt0()}
    with open(path, "wb") as f:
        pickle.dump(blob, f, protocol=4)


def load(world, path):
    with open(path, "rb") as f:
        blob = pickle.load(f)
    api, w = world._api, world._w

    def arr(typ, n, raw):
        a = (typ * max(n, 1))()
        C.memmove(a, raw, len(raw))
        return a
    assert api.world_write_bodies(w, arr(A.BodyState, blob["nb"], blob["bodies"]), blob["nb"]) == blob["nb"]
    assert api.world_write_proxies(w, arr(A.ProxyRec, blob["np"], blob["proxies"]), blob["np"]) == blob["np"]
    if blob["nj"]:
        assert api.world_write_joints(w, arr(A.JointState, blob["nj"], blob["joints"]), blob["nj"]) == blob["nj"]
    rc = api.world_write_contacts(w, arr(A.ContactRec, blob["nc"], blob["contacts"]), blob["nc"])
    assert rc == blob["nc"], (rc, api.last_error())
    moves = blob["moves"]
    flat = (C.c_int32 * max(2 * len(moves), 2))(*[x for m in moves for x in m])
    assert api.world_write_moves(w, flat, len(moves)) == len(moves)
    api.world_set_inv_dt0(w, blob["inv_dt0"])


def capture(src):
    """The dynamic state of world `src` as the ABI's bulk records (ctypes arrays + counts)."""
    bodies, nb = src.read_bodies()
    prox, npx = src.read_proxies()
    joints, nj = src.read_joints()
    cons, nc = src.read_contacts()
    return {"bodies": bodies, "nb": nb, "proxies": prox, "np": npx, "joints": joints, "nj": nj, "contacts": cons, "nc": nc,
            "moves": src.read_moves(), "inv_dt0": src.get_inv_dt0()}


def apply(dst, snap):
    """Write a `capture` into world `dst` (the same scene built on it).  Works with any library that exports the ABI's
    write_* calls -- the CUDA library, or the CPU oracle in tests / bench.py's reference arm."""
    api, w = dst._api, dst._w
    rc = api.world_write_bodies(w, snap["bodies"], snap["nb"])
    assert rc == snap["nb"], ("write_bodies", rc, snap["nb"])
    rc = api.world_write_proxies(w, snap["proxies"], snap["np"])
    assert rc == snap["np"], ("write_proxies", rc, snap["np"])
    if snap["nj"]:
        rc = api.world_write_joints(w, snap["joints"], snap["nj"])
        assert rc == snap["nj"], ("write_joints", rc, snap["nj"])
    rc = api.world_write_contacts(w, snap["contacts"], snap["nc"])
    assert rc == snap["nc"], ("write_contacts", rc, snap["nc"])
    moves = snap["moves"]
    flat = (C.c_int32 * max(2 * len(moves), 2))(*[x for m in moves for x in m])
    rc = api.world_write_moves(w, flat, len(moves))
    assert rc == len(moves), ("write_moves", rc, len(moves))
    api.world_set_inv_dt0(w, snap["inv_dt0"])
    return {"bodies": snap["nb"], "proxies": snap["np"], "joints": snap["nj"], "contacts": snap["nc"], "moves": len(moves)}


def transplant(src, dst):
    """Copy the dynamic state of world `src` into world `dst` (the same scene built on both): bodies, proxies (tight + fat
    AABBs), joints' accumulated impulses, the contact cache (manifolds, impulses, flags), the pending move buffer, inv_dt0."""
    return apply(dst, capture(src))
