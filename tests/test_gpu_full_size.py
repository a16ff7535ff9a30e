"""BASELINE config 4 at its FULL size (100 000-body pile with joint chains) through size-independent properties: the oracle
cannot step a world this big in seconds, so what is checked is what must hold whatever the size --
* the broadphase + pair cache are exact: after a step every pair of proxies whose persistent fat AABBs overlap (different
  bodies, not vetoed by a joint: b2body.d:476-500, b2contactmanager.d:52-176) owns a contact, and after the next Collide no
  contact is left whose fat boxes do not overlap (b2contactmanager.d:299-308).  The checker is an independent numpy / k-d tree
  sweep over the proxy boxes read back through the ABI;
* the solver schedule is a proper colouring (no two constraints of a colour share a dynamic body);
* replay is exact: a snapshot taken mid-run, restored and stepped again reproduces the run bit for bit;
* the state stays finite and inside the container."""
import ctypes as C

import numpy as np
import pytest

from dbox_b200 import _abi as A
from dbox_b200 import scenes

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0


def _np(buf, n, typ):
    return np.frombuffer(buf, dtype=np.uint8, count=n * C.sizeof(typ)).reshape(n, C.sizeof(typ))


def _proxy_arrays(world):
    buf, n = world.read_proxies()
    raw = _np(buf, n, A.ProxyRec)
    ids = raw[:, :12].copy().view(np.int32)          # fixture child proxyId
    fat = raw[:, 28:44].copy().view(np.float32)      # lo.x lo.y hi.x hi.y
    return ids[:, 0], ids[:, 1], fat


def _contact_pairs(world):
    buf, n = world.read_contacts()
    ids = _np(buf, n, A.ContactRec)[:, :16].copy().view(np.int32)   # fixtureA fixtureB childA childB
    return ids


def _overlapping_fat_pairs(fat):
    """indices (i < j) of boxes that overlap (b2TestOverlap: touching edges count), by an independent method"""
    from scipy.spatial import cKDTree
    ext = np.maximum(fat[:, 2] - fat[:, 0], fat[:, 3] - fat[:, 1])
    small = np.nonzero(ext < 4.0)[0]
    big = np.nonzero(ext >= 4.0)[0]
    c = 0.5 * (fat[small, :2] + fat[small, 2:])
    r = float(np.sqrt(2.0) * ext[small].max()) + 1e-3
    cand = cKDTree(c).query_pairs(r, output_type="ndarray")
    i, j = small[cand[:, 0]], small[cand[:, 1]]

    def ov(a, b):
        return (fat[a, 0] <= fat[b, 2]) & (fat[b, 0] <= fat[a, 2]) & (fat[a, 1] <= fat[b, 3]) & (fat[b, 1] <= fat[a, 3])
    m = ov(i, j)
    pairs = [np.stack([i[m], j[m]], 1)]
    allidx = np.arange(fat.shape[0])
    for b in big:
        hit = allidx[ov(np.full(allidx.shape, b), allidx) & (allidx != b)]
        hit = hit[(ext[hit] < 4.0) | (hit > b)]       # big-big pairs once
        pairs.append(np.stack([np.full(hit.shape, b), hit], 1))
    p = np.concatenate(pairs, 0)
    return np.stack([p.min(1), p.max(1)], 1)


def _key(a, b):
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    return lo.astype(np.int64) * (1 << 32) + hi.astype(np.int64)


def test_pile_100k_full_size_properties(gpu_api):
    n, columns = 100000, 1000
    w, bodies, njoints = scenes.pile(api=gpu_api, n=n, columns=columns)
    assert njoints > 15000
    w.SetAllowSleeping(False)
    w.StepN(DT, 8, 3, 200)                                     # the grid's 0.05 gaps close, the pile carries its weight
    cnt = w.counts()
    assert cnt.bodies == n + 1 and cnt.contacts > 2 * n and cnt.touching > n and cnt.awakeBodies == n
    assert w._api.world_debug_colour_conflicts(w._w) == 0

    # ---- broadphase / pair-cache exactness at full size
    fix, child, fat = _proxy_arrays(w)
    assert fat.shape[0] == cnt.proxies and np.isfinite(fat).all()
    fix_body = np.full(int(fix.max()) + 1, -1, np.int64)
    for f in w._fixtures.values():
        fix_body[f.id] = f.body.id
    body = fix_body[fix]
    ov = _overlapping_fat_pairs(fat)
    bi, bj = body[ov[:, 0]], body[ov[:, 1]]
    veto = np.array(sorted(_key(np.array([j.bodyA.id]), np.array([j.bodyB.id]))[0] for j in w._joints.values() if not j.collideConnected), np.int64)
    assert veto.size > 5000
    keep = (bi != bj) & ~np.isin(_key(bi, bj), veto)
    # proxies are identified by (fixture, child); child < 2^12 here
    pid = fix.astype(np.int64) * 4096 + child
    expected = np.unique(_key(pid[ov[keep, 0]], pid[ov[keep, 1]]))
    c = _contact_pairs(w)
    have = np.unique(_key(c[:, 0].astype(np.int64) * 4096 + c[:, 2], c[:, 1].astype(np.int64) * 4096 + c[:, 3]))
    assert have.size == c.shape[0] == cnt.contacts            # one contact per proxy pair
    missing = np.setdiff1d(expected, have)
    assert missing.size == 0, "overlapping proxy pairs without a contact: %d of %d" % (missing.size, expected.size)
    stale = np.setdiff1d(have, expected)
    assert stale.size < 0.05 * have.size                       # boxes that drifted apart during the last step: gone after Collide
    assert w._api.world_stage_collide(w._w) >= 0
    c2 = _contact_pairs(w)
    have2 = np.unique(_key(c2[:, 0].astype(np.int64) * 4096 + c2[:, 2], c2[:, 1].astype(np.int64) * 4096 + c2[:, 3]))
    assert np.array_equal(have2, expected), (have2.size, expected.size)

    # ---- exact replay from a snapshot
    need = gpu_api.world_export_state(w._w, None, 0)
    blob = (C.c_char * need)()
    assert gpu_api.world_export_state(w._w, blob, need) == need
    w.StepN(DT, 8, 3, 12)
    buf, nb = w.read_bodies()
    first = _np(buf, nb, A.BodyState).copy()
    counts1 = w.counts()
    assert gpu_api.world_import_state(w._w, blob, need) == 0
    w.StepN(DT, 8, 3, 12)
    buf, nb2 = w.read_bodies()
    second = _np(buf, nb2, A.BodyState)
    counts2 = w.counts()
    assert nb == nb2 and np.array_equal(first, second)
    assert (counts1.contacts, counts1.touching) == (counts2.contacts, counts2.touching)

    # ---- the state is finite and inside the container
    st = np.frombuffer(second.tobytes(), dtype=np.float32).reshape(nb, -1)
    assert np.isfinite(st).all()
    half_w = 1.05 * columns * 0.5 + 5.0
    px, py = st[1:, 2], st[1:, 3]                             # BodyState.p of the dynamic bodies (body 0 is the ground)
    assert px.min() > -half_w - 0.6 and px.max() < half_w + 0.6 and py.min() > 0.3
    assert w._api.world_debug_colour_conflicts(w._w) == 0


def test_batched_65536_pyramids_full_size_properties(gpu_api):
    """BASELINE config 5 at its full size on one GPU: 65,536 Pyramid worlds (13.9 M bodies) as replicas of one device world.
    Size-independent properties: worlds that start identical stay identical BIT FOR BIT (every replica equals replica 0, which
    tests/test_gpu_parity.py compares with a world stepped alone), the populations are exact multiples of one world's, the
    solver schedule is a proper colouring, and worlds that are then pushed apart diverge without touching their neighbours"""
    worlds = 65536
    single, _ = scenes.pyramid(api=gpu_api)
    caps = A.Caps(); caps.maxContacts = worlds * 700
    batch, _ = scenes.pyramid(api=gpu_api, caps=caps)
    nb = batch.counts().bodies
    batch.Replicate(worlds)
    assert gpu_api.world_replica_count(batch._w) == worlds
    single.StepN(DT, 8, 3, 40); batch.StepN(DT, 8, 3, 40)
    cs, cb = single.counts(), batch.counts()
    assert cb.bodies == nb * worlds
    assert (cb.contacts, cb.touching, cb.awakeBodies) == (cs.contacts * worlds, cs.touching * worlds, cs.awakeBodies * worlds)
    assert gpu_api.world_debug_colour_conflicts(batch._w) == 0
    xf = np.empty((nb * worlds, 4), np.float32)
    assert gpu_api.world_read_transforms(batch._w, xf.ctypes.data, nb * worlds) == nb * worlds
    xf = xf.reshape(worlds, nb, 4).view(np.uint32)
    assert (xf == xf[0]).all()
    one = np.empty((nb, 4), np.float32)
    assert gpu_api.world_read_transforms(single._w, one.ctypes.data, nb) == nb
    assert np.array_equal(one.view(np.uint32), xf[0])
    # push every 4096th world: it diverges, the worlds next to it do not
    vel = np.zeros((worlds // 4096, 4), np.float32); vel[:, 0] = 3.0
    ids = (np.arange(worlds // 4096, dtype=np.int32) * 4096 * nb + nb - 1)          # the top box of those worlds
    batch.SetBodyStates(ids=ids, vel=vel)
    batch.StepN(DT, 8, 3, 20)
    xf2 = np.empty((nb * worlds, 4), np.float32)
    assert gpu_api.world_read_transforms(batch._w, xf2.ctypes.data, nb * worlds) == nb * worlds
    xf2 = xf2.reshape(worlds, nb, 4).view(np.uint32)
    same = (xf2 == xf2[1]).all(axis=(1, 2))
    assert not same[::4096].any() and same.sum() == worlds - worlds // 4096
    batch.close(); single.close()


def test_tumbler_20k_full_size_invariants(gpu_api):
    """BASELINE config 3 at its full size: the Tumbler grown to 20,000 bodies (container x5); bodies are added
    while the world runs (the reference adds one per step; here 16 side by side so that the scene is full after 1,250 steps).
    The container is ONE dynamic body under hundreds of contacts (overflow colour lanes).  Invariants: nothing leaks out of the
    container, the schedule stays a proper colouring, the motor keeps its speed, populations are sane, the state is finite"""
    n = 20000
    t = scenes.Tumbler(api=gpu_api, count=n, scale=5.0)          # default pools: the contact pool grows by its watermark while the scene fills
    steps = 0
    while t.m_count < n:
        t.Step(spawn_per_step=16); steps += 1
    for _ in range(150):
        t.Step(); steps += 1
    w = t.world
    cnt = w.counts()
    assert cnt.bodies == n + 3 and cnt.joints == 1 and cnt.contacts > n and cnt.touching > n // 2
    assert gpu_api.world_debug_colour_conflicts(w._w) == 0
    buf, nb = w.read_bodies()
    st = np.frombuffer(buf, dtype=np.float32, count=nb * 29).reshape(nb, 29)
    assert np.isfinite(st).all()
    c = t.container._state()
    p = st[3:, 2:4]                                               # the spawned bodies (ids 0-2: demo body, ground, container)
    dx, dy = p[:, 0] - c.p.x, p[:, 1] - c.p.y
    lx, ly = c.qc * dx + c.qs * dy, -c.qs * dx + c.qc * dy        # container frame
    assert np.abs(lx).max() < 53.0 and np.abs(ly).max() < 53.0
    assert abs(t.container.GetAngularVelocity() - 0.05 * np.pi) < 2e-3
    assert abs(t.container.GetAngle() - steps * DT * 0.05 * np.pi) < 0.05


# ---------------------------------------------------------------------------------------------------------------------
# Config 4 at full size against the oracle, one step: the settled 100,000-body device world is transplanted into the CPU
# oracle (orc_world_write_*), both step once from the identical state, the oracle walking its constraints in the schedule
# the device's colouring produced (parity.hand_device_order_to_oracle; the DFS order of a 100k-body island would need far
# more levels than the device's level override holds, so the order travels device -> oracle here, and oracle -> device in
# tests/test_gpu_parity.py at sizes where it fits).  north_star tolerances: contact set / feature keys exact, manifolds
# 1e-5, velocities and impulses 1e-4 relative.
def _recs(buf, n, typ):
    return np.frombuffer(buf, dtype=np.dtype(typ), count=n)


def _ckey(r):
    return ((r["fixtureA"].astype(np.int64) * 64 + r["childA"]) << 32) | (r["fixtureB"].astype(np.int64) * 64 + r["childB"])


def _rel(a, b, floor):
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)


def _compare_step(wg, wo, label):
    """both worlds have just stepped once from the same state; returns, per quantity, (worst relative error, fraction of
    items above the north_star tolerance)"""
    def stat(errs, tol):
        e = np.concatenate([np.ravel(x) for x in errs]) if errs else np.zeros(1)
        return float(e.max()) if e.size else 0.0, float((e > tol).mean()) if e.size else 0.0
    bg, nb = wg.read_bodies(); bo, nbo = wo.read_bodies()
    assert nb == nbo
    G, O = _recs(bg, nb, A.BodyState), _recs(bo, nb, A.BodyState)
    dyn = O["type"] != A.STATIC_BODY
    pos = stat([_rel(G["c"]["x"][dyn], O["c"]["x"][dyn], 1.0), _rel(G["c"]["y"][dyn], O["c"]["y"][dyn], 1.0), _rel(G["a"][dyn], O["a"][dyn], 1.0)], 1e-5)
    # per body: the worst of its three velocity components
    evb = np.maximum(np.maximum(_rel(G["v"]["x"][dyn], O["v"]["x"][dyn], 1.0), _rel(G["v"]["y"][dyn], O["v"]["y"][dyn], 1.0)), _rel(G["w"][dyn], O["w"][dyn], 1.0))
    vel = stat([evb], 1e-4)
    assert np.array_equal(G["flags"][dyn] & 0x2, O["flags"][dyn] & 0x2), "awake flags differ"
    cg, ng = wg.read_contacts(); co, no = wo.read_contacts()
    CG, CO = _recs(cg, ng, A.ContactRec), _recs(co, no, A.ContactRec)
    kg, ko = _ckey(CG), _ckey(CO)
    ig, io = np.argsort(kg), np.argsort(ko)
    assert ng == no and np.array_equal(kg[ig], ko[io]), "%s: contact sets differ (%d vs %d)" % (label, ng, no)
    CG, CO = CG[ig], CO[io]
    assert np.array_equal(CG["flags"] & 0x6, CO["flags"] & 0x6), "touching / enabled flags differ"
    mg, mo = CG["manifold"], CO["manifold"]
    assert np.array_equal(mg["pointCount"], mo["pointCount"])
    live = mo["pointCount"] > 0          # the type of an empty manifold is whatever the last Evaluate left behind
    assert np.array_equal(mg["type"][live], mo["type"][live])
    em, ei = [], []
    for k in range(2):
        live = mo["pointCount"] > k
        assert np.array_equal(mg["points"]["key"][:, k][live], mo["points"]["key"][:, k][live]), "feature keys differ"
        for ax in ("x", "y"):
            em.append(_rel(mg["points"]["localPoint"][ax][:, k][live], mo["points"]["localPoint"][ax][:, k][live], 1.0))
        for f in ("normalImpulse", "tangentImpulse"):
            ei.append(_rel(mg["points"][f][:, k][live], mo["points"][f][:, k][live], 1.0))
    touching = mo["pointCount"] > 0
    for f in ("localNormal", "localPoint"):
        for ax in ("x", "y"):
            em.append(_rel(mg[f][ax][touching], mo[f][ax][touching], 1.0))
    jg, nj = wg.read_joints(); jo, njo = wo.read_joints()
    assert nj == njo
    JG, JO = _recs(jg, nj, A.JointState), _recs(jo, nj, A.JointState)
    assert np.array_equal(JG["limitState"], JO["limitState"])
    return {"pos": pos, "vel": vel, "manifold": stat(em, 1e-5), "contact_impulse": stat(ei, 1e-4),
            "joint_impulse": stat([_rel(JG["impulse"], JO["impulse"], 1.0)] if nj else [], 1e-4),
            "contacts": int(ng), "touching": int(touching.sum()), "joints": int(nj)}


def _one_ulp_twin(snap):
    """the snapshot with every body angle moved by one ulp (sin / cos left alone): what a 1-ulp difference between CUDA's and
    glibc's sinf / cosf does to a step, measured on the oracle itself"""
    B = np.frombuffer(snap["bodies"], dtype=np.dtype(A.BodyState), count=snap["nb"]).copy()
    B["a"] = np.nextafter(B["a"], np.float32(10.0)).astype(np.float32)
    twin = dict(snap)
    twin["bodies"] = (A.BodyState * snap["nb"]).from_buffer_copy(B.tobytes())
    return twin


@pytest.mark.parametrize("n,columns,settle", [(3000, 100, 300), (100000, 1000, 600)])
def test_pile_single_step_matches_oracle(gpu_api, oracle_api, n, columns, settle):
    _pile_single_step_vs_oracle(gpu_api, oracle_api, n, columns, settle)


@pytest.mark.parametrize("flags,what", [(128, "grid barriers instead of neighbour handshakes"),
                                        (2048, "joints stay in the global arrays"),
                                        (4096, "the neighbour's bodies are reached through L2 (no staging, no staged boundary rows)"),
                                        (8192, "boundary rows stay in the global arrays"),
                                        (32768, "more neighbour bodies than slots for them: the whole tile falls back to L2"),
                                        (64, "tile solver off: k_solve")])
def test_pile_single_step_matches_oracle_on_the_fallback_paths(gpu_api, oracle_api, flags, what):
    """k_solve_tiles takes these branches when something does not fit in a tile's shared memory (a tile richer in joints or
    boundary rows than the pile's); the DBX_DEBUG bits force them so that they are held to the same oracle as the fast paths.
    The library reads the variable whenever it rebuilds its device view (world creation included)."""
    import os
    old = os.environ.get("DBX_DEBUG")
    os.environ["DBX_DEBUG"] = str(flags)
    try:
        _pile_single_step_vs_oracle(gpu_api, oracle_api, 3000, 100, 300)
    finally:
        if old is None:
            del os.environ["DBX_DEBUG"]
        else:
            os.environ["DBX_DEBUG"] = old


def test_pile_with_global_constraints_matches_oracle(gpu_api, oracle_api):
    """Distance joints across half the pile: constraints of the tile solver's global class (grid-wide phases after the local and
    boundary ones, grid barriers instead of neighbour handshakes), held to the same oracle."""
    r = _pile_single_step_vs_oracle(gpu_api, oracle_api, 3000, 100, 300, long_links=12)
    assert r[False]["tiles"] > 0 and r[False]["global_rows"] > 0, r


def test_tiles_without_any_boundary_constraint_match_oracle(gpu_api, oracle_api):
    """Twelve separate clusters of 256 boxes: twelve tiles, not one boundary or global constraint -- the tile solver's passes then
    run with no exchange and no barrier between tiles at all (the path a world of scattered piles takes)."""
    from dbox_b200 import state
    from tests.parity import hand_device_order_to_oracle
    wg, _ = scenes.islands_of_boxes(api=gpu_api)
    wo, _ = scenes.islands_of_boxes(api=oracle_api)
    for w in (wg, wo):
        w.SetAllowSleeping(False)
    wg.StepN(DT, 8, 3, 150)
    snap = state.capture(wg)
    for continuous in (False, True):
        for w in (wg, wo):
            w.SetContinuousPhysics(continuous)
            state.apply(w, snap)
        wg.Step(DT, 8, 3)
        found, info = hand_device_order_to_oracle(oracle_api, wg, wo)
        wo.Step(DT, 8, 3)
        hb = (C.c_int32 * 2400)()
        gpu_api.world_debug_header(wg._w, hb, 9600)
        assert info[3] == 12 and hb[1124] == 0 and hb[1125] == 0, (info, hb[1124], hb[1125])      # tiles, boundary rows, global rows
        cg, co = wg.counts(), wo.counts()
        assert cg.touching == co.touching and cg.contacts == co.contacts and cg.islands == co.islands >= 12, (cg.touching, co.touching, cg.contacts, co.contacts, cg.islands, co.islands)
        r = _compare_step(wg, wo, "continuous=%s" % continuous)
        assert r["manifold"][0] < 1e-5 and r["pos"][0] < 2e-5 and r["vel"][0] < 2e-3, r
    wg.close(); wo.close()


def _pile_single_step_vs_oracle(gpu_api, oracle_api, n, columns, settle, **scene):
    """One step of the settled pile on the device against the sequential oracle walking the device's own Gauss-Seidel order.
    Exact: contact set, touching flags, manifold types, feature keys, manifolds (bit for bit), island count.  Velocities,
    positions and impulses: north_star's 1e-4 / 1e-5 relative for all but a fraction of a per cent of the bodies, and for
    those the yardstick is the solver's own conditioning -- the block solver accepts 2-point contacts whose K matrix has a
    condition number up to 1000 (b2contactsolver.d:418-448), deep stacks chain such contacts, and the oracle run against ITSELF
    with every body angle moved by one ulp spreads just as far.  The device must stay within a small factor of that spread.
    (A wrong order or a wrong warm start shows as 1e-2 .. 1 on hundreds of bodies: both were found with this test.)"""
    from dbox_b200 import state
    from tests.parity import hand_device_order_to_oracle
    wg, _, nj = scenes.pile(api=gpu_api, n=n, columns=columns, **scene)
    wo, _, _ = scenes.pile(api=oracle_api, n=n, columns=columns, **scene)
    wt, _, _ = scenes.pile(api=oracle_api, n=n, columns=columns, **scene)
    for w in (wg, wo, wt):
        w.SetAllowSleeping(False)
    wg.StepN(DT, 8, 3, settle)
    report = {}
    snap = state.capture(wg)
    twin = _one_ulp_twin(snap)
    assert snap["nb"] == n + 1 and snap["nj"] == nj
    for continuous in (False, True):
        for w in (wg, wo, wt):
            w.SetContinuousPhysics(continuous)
            state.apply(w, twin if w is wt else snap)
        assert oracle_api.world_tree_validate(wo._w) == 1
        wg.Step(DT, 8, 3)
        found, info = hand_device_order_to_oracle(oracle_api, wg, wo)
        hand_device_order_to_oracle(oracle_api, wg, wt)
        wo.Step(DT, 8, 3); wt.Step(DT, 8, 3)
        cg, co = wg.counts(), wo.counts()
        assert found > 0
        assert cg.islands == co.islands, ("island count", cg.islands, co.islands)          # row a14
        assert cg.touching == co.touching and cg.contacts == co.contacts
        r = _compare_step(wg, wo, "continuous=%s" % continuous)
        try:
            own = _compare_step(wt, wo, "oracle twin")
        except AssertionError:          # the ulp flipped a feature somewhere: no yardstick from this pair
            own = None
        hb = (C.c_int32 * 2400)()
        gpu_api.world_debug_header(wg._w, hb, 9600)        # Header::nTileB, nTileG (dbx_device.cuh) sit at ints 1124, 1125
        r.update(order_found=found, position_backwards=info[0], colours=info[1], joint_colours=info[2], tiles=info[3], islands=cg.islands,
                 boundary_rows=hb[1124], global_rows=hb[1125])
        report[continuous] = r
        print("pile %d single step, device vs oracle, continuous=%s: %s" % (n, continuous, r))
        print("pile %d single step, oracle vs oracle with angles + 1 ulp:  %s" % (n, own))
        # bit for bit after Solve; with the TOI sub-steps in, contacts are re-evaluated at positions that carry the deviation below
        assert r["manifold"][0] == 0.0 if not continuous else r["manifold"][0] < 1e-5, r
        for q, tol in (("pos", 1e-5), ("vel", 1e-4), ("contact_impulse", 1e-4), ("joint_impulse", 1e-4)):
            worst, frac = r[q]
            own_worst, own_frac = own[q] if own else (0.0, 0.0)
            assert frac <= max(8.0 * own_frac, 1e-2), (q, r[q], own and own[q])        # how many items are beyond the tolerance
            assert worst <= max(8.0 * own_worst, 20.0 * tol), (q, r[q], own and own[q])    # and how far
    wg.close(); wo.close(); wt.close()
    return report
