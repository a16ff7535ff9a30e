// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.b2world;

import dbox.common;
import dbox.collision;
import dbox.collision.shapes;
import dbox.dynamics.b2body;
import dbox.dynamics.b2fixture;
import dbox.dynamics.b2timestep;
import dbox.dynamics.b2worldcallbacks;
import dbox.dynamics.contacts.b2contact;
import dbox.dynamics.joints.b2joint;
import dbox_b200_c;

/// reference: dynamics/b2world.d:34-1630.  The world lives on the GPU behind a dbx_world handle; this struct keeps the host-side
/// object graph a D program walks (bodies, fixtures, joints as stable heap objects) and forwards everything else.
struct b2World
{
    @disable this();
    @disable this(this);

    /// reference: :865-919
    this(b2Vec2 gravity)
    {
        m_handle = dbx_world_create(gravity.x, gravity.y, /*device*/ 0, /*caps*/ null);
        if (m_handle is null)
            throw new Exception("dbox_b200: " ~ fromCString(dbx_last_error()));   // e.g. "no CUDA device": there is no CPU fallback
        m_gravity = gravity;
    }

    ~this()
    {
        if (m_handle !is null)
            dbx_world_destroy(cast(dbx_world*) m_handle);
        m_handle = null;
    }

    void SetDestructionListener(b2DestructionListener listener) { m_destructionListener = listener; }      /// :44-47
    void SetContactFilter(b2ContactFilter filter)                                                           /// :52-56
    {
        m_contactFilter = filter;
        dbx_world_set_user_filter(h, filter is null ? 0 : (DBX_FILTER_LOG | DBX_FILTER_REPLACES_DEFAULT));
    }
    void SetContactListener(b2ContactListener listener)                                                     /// :59-63
    {
        m_contactListener = listener;
        dbx_world_enable_contact_events(h, listener is null ? 0 : 1 << 16);
        dbx_world_enable_post_solve(h, listener is null ? 0 : 1 << 16);
    }

    /// reference: :75-99
    b2Body* CreateBody(const(b2BodyDef)* def)
    {
        dbx_body_def bd;
        dbx_default_body_def(&bd);
        bd.type = cast(int) def.type;
        bd.position = dbx_vec2(def.position.x, def.position.y);
        bd.angle = def.angle;
        bd.linearVelocity = dbx_vec2(def.linearVelocity.x, def.linearVelocity.y);
        bd.angularVelocity = def.angularVelocity;
        bd.linearDamping = def.linearDamping; bd.angularDamping = def.angularDamping;
        bd.allowSleep = def.allowSleep ? 1 : 0; bd.awake = def.awake ? 1 : 0; bd.fixedRotation = def.fixedRotation ? 1 : 0;
        bd.bullet = def.bullet ? 1 : 0; bd.active = def.active ? 1 : 0;
        bd.gravityScale = def.gravityScale;
        const int id = dbx_body_create(h, &bd);
        if (id < 0)
            return null;                      // DBX_E_LOCKED inside a callback, like the reference's IsLocked() early-out
        auto b = new b2Body;
        b.m_id = id; b.m_world = &this; b.m_userData = cast(void*) def.userData;
        b.m_next = m_bodyList;
        if (m_bodyList !is null) m_bodyList.m_prev = b;
        m_bodyList = b;
        ++m_bodyCount;
        return b;
    }

    /// reference: :105-191 (joints and fixtures of the body go with it; the listeners hear about them first)
    void DestroyBody(b2Body* b)
    {
        for (b2Joint j = m_jointList; j !is null;)
        {
            b2Joint next = j.m_next;
            if (j.m_bodyA == b || j.m_bodyB == b)
            {
                if (m_destructionListener !is null) m_destructionListener.SayGoodbye(j);
                DestroyJoint(j);
            }
            j = next;
        }
        for (b2Fixture* f = b.m_fixtureList; f !is null; f = f.m_next)
        {
            if (m_destructionListener !is null) m_destructionListener.SayGoodbye(f);
            unregisterFixture(f);
        }
        dbx_body_destroy(h, b.m_id);
        if (b.m_prev !is null) b.m_prev.m_next = b.m_next;
        if (b.m_next !is null) b.m_next.m_prev = b.m_prev;
        if (b == m_bodyList) m_bodyList = b.m_next;
        --m_bodyCount;
    }

    /// reference: :196-261.  The joint definition classes (dbox.dynamics.joints.*) fill a dbx_joint_def.
    b2Joint CreateJoint(const(b2JointDef) def)
    {
        dbx_joint_def jd = def.toDevice();
        const int id = dbx_joint_create(h, &jd);
        if (id < 0)
            return null;
        auto j = new b2Joint(id, cast(b2JointDef) def, &this);
        j.m_next = m_jointList;
        if (m_jointList !is null) m_jointList.m_prev = j;
        m_jointList = j;
        ++m_jointCount;
        return j;
    }

    /// reference: :265-360
    void DestroyJoint(b2Joint j)
    {
        dbx_joint_destroy(h, j.m_id);
        if (j.m_prev !is null) j.m_prev.m_next = j.m_next;
        if (j.m_next !is null) j.m_next.m_prev = j.m_prev;
        if (j is m_jointList) m_jointList = j.m_next;
        --m_jointCount;
    }

    /// reference: :367-434.  Without a listener: one call.  With one: the step is cut where the reference calls PreSolve
    /// (dbx_world_step_begin = Collide, dbx_world_step_end = Solve + SolveTOI), and Begin / End / PostSolve are replayed from the
    /// device's records in the reference's call order.
    void Step(float32 dt, int32 velocityIterations, int32 positionIterations)
    {
        if (m_contactListener is null && m_contactFilter is null)
        {
            check(dbx_world_step(h, dt, velocityIterations, positionIterations));
            return;
        }
        check(dbx_world_step_begin(h, dt, velocityIterations, positionIterations));
        applyUserFilter();
        if (m_contactListener !is null)
        {
            dispatchContactEvents();
            preSolve();
        }
        check(dbx_world_step_end(h));
        if (m_contactListener !is null)
        {
            dispatchContactEvents();
            postSolve();
        }
    }

    void ClearForces() { dbx_world_clear_forces(h); }                                                       /// :443-450

    /// reference: :563-575.  Fat-AABB query on the device tree; the callback sees fixtures until it returns false.
    void QueryAABB(b2QueryCallback callback, b2AABB aabb)
    {
        dbx_aabb box = dbx_aabb(dbx_vec2(aabb.lowerBound.x, aabb.lowerBound.y), dbx_vec2(aabb.upperBound.x, aabb.upperBound.y));
        int cap = 256;
        for (;;)
        {
            auto hits = new int[2 * cap];
            int count;
            dbx_world_query_aabb(h, &box, 1, cap, &count, hits.ptr);
            if (count > cap) { cap = count; continue; }
            foreach (k; 0 .. count)
                if (auto f = hits[2 * k] in m_fixtures)
                    if (!callback.ReportFixture(*f)) return;
            return;
        }
    }

    /// reference: :577-590 + b2dynamictree.d:239-331.  All hits along the ray come back from the device in one call; the callback's
    /// return value clips the ray exactly as the reference's tree walk does (0 stop, fraction clip, 1 keep, -1 ignore).
    void RayCast(b2RayCastCallback callback, b2Vec2 point1, b2Vec2 point2)
    {
        dbx_ray ray = dbx_ray(dbx_vec2(point1.x, point1.y), dbx_vec2(point2.x, point2.y));
        int cap = 256;
        for (;;)
        {
            auto hits = new dbx_ray_hit[cap];
            int count;
            dbx_world_raycast_all(h, &ray, 1, cap, &count, hits.ptr);
            if (count > cap) { cap = count; continue; }
            float32 maxFraction = 1.0f;
            foreach (k; 0 .. count)
            {
                if (hits[k].fraction > maxFraction) continue;
                auto f = hits[k].fixture in m_fixtures;
                if (f is null) continue;
                const float32 r = callback.ReportFixture(*f, b2Vec2(hits[k].point.x, hits[k].point.y), b2Vec2(hits[k].normal.x, hits[k].normal.y), hits[k].fraction);
                if (r == 0.0f) return;
                if (r > 0.0f) maxFraction = r;
            }
            return;
        }
    }

    inout(b2Body)* GetBodyList() inout { return m_bodyList; }
    inout(b2Joint) GetJointList() inout { return m_jointList; }
    /// reference: :610-613.  Rebuilt from the device pair cache on demand (newest first is not kept: pair-key order).
    b2Contact GetContactList() { refreshContacts(); return m_contactList; }

    bool GetAllowSleeping() const { return (flags() & DBX_WORLD_ALLOW_SLEEP) != 0; }
    void SetAllowSleeping(bool flag) { setFlag(DBX_WORLD_ALLOW_SLEEP, flag); }                              /// :622-636
    bool GetWarmStarting() const { return (flags() & DBX_WORLD_WARM_STARTING) != 0; }
    void SetWarmStarting(bool flag) { setFlag(DBX_WORLD_WARM_STARTING, flag); }
    bool GetContinuousPhysics() const { return (flags() & DBX_WORLD_CONTINUOUS) != 0; }
    void SetContinuousPhysics(bool flag) { setFlag(DBX_WORLD_CONTINUOUS, flag); }
    bool GetSubStepping() const { return (flags() & DBX_WORLD_SUB_STEPPING) != 0; }
    void SetSubStepping(bool flag) { setFlag(DBX_WORLD_SUB_STEPPING, flag); }
    bool GetAutoClearForces() const { return (flags() & DBX_WORLD_AUTO_CLEAR_FORCES) != 0; }
    void SetAutoClearForces(bool flag) { setFlag(DBX_WORLD_AUTO_CLEAR_FORCES, flag); }
    int32 GetProxyCount() const { return counts().proxies; }
    int32 GetBodyCount() const { return m_bodyCount; }
    int32 GetJointCount() const { return m_jointCount; }
    int32 GetContactCount() const { return counts().contacts; }
    int32 GetTreeHeight() const { int hgt, bal; float q; dbx_world_tree_stats(cast(dbx_world*) m_handle, &hgt, &bal, &q); return hgt; }
    int32 GetTreeBalance() const { int hgt, bal; float q; dbx_world_tree_stats(cast(dbx_world*) m_handle, &hgt, &bal, &q); return bal; }
    float32 GetTreeQuality() const { int hgt, bal; float q; dbx_world_tree_stats(cast(dbx_world*) m_handle, &hgt, &bal, &q); return q; }
    b2Vec2 GetGravity() const { return m_gravity; }
    void SetGravity(b2Vec2 gravity) { m_gravity = gravity; dbx_world_set_gravity(h, gravity.x, gravity.y); }
    bool IsLocked() const { return false; }          /// callbacks run between device launches, never inside one
    void ShiftOrigin(b2Vec2 newOrigin) { dbx_world_shift_origin(h, newOrigin.x, newOrigin.y); }             /// :758-780
    b2Profile GetProfile() const                                                                            /// :789-792
    {
        dbx_profile p;
        dbx_world_profile(cast(dbx_world*) m_handle, &p);
        return b2Profile(p.step, p.collide, p.solve, p.solveInit, p.solveVelocity, p.solvePosition, p.broadphase, p.solveTOI);
    }

    // ---- shim internals
    void registerFixture(b2Fixture* f) { m_fixtures[f.m_id] = f; }
    void unregisterFixture(b2Fixture* f) { m_fixtures.remove(f.m_id); }

    void* m_handle;                  /// dbx_world*

private:
    @property dbx_world* h() { return cast(dbx_world*) m_handle; }
    uint flags() const { return dbx_world_get_flags(cast(dbx_world*) m_handle); }
    void setFlag(uint bit, bool on) { const f = flags(); dbx_world_set_flags(h, on ? (f | bit) : (f & ~bit)); }
    dbx_counts counts() const { dbx_counts c; dbx_world_counts(cast(dbx_world*) m_handle, &c); return c; }
    static void check(int rc) { if (rc < 0) throw new Exception("dbox_b200: " ~ fromCString(dbx_last_error())); }
    static string fromCString(const(char)* s) { import core.stdc.string : strlen; return s is null ? "" : s[0 .. strlen(s)].idup; }

    static ulong pairKey(int fixtureA, int childA, int fixtureB, int childB)
    {
        return (cast(ulong) cast(uint) (fixtureA * 64 + childA) << 32) | cast(uint) (fixtureB * 64 + childB);
    }
    b2Contact contactFor(int fixtureA, int childA, int fixtureB, int childB)
    {
        const key = pairKey(fixtureA, childA, fixtureB, childB);
        if (auto c = key in m_contacts) return *c;
        auto c = new b2Contact;
        auto fa = fixtureA in m_fixtures, fb = fixtureB in m_fixtures;
        c.m_fixtureA = fa is null ? null : *fa; c.m_fixtureB = fb is null ? null : *fb;
        c.m_indexA = childA; c.m_indexB = childB;
        m_contacts[key] = c;
        return c;
    }
    void fill(b2Contact c, ref const dbx_contact_rec r)
    {
        c.m_flags = r.flags; c.m_friction = r.friction; c.m_restitution = r.restitution; c.m_tangentSpeed = r.tangentSpeed; c.m_patchMask = 0;
        c.m_manifold.type = cast(b2Manifold.Type) r.manifold.type;
        c.m_manifold.pointCount = r.manifold.pointCount;
        c.m_manifold.localNormal = b2Vec2(r.manifold.localNormal.x, r.manifold.localNormal.y);
        c.m_manifold.localPoint = b2Vec2(r.manifold.localPoint.x, r.manifold.localPoint.y);
        foreach (k; 0 .. 2)
        {
            c.m_manifold.points[k].localPoint = b2Vec2(r.manifold.points[k].localPoint.x, r.manifold.points[k].localPoint.y);
            c.m_manifold.points[k].normalImpulse = r.manifold.points[k].normalImpulse;
            c.m_manifold.points[k].tangentImpulse = r.manifold.points[k].tangentImpulse;
            c.m_manifold.points[k].id.key = r.manifold.points[k].key;
        }
    }
    dbx_contact_rec[] readContacts()
    {
        const n = dbx_world_read_contacts(h, null, 0);
        auto recs = new dbx_contact_rec[n > 0 ? n : 1];
        const m = dbx_world_read_contacts(h, recs.ptr, n);
        return recs[0 .. (m < n ? m : n)];
    }
    void refreshContacts()
    {
        m_contactList = null;
        b2Contact[ulong] alive;
        foreach_reverse (ref r; readContacts())
        {
            auto c = contactFor(r.fixtureA, r.childA, r.fixtureB, r.childB);
            fill(c, r);
            c.m_next = m_contactList; m_contactList = c;
            alive[pairKey(r.fixtureA, r.childA, r.fixtureB, r.childB)] = c;
        }
        m_contacts = alive;
    }
    /// BeginContact / EndContact in the reference's call order (b2contact.d:338-346, b2contactmanager.d:60-63)
    void dispatchContactEvents()
    {
        dbx_contact_event[256] ev;
        for (;;)
        {
            const n = dbx_world_poll_contact_events(h, ev.ptr, cast(int) ev.length);
            if (n <= 0) return;
            foreach (ref e; ev[0 .. (n < ev.length ? n : ev.length)])
            {
                auto c = contactFor(e.fixtureA, e.childA, e.fixtureB, e.childB);
                if (e.type == DBX_CONTACT_BEGIN) { c.m_flags |= 0x0002; m_contactListener.BeginContact(c); }
                else { c.m_flags &= ~0x0002; m_contactListener.EndContact(c); }
            }
            if (n <= ev.length) return;
        }
    }
    /// PreSolve for every touching, enabled, non-sensor contact after Collide (b2contact.d:348-355); what the callback changed on
    /// the contact goes back to the device as patches before Solve runs.  SetEnabled(false) then holds for the whole step, the
    /// TOI loop's re-evaluations included (where the reference would call PreSolve again, b2world.d:1295,1379).
    /// preSolveLookahead (off by default: the reference never does this): with continuous physics on, also ask about contacts that
    /// are not touching yet -- their manifold is empty -- because a fast body can first touch INSIDE the TOI loop, where the
    /// listener cannot be reached; the answer given here is the one that loop applies (include/dbox_b200.h, "PreSolve").
    bool preSolveLookahead = false;
    void preSolve()
    {
        dbx_contact_patch[] patches;
        b2Manifold oldManifold;          // (the manifold before this step's Update is not kept on the device)
        foreach (ref r; readContacts())
        {
            if ((r.flags & 0x0002) == 0 && !(preSolveLookahead && GetContinuousPhysics())) continue;
            auto c = contactFor(r.fixtureA, r.childA, r.fixtureB, r.childB);
            if (c.m_fixtureA is null || c.m_fixtureB is null || c.m_fixtureA.IsSensor() || c.m_fixtureB.IsSensor()) continue;
            fill(c, r);
            m_contactListener.PreSolve(c, &oldManifold);
            if (c.m_patchMask != 0)
                patches ~= dbx_contact_patch(r.fixtureA, r.childA, r.fixtureB, r.childB, cast(int) c.m_patchMask, c.IsEnabled() ? 1 : 0,
                                             c.m_friction, c.m_restitution, c.m_tangentSpeed);
        }
        if (patches.length > 0)
            check(dbx_world_patch_contacts(h, patches.ptr, cast(int) patches.length));
    }
    /// PostSolve from the records the island solve and the TOI sub-steps left (b2island.d:438-462)
    void postSolve()
    {
        const n = dbx_world_read_post_solve(h, null, 0);
        if (n <= 0) return;
        auto recs = new dbx_post_solve[n];
        const m = dbx_world_read_post_solve(h, recs.ptr, n);
        foreach (ref r; recs[0 .. (m < n ? m : n)])
        {
            b2ContactImpulse impulse;
            impulse.count = r.count;
            foreach (k; 0 .. r.count) { impulse.normalImpulses[k] = r.normalImpulses[k]; impulse.tangentImpulses[k] = r.tangentImpulses[k]; }
            m_contactListener.PostSolve(contactFor(r.fixtureA, r.childA, r.fixtureB, r.childB), &impulse);
        }
    }
    /// a user b2ContactFilter sees the contacts this step's broadphase created and vetoes them before they are ever solved
    void applyUserFilter()
    {
        if (m_contactFilter is null) return;
        const n = dbx_world_poll_new_contacts(h, null, 0);
        if (n <= 0) return;
        auto quad = new int[4 * n];
        const m = dbx_world_poll_new_contacts(h, quad.ptr, n);
        dbx_contact_patch[] veto;
        foreach (k; 0 .. (m < n ? m : n))
        {
            auto fa = quad[4 * k] in m_fixtures, fb = quad[4 * k + 2] in m_fixtures;
            if (fa is null || fb is null) continue;
            if (!m_contactFilter.ShouldCollide(*fa, *fb))
                veto ~= dbx_contact_patch(quad[4 * k], quad[4 * k + 1], quad[4 * k + 2], quad[4 * k + 3], DBX_PATCH_DESTROY, 0, 0, 0, 0);
        }
        if (veto.length > 0)
            check(dbx_world_patch_contacts(h, veto.ptr, cast(int) veto.length));
    }

    b2Vec2 m_gravity;
    b2Body* m_bodyList;
    b2Joint m_jointList;
    b2Contact m_contactList;
    int32 m_bodyCount, m_jointCount;
    b2Fixture*[int] m_fixtures;      /// fixture handle -> host object
    b2Contact[ulong] m_contacts;     /// (fixture, child) pair -> the view handed to the listeners
    b2DestructionListener m_destructionListener;
    b2ContactFilter m_contactFilter;
    b2ContactListener m_contactListener;
}
