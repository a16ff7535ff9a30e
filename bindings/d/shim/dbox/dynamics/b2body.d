// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.b2body;

import dbox.common;
import dbox.collision.shapes;
import dbox.dynamics.b2fixture;
import dbox.dynamics.b2world;
import dbox_b200_c;

/// reference: dynamics/b2body.d:35-47
enum b2BodyType
{
    b2_staticBody = 0,
    b2_kinematicBody,
    b2_dynamicBody
}

alias b2_staticBody    = b2BodyType.b2_staticBody;
alias b2_kinematicBody = b2BodyType.b2_kinematicBody;
alias b2_dynamicBody   = b2BodyType.b2_dynamicBody;

/// reference: :51-103
struct b2BodyDef
{
    b2BodyType type = b2BodyType.b2_staticBody;
    b2Vec2 position = b2Vec2(0, 0);
    float32 angle = 0;
    b2Vec2 linearVelocity = b2Vec2(0, 0);
    float32 angularVelocity = 0;
    float32 linearDamping = 0;
    float32 angularDamping = 0;
    bool allowSleep = true;
    bool awake = true;
    bool fixedRotation;
    bool bullet;
    bool active = true;
    void* userData;
    float32 gravityScale = 1.0;
}

/// reference: :106-1218.  A handle: the state lives in the device world; accessors read it through dbx_body_get_state (one bulk
/// device-to-host copy per step, cached by the library until the next step).
struct b2Body
{
    /// reference: :116-155
    b2Fixture* CreateFixture(const(b2FixtureDef)* def)
    {
        dbx_fixture_def fd;
        dbx_default_fixture_def(&fd);
        fd.friction = def.friction; fd.restitution = def.restitution; fd.density = def.density;
        fd.isSensor = def.isSensor ? 1 : 0;
        fd.categoryBits = def.filter.categoryBits; fd.maskBits = def.filter.maskBits; fd.groupIndex = def.filter.groupIndex;
        dbx_shape s = toDeviceShape(def.shape);
        const int id = dbx_fixture_create(worldHandle(), m_id, &fd, &s);
        if (id < 0)
            return null;
        auto f = new b2Fixture;
        f.m_id = id; f.m_body = &this; f.m_shape = cast(b2Shape) def.shape; f.m_density = def.density; f.m_friction = def.friction;
        f.m_restitution = def.restitution; f.m_isSensor = def.isSensor; f.m_filter = def.filter; f.m_userData = cast(void*) def.userData;
        f.m_next = m_fixtureList; m_fixtureList = f; ++m_fixtureCount;
        m_world.registerFixture(f);
        return f;
    }

    /// reference: :164-171
    b2Fixture* CreateFixture(const(b2Shape) shape, float32 density)
    {
        b2FixtureDef def;
        def.shape = cast(b2Shape) shape;
        def.density = density;
        return CreateFixture(&def);
    }

    /// reference: :179-247
    void DestroyFixture(b2Fixture* fixture)
    {
        dbx_fixture_destroy(worldHandle(), fixture.m_id);
        b2Fixture** node = &m_fixtureList;
        while (*node !is null)
        {
            if (*node == fixture) { *node = fixture.m_next; break; }
            node = &(*node).m_next;
        }
        m_world.unregisterFixture(fixture);
        --m_fixtureCount;
    }

    b2Transform GetTransform() const { const s = state(); b2Transform xf; xf.p = b2Vec2(s.p.x, s.p.y); xf.q.s = s.qs; xf.q.c = s.qc; return xf; }
    void SetTransform(b2Vec2 position, float32 angle) { dbx_body_set_transform(worldHandle(), m_id, position.x, position.y, angle); }   /// :261-285
    b2Vec2 GetPosition() const { const s = state(); return b2Vec2(s.p.x, s.p.y); }
    float32 GetAngle() const { return state().a; }
    b2Vec2 GetWorldCenter() const { const s = state(); return b2Vec2(s.c.x, s.c.y); }
    b2Vec2 GetLocalCenter() const { const s = state(); return b2Vec2(s.localCenter.x, s.localCenter.y); }
    b2Vec2 GetLinearVelocity() const { const s = state(); return b2Vec2(s.v.x, s.v.y); }
    void SetLinearVelocity(b2Vec2 v) { dbx_body_set_linear_velocity(worldHandle(), m_id, v.x, v.y); }                                   /// :322-335
    float32 GetAngularVelocity() const { return state().w; }
    void SetAngularVelocity(float32 w) { dbx_body_set_angular_velocity(worldHandle(), m_id, w); }                                       /// :346-359
    void ApplyForce(b2Vec2 force, b2Vec2 point, bool wake) { dbx_body_apply_force(worldHandle(), m_id, force.x, force.y, point.x, point.y, wake ? 1 : 0); }
    void ApplyForceToCenter(b2Vec2 force, bool wake) { const c = GetWorldCenter(); ApplyForce(force, c, wake); }
    void ApplyTorque(float32 torque, bool wake) { dbx_body_apply_torque(worldHandle(), m_id, torque, wake ? 1 : 0); }
    void ApplyLinearImpulse(b2Vec2 impulse, b2Vec2 point, bool wake) { dbx_body_apply_linear_impulse(worldHandle(), m_id, impulse.x, impulse.y, point.x, point.y, wake ? 1 : 0); }
    void ApplyAngularImpulse(float32 impulse, bool wake) { dbx_body_apply_angular_impulse(worldHandle(), m_id, impulse, wake ? 1 : 0); }
    float32 GetMass() const { return state().mass; }
    float32 GetInertia() const { const s = state(); return s.I + s.mass * (s.localCenter.x * s.localCenter.x + s.localCenter.y * s.localCenter.y); }   /// :490-493
    void GetMassData(b2MassData* data) const { const s = state(); data.mass = s.mass; data.I = GetInertia(); data.center = b2Vec2(s.localCenter.x, s.localCenter.y); }
    void SetMassData(const(b2MassData)* d) { dbx_body_set_mass_data(worldHandle(), m_id, d.mass, d.center.x, d.center.y, d.I); }          /// :502-540
    void ResetMassData() { dbx_body_reset_mass_data(worldHandle(), m_id); }                                                               /// :555-625
    b2Vec2 GetWorldPoint(b2Vec2 localPoint) const { return b2Mul(GetTransform(), localPoint); }
    b2Vec2 GetWorldVector(b2Vec2 localVector) const { return b2Mul(GetTransform().q, localVector); }
    b2Vec2 GetLocalPoint(b2Vec2 worldPoint) const { return b2MulT(GetTransform(), worldPoint); }
    b2Vec2 GetLocalVector(b2Vec2 worldVector) const { return b2MulT(GetTransform().q, worldVector); }
    b2Vec2 GetLinearVelocityFromWorldPoint(b2Vec2 worldPoint) const
    {
        const s = state();
        return b2Vec2(s.v.x, s.v.y) + b2Cross(s.w, worldPoint - b2Vec2(s.c.x, s.c.y));
    }
    b2Vec2 GetLinearVelocityFromLocalPoint(b2Vec2 localPoint) const { return GetLinearVelocityFromWorldPoint(GetWorldPoint(localPoint)); }
    float32 GetLinearDamping() const { return state().linearDamping; }
    void SetLinearDamping(float32 d) { dbx_body_set_linear_damping(worldHandle(), m_id, d); }
    float32 GetAngularDamping() const { return state().angularDamping; }
    void SetAngularDamping(float32 d) { dbx_body_set_angular_damping(worldHandle(), m_id, d); }
    float32 GetGravityScale() const { return state().gravityScale; }
    void SetGravityScale(float32 scale) { dbx_body_set_gravity_scale(worldHandle(), m_id, scale); }
    b2BodyType GetType() const { return cast(b2BodyType) state().type; }
    void SetType(b2BodyType type) { dbx_body_set_type(worldHandle(), m_id, cast(int) type); }                                             /// :718-775
    bool IsBullet() const { return (state().flags & 0x0008) != 0; }
    void SetBullet(bool flag) { dbx_body_set_bullet(worldHandle(), m_id, flag ? 1 : 0); }
    bool IsSleepingAllowed() const { return (state().flags & 0x0004) != 0; }
    void SetSleepingAllowed(bool flag) { dbx_body_set_sleeping_allowed(worldHandle(), m_id, flag ? 1 : 0); }
    bool IsAwake() const { return (state().flags & 0x0002) != 0; }
    void SetAwake(bool flag) { dbx_body_set_awake(worldHandle(), m_id, flag ? 1 : 0); }                                                  /// :827-846
    bool IsActive() const { return (state().flags & 0x0020) != 0; }
    void SetActive(bool flag) { dbx_body_set_active(worldHandle(), m_id, flag ? 1 : 0); }                                                /// :867-914
    bool IsFixedRotation() const { return (state().flags & 0x0010) != 0; }
    void SetFixedRotation(bool flag) { dbx_body_set_fixed_rotation(worldHandle(), m_id, flag ? 1 : 0); }                                 /// :924-945
    inout(b2Fixture)* GetFixtureList() inout { return m_fixtureList; }
    inout(b2Body)* GetNext() inout { return m_next; }
    void* GetUserData() const { return cast(void*) m_userData; }
    void SetUserData(void* data) { m_userData = data; }
    inout(b2World)* GetWorld() inout { return m_world; }

    dbx_world* worldHandle() const { return cast(dbx_world*) m_world.m_handle; }
    private dbx_body_state state() const
    {
        dbx_body_state s;
        dbx_body_get_state(worldHandle(), m_id, &s);
        return s;
    }

    int m_id = -1;               /// body handle of the C ABI
    b2World* m_world;
    b2Body* m_prev, m_next;
    b2Fixture* m_fixtureList;
    int32 m_fixtureCount;
    void* m_userData;
}
