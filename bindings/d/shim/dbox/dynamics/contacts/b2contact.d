// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.contacts.b2contact;

import dbox.common;
import dbox.collision;
import dbox.dynamics.b2fixture;
import dbox_b200_c;

/// reference: dynamics/contacts/b2contact.d:77-205.  A view of one device contact, rebuilt from dbx_world_read_contacts /
/// the event records each time the host looks (the contact itself lives in the device pair cache).
class b2Contact
{
    inout(b2Manifold)* GetManifold() inout { return &m_manifold; }
    bool IsTouching() const { return (m_flags & 0x0002) != 0; }
    bool IsEnabled() const { return (m_flags & 0x0004) != 0; }
    /// takes effect through dbx_world_patch_contacts when called from PreSolve (b2contact.d:137-149)
    void SetEnabled(bool flag) { if (flag) m_flags |= 0x0004; else m_flags &= ~0x0004; m_patchMask |= DBX_PATCH_ENABLED; }
    inout(b2Contact) GetNext() inout { return m_next; }
    inout(b2Fixture)* GetFixtureA() inout { return m_fixtureA; }
    int32 GetChildIndexA() const { return m_indexA; }
    inout(b2Fixture)* GetFixtureB() inout { return m_fixtureB; }
    int32 GetChildIndexB() const { return m_indexB; }
    void SetFriction(float32 friction) { m_friction = friction; m_patchMask |= DBX_PATCH_FRICTION; }
    float32 GetFriction() const { return m_friction; }
    void ResetFriction() { SetFriction(b2MixFriction(m_fixtureA.GetFriction(), m_fixtureB.GetFriction())); }
    void SetRestitution(float32 restitution) { m_restitution = restitution; m_patchMask |= DBX_PATCH_RESTITUTION; }
    float32 GetRestitution() const { return m_restitution; }
    void ResetRestitution() { SetRestitution(b2MixRestitution(m_fixtureA.GetRestitution(), m_fixtureB.GetRestitution())); }
    void SetTangentSpeed(float32 speed) { m_tangentSpeed = speed; m_patchMask |= DBX_PATCH_TANGENT_SPEED; }
    float32 GetTangentSpeed() const { return m_tangentSpeed; }

    uint m_flags, m_patchMask;
    b2Contact m_next;
    b2Fixture* m_fixtureA, m_fixtureB;
    int32 m_indexA, m_indexB;
    b2Manifold m_manifold;
    float32 m_friction = 0, m_restitution = 0, m_tangentSpeed = 0;
}

/// reference: b2contact.d:32-42
float32 b2MixFriction(float32 friction1, float32 friction2) { return b2Sqrt(friction1 * friction2); }
float32 b2MixRestitution(float32 restitution1, float32 restitution2) { return restitution1 > restitution2 ? restitution1 : restitution2; }
