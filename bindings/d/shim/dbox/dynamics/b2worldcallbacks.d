// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.b2worldcallbacks;

import dbox.common;
import dbox.collision;
import dbox.dynamics.b2fixture;
import dbox.dynamics.contacts.b2contact;
import dbox.dynamics.joints.b2joint;

/// reference: dynamics/b2worldcallbacks.d:34-50.  Called from b2World.DestroyBody for the joints / fixtures that go with a body.
class b2DestructionListener
{
    void SayGoodbye(b2Joint joint) { }
    void SayGoodbye(b2Fixture* fixture) { }
}

/// reference: :52-66.  The device evaluates the default rule itself; a user filter is consulted for the contacts the step
/// created (dbx_world_set_user_filter / dbx_world_poll_new_contacts) and vetoes through dbx_world_patch_contacts.
class b2ContactFilter
{
    bool ShouldCollide(b2Fixture* fixtureA, b2Fixture* fixtureB)
    {
        const b2Filter filterA = fixtureA.GetFilterData();
        const b2Filter filterB = fixtureB.GetFilterData();
        if (filterA.groupIndex == filterB.groupIndex && filterA.groupIndex != 0)
            return filterA.groupIndex > 0;
        return (filterA.maskBits & filterB.categoryBits) != 0 && (filterA.categoryBits & filterB.maskBits) != 0;
    }
}

/// reference: :71-76
struct b2ContactImpulse
{
    float32[b2_maxManifoldPoints] normalImpulses;
    float32[b2_maxManifoldPoints] tangentImpulses;
    int32 count;
}

/// reference: :87-130.  BeginContact / EndContact arrive after the step that produced them, in the reference's call order
/// (b2World.Step polls dbx_world_poll_contact_events); PreSolve runs between dbx_world_step_begin and dbx_world_step_end;
/// PostSolve is fed from dbx_world_read_post_solve.
class b2ContactListener
{
    void BeginContact(b2Contact contact) { }
    void EndContact(b2Contact contact) { }
    void PreSolve(b2Contact contact, const(b2Manifold)* oldManifold) { }
    void PostSolve(b2Contact contact, const(b2ContactImpulse)* impulse) { }
}

/// reference: :134-141
class b2QueryCallback
{
    abstract bool ReportFixture(b2Fixture* fixture);
}

/// reference: :145-156
class b2RayCastCallback
{
    abstract float32 ReportFixture(b2Fixture* fixture, b2Vec2 point, b2Vec2 normal, float32 fraction);
}
