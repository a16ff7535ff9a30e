// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.joints.b2pulleyjoint;

import dbox.common;
import dbox.dynamics.b2body;
import dbox.dynamics.joints.b2joint;
import dbox_b200_c;

/// reference: dynamics/joints/b2pulleyjoint.d:41-104 (same fields and defaults; the joint itself is solved on the device: dbx_solver.cuh, dbx_joints2.cuh)
class b2PulleyJointDef : b2JointDef
{
    this() { type = b2JointType.e_pulleyJoint; }

    void Initialize(b2Body* bA, b2Body* bB, b2Vec2 groundA, b2Vec2 groundB, b2Vec2 anchorA, b2Vec2 anchorB, float32 r)
    {
        bodyA = bA; bodyB = bB;
        groundAnchorA = groundA; groundAnchorB = groundB;
        localAnchorA = bodyA.GetLocalPoint(anchorA);
        localAnchorB = bodyB.GetLocalPoint(anchorB);
        b2Vec2 dA = anchorA - groundA;
        lengthA = dA.Length();
        b2Vec2 dB = anchorB - groundB;
        lengthB = dB.Length();
        ratio = r;
        collideConnected = true;
    }
    b2Vec2 groundAnchorA = b2Vec2(-1.0f, 1.0f);
    b2Vec2 groundAnchorB = b2Vec2(1.0f, 1.0f);
    b2Vec2 localAnchorA = b2Vec2(-1.0f, 0.0f);
    b2Vec2 localAnchorB = b2Vec2(1.0f, 0.0f);
    float32 lengthA = 0;
    float32 lengthB = 0;
    float32 ratio = 1.0f;

    override dbx_joint_def toDevice() const
    {
        dbx_joint_def d = super.toDevice();
        d.groundAnchorA = dbx_vec2(groundAnchorA.x, groundAnchorA.y);
        d.groundAnchorB = dbx_vec2(groundAnchorB.x, groundAnchorB.y);
        d.localAnchorA = dbx_vec2(localAnchorA.x, localAnchorA.y);
        d.localAnchorB = dbx_vec2(localAnchorB.x, localAnchorB.y);
        d.lengthA = lengthA;
        d.lengthB = lengthB;
        d.ratio = ratio;
        return d;
    }
}
