// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.joints.b2distancejoint;

import dbox.common;
import dbox.dynamics.b2body;
import dbox.dynamics.joints.b2joint;
import dbox_b200_c;

/// reference: dynamics/joints/b2distancejoint.d:37-92 (same fields and defaults; the joint itself is solved on the device: dbx_solver.cuh, dbx_joints2.cuh)
class b2DistanceJointDef : b2JointDef
{
    this() { type = b2JointType.e_distanceJoint; }

    /// reference: :54-64
    void Initialize(b2Body* b1, b2Body* b2, b2Vec2 anchor1, b2Vec2 anchor2)
    {
        bodyA = b1; bodyB = b2;
        localAnchorA = bodyA.GetLocalPoint(anchor1);
        localAnchorB = bodyB.GetLocalPoint(anchor2);
        b2Vec2 d = anchor2 - anchor1;
        length = d.Length();
    }
    b2Vec2 localAnchorA = b2Vec2(0, 0);
    b2Vec2 localAnchorB = b2Vec2(0, 0);
    float32 length = 1.0f;
    float32 frequencyHz = 0;
    float32 dampingRatio = 0;

    override dbx_joint_def toDevice() const
    {
        dbx_joint_def d = super.toDevice();
        d.localAnchorA = dbx_vec2(localAnchorA.x, localAnchorA.y);
        d.localAnchorB = dbx_vec2(localAnchorB.x, localAnchorB.y);
        d.length = length;
        d.frequencyHz = frequencyHz;
        d.dampingRatio = dampingRatio;
        return d;
    }
}
