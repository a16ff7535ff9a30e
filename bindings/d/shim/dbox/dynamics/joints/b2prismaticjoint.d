// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.joints.b2prismaticjoint;

import dbox.common;
import dbox.dynamics.b2body;
import dbox.dynamics.joints.b2joint;
import dbox_b200_c;

/// reference: dynamics/joints/b2prismaticjoint.d:39-160 (same fields and defaults; the joint itself is solved on the device: dbx_solver.cuh, dbx_joints2.cuh)
class b2PrismaticJointDef : b2JointDef
{
    this() { type = b2JointType.e_prismaticJoint; }

    void Initialize(b2Body* bA, b2Body* bB, b2Vec2 anchor, b2Vec2 axis)
    {
        bodyA = bA; bodyB = bB;
        localAnchorA = bodyA.GetLocalPoint(anchor);
        localAnchorB = bodyB.GetLocalPoint(anchor);
        localAxisA = bodyA.GetLocalVector(axis);
        referenceAngle = bodyB.GetAngle() - bodyA.GetAngle();
    }
    b2Vec2 localAnchorA = b2Vec2(0, 0);
    b2Vec2 localAnchorB = b2Vec2(0, 0);
    b2Vec2 localAxisA = b2Vec2(1.0f, 0.0f);
    float32 referenceAngle = 0;
    bool enableLimit = false;
    float32 lowerTranslation = 0;
    float32 upperTranslation = 0;
    bool enableMotor = false;
    float32 maxMotorForce = 0;
    float32 motorSpeed = 0;

    override dbx_joint_def toDevice() const
    {
        dbx_joint_def d = super.toDevice();
        d.localAnchorA = dbx_vec2(localAnchorA.x, localAnchorA.y);
        d.localAnchorB = dbx_vec2(localAnchorB.x, localAnchorB.y);
        d.localAxisA = dbx_vec2(localAxisA.x, localAxisA.y);
        d.referenceAngle = referenceAngle;
        d.enableLimit = enableLimit ? 1 : 0;
        d.lowerTranslation = lowerTranslation;
        d.upperTranslation = upperTranslation;
        d.enableMotor = enableMotor ? 1 : 0;
        d.maxMotorForce = maxMotorForce;
        d.motorSpeed = motorSpeed;
        return d;
    }
}
