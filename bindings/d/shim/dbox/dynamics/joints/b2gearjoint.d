// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.joints.b2gearjoint;

import dbox.common;
import dbox.dynamics.joints.b2joint;
import dbox_b200_c;

/// reference: dynamics/joints/b2gearjoint.d:36-58 (joint1 / joint2 must be revolute or prismatic joints)
class b2GearJointDef : b2JointDef
{
    this() { type = b2JointType.e_gearJoint; }

    b2Joint joint1;
    b2Joint joint2;
    float32 ratio = 1.0f;

    override dbx_joint_def toDevice() const
    {
        dbx_joint_def d = super.toDevice();
        d.joint1 = joint1 is null ? -1 : joint1.m_id;
        d.joint2 = joint2 is null ? -1 : joint2.m_id;
        d.bodyA = joint1 is null ? -1 : joint1.m_bodyB.m_id;     // b2gearjoint.d:84-160: the gear's own pair is (joint1.bodyB, joint2.bodyB)
        d.bodyB = joint2 is null ? -1 : joint2.m_bodyB.m_id;
        d.ratio = ratio;
        return d;
    }
}
