// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.joints.b2mousejoint;

import dbox.common;
import dbox.dynamics.b2body;
import dbox.dynamics.joints.b2joint;
import dbox_b200_c;

/// reference: dynamics/joints/b2mousejoint.d:36-60 (same fields and defaults; the joint itself is solved on the device: dbx_solver.cuh, dbx_joints2.cuh)
class b2MouseJointDef : b2JointDef
{
    this() { type = b2JointType.e_mouseJoint; }

    b2Vec2 target = b2Vec2(0, 0);
    float32 maxForce = 0;
    float32 frequencyHz = 5.0f;
    float32 dampingRatio = 0.7f;

    override dbx_joint_def toDevice() const
    {
        dbx_joint_def d = super.toDevice();
        d.target = dbx_vec2(target.x, target.y);
        d.maxForce = maxForce;
        d.frequencyHz = frequencyHz;
        d.dampingRatio = dampingRatio;
        return d;
    }
}
