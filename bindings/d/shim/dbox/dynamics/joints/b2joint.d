// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.joints.b2joint;

import dbox.common;
import dbox.dynamics.b2body;
import dbox.dynamics.b2world;
import dbox_b200_c;

/// reference: dynamics/joints/b2joint.d:28-42 (the numbering is the C ABI's DBX_JOINT_* too)
enum b2JointType
{
    e_unknownJoint,
    e_revoluteJoint,
    e_prismaticJoint,
    e_distanceJoint,
    e_pulleyJoint,
    e_mouseJoint,
    e_gearJoint,
    e_wheelJoint,
    e_weldJoint,
    e_frictionJoint,
    e_ropeJoint,
    e_motorJoint
}
alias e_unknownJoint = b2JointType.e_unknownJoint;
alias e_revoluteJoint = b2JointType.e_revoluteJoint;
alias e_prismaticJoint = b2JointType.e_prismaticJoint;
alias e_distanceJoint = b2JointType.e_distanceJoint;
alias e_pulleyJoint = b2JointType.e_pulleyJoint;
alias e_mouseJoint = b2JointType.e_mouseJoint;
alias e_gearJoint = b2JointType.e_gearJoint;
alias e_wheelJoint = b2JointType.e_wheelJoint;
alias e_weldJoint = b2JointType.e_weldJoint;
alias e_frictionJoint = b2JointType.e_frictionJoint;
alias e_ropeJoint = b2JointType.e_ropeJoint;
alias e_motorJoint = b2JointType.e_motorJoint;

/// reference: :46-51
enum b2LimitState
{
    e_inactiveLimit,
    e_atLowerLimit,
    e_atUpperLimit,
    e_equalLimits
}

/// reference: :77-93.  Every concrete definition fills the one flat record the C ABI takes.
class b2JointDef
{
    b2JointType type = b2JointType.e_unknownJoint;
    void* userData;
    b2Body* bodyA;
    b2Body* bodyB;
    bool collideConnected;

    /// the device record: common part here, the rest by the subclasses
    dbx_joint_def toDevice() const
    {
        dbx_joint_def d;
        dbx_default_joint_def(&d, cast(int) type);
        d.bodyA = bodyA is null ? -1 : bodyA.m_id;
        d.bodyB = bodyB is null ? -1 : bodyB.m_id;
        d.collideConnected = collideConnected ? 1 : 0;
        return d;
    }
}

/// reference: :97-260.  A handle to the device joint; the run-time setters of the concrete joint classes go through
/// dbx_joint_set_params (motor speed / torque / on-off, limits, springs, lengths, offsets) with the reference's wake-ups.
class b2Joint
{
    this(int id, b2JointDef def, b2World* world)
    {
        m_id = id; m_def = def; m_type = def.type; m_bodyA = def.bodyA; m_bodyB = def.bodyB; m_world = world;
        m_collideConnected = def.collideConnected; m_userData = def.userData;
    }
    b2JointType GetType() const { return m_type; }
    inout(b2Body)* GetBodyA() inout { return m_bodyA; }
    inout(b2Body)* GetBodyB() inout { return m_bodyB; }
    inout(b2Joint) GetNext() inout { return m_next; }
    void* GetUserData() const { return cast(void*) m_userData; }
    void SetUserData(void* data) { m_userData = data; }
    bool IsActive() const { return m_bodyA.IsActive() && m_bodyB.IsActive(); }
    bool GetCollideConnected() const { return m_collideConnected; }

    /// accumulated impulses and limit state as the last step left them
    dbx_joint_state deviceState() const
    {
        auto all = new dbx_joint_state[m_id + 1];
        dbx_world_read_joints(cast(dbx_world*) m_world.m_handle, all.ptr, m_id + 1);
        return all[m_id];
    }
    /// b2RevoluteJoint / b2PrismaticJoint / b2WheelJoint .SetMotorSpeed (e.g. b2revolutejoint.d:268-273): wakes both bodies
    void SetMotorSpeed(float32 speed)
    {
        dbx_joint_def d = m_def.toDevice();
        d.motorSpeed = speed;
        dbx_joint_set_params(cast(dbx_world*) m_world.m_handle, m_id, &d, DBX_JP_MOTOR_SPEED);
    }
    /// b2MouseJoint.SetTarget (b2mousejoint.d:112-120)
    void SetTarget(b2Vec2 target) { dbx_joint_set_target(cast(dbx_world*) m_world.m_handle, m_id, target.x, target.y); }

    int m_id = -1;               /// joint handle of the C ABI
    b2JointDef m_def;
    b2JointType m_type;
    b2Joint m_prev, m_next;
    b2Body* m_bodyA, m_bodyB;
    b2World* m_world;
    bool m_collideConnected;
    void* m_userData;
}
