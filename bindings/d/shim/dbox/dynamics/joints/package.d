// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.joints;

public
{
    import dbox.dynamics.joints.b2joint;
    import dbox.dynamics.joints.b2distancejoint;
    import dbox.dynamics.joints.b2frictionjoint;
    import dbox.dynamics.joints.b2gearjoint;
    import dbox.dynamics.joints.b2motorjoint;
    import dbox.dynamics.joints.b2mousejoint;
    import dbox.dynamics.joints.b2prismaticjoint;
    import dbox.dynamics.joints.b2pulleyjoint;
    import dbox.dynamics.joints.b2revolutejoint;
    import dbox.dynamics.joints.b2ropejoint;
    import dbox.dynamics.joints.b2weldjoint;
    import dbox.dynamics.joints.b2wheeljoint;
}
