// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics;

public
{
    import dbox.dynamics.b2body;
    import dbox.dynamics.b2fixture;
    import dbox.dynamics.b2timestep;
    import dbox.dynamics.b2world;
    import dbox.dynamics.b2worldcallbacks;
}
