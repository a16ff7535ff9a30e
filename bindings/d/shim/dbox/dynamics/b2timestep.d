// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.b2timestep;

import dbox.common;

/// Profiling data, times in milliseconds (reference: dynamics/b2timestep.d:37-47); filled from CUDA events by dbx_world_profile.
struct b2Profile
{
    float32 step = 0;
    float32 collide = 0;
    float32 solve = 0;
    float32 solveInit = 0;
    float32 solveVelocity = 0;
    float32 solvePosition = 0;
    float32 broadphase = 0;
    float32 solveTOI = 0;
}
