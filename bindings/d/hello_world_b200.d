// The scene of examples/hello_world (a 100 x 20 static slab, a 2 x 2 box dropped from y = 4, sixty steps at 60 Hz with 6 velocity and
// 2 position iterations) written straight against the C ABI: the smallest D program that steps a world on the B200.  It is what the
// forwarding bodies of INTEGRATION.md section 3 boil down to for that example; tests/test_gpu_parity.py::test_hello_world_trajectory
// runs the same calls through the Python mirror and compares the trajectory with the oracle.
//
//   ldc2 -O hello_world_b200.d dbox_b200_c.d -L-L../../dbox_b200 -L-ldbox_b200 -L-rpath=../../dbox_b200
//
// (No D compiler exists in the image this repository was built in: the file is kept next to the generated binding as the
// integration recipe, not as a tested artefact.)
module hello_world_b200;

import core.stdc.stdio : printf;
import dbox_b200_c;

/// one body with one box fixture; returns the body id (or a negative DBX_E_* code)
int addBox(dbx_world* w, int bodyType, float x, float y, float halfW, float halfH, float density, float friction)
{
    dbx_body_def bd;
    dbx_default_body_def(&bd);
    bd.type = bodyType;
    bd.position = dbx_vec2(x, y);
    const int body = dbx_body_create(w, &bd);
    if (body < 0)
        return body;

    dbx_shape box;
    dbx_shape_set_box(&box, halfW, halfH);
    dbx_fixture_def fd;
    dbx_default_fixture_def(&fd);
    fd.density = density;
    fd.friction = friction;
    const int fixture = dbx_fixture_create(w, body, &fd, &box);
    return fixture < 0 ? fixture : body;
}

int main()
{
    dbx_world* w = dbx_world_create(0.0f, -10.0f, /*device*/ 0, /*caps*/ null);
    if (w is null)
    {
        printf("dbx_world_create: %s\n", dbx_last_error());   // e.g. "no CUDA device": there is no CPU fallback
        return 1;
    }
    addBox(w, DBX_STATIC_BODY, 0.0f, -10.0f, 50.0f, 10.0f, 0.0f, 0.2f);
    const int falling = addBox(w, DBX_DYNAMIC_BODY, 0.0f, 4.0f, 1.0f, 1.0f, 1.0f, 0.3f);

    foreach (step; 0 .. 60)
    {
        if (dbx_world_step(w, 1.0f / 60.0f, 6, 2) < 0)
        {
            printf("dbx_world_step: %s\n", dbx_last_error());
            return 1;
        }
        dbx_body_state s;
        dbx_body_get_state(w, falling, &s);
        printf("%4.2f %4.2f %4.2f\n", s.p.x, s.p.y, s.a);
    }
    dbx_world_destroy(w);
    return 0;
}
