// examples/hello_world/hello_world.d (reference lines 31-103) written straight against the C ABI: the smallest D program
// that steps a world on the B200.  It is what the forwarding bodies of INTEGRATION.md section 3 boil down to for this
// example; tests/test_gpu_parity.py::test_hello_world_trajectory runs the same calls through the Python mirror.
//
//   ldc2 -O hello_world_b200.d dbox_b200_c.d -L-L../../dbox_b200 -L-ldbox_b200 -L-rpath=../../dbox_b200
//
// (No D compiler exists in the image this repository was built in: the file is kept next to the generated binding as the
// integration recipe, not as a tested artefact.)
module hello_world_b200;

import core.stdc.stdio : printf;
import dbox_b200_c;

int main()
{
    dbx_world* world = dbx_world_create(0.0f, -10.0f, 0, null);          // b2World(gravity)
    if (world is null)
    {
        printf("dbx_world_create: %s\n", dbx_last_error());               // e.g. "no CUDA device": there is no CPU fallback
        return 1;
    }

    dbx_body_def groundBodyDef;
    dbx_default_body_def(&groundBodyDef);
    groundBodyDef.position = dbx_vec2(0.0f, -10.0f);
    int groundBody = dbx_body_create(world, &groundBodyDef);             // world.CreateBody(&groundBodyDef)

    dbx_shape groundBox;
    dbx_shape_set_box(&groundBox, 50.0f, 10.0f);                         // groundBox.SetAsBox(50, 10)
    dbx_fixture_def groundFixture;
    dbx_default_fixture_def(&groundFixture);
    groundFixture.density = 0.0f;
    dbx_fixture_create(world, groundBody, &groundFixture, &groundBox);   // groundBody.CreateFixture(groundBox, 0)

    dbx_body_def bodyDef;
    dbx_default_body_def(&bodyDef);
    bodyDef.type = DBX_DYNAMIC_BODY;
    bodyDef.position = dbx_vec2(0.0f, 4.0f);
    int worldBody = dbx_body_create(world, &bodyDef);

    dbx_shape dynamicBox;
    dbx_shape_set_box(&dynamicBox, 1.0f, 1.0f);
    dbx_fixture_def fixtureDef;
    dbx_default_fixture_def(&fixtureDef);
    fixtureDef.density = 1.0f;
    fixtureDef.friction = 0.3f;
    dbx_fixture_create(world, worldBody, &fixtureDef, &dynamicBox);      // worldBody.CreateFixture(&fixtureDef)

    const float timeStep = 1.0f / 60.0f;
    const int velocityIterations = 6;
    const int positionIterations = 2;
    for (int i = 0; i < 60; ++i)
    {
        if (dbx_world_step(world, timeStep, velocityIterations, positionIterations) < 0)   // world.Step(...)
        {
            printf("dbx_world_step: %s\n", dbx_last_error());
            return 1;
        }
        dbx_body_state s;
        dbx_body_get_state(world, worldBody, &s);                        // GetPosition() / GetAngle()
        printf("%4.2f %4.2f %4.2f\n", s.p.x, s.p.y, s.a);
    }
    dbx_world_destroy(world);
    return 0;
}
