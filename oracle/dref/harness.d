/// oracle/dref/harness.d -- TEST INFRASTRUCTURE: pins the CPU oracle to the real dbox.
///
/// Links against the UNMODIFIED reference library (built from /root/reference/src by oracle/dref/build.sh, outputs only
/// under oracle/_ref/) and prints, as JSON on stdout, the results of the reference's own fixed inputs:
///   examples/hello_world/hello_world.d:31-103      x, y, angle per step for 60 steps
///   examples/demo/tests/pyramid.d:39-78            contact / touching / awake counts at fixed steps, sleep step, top box
///   examples/demo/tests/polycollision.d:40-57      b2CollidePolygons on the demo's boxes (+ the same boxes brought into contact)
///   examples/demo/tests/distancetest.d:43-51       b2Distance
///   examples/demo/tests/timeofimpact.d:40-74       b2TimeOfImpact
/// Every float is printed as the 8 hex digits of its IEEE-754 bit pattern, so the comparison with the oracle
/// (tests/test_oracle.py::test_oracle_matches_reference_golden) is bit for bit.  Same keys as tests/golden/oracle_golden.json.
module harness;

import std.stdio;
import std.format;
import std.array;

import dbox;

string bits(float x)
{
    uint u = *cast(uint*)&x;
    return format("\"%08x\"", u);
}

string jarr(string[] items)
{
    return "[" ~ items.join(", ") ~ "]";
}

string helloWorld()
{
    b2Vec2 gravity = b2Vec2(0.0f, -10.0f);
    b2World world = b2World(gravity);
    b2BodyDef groundBodyDef;
    groundBodyDef.position.Set(0.0f, -10.0f);
    b2Body* groundBody = world.CreateBody(&groundBodyDef);
    b2PolygonShape groundBox = new b2PolygonShape;
    groundBox.SetAsBox(50.0f, 10.0f);
    groundBody.CreateFixture(groundBox, 0.0f);
    b2BodyDef bodyDef;
    bodyDef.type = b2_dynamicBody;
    bodyDef.position.Set(0.0f, 4.0f);
    b2Body* worldBody = world.CreateBody(&bodyDef);
    b2PolygonShape dynamicBox = new b2PolygonShape;
    dynamicBox.SetAsBox(1.0f, 1.0f);
    b2FixtureDef fixtureDef;
    fixtureDef.shape = dynamicBox;
    fixtureDef.density = 1.0f;
    fixtureDef.friction = 0.3f;
    worldBody.CreateFixture(&fixtureDef);
    float32 timeStep = 1.0f / 60.0f;
    string[] rows;
    for (int32 i = 0; i < 60; ++i)
    {
        world.Step(timeStep, 6, 2);
        b2Vec2 position = worldBody.GetPosition();
        float32 angle = worldBody.GetAngle();
        rows ~= jarr([bits(position.x), bits(position.y), bits(angle)]);
    }
    return jarr(rows);
}

string pyramid()
{
    b2World world = b2World(b2Vec2(0.0f, -10.0f));
    {
        // the demo base class creates a fixture-less static body first (examples/demo/framework/test.d:160-161)
        b2BodyDef bodyDef;
        world.CreateBody(&bodyDef);
    }
    {
        b2BodyDef bd;
        b2Body* ground = world.CreateBody(&bd);
        auto shape = new b2EdgeShape();
        shape.Set(b2Vec2(-40.0f, 0.0f), b2Vec2(40.0f, 0.0f));
        ground.CreateFixture(shape, 0.0f);
    }
    b2Body* last = null;
    {
        float32 a = 0.5f;
        auto shape = new b2PolygonShape();
        shape.SetAsBox(a, a);
        b2Vec2 x = b2Vec2(-7.0f, 0.75f);
        b2Vec2 y;
        b2Vec2 deltaX = b2Vec2(0.5625f, 1.25f);
        b2Vec2 deltaY = b2Vec2(1.125f, 0.0f);
        for (int32 i = 0; i < 20; ++i)
        {
            y = x;
            for (int32 j = i; j < 20; ++j)
            {
                b2BodyDef bd;
                bd.type = b2_dynamicBody;
                bd.position = y;
                b2Body* body_ = world.CreateBody(&bd);
                body_.CreateFixture(shape, 5.0f);
                last = body_;
                y += deltaY;
            }
            x += deltaX;
        }
    }
    string[] hist;
    int sleepStep = -1;
    for (int i = 0; i < 400; ++i)
    {
        world.Step(1.0f / 60.0f, 8, 3);
        int touching = 0, awake = 0;
        for (b2Contact c = world.GetContactList(); c; c = c.GetNext())
            if (c.IsTouching()) ++touching;
        for (b2Body* b = world.GetBodyList(); b; b = b.GetNext())
            if (b.IsAwake() && b.GetType() != b2_staticBody) ++awake;
        if (i == 0 || i == 10 || i == 20 || i == 50 || i == 100 || i == 399)
            hist ~= format("[%d, %d, %d, %d]", i, world.GetContactCount(), touching, awake);
        if (sleepStep < 0 && awake == 0) sleepStep = i;
    }
    b2Vec2 p = last.GetPosition();
    return format("{\"history\": %s, \"sleep_step\": %d, \"top\": %s}", jarr(hist), sleepStep,
                  jarr([bits(p.x), bits(p.y), bits(last.GetAngle())]));
}

string manifoldJson(ref b2Manifold m)
{
    string[] pts;
    for (int i = 0; i < m.pointCount; ++i)
        pts ~= format("[%s, %s, %d]", bits(m.points[i].localPoint.x), bits(m.points[i].localPoint.y), m.points[i].id.key);
    return format("{\"pointCount\": %d, \"type\": %d, \"localNormal\": %s, \"localPoint\": %s, \"points\": %s}", m.pointCount, cast(int)m.type,
                  jarr([bits(m.localNormal.x), bits(m.localNormal.y)]), jarr([bits(m.localPoint.x), bits(m.localPoint.y)]), jarr(pts));
}

string polyCollision(float bx, float by, float angle)
{
    auto polygonA = new b2PolygonShape();
    auto polygonB = new b2PolygonShape();
    b2Transform xfA, xfB;
    polygonA.SetAsBox(0.2f, 0.4f);
    xfA.Set(b2Vec2(0.0f, 0.0f), 0.0f);
    polygonB.SetAsBox(0.5f, 0.5f);
    xfB.Set(b2Vec2(bx, by), angle);
    b2Manifold manifold;
    b2CollidePolygons(&manifold, polygonA, xfA, polygonB, xfB);
    return manifoldJson(manifold);
}

string distanceTest()
{
    auto polygonA = new b2PolygonShape();
    auto polygonB = new b2PolygonShape();
    b2Transform xfA, xfB;
    xfA.SetIdentity();
    xfA.p.Set(0.0f, -0.2f);
    polygonA.SetAsBox(10.0f, 0.2f);
    xfB.Set(b2Vec2(12.017401f, 0.13678508f), -0.0109265f);
    polygonB.SetAsBox(2.0f, 0.1f);
    b2DistanceInput input;
    input.proxyA.Set(polygonA, 0);
    input.proxyB.Set(polygonB, 0);
    input.transformA = xfA;
    input.transformB = xfB;
    input.useRadii = true;
    b2SimplexCache cache;
    cache.count = 0;
    b2DistanceOutput output;
    b2Distance(&output, &cache, &input);
    return format("{\"distance\": %s, \"iterations\": %d, \"pointA\": %s, \"pointB\": %s}", bits(output.distance), output.iterations,
                  jarr([bits(output.pointA.x), bits(output.pointA.y)]), jarr([bits(output.pointB.x), bits(output.pointB.y)]));
}

string timeOfImpact()
{
    auto shapeA = new b2PolygonShape();
    auto shapeB = new b2PolygonShape();
    shapeA.SetAsBox(25.0f, 5.0f);
    shapeB.SetAsBox(2.5f, 2.5f);
    b2Sweep sweepA;
    sweepA.c0.Set(24.0f, -60.0f);
    sweepA.a0 = 2.95f;
    sweepA.c = sweepA.c0;
    sweepA.a = sweepA.a0;
    sweepA.localCenter.SetZero();
    b2Sweep sweepB;
    sweepB.c0.Set(53.474274f, -50.252514f);
    sweepB.a0 = 513.36676f;
    sweepB.c.Set(54.595478f, -51.083473f);
    sweepB.a = 513.62781f;
    sweepB.localCenter.SetZero();
    b2TOIInput input;
    input.proxyA.Set(shapeA, 0);
    input.proxyB.Set(shapeB, 0);
    input.sweepA = sweepA;
    input.sweepB = sweepB;
    input.tMax = 1.0f;
    b2TOIOutput output;
    b2TimeOfImpact(&output, &input);
    return format("{\"state\": %d, \"t\": %s}", cast(int)output.state, bits(output.t));
}

void main()
{
    writeln("{");
    writefln(" \"constants\": {\"angularSlop\": %s, \"maxAngularCorrection\": %s},", bits(b2_angularSlop), bits(b2_maxAngularCorrection));
    writefln(" \"polycollision\": %s,", polyCollision(19.345284f, 1.5632932f, 1.9160721f));
    writefln(" \"polycollision_touching\": %s,", polyCollision(0.55f, 0.3f, 1.9160721f));
    writefln(" \"distancetest\": %s,", distanceTest());
    writefln(" \"timeofimpact\": %s,", timeOfImpact());
    writefln(" \"hello_world\": %s,", helloWorld());
    writefln(" \"pyramid\": %s", pyramid());
    writeln("}");
}
