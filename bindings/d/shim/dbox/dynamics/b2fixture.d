// dbox_b200 D shim -- replaces the module of the same name in d-gamedev-team/dbox (src/dbox/dynamics/...): same public names and
// signatures, bodies forwarding to the extern(C) ABI of libdbox_b200.so (bindings/d/dbox_b200_c.d, generated from
// include/dbox_b200.h).  Build recipe: INTEGRATION.md section 3.  No D compiler exists in the image this repository is built
// in, so this file has not been compiled here; it is written against the reference's own declarations (cited per member).
module dbox.dynamics.b2fixture;

import dbox.common;
import dbox.collision.shapes;
import dbox.dynamics.b2body;
import dbox_b200_c;

/// reference: dynamics/b2fixture.d:32-45
struct b2Filter
{
    uint16 categoryBits = 0x0001;
    uint16 maskBits = 0xFFFF;
    int16 groupIndex = 0;
}

/// reference: :48-72
struct b2FixtureDef
{
    b2Shape shape;
    void* userData;
    float32 friction = 0.2;
    float32 restitution = 0;
    float32 density = 0;
    bool isSensor;
    b2Filter filter;
}

/// the device record of a reference shape object (collision/shapes/*.d): geometry in the body frame
dbx_shape toDeviceShape(const(b2Shape) shape)
{
    dbx_shape s;
    s.radius = shape.m_radius;
    if (auto c = cast(const(b2CircleShape)) shape)
    {
        s.type = DBX_SHAPE_CIRCLE;
        s.p = dbx_vec2(c.m_p.x, c.m_p.y);
    }
    else if (auto e = cast(const(b2EdgeShape)) shape)
    {
        s.type = DBX_SHAPE_EDGE;
        s.v0 = dbx_vec2(e.m_vertex0.x, e.m_vertex0.y); s.v1 = dbx_vec2(e.m_vertex1.x, e.m_vertex1.y);
        s.v2 = dbx_vec2(e.m_vertex2.x, e.m_vertex2.y); s.v3 = dbx_vec2(e.m_vertex3.x, e.m_vertex3.y);
        s.hasV0 = e.m_hasVertex0 ? 1 : 0; s.hasV3 = e.m_hasVertex3 ? 1 : 0;
    }
    else if (auto p = cast(const(b2PolygonShape)) shape)
    {
        s.type = DBX_SHAPE_POLYGON;
        s.centroid = dbx_vec2(p.m_centroid.x, p.m_centroid.y);
        s.count = p.m_count;
        foreach (i; 0 .. p.m_count)
        {
            s.vertices[i] = dbx_vec2(p.m_vertices[i].x, p.m_vertices[i].y);
            s.normals[i] = dbx_vec2(p.m_normals[i].x, p.m_normals[i].y);
        }
    }
    else if (auto ch = cast(const(b2ChainShape)) shape)
    {
        s.type = DBX_SHAPE_CHAIN;
        s.chainVertices = cast(const(dbx_vec2)*) ch.m_vertices;      // b2Vec2 and dbx_vec2 are both two floats; copied by fixture_create
        s.chainCount = ch.m_count;
        s.prevVertex = dbx_vec2(ch.m_prevVertex.x, ch.m_prevVertex.y); s.nextVertex = dbx_vec2(ch.m_nextVertex.x, ch.m_nextVertex.y);
        s.hasPrev = ch.m_hasPrevVertex ? 1 : 0; s.hasNext = ch.m_hasNextVertex ? 1 : 0;
    }
    return s;
}

/// reference: :84-300.  A handle: the fixture lives in the device world.
struct b2Fixture
{
    b2Shape.Type GetType() const { return m_shape.GetType(); }
    inout(b2Shape) GetShape() inout { return m_shape; }          /// the host copy made at creation (b2fixture.d:380 clones it too)
    bool IsSensor() const { return m_isSensor; }
    void SetSensor(bool sensor) { m_isSensor = sensor; dbx_fixture_set_sensor(m_body.worldHandle(), m_id, sensor ? 1 : 0); }
    b2Filter GetFilterData() const { return m_filter; }
    void SetFilterData(b2Filter filter)
    {
        m_filter = filter;
        dbx_fixture_set_filter(m_body.worldHandle(), m_id, filter.categoryBits, filter.maskBits, filter.groupIndex);
    }
    void Refilter() { SetFilterData(m_filter); }
    inout(b2Body)* GetBody() inout { return m_body; }
    inout(b2Fixture)* GetNext() inout { return m_next; }
    void* GetUserData() const { return cast(void*) m_userData; }
    void SetUserData(void* data) { m_userData = data; }
    float32 GetDensity() const { return m_density; }
    void SetDensity(float32 density) { m_density = density; dbx_fixture_set_density(m_body.worldHandle(), m_id, density); }
    float32 GetFriction() const { return m_friction; }
    void SetFriction(float32 friction) { m_friction = friction; dbx_fixture_set_friction(m_body.worldHandle(), m_id, friction); }
    float32 GetRestitution() const { return m_restitution; }
    void SetRestitution(float32 restitution) { m_restitution = restitution; dbx_fixture_set_restitution(m_body.worldHandle(), m_id, restitution); }
    bool TestPoint(b2Vec2 p)
    {
        int id = m_id, inside = 0;
        dbx_vec2 q = dbx_vec2(p.x, p.y);
        dbx_world_test_points(m_body.worldHandle(), &id, &q, 1, &inside);
        return inside != 0;
    }

    int m_id = -1;               /// fixture handle of the C ABI
    b2Body* m_body;
    b2Fixture* m_next;
    b2Shape m_shape;
    float32 m_density = 0, m_friction = 0, m_restitution = 0;
    bool m_isSensor;
    b2Filter m_filter;
    void* m_userData;
}
