import sys, ctypes as C
sys.path.insert(0, '.')
from dbox_b200 import scenes, lib
ga = lib.api()
n, steps = int(sys.argv[1]), int(sys.argv[2])
t = scenes.Tumbler(api=ga, count=n)
out = set()
for k in range(steps):
    t.Step()
    if k % 10 == 9 or k == steps - 1:
        st, nb = t.world.read_bodies()
        for b in t.bodies:
            s = st[b.id]
            # position in the container frame
            import math
            ca = t.container._state()
            dx, dy = s.p.x - ca.p.x, s.p.y - ca.p.y
            lx, ly = ca.qc * dx + ca.qs * dy, -ca.qs * dx + ca.qc * dy
            if (abs(lx) > 10.6 or abs(ly) > 10.6) and b.id not in out:
                out.add(b.id)
                print("step %d body %d left: local (%.2f, %.2f) v (%.2f, %.2f) type %s" % (k, b.id, lx, ly, s.v.x, s.v.y, "circle" if (b.id - 3) & 1 else "box"))
c = t.world.counts()
print("out", len(out), "contacts", c.contacts, "touching", c.touching, "colours", c.colours, "conflicts", ga.world_debug_colour_conflicts(t.world._w))
