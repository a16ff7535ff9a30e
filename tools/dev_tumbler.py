import sys, time, ctypes as C
sys.path.insert(0, '.')
from dbox_b200 import scenes, lib
from oracle import orc
count = int(sys.argv[1]) if len(sys.argv) > 1 else 800
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
spawn = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ga = lib.api()
t = scenes.Tumbler(api=ga, count=count)
t0 = time.time()
for k in range(steps):
    try:
        t.Step(spawn_per_step=spawn)
    except Exception as e:
        print("step", k, "FAILED", e)
        hb = (C.c_int32 * 64)(); ga.world_debug_header(t.world._w, hb, 256)
        names = "cHigh nFree nContacts nMoved nPairs nSolve nColours nTouching nUncoloured nUncoloured2 error nTomb nIslands nAwake toiEvents nEvents tailStart nCtEvents".split()
        print({n: hb[i] for i, n in enumerate(names)}); break
    if (k + 1) % 200 == 0:
        c = t.world.counts()
        print("step %d bodies %d contacts %d touching %d colours %d  %.2f ms/step wall" % (k + 1, c.bodies, c.contacts, c.touching, c.colours, 1e3 * (time.time() - t0) / (k + 1)), flush=True)
# oracle side for the same number of steps
oa = orc.api()
to = scenes.Tumbler(api=oa, count=count)
t0 = time.time()
for k in range(steps):
    to.Step(spawn_per_step=spawn)
print("oracle %.2f ms/step wall" % (1e3 * (time.time() - t0) / steps))
co = to.world.counts(); cg = t.world.counts()
print("oracle bodies %d contacts %d touching %d | gpu contacts %d touching %d" % (co.bodies, co.contacts, co.touching, cg.contacts, cg.touching))
import math
def stats(w, bodies):
    ys = [b.GetPosition() for b in bodies]
    inside = sum(1 for p in ys if abs(p.x) < 10.6 and -0.6 < p.y < 20.6)
    return inside, sum(p.y for p in ys) / max(1, len(ys))
print("gpu inside/meanY", stats(t.world, t.bodies), "oracle", stats(to.world, to.bodies))
print("container angle gpu %.5f oracle %.5f" % (t.container.GetAngle(), to.container.GetAngle()))
