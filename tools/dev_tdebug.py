import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbox_b200 import scenes, lib
ga = lib.api()
w, b, nj = scenes.pile(api=ga, n=3000, columns=100)
w.SetAllowSleeping(False)
hb = (C.c_int32 * 2400)()
for k in range(int(os.environ.get("STEPS", "12"))):
    w.Step(1 / 60., 8, 3)
    c = w.counts()
    ga.world_debug_header(w._w, hb, 9600)
    print(k, "contacts", c.contacts, "touching", c.touching, "nSolve", hb[5], "nColours", hb[6], "nTileB/G", hb[1124], hb[1125], "err", hb[10], "conflicts", ga.world_debug_colour_conflicts(w._w), flush=True)
