import sys; sys.path.insert(0, '.')
from dbox_b200 import scenes, lib
from oracle import orc
ga, oa = lib.api(), orc.api()
g, _ = scenes.pyramid(api=ga, count=8); o, _ = scenes.pyramid(api=oa, count=8)
g.EnableContactEvents(4096); o.EnableContactEvents(4096)
K = lambda e: (e[0], e[1], e[3], e[4], e[5], e[6], e[7], e[8])
for step in range(60):
    g.Step(1 / 60., 8, 3); o.Step(1 / 60., 8, 3)
    eg, eo = sorted(map(K, g.PollContactEvents())), sorted(map(K, o.PollContactEvents()))
    if eg != eo:
        print("step", step, "gpu-only", [e for e in eg if e not in eo], "oracle-only", [e for e in eo if e not in eg])
    elif eg:
        print("step", step, "ok", len(eg))
