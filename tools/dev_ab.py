"""dev: A/B of DBX_DEBUG variants on the settled 100k pile, alternating, stage times from world_time_steps (200 steps, L2 flush)"""
import sys, os, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import ctypes as C
    from dbox_b200 import scenes, lib, state
    ga = lib.api()
    w, b, nj = scenes.pile(api=ga, n=100000, columns=1000)
    w.SetAllowSleeping(False)
    p = "/tmp/pile100k_settled.pkl"      # (scratch on the GPU box: gpurun_out/ is copied back and capped at 64 MiB)
    if os.path.exists(p):
        state.load(w, p); w.StepN(1 / 60., 8, 3, 30)
    else:
        w.StepN(1 / 60., 8, 3, 600)
        os.makedirs("gpurun_out", exist_ok=True); state.save(w, p)
    tot = C.c_float(); st = (C.c_float * 9)()
    out = []
    for rep in range(3):
        ga.world_time_steps(w._w, 1 / 60., 8, 3, 200, 1, C.byref(tot), st)
        out.append((round(tot.value / 200, 4), round(st[1], 4), round(st[2], 4), round(st[4], 4)))
    print(json.dumps(out))
else:
    variants = [int(x) for x in sys.argv[1:]] or [0, 2048]
    for rnd in range(2):
        for v in variants:
            env = dict(os.environ, DBX_DEBUG=str(v))
            r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
            print("DBX_DEBUG=%d" % v, "(ms/step, islands, colour+sort, solve) x3:", line, flush=True)
