"""dev: phase stamps of k_solve_tiles on the 100k pile (CTA 0 and the middle CTA) + class sizes"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbox_b200 import scenes, lib
ga = lib.api()
n = int(os.environ.get("N", "100000")); cols = int(os.environ.get("COLS", "1000"))
w, b, nj = scenes.pile(api=ga, n=n, columns=cols, joints=os.environ.get('NOJ') is None)
w.SetAllowSleeping(False)
w.StepN(1 / 60., 8, 3, int(os.environ.get("SETTLE", "400")))
buf = (C.c_uint64 * 4096)()
ga.world_debug_phase_times(w._w, buf, 4096)
w.StepN(1 / 60., 8, 3, 3)
m = ga.world_debug_phase_times(w._w, buf, 4096)
for base, name in ((0, "CTA 0"), (1024, "middle CTA")):
    ts = [buf[base + i] for i in range(1000) if buf[base + i]]
    d = [(ts[i + 1] - ts[i]) / 1000. for i in range(len(ts) - 1)]
    print(name, "marks", len(ts), "total us %.1f" % sum(d))
    print("  prologue: offsets %.1f  recolour %.1f  bodies %.1f  rows wait %.1f" % tuple(d[:4]))
    rest = d[4:]
    # forward sweeps: 4 intervals each (L, publish+GB, B, GB+readback)
    for k in range(0, min(len(rest), 4 * 9), 4):
        print("  sweep %d: L %.1f  pub+GB %.1f  B %.1f  GB+rb %.1f" % tuple([k // 4] + rest[k:k + 4]) if len(rest) >= k + 4 else rest[k:])
    print("  after the velocity passes (store impulses + integrate + GB | position passes ... | write-back + sleep):", " ".join("%.1f" % x for x in rest[36:]))
for base, name in ((2048, "CTA 0"), (2048 + 128, "middle CTA")):
    print(name, "velocity pass 4, per local colour: (colour, items, cycles own item, cycles waiting at the barrier)")
    print("  ", [(int(buf[base + 4 * k + 3]), int(buf[base + 4 * k + 2]), int(buf[base + 4 * k]), int(buf[base + 4 * k + 1])) for k in range(12)])
st = [buf[3300 + i] for i in range(148) if buf[3300 + i]]
print('CTA start skew: %d CTAs, last - first = %.1f us; sorted offsets (us):' % (len(st), (max(st) - min(st)) / 1000.), ' '.join('%.0f' % ((x - min(st)) / 1000.) for x in sorted(st)[::12]))
hb = (C.c_int32 * 2400)()
ga.world_debug_header(w._w, hb, 9600)
c = w.counts()
print("colours", c.colours, "touching", c.touching, "tail ints of header (.., nTileB, nTileG):", list(hb[1120:1140]))
tot = C.c_float(); st = (C.c_float * 9)()
ga.world_time_steps(w._w, 1 / 60., 8, 3, 50, 1, C.byref(tot), st)
print("ms/step %.4f" % (tot.value / 50), "stages", ["%.3f" % x for x in st])
tt = [buf[3000 + i] for i in range(64) if buf[3000 + i]]
print("k_toi marks of thread 0 (us between):", " ".join("%.1f" % ((tt[i + 1] - tt[i]) / 1000.) for i in range(len(tt) - 1)))
for base, name in ((3700, "CTA 0 vel"), (3732, "middle vel"), (3764, "CTA 0 pos"), (3796, "middle pos")):
    print(name, "boundary colours (items, joints, cycles own, cycles wait):", [(int(buf[base + 4 * k + 2]), int(buf[base + 4 * k + 3]), int(buf[base + 4 * k]), int(buf[base + 4 * k + 1])) for k in range(6)])
for base, name in ((2304, "CTA 0"), (2304 + 128, "middle CTA")):
    print(name, "position pass 2, per local colour: (colour, items, cycles own item, cycles waiting at the barrier)")
    print("  ", [(int(buf[base + 4 * k + 3]), int(buf[base + 4 * k + 2]), int(buf[base + 4 * k]), int(buf[base + 4 * k + 1])) for k in range(12)])
print("prologue cycles (thread 0) after: rbd/rpc, joints staged, sW patched, bodies loaded, bodies synced:", [int(buf[3900 + i]) for i in range(12)], [int(buf[3916 + i]) for i in range(12)])
arr = [(buf[3500 + i], i) for i in range(148) if buf[3500 + i]]
if arr:
    t0 = min(a for a, _ in arr)
    print("arrival at the first grid barrier (position pass 2), us after the first CTA; (tile: local colours, local rows, boundary colours x1000 + items):")
    for a, i in sorted(arr)[::-1][:12] + sorted(arr)[:4]:
        v = int(buf[3940 + i]); hi = v >> 32; v &= 0xFFFFFFFF
        print("   nbr bodies %d joints staged %d (local: %d) boundary rows staged %d" % (hi & 1023, (hi >> 10) & 1023, (hi >> 20) & 1, hi >> 21), end="")
        print("   tile %3d  +%.1f us  colours %d rows %d boundary %d" % (i, (a - t0) / 1000., v // 100000000, (v // 10000) % 10000, v % 10000))
if arr:
    his = [int(buf[3940 + i]) >> 32 for _, i in arr]
    print("max neighbour bodies", max(h & 1023 for h in his), "tiles without joints in shared memory", sum(1 for h in his if not (h >> 20) & 1), "min/max boundary rows staged", min(h >> 21 for h in his), max(h >> 21 for h in his))
if arr:
    t0 = min(a for a, _ in arr)
    print("tile: smid, arrival us")
    print(" ".join("%d:%d:%.0f" % (i, int(buf[3650 + i]) - 1, (a - t0) / 1000.) for a, i in sorted(arr, key=lambda x: x[1])))
