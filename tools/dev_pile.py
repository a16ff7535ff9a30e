import sys, time, ctypes as C; sys.path.insert(0,'.')
from dbox_b200 import scenes, lib, _abi as A
ga = lib.api()
n=int(sys.argv[1]); cols=int(sys.argv[2]); steps=int(sys.argv[3]) if len(sys.argv)>3 else 600
t=time.time(); w,b,nj=scenes.pile(api=ga, n=n, columns=cols); print("build %.1fs joints %d" % (time.time()-t, nj), flush=True)
w.SetAllowSleeping(False)
names=["collide","islands","colour","prepare","solve","sync","findnew","toi","clear"]
tot=C.c_float(); st=(C.c_float*9)()
done=0
while done < steps:
    k=min(50, steps-done)
    t=time.time(); rc=ga.world_time_steps(w._w, 1/60., 8, 3, k, 0, C.byref(tot), st); wall=time.time()-t
    if rc<0: print("ERR", rc, ga.last_error()); break
    done+=k
    c=w.counts()
    print("step %d: %.3f ms/step (wall %.3f) contacts %d touching %d awake %d islands %d colours %d | " % (done, tot.value/k, wall*1e3/k, c.contacts, c.touching, c.awakeBodies, c.islands, c.colours) + " ".join("%s %.3f" % (names[i], st[i]) for i in range(9)), flush=True)
print("conflicts", ga.world_debug_colour_conflicts(w._w), "launches", ga.world_launch_count(w._w))
s,nb=w.read_bodies()
ys=[s[i].c.y for i in range(nb) if s[i].type==2]
print("max y %.2f min y %.3f" % (max(ys), min(ys)))
