import sys, time, ctypes as C; sys.path.insert(0,'.')
from dbox_b200 import scenes, lib, _abi as A
ga=lib.api()
copies=int(sys.argv[1]); steps=int(sys.argv[2]) if len(sys.argv)>2 else 300
caps=A.Caps(); caps.maxContacts=int(copies*620)
w,_=scenes.pyramid(api=ga, caps=caps)
t=time.time(); w.Replicate(copies); print("replicate %.2fs" % (time.time()-t), flush=True)
names=["collide","islands","colour","prepare","solve","sync","findnew","toi","clear"]
tot=C.c_float(); st=(C.c_float*9)()
done=0
while done<steps:
    k=min(50,steps-done)
    t=time.time(); rc=ga.world_time_steps(w._w,1/60.,8,3,k,0,C.byref(tot),st); wall=time.time()-t
    if rc<0: print("ERR",rc,ga.last_error()); break
    done+=k; c=w.counts()
    print("step %d: %.3f ms/step (wall %.3f) -> %.2f M world-steps/s | contacts %d touching %d awake %d colours %d | " % (done, tot.value/k, wall*1e3/k, copies/(tot.value/k)/1e3, c.contacts, c.touching, c.awakeBodies, c.colours)+" ".join("%s %.2f"%(names[i],st[i]) for i in range(9)), flush=True)
