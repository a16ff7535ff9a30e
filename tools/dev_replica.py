import sys; sys.path.insert(0,'.')
from dbox_b200 import scenes, lib
ga=lib.api()
copies=int(sys.argv[1]) if len(sys.argv)>1 else 16
w,_=scenes.pyramid(api=ga); w.Replicate(copies)
nb=w.counts().bodies//copies
for step in range(60):
    w.Step(1/60.,8,3)
    s,n=w.read_bodies()
    bad=[]
    for r in range(1,copies):
        for i in range(nb):
            a,b=s[i],s[r*nb+i]
            if (a.c.x,a.c.y,a.a,a.v.x,a.v.y,a.w,a.flags)!=(b.c.x,b.c.y,b.a,b.v.x,b.v.y,b.w,b.flags):
                bad.append((r,i,a.c.x-b.c.x,a.c.y-b.c.y,a.a-b.a,a.v.x-b.v.x,a.v.y-b.v.y,a.w-b.w,a.flags,b.flags))
    c=w.counts()
    if bad or step%10==0: print("step",step,"contacts",c.contacts,"touching",c.touching,"colours",c.colours,"mismatches:",len(bad),bad[:8])
    if bad: break
