import sys, ctypes as C; sys.path.insert(0,'.')
from dbox_b200 import scenes, lib
ga = lib.api()
import os
w,b,nj=scenes.pile(api=ga, n=100000, columns=1000, joints=os.environ.get('NOJ') is None)
w.SetAllowSleeping(False)
w.StepN(1/60.,8,3,300)
buf=(C.c_uint64*4096)()
ga.world_debug_phase_times(w._w, buf, 4096)
w.StepN(1/60.,8,3,3)
n=ga.world_debug_phase_times(w._w, buf, 4096)
ts=[buf[i] for i in range(min(n,512)) if buf[i]]
d=[(ts[i+1]-ts[i])/1000. for i in range(len(ts)-1)]
print("phases", len(d), "total us %.1f" % sum(d))
print(" ".join("%.1f" % x for x in d))
c=w.counts(); print(c.colours, c.touching)

hb=(C.c_int32*1200)()
ga.world_debug_header(w._w, hb, 4800)
# Header: 18 ints, barrier, epoch, toiMin(8B), bounds[4] -> colourOff starts at int index 16+2+2+4 = 24
off=[hb[26+i] for i in range(14)]
print("colour sizes", [off[i+1]-off[i] for i in range(12)], "nSolve", hb[5], "nColours", hb[6])

import os
if int(os.environ.get("DBX_DEBUG","0")) & 2:
    p0 = int(os.environ["DBX_DEBUG"]) >> 8
    print('nonzero window entries', sum(1 for i in range(512,4096) if buf[i]))
    base = min(buf[512 + i] for i in range(8*148*2) if buf[512+i])
    for ph in range(8):
        arr=[(buf[512+(ph*148+b)*2]-base)/1000. for b in range(148)]
        rel=[(buf[512+(ph*148+b)*2+1]-base)/1000. for b in range(148)]
        ja=arr[:17]; ca=arr[17:]
        print("bar %d: joint-blk arrive %.2f..%.2f  contact-blk arrive %.2f..%.2f (median %.2f) | release %.2f..%.2f" % (p0+ph, min(ja), max(ja), min(ca), max(ca), sorted(ca)[len(ca)//2], min(rel), max(rel)))

tt=[buf[3000+i] for i in range(64) if buf[3000+i]]
print("toi marks (us):", " ".join("%.1f" % ((tt[i+1]-tt[i])/1000.) for i in range(len(tt)-1)))
