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
ts=[buf[i] for i in range(n) if buf[i]]
d=[(ts[i+1]-ts[i])/1000. for i in range(len(ts)-1)]
print("phases", len(d), "total us %.1f" % sum(d))
print(" ".join("%.1f" % x for x in d))
c=w.counts(); print(c.colours, c.touching)

hb=(C.c_int32*1200)()
ga.world_debug_header(w._w, hb, 4800)
# Header: 16 ints, barrier, epoch, toiMin(8B), bounds[4] -> colourOff starts at int index 16+2+2+4 = 24
off=[hb[24+i] for i in range(14)]
print("colour sizes", [off[i+1]-off[i] for i in range(12)], "nSolve", hb[5], "nColours", hb[6])
