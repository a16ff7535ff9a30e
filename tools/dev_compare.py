import sys, time; sys.path.insert(0,'/root/repo')
from oracle import orc
from dbox_b200 import scenes, lib
import ctypes as C
oa = orc.api(); ga = lib.api()
print("devices", ga.device_count())
# hello world
wo, bo = scenes.hello_world(api=oa)
wg, bg = scenes.hello_world(api=ga)
for i in range(60):
    wo.Step(1/60., 6, 2); wg.Step(1/60., 6, 2)
    po, pg = bo.GetPosition(), bg.GetPosition()
    if i % 5 == 4 or abs(po.y-pg.y) > 1e-5:
        print(i, "orc %.6f %.6f | gpu %.6f %.6f %.6f" % (po.y, bo.GetAngle(), pg.y, bg.GetAngle(), pg.x), wg.counts().contacts, wg.counts().touching)
# pyramid
wo, bso = scenes.pyramid(api=oa); wg, bsg = scenes.pyramid(api=ga)
t=time.time()
for i in range(300):
    wo.Step(1/60., 8, 3); wg.Step(1/60., 8, 3)
    if i in (0,1,2,5,10,20,50,100,200,299):
        co, cg = wo.counts(), wg.counts()
        so,_ = wo.read_bodies(); sg,_ = wg.read_bodies()
        err = max(abs(so[k].c.y-sg[k].c.y)+abs(so[k].c.x-sg[k].c.x) for k in range(co.bodies))
        print(i, "orc c=%d t=%d aw=%d isl=%d | gpu c=%d t=%d aw=%d isl=%d col=%d | maxposerr %.3g" % (co.contacts, co.touching, co.awakeBodies, co.islands, cg.contacts, cg.touching, cg.awakeBodies, cg.islands, cg.colours, err))
print("time", time.time()-t)
p = wg.GetProfile(); print("profile ms step %.3f collide %.3f solveInit %.3f solveVel %.3f bp %.3f" % (p.step, p.collide, p.solveInit, p.solveVelocity, p.broadphase))
