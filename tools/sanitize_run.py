"""short GPU workload for compute-sanitizer (racecheck / initcheck): every shared-memory kernel of the step on small scenes --
k_solve_tiles (3,000-body jointed pile), k_solve_worlds (replicated pyramids), k_solve / k_colour / k_toi / k_query (pyramid,
bullet), with sleeping and continuous physics on"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbox_b200 import scenes, lib
from dbox_b200.world import b2BodyDef, b2CircleShape, b2_dynamicBody
ga = lib.api()
DT = 1 / 60.
n = int(os.environ.get("STEPS", "6"))
w, b, nj = scenes.pile(api=ga, n=3000, columns=100)
w.StepN(DT, 8, 3, n)
for k in range(3):                      # row-granular body sync (k_body_rows_get / _set) between steps
    b[7 + k].ApplyTorque(0.1); b[40].SetLinearVelocity((0.0, 0.1)); b[7 + k].GetPosition()
    w.Step(DT, 8, 3)
print("pile 3000 (tile solver):", w.counts().touching, "touching")
w.close()
w, b = scenes.islands_of_boxes(api=ga)          # twelve tiles without a boundary constraint between them
w.StepN(DT, 8, 3, n)
print("12 separate clusters (tile solver, no exchange):", w.counts().touching, "touching")
w.close()
w, b = scenes.pyramid(api=ga)
bd = b2BodyDef(); bd.type = b2_dynamicBody; bd.bullet = True; bd.position.Set(-30.0, 5.0); bd.linearVelocity.Set(200.0, 0.0)
bullet = w.CreateBody(bd); s = b2CircleShape(ga); s.m_radius = 0.25; bullet.CreateFixture(s, 20.0)
w.StepN(DT, 8, 3, 4 * n)
print("pyramid + bullet (k_solve, k_toi):", w.counts().touching, "touching")
w.close()
w, b = scenes.pyramid(api=ga)
ga.world_replicate(w._w, 24)
w.StepN(DT, 8, 3, 2 * n)
print("24 replicated pyramids (k_solve_worlds):", w.counts().touching, "touching")
w.close()
