import sys; sys.path.insert(0,'.')
from dbox_b200 import lib
a=lib.api()
for stores in (0, 1):
    for mode in (0, 1, 2):
        print("stores", stores, "mode", mode, "148x512: %.3f us" % a.debug_barrier_us(0, 148, 512, (stores*10+mode)*100000 + 2000))
