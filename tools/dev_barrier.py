import sys; sys.path.insert(0,'.')
from dbox_b200 import lib
a=lib.api()
for blocks in (148, 296, 74, 37, 16, 1):
    for threads in (512, 128):
        print(blocks, threads, "%.3f us" % a.debug_barrier_us(0, blocks, threads, 2000))
