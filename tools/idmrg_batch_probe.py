import sys, json, time
sys.path.insert(0, '.')
import __graft_entry__ as g; g.build()
import bench
for n in (64, 256):
    t=time.time(); out = bench.idmrg_batch_leg(n, 128); print(json.dumps(out), time.time()-t, flush=True)
