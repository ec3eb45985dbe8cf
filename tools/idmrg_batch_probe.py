"""Probe of the batched iDMRG growth step (development aid): python tools/idmrg_batch_probe.py [n_chains ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g  # noqa: E402

g.build()
import bench  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [64, 256]:
    t = time.time()
    out = bench.idmrg_batch_leg(n, 128)
    print(json.dumps(out), f"{time.time() - t:.1f} s", flush=True)
