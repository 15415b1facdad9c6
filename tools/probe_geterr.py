"""Device time of geterr over n energies: the plain chain of one warp against the block-parallel routine (development aid)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import numpy as np
pkg = entry.load_package()
rng = np.random.default_rng(1)
for n in (256, 1000, 1280, 1500, 3000, 5000, 10000, 40000):
    x = np.abs(rng.normal(1.5e9 / n, 0.5e9 / n, n)).astype(np.int32)
    out = {"n": n}
    for route, name in ((0, "chain_us"), (1, "block_us")):
        ts = []
        for _ in range(7):
            pkg.geterr(x, route=route)
            ts.append(1000 * pkg.mview_last_ms())
        out[name] = round(sorted(ts)[3], 2)
    print(json.dumps(out))
