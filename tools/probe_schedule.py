"""Progress of the full-size triangulate schedule under a time cap (development aid): writes the
program's own log (NT + total energy at every mesh update, like the reference prints) to gpurun_out/."""
import os, sys, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as entry
import importlib
pkg = entry.load_package()
synth = importlib.import_module("tpose_b200.synth")
w, h, frames = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
native = "native" in sys.argv[4:]
auto = "auto" in sys.argv[4:]
out = os.path.join(ROOT, "gpurun_out"); os.makedirs(out, exist_ok=True)
img = synth.make_image(w, h, seed=1000)
p = os.path.join(out, f"sched_{w}x{h}.ppm")
with open(p, "wb") as f:
    f.write(b"P6\n%d %d\n255\n" % (w, h)); f.write(np.ascontiguousarray(img[..., :3]).tobytes())
if os.path.exists(p + ".tri"): os.remove(p + ".tri")
t0 = time.time()
r = subprocess.run([os.path.join(ROOT, "t-pose_b200", "bin", "triangulate"), "-i", p] + (["-native"] if native else []) + (["-stepscale", "auto"] if auto else []) + ["-max", "3000", "-frames", frames],
                   capture_output=True, text=True)
log = r.stdout.splitlines()
open(os.path.join(out, f"sched_{w}x{h}.log"), "w").write(r.stdout[-200000:])
print("wall", time.time() - t0, "lines", len(log))
print("\n".join(log[:3])); print("..."); print("\n".join(log[-6:]))
os.remove(p)
