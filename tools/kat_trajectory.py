"""Informational (build container only: needs /root/reference): the oracle's schedule loop on the reference's
resource/imageB.png at the 800x450 window, next to the reference's own energy traces
(tests/opt_min_energy_edgeflip/strat/flip_set_maxenergy*.txt: "NT  sum of energies" at every topology change).
The first line is the soft pin of tests/test_kat_cpu.py (0.2 %); later lines depend on hundreds of chaotic
descent frames and on the exact (unrecorded) GL raster, so they are compared by eye: same NT sequence at the
start, energies within +-20 % early and within a few per cent from ~60 triangles on for the un-cooled run.
Output kept in profiles/r02_kat_trajectory.txt."""
import sys, time, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from PIL import Image
from oracle import oracle as O, schedule_oracle as S
import hostlib as H
for name, law, trace in (('imageB.png', 2, 'flip_set_maxenergy.txt'), ('imageB.png', 1, 'flip_set_maxenergy_dp_10.txt')):
    im = np.array(Image.open('/root/reference/resource/'+name).convert('RGBA'))
    h, w = im.shape[:2]
    img = O.resample(im, int(w/1.5), int(h/1.5))
    ratio = float(np.float32(w)/np.float32(h))
    O.set_dp_law(law)
    ups=[]
    t0=time.time()
    out = S.triangulate_loop(H.ref(), H.Tri, img, ratio, [10**9], '/tmp/kat.tri', 1500, updates=ups)
    O.set_dp_law(-1)
    ref=[l.split() for l in open('/root/reference/tests/opt_min_energy_edgeflip/strat/'+trace).read().splitlines()[:len(ups)]]
    print(trace, 'frames', out['frames'], 'time', round(time.time()-t0,1))
    for (nt,tot,fr),r in zip(ups,ref): print('  ours', nt, int(tot), 'frame', fr, '| ref', r[0], r[1], 'rel', round(tot/float(r[1])-1,4))
