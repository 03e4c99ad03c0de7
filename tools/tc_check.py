import os, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
os.environ["LST_NCC_ENGINE"] = sys.argv[1] if len(sys.argv) > 1 else "tc"
from laser_spot_track_b200 import LaserSpotTrack4
from laser_spot_track_b200.synth import StackConfig, SyntheticStack, CONFIGS
from oracle import lst_oracle as O
import util
cfg = CONFIGS["A"]; st = SyntheticStack(cfg)
frames = [st.frame(i) for i in range(0, 9)]
op = util.oracle_params(cfg, match_intensity=False)
otr = O.OracleTracker(op); otr.set_reference(frames[0], st.ref_xy); want = otr.run(frames[1:])
with LaserSpotTrack4(cfg.width, cfg.height, 8, templSize=32, sArea=32, matchIntensity=False) as t:
    t.set_reference(frames[0], st.ref_xy)
    got = t.track(np.stack(frames[1:]))
    bad = 0
    for i,(g,w) in enumerate(zip(got,want)):
        for k in range(5):
            a,b = g.tm[k], w.tm[k]
            if (a.ix,a.iy,a.score)!=(b.ix,b.iy,b.score):
                bad += 1
                if bad < 6: print("frame",i,"tpl",k,"got",(a.ix,a.iy,a.score),"want",(b.ix,b.iy,b.score))
    print("engine", os.environ["LST_NCC_ENGINE"], "mismatches", bad, "of", len(want)*5)
