#!/bin/bash
out=gpurun_out/r2c38; mkdir -p $out
python - <<'PY' 2>&1 | tee gpurun_out/r2c38/c3_profile.txt
import cProfile, pstats, sys, io, time
sys.path.insert(0,'scripts')
import bench_configs as bc
from scrutiny_b200.simulate import LineageSimulation
orig=LineageSimulation.run_all
calls=[]
def prof(self,*a,**k):
    pr=cProfile.Profile(); pr.enable(); r=orig(self,*a,**k); pr.disable(); calls.append(pr); return r
LineageSimulation.run_all=prof
bc.run('c3')
s=io.StringIO(); pstats.Stats(calls[-1],stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:6000])
PY
