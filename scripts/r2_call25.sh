#!/bin/bash
out=gpurun_out/r2c25; mkdir -p $out
python scripts/time_breakdown.py 12500 2>&1 | grep rep | tee $out/breakdown_c3.txt
python scripts/time_breakdown.py 12500 host 2>&1 | grep rep | tee $out/breakdown_c3_host.txt
python - <<'PY' 2>&1 | tee gpurun_out/r2c25/c1_profile.txt
import cProfile, pstats, sys, io, time
sys.argv=['x','c1']
sys.path.insert(0,'scripts')
import bench_configs as bc
import numpy as np
# profile the c1 run_all only: monkeypatch LineageSimulation.run_all
from scrutiny_b200.simulate import LineageSimulation
orig=LineageSimulation.run_all
calls=[]
def prof(self,*a,**k):
    pr=cProfile.Profile(); pr.enable(); r=orig(self,*a,**k); pr.disable(); calls.append(pr); return r
LineageSimulation.run_all=prof
bc.run('c1')
s=io.StringIO(); pstats.Stats(calls[-1],stream=s).sort_stats('cumulative').print_stats(25); print(s.getvalue()[:5000])
PY
