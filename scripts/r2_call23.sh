#!/bin/bash
out=gpurun_out/r2c24; mkdir -p $out; rm -f $out/trace.txt
( timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_bench_configs.py -q -m gpu -x -k "dense or c4" ) > $out/pytest_dense.log 2>&1; tail -3 $out/pytest_dense.log
SCR_DENSE_TRACE=$out/trace.txt SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/trace.so timeout 300 python scripts/quick_bench.py --net dense500 --cells 1000 --steps 400 --alpha 1.0 --reps 2 2>&1 | grep "rep\|rror"
python - <<'PY'
import numpy as np
rows=[l.split() for l in open('gpurun_out/r2c24/trace.txt') if not l.startswith('#')]
a=np.array(rows,dtype=np.int64)
a=a[-400:]          # last launch
a=a[50:350]
names=['poll done','A issued','first stage full','last MMA issued','tfull seen','stores issued','fence+bar done','release done']
t0=a[:,0]
print('step period (poll done -> next poll done): median %.0f ns'%np.median(np.diff(t0)))
for k in range(1,8):
    print('%-18s +%6.0f ns (median, from poll done)'%(names[k], np.median(a[:,k]-t0)))
print('release -> next poll done: %.0f ns'%np.median(a[1:,0]-a[:-1,7]))
PY
