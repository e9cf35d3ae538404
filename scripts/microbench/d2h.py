import torch, time, numpy as np
t0=time.time()
try:
    h=torch.empty((1000000,3000),dtype=torch.int32,pin_memory=True); ok=True
except RuntimeError as e:
    print("pin failed",e); h=torch.empty((1000000,3000),dtype=torch.int32); ok=False
print("alloc pinned=%s %.2fs"%(ok,time.time()-t0))
d=torch.zeros((1000000,3000),dtype=torch.int32,device="cuda")
for i in range(3):
    torch.cuda.synchronize(); t0=time.time(); h.copy_(d,non_blocking=True); torch.cuda.synchronize(); dt=time.time()-t0
    print("D2H 12GB %.3fs %.1f GB/s"%(dt,12/dt))
# chunked 256MB like scr_counts
s=torch.cuda.Stream()
for i in range(2):
    torch.cuda.synchronize(); t0=time.time()
    with torch.cuda.stream(s):
        for c0 in range(0,1000000,22369):
            h[c0:c0+22369].copy_(d[c0:c0+22369],non_blocking=True)
    torch.cuda.synchronize(); dt=time.time()-t0
    print("chunked D2H %.3fs %.1f GB/s"%(dt,12/dt))
import subprocess; print(subprocess.run("nvidia-smi topo -m | head -5; numactl -H 2>/dev/null | head -5; nproc; free -g | head -2",shell=True,capture_output=True,text=True).stdout)
