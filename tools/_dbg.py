import os, sys, time
sys.path.insert(0,'.')
import numpy as np, torch
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import scenes
import bench
L=_lib.lib(); N=1<<20
env=scenes.build_product_env(scenes.table_shelf_scene()); h=env.handle
host=[scenes.random_configs("panda",N,seed=b) for b in range(8)]
qs=[torch.from_numpy(x).pin_memory() for x in host[:4]]
bits=torch.zeros((N+31)//32,dtype=torch.int32).pin_memory()
def e2e(tag):
    for i in range(3): _lib.check(L.vmv_validate_configs(vmv.panda.id,h,qs[i%4].data_ptr(),N,bits.data_ptr()))
    t0=time.perf_counter()
    for i in range(20): _lib.check(L.vmv_validate_configs(vmv.panda.id,h,qs[i%4].data_ptr(),N,bits.data_ptr()))
    print(f"{tag:40s} {(time.perf_counter()-t0)/20*1e3:.3f} ms", flush=True)
e2e("fresh")
dev=[torch.from_numpy(x).cuda() for x in host]
e2e("after 8 device batches")
db=torch.zeros((N+31)//32,dtype=torch.int32,device="cuda")
st=torch.cuda.current_stream().cuda_stream
for i in range(45): _lib.check(L.vmv_validate_configs_dev(vmv.panda.id,h,dev[i%8].data_ptr(),N,db.data_ptr(),st))
torch.cuda.synchronize()
e2e("after 45 launches on torch's stream")
s=bench.ClockSampler(0); s.start(); time.sleep(0.05); print(s.stop())
e2e("after the clock sampler")
time.sleep(1.0)
e2e("one second later")
