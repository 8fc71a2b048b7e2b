import ctypes as C, os, sys
sys.path.insert(0,'.')
import numpy as np
import vamp_mvt_b200 as vmv
from vamp_mvt_b200 import _lib
from tests import workloads
L=_lib.lib()
rng=np.random.default_rng(5)
clouds=[rng.random((n,3)).astype(np.float32)*np.float32([1.5,1.5,1.0]) for n in (1,2,3,5,64,65,1000,2048,2049,5000)]
clouds.append(workloads.synth_pointcloud(100_000,0.55))
for pts in clouds:
    dg=[]
    for host in (False,True):
        if host: os.environ["VMV_CAPT_HOST_BUILD"]="1"
        else: os.environ.pop("VMV_CAPT_HOST_BUILD",None)
        env=vmv.Environment(); env.add_capt_pointcloud(pts,0.03,0.24,vmv.POINT_RADIUS)
        out=(C.c_uint64*4)(); _lib.check(L.vmv_env_capt_digest(env.handle,0,out)); dg.append(tuple(out))
    print(len(pts), [a==b for a,b in zip(*dg)])
