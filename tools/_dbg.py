import os, sys, time
sys.path.insert(0,'.')
import numpy as np
import vamp_mvt_b200 as vmv
from tests import scenes, workloads
env=scenes.build_product_env(scenes.table_shelf_scene())
q=scenes.random_configs("panda",1<<16,seed=0)
vmv.panda.validate_batch(q,env)   # CUDA warm
pts=workloads.synth_pointcloud(100_000,0.55)
for k in range(3):
    e=vmv.Environment(); t0=time.perf_counter(); e.add_capt_pointcloud(pts,0.03,0.24,0.0025); print("build",k,(time.perf_counter()-t0)*1e3,"ms",flush=True)
