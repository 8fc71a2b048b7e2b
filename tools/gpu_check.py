#!/usr/bin/env python3
"""Ad-hoc GPU parity sweep (dev tool): product vs oracle (and vs oracle/_ref when present)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import vamp_mvt_b200 as vmv
from oracle import pyoracle as po
from tests import scenes

def main():
    po.build()
    have_ref = po.ref_available()
    total_bad = 0
    for robot in ["panda", "ur5", "fetch", "baxter"]:
        R = getattr(vmv, robot); O = po.Oracle(robot)
        ref = po.Ref(robot) if have_ref else None
        keep = {"panda": 0.0, "ur5": 0.0, "fetch": 0.45, "baxter": 0.5}[robot]
        named = [("cage", scenes.sphere_cage()), ("table", scenes.table_shelf_scene()), ("box", scenes.box_scene())] if robot == "panda" else []
        named += [(f"rand{s}", scenes.random_scene(s, keep_out=keep)) for s in range(3)]
        named += [("empty", {"order": []})]
        for name, sc in named:
            env = scenes.build_product_env(sc)
            oenv = po.add_scene(po.OracleEnv(), scenes.packed(sc))
            N = 20000
            q = scenes.random_configs(robot, N, seed=1)
            # a quarter of the configurations leave the joint limits (exercises the unpruned self-collision path)
            rng = np.random.default_rng(5)
            q[: N // 4] += rng.uniform(-0.6, 0.6, size=(N // 4, q.shape[1])).astype(np.float32)
            from vamp_mvt_b200 import _lib
            _lib.lib().vmv_force_kernel_path(1)
            got1 = R.validate_batch(q, env)
            _lib.lib().vmv_force_kernel_path(0)
            t = time.time(); got = R.validate_batch(q, env); tg = time.time() - t
            if (got1 != got).any():
                print("   !! v1/v2 kernels disagree on", int((got1 != got).sum()), "configs")
            want = O.validate_configs(oenv, q)
            mism = np.nonzero(got != want)[0]
            cl = O.min_clearance(oenv, q[mism]) if len(mism) else np.zeros(0)
            out = int((np.abs(cl) > 1e-5).sum())
            msg = f"{robot:7s} {name:6s} cfg valid={want.mean():.3f} mism={len(mism)} outside_band={out}"
            if ref is not None:
                renv = po.add_scene(po.RefEnv(), scenes.packed(sc))
                wr = ref.validate_configs(renv, q, threads=8)
                msg += f" | vs_ref mism={(got != wr).sum()} oracle_vs_ref={(want != wr).sum()}"
            fk = R.fk_batch(q[:512]); fo = O.sphere_fk(q[:512])
            msg += f" fk_err={np.abs(fk - fo).max():.2e}"
            M = 4000
            a, b = scenes.random_edges(robot, M, seed=2)
            ge = R.validate_motion_batch(a, b, env)
            we = O.validate_edges(oenv, a, b)
            msg += f" | edges valid={we.mean():.3f} mism={(ge != we).sum()}"
            if ref is not None:
                wre = ref.validate_edges(renv, a, b, threads=8)
                msg += f" vs_ref={(ge != wre).sum()}"
            print(msg, flush=True)
            total_bad += out
    print("TOTAL outside band:", total_bad)

if __name__ == "__main__":
    main()
