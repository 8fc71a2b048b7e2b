#!/usr/bin/env python3
"""bench.py -- headline benchmark of the validation hot path (BASELINE.json).

Workload at every N (weak scaling, per-GPU work fixed): BASELINE config 2 -- 2^20 random Panda
configurations (uniform in the joint bounds, seeded) validated against a synthetic
MotionBenchMaker-style table + shelf scene of cuboids and cylinders (tests/scenes.py).  One "step"
is one pass of the hot path (FK + environment + self collision -> 1 verdict bit per configuration)
over one batch; consecutive steps rotate through 8 distinct resident batches (8 x 29 MB > the 126 MB
L2) so no step re-reads inputs the previous step left in L2.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
    python bench.py --impl reference ...                           the reference's own CPU path

Prints ONE JSON line (rank 0).  `value` is timed on the device with inputs resident in HBM; `e2e`
goes through the host-buffer C-ABI call (pinned host inputs, H2D + kernels + D2H inside the timed
region); `roofline` is the FP32-pipe roofline of the dominant kernel; `cpu_baseline` is the
reference's own AVX2 code (oracle/_ref, compiled from the reference sources) on this box's host
cores.  The secondary C3 workload (2^18 edges, box scene) is reported under "edges".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

N_CONFIGS = 1 << 20
N_EDGES = 1 << 18
N_BATCHES = 8
ROBOT = "panda"

# source-level flop counts of the reference program: FK 1147 + 14 trig x 32;
# sphere 11, capsule 28, z-capsule 18, cuboid 33, z-cuboid 24, heightfield 18, self pair 11,
# max_extent 8 per swept sphere, 1 per visited object
FLOPS = dict(fk=1147 + 14 * 32, extent=8, visited=1, t_sphere=11, t_capsule=28, t_zcapsule=18, t_cuboid=33,
             t_zcuboid=24, t_heightfield=18, t_self=11)


def algorithmic_flops(counters: dict) -> float:
    return float(sum(FLOPS[k] * counters[k] for k in FLOPS))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def reference_arm(args) -> int:
    """The reference's own CPU implementation of the path (oracle/_ref = the reference headers
    compiled in place) with all host threads, on the same workload, metric and unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pyoracle as po
    from tests import scenes

    scene = scenes.table_shelf_scene()
    threads = po.host_threads()
    q = scenes.random_configs(ROBOT, N_CONFIGS, seed=0)
    line = {"impl": "reference", "metric": "panda_collision_checked_configs_per_sec", "unit": "configs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: 2^20 random Panda configs vs synthetic MBM-style table/shelf scene (13 cuboids + 4 cylinders)",
                       "robot": ROBOT, "n_configs": N_CONFIGS}}
    if po.ref_available():
        ref = po.Ref(ROBOT)
        env = po.add_scene(po.RefEnv(), scenes.packed(scene))
        kind = "reference"
        run = lambda: ref.time_configs(env, q, threads, reps=1)
    else:
        oracle = po.Oracle(ROBOT)
        env = po.add_scene(po.OracleEnv(), scenes.packed(scene))
        kind, threads = "port", 1
        sub = q[: 1 << 17]

        def run():
            t0 = time.perf_counter()
            oracle.validate_configs(env, sub)
            return (time.perf_counter() - t0) * (len(q) / len(sub))

    for _ in range(args.warmup):
        run()
    times = [run() for _ in range(args.steps)]
    t = float(np.mean(times))
    value = N_CONFIGS / t
    line.update({
        "value": value, "ms_per_step": t * 1e3,
        "cpu_baseline": {"value": value, "unit": "configs/s", "cores": threads, "kind": kind,
                         "sample": f"each step = the full 2^20-config batch, {threads} std::thread static partition"},
        "e2e": {"value": value, "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    print(json.dumps(line), flush=True)
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-gather", action="store_true", help="skip the NCCL verdict-bitmask all-gather at N>1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-edges", action="store_true", help="skip the secondary edges workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    import vamp_mvt_b200 as vmv
    from vamp_mvt_b200 import _lib
    from tests import scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _lib.lib()
    _lib.check(L.vmv_set_device(local_rank))
    robot = getattr(vmv, ROBOT)
    dof = robot.dimension()
    stream = torch.cuda.current_stream().cuda_stream

    scene = scenes.table_shelf_scene()
    env = scenes.build_product_env(scene)
    h_env = env.handle

    # ---- resident inputs: N_BATCHES distinct batches per rank ------------------------------------
    host_batches = [scenes.random_configs(ROBOT, N_CONFIGS, seed=1000 * rank + b) for b in range(N_BATCHES)]
    dev_batches = [torch.from_numpy(h).cuda() for h in host_batches]
    n_words = (N_CONFIGS + 31) // 32
    # two verdict buffers: at N > 1 the all-gather of step i (NCCL stream) overlaps the kernel of step
    # i + 1 (compute stream); a buffer is reused only after its gather has completed
    bits2 = [torch.zeros(n_words, dtype=torch.int32, device="cuda") for _ in range(2)]
    gathered2 = [torch.zeros(world * n_words, dtype=torch.int32, device="cuda") for _ in range(2)] if world > 1 else None
    pending = [None, None]
    dev_bits = bits2[0]

    def gather_async(i: int):
        if gathered2 is not None and not args.no_gather:
            pending[i % 2] = dist.all_gather_into_tensor(gathered2[i % 2], bits2[i % 2], async_op=True)

    def wait_slot(i: int):
        if pending[i % 2] is not None:
            pending[i % 2].wait()
            pending[i % 2] = None

    def step(i: int):
        qd = dev_batches[i % N_BATCHES]
        wait_slot(i)
        _lib.check(L.vmv_validate_configs_dev(robot.id, h_env, qd.data_ptr(), N_CONFIGS, bits2[i % 2].data_ptr(), stream))
        gather_async(i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.vmv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # dominant-kernel time: events tightly around each launch, accumulated over the timed region
    kernel_events = []
    barrier()
    e0.record()
    for i in range(args.steps):
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ka.record()
        qd = dev_batches[i % N_BATCHES]
        wait_slot(i)
        _lib.check(L.vmv_validate_configs_dev(robot.id, h_env, qd.data_ptr(), N_CONFIGS, bits2[i % 2].data_ptr(), stream))
        kb.record()
        kernel_events.append((ka, kb))
        gather_async(i)
    wait_slot(0), wait_slot(1)  # every gather has completed (the compute stream waits on them) before the end event
    e1.record()
    barrier()
    launches = int(L.vmv_launch_count() - launches0)
    total_ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * N_CONFIGS / (ms_per_step * 1e-3)
    valid_fraction = float(np.unpackbits(dev_bits.cpu().numpy().view(np.uint8)).mean())

    # ---- e2e: host buffers through the public C-ABI call (H2D + kernels + D2H inside) ------------
    pin_q = [torch.from_numpy(h).pin_memory() for h in host_batches[:4]]
    pin_bits = torch.zeros(n_words, dtype=torch.int32).pin_memory()

    def e2e_step(i: int):
        _lib.check(L.vmv_validate_configs(robot.id, h_env, pin_q[i % 4].data_ptr(), N_CONFIGS, pin_bits.data_ptr()))

    for i in range(3):
        e2e_step(i)
    barrier()
    e2e_steps = max(5, args.steps // 2)
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * N_CONFIGS / e2e_s

    # ---- secondary workload: C3 edges ------------------------------------------------------------
    edges = None
    if not args.no_edges:
        box_env = scenes.build_product_env(scenes.box_scene())
        hb = box_env.handle
        a_h, b_h = scenes.random_edges(ROBOT, N_EDGES, seed=rank)
        a_d, b_d = torch.from_numpy(a_h).cuda(), torch.from_numpy(b_h).cuda()
        ebits = torch.zeros((N_EDGES + 31) // 32, dtype=torch.int32, device="cuda")
        run = lambda: _lib.check(L.vmv_validate_edges_dev(robot.id, hb, a_d.data_ptr(), b_d.data_ptr(), N_EDGES, 0, ebits.data_ptr(), stream))
        for _ in range(3):
            run()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(5, args.steps // 4)
        s0.record()
        for _ in range(reps):
            run()
        s1.record()
        barrier()
        ems = s0.elapsed_time(s1) / reps
        if world > 1:
            t = torch.tensor([ems], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        # the same through the host-buffer call (pinned host endpoints, H2D + kernel + D2H inside)
        pa, pb = torch.from_numpy(a_h).pin_memory(), torch.from_numpy(b_h).pin_memory()
        pe = torch.zeros((N_EDGES + 31) // 32, dtype=torch.int32).pin_memory()
        run_h = lambda: _lib.check(L.vmv_validate_edges(robot.id, hb, pa.data_ptr(), pb.data_ptr(), N_EDGES, 0, pe.data_ptr()))
        for _ in range(2):
            run_h()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            run_h()
        barrier()
        e2e_edges_s = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([e2e_edges_s], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_edges_s = float(t.item())
        edges = {"workload": "C3: 2^18 Panda edges (|b-a| ~ U(0.25,2) rad), resolution 32, synthetic box scene",
                 "value": world * N_EDGES / (ems * 1e-3), "unit": "edges/s", "ms_per_step": ems,
                 "e2e": {"value": world * N_EDGES / e2e_edges_s, "unit": "edges/s", "h2d_bytes_per_step": 2 * N_EDGES * dof * 4,
                         "d2h_bytes_per_step": (N_EDGES + 31) // 32 * 4, "api": "vmv_validate_edges (host pointers, pinned)"},
                 "valid_fraction": float(np.unpackbits(ebits.cpu().numpy().view(np.uint8)).mean())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline (FP32 pipe) + cpu baseline: rank 0 only -----------------------------------------
    from oracle import pyoracle as po

    roofline, cpu_baseline = None, None
    try:
        po.build()
        oracle = po.Oracle(ROBOT)
        oenv = po.add_scene(po.OracleEnv(), scenes.packed(scene))
        sample = host_batches[0][: 1 << 16]
        cnt = po.Counters()
        oracle.validate_configs(oenv, sample, cnt)
        flops_per_config = algorithmic_flops(cnt.as_dict()) / len(sample)
        peaks = {}
        try:
            peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        peak_tflops = n_sm * 128 * 2 * sm_max * 1e6 / 1e12
        achieved = flops_per_config * N_CONFIGS / (kernel_ms * 1e-3) / 1e12
        bytes_per_launch = N_CONFIGS * (dof * 4 + 1 / 8)
        # DRAM bytes per launch and pipe activity of the same kernel on the same workload, from the
        # committed `ncu --set full` capture (profiles/): a number taken under the profiler, reported
        # as such, never mixed into the timings
        captured = {}
        try:
            captured = json.loads((REPO / "profiles" / "kernel_metrics.json").read_text()).get("k_validate_configs_v4", {})
        except Exception:
            pass
        roofline = {
            "bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
            "traffic": captured.get("dram_bytes_per_launch"),
            "traffic_source": captured.get("capture"),
            "pipe_fma_pct_of_peak": captured.get("pipe_fma_pct"), "pipe_alu_pct_of_peak": captured.get("pipe_alu_pct"),
            "kernel": "k_validate_configs_v4<panda, u32 masks, verdict tables> (vmv_kernels_v4.cuh)", "kernel_ms": kernel_ms,
            "algorithmic_flops_per_config": flops_per_config,
            "peak_source": f"FP32 pipe: {n_sm} SMs x 128 lanes x 2 flop x {sm_max:.0f} MHz (clocks.max.sm); no tensor-core work on this path",
            "hbm": {"algorithmic_bytes_per_launch": bytes_per_launch,
                    "achieved_gbs": bytes_per_launch / (kernel_ms * 1e-3) / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs"), "peak_source": "MEASURED_PEAKS.json" if peaks else "absent"},
        }
        if not args.no_cpu and world == 1:
            threads = po.host_threads()
            if po.ref_available():
                ref = po.Ref(ROBOT)
                renv = po.add_scene(po.RefEnv(), scenes.packed(scene))
                ref.time_configs(renv, host_batches[0][: 1 << 16], threads, reps=1)
                t = ref.time_configs(renv, host_batches[0], threads, reps=3)
                cpu_baseline = {"value": N_CONFIGS / t, "unit": "configs/s", "cores": threads, "kind": "reference",
                                "sample": f"one full 2^20-config batch (same inputs as the GPU step), best of 3, {threads} threads; "
                                          "reference AVX2 code compiled in place (oracle/_ref)"}
                if edges is not None:
                    benv = po.add_scene(po.RefEnv(), scenes.packed(scenes.box_scene()))
                    te = ref.time_edges(benv, a_h, b_h, threads, reps=2)
                    edges["cpu_baseline"] = {"value": N_EDGES / te, "unit": "edges/s", "cores": threads, "kind": "reference",
                                             "sample": f"the same 2^18 edges, best of 2, {threads} threads"}
            else:
                sub = host_batches[0][: 1 << 18]
                t0 = time.perf_counter()
                oracle.validate_configs(oenv, sub)
                t = time.perf_counter() - t0
                cpu_baseline = {"value": len(sub) / t, "unit": "configs/s", "cores": 1, "kind": "port",
                                "sample": "2^18 configs of batch 0 through the scalar C oracle, one thread"}
    except Exception as ex:  # the checker is optional for the number itself
        roofline = roofline or {"error": str(ex)}

    line = {
        "metric": "panda_collision_checked_configs_per_sec",
        "value": value,
        "unit": "configs/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": "C2: 2^20 random Panda configs vs synthetic MBM-style table/shelf scene (13 cuboids + 4 cylinders)",
            "robot": ROBOT, "n_configs_per_gpu": N_CONFIGS, "valid_fraction": valid_fraction,
            "l2": f"steps rotate through {N_BATCHES} distinct resident batches ({N_BATCHES * N_CONFIGS * dof * 4 / 1e6:.0f} MB > 126 MB L2)",
            "parallelism": f"dp{world}: independent shards per GPU" + ("" if world == 1 or args.no_gather else " + NCCL all_gather of verdict bitmasks"),
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "configs/s", "h2d_bytes_per_step": N_CONFIGS * dof * 4, "d2h_bytes_per_step": n_words * 4,
                "ms_per_step": e2e_s * 1e3, "api": "vmv_validate_configs (host pointers, pinned)"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "edges": edges,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
