#!/usr/bin/env python
"""bench.py -- body-steps/sec of the Ballbox PhysicsStep on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line on rank 0.  A "step" is one PhysicsStep over the whole
scene, device resident.  Workloads (BASELINE.json configs):

  c3  mixed pile: 500k spheres + 500k rotated boxes in six static walls (the
      configuration the headline metric "body-steps/sec at 1M bodies" is quoted
      on).  A single world does not shard (sequential-order Gauss-Seidel), so at
      N > 1 every rank steps its own replica ("replicas only", weak scaling).
  c5  4096 independent 256-body worlds per GPU, batched RL style; the batch is
      the workload that shards (no collective on the step path, one NCCL
      all-reduce of the statistics after the last step).  Reported next to c3
      under "batched_worlds" at every N.

`--impl reference` times the reference's own CPU PhysicsStep (oracle/_ref, the
unmodified header; falls back to the C port under oracle/) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = 1.0 / 60.0
WORKLOAD_C3 = "c3_mixed_pile_500k_spheres_500k_boxes_6_walls"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm
def checker_lib():
    from tests import oracles
    libs = oracles.load_all()
    if "ref" in libs:
        return libs["ref"], "reference"
    return libs["oracle"], "port"


def cpu_sample(n_bodies=4000, steps=12, threads=1):
    """The reference's full O(N^2) PhysicsStep on a C3-shaped scene of n_bodies."""
    from ballbox_b200 import binding as bb
    lib, kind = checker_lib()
    w = lib.build_scene(bb.SCENE_MIXED, n_bodies, 20261019)
    bodies = w.n_spheres + w.n_boxes
    w.step(DT)                                   # touch every page once
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step(DT)
    dt = time.perf_counter() - t0
    return {"value": bodies * steps / dt, "unit": "body-steps/s", "cores": threads, "kind": kind,
            "sample": f"full all-pairs PhysicsStep, C3-shaped scene of {bodies} bodies (same density, friction 0.3, "
                      f"restitution 0.3, 8 iterations), {steps} steps, 1 thread (the reference is single-threaded); "
                      f"{dt / steps * 1e3:.1f} ms/step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    from ballbox_b200 import binding as bb
    lib, kind = checker_lib()
    n = args.ref_bodies
    w = lib.build_scene(bb.SCENE_MIXED, n, 20261019)
    bodies = w.n_spheres + w.n_boxes
    for _ in range(max(1, args.warmup)):
        w.step(DT)
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step(DT)
    el = time.perf_counter() - t0
    v = bodies * steps / el
    sample = (f"full all-pairs PhysicsStep on a C3-shaped scene of {bodies} bodies per step (the 1M-body scene costs "
              f"~9 h per step on one core, SURVEY.md section 6), 1 thread: the reference has no threading")
    line = {"impl": "reference", "metric": "body-steps/sec", "value": v, "unit": "body-steps/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": el / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_C3, "sample_bodies": bodies},
            "cpu_baseline": {"value": v, "unit": "body-steps/s", "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def step_bytes(ns, nb, P, C, D, I):
    """Algorithmic bytes of one step, SURVEY.md section 8d."""
    return (ns * 165 + nb * 205) + (ns * 100 + nb * 132) + P * 16 + C * 104 + I * (C * 116 + D * 48)


def solver_bytes(C, D, I):
    """Algorithmic bytes of the solve kernel (one launch = all I iterations of a step)."""
    return I * (C * 116 + D * 48)


def time_resident(world, steps, warmup, torch, barrier):
    """W untimed + K timed resident steps; CUDA events on the step stream."""
    stream = torch.cuda.current_stream()
    world.set_stream(stream.cuda_stream)
    world.step_n(DT, warmup)
    world.synchronize()
    world.set_stage_timing(True)
    world.stage_timing()                         # reset accumulation
    s0 = world.stats()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    world.step_n(DT, steps)
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    stages, covered = world.stage_timing()
    world.set_stage_timing(False)
    s1 = world.stats()
    return ms, stages, covered, s0, s1


def run_c3(args, lib, torch, barrier, rank):
    from ballbox_b200 import binding as bb
    n = args.bodies
    t_build = time.perf_counter()
    w = lib.build_scene(bb.SCENE_MIXED, n, 20261019 + rank)
    w.set_sync_mode(bb.SYNC_RESIDENT)
    lib.dll.BallboxSetDevice(w.ptr, torch.cuda.current_device())
    w.upload()
    t_build = time.perf_counter() - t_build
    t_settle = time.perf_counter()
    done = 0
    while done < args.settle:                   # settle on the GPU, untimed, in slices so a hang is bounded
        k = min(50, args.settle - done)
        w.step_n(DT, k); w.synchronize(); done += k
    t_settle = time.perf_counter() - t_settle
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    ms, stages, covered, s0, s1 = time_resident(w, args.steps, args.warmup, torch, barrier)
    clocks = sampler.stop()
    bodies = s1.spheres + s1.boxes
    iters = w.w.settings.solver_iterations
    res = {"ms": ms, "bodies": bodies, "stages_ms_per_step": {k: v / max(covered, 1) for k, v in stages.items()},
           "launches": int(s1.kernel_launches - s0.kernel_launches), "clocks": clocks,
           "contacts": s1.contacts, "solver_contacts": s1.solver_contacts, "touched": s1.touched_bodies,
           "levels": s1.levels, "pairs": s1.candidate_pairs, "iters": iters, "ns": s1.spheres, "nb": s1.boxes,
           "overflow": s1.overflow, "build_s": t_build, "settle_s": t_settle, "world": w}
    return res


def run_e2e(res, steps):
    """The drop-in call: PhysicsStep(world, dt) with HOST arrays; every step uploads the
    host SoA arrays, steps, and downloads bodies + the constraint list."""
    from ballbox_b200 import binding as bb
    w = res["world"]
    w.download(bb.DL_BODIES)
    w.set_sync_mode(bb.SYNC_COMPAT)
    w.step(DT)                                   # warm the path (allocates the export buffer)
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step(DT)
    el = time.perf_counter() - t0
    ns, nb, C = res["ns"], res["nb"], len(w.constraints())
    h2d = ns * (12 * 5 + 4 * 4 + 2) + nb * (12 * 7 + 16 + 4 * 2 + 2)
    d2h_bodies = ns * (12 * 3 + 4 + 1) + nb * (12 * 3 + 16 + 4 + 1)
    d2h = d2h_bodies + C * 104
    out = {"value": res["bodies"] * steps / el, "unit": "body-steps/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": steps, "ms_per_step": el / steps * 1e3,
           "mode": "BBX_SYNC_COMPAT (default drop-in: bodies up, bodies + world->constraints down, every step)"}
    # the same call when the user opts out of the per-step constraint-list download (ballbox_ext.h)
    w.set_sync_mode(bb.SYNC_COMPAT_BODIES)
    w.step(DT)
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step(DT)
    el2 = time.perf_counter() - t0
    out["bodies_only"] = {"value": res["bodies"] * steps / el2, "unit": "body-steps/s", "ms_per_step": el2 / steps * 1e3,
                          "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_bodies),
                          "mode": "BBX_SYNC_COMPAT_BODIES (constraint list stays on the device until asked for)"}
    return out


def run_small_config(lib, torch, kind, n, settle, steps, label):
    """Resident stepping of one of the other BASELINE configs (single world), CUDA events on the step stream."""
    from ballbox_b200 import binding as bb
    w = lib.build_scene(kind, n, 20261019)
    w.set_sync_mode(bb.SYNC_RESIDENT)
    lib.dll.BallboxSetDevice(w.ptr, torch.cuda.current_device())
    stream = torch.cuda.current_stream()
    w.set_stream(stream.cuda_stream)
    w.upload()
    done = 0
    while done < settle:
        k = min(50, settle - done); w.step_n(DT, k); w.synchronize(); done += k
    w.set_stage_timing(True); w.stage_timing()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); w.step_n(DT, steps); e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    stages, covered = w.stage_timing()
    st = w.stats()
    bodies = st.spheres + st.boxes
    out = {"workload": label, "value": bodies * steps / (ms * 1e-3), "unit": "body-steps/s", "ms_per_step": ms / steps,
           "bodies": bodies, "contacts": st.contacts, "levels": st.levels, "settle_steps": settle,
           "stages_ms_per_step": {k: round(v / max(covered, 1), 4) for k, v in stages.items()}}
    w.destroy()
    return out


def run_c5(args, lib, torch, barrier, rank):
    """4096 independent 256-body worlds per GPU."""
    import ctypes as C
    from ballbox_b200 import binding as bb
    n_worlds = args.worlds
    s, b, c = lib.scene_capacity(bb.SCENE_RLWORLD, 0)
    batch = lib.dll.CreatePhysicsWorldBatch(n_worlds, s, b, c)
    if not batch:
        raise RuntimeError("CreatePhysicsWorldBatch failed")
    for i in range(n_worlds):
        lib.dll.BbxSceneBuild(lib.dll.BatchWorld(batch, i), bb.SCENE_RLWORLD, 0, 7000 + rank * n_worlds + i)
    lib.dll.BatchSetDevice(batch, torch.cuda.current_device())
    stream = torch.cuda.current_stream()
    lib.dll.BatchSetStream(batch, C.c_void_p(stream.cuda_stream))

    def chk():
        e = lib.dll.BatchGetLastError(batch)
        if e:
            raise RuntimeError(e.decode())
    lib.dll.BatchUpload(batch); chk()
    done = 0
    while done < args.settle:
        k = min(50, args.settle - done)
        lib.dll.PhysicsStepBatch(batch, DT, k); lib.dll.BatchSynchronize(batch); chk(); done += k
    lib.dll.PhysicsStepBatch(batch, DT, args.warmup); lib.dll.BatchSynchronize(batch); chk()
    st0 = bb.BallboxStats(); lib.dll.BatchGetStats(batch, C.byref(st0))
    barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    lib.dll.PhysicsStepBatch(batch, DT, args.steps)
    e1.record(stream)
    torch.cuda.synchronize(); barrier(); chk()
    ms = e0.elapsed_time(e1)
    st = bb.BallboxStats(); lib.dll.BatchGetStats(batch, C.byref(st))
    out = {"ms": ms, "bodies": st.spheres + st.boxes, "contacts": st.contacts, "levels": st.levels, "worlds": n_worlds,
           "launches": int(st.kernel_launches - st0.kernel_launches), "overflow": st.overflow}
    lib.dll.DestroyPhysicsWorldBatch(batch)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bodies", type=int, default=1_000_000, help="dynamic bodies of the C3 scene")
    ap.add_argument("--worlds", type=int, default=4096, help="C5 worlds per GPU")
    ap.add_argument("--settle", type=int, default=1500, help="untimed settle steps before measuring (a 160-unit tall pile needs ~1500)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-bodies", type=int, default=4000, help="size of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--skip-batched", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-others", action="store_true", help="do not time configs 2 and 4 (N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import ballbox_b200
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        # stdout carries the ONE JSON line: NCCL's own banner (NCCL_DEBUG=VERSION on some boxes) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)       # the one collective of the design: statistics, after the last step
        return float(t.item())

    lib = ballbox_b200.load()
    peak, peak_src = load_peaks()

    r = run_c3(args, lib, torch, barrier, rank)
    if r["overflow"]:
        raise SystemExit(f"bench.py: device buffer overflow flags {r['overflow']}")
    ms = max_over_ranks(r["ms"])
    total_bodies = sum_over_ranks(r["bodies"])
    value = total_bodies * args.steps / (ms * 1e-3)
    I = r["iters"]
    C_, D_, P_ = r["solver_contacts"], r["touched"], r["pairs"]
    solve_ms = r["stages_ms_per_step"]["solve_kernel"]
    sb = solver_bytes(C_, D_, I)
    achieved = sb / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
    stepb = step_bytes(r["ns"], r["nb"], P_, r["contacts"], D_, I)
    step_ach = stepb / (r["ms"] / args.steps * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of k_solve_levels for this exact scene (C3, 1M bodies, settle
    # 1500) from the ncu capture summarised in profiles/r01i_launches_c3_1M.md; null for any other scene
    traffic = 5021.3e6 + 1595.0e6 if (args.bodies == 1_000_000 and args.settle == 1500) else None
    roofline = {"bound": "hbm", "kernel": "k_solve_levels", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "note": "bound by waiting (barriers 48 % of stall samples, issue slots 23 % active), not by HBM: profiles/r01i_ncu_solve_levels.md, profiles/r01_notes.md",
                "algorithmic_bytes_per_launch": sb, "ms_per_launch": solve_ms,
                "whole_step": {"algorithmic_bytes": stepb, "achieved": step_ach, "frac": step_ach / peak}}

    e2e = run_e2e(r, args.e2e_steps) if rank == 0 or dist is not None else None
    if dist is not None:
        e2e_total = sum_over_ranks(e2e["value"])
        e2e = dict(e2e, value=e2e_total)
    r["world"].destroy()

    batched = None
    if not args.skip_batched:
        b = run_c5(args, lib, torch, barrier, rank)
        bms = max_over_ranks(b["ms"])
        btot = sum_over_ranks(b["bodies"])
        batched = {"workload": f"c5_{args.worlds}_worlds_x_256_bodies_per_gpu", "value": btot * args.steps / (bms * 1e-3),
                   "unit": "body-steps/s", "ms_per_step": bms / args.steps, "worlds_per_gpu": args.worlds,
                   "bodies_total": int(btot), "contacts_rank0": b["contacts"], "levels_rank0": b["levels"],
                   "launches": b["launches"], "scaling": "weak"}

    others = None
    if rank == 0 and world_size == 1 and not args.skip_others:
        from ballbox_b200 import binding as bb
        others = [run_small_config(lib, torch, bb.SCENE_BALLS, 100_000, 300, args.steps,
                                   "c2_balls_in_a_box_100k_spheres_6_walls_restitution_0.3"),
                  run_small_config(lib, torch, bb.SCENE_SDF, 200_000, 400, args.steps,
                                   "c4_sdf_terrain_100k_spheres_100k_boxes_on_wavy_sdf_4_walls")]

    cpu = None
    if rank == 0 and world_size == 1 and not args.skip_cpu:
        cpu = cpu_sample(args.ref_bodies, args.cpu_steps)

    if rank == 0:
        line = {"metric": "body-steps/sec", "value": value, "unit": "body-steps/s", "n_gpus": world_size,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_C3, "bodies_per_gpu": r["bodies"], "replicas": world_size,
                           "parallelism": "replicas only (a single world does not shard)",
                           "settle_steps": args.settle, "solver": "exact level-scheduled Gauss-Seidel",
                           "solver_iterations": I, "l2_policy": "inputs larger than L2 (state + constraints > 126 MB)",
                           "contacts": r["contacts"], "candidate_pairs": P_, "touched_bodies": D_, "levels": r["levels"]},
                "roofline": roofline, "e2e": e2e, "gpu_launches": r["launches"], "clocks": r["clocks"],
                "stages_ms_per_step": r["stages_ms_per_step"], "batched_worlds": batched,
                "setup": {"scene_build_s": r["build_s"], "settle_s": r["settle_s"]}}
        if others:
            line["other_configs"] = others
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
