#!/usr/bin/env python
"""bench.py -- body-steps/sec of the Ballbox PhysicsStep on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line on rank 0.  A "step" is one PhysicsStep over the whole
scene, device resident.  Workloads (BASELINE.json configs):

  c3  mixed pile: 500k spheres + 500k rotated boxes in six static walls (the
      configuration the headline metric "body-steps/sec at 1M bodies" is quoted
      on).  A single world does not shard (sequential-order Gauss-Seidel), so at
      N > 1 every rank steps its own replica ("replicas only", weak scaling).
  c5  ONE batch of 4096 independent 256-body worlds, batched RL style, partitioned
      over the N GPUs (world w -> rank w // (4096/N), seeds by global world id:
      ballbox_b200/sharding.py): the workload that shards.  No collective on the
      step path, one NCCL all-reduce of the statistics after the last step.
      Reported under "batched_worlds" at every N: "strong" (4096 worlds in total,
      4096/N per GPU, with an order-independent 64-bit checksum of all worlds'
      state hashes that must be identical at every N) and, at N > 1, "weak"
      (4096 worlds per GPU).

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
SOLVE_KERNEL = "k_replay_staged"


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


def cpu_sample_c5(settle=30, steps=30):
    """Config 5 on ALL host cores: one reference world per thread (the reference is single-threaded and
    has no global state, so independent worlds run side by side; ctypes releases the GIL during a call)."""
    from concurrent.futures import ThreadPoolExecutor
    from ballbox_b200 import binding as bb, sharding
    lib, kind = checker_lib()
    threads = os.cpu_count() or 1
    worlds = [lib.build_scene(bb.SCENE_RLWORLD, 0, sharding.world_seed(i)) for i in range(threads)]

    def run(w, n):
        for _ in range(n):
            w.step(DT)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda w: run(w, settle), worlds))
        t0 = time.perf_counter()
        list(ex.map(lambda w: run(w, steps), worlds))
        el = time.perf_counter() - t0
    bodies = sum(w.n_spheres + w.n_boxes for w in worlds)
    for w in worlds:
        w.destroy()
    return {"value": bodies * steps / el, "unit": "body-steps/s", "cores": threads, "kind": kind,
            "sample": f"config 5: {threads} worlds of 256 bodies (seeds of worlds 0..{threads - 1}), one world per thread on "
                      f"{threads} threads, {settle} untimed + {steps} timed full all-pairs PhysicsStep calls each; "
                      f"{el / steps * 1e3:.2f} ms per step of all {threads} worlds"}


STATE_S = ("pos", "vel", "angvel", "sleeping", "timer", "force", "torque")
STATE_B = ("pos", "rot", "vel", "angvel", "sleeping", "timer", "force", "torque")


def cpu_pruned_same_size(res, steps=1):
    """The reference's own stage functions at the SAME size and state as the GPU arm (C3, 1M bodies, settled):
    detection through the CPU grid of oracle/ref_shim.c (RefPhysicsStepPruned, identical results to the
    all-pairs loops, tests/test_oracle.py) because the stock O(N^2) loops need ~9 h per step at this size."""
    from ballbox_b200 import binding as bb
    lib, kind = checker_lib()
    g = res["world"]
    g.download(bb.DL_BODIES)
    c = lib.build_scene(bb.SCENE_MIXED, res["n_dynamic"], res["seed"])
    sa, da = g.sphere_arrays(), c.sphere_arrays()
    for k in STATE_S:
        da[k][...] = sa[k]
    sa, da = g.box_arrays(), c.box_arrays()
    for k in STATE_B:
        da[k][...] = sa[k]
    t0 = time.perf_counter()
    for _ in range(steps):
        c.step(DT, pruned=True)
    el = time.perf_counter() - t0
    bodies = c.n_spheres + c.n_boxes
    contacts = len(c.constraints())
    c.destroy()
    return {"value": bodies * steps / el, "unit": "body-steps/s", "cores": 1, "kind": kind + "+cpu_grid",
            "ms_per_step": el / steps * 1e3, "contacts": contacts,
            "sample": f"same size and state as the GPU arm: {bodies} bodies, state copied from the settled GPU world, {steps} "
                      f"step(s) of the reference's stage functions with the all-pairs loops replaced by a CPU grid "
                      f"(oracle/ref_shim.c), 1 thread, no warm-up step"}


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
    torch.cuda.profiler.start()                  # cudaProfilerStart: `ncu --profile-from-start off python bench.py ...` lists exactly the timed launches
    e0.record(stream)
    world.step_n(DT, steps)
    e1.record(stream)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    barrier()
    ms = e0.elapsed_time(e1)
    stages, covered = world.stage_timing()
    replay_ms, replay_steps = world.replay_timing()
    stages["replay_kernel"] = replay_ms * covered / max(replay_steps, 1)      # part of solve_kernel (same normalisation)
    world.set_stage_timing(False)
    s1 = world.stats()
    return ms, stages, covered, s0, s1


def run_c3(args, lib, torch, barrier, rank):
    from ballbox_b200 import binding as bb
    n = args.bodies
    t_build = time.perf_counter()
    seed = 20261019 + rank
    w = lib.build_scene(bb.SCENE_MIXED, n, seed)
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
           "overflow": s1.overflow, "build_s": t_build, "settle_s": t_settle, "world": w, "seed": seed, "n_dynamic": n}
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


def run_batched(args, lib, torch, barrier, ids):
    """One rank's shard (global world ids `ids`) of a batch of independent 256-body worlds."""
    from ballbox_b200 import binding as bb, sharding
    stream = torch.cuda.current_stream()
    shard = sharding.WorldShard(lib, ids, bb.SCENE_RLWORLD, device=torch.cuda.current_device(), stream=stream.cuda_stream)
    done = 0
    while done < args.settle_worlds:
        k = min(50, args.settle_worlds - done)
        shard.step(DT, k); done += k
    hash_settled = sharding.combine_hashes(shard.world_hashes().values())
    shard.step(DT, args.warmup)
    st0 = shard.stats()
    barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    shard.step(DT, args.steps, sync=False)
    e1.record(stream)
    torch.cuda.synchronize(); barrier(); shard.synchronize()
    ms = e0.elapsed_time(e1)
    st = shard.stats()
    hash_final = sharding.combine_hashes(shard.world_hashes().values())
    out = {"ms": ms, "bodies": st.spheres + st.boxes, "contacts": st.contacts, "levels": st.levels, "worlds": len(ids),
           "launches": int(st.kernel_launches - st0.kernel_launches), "overflow": st.overflow,
           "hash_settled": hash_settled, "hash_final": hash_final}
    shard.destroy()
    return out


def pcie_probe(torch, barrier, mbytes=256, reps=4):
    """Pinned-memory copy bandwidth of THIS rank while every rank copies at the same time (GB/s, CUDA events):
    the host-side limit the drop-in call's 621 MB/step constraint download runs into at N > 1."""
    n = mbytes << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    out = {}
    for name, (dst, src) in (("d2h_gbs", (host, dev)), ("h2d_gbs", (dev, host))):
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record(); torch.cuda.synchronize(); barrier()
        out[name] = n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    return out


def load_traffic(args, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture
    (profiles/solver_traffic.json), only when it was taken on exactly this scene and kernel; else null."""
    p = os.path.join(ROOT, "profiles", "solver_traffic.json")
    try:
        t = json.load(open(p))
    except Exception:
        return None, None
    if t.get("bodies") != args.bodies or t.get("settle") != args.settle or t.get("kernel") != kernel:
        return None, None
    return float(t["dram_read_bytes"]) + float(t["dram_write_bytes"]), t.get("source")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bodies", type=int, default=1_000_000, help="dynamic bodies of the C3 scene")
    ap.add_argument("--worlds", type=int, default=4096, help="C5: worlds of the ONE batch that is partitioned over the GPUs "
                    "(strong scaling); the weak figure at N > 1 uses this many worlds per GPU")
    ap.add_argument("--settle", type=int, default=1500, help="untimed settle steps before measuring (a 160-unit tall pile needs ~1500)")
    ap.add_argument("--settle-worlds", type=int, default=300, help="untimed settle steps of the batched worlds")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--ref-bodies", type=int, default=4000, help="size of the bounded CPU sample")
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--skip-batched", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-pruned", action="store_true", help="skip the same-size CPU leg (reference stage functions + CPU grid at 1M bodies, ~30 s)")
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
    # The solve is two launches: k_solve_levels (iteration 0: level discovery fused with the first sweep) and
    # k_replay_staged (iterations 1..I-1), the dominant kernel of the step.  Both are timed live with CUDA events on the
    # step stream; algorithmic bytes per SURVEY.md 8d: one iteration = C*116 + D*48.
    solve_ms = r["stages_ms_per_step"]["solve_kernel"]
    replay_ms = r["stages_ms_per_step"]["replay_kernel"]
    sb = solver_bytes(C_, D_, I)
    rb = solver_bytes(C_, D_, max(I - 1, 0))
    achieved = rb / (replay_ms * 1e-3) / 1e9 if replay_ms > 0 else 0.0
    solve_ach = sb / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else 0.0
    stepb = step_bytes(r["ns"], r["nb"], P_, r["contacts"], D_, I)
    step_ach = stepb / (r["ms"] / args.steps * 1e-3) / 1e9
    traffic, traffic_src = load_traffic(args, SOLVE_KERNEL)
    roofline = {"bound": "hbm", "kernel": SOLVE_KERNEL, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "note": "iterations 1..I-1 of the exact level schedule; bound by the 51 x 7 grid-synchronised levels (about half of "
                        "every level is barrier + first gather) and by instruction issue while sweeping (~950 warp instructions "
                        "per constraint, reference arithmetic without FMA contraction), not by HBM: profiles/r02_notes.md",
                "algorithmic_bytes_per_launch": rb, "ms_per_launch": replay_ms,
                "whole_solve": {"kernels": "k_solve_levels + k_replay_staged", "algorithmic_bytes": sb, "ms": solve_ms,
                                "achieved": solve_ach, "frac": solve_ach / peak},
                "whole_step": {"algorithmic_bytes": stepb, "achieved": step_ach, "frac": step_ach / peak}}

    barrier()
    e2e = run_e2e(r, args.e2e_steps)
    if dist is not None:
        # every rank makes the drop-in call at the same time: the whole-job figure is the sum; the per-rank times and the
        # pinned-copy bandwidth each rank sees while all ranks copy show what saturates (one host's PCIe / memory path)
        per_rank = [None] * world_size
        dist.all_gather_object(per_rank, {"ms_per_step": e2e["ms_per_step"], "bodies_only_ms_per_step": e2e["bodies_only"]["ms_per_step"]})
        probe = pcie_probe(torch, barrier)
        probes = [None] * world_size
        dist.all_gather_object(probes, probe)
        e2e = dict(e2e, value=sum_over_ranks(e2e["value"]), ms_per_step=max(p["ms_per_step"] for p in per_rank),
                   per_rank_ms_per_step=[round(p["ms_per_step"], 3) for p in per_rank],
                   concurrent_pinned_copy_gbs={"d2h_per_rank": [round(p["d2h_gbs"], 1) for p in probes],
                                               "h2d_per_rank": [round(p["h2d_gbs"], 1) for p in probes],
                                               "d2h_sum": round(sum(p["d2h_gbs"] for p in probes), 1)})
        e2e["bodies_only"] = dict(e2e["bodies_only"], value=sum_over_ranks(e2e["bodies_only"]["value"]))
    elif rank == 0:
        e2e["pinned_copy_gbs"] = {k: round(v, 1) for k, v in pcie_probe(torch, barrier).items()}
    same_size_cpu = None
    if rank == 0 and world_size == 1 and not args.skip_cpu and not args.skip_pruned:
        same_size_cpu = cpu_pruned_same_size(r)
    r["world"].destroy()

    batched = None
    if not args.skip_batched:
        from ballbox_b200 import sharding

        def leg(ids, scaling):
            b = run_batched(args, lib, torch, barrier, ids)
            if b["overflow"]:
                raise SystemExit(f"bench.py: device buffer overflow flags {b['overflow']} in the batched worlds")
            bms = max_over_ranks(b["ms"])
            # the ONE collective of the design: statistics (and the checksum of checksums) after the last step
            tot = sharding.reduce_stats(sharding.pack_stats(b["bodies"] * args.steps, b["contacts"], b["worlds"], b["hash_final"]),
                                        dist, device="cuda")
            tot0 = sharding.reduce_stats(sharding.pack_stats(0, 0, 0, b["hash_settled"]), dist, device="cuda")
            return {"scaling": scaling, "worlds_total": int(tot["worlds"]), "worlds_per_gpu": b["worlds"],
                    "value": tot["body_steps"] / (bms * 1e-3), "unit": "body-steps/s", "ms_per_step": bms / args.steps,
                    "bodies_total": int(tot["body_steps"] / args.steps), "contacts_total": int(tot["contacts"]),
                    "levels_rank0": b["levels"], "launches_rank0": b["launches"],
                    "checksum_after_settle": "%016x" % sharding.unpack_hash(tot0),
                    "checksum_final": "%016x" % sharding.unpack_hash(tot),
                    "checksum_steps": args.settle_worlds + args.warmup + args.steps}
        strong = leg(sharding.strong_world_ids(rank, world_size, args.worlds), "strong")
        batched = {"workload": f"c5_one_batch_of_{args.worlds}_worlds_x_256_bodies_partitioned_over_{world_size}_gpus",
                   "value": strong["value"], "unit": "body-steps/s", "ms_per_step": strong["ms_per_step"],
                   "settle_steps": args.settle_worlds, "strong": strong,
                   "checksum_note": "sum modulo 2^64 of the FNV-1a state hashes of all worlds (seeded by global world id): "
                                    "identical at every N for equal --steps/--warmup; checksum_after_settle does not depend on them"}
        if world_size > 1:
            batched["weak"] = leg(sharding.rank_world_ids(rank, world_size, args.worlds), "weak")
        else:
            batched["weak"] = dict(strong, scaling="weak")

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
        cpu["batched_worlds_all_cores"] = cpu_sample_c5()
        if same_size_cpu:
            cpu["reference_pruned_1M"] = same_size_cpu

    if rank == 0:
        line = {"metric": "body-steps/sec", "value": value, "unit": "body-steps/s", "n_gpus": world_size,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_C3, "bodies_per_gpu": r["bodies"], "replicas": world_size,
                           "parallelism": "replicas only (a single world does not shard)",
                           "settle_steps": args.settle, "solver": "exact level-scheduled Gauss-Seidel (k_solve_levels + k_replay_staged)",
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
