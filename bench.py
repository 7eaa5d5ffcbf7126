#!/usr/bin/env python
"""bench.py -- headline benchmark of the LinearSFM hot path on B200.

Metric (BASELINE.json): end-to-end solve seconds of the hierarchical map-joining tree, stereo,
NC3500 shape (3499 local maps, ~128 new landmarks per frame; the real NC3500_C dataset is not
shipped with the reference, so the workload is the seeded synthetic scene of SURVEY 8(d)).
A "step" = one full merge-tree solve (all levels, final re-base), the region the reference times
(LinearSFMImp.cpp:1929 -> 2068).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--maps M]

Prints ONE JSON line (rank 0).  `value` = seconds per solve with the leaf maps already resident in
HBM; `e2e` = the same solve through the host-buffer C-ABI path (pack + H2D of all leaf maps, solve,
D2H of the final state inside the timed region); `roofline` = the dominant kernel's achieved HBM
GB/s (algorithmic bytes / CUDA-event time) against MEASURED_PEAKS.json; `cpu_baseline` = the
reference's own CPU code (oracle/_ref) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: file descriptor 1 is pointed at stderr for everything else
# (NCCL's version banner, the reference's printf progress lines, library chatter); the JSON line
# goes to a private duplicate of the original stdout
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

FULL_MAPS = 3499
FEATS = 128
METRIC = "end-to-end solve time, NC3500-shape stereo merge tree (3499 local maps)"
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# FP64 peak of this pool's B200s, measured with tools/ubench/dmma.cu (profiles/r2_dmma_ubench.txt):
# mma.sync.m8n8k4.f64 (DMMA) 37.07 TFLOP/s, plain DFMA 33.6 TFLOP/s -- MEASURED_PEAKS.json has no FP64 entry
FP64_PEAK_TFLOPS = 37.07


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


REF_BUDGET_S = 150.0     # wall-clock budget of the reference arm's timed solves


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation (oracle/_ref = unmodified
    LinearSFMImp.cpp + CHOLMOD shim), single-threaded by construction, on the FULL workload (every
    local map of the config; nothing is extrapolated).  The CPU code is deterministic and has no
    caches to warm, so there is no warm-up; solves are timed one by one until `--steps` are done or
    the next one would exceed REF_BUDGET_S, and the line reports how many were timed in `steps`
    (`requested` keeps the command line)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_oracle as ro
    from linearsfm_b200 import synth
    ro.build()
    nmaps = args.maps
    maps = make_scene(args)
    times, cpu_times = [], []
    t_begin = time.perf_counter()
    while len(times) < max(1, args.steps):
        if times and (time.perf_counter() - t_begin) + max(times) > REF_BUDGET_S:
            break
        _, t_ref, t_wall = ro.run_tree_stereo(maps)
        times.append(t_wall)
        cpu_times.append(t_ref)
    t = float(np.mean(times))
    line = {
        "impl": "reference", "metric": METRIC, "value": t, "unit": "s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": 0, "requested": {"steps": args.steps, "warmup": args.warmup},
        "ms_per_step": t * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "extrapolated": False,
        "config": {"workload": workload_name(nmaps, args.feats, args.scene), "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": t, "unit": "s", "cores": 1, "kind": "reference", "host_cores": os.cpu_count(),
                         "sample": f"all {nmaps} local maps, {len(times)} solve(s) timed, no warm-up "
                                   f"(wall {min(times):.2f}..{max(times):.2f} s; the reference's own clock() figure "
                                   f"{float(np.mean(cpu_times)):.2f} s)"},
        "e2e": {"value": t, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def workload_name(nmaps, feats=FEATS, scene="default"):
    extra = "" if scene == "default" else ", loop closures every 500 frames (revisit 0.1), gated landmarks <= 15 m"
    return f"synthetic NC3500-shape stereo scene, {nmaps} local maps, {feats} new landmarks/frame{extra}"


def make_scene(args):
    from linearsfm_b200 import synth
    if args.scene == "closed":
        return synth.make_stereo_scene(args.maps, feats_per_frame=args.feats, revisit=0.1, lap=500, max_depth=15.0,
                                       gate=True)
    return synth.make_stereo_scene(args.maps, feats_per_frame=args.feats)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--maps", type=int, default=FULL_MAPS)
    ap.add_argument("--feats", type=int, default=FEATS, help="new landmarks per frame")
    ap.add_argument("--scene", default="default", choices=["default", "closed"],
                    help="default: the round-1 open sequential chain; closed: loop closures every 500 frames, "
                         "outlier-gated landmarks (well conditioned; the 50k-map config uses it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    from linearsfm_b200 import api, synth, dist as lsd
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api.init(local_rank)
    dev = torch.device("cuda", local_rank)

    nmaps = args.maps
    maps_all = make_scene(args)
    lo, hi = lsd.slice_of(nmaps, world, rank)
    mine = maps_all[lo:hi]
    arr, keep = api.to_c_array(mine)
    h2d_bytes = sum(12 * m.r + 296 * m.nU + 152 * m.nW + 76 * m.n for m in mine)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    be = lsd.TreeBackend(api, mine)

    def step_resident():
        """inputs already resident in HBM: re-adopt the uploaded leaves, run the (sharded) tree."""
        be.tree.reset()
        if world == 1:
            be.tree.solve()
        else:
            lsd.run_sharded(be, nmaps, rank, world, dev)

    def step_e2e():
        be.tree.set_maps_c(arr, len(mine))
        if world == 1:
            be.tree.solve()
            if rank == 0:
                be.tree.download_state(0)
        else:
            root = lsd.run_sharded(be, nmaps, rank, world, dev)
            if root:
                be.tree.download_state(0)

    for _ in range(args.warmup):
        step_resident()
    api.stats_reset(stage_timing=False)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        step_resident()
        dev_ms += be.tree.last_solve_ms()
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None
    launches = api.stats()["launches"]
    sec = (t1 - t0) / args.steps
    tt = torch.tensor([sec, float(launches)], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        sec = float(mx[0]); launches = float(sm[1])

    # ---- e2e through host buffers ----
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_sec = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_sec, float(h2d_bytes)], dtype=torch.float64, device=dev)
    if dist is not None:
        mx = te.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = te.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_sec = float(mx[0]); h2d_bytes = float(sm[1])
    d2h_bytes = 0
    if rank == 0:
        s = be.tree.result_shape(0)
        d2h_bytes = 12 * s.r

    # ---- per-stage pass (CUDA events around every stage on the library's stream) ----
    roof = None
    stages = {}
    if world == 1:
        step_resident()                      # re-warm the stream-ordered pool after the e2e uploads
        api.stats_reset(stage_timing=True)
        step_resident()
        st = api.stats()["stages"]
        stages = {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                      "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                      "GFps": round(v["flops"] / max(v["ms"], 1e-9) / 1e6, 2)} for k, v in st.items()}
        peak, how = load_peaks()
        # stages that are ONE kernel launched once per tree level (CUDA events on the library's
        # stream around that launch): the roofline object is the one with the most time
        single = {"transform.wv": "tfc::k_tf_chunk", "solve.schur": "schur_lock::k_schur_lock",
                  "join.values": "k_join_w", "solve.backsub": "k_backsub"}
        cand = {k: v for k, v in st.items() if k in single and v["launches"] > 0}
        top = max(cand.items(), key=lambda kv: kv[1]["ms"])
        ach = top[1]["bytes"] / max(top[1]["ms"], 1e-9) / 1e6
        traffic = None
        traffic_note = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):     # dram bytes per launch from the committed ncu --set full capture
            try:
                tj = json.load(open(tpath)).get(single[top[0]], {})
                traffic = tj.get("dram_bytes_per_launch")
                traffic_note = {k: tj[k] for k in ("launch", "time_ms", "algorithmic_bytes_that_launch") if k in tj}
            except Exception:
                traffic = None
        # share of the GPU time of a step: the host-only stage (symbolic analysis, overlapped with
        # the Schur kernel when the stage timers are off) is left out, like in an ncu launch list
        total_ms = max(sum(v["ms"] for k, v in st.items() if k != "solve.symbolic"), 1e-9)
        fp64 = None
        if top[1].get("flops", 0) > 0:
            tf = top[1]["flops"] / max(top[1]["ms"], 1e-9) / 1e9
            fp64 = {"achieved_tflops": round(tf, 3), "peak_tflops": FP64_PEAK_TFLOPS, "frac": round(tf / FP64_PEAK_TFLOPS, 4),
                    "algorithmic_flops_per_launch": round(top[1]["flops"] / top[1]["launches"]),
                    "peak_source": "measured: tools/ubench/dmma.cu, profiles/r2_dmma_ubench.txt (DMMA m8n8k4)"}
        roof = {"kernel": single[top[0]], "stage": top[0], "bound": "hbm", "achieved": round(ach, 1), "peak": peak,
                "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic, "traffic_launch": traffic_note,
                "peak_source": how,
                "launches": top[1]["launches"], "ms_total": round(top[1]["ms"], 3),
                "avg_launch_ms": round(top[1]["ms"] / top[1]["launches"], 4),
                "algorithmic_bytes_per_launch": round(top[1]["bytes"] / top[1]["launches"]),
                "share_of_step": round(top[1]["ms"] / total_ms, 3),
                "fp64": fp64,
                "other_kernels": {single[k]: {"GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                                              "frac": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peak, 4),
                                              "ms_total": round(v["ms"], 3), "share_of_step": round(v["ms"] / total_ms, 3)}
                                  for k, v in cand.items() if k != top[0]}}
        # A kernel whose algorithmic intensity (flop / byte) lies above the machine balance (measured FP64
        # tensor-pipe peak / measured HBM peak) is bounded by the FP64 pipe, not by HBM: report it against that
        # roof ("tensor": the FP64 tensor pipe, DMMA m8n8k4 -- the DFMA pipe peaks 10 % lower) and keep the HBM view.
        if fp64 is not None and top[1]["flops"] / max(top[1]["bytes"], 1.0) > FP64_PEAK_TFLOPS * 1e3 / peak:
            roof["hbm"] = {"achieved": roof["achieved"], "peak": peak, "unit": "GB/s", "frac": roof["frac"],
                           "peak_source": how}
            roof.update({"bound": "tensor", "achieved": fp64["achieved_tflops"], "peak": FP64_PEAK_TFLOPS,
                         "unit": "TFLOP/s", "frac": fp64["frac"], "peak_source": fp64["peak_source"],
                         "intensity_flop_per_byte": round(top[1]["flops"] / max(top[1]["bytes"], 1.0), 2),
                         "machine_balance_flop_per_byte": round(FP64_PEAK_TFLOPS * 1e3 / peak, 2)})
        api.stats_reset(stage_timing=False)

    # ---- N > 1: the sharded result against a single-GPU solve of the same leaves (rank 0, untimed) ----
    verify = None
    if world > 1:
        be.tree.reset()
        root = lsd.run_sharded(be, nmaps, rank, world, dev)
        barrier()
        if rank == 0:
            assert root
            stno_s, st_s = be.tree.download_state(0)
            single = api.Tree(maps_all)
            single.solve()
            stno_1, st_1 = single.download_state(0)
            single.close()
            ints_equal = bool(np.array_equal(stno_s, stno_1))
            scale = float(np.max(np.abs(st_1)))
            diff = float(np.max(np.abs(st_s - st_1)) / scale) if ints_equal else None
            verify = {"against": "single-GPU solve of the same leaves on rank 0 (untimed)", "ints_equal": ints_equal,
                      "state_max_rel_diff": diff, "bit_identical": bool(ints_equal and np.array_equal(st_s, st_1)),
                      "note": "sharding changes which joins share a launch, hence the order of some fixed-order sums; "
                              "on the default open 3499-frame chain last-bit differences are amplified to ~1e-3 "
                              "(the reference's own result moves by 1.6e-3 under 1e-15 input noise there, "
                              "DESIGN.md section 3); --scene closed is the well-conditioned scene"}
        barrier()

    # ---- the monocular configs of BASELINE.json (RS90_C / RS468_C shapes), host buffers in and out ----
    mono = None
    if world == 1 and args.scene == "default":
        mono = {}
        for tag, n in (("rs90_shape_88_maps", 88), ("rs468_shape_466_maps", 466)):
            mm = synth.make_mono_scene(n, feats_per_frame=40, style="aerial", seed=n + 2)   # seeds 90 / 468 as in tests/test_gpu_mono.py
            api.run_mono(mm)                                   # warm-up
            ts = []
            for _ in range(3):
                torch.cuda.synchronize()
                t0m = time.perf_counter()
                api.run_mono(mm)
                ts.append(time.perf_counter() - t0m)
            mono[tag] = {"value": min(ts), "unit": "s",
                         "note": "lsfm_run_mono through host buffers (H2D + merge tree + D2H), best of 3"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import ref_oracle as ro
            _, t_ref, t_wall = ro.run_tree_stereo(maps_all)
            cpu = {"value": t_wall, "unit": "s", "cores": 1, "kind": "reference",
                   "host_cores": os.cpu_count(),
                   "sample": f"all {nmaps} local maps, one solve through oracle/_ref (unmodified LinearSFMImp.cpp + CHOLMOD shim), nothing extrapolated"}
        except Exception as e:  # the oracle is a checker; its absence must not kill the bench
            cpu = {"value": None, "unit": "s", "cores": 1, "kind": "reference", "sample": f"unavailable: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": sec, "unit": "s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(nmaps, args.feats, args.scene),
                       "l2": "inputs larger than L2 (leaf maps ~0.4 GB, upper levels > 1 GB)",
                       "parallelism": f"tree-level sharding x{world}" if world > 1 else "single GPU"},
            "device_ms_per_step": (dev_ms / args.steps) if world == 1 else None,
            "verify": verify,
            "e2e": {"value": e2e_sec, "unit": "s", "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes)},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if roof:
            line["roofline"] = roof
            line["stages"] = stages
        if cpu:
            line["cpu_baseline"] = cpu
        if mono:
            line["mono_configs"] = mono
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
