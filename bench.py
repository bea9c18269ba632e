#!/usr/bin/env python
"""bench.py - scenes/sec of the Uni3DETR per-scene forward hot path (BASELINE.json metric) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--dtype bf16|fp32]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N
  python bench.py --impl reference      # CPU arm: the oracle port on the host cores

One "step" = one pass of the hot path (voxelize -> sparse encoder -> dense CNN -> 2x FPS -> decoder
-> heads -> NMSFreeCoder top-k; --postprocess adds the device-side per-class NMS) over a batch of B synthetic 20k-point SUN-RGBD-shaped scenes
(BASELINE config 2: uni3detr_sunrgbd.py, 300 queries x 4 groups, 3 decoder layers, bf16).
Scenes are sharded whole across ranks (weak scaling: B scenes per rank per step, no data-path
collective); one all-reduce of the metrics vector ends the job.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# The driver parses ONE JSON line from stdout: keep the real stdout for that line only and send
# everything else that libraries print on fd 1 (NCCL's version banner, cuDNN notices) to stderr.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


WORKLOAD = "sunrgbd"
# BASELINE.json configs 2-5: workload -> (dtype, scenes per step per GPU, description). Config 2 is the metric's
# configuration (the headline line); 3-5 are reported as extra blocks (`configs`) at their stated dtypes.
WORKLOADS = {
    "sunrgbd": ("bf16", 32, "uni3detr_sunrgbd.py synthetic 20k-pt scenes, 300 queries x4 groups, 3 decoder layers"),
    "scannet_large": ("fp32", 4, "uni3detr_scannet_large.py synthetic 100k-pt scenes, dynamic voxelization, 300 queries x4 groups"),
    "kitti": ("bf16", 8, "uni3detr_kitti_3classes.py synthetic 20k-pt lidar scenes, grid 41x1600x1408, 9 decoder layers"),
    "nuscenes": ("fp32", 2, "uni3detr_nuscenes.py synthetic 200k-pt multi-sweep scenes, 900 queries x4 groups"),
}
RIDGE_NOTE = "bound = tensor if FLOP/byte of the launch exceeds measured bf16 peak / measured HBM peak, else hbm"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm ----
def cpu_port_rate(n_scenes, threads, seed=0, warm=0, workload=None):
    """Oracle (CPU port of the reference path, oracle/model.py) on `n_scenes` 20k-point scenes.
    Returns (scenes/sec, seconds). bench.py is one of the places allowed to execute oracle/."""
    import torch
    from oracle import model as M
    from uni3detr_b200 import synth
    torch.set_num_threads(threads)
    workload = workload or WORKLOAD
    model, cfg = synth.build_model(workload, seed=0)
    sd = model.state_dict()
    rp = torch.rand(1, cfg["pts_bbox_head"]["num_query"], 3, generator=torch.Generator().manual_seed(seed))
    scenes = [synth.make_scene(workload, 100 + i) for i in range(n_scenes + warm)]
    for s in scenes[:warm]:
        M.forward(sd, cfg, [s], random_point=rp)
    t0 = time.perf_counter()
    for s in scenes[warm:]:
        outs, _, _ = M.forward(sd, cfg, [s], random_point=rp)
        M.nms_free_decode(outs, cfg["pts_bbox_head"]["bbox_coder"])
    dt = time.perf_counter() - t0
    return n_scenes / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path. Its arithmetic lives in
    mmcv/mmdet3d/spconv, which cannot be installed here, so this arm times the oracle
    port with every host thread, one scene per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warm = max(1, args.steps), min(args.warmup, 1)   # one 20k-point scene per step (~1.1 s on 16 cores)
    rate, dt = cpu_port_rate(steps, cores, warm=warm, workload=args.workload)
    line = {"impl": "reference", "metric": "scenes/sec", "value": rate, "unit": "scenes/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "note": "warm-up clipped to 1 scene (each CPU step is ~1 s); --steps honoured",
            "config": {"workload": WORKLOADS[args.workload][2] + ", full forward", "scenes_per_step": 1},
            "cpu_baseline": {"value": rate, "unit": "scenes/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} scene(s), one per step, oracle/model.py full forward + decode"},
            "e2e": {"value": rate, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------ GPU arm ----
def spconv_cost(info, n_in):
    """Algorithmic bytes / FLOPs of one sparse-conv launch:
    (N_in*Cin + N_out*Cout)*sizeof(act) + K*Cin*Cout*sizeof(w) + pairs*8 B; 2*pairs*Cin*Cout FLOP.
    N_in / N_out are the LIVE row counts (not buffer capacities)."""
    es, pairs, n_out = info["esize"], info["pairs"], info["n_out"]
    by = (n_in * info["Cin"] + n_out * info["Cout"]) * es + info["K"] * info["Cin"] * info["Cout"] * es + pairs * 8
    return by, 2.0 * pairs * info["Cin"] * info["Cout"]


def dominant_kernel(kernels):
    """The record the roofline is reported for. Records are (kernel function, layer shape) groups - `spconv_tc[27x64->64]`,
    `linear_tc[256->256]`, `fps[...]` -; the dominant KERNEL is the function whose groups sum to the most time in the step
    (one sparse-conv kernel runs 25 launches over ten shapes), and within it the heaviest shape is the one reported.
    Returns (record, summed ms of the function)."""
    fam = {}
    for k in kernels:
        f = k["kernel"].split("[")[0]
        fam[f] = fam.get(f, 0.0) + k["ms_per_step"]
    best = max(fam, key=fam.get)
    rec = max((k for k in kernels if k["kernel"].split("[")[0] == best), key=lambda k: k["ms_per_step"])
    return rec, fam[best]


def roofline_of(top, peaks):
    """Roofline of one kernel record of the profile pass (pure arithmetic; tests/test_abi.py exercises it)."""
    # every op of this pass is timed ALONE (a sync before each launch): the GPU is not under the sustained
    # power-capped load of a long step, so the tensor denominator is the BURST cuBLAS figure of
    # MEASURED_PEAKS.json (B200_PROFILING.md); the sustained figure is reported beside it
    tf_peak = peaks["tf_burst"]
    ridge = tf_peak * 1e12 / (peaks["hbm"] * 1e9)
    intensity = top["flops_per_launch"] / max(top["bytes_per_launch"], 1.0)
    if intensity > ridge:
        roof = {"bound": "tensor", "achieved": top["tflops"], "peak": tf_peak, "unit": "TFLOP/s",
                "frac": top["tflops"] / tf_peak, "frac_of_sustained_peak": top["tflops"] / peaks["tf_sustained"],
                "sustained_peak": peaks["tf_sustained"]}
    else:
        roof = {"bound": "hbm", "achieved": top["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                "frac": top["gbs"] / peaks["hbm"]}
    traffic = None   # dram__bytes_read+write per launch from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f).get(top["kernel"], {})
        traffic = tj.get("dram_bytes_per_launch")
        if tj.get("commit"):
            roof["traffic_capture_commit"] = tj["commit"]      # binary the ncu capture was taken from
        if tj.get("tensor_pipe_active_pct") is not None:   # same committed ncu capture
            roof["tensor_pipe_active_pct_ncu"] = tj["tensor_pipe_active_pct"]
    roof.update({"traffic": traffic, "kernel": top["kernel"], "avg_launch_ms": top["avg_launch_ms"],
                 "algorithmic_bytes_per_launch": top["bytes_per_launch"],
                 "algorithmic_flops_per_launch": top["flops_per_launch"],
                 "peak_source": peaks["source"] + " (burst cuBLAS bf16 for a kernel timed alone / copy bandwidth)", "rule": RIDGE_NOTE})
    return roof


def roofline_pass(step_fn, peaks, reps=3):
    """Per-op device times (CUDA events on the launching stream) of the libu3d kernels inside the
    step, then the roofline of the dominant one."""
    import torch
    from uni3detr_b200 import ops
    step_fn()
    torch.cuda.synchronize()
    agg = {}
    for rep in range(reps):
        ops.profile_begin()
        step_fn()
        live_rows = 0   # live rows of the current sparse level = input rows of the next conv
        seen = {}
        for name, info, ms in ops.profile_end():
            if name in ("spconv_fwd", "spconv_fwd_packed", "spconv_fwd_packed_x3"):
                kind = {"spconv_fwd": "spconv_simt", "spconv_fwd_packed": "spconv_tc", "spconv_fwd_packed_x3": "spconv_tc_x3"}[name]
                key = f"{kind}[{info['K']}x{info['Cin']}->{info['Cout']}]"   # x3: fp32 layer as 3 bf16 tensor-core passes
                name = "spconv_fwd"
            elif name == "fps":
                key = f"fps[n<={info['max_n']},nq={info['nq']}]"
            elif name == "linear_tc":
                key = f"linear_tc[{info['K']}->{info['N']}]"
            else:
                key = name
            a = agg.setdefault(key, dict(ms=[], calls=0, bytes=0.0, flops=0.0, name=name))
            idx = seen.get(key, 0)
            seen[key] = idx + 1
            # the profile pass runs `reps` times; a call's time is its MINIMUM over the repetitions (the
            # first repetition can include allocator growth on the host side of an op)
            if idx < len(a["ms"]):
                a["ms"][idx] = min(a["ms"][idx], ms)
                continue_cost = False
            else:
                a["ms"].append(ms)
                continue_cost = True
            if name == "spconv_fwd":
                by, fl = spconv_cost(info, live_rows)
                live_rows = info["n_out"]
            elif name in ("voxelize_hard", "voxelize_dynamic"):
                live_rows = info["n_voxels"]
                by, fl = info["n_points"] * info["C"] * 4 + live_rows * (info["C"] * 4 + 16), 0.0
            elif name in ("rulebook_subm", "rulebook_down"):
                by, fl = info["n_in"] * 16 + info["pairs"] * 8 + 2 * info["n_in"] * 8, 0.0
            elif name in ("sparse_to_dense", "sine_embed"):
                by, fl = info["bytes"], 0.0
            elif name == "fps":
                by, fl = info["B"] * (info["max_n"] * 12 + info["nq"] * 16), 0.0
            elif name == "cross_sample":
                by, fl = info["rows"] * info["C"] * info["esize"] * (8 + 2), 0.0
            elif name == "linear_tc":
                by = (info["rows"] * (info["K"] + info["N"]) + info["N"] * info["K"]) * 2
                fl = 2.0 * info["rows"] * info["K"] * info["N"]
            elif name == "mha_core":
                r = info["n_seq"] * info["seq_len"]
                by = r * info["heads"] * 32 * info["esize"] * 4
                fl = 4.0 * info["n_seq"] * info["heads"] * info["seq_len"] ** 2 * 32
            else:
                by, fl = float(info.get("bytes", 0.0)), 0.0
            if continue_cost:        # algorithmic cost counted once per call (identical in every repetition)
                a["calls"] += 1
                a["bytes"] += by
                a["flops"] += fl
    for a in agg.values():
        a["ms"] = sum(a["ms"])
    kernels = []
    for key, a in agg.items():
        per_ms = a["ms"] / a["calls"]
        kernels.append({"kernel": key, "calls_per_step": a["calls"], "ms_per_step": a["ms"],
                        "avg_launch_ms": per_ms, "gbs": a["bytes"] / a["ms"] / 1e6 if a["ms"] else 0.0,
                        "tflops": a["flops"] / a["ms"] / 1e9 if a["ms"] else 0.0,
                        "bytes_per_launch": a["bytes"] / a["calls"], "flops_per_launch": a["flops"] / a["calls"]})
    kernels.sort(key=lambda k: -k["ms_per_step"])
    top, fam_ms = dominant_kernel(kernels)
    roof = roofline_of(top, peaks)
    fam = top["kernel"].split("[")[0]
    roof["selection"] = ("dominant kernel = the libu3d kernel function with the largest summed time in the step (%s: %.2f ms "
                         "over all its layer shapes); the roofline is that of its heaviest shape. Every row of `kernels` "
                         "carries its own GB/s and TFLOP/s." % (fam, fam_ms))
    ours_ms = sum(k["ms_per_step"] for k in kernels)
    return roof, kernels[:12], ours_ms


def measure(workload, B, K, W, dtype_name, args, world, rank, dev, peaks, with_e2e=True, with_roofline=True):
    """One workload: device-resident scenes/s (CUDA events per step, L2 flush between steps, max over ranks),
    end-to-end scenes/s from pinned host points, roofline of the dominant libu3d kernel. Returns a dict."""
    import torch
    import torch.distributed as dist
    from uni3detr_b200 import ops, sharding, synth

    dtype = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    model, cfg = synth.build_model(workload, seed=0)
    model = model.to(dev)
    model.set_compute_dtype(dtype)
    nq = cfg["pts_bbox_head"]["num_query"]
    coder = model.pts_bbox_head.bbox_coder
    # rank r owns scenes {i : i mod world == r} of the job's B*world scene pool (weak scaling)
    mine = sharding.scene_indices(B * world, rank, world)
    host_pts = [torch.from_numpy(synth.make_scene(workload, i)).pin_memory() for i in mine]
    dev_pts = [p.to(dev) for p in host_pts]
    rp = torch.rand(B, nq, 3, generator=torch.Generator().manual_seed(1234 + rank)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    host_batch = torch.cat(host_pts, 0).pin_memory()                # (B*n_points, C) pinned

    def eager_step(points):
        outs, _ = model.forward_raw(points, random_point=rp)
        if args.postprocess:      # + bottom-centre shift, per-class NMS, thresholds (the step after the path)
            return model.pts_bbox_head.postprocess_fixed(outs)
        return coder.decode_fixed(outs)

    graphed = None
    if not args.no_graph:
        # the public serving call: the whole forward captured once as a CUDA graph, replayed per batch
        from uni3detr_b200 import GraphedForward
        graphed = GraphedForward(model, [p.shape[0] for p in host_pts], host_pts[0].shape[1], random_point=rp,
                                 postprocess=args.postprocess)
        graphed.load(host_batch)

    def step(points):
        if graphed is not None:
            return graphed.run(points if torch.is_tensor(points) else None)
        return eager_step(points)

    # ---- device-resident throughput ("value")
    for _ in range(W):
        step(dev_pts)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    n0 = ops.launch_count()
    torch.cuda.synchronize()
    for s, e in evs:
        flush.zero_()                      # untimed L2 flush between timed iterations
        s.record()
        res = step(dev_pts)
        e.record()
    torch.cuda.synchronize()
    launches = ops.launch_count() - n0
    if graphed is not None:
        launches = graphed.launches_per_replay * K      # kernels of the replayed graph
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_s = sum(s.elapsed_time(e) for s, e in evs) / 1e3
    checksum = float(res[1].float().sum())
    scenes, t_max, checksum = sharding.reduce_metrics(B * K, dev_s, checksum, device=dev)
    out = {"workload": workload, "value": scenes / t_max, "unit": "scenes/s", "ms_per_step": 1e3 * t_max / K,
           "dtype": dtype_name, "scenes_per_step_per_gpu": B, "points_per_scene": int(host_pts[0].shape[0]),
           "steps": K, "gpu_launches": launches, "clocks": clocks, "checksum": checksum,
           "launch": "eager" if graphed is None else "CUDA graph replay (uni3detr_b200.GraphedForward)"}

    # ---- end to end through the reference-facing call with HOST buffers ("e2e")
    if with_e2e:
        out_host = None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if graphed is not None and not args.no_pipeline:
            # the serving loop a user runs: every step copies its batch from pinned host memory and reads its
            # boxes back to the host; the copies of neighbouring steps overlap the replay (GraphedForward.submit)
            graphed.start_pipeline()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(K):
                graphed.submit(host_batch)                                   # H2D + replay + D2H, all enqueued
                # (collect() of step i-1 would go here in a server; its wait is hidden behind step i's replay)
            out_host = graphed.collect()                                     # results of the last step on the host
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
        else:
            t0 = time.perf_counter()
            for _ in range(K):
                if graphed is not None:
                    boxes, scores, labels, mask = step(host_batch)                # H2D of the batch inside
                else:
                    boxes, scores, labels, mask = step([p.to(dev, non_blocking=True) for p in host_pts])
                out_host = [t.cpu() for t in (boxes, scores, labels, mask)]      # D2H read of the step's result
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
        _, e2e_max, _ = sharding.reduce_metrics(B * K, e2e_s, 0.0, device=dev)
        out["e2e"] = {"value": scenes / e2e_max, "unit": "scenes/s",
                      "h2d_bytes_per_step": sum(p.numel() * p.element_size() for p in host_pts),
                      "d2h_bytes_per_step": sum(t.numel() * t.element_size() for t in out_host)}
    # ---- rank 0 only: roofline of the dominant libu3d kernel
    if rank == 0 and with_roofline and not args.no_roofline:
        out["roofline"], out["kernels"], out["libu3d_ms_per_step"] = roofline_pass(lambda: eager_step(dev_pts), peaks)
    del graphed, model, dev_pts, flush
    torch.cuda.empty_cache()
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback); use --impl reference")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)
    peaks = load_peaks()
    wl = args.workload
    dtype_name = args.dtype or WORKLOADS[wl][0]
    B = args.batch or WORKLOADS[wl][1]
    main = measure(wl, B, K, W, dtype_name, args, world, rank, dev, peaks)

    # extra blocks (every rank takes part: the metrics all-reduce is collective): the other BASELINE configs at
    # their stated sizes / dtypes, and the headline workload at the reference's batch sizes 1 and 4 (latency)
    extra, batches = [], {}
    if not args.no_extra:
        k2 = max(3, min(K, 5))
        for b in (1, 4):
            if wl == "sunrgbd" and b != B:
                try:
                    r = measure(wl, b, k2, 3, dtype_name, args, world, rank, dev, peaks, with_roofline=False)
                    batches[str(b)] = {"value": r["value"], "ms_per_step": r["ms_per_step"],
                                       "ms_per_scene": r["ms_per_step"] / b, "e2e": r["e2e"]["value"]}
                except Exception as ex:   # noqa: BLE001  (report, never lose the headline line)
                    batches[str(b)] = {"error": repr(ex)[:200]}
        for name, (dt, b, desc) in WORKLOADS.items():
            if name == wl:
                continue
            try:
                r = measure(name, b, k2, 3, dt, args, world, rank, dev, peaks)
                r["config"] = desc
                r["kernels"] = (r.get("kernels") or [])[:8]
                extra.append(r)
            except Exception as ex:       # noqa: BLE001
                extra.append({"workload": name, "error": repr(ex)[:300]})
                torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cores = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, dt = cpu_port_rate(args.cpu_scenes, cores, workload=wl)
        cpu = {"value": rate, "unit": "scenes/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_scenes} scene(s) of the same workload, oracle/model.py full forward + decode, {dt:.1f} s"}
    line = {"metric": "scenes/sec", "value": main["value"], "unit": "scenes/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype_name, "data": "synthetic",
            "config": {"workload": WORKLOADS[wl][2] + ", full forward (voxelize+sparse encoder+dense CNN+FPS+decoder+top-k)"
                                   + ("+per-class NMS" if args.postprocess else ""),
                       "scenes_per_step_per_gpu": B, "points_per_scene": main["points_per_scene"],
                       "parallelism": f"scenes sharded x{world}",
                       "l2": "256 MiB flush write between timed steps (untimed)", "timing": "CUDA events per step, max over ranks",
                       "launch": main["launch"]},
            "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "clocks": main["clocks"],
            "roofline": main.get("roofline"), "cpu_baseline": cpu, "kernels": main.get("kernels", []),
            "libu3d_ms_per_step": main.get("libu3d_ms_per_step"), "checksum": main["checksum"],
            "batches": batches, "configs": extra}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """--train: data-parallel TRAINING steps of the same model (SURVEY.md 8f rank 2 / BASELINE configs 4-5 "DDP"):
    forward_train under autograd (hand-written sparse-conv / cross-sample backward, device matcher + losses),
    ONE flat gradient all-reduce per step over NCCL (uni3detr_b200/train.py), AdamW. Not the headline metric:
    reported as its own JSON line with `"mode": "train"`."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from uni3detr_b200 import ops, sharding, synth
    from uni3detr_b200.train import DataParallelTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --train: no CUDA device")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = args.workload
    B = args.batch or 4                                   # samples_per_gpu of the reference configs
    K, W = args.steps, max(args.warmup, 3)
    model, cfg = synth.build_model(wl, seed=0)
    model = model.to(dev).train()
    trainer = DataParallelTrainer(model)
    pcr = np.asarray(cfg["pts_voxel_layer"]["point_cloud_range"], np.float64)
    ncls = cfg["pts_bbox_head"]["num_classes"]
    mine = sharding.scene_indices(B * world, rank, world)
    pts = [torch.from_numpy(synth.make_scene(wl, i)).to(dev) for i in mine]
    rng = np.random.default_rng(7 + rank)
    gts, gls = [], []
    for _ in mine:
        n = int(rng.integers(3, 9))
        ctr = pcr[:3] + (0.2 + 0.6 * rng.random((n, 3))) * (pcr[3:] - pcr[:3])
        box = np.concatenate([ctr, 0.4 + rng.random((n, 3)), (rng.random((n, 1)) - 0.5) * 3], 1).astype(np.float32)
        gts.append(torch.from_numpy(box).to(dev))
        gls.append(torch.from_numpy(rng.integers(0, ncls, n)).to(dev))
    first = None
    for _ in range(W):
        l = trainer.step(pts, gts, gls)
        first = first or {k: float(v) for k, v in l.items()}
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = ops.launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(K):
        l = trainer.step(pts, gts, gls)
    e.record()
    torch.cuda.synchronize()
    launches = ops.launch_count() - n0
    scenes, t_max, _ = sharding.reduce_metrics(B * K, s.elapsed_time(e) / 1e3, 0.0, device=dev)
    if rank == 0:
        last = {k: float(v) for k, v in l.items()}
        emit({"mode": "train", "metric": "train scenes/sec", "value": scenes / t_max, "unit": "scenes/s", "n_gpus": world,
              "steps": K, "warmup": W, "ms_per_step": 1e3 * t_max / K, "higher_is_better": True, "scaling": "weak",
              "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
              "config": {"workload": WORKLOADS[wl][2] + ", forward_train + backward + AdamW", "scenes_per_step_per_gpu": B,
                         "parallelism": f"scenes sharded x{world}, one flat gradient all-reduce per step"},
              "collective": {"all_reduce_per_step": trainer.collectives_per_step, "bytes": int(trainer.flat.numel() * 4),
                             "backend": "nccl" if world > 1 else None},
              "gpu_launches": launches, "loss_total_first": sum(first.values()), "loss_total_last": sum(last.values())})
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=WORKLOAD, choices=sorted(WORKLOADS),
                    help="BASELINE config of the headline line (default: config 2, the metric's configuration)")
    ap.add_argument("--batch", type=int, default=0, help="scenes per step per GPU (default: per workload, 32 for sunrgbd)")
    ap.add_argument("--dtype", default=None, choices=["bf16", "fp32"], help="default: the workload's stated dtype")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the extra blocks (batch 1 / 4 of the headline workload, the other BASELINE configs)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-scenes", type=int, default=10, help="bounded CPU-baseline sample (scenes, ~1.1 s each on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the per-op roofline pass (ncu runs)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--postprocess", action="store_true",
                    help="also run get_bboxes' device post-processing (per-class NMS) inside the step")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="e2e: serialise H2D -> replay -> D2H per step instead of overlapping neighbouring steps")
    ap.add_argument("--train", action="store_true", help="time data-parallel training steps instead of the forward")
    args = ap.parse_args()
    if args.train:
        return run_train(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
