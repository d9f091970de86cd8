#!/usr/bin/env python
"""bench.py -- BEV input generation throughput on B200 (BASELINE.json metric: BEV frames/s, Gpoints/s,
fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5]

One "step" = one pass of the hot path over one batch of synthetic frames per GPU.  Default workload =
BASELINE.json configs[1]: multi-sweep aggregation, 10 sweeps x 100 000 points, batch 32 per GPU
(the other configs are selectable and are parity-test cases).  For N > 1 launch under torchrun (one
process per GPU); frames shard independently, no collective on the data path (weak scaling); NCCL only
gathers the timing (max over ranks).

Printed JSON (rank 0, one line): the driver contract plus
  roofline      dominant kernel: achieved algorithmic GB/s (live CUDA-event time of that kernel inside the
                timed region) vs the measured HBM peak in MEASURED_PEAKS.json
  kernels       per-kernel ms/step and launches/step measured the same way
  e2e           same metric through the public host-facing API: pinned host buffers -> H2D -> kernels ->
                device-resident model inputs + D2H read of a per-frame probe, all inside the timed region
  cpu_baseline  the reference's numpy/OpenCV path (oracle/numpy_port.py) timed on this box's host cores
`--impl reference` times that CPU path alone (all host cores, one frame per worker process per step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

LIDAR_FRAME_BYTES = 290 * 400 * 720 * 4      # 334 080 000
MAP_FRAME_BYTES = 9 * 400 * 720 * 4          # 10 368 000
FALLBACK_HBM_GBS = 6650.0                    # /opt/skills/guides/B200_PROFILING.md fallback

WORKLOADS = {
    "c1": dict(desc="configs[0]: the reference's CPU-runnable case: one frame, ONE loaded sweep of 100k pts (output still 290 channels)",
               batch=1, map=False, aug=False, loaded_sweeps=1),
    "c2": dict(desc="configs[1]: multi-sweep aggregation, 10 sweeps x 100k pts, synthetic ego poses", batch=32, map=False, aug=False),
    "c3": dict(desc="configs[2]: LiDAR + HD-map rasterized channels (425 lanes, 20 crosswalks)", batch=32, map=True, aug=False),
    "c4": dict(desc="configs[3]: dense/skewed scene, 1M pts/frame in 4 clusters", batch=32, map=False, aug=False),
    "c5": dict(desc="configs[4]: LiDAR + map + flip/rotate/scale/dropout augmentation", batch=16, map=True, aug=True),
}


# ----------------------------------------------------------------------------------- workload
def make_frames(workload: str, batch: int, points: int, rank: int):
    import bevgen_b200 as bg
    syn = bg.synthetic
    md = syn.synthetic_map() if WORKLOADS[workload]["map"] else None
    frames = []
    for i in range(batch):
        idx = rank * batch + i
        if workload == "c1":
            frames.append(syn.frame_c1(idx, n=points))
        elif workload == "c2":
            frames.append(syn.frame_c2(idx, n=points))
        elif workload == "c4":
            frames.append(syn.frame_c4(idx, n=points))
        else:
            frames.append(syn.frame_c3(idx, n=points, map_data=md))
    return frames, md


POINT_FORMATS = {
    "av2": (7, "AV2 sweep columns as stored on disk and read by dataset.py:311-321: float16 x,y,z + uint8 intensity, 7 B/point"),
    "f16": (8, "rows of 4 x float16 (x,y,z,intensity), 8 B/point"),
    "f32": (16, "rows of 4 x float32 (x,y,z,intensity), 16 B/point"),
}


def workload_config(args) -> dict:
    """The workload, identically on both arms (`--impl ours` / `--impl reference`): what one step processes per GPU."""
    wl = WORKLOADS[args.workload]
    B = args.batch or wl["batch"]
    pb = POINT_FORMATS[args.points_dtype][0]
    loaded = wl.get("loaded_sweeps", 10)
    inputs_mb = B * loaded * args.points * pb / 1e6
    return {"workload": wl["desc"], "batch_per_gpu": B, "sweeps": 10, "loaded_sweeps": loaded, "points_per_sweep": args.points,
            "points_per_frame": loaded * args.points, "points_format": POINT_FORMATS[args.points_dtype][1],
            "point_values": "synthetic float32 draws" + ("" if args.points_dtype == "f32" else
                                                         " rounded to float16-representable coordinates (same values on both arms)"),
            "arithmetic": "float64 pose transform and binning, float32 intensities (the reference's own precision and op order)",
            "l2": f"no flush needed: per step inputs {inputs_mb:.0f} MB + outputs {B * LIDAR_FRAME_BYTES / 1e9:.1f} GB >> 126 MB L2",
            "sharding": "frames independent per rank, no data-path collective"}


def frame_bytes(frame_points: int, with_map: bool, map_verts: int, point_bytes: int = 16) -> int:
    """algorithmic bytes per frame."""
    b = frame_points * point_bytes + 10 * 128 + LIDAR_FRAME_BYTES
    if with_map:
        b += MAP_FRAME_BYTES + map_verts * 16
    return b


# ----------------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    """One frame through the numpy/OpenCV port, exactly the stages Dataset.__getitem__ runs."""
    workload, points, idx, map_path, seed, quantize = args
    import random as _r
    import bevgen_b200 as bg
    from oracle import numpy_port as npp
    syn = bg.synthetic
    f = syn.frame_c1(idx, n=points) if workload == "c1" else syn.frame_c4(idx, n=points) if workload == "c4" else \
        (syn.frame_c2(idx, n=points) if workload == "c2" else syn.frame_c3(idx, n=points, map_data={}))
    if quantize:
        f = syn.quantize_f16(f)          # the same point values our arm consumes
    t0 = time.perf_counter()
    lid = npp.frame_lidar(f.points4, f.sweep_offsets, f.sweep_poses, f.current_pose)
    if map_path:
        keys = ["tx_m", "ty_m", "tz_m", "qx", "qy", "qz", "qw"]
        mp = npp.map_raster(map_path, dict(zip(keys, f.current_pose.tolist())))
        if workload == "c5":
            _r.seed(seed)
            p = bg.augment.draw_aug_params(_r)
            npp.augment(lid, mp, p.flip, p.angle_deg, p.scale, p.rects)
    return time.perf_counter() - t0


def cpu_reference(workload: str, points: int, steps: int, warmup: int, workers: int, budget_s: float = 150.0,
                  quantize: bool = False):
    """frames/s of the CPU path with `workers` processes, each step = one frame per worker."""
    import multiprocessing as mp
    import tempfile
    import bevgen_b200 as bg
    map_path = None
    if WORKLOADS[workload]["map"]:
        map_path = os.path.join(tempfile.gettempdir(), f"bevgen_bench_map_{os.getpid()}.json")
        bg.synthetic.write_map_json(bg.synthetic.synthetic_map(), map_path)
    ctx = mp.get_context("fork")
    step_times = []
    t_begin = time.perf_counter()
    with ctx.Pool(workers) as pool:
        for s in range(warmup + steps):
            jobs = [(workload, points, s * workers + w, map_path, 1000 * 5 + s * workers + w, quantize) for w in range(workers)]
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs, chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup:
                step_times.append(dt)
            if time.perf_counter() - t_begin > budget_s and len(step_times) >= 1:
                break
    if map_path:
        os.unlink(map_path)
    total = sum(step_times)
    return workers * len(step_times) / total, total / len(step_times), len(step_times)


def cpu_single_process(workload: str, points: int, frames: int = 2, quantize: bool = False):
    """frames/s of the CPU path as the reference itself runs it: ONE process (DataLoader num_workers=0, train_vit.py:20),
    library-default threading (OpenBLAS / OpenCV pools) -- SURVEY.md 8(d)'s variant (i).  Best of `frames` frames."""
    import tempfile
    import bevgen_b200 as bg
    map_path = None
    if WORKLOADS[workload]["map"]:
        map_path = os.path.join(tempfile.gettempdir(), f"bevgen_bench_map1_{os.getpid()}.json")
        bg.synthetic.write_map_json(bg.synthetic.synthetic_map(), map_path)
    best = min(_cpu_worker((workload, points, 900 + i, map_path, 1000 * 5 + 900 + i, quantize)) for i in range(frames))
    if map_path:
        os.unlink(map_path)
    return 1.0 / best


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = min(host_threads(), 32)
    wl = WORKLOADS[args.workload]
    fps, step_s, nsteps = cpu_reference(args.workload, args.points, args.steps, args.warmup, workers,
                                        quantize=args.points_dtype != "f32")
    pts_frame = wl.get("loaded_sweeps", 10) * args.points
    single = cpu_single_process(args.workload, args.points, 2, quantize=args.points_dtype != "f32")
    line = {
        "impl": "reference", "metric": "bev_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": nsteps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "gpoints_per_s": fps * pts_frame / 1e9,
        "config": workload_config(args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": workers, "kind": "port",
                         "sample": f"{nsteps} steps x {workers} frames of the configured workload: the reference's numpy/OpenCV call "
                                   f"sequence (oracle/numpy_port.py), one frame per worker process per step, {workers} worker "
                                   "processes with library-default threads each (the reference itself is single-process)",
                         "single_process_value": single,
                         "single_process_note": "the same call sequence in ONE process with library-default threading, as the "
                                                "reference's own loader runs it (num_workers=0, train_vit.py:20); best of 2 frames"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int) -> None:
        self.proc = None
        self.idx = gpu_index

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.splitlines():
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------- ours
def run_ours(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist
    import bevgen_b200 as bg
    from bevgen_b200 import _native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _native.lib()
    numa = bg.sharding.bind_to_gpu_numa_node(local)

    wl = WORKLOADS[args.workload]
    B = args.batch or wl["batch"]
    algo = {"auto": bg.ALGO_AUTO, "atomic": bg.ALGO_ATOMIC, "tiled": bg.ALGO_TILED, "memset": bg.ALGO_MEMSET_ATOMIC}[args.algo]
    frames, md = make_frames(args.workload, B, args.points, rank)
    if args.points_dtype != "f32":
        frames = [bg.synthetic.quantize_f16(f) for f in frames]
    table = bg.MapTable([bg.maps.parse_map_dict(md)]) if wl["map"] else None
    aug = None
    if wl["aug"]:
        import random
        aug = []
        for i in range(B):
            random.seed(1000 * 5 + rank * B + i)
            aug.append(bg.draw_aug_params(random))
        frames, aug, _ = bg.partition_for_augmentation(frames, aug)     # resampled frames first: no identity copies
    batch = bg.FrameBatch.pack(frames, 10, map_ids=[0] * B if wl["map"] else None,
                               points_dtype={"f16": np.float16, "f32": np.float32, "av2": "av2"}[args.points_dtype])
    gen = bg.BevGenerator(table, algo=algo, device=dev, fused_augment=not args.dense_augment)
    S = 10
    out_l = torch.empty((B, S * 29, 400, 720), dtype=torch.float32, device=dev)
    out_m = torch.empty((B, 9, 400, 720), dtype=torch.float32, device=dev) if wl["map"] else None
    dvc = gen.upload(batch)                       # device-resident inputs for the kernel-only metric
    torch.cuda.synchronize()

    # ---- parity gate before any timing: frame 0 of this rank vs the C oracle, bit for bit ----------------
    parity = "skipped"
    if not args.no_parity:
        from oracle import c_oracle as co
        r = gen.generate(batch, aug, dev=dvc, with_map=wl["map"], out_lidar=out_l, out_map=out_m)
        f0 = frames[0]
        want = co.lidar_bev_fused(f0.points4, f0.sweep_offsets, f0.rel_tf, S)
        wantm = None
        if wl["map"]:
            T3 = bg.pose.map_matrix_from_pose(f0.current_pose[0], f0.current_pose[1], f0.current_pose[3:7])
            wantm = co.map_raster_flat(co.flatten_map(md), T3)
        if aug is not None:
            p = aug[0]
            for name in ("l", "m"):
                g = want if name == "l" else wantm
                if p.flip: g = co.flip_w(g)
                if p.angle_deg is not None: g = co.rotate(g, p.angle_deg)
                if p.scale is not None: g = co.scale(g, p.scale)
                g = np.ascontiguousarray(g).copy()
                if p.rects: co.dropout(g, p.rects)
                if name == "l": want = g
                else: wantm = g
        ok = bool(torch.equal(r["lidar_bev"][0].cpu(), torch.from_numpy(want)))
        if wl["map"]:
            ok = ok and bool(torch.equal(r["map_bev"][0].cpu(), torch.from_numpy(wantm)))
        if not ok:
            raise RuntimeError("parity check against the C oracle FAILED; refusing to time a wrong kernel")
        parity = "bit-exact vs oracle (frame 0)"

    def step_resident():
        gen.generate(batch, aug, dev=dvc, with_map=wl["map"], out_lidar=out_l, out_map=out_m)

    probe_host = torch.empty((B, 720), dtype=torch.float32).pin_memory()

    def step_e2e():
        r = gen.generate(batch, aug, with_map=wl["map"], out_lidar=out_l, out_map=out_m)   # H2D from pinned inside
        probe_host.copy_(r["lidar_bev"][:, 14, 200, :], non_blocking=True)              # D2H read of the result

    # pipelined e2e: H2D of step i+1 (copy stream) overlaps the kernels of step i; every step still uploads its
    # inputs from pinned host memory and reads its probe back, all inside the timed region
    pf = bg.PrefetchingGenerator(gen)
    out_l2 = torch.empty_like(out_l)
    out_m2 = torch.empty_like(out_m) if out_m is not None else None

    def run_e2e_pipelined(steps):
        h = pf.submit(batch)
        done = []                                   # completion events: the host stays at most --e2e-lookahead steps ahead of the GPU
        for i in range(steps):                      # (bounded staging memory, as a training loop's own sync would impose)
            if args.e2e_lookahead and len(done) >= args.e2e_lookahead:
                done.pop(0).synchronize()
            cur, h = h, (pf.submit(batch) if i + 1 < steps else None)
            ol, om = (out_l, out_m) if i % 2 == 0 else (out_l2, out_m2)
            r = pf.run(cur, aug, with_map=wl["map"], out_lidar=ol, out_map=om)
            probe_host.copy_(r["lidar_bev"][:, 14, 200, :], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            done.append(ev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- kernel-only metric (inputs resident in HBM), per-kernel event timing on --------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                           # nvidia-smi sampling spans warm-up + all timed regions below
        time.sleep(0.3)
    launches0 = _native.launch_count()
    _native.timing_enable(True)
    _native.timing_collect()
    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    _native.timing_collect()                      # drop warm-up records
    launches0 = _native.launch_count()
    total_ms = timed(step_resident, args.steps, 0)
    launches = _native.launch_count() - launches0
    ktimes = _native.timing_collect()
    _native.timing_enable(False)
    # ---- same loop without the timing events: the headline device-resident number ------------------------
    total_ms_plain = timed(step_resident, args.steps, 1)
    # ---- end to end through the host-facing API -----------------------------------------------------------
    # strict reference signature (numpy lists in -> fresh numpy array out, one frame per call): includes the 334 MB D2H
    strict_fps = install_fps = None
    if world == 1 and args.workload in ("c1", "c2", "c4"):
        f0 = frames[0]
        S0 = f0.num_sweeps
        # the arrays dataset.py:320-321 hands over: sweep_df[['x','y','z']].values (column-major, float16 in AV2) and intensity
        rdt = np.float32 if args.points_dtype == "f32" else np.float16
        raw = [np.asfortranarray(f0.sweep(s)[:, :3].astype(rdt)) for s in range(S0)]
        pl = [bg.transform_points(raw[s], f0.rel_tf[s]) for s in range(S0)]
        il = [f0.sweep(s)[:, 3].copy() for s in range(S0)]
        for _ in range(2):                          # warm-up: the pinned result buffers are allocated once, then recycled
            bg.create_intentnet_lidar_bev(pl, il)
        t0 = time.perf_counter()
        for _ in range(6):
            pl = [bg.transform_points(raw[s], f0.rel_tf[s]) for s in range(S0)]
            bg.create_intentnet_lidar_bev(pl, il)
        strict_fps = 6 / (time.perf_counter() - t0)
        # the same calls as dataset_shim.install() binds them inside the reference's dataset module: transform_points defers
        # (dataset.py:340) and create_intentnet_lidar_bev (dataset.py:347) runs the fused launch on the raw points
        tp = bg.bev_ops.transform_points_deferred
        for _ in range(2):
            bg.create_intentnet_lidar_bev([tp(raw[s], f0.rel_tf[s]) for s in range(S0)], il)
        t0 = time.perf_counter()
        for _ in range(6):
            bg.create_intentnet_lidar_bev([tp(raw[s], f0.rel_tf[s]) for s in range(S0)], il)
        install_fps = 6 / (time.perf_counter() - t0)
    # ---- sliding-window stream through the device-resident sweep store (AV2 columns only) --------------------------------
    # The reference's samples are 10-sweep sliding windows over a log (dataset.py:174-211): consecutive samples share 9 of
    # their 10 sweeps.  Here a step is B CONSECUTIVE samples of a synthetic log; every sweep is uploaded once (store.put, own
    # copy stream) and named by the up to 10 samples that use it, so a step moves B new sweeps over PCIe instead of 10 B.
    # Same kernels, same per-frame work (every sample still transforms and bins all of its 10 sweeps).
    sw = None
    if args.points_dtype == "av2" and not args.no_sliding_window and args.workload != "c1":
        sw = {}
        pool_n = 64                                                   # distinct synthetic sweeps, cycled (every key is a new upload)
        sweeps, poses = bg.synthetic.sliding_log(pool_n, n=args.points, seed=7 + rank, dense=args.workload == "c4")
        pool = [bg.PointColumns.from_points4(p) for p in sweeps]     # pinned host columns, as sweep_io's column layout gives them
        def col(pc, k):
            b, st, n = pc.buf, pc.stride, pc.n
            return (b[2 * k * st:2 * k * st + 2 * n].view(torch.float16) if k < 3 else b[6 * st:6 * st + n])
        pool_cols = [tuple(col(pc, k) for k in range(4)) for pc in pool]
        total_steps = args.warmup + args.steps * max(1, args.e2e_repeats) + 4
        n_log = total_steps * B + S
        log_poses = bg.synthetic._poses(n_log)
        if wl["map"]:
            log_poses[:, 0] += bg.synthetic.EGO_XY[0]; log_poses[:, 1] += bg.synthetic.EGO_XY[1]
        store = bg.SweepStore(capacity_points=(3 * B + S) * (args.points + 8), device=dev)
        # pose prep (host, the reference's own numpy / scipy calls) is done up front, like FrameBatch.pack for the other loops
        def sample(i):
            rel = bg.pose.sweep_rel_transforms(log_poses[i + S - 1], log_poses[i:i + S])
            cp = log_poses[i + S - 1]
            return bg.StoreFrame([j for j in range(i, i + S)], log_poses[i:i + S], cp, rel_tf=rel,
                                 ego_T=bg.pose.map_matrix_from_pose(cp[0], cp[1], cp[3:7]))
        samples = [sample(i) for i in range(total_steps * B)]
        state = {"next_sweep": 0, "step": 0}
        def step_sliding(ol, om):
            i0 = state["step"] * B
            while state["next_sweep"] < i0 + B + S - 1:              # B new sweeps per step (B + 9 on the first)
                j = state["next_sweep"]
                store.put(j, *pool_cols[j % pool_n])
                state["next_sweep"] += 1
            sb = bg.StoreBatch.pack(store, samples[i0:i0 + B], S, map_ids=[0] * B if wl["map"] else None)
            r = gen.generate(sb, aug, with_map=wl["map"], out_lidar=ol, out_map=om)
            probe_host.copy_(r["lidar_bev"][:, 14, 200, :], non_blocking=True)
            state["step"] += 1
            return r
        # parity of the store path: its first step against the packed form of the same samples
        r = step_sliding(out_l, out_m)
        fr = [bg.synthetic.window_frame([sweeps[j % pool_n] for j in range(S + 1)], log_poses, i) for i in range(2)]
        chk = gen.generate(bg.FrameBatch.pack(fr, S, map_ids=[0, 0] if wl["map"] else None, points_dtype="av2"),
                           aug[:2] if aug is not None else None, with_map=wl["map"])
        if not torch.equal(chk["lidar_bev"], r["lidar_bev"][:2]):
            raise RuntimeError("sweep-store path differs from the packed path; refusing to time it")
        def run_sliding(steps):
            done = []
            h0 = store.h2d_bytes
            for i in range(steps):
                if args.e2e_lookahead and len(done) >= args.e2e_lookahead:
                    done.pop(0).synchronize()
                ol, om = (out_l, out_m) if i % 2 == 0 else (out_l2, out_m2)
                step_sliding(ol, om)
                ev = torch.cuda.Event(); ev.record(); done.append(ev)
            sw["h2d_bytes_per_step"] = (store.h2d_bytes - h0) / steps
    e2e_serial_ms = timed(step_e2e, args.steps, max(1, min(args.warmup, 2)))
    run_e2e_pipelined(2)
    e2e_runs = [timed(lambda: run_e2e_pipelined(args.steps), 1, 0) for _ in range(max(1, args.e2e_repeats))]
    e2e_ms = min(e2e_runs)
    if sw is not None:
        run_sliding(2)
        sw_runs = [timed(lambda: run_sliding(args.steps), 1, 0) for _ in range(max(1, args.e2e_repeats))]
        sw["ms"] = min(sw_runs); sw["runs"] = sw_runs
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    K = args.steps
    frames_total = world * B * K
    value = frames_total / (total_ms_plain / 1e3)
    pts_frame = int(batch.num_points / B)
    verts = int(table.verts.shape[0]) if table is not None else 0
    pbytes = POINT_FORMATS[args.points_dtype][0]
    fbytes = frame_bytes(pts_frame, wl["map"], verts, pbytes)

    kernels = {n: {"ms_per_step": ms / K, "launches_per_step": c / K} for n, (c, ms) in ktimes.items()}
    own_bytes = {   # algorithmic bytes one launch of each kernel must move, per frame
        "k_zero_fill_tma": LIDAR_FRAME_BYTES, "cudaMemsetAsync": LIDAR_FRAME_BYTES, "k_scatter_atomic": pts_frame * pbytes,
        "k_bin_fixed": pts_frame * pbytes, "k_tile_store": LIDAR_FRAME_BYTES, "k_overflow_red": 0,
        "k_map_expand": MAP_FRAME_BYTES, "k_augment": 2 * (LIDAR_FRAME_BYTES + (MAP_FRAME_BYTES if wl["map"] else 0)),
        "k_tile_aug": LIDAR_FRAME_BYTES, "k_rebin_aug": 0, "k_overflow_total": 0,
    }
    n_gather = sum(1 for p in (aug or []) if p.needs_gather)
    fused_aug = bool(aug) and "k_tile_aug" in kernels
    if fused_aug:       # the gather then only writes the map grids of the resampled frames (source: L2-resident bit planes)
        own_bytes["k_augment"] = MAP_FRAME_BYTES
    peak = FALLBACK_HBM_GBS
    peak_src = "fallback"
    mp_path = ROOT / "MEASURED_PEAKS.json"
    if mp_path.exists():
        try:
            peak = float(json.loads(mp_path.read_text())["hbm_gbs"]); peak_src = "measured"
        except Exception:
            pass
    roofline = None
    mine = {n: k for n, k in kernels.items() if n != "cudaMemsetAsync"} or kernels
    if mine:
        dom = max(mine, key=lambda n: mine[n]["ms_per_step"])
        k = kernels[dom]
        # algorithmic bytes the kernel must move per STEP (it may take several launches: sub-ranges, frame classes)
        if dom in ("k_augment", "k_tile_aug"):
            bytes_step = n_gather * own_bytes[dom]
        elif dom == "k_tile_store" and fused_aug:
            bytes_step = (B - n_gather) * own_bytes[dom]
        else:
            bytes_step = own_bytes.get(dom, fbytes) * B
        per_step = max(k["launches_per_step"], 1e-9)
        achieved = bytes_step / (k["ms_per_step"] / 1e3) / 1e9
        traffic = None
        tpath = ROOT / "profiles" / "traffic.json"
        if tpath.exists() and args.workload == "c2" and B == WORKLOADS["c2"]["batch"]:
            try:
                traffic = json.loads(tpath.read_text()).get(dom)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": bytes_step / per_step,
                    "ms_per_launch": k["ms_per_step"] / per_step,
                    "peak_note": "peak = measured copy bandwidth (read + write); a write-only stream can exceed it "
                                 "(zero-fill alone: 7.5 TB/s, profiles/r1_fillbench.txt)"}
    step_gbs = fbytes * B / (total_ms_plain / K / 1e3) / 1e9

    e2e_value = frames_total / (e2e_ms / 1e3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        workers = min(host_threads(), 32)
        cfps, cstep, cn = cpu_reference(args.workload, args.points, 2, 1, workers, budget_s=60.0,
                                        quantize=args.points_dtype != "f32")
        cpu = {"value": cfps, "unit": "frames/s", "cores": workers, "kind": "port",
               "sample": f"{cn} steps x {workers} frames of the same workload through oracle/numpy_port.py "
                         f"({workers} worker processes, one frame each per step; {cstep:.2f} s/step)",
               "single_process_value": cpu_single_process(args.workload, args.points, 2, quantize=args.points_dtype != "f32"),
               "single_process_note": "the same call sequence in ONE process with library-default threading, as the reference's "
                                      "own loader runs it (num_workers=0, train_vit.py:20); best of 2 frames"}

    line = {
        "metric": "bev_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": total_ms_plain / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "impl_notes": {"algo": args.algo, "points_dtype": args.points_dtype, "parity": parity, "host_affinity": numa},
        "gpoints_per_s": value * pts_frame / 1e9,
        "bytes_per_frame": fbytes, "step_algorithmic_gbs": step_gbs, "step_frac_of_peak": step_gbs / peak,
        "roofline": roofline, "kernels": kernels, "ms_per_step_with_event_timing": total_ms / K,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": batch.h2d_bytes,
                "d2h_bytes_per_step": int(probe_host.numel() * 4), "ms_per_step": e2e_ms / K,
                "serial_value": frames_total / (e2e_serial_ms / 1e3), "strict_numpy_signature_value": strict_fps, "install_mode_value": install_fps,
                "runs_ms_per_step": [m / K for m in e2e_runs],
                "sliding_window": None if sw is None else {
                    "value": frames_total / (sw["ms"] / 1e3), "unit": "frames/s", "ms_per_step": sw["ms"] / K,
                    "h2d_bytes_per_step": int(sw["h2d_bytes_per_step"]) + B * (S * 4 + (S + 1) * 4 + S * 128 + 72),
                    "runs_ms_per_step": [m / K for m in sw["runs"]],
                    "note": "same metric for a STREAM of consecutive 10-sweep sliding-window samples of one log (the reference's own "
                            "sample structure, dataset.py:174-211): every sweep is uploaded once into a device-resident store "
                            "(bevgen_lidar_bev_store_cols) and named by the up to 10 samples that use it, so a step uploads B new sweeps "
                            "instead of 10 B; per-frame work is unchanged.  NOT the headline e2e, which uploads all 10 sweeps of every frame"},
                "note": "pinned host frames -> H2D (copy stream, overlapped with the previous step's kernels) -> kernels -> "
                        "device-resident [B,290,400,720] model input; D2H = per-frame probe row; serial_value = same without overlap; "
                        "strict_numpy_signature_value = transform_points x10 + create_intentnet_lidar_bev per frame with numpy in/out (334 MB D2H per frame); "
                        "install_mode_value = the same two calls as dataset_shim.install() binds them (transform deferred into the fused launch)"},
        "gpu_launches": int(launches), "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--algo", default="auto", choices=["auto", "atomic", "tiled", "memset"],
                    help="auto (default): tiled, switching to atomic while the scenes keep overflowing the tile bins (config 4)")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--points", type=int, default=100_000, help="points per sweep")
    ap.add_argument("--points-dtype", default="av2", choices=["av2", "f16", "f32"],
                    help="host/device format of the input points.  av2 (default): AV2's on-disk sweep columns, float16 x,y,z + uint8 "
                         "intensity, 7 B/point -- the dtypes the reference's Dataset reads (dataset.py:311-321); f16: 4 x float16 rows, "
                         "8 B/point; f32: float32 rows, 16 B/point with unrounded coordinates.  av2 / f16 round the synthetic coordinates "
                         "to float16-representable values on BOTH arms; all arithmetic stays float64 / float32 as in the reference")
    ap.add_argument("--dense-augment", action="store_true",
                    help="config 5: take the dense route (generate -> gather) instead of the fused generation+augmentation kernel")
    ap.add_argument("--e2e-repeats", type=int, default=3,
                    help="the pipelined e2e region is timed this many times; e2e.value is the fastest, e2e.runs_ms_per_step lists all "
                         "(the host->device rate of the shared box dips intermittently by 20-40 %%)")
    ap.add_argument("--e2e-lookahead", type=int, default=2,
                    help="e2e loop: steps the host may run ahead of the GPU (bounds the staging memory; 0 = unbounded)")
    ap.add_argument("--no-sliding-window", action="store_true", help="skip the sweep-store (sliding-window stream) e2e measurement")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
