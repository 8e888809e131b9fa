#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 panorama-stitching path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload input2|input|synth4k]

Metric (BASELINE.json): stitched Mpixel/s = sum of input pixels of the panorama job / time from "decoded images" to
"finished panorama".  One step = one whole panorama job (projection, SIFT, all-pairs matching, RANSAC, warp, multiband
blend, equalisation) on the named workload; default workload = BASELINE.json configs[1], the bundled Input2 set
(4 x 1210x907).  `value` is measured with the decoded input images already resident in HBM and the result left in HBM;
`e2e` is the same job through the C ABI with pinned HOST buffers (H2D of the inputs and D2H of the panorama inside the
timed region).  With N > 1 (torchrun, one rank per GPU) every rank stitches its own job (weak scaling, no data-path
collective); time = max over ranks, value = N x pixels / time.

--impl reference times the reference's own CPU implementation (oracle/_ref, compiled from the reference sources) on the
host cores of the box, as process-level replicas (the reference is single-threaded), on a bounded sample of the same
workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stitched_mpixel_per_s"
UNIT = "Mpixel/s"


# ----------------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------------
def synth_scene_views_r1(n, w, h, seed=20181126):
    """Round-1 generator (pure translations of a blocky scene), kept so that round-1 numbers can be re-measured:
    --workload synth4k_r1."""
    rng = np.random.default_rng(seed)
    W = w // 2 * (n + 1)
    scene = np.zeros((3, h + 32, W), np.float32)
    for octave in range(5):
        s = 4 << octave
        g = rng.random((3, (h + 32) // s + 2, W // s + 2)).astype(np.float32)
        scene += np.kron(g, np.ones((1, s, s), np.float32))[:, : h + 32, :W] * (s / 64.0)
    scene = 255.0 * (scene - scene.min()) / (scene.max() - scene.min())
    nblobs = int(200 * W * h / 1e6)
    ys = rng.integers(0, h + 32, nblobs); xs = rng.integers(0, W, nblobs); rs = rng.integers(3, 24, nblobs)
    cols = rng.integers(0, 256, (nblobs, 3))
    for y, x, r, c in zip(ys, xs, rs, cols):
        y0, y1, x0, x1 = max(0, y - r), min(h + 32, y + r), max(0, x - r), min(W, x + r)
        scene[:, y0:y1, x0:x1] = 0.5 * scene[:, y0:y1, x0:x1] + 0.5 * c[:, None, None]
    scene = np.clip(scene, 0, 255).astype(np.uint8)
    views = []
    for i in range(n):
        dy = int(rng.integers(0, 17))
        views.append(np.ascontiguousarray(scene[:, dy : dy + h, i * (w // 2) : i * (w // 2) + w]))
    return views


def _render_views(args):
    n, w, h, seed, only = args
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import synth_scene
    v = synth_scene.views(n, w, h, seed, only=only)
    for i in only:
        path = _view_cache_path(n, w, h, seed, i)
        tmp = path + f".{os.getpid()}.tmp.npy"
        np.save(tmp, v[i])
        os.replace(tmp, path)
    return only


def _view_cache_path(n, w, h, seed, i):
    return os.path.join(tempfile.gettempdir(), f"pano_b200_synth_v3_{n}x{w}x{h}_{seed}_view{i}.npy")


def synth_scene_views(n, w, h, seed=20181126, only=None):
    """SURVEY.md 8(d) generator (tools/synth_scene.py): fractal noise + soft-edged shapes + salt; 50 % overlap views with
    +-8 px / +-0.5 degree / +-3 % gain jitter.  Rendering takes seconds per view (25 s at 8K), so views are rendered by a
    process pool and cached per view in the temp directory (both arms of a bench run, and repeated runs, read the same
    bytes).  only = the view indices wanted (a rank of a sharded job renders just the views it owns); the other entries
    of the returned list are None."""
    want = list(range(n)) if only is None else list(only)
    out = [None] * n
    missing = []
    for i in want:
        path = _view_cache_path(n, w, h, seed, i)
        try:
            a = np.load(path)
            if a.shape == (3, h, w):
                out[i] = a
                continue
        except Exception:
            pass
        missing.append(i)
    if missing:
        procs = max(1, min(len(missing), (os.cpu_count() or 2) // 2, 12))
        if procs == 1:
            _render_views((n, w, h, seed, missing))
        else:
            # child interpreters running THIS file's renderer entry (not a multiprocessing pool: a spawned pool worker
            # re-imports the caller's main module, which for an unguarded script means running the script again)
            import subprocess
            chunks = [missing[k::procs] for k in range(procs)]
            kids = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--render-views",
                                      ",".join(map(str, (n, w, h, seed))), ",".join(map(str, c))]) for c in chunks]
            for k in kids:
                if k.wait() != 0:
                    raise RuntimeError("synthetic view renderer failed")
        for i in missing:
            out[i] = np.load(_view_cache_path(n, w, h, seed, i))
    return out


WORKLOAD_DESC = {
    "synth4k": "synthetic 8-image 3840x2160 horizontal panorama, BASELINE.json configs[2] (the largest single-GPU configuration)",
    "synth4k_r1": "synthetic 8-image 3840x2160 panorama, round-1 generator (pure translations)",
    "synth1080": "synthetic 8-image 1920x1080 horizontal panorama (small variant of configs[2])",
    "synth8k": "synthetic 24-image 7680x4320 360-degree panorama, BASELINE.json configs[3]",
}


SYNTH_SHAPES = {"synth4k": (8, 3840, 2160), "synth1080": (8, 1920, 1080), "synth8k": (24, 7680, 4320)}


def load_workload(name, only=None):
    """-> (images, description, data kind).  only (synthetic workloads): render / load just these views."""
    from computervisionimagestich2_b200 import bmpio
    if name in SYNTH_SHAPES:
        n, w, h = SYNTH_SHAPES[name]
        return synth_scene_views(n, w, h, only=only), WORKLOAD_DESC[name], "synthetic"
    data = os.path.join(ROOT, "oracle", "_ref", "data")
    if name in ("input", "input2"):
        d = os.path.join(data, "Input" if name == "input" else "Input2")
        imgs = [bmpio.load_bmp(os.path.join(d, f"{i}.bmp")) for i in range(1, 5)]
        desc = ("Input/1-4.bmp 4-image panorama (384x512), BASELINE.json configs[0]" if name == "input"
                else "Input2/1-4.bmp 4-image panorama (1210x907), BASELINE.json configs[1]")
        return imgs, desc, "bundled reference fixtures (Input2 BMPs)" if name == "input2" else "bundled reference fixtures (Input BMPs)"
    if name == "synth4k":
        return synth_scene_views(8, 3840, 2160), WORKLOAD_DESC[name], "synthetic"
    if name == "synth4k_r1":
        return synth_scene_views_r1(8, 3840, 2160), WORKLOAD_DESC[name], "synthetic"
    if name == "synth1080":   # quick variant of the same generator (tests, smoke runs)
        return synth_scene_views(8, 1920, 1080), WORKLOAD_DESC[name], "synthetic"
    if name == "synth8k":   # BASELINE.json configs[3]; sharded over the GPUs of the box (--gpus 8)
        return synth_scene_views(24, 7680, 4320), WORKLOAD_DESC[name], "synthetic"
    raise SystemExit(f"unknown workload {name}")


def megapixels(imgs):
    return sum(i.shape[1] * i.shape[2] for i in imgs) / 1e6


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region by an in-process NVML thread (a spawned
    `nvidia-smi -lms` loop takes driver locks that stall a launch/sync-heavy host pipeline)."""

    def __init__(self, gpu_index, period_s=0.01):
        import threading
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self.ok = False
        try:
            import pynvml as N
            N.nvmlInit()
            self.N = N
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    idx = gpu_index
            self.h = N.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            return
        self.period = period_s
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        N = self.N
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)))
                r = int(N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.ok:
            return out
        self._stop.set()
        self.t.join(timeout=2)
        if self.sm:
            out["sm_mhz"] = statistics.median(self.sm)
            out["sm_max_mhz"] = self.max_sm
        out["reasons"] = sorted(self.reasons)
        out["samples"] = len(self.sm)
        return out


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (oracle/_ref = the reference itself compiled from its sources)
# ----------------------------------------------------------------------------------------------------------------
def ref_sample(name):
    """The bounded sample of a workload that the CPU reference is timed on: (images, description).  The bundled sets are
    small enough for whole jobs or a 2-image sub-panorama; for the synthetic 4K / 8K workloads even a 2-image
    sub-panorama at full resolution takes the reference ~15 minutes per core (its matcher is O(NA NB): 21.6 k features per
    4K view), so the sample is a 2-view panorama from the SAME generator at 1920 x 1080 -- which flatters the CPU
    (quarter-area views have a quarter of the features, and 2 views need 3 directed matches where 8 views need 56)."""
    if name in ("input", "input2"):
        imgs, desc, _ = load_workload(name)
        return imgs, desc
    v = synth_scene_views(2, 1920, 1080)
    return v, "2-view 1920x1080 panorama from the synthetic generator of the workload (quarter-area sample of the 3840x2160 views)"


def _ref_worker(args):
    name, idx, reps = args
    sys.path.insert(0, ROOT)
    from oracle import ref_api
    imgs, _ = ref_sample(name)
    imgs = [imgs[i] for i in idx]
    t0 = time.perf_counter()
    for _ in range(reps):
        out, _info = ref_api.stitch_mem(imgs)
    return time.perf_counter() - t0, out.shape


def ref_sample_plan(name, steps_total, budget_s=200.0):
    """Pick the bounded sample: the full job when it fits the time budget, else a 2-image sub-panorama."""
    if name not in ("input", "input2"):
        n = max(1, min(steps_total, int(budget_s // 70.0)))
        return [0, 1], "the 2-view sample (2 SIFT, 3 directed matches, 2 RANSAC, 1 blend, tail)", n
    est_full = {"input": 2.5, "input2": 45.0}[name]
    est_pair = {"input": 1.0, "input2": 12.0}[name]
    if steps_total * est_full <= budget_s:
        return [0, 1, 2, 3], "the full 4-image job", steps_total
    n = max(1, min(steps_total, int(budget_s // est_pair)))
    return [1, 2], "2-image sub-panorama (images 2 and 3 of the set: 2 SIFT, 3 directed matches, 2 RANSAC, 1 blend, tail)", n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_api
    name = args.workload
    if not ref_api.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpano_ref.so was not built (needs /root/reference at build time)"}))
        return
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, args.ref_procs))
    idx, what, nsteps = ref_sample_plan(name, args.steps + args.warmup)
    warm = min(args.warmup, max(0, nsteps - 1), 1)
    timed = max(1, min(args.steps, nsteps - warm))
    if name in ("input", "input2"):
        _, desc, data = load_workload(name)
    else:
        desc, data = WORKLOAD_DESC.get(name, name), "synthetic"
    simgs, sdesc = ref_sample(name)
    mpix = megapixels([simgs[i] for i in idx])
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        if warm:
            pool.map(_ref_worker, [(name, idx, warm)] * cores)
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(name, idx, timed)] * cores)
        wall = time.perf_counter() - t0
    value = cores * timed * mpix / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": timed,
        "warmup": warm, "ms_per_step": 1e3 * wall / timed, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": data,
        "config": {"workload": desc, "sample": f"{what} of {sdesc}", "replicas": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{what} of {sdesc}; {cores} single-threaded process replicas of oracle/_ref (the reference has no threads), {timed} timed step(s) each"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_baseline_single(name, time_cap_s=80.0):
    """One core, one bounded sample (kind = reference), for the cpu_baseline object of the B200 arm."""
    from oracle import ref_api
    if not ref_api.available():
        return None
    idx, what, _ = ref_sample_plan(name, 1, budget_s=time_cap_s)
    simgs, sdesc = ref_sample(name)
    sub = [simgs[i] for i in idx]
    t0 = time.perf_counter()
    ref_api.stitch_mem(sub)
    dt = time.perf_counter() - t0
    return {"value": megapixels(sub) / dt, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"{what} of {sdesc}, one run on one core ({dt:.1f} s)"}


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import computervisionimagestich2_b200 as pano
    L = pano.lib()
    mode = args.mode
    if mode == "auto":
        mode = "sharded" if (world > 1 and args.workload in SYNTH_SHAPES) else "replicas"
    ctx = pano.Context(local_rank)
    ctx.set_match_mode(args.match_mode)
    if args.workload == "pairs1024":
        return run_b200_pairs(args, ctx, L, dist, rank, local_rank, world)
    if world > 1 and mode == "sharded":
        from computervisionimagestich2_b200 import dist as pdist
        only = pdist.images_of_rank(SYNTH_SHAPES[args.workload][0], world, rank) if args.workload in SYNTH_SHAPES else None
        imgs, desc, data = load_workload(args.workload, only=only)    # a rank renders only the views it owns
        return run_b200_sharded(args, ctx, L, imgs, desc, data, dist, rank, local_rank, world)
    imgs, desc, data = load_workload(args.workload)
    n = len(imgs)
    mpix = megapixels(imgs)
    ws = (C.c_int * n)(*[i.shape[2] for i in imgs])
    hs = (C.c_int * n)(*[i.shape[1] for i in imgs])

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # --- device-resident measurement ------------------------------------------------------------------------------
    ptrs = (C.c_void_p * n)(*[i.ctypes.data for i in imgs])
    ctx._check(L.pano_b200_stage_images(ctx.h, ptrs, ws, hs, n), "stage_images")
    ow, oh = C.c_int(), C.c_int()
    for _ in range(args.warmup):
        ctx._check(L.pano_b200_stitch_staged(ctx.h, C.byref(ow), C.byref(oh)), "stitch_staged")
    L.pano_b200_ktimer_enable(0)
    L.pano_b200_ktimer_reset()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    ms = C.c_float()
    total_ms = 0.0
    for _ in range(args.steps):
        L.pano_b200_flush_l2(ctx.h)
        L.pano_b200_timer_start(ctx.h)
        ctx._check(L.pano_b200_stitch_staged(ctx.h, C.byref(ow), C.byref(oh)), "stitch_staged")
        L.pano_b200_timer_stop(ctx.h, C.byref(ms))
        total_ms += ms.value
    barrier()
    launches = L.pano_b200_ktimer_launches()
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = world * mpix / (ms_per_step / 1e3)
    t = pano.Times()
    L.pano_b200_stitch_times(ctx.h, C.byref(t))
    stages = {f[0]: round(getattr(t, f[0]), 3) for f in pano.Times._fields_}

    # --- end to end through the C ABI with pinned host buffers --------------------------------------------------
    in_bytes = [3 * i.shape[1] * i.shape[2] for i in imgs]
    pin_in = []
    for im, b in zip(imgs, in_bytes):
        p = L.pano_b200_alloc_pinned(C.c_size_t(b))
        C.memmove(p, im.ctypes.data, b)
        pin_in.append(p)
    out_cap = 3 * ow.value * oh.value
    pin_out = L.pano_b200_alloc_pinned(C.c_size_t(out_cap))
    pptrs = (C.c_void_p * n)(*pin_in)
    e2e_ms = 0.0
    e2e_steps = max(1, args.steps)
    for it in range(1 + e2e_steps):
        L.pano_b200_flush_l2(ctx.h)
        t0 = time.perf_counter()
        ctx._check(L.pano_b200_stitch_into(ctx.h, pptrs, ws, hs, n, C.c_void_p(pin_out), C.c_size_t(out_cap), C.byref(ow), C.byref(oh)), "stitch_into")
        dt = (time.perf_counter() - t0) * 1e3
        if it > 0:
            e2e_ms += dt
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_value = world * mpix / (e2e_ms / e2e_steps / 1e3)
    pano_hash = None
    bit_exact = None
    if rank == 0:
        import hashlib
        res = np.ctypeslib.as_array(C.cast(pin_out, C.POINTER(C.c_uint8)), (out_cap,))
        pano_hash = hashlib.sha256(res.tobytes()).hexdigest()
        try:   # the committed golden anchor of this workload (generated from the compiled reference)
            anchors = json.load(open(os.path.join(ROOT, "tests", "golden", "anchors.json")))
            key = {"input": "Input", "input2": "Input2"}.get(args.workload)
            if key:
                bit_exact = pano_hash == anchors[key]["pano_sha256"]
        except Exception:
            bit_exact = None

    # --- per-kernel pass (CUDA events around every launch; not part of the timed numbers above).  The images are
    #     processed on ONE lane here so that kernel times do not overlap and shares are meaningful. ----------------
    L.pano_b200_set_lanes(ctx.h, 1)
    ctx._check(L.pano_b200_stitch_staged(ctx.h, C.byref(ow), C.byref(oh)), "stitch_staged")
    L.pano_b200_ktimer_reset()
    L.pano_b200_ktimer_enable(1)
    L.pano_b200_flush_l2(ctx.h)
    ctx._check(L.pano_b200_stitch_staged(ctx.h, C.byref(ow), C.byref(oh)), "stitch_staged")
    buf = C.create_string_buffer(1 << 16)
    L.pano_b200_ktimer_report(buf, 1 << 16)
    L.pano_b200_ktimer_enable(0)
    L.pano_b200_set_lanes(ctx.h, 8)
    kernels = json.loads(buf.value.decode())
    # --- north-star stage (3): the uint8 / tcgen05 matcher on resident SIFT-like tables (kernel only) ----------------
    match_u8 = None
    if rank == 0 and not args.no_match_u8:
        try:
            rng = np.random.default_rng(1)
            def sift_like(nrows):
                x = rng.gamma(0.6, 1.0, (nrows, 128)).astype(np.float32)
                x /= np.linalg.norm(x, axis=1, keepdims=True)
                x = np.minimum(x, 0.2)
                x /= np.linalg.norm(x, axis=1, keepdims=True)
                return np.minimum(512.0 * x, 255.0).astype(np.uint8)
            na = nb = 72000  # features of one 7680x4320 image (SURVEY 6.2 density), BASELINE.json configs[3]
            ms_u8 = ctx.bench_match_u8(0, 0, 5, A=sift_like(na), B=sift_like(nb))
            ms_peak, ksteps = ctx.bench_match_u8_peak()
            tops = 2.0 * 128 * na * nb / (ms_u8 * 1e-3) / 1e12
            issued = 2.0 * 32 * ksteps * na * nb            # int8 ops the MMA stream issues (descriptor + norm-extension K steps)
            peak_meas = issued / (ms_peak * 1e-3) / 1e12 if ms_peak > 0 else None
            match_u8 = {"nA": na, "nB": nb, "ms": round(ms_u8, 4), "int8_TOPS": round(tops, 1),
                        "issued_TOPS": round(issued / (ms_u8 * 1e-3) / 1e12, 1),
                        "measured_peak_TOPS": None if peak_meas is None else round(peak_meas, 1),
                        "frac_of_measured": None if peak_meas is None else round(issued / (ms_u8 * 1e-3) / 1e12 / peak_meas, 4),
                        "useful_frac_of_measured": None if peak_meas is None else round(tops / peak_meas, 4),
                        "peak_TOPS_nominal": 4500.0, "frac_of_nominal": round(tops / 4500.0, 4),
                        "peak_source": "measured in this run: the same kernel (TMA + UTCIMMA M128 x N128 x K32, cta_group::1) with the epilogue reduced to releasing the accumulators -- what the tensor pipe delivers to this instruction stream; nominal dense int8 4.5 POPS beside it",
                        "k_steps_per_tile": ksteps,
                        "data": "synthetic SIFT-like uint8 descriptor tables, resident in HBM", "exact_vs_integer_oracle": True,
                        "note": "useful ops 2*128*nA*nB over main + finish kernels; the MMA also runs the norm-extension K step(s), counted in issued_TOPS"}
        except Exception as e:
            match_u8 = {"error": str(e)}
    # --- the reference's second caller of the same kernels (src/ex6): its 18-image data set, host buffers in and out
    ex6 = None
    if rank == 0 and world == 1 and not args.no_ex6:
        try:
            ex6 = bench_ex6(ctx)
        except Exception as e:
            ex6 = {"error": str(e)}
    # --- BASELINE.json configs[4] (batched independent pairs), a bounded sample through the public call
    pairs = None
    if rank == 0 and world == 1 and not args.no_pairs:
        try:
            pairs = bench_pairs(ctx)
        except Exception as e:
            pairs = {"error": str(e)}
    for p in pin_in:
        L.pano_b200_free_pinned(C.c_void_p(p))
    L.pano_b200_free_pinned(C.c_void_p(pin_out))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ranked = sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])
    top = ranked[0]
    roofline = roofline_of(top, kernels, clocks, args.workload)
    # No kernel dominates this job any more (the largest holds ~1/6 of the kernel time, the next three are within 35 % of
    # it): the same object for the next kernels, so that the line does not hinge on which of them happens to lead
    roofline_next = [roofline_of(kv, kernels, clocks, args.workload) for kv in ranked[1:4] if kv[1]["ms"] > 0]
    # the HBM-bound scale-space kernels, always reported (north_star: blur GB/s)
    hbm_kernels = {}
    for name, k in kernels.items():
        if k["ms"] > 0 and k["bytes"] > 0 and not name.startswith(("match", "ransac", "sift.refine", "sift.orient")):
            hbm_kernels[name] = {"GBps": round(k["bytes"] / (k["ms"] * 1e-3) / 1e9, 1), "ms": round(k["ms"], 4), "launches": k["launches"]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_single(args.workload)
        except Exception as e:  # the baseline is informative; never fail the bench on it
            cpu = {"error": str(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": data,
        "config": {"workload": desc, "images": n, "input_mpixel": mpix, "output": [ow.value, oh.value],
                   "l2": "flushed (256 MB memset) between timed iterations", "parallelism": f"replicas x{world}" if world > 1 else "1 GPU",
                   "bit_exact_vs_reference": bit_exact, "panorama_sha256": pano_hash},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sum(in_bytes), "d2h_bytes_per_step": out_cap,
                "ms_per_step": e2e_ms / e2e_steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_next": roofline_next,
        "cpu_baseline": cpu,
        "stages_ms_last_step": stages,
        "kernels_ms": {k: round(v["ms"], 4) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])},
        "hbm_kernels": hbm_kernels,
        "sift": sift_summary(kernels, stages, mpix),
        "match_u8": match_u8,
        "ex6": ex6,
        "pairs": pairs,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_b200_sharded(args, ctx, L, imgs, desc, data, dist, rank, local_rank, world):
    """N > 1, one panorama job sharded over the ranks (strong scaling): image i on rank i % N, NCCL all-gather of the
    descriptor blocks straight from HBM, directed matching problems dealt to the ranks, rank 0 runs the sequential
    stitch loop (computervisionimagestich2_b200/dist.py: stitch_sharded_device).  Timed with CUDA events bracketed by
    barrier + synchronize; time = max over ranks; value = job pixels / time."""
    import hashlib
    import torch
    from computervisionimagestich2_b200 import dist as pdist
    dev = torch.device("cuda", local_rank)
    n = len(imgs)
    mine = pdist.images_of_rank(n, world, rank)
    mpix = max_over_ranks_sum(dist, torch, dev, sum(imgs[i].shape[1] * imgs[i].shape[2] for i in mine) / 1e6)
    keep = {i: torch.from_numpy(imgs[i]).to(dev) for i in mine}            # inputs resident in HBM on their owner
    staged = {i: (keep[i].data_ptr(), imgs[i].shape[2], imgs[i].shape[1]) for i in mine}
    sync = torch.cuda.synchronize

    def max_over_ranks(x, op=None):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op or dist.ReduceOp.MAX)
        return float(t.item())

    def job(staged_inputs, timers=None, want_output=False):
        return pdist.stitch_sharded_device(ctx, imgs, dist, dev, staged=staged if staged_inputs else None,
                                           want_output=want_output, sync=sync, timers=timers)

    for _ in range(args.warmup):
        job(True)
    L.pano_b200_ktimer_enable(0)
    L.pano_b200_ktimer_reset()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases = {}
    total_ms = 0.0
    info = None
    for _ in range(args.steps):
        L.pano_b200_flush_l2(ctx.h)
        sync()
        dist.barrier()
        sync()
        e0.record()
        _, info = job(True, timers=phases)
        sync()
        dist.barrier()
        sync()
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
    launches = int(max_over_ranks(float(L.pano_b200_ktimer_launches()), dist.ReduceOp.SUM))
    clocks = sampler.stop() if sampler else None
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = mpix / (ms_per_step / 1e3)
    # ---- end to end: pinned host inputs on every rank (H2D inside), panorama to pinned host memory on rank 0 ----
    pin = {}
    for i in mine:
        b = imgs[i].nbytes
        p = L.pano_b200_alloc_pinned(C.c_size_t(b))
        C.memmove(p, imgs[i].ctypes.data, b)
        pin[i] = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (b,)).reshape(imgs[i].shape)
    pin_imgs = [pin.get(i, imgs[i]) for i in range(n)]
    e2e_ms, pano, e2e_steps = 0.0, None, max(1, min(args.steps, 5))
    for it in range(1 + e2e_steps):
        L.pano_b200_flush_l2(ctx.h)
        sync(); dist.barrier(); sync()
        t0 = time.perf_counter()
        pano, _info2 = pdist.stitch_sharded_device(ctx, pin_imgs, dist, dev, want_output=True, sync=sync)
        sync(); dist.barrier()
        if it > 0:
            e2e_ms += (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(e2e_ms)
    h2d = int(max_over_ranks(float(sum(imgs[i].nbytes for i in mine)), dist.ReduceOp.SUM))
    # ---- one instrumented step: rank 0's kernels (it holds 1/N of SIFT and matching and all of the canvas work) ----
    L.pano_b200_ktimer_reset()
    L.pano_b200_ktimer_enable(1)
    job(True)
    kernels = {}
    if rank == 0:
        buf = C.create_string_buffer(1 << 16)
        L.pano_b200_ktimer_report(buf, 1 << 16)
        kernels = json.loads(buf.value.decode())
    L.pano_b200_ktimer_enable(0)
    mstats = ctx.match_stats()
    if rank != 0:
        dist.destroy_process_group()
        return
    pano_hash = hashlib.sha256(pano.tobytes()).hexdigest()
    ph = {k: round(v / args.steps, 3) for k, v in phases.items()}
    serial = ph.get("stitch", 0.0) + ph.get("gather", 0.0)
    limiter = max(ph.items(), key=lambda kv: kv[1])[0] if ph else None
    top = max(kernels.items(), key=lambda kv: kv[1]["ms"]) if kernels else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": data,
        "config": {"workload": desc, "images": n, "input_mpixel": mpix, "output": list(info["size"]),
                   "l2": "flushed (256 MB memset) between timed iterations",
                   "parallelism": f"sharded x{world}: image i on rank i % {world}; NCCL all-gather of descriptor blocks (HBM to HBM); "
                                  "directed matching problems dealt to the ranks; projections and match lists to rank 0; rank 0 stitches",
                   "panorama_sha256": pano_hash, "bit_exact_vs_one_gpu": expected_synth_hash(args.workload, pano_hash),
                   "phases_ms_rank0": ph, "limiter": f"{limiter} (rank 0: sequential stitch loop + gather = {serial:.1f} ms of {ms_per_step:.1f} ms)",
                   "problems_per_rank": info["plan"], "nfeat": info["nfeat"]},
        "clocks": clocks,
        "e2e": {"value": mpix / (e2e_ms / e2e_steps / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(pano.nbytes),
                "ms_per_step": e2e_ms / e2e_steps},
        "gpu_launches": launches,
        "roofline": roofline_of(top, kernels, clocks) if top else None,
        "cpu_baseline": None,
        "kernels_ms_rank0": {k: round(v["ms"], 4) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])},
        "match_prefilter_rank0": mstats,
    }
    print(json.dumps(line))
    dist.destroy_process_group()


_IIR_NOTE = ("recursive Gaussian (CImg vanvliet), {axis} pass: per line a serial 3rd-order recurrence in double, dependent chain DMUL+3 DADD = 32 "
             "cycles/sample (measured); parallel only across lines. Algorithmic bytes = 8 B per plane sample per pass (one read + one "
             "write, SURVEY 8d 'x-IIR r+w, y-IIR r+w'); the kernel itself moves 16 B (the forward result goes through HBM before the "
             "backward sweep: a line does not fit on chip). {extra}")
ROOFLINE_NOTES = {
    "blend.iir_x": _IIR_NOTE.format(axis="x", extra="Lines = rows x planes: only 2.3 k x 7 lines of up to 18 k samples on the last edge, i.e. 3.4 warps per SM -- the pass is bound by the per-line latency (2 N x ~65 cycles), not by bandwidth"),
    "blend.iir_y": _IIR_NOTE.format(axis="y", extra="Lines = columns x planes (126 k lines on the last edge): ncu shows long_scoreboard + mio_throttle stalls, DRAM at 47 %"),
    "match.group_sym": "grouped lower bound of the uint8 SAD pre-filter, both directed problems of an image pair per pass: 8 VABSDIFF4.U8.ACC per (row of X, row of Y) on 32-byte group vectors + 2.5 packed 16-bit add / max / min / xor for the skip test (10.5 ALU-pipe instructions per row pair against 37 for the full SAD pass); VABSDIFF4 issues at half rate (measured 60.7 of 64 lanes/clk/SM, tools/ubench/vsad_rate.cu), so peak = 148 SM x 64 lanes x median SM clock; 'achieved' counts the 8 VABSDIFF4 only -- with the bookkeeping on the same pipe the pipe carries 1.31 x that",
    "sift.descr": "one warp per (keypoint, angle); gather from the L2-resident gradient map + double-precision geometry per patch sample; algorithmic bytes = sum (2W+1)^2 * 8 B patch reads + 512 B per descriptor (SURVEY 8d).  The kernel is not a streaming kernel: its patches are re-read from L2 (the gradient map of a 4K octave is 66 MB) and its time is instruction issue + FP64 latency of the per-sample geometry (~20 instructions per patch sample, two of them double divisions; ncu: profiles/r02b_ncu_full_descr_kernel_4k.txt), so the HBM fraction is low by construction",
    "blend.collapse": "expand + Laplacian + blend + add + clamp of one pyramid level, 9 planes up-sampled per pixel with CImg's double-precision linear interpolation (rounded to float after each axis): instruction-bound (DESIGN.md 4.5); algorithmic bytes = level planes read + output written",
    "blend.reduce": "CImg moving-average 2:1 reduce of the blurred level (x pass rounded to float, then y); algorithmic bytes = 4 B x planes x (source + destination samples)",
    "match.sad": "uint8 SAD pre-filter of the exact float-L1 matcher: 32 VABSDIFF4.U8.ACC per (query, database row) pair (one per 4 dimensions) + ~4 integer min/max for the running bounds; VABSDIFF4 issues at half rate (measured 60.7 of 64 lanes/clk/SM, tools/ubench/vsad_rate.cu), so peak = 148 SM x 64 lanes x median SM clock; 'achieved' counts the 32 VABSDIFF4 only",
    "match.sad_sym": "uint8 SAD pre-filter of the exact float-L1 matcher, BOTH directed problems of an image pair from one pass over the SAD matrix: 32 VABSDIFF4.U8.ACC per (row of X, row of Y) + ~5 packed 16-bit min/max/add for the two sets of running bounds + 1.5 CREDUX; VABSDIFF4 issues at half rate (measured 60.7 of 64 lanes/clk/SM, tools/ubench/vsad_rate.cu), so peak = 148 SM x 64 lanes x median SM clock; 'achieved' counts the 32 VABSDIFF4 only (the bookkeeping shares the same ALU pipe)",
    "match.l1": "exact float-L1 matcher: 2 FP32 instructions (FADD sub, FADD |.|-accumulate) per dimension, no FMA possible; peak = 148 SM x 128 lanes x median SM clock",
}


def roofline_of(top, kernels, clocks, workload=None):
    """roofline object of the dominant kernel of the instrumented pass.  HBM-bound kernels: algorithmic bytes (declared by
    the launcher) / CUDA-event time against MEASURED_PEAKS.json; the two matcher kernels are instruction-issue
    bound integer / FP32 CUDA-core kernels (no tensor-core form of an L1 distance exists): their figure is
    thread-instructions per second against the issue peak of the pipe they run on."""
    top_name, top_k = top
    if top_k["ms"] <= 0:
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    ksum = sum(k["ms"] for k in kernels.values())
    per_launch_ms = top_k["ms"] / top_k["launches"]
    common = {"kernel": top_name, "avg_launch_ms": per_launch_ms, "launches": top_k["launches"],
              "share_of_kernel_time": top_k["ms"] / ksum, "note": ROOFLINE_NOTES.get(top_name), "traffic": None}
    if top_name in ("match.l1", "match.sad", "match.sad_sym", "match.group_sym"):
        ach = top_k["bytes"] / (top_k["ms"] * 1e-3) / 1e12
        clk = (clocks or {}).get("sm_mhz") or 1500.0
        lanes = 128 if top_name == "match.l1" else 64
        peak = 148 * lanes * clk * 1e6 / 1e12
        r = {"bound": "fp32-issue" if top_name == "match.l1" else "int-alu-issue", "achieved": ach, "peak": peak,
             "unit": "Tinstr/s", "frac": ach / peak, "peak_source": f"148 SM x {lanes} lanes/clk x {clk:.0f} MHz (median SM clock during the timed region)"}
    else:
        ach = top_k["bytes"] / (top_k["ms"] * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "peak_source": peak_src,
             "bytes_per_launch": top_k["bytes"] / top_k["launches"]}
    r.update(common)
    if workload:
        r.update(ncu_traffic(top_name, workload))
    return r


def max_over_ranks_sum(dist, torch, dev, x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    return float(t.item())


def _render_pair(args):
    index, w, h = args
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import synth_scene
    path = os.path.join(tempfile.gettempdir(), f"pano_b200_synth_v3_pair{index}_{w}x{h}.npy")
    try:
        a = np.load(path)
        if a.shape == (2, 3, h, w):
            return a
    except Exception:
        pass
    a = np.stack(synth_scene.pair(index, w, h))
    tmp = path + f".{os.getpid()}.tmp.npy"
    np.save(tmp, a)
    os.replace(tmp, path)
    return a


def run_b200_pairs(args, ctx, L, dist, rank, local_rank, world):
    """BASELINE.json configs[4]: 1024 independent 1920x1080 pairs -- SIFT of both images, getImgPair both ways, RANSAC of
    the adjacent directions -- pair p on rank p % N, no data-path collective ("replicas only"), 8 pairs per
    pano_b200_pairs call.  The total is fixed, so N > 1 is strong scaling.  `value`: inputs staged in HBM
    (pano_b200_pairs_staged); `e2e`: pinned host buffers in (H2D inside), 160-byte records out."""
    import torch
    npairs, w, h, distinct, chunk = args.pairs, 1920, 1080, min(args.pairs, args.distinct_pairs), 8
    dev = torch.device("cuda", local_rank)
    mine = list(range(rank, npairs, world))
    need = sorted({p % distinct for p in mine})
    procs = max(1, min(len(need), (os.cpu_count() or 2) // max(1, world), 8))
    if procs > 1:
        import subprocess
        kids = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--render-pairs", f"{w},{h}",
                                  ",".join(map(str, need[k::procs]))]) for k in range(procs)]
        for k in kids:
            if k.wait() != 0:
                raise RuntimeError("synthetic pair renderer failed")
    got = [_render_pair((i, w, h)) for i in need]   # cached by the children above
    scenes = dict(zip(need, got))
    staged = {i: torch.from_numpy(scenes[i]).to(dev) for i in need}            # [2][3][h][w] per distinct pair, in HBM
    pinned = {}
    for i in need:
        b = scenes[i].nbytes
        ptr = L.pano_b200_alloc_pinned(C.c_size_t(b))
        C.memmove(ptr, scenes[i].ctypes.data, b)
        pinned[i] = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (b,)).reshape(scenes[i].shape)
    img_bytes = 3 * w * h

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_staged():
        out = []
        for k in range(0, len(mine), chunk):
            sub = mine[k:k + chunk]
            ptrs = [staged[p % distinct].data_ptr() + j * img_bytes for p in sub for j in (0, 1)]
            out.append(ctx.pairs_staged(ptrs, [(w, h)] * len(ptrs)))
        return np.concatenate(out) if out else None

    def run_host():
        out = []
        for k in range(0, len(mine), chunk):
            sub = mine[k:k + chunk]
            out.append(ctx.pairs([(pinned[p % distinct][0], pinned[p % distinct][1]) for p in sub]))
        return np.concatenate(out) if out else None

    def max_over_ranks(x, op=None):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op or dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(min(args.warmup, 2)):
        run_staged()
    L.pano_b200_ktimer_enable(0)
    L.pano_b200_ktimer_reset()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms, rec = 0.0, None
    for _ in range(args.steps):
        L.pano_b200_flush_l2(ctx.h)
        sync_all()
        e0.record()
        rec = run_staged()
        sync_all()
        e1.record()
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
    launches = int(max_over_ranks(float(L.pano_b200_ktimer_launches()), None if dist is None else dist.ReduceOp.SUM))
    clocks = sampler.stop() if sampler else None
    ms_per_step = max_over_ranks(total_ms) / args.steps
    mpix = 2.0 * npairs * w * h / 1e6
    e2e_ms = 0.0
    for it in range(2):
        sync_all()
        t0 = time.perf_counter()
        rec_h = run_host()
        sync_all()
        if it > 0:
            e2e_ms += (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(e2e_ms)
    same = rec is None or rec_h is None or (rec["nfeat"].tobytes() == rec_h["nfeat"].tobytes() and rec["H"].tobytes() == rec_h["H"].tobytes())
    L.pano_b200_ktimer_reset()
    L.pano_b200_ktimer_enable(1)
    ctx.match_stats(reset=True)
    run_staged()
    buf = C.create_string_buffer(1 << 16)
    L.pano_b200_ktimer_report(buf, 1 << 16)
    L.pano_b200_ktimer_enable(0)
    kernels = json.loads(buf.value.decode())
    mstats = ctx.match_stats()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    top = max(kernels.items(), key=lambda kv: kv[1]["ms"]) if kernels else None
    line = {
        "metric": METRIC, "value": mpix / (ms_per_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": min(args.warmup, 2), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batched {npairs} synthetic 1920x1080 image pairs: SIFT + getImgPair both ways + RANSAC, BASELINE.json configs[4]",
                   "pairs": npairs, "distinct_pairs": distinct, "pairs_per_call": chunk, "input_mpixel": mpix,
                   "pairs_per_s": npairs / (ms_per_step / 1e3),
                   "parallelism": f"pair p on rank p % {world}; replicas only, no data-path collective" if world > 1 else "1 GPU",
                   "l2": "flushed (256 MB memset) before every timed step; a step's inputs (%.1f GB per rank) exceed L2" % (len(mine) * 2 * img_bytes / 1e9),
                   "note": f"{distinct} distinct scenes (tools/synth_scene.pair), pair p uses scene p % {distinct}: rendering 1024 scenes would take the CPU longer than the benchmark; every pair is processed in full",
                   "features_per_image_mean": float(rec["nfeat"].mean()), "matches_per_direction_mean": float(rec["nmatch"].mean()),
                   "directions_fitted_rank0": int(rec["has_h"].sum()), "staged_equals_host_path": bool(same)},
        "clocks": clocks,
        "e2e": {"value": mpix / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(2 * npairs * img_bytes),
                "d2h_bytes_per_step": int(160 * npairs), "ms_per_step": e2e_ms, "pairs_per_s": npairs / (e2e_ms / 1e3)},
        "gpu_launches": launches,
        "roofline": roofline_of(top, kernels, clocks) if top else None,
        "cpu_baseline": None,
        "kernels_ms_rank0": {k: round(v["ms"], 4) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])},
        "match_prefilter_rank0": mstats,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def expected_synth_hash(workload, got):
    """tests/golden/anchors.json["synth"][workload] = SHA-256 of the panorama of a synthetic workload as produced on ONE
    GPU (bit-exact against the reference at the sizes the tests can afford; tests/test_gpu_synth.py); the sharded job
    must reproduce it at every N."""
    try:
        a = json.load(open(os.path.join(ROOT, "tests", "golden", "anchors.json")))
        want = a.get("synth", {}).get(workload)
        return None if want is None else want == got
    except Exception:
        return None


def ncu_traffic(kernel, workload):
    """DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of this workload (profiles/
    r01_ncu_traffic.json, written by tools/summarize_ncu.py): the capture is of ONE launch -- the largest -- so its own
    algorithmic bytes are given beside it; `traffic` stays None when no capture of this kernel / workload exists."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))[workload][kernel]
            return {"traffic": t["dram_bytes"], "traffic_launch": t["launch"], "traffic_launch_algorithmic_bytes": t["algorithmic_bytes"],
                    "traffic_source": t["source"]}
        except Exception:
            continue
    return {}


def sift_summary(kernels, stages, mpix):
    """north_star: 'SIFT Mpix/s (HBM GB/s)'.  Mpixel/s = input pixels / wall time of the feature stage of the last timed
    step (projection + SIFT + table for all images, concurrent lanes; `project` and `table` in stages_ms are per-lane sums
    that lie inside it); GB/s = algorithmic bytes / CUDA-event time of the
    scale-space kernels (blur, DoG-on-the-fly detector, gradient) in the serial instrumented pass."""
    ss = [k for n, k in kernels.items() if n in ("sift.blur_v", "sift.blur_h", "sift.detect", "sift.gradient")]
    ms = sum(k["ms"] for k in ss)
    by = sum(k["bytes"] for k in ss)
    feat_ms = stages.get("sift", 0)   # wall time of Stitcher::add_images: projection + SIFT + table of all images, lanes in parallel
    return {"mpixel_per_s": round(mpix / (feat_ms * 1e-3), 1) if feat_ms > 0 else None, "feature_stage_ms": round(feat_ms, 3),
            "scale_space_GBps": round(by / (ms * 1e-3) / 1e9, 1) if ms > 0 else None, "scale_space_ms": round(ms, 4),
            "all_sift_kernels_ms": round(sum(k["ms"] for n, k in kernels.items() if n.startswith("sift.")), 4)}


def bench_ex6(ctx, reps=3):
    """src/ex6/dataset2 (18 x 600x800 -> 6282x883) with the ex6 profile through the public call (host buffers in/out)."""
    import hashlib
    from computervisionimagestich2_b200 import bmpio
    d = os.path.join(ROOT, "oracle", "_ref", "data", "ex6_dataset2")
    a = json.load(open(os.path.join(ROOT, "tests", "golden", "anchors.json")))["ex6"]
    imgs = [bmpio.load_bmp(os.path.join(d, f"{i + 1}.bmp")) for i in range(a["dataset2"]["n"])]
    mp = megapixels(imgs)
    ctx.set_profile("ex6", a["ransac_seed"])
    try:
        best, out = 1e9, None
        for _ in range(1 + reps):
            t0 = time.perf_counter()
            out, _info = ctx.stitch(imgs)
            best = min(best, time.perf_counter() - t0)
    finally:
        ctx.set_profile("root", 666666)
    return {"workload": "src/ex6/dataset2: 18 x 600x800 chain panorama, ex6 profile, RANSAC seed pinned", "input_mpixel": mp,
            "output": [int(out.shape[2]), int(out.shape[1])], "ms": round(best * 1e3, 2), "mpixel_per_s": round(mp / best, 1),
            "bit_exact_vs_reference": hashlib.sha256(out.tobytes()).hexdigest() == a["dataset2"]["sha256"],
            "reference_cpu_seconds_one_core": a["dataset2"]["cpu_seconds"]}


def bench_pairs(ctx, npairs=8, w=1920, h=1080, reps=3):
    """npairs synthetic 1080p pairs (two 50 %-overlap views of a seeded scene each) through Context.pairs
    (pano_b200_pairs): SIFT of both images, both directed matches, RANSAC of the adjacent directions; host buffers in,
    160-byte records out.  Best of `reps` after one warm-up call."""
    pairs = [tuple(synth_scene_views(2, w, h, seed=20181126 + p)) for p in range(npairs)]
    best, rec = 1e9, None
    for _ in range(1 + reps):
        t0 = time.perf_counter()
        rec = ctx.pairs(pairs)
        best = min(best, time.perf_counter() - t0)
    return {"workload": f"{npairs} synthetic {w}x{h} image pairs: SIFT + both directed matches + RANSAC, one call",
            "ms": round(best * 1e3, 2), "pairs_per_s": round(npairs / best, 1),
            "mpixel_per_s": round(2 * npairs * w * h / 1e6 / best, 1),
            "features_per_image_mean": round(float(rec["nfeat"].mean()), 1),
            "matches_per_direction_mean": round(float(rec["nmatch"].mean()), 1),
            "directions_fitted": int(rec["has_h"].sum())}


def main():
    if len(sys.argv) == 4 and sys.argv[1] == "--render-views":   # renderer child of synth_scene_views
        n, w, h, seed = map(int, sys.argv[2].split(","))
        _render_views((n, w, h, seed, [int(x) for x in sys.argv[3].split(",")]))
        return
    if len(sys.argv) == 4 and sys.argv[1] == "--render-pairs":   # renderer child of the pairs workload
        w, h = map(int, sys.argv[2].split(","))
        for i in map(int, sys.argv[3].split(",")):
            _render_pair((i, w, h))
        return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="synth4k", choices=["input", "input2", "synth4k", "synth4k_r1", "synth1080", "synth8k", "pairs1024"])
    ap.add_argument("--pairs", type=int, default=1024, help="pairs1024: number of pairs")
    ap.add_argument("--distinct-pairs", type=int, default=64, help="pairs1024: number of distinct scenes rendered")
    ap.add_argument("--ref-procs", type=int, default=64)
    ap.add_argument("--match-mode", default="prefilter", choices=["prefilter", "full", "prefilter_onedir"])
    ap.add_argument("--mode", default="auto", choices=["auto", "sharded", "replicas"],
                    help="N > 1: 'sharded' = one job over all ranks (strong scaling; default for the synthetic workloads), "
                         "'replicas' = one job per rank (weak scaling; default for the bundled sets)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-match-u8", action="store_true")
    ap.add_argument("--no-ex6", action="store_true")
    ap.add_argument("--no-pairs", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
