"""Sharded panorama job over N GPUs (torchrun, NCCL):  images -> ranks, all-gather of feature blocks, pair-sharded
matching, rank 0 stitches.  Prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_sharded.py [input|input2|synth4k|synth8k|ex6_dataset2|ex6_dataset3] [reps]

ex6_dataset2 / ex6_dataset3 run the reference's src/ex6 sets (18 / 11 images) with the ex6 profile; the SHA-256 is
checked against tests/golden/anchors.json.
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import computervisionimagestich2_b200 as pano  # noqa: E402
from computervisionimagestich2_b200 import dist as pdist  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "input2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
profile, golden = "root", None
if name.startswith("ex6_dataset"):
    from computervisionimagestich2_b200 import bmpio
    k = int(name[-1])
    a = json.load(open(os.path.join(ROOT, "tests", "golden", "anchors.json")))["ex6"]
    d = os.path.join(ROOT, "oracle", "_ref", "data", name)
    imgs = [bmpio.load_bmp(os.path.join(d, f"{i + 1}.bmp")) for i in range(a[f"dataset{k}"]["n"])]
    desc, profile, golden = f"src/ex6/dataset{k} ({len(imgs)} images), ex6 profile", "ex6", a[f"dataset{k}"]["sha256"]
else:
    imgs, desc, _ = bench.load_workload(name)
ctx = pano.Context(local)
if profile == "ex6":
    ctx.set_profile("ex6", 666666)
times = []
for r in range(reps + 1):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, info = pdist.stitch_sharded(ctx, imgs, dist=dist if world > 1 else None, device=f"cuda:{local}", profile=profile)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if r > 0:
        times.append(time.perf_counter() - t0)
if rank == 0:
    print(json.dumps({"workload": desc, "world": world, "ms_per_job": round(1e3 * float(np.median(times)), 2),
                      "mpix_per_s": round(bench.megapixels(imgs) / float(np.median(times)), 1), "nfeat": info["nfeat"],
                      "log": info["log"].split(), "panorama": list(out.shape), "sha256": hashlib.sha256(out.tobytes()).hexdigest(),
                      "matches_golden": None if golden is None else hashlib.sha256(out.tobytes()).hexdigest() == golden}))
if world > 1:
    dist.destroy_process_group()
