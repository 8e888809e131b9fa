"""Batched image pairs (BASELINE.json configs[4]): SIFT of both images + both directed matches + RANSAC per pair,
pair p on rank p % world, no data-path collective ("replicas only", SURVEY.md 8e last row).  Prints one JSON line on
rank 0; time = max over ranks of the host wall clock around the rank's share (device synchronised on both sides, every
call of the pair job goes through the C ABI with host buffers, so this is an end-to-end figure).

    python tools/run_pairs.py [npairs=64] [width=1920] [height=1080] [reps=2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/run_pairs.py 64

The pairs are two 50 %-overlap views of a seeded synthetic scene (bench.synth_scene_views, seed 20181126 + p); each
rank renders only its own pairs.
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

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
h = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pairs = [None] * npairs
for p in pdist.pairs_of_rank(npairs, world, rank):
    pairs[p] = tuple(bench.synth_scene_views(2, w, h, seed=20181126 + p))
ctx = pano.Context(local)
times = []
for r in range(reps + 1):   # the first pass is the warm-up (allocations, pinned staging)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mine = pdist.pairs_batch(ctx, pairs, dist=dist if world > 1 else None, device=f"cuda:{local}", gather=False)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if r > 0:
        times.append(float(t.item()))
table = mine
if world > 1:   # result report (160 B per pair), outside the timed region
    table = pdist._all_gather_ragged(dist, mine.view(np.uint8), f"cuda:{local}")
    table = np.concatenate([np.frombuffer(b.tobytes(), pdist.PAIR_RECORD) for b in table])
    table = table[np.argsort(table["pair"], kind="stable")]
if rank == 0:
    t = float(np.median(times))
    print(json.dumps({"workload": f"{npairs} synthetic {w}x{h} image pairs: SIFT + both directed matches + RANSAC",
                      "world": world, "s_per_batch": round(t, 4), "pairs_per_s": round(npairs / t, 2),
                      "mpix_per_s": round(2 * npairs * w * h / 1e6 / t, 1),
                      "features_per_image_mean": round(float(table["nfeat"].mean()), 1),
                      "matches_per_direction_mean": round(float(table["nmatch"].mean()), 1),
                      "directions_fitted": int(table["has_h"].sum()),
                      "table_sha256": hashlib.sha256(table.tobytes()).hexdigest()}))
ctx.close()
if world > 1:
    dist.destroy_process_group()
