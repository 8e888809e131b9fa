"""CPU tier: the host logic of the sharded job (computervisionimagestich2_b200/dist.py) with world_size = 2 over gloo.
The engine is the oracle (oracle/_ref) instead of the CUDA library: what is checked here is the sharding plan, the ragged
all-gathers and the wave logic -- every rank must end up with the same feature tables and match lists as a single
process, and those must equal the golden anchors."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class OracleEngine:
    """extract / match_idx backed by the compiled reference; stitch_features only records what it was given."""

    def __init__(self):
        from oracle import ref_api
        self.ref = ref_api

    def extract(self, img):
        p = self.ref.project(img)
        d, k = self.ref.sift_features(self.ref.gray(p))
        return p, d, k

    def match_idx(self, dA, dB):
        sys.path.insert(0, HERE)
        import emul_api
        return emul_api.match_idx(dA, dB)

    def stitch_features(self, projs, feats, midx):
        return None, dict(npairs=len(midx), shapes=[p.shape for p in projs])


def _worker(rank, world, port, out_dir, profile="root"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    from computervisionimagestich2_b200 import dist as pdist
    from oracle import ref_api
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    imgs = [ref_api.load_bmp(os.path.join(ref_api.REF_DATA, "Input", f"{i}.bmp")) for i in range(1, 5)]
    _, info = pdist.stitch_sharded(OracleEngine(), imgs, dist=dist, device="cpu", profile=profile)
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump({"nfeat": info["nfeat"], "match_counts": {f"{i},{j}": c for (i, j), c in info["match_counts"].items()},
                   "npairs": info.get("npairs")}, f)
    dist.barrier()
    dist.destroy_process_group()


def test_plan_helpers():
    from computervisionimagestich2_b200 import dist as pdist
    assert pdist.images_of_rank(5, 2, 0) == [0, 2, 4] and pdist.images_of_rank(5, 2, 1) == [1, 3]
    assert pdist.wave1(3) == [(0, 1), (0, 2), (1, 2)]
    counts = {(0, 1): 87, (0, 2): 3, (1, 2): 54}
    assert pdist.wave2(3, counts) == [(2, 0)]
    assert pdist.chain_wave(4) == [(0, 1), (1, 0), (1, 2), (2, 1), (2, 3), (3, 2)] and pdist.chain_wave(1) == []
    items = [((2, 0), np.array([1, -1, 5], np.int32)), ((1, 0), np.zeros(0, np.int32))]
    back = pdist._unpack_matches(pdist._pack_matches(items))
    assert np.array_equal(back[(2, 0)], items[0][1]) and len(back[(1, 0)]) == 0


def test_sharded_job_world2_gloo(ref, tmp_path):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    anchors = json.load(open(os.path.join(HERE, "golden", "anchors.json")))["Input"]
    r0 = json.load(open(tmp_path / "rank0.json"))
    r1 = json.load(open(tmp_path / "rank1.json"))
    assert r0["nfeat"] == r1["nfeat"] == anchors["nfeat"]
    assert r0["match_counts"] == r1["match_counts"]
    # wave 1 = all i < j; wave 2 = the i > j whose mirror found fewer than 20 matches
    expect = {k: v for k, v in anchors["match_counts"].items()
              if int(k.split(",")[0]) < int(k.split(",")[1])
              or anchors["match_counts"][",".join(reversed(k.split(",")))] < 20}
    assert r0["match_counts"] == expect
    assert r0["npairs"] == len(expect) and r1["npairs"] is None


def test_sharded_job_world2_gloo_ex6_chain(ref, tmp_path):
    """src/ex6 profile: only the neighbours of the fixed chain are matched, both directions, in one wave."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path), "ex6"), nprocs=2, join=True)
    anchors = json.load(open(os.path.join(HERE, "golden", "anchors.json")))["Input"]
    r0 = json.load(open(tmp_path / "rank0.json"))
    r1 = json.load(open(tmp_path / "rank1.json"))
    assert r0["nfeat"] == r1["nfeat"] == anchors["nfeat"]
    expect = {k: anchors["match_counts"][k] for k in ("0,1", "1,0", "1,2", "2,1", "2,3", "3,2")}
    assert r0["match_counts"] == r1["match_counts"] == expect
    assert r0["npairs"] == 6 and r1["npairs"] is None
