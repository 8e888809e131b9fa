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


# ---- batched pairs (BASELINE.json configs[4]): replicas only ----------------------------------------------------------
PAIR_LIST = [(0, 1), (1, 2), (2, 3), (0, 3), (3, 2)]


class OraclePairEngine(OracleEngine):
    def ransac(self, src, dst):
        return self.ref.ransac(src, dst)

    def pairs(self, pairs):
        from computervisionimagestich2_b200 import dist as pdist
        return pdist.pairs_by_calls(self, pairs)


def _pairs_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    from computervisionimagestich2_b200 import dist as pdist
    from oracle import ref_api
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    imgs = [ref_api.load_bmp(os.path.join(ref_api.REF_DATA, "Input", f"{i}.bmp")) for i in range(1, 5)]
    table = pdist.pairs_batch(OraclePairEngine(), [(imgs[a], imgs[b]) for a, b in PAIR_LIST], dist=dist, device="cpu", chunk=2)
    own = pdist.pairs_batch(OraclePairEngine(), [(imgs[a], imgs[b]) for a, b in PAIR_LIST[:3]], dist=dist, device="cpu",
                            gather=False)
    np.save(os.path.join(out_dir, f"pairs_rank{rank}.npy"), table)
    np.save(os.path.join(out_dir, f"own_rank{rank}.npy"), own)
    dist.barrier()
    dist.destroy_process_group()


def test_pairs_plan():
    from computervisionimagestich2_b200 import dist as pdist
    assert pdist.pairs_of_rank(5, 2, 0) == [0, 2, 4] and pdist.pairs_of_rank(5, 2, 1) == [1, 3]
    assert pdist.pairs_of_rank(1, 8, 3) == [] and pdist.PAIR_RECORD.itemsize == 160


def test_pairs_batch_world2_gloo(ref, tmp_path):
    """Pair p runs on rank p % 2; both ranks end with the same table, equal to a single-process run and to the anchors;
    the fitted coefficients are the reference's RANSAC on the reference's match list, bit for bit."""
    import socket
    import torch.multiprocessing as mp
    from computervisionimagestich2_b200 import dist as pdist
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_pairs_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0 = np.load(tmp_path / "pairs_rank0.npy")
    t1 = np.load(tmp_path / "pairs_rank1.npy")
    assert t0.tobytes() == t1.tobytes() and len(t0) == len(PAIR_LIST)
    assert [int(p) for p in np.load(tmp_path / "own_rank0.npy")["pair"]] == [0, 2]
    assert [int(p) for p in np.load(tmp_path / "own_rank1.npy")["pair"]] == [1]
    imgs = [ref.load_bmp(os.path.join(ref.REF_DATA, "Input", f"{i}.bmp")) for i in range(1, 5)]
    single = pdist.pairs_batch(OraclePairEngine(), [(imgs[a], imgs[b]) for a, b in PAIR_LIST])
    assert single.tobytes() == t0.tobytes()
    anchors = json.load(open(os.path.join(HERE, "golden", "anchors.json")))["Input"]
    feats = [ref.sift_features(ref.gray(ref.project(im))) for im in imgs]
    for rec, (a, b) in zip(t0, PAIR_LIST):
        assert list(rec["nfeat"]) == [anchors["nfeat"][a], anchors["nfeat"][b]]
        assert list(rec["nmatch"]) == [anchors["match_counts"][f"{a},{b}"], anchors["match_counts"][f"{b},{a}"]]
        for d, (i, j) in enumerate(((a, b), (b, a))):
            assert bool(rec["has_h"][d]) == (rec["nmatch"][d] >= 20)
            if rec["has_h"][d]:
                src, dst = ref.match(feats[i][0], feats[i][1], feats[j][0], feats[j][1])
                assert ref.ransac(src, dst).tobytes() == rec["H"][d].tobytes()
            else:
                assert not rec["H"][d].any()
    assert t0["has_h"].sum() >= 6   # the chain neighbours of Input are adjacent in both directions


# ---- the device-resident exchange (dist.stitch_sharded_device): same host logic, tensors on the CPU over gloo ----------
class OracleShardEngine:
    """shard_* interface of Context backed by the compiled reference; "device" tensors are CPU torch tensors."""

    def __init__(self):
        from oracle import ref_api
        self.ref = ref_api

    def shard_begin(self, n):
        self.im, self.presets, self.n = {}, {}, n

    def shard_extract(self, imgs, slots, staged_ptrs=None, sizes=None):
        for img, s in zip(imgs, slots):
            p = self.ref.project(img)
            d, k = self.ref.sift_features(self.ref.gray(p))
            self.im[s] = dict(w=img.shape[2], h=img.shape[1], proj=p, d=d, k=k)

    def nfeatures(self, i):
        return len(self.im[i]["k"])

    def shard_export(self, i, descr_t, keys, proj_t):
        if descr_t is not None:
            descr_t.numpy()[:] = self.im[i]["d"]
            keys[:] = self.im[i]["k"]
        if proj_t is not None:
            proj_t.numpy()[:] = self.im[i]["proj"].ravel()

    def shard_import(self, i, w, h, n, descr_t, keys, proj_t):
        self.im[i] = dict(w=w, h=h, d=descr_t.numpy()[:n].copy().reshape(n, 128), k=np.array(keys).copy(),
                          proj=None if proj_t is None else proj_t.numpy().copy().reshape(3, h, w))

    def shard_match(self, I, J, out_t):
        sys.path.insert(0, HERE)
        import emul_api
        off = 0
        for i, j in zip(I, J):
            idx = emul_api.match_idx(self.im[i]["d"], self.im[j]["d"])
            out_t.numpy()[off:off + len(idx)] = idx
            off += len(idx)

    def shard_preset(self, i, j, idx):
        self.presets[(i, j)] = np.array(idx).copy()

    def shard_stitch(self, want_output=True):
        return None, dict(log="")


def _shard_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import hashlib
    import torch.distributed as dist
    from computervisionimagestich2_b200 import dist as pdist
    from oracle import ref_api
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    imgs = [ref_api.load_bmp(os.path.join(ref_api.REF_DATA, "Input", f"{i}.bmp")) for i in range(1, 5)]
    eng = OracleShardEngine()
    tm = {}
    _, info = pdist.stitch_sharded_device(eng, imgs, dist, "cpu", timers=tm)
    rec = {"nfeat": info["nfeat"], "plan": info["plan"], "timers": sorted(tm),
           "have_proj": [eng.im[i]["proj"] is not None for i in range(4)],
           "proj_sha": [hashlib.sha256(eng.im[i]["proj"].tobytes()).hexdigest() if eng.im[i]["proj"] is not None else None for i in range(4)],
           "descr_sha": [hashlib.sha256(eng.im[i]["d"].tobytes()).hexdigest() for i in range(4)],
           "keys_sha": [hashlib.sha256(eng.im[i]["k"].tobytes()).hexdigest() for i in range(4)],
           "presets": {f"{i},{j}": hashlib.sha256(v.tobytes()).hexdigest() for (i, j), v in eng.presets.items()},
           "match_counts": {f"{i},{j}": c for (i, j), c in info.get("match_counts", {}).items()}}
    with open(os.path.join(out_dir, f"shard{rank}.json"), "w") as f:
        json.dump(rec, f)
    dist.destroy_process_group()


def test_deal_problems_plan():
    from computervisionimagestich2_b200 import dist as pdist
    probs = pdist.all_directed(4)
    assert len(probs) == 12 and (0, 0) not in probs
    plan = pdist.deal_problems(probs, [100, 200, 300, 400], 3)
    flat = sorted(p for r in plan for p in r)
    assert flat == sorted(probs)                                  # every problem exactly once
    for r in plan:                                                # both directions of a pair on the same rank
        assert all((j, i) in r for (i, j) in r)
    cost = [sum(100 * (i + 1) * 100 * (j + 1) for (i, j) in r) for r in plan]
    assert max(cost) <= 1.5 * (sum(cost) / 3)                     # and roughly balanced
    assert pdist.deal_problems(probs, [0, 0, 0, 0], 2) == pdist.deal_problems(probs, [0, 0, 0, 0], 2)   # deterministic
    assert [len(r) for r in pdist.deal_problems(pdist.chain_wave(3), [5, 5, 5], 8)].count(2) == 2


def test_sharded_device_exchange_world2_gloo(ref, tmp_path):
    """Descriptors and keypoints reach every rank, projections only rank 0, every directed match list reaches rank 0 --
    all byte-identical to a single process computing them from the images."""
    import hashlib
    import socket
    import torch.multiprocessing as mp
    sys.path.insert(0, HERE)
    import emul_api
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = json.load(open(tmp_path / "shard0.json"))
    r1 = json.load(open(tmp_path / "shard1.json"))
    anchors = json.load(open(os.path.join(HERE, "golden", "anchors.json")))["Input"]
    imgs = [ref.load_bmp(os.path.join(ref.REF_DATA, "Input", f"{i}.bmp")) for i in range(1, 5)]
    projs = [ref.project(im) for im in imgs]
    feats = [ref.sift_features(ref.gray(p)) for p in projs]
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert r0["nfeat"] == r1["nfeat"] == anchors["nfeat"]
    assert r0["descr_sha"] == r1["descr_sha"] == [sha(f[0]) for f in feats]
    assert r0["keys_sha"] == r1["keys_sha"] == [sha(f[1]) for f in feats]
    assert r0["have_proj"] == [True] * 4 and r0["proj_sha"] == [sha(p) for p in projs]
    assert r1["have_proj"] == [False, True, False, True]          # rank 1 keeps its own, receives none
    assert sum(r0["plan"]) == 12 and r0["plan"] == r1["plan"]
    assert r0["match_counts"] == anchors["match_counts"] and r1["match_counts"] == {}
    for i in range(4):
        for j in range(4):
            if i != j:
                assert r0["presets"][f"{i},{j}"] == sha(emul_api.match_idx(feats[i][0], feats[j][0]).astype(np.int32))
    assert r1["presets"] == {} and {"extract", "exchange", "match", "gather"} <= set(r0["timers"])


# ---- plane-sharded canvas stages: world_size 3 over gloo, fake engine ------------------------------------------------
class OraclePlaneEngine(OracleShardEngine):
    """adds the plane-stitch interface: records the lock-step seam exchange and the plane collection"""

    def shard_stitch_planes(self, first, count, exchange):
        self.first = first
        self.seen = []
        for edge in range(self.n - 1):                    # one exchange per stitched edge, in lock step on all three ranks
            vals = exchange([100 + edge, 7, 50 + edge, 3] if first == 0 else None, first == 0)
            self.seen.append([int(v) for v in vals])
        return dict(log="", size=(5, 4))

    def shard_plane_export(self, k, out_t):
        out_t.numpy()[:] = 10 + self.first

    def shard_plane_import(self, channel, in_t):
        self.imported = getattr(self, "imported", {})
        self.imported[channel] = int(in_t.numpy()[0])

    def shard_tail(self, want_output=True):
        return None, dict(tail=True)


def _plane_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    from computervisionimagestich2_b200 import dist as pdist
    from oracle import ref_api
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    imgs = [ref_api.load_bmp(os.path.join(ref_api.REF_DATA, "Input", f"{i}.bmp")) for i in range(1, 5)]
    eng = OraclePlaneEngine()
    _, info = pdist.stitch_sharded_device(eng, imgs, dist, "cpu")
    rec = {"canvas_ranks": info["canvas_ranks"], "seen": getattr(eng, "seen", None), "imported": getattr(eng, "imported", None),
           "npresets": len(eng.presets), "have_proj": [eng.im[i]["proj"] is not None for i in range(4)], "tail": info.get("tail")}
    with open(os.path.join(out_dir, f"plane{rank}.json"), "w") as f:
        json.dump(rec, f)
    dist.destroy_process_group()


def test_plane_sharded_stitch_world3_gloo(ref, tmp_path):
    """With three ranks the canvas stages split by colour plane: every canvas rank gets all projections and all 12 match
    lists, the seam statistics of plane 0 reach ranks 1 and 2 edge by edge, rank 0 collects planes 1 and 2 and runs the tail."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_plane_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    r = [json.load(open(tmp_path / f"plane{k}.json")) for k in range(3)]
    want = [[100 + e, 7, 50 + e, 3] for e in range(3)]
    for k in range(3):
        assert r[k]["canvas_ranks"] == [0, 1, 2] and r[k]["npresets"] == 12 and r[k]["have_proj"] == [True] * 4
        assert r[k]["seen"] == want
    assert r[0]["imported"] == {"1": 11, "2": 12} and r[0]["tail"] is True
    assert r[1]["imported"] is None and r[2]["imported"] is None
