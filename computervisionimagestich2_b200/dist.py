"""Sharded panorama job: one process per GPU, `torch.distributed` for the plumbing.

What shards, and how (reference: ImageProcess.cpp:12-23, 117-137):
  * readFile -- load / project / SIFT / feature table of image i is independent of every other image: image i is
    processed by rank i % world.  ONE exchange follows: an all-gather of the per-image blocks (feature count,
    descriptor table [n][128] f32, keypoints [n] x 32 B, projected image), ragged sizes padded to the maximum.
  * all-pairs matching -- directed problem (i, j) is independent of every other: the problems of a wave are dealt
    round-robin to the ranks, the match lists (nfeat[j] int32 each) are all-gathered.  The two waves reproduce the set of
    problems the reference evaluates: (i, j) for i < j always, (i, j) for i > j only when (j, i) found fewer than 20 matches.
  * the BFS stitching loop is sequential (canvas k depends on canvas k-1): rank 0 runs it (pano_b200_stitch_features)
    with every match list preset; the few tree-edge directions the discovery skipped are evaluated there.
No other collective is used.  The engine is injected so that the host logic can be tested on CPU with the gloo backend
(tests/test_cpu_dist.py drives it with the oracle); the product engine is `Context` (CUDA, no fallback).
"""
from __future__ import annotations

import numpy as np

from . import KEY_DTYPE

THRESHOLD = 20  # ImageProcess.h:18


def images_of_rank(n: int, world: int, rank: int):
    return list(range(rank, n, world))


def wave1(n: int):
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


def wave2(n: int, counts: dict):
    return [(i, j) for i in range(n) for j in range(i) if counts[(j, i)] < THRESHOLD]


def chain_wave(n: int):
    """The src/ex6 profile matches only neighbours of the fixed chain 0-1-...-(n-1), both directions
    (src/ex6/ImageProcess.cpp:150-157, 195-196): one wave, no adjacency discovery."""
    return [p for i in range(n - 1) for p in ((i, i + 1), (i + 1, i))]


def _all_gather_ragged(dist, arr: np.ndarray, device):
    """all-gather of one byte blob per rank (ragged): sizes first, then blobs padded to the maximum."""
    import torch
    world = dist.get_world_size()
    raw = np.frombuffer(np.ascontiguousarray(arr).tobytes(), np.uint8)
    size = torch.tensor([raw.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, size)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    buf = torch.zeros(cap, dtype=torch.uint8, device=device)
    if raw.size:
        buf[: raw.size] = torch.from_numpy(raw.copy()).to(device)
    outs = [torch.empty(cap, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, buf)
    return [o[:s].cpu().numpy() for o, s in zip(outs, sizes)]


def _pack_images(items):
    """[(index, proj, descr, keys)] -> one byte blob"""
    parts = [np.array([len(items)], np.int64).tobytes()]
    for i, proj, descr, keys in items:
        hdr = np.array([i, proj.shape[1], proj.shape[2], len(keys)], np.int64)
        parts += [hdr.tobytes(), np.ascontiguousarray(proj, np.uint8).tobytes(),
                  np.ascontiguousarray(descr, np.float32).tobytes(), np.ascontiguousarray(keys, KEY_DTYPE).tobytes()]
    return np.frombuffer(b"".join(parts), np.uint8)


def _unpack_images(blob):
    b = blob.tobytes()
    k = int(np.frombuffer(b, np.int64, 1, 0)[0])
    off = 8
    out = []
    for _ in range(k):
        i, h, w, n = (int(x) for x in np.frombuffer(b, np.int64, 4, off))
        off += 32
        proj = np.frombuffer(b, np.uint8, 3 * h * w, off).reshape(3, h, w).copy()
        off += 3 * h * w
        descr = np.frombuffer(b, np.float32, n * 128, off).reshape(n, 128).copy()
        off += n * 512
        keys = np.frombuffer(b, KEY_DTYPE, n, off).copy()
        off += n * KEY_DTYPE.itemsize
        out.append((i, proj, descr, keys))
    return out


def _pack_matches(items):
    parts = [np.array([len(items)], np.int64).tobytes()]
    for (i, j), idx in items:
        parts += [np.array([i, j, len(idx)], np.int64).tobytes(), np.ascontiguousarray(idx, np.int32).tobytes()]
    return np.frombuffer(b"".join(parts), np.uint8)


def _unpack_matches(blob):
    b = blob.tobytes()
    k = int(np.frombuffer(b, np.int64, 1, 0)[0])
    off = 8
    out = {}
    for _ in range(k):
        i, j, n = (int(x) for x in np.frombuffer(b, np.int64, 3, off))
        off += 24
        out[(i, j)] = np.frombuffer(b, np.int32, n, off).copy()
        off += 4 * n
    return out


def stitch_sharded(engine, images, dist=None, device="cpu", profile="root"):
    """images: the full list of planar uint8 [3][H][W] inputs (every rank holds it, or at least its own share at the
    right positions).  Returns (panorama, info) on rank 0 and (None, info) elsewhere; info carries the exchanged
    feature counts and match counts so that callers / tests can check them on every rank.  profile: "root" (all-pairs
    discovery, two waves) or "ex6" (chain neighbours, one wave); the engine must have been switched to the same profile
    (Context.set_profile) so that rank 0 stitches accordingly."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    n = len(images)
    # 1. features of my images
    mine = []
    for i in images_of_rank(n, world, rank):
        proj, descr, keys = engine.extract(images[i])
        mine.append((i, proj, descr, keys))
    # 2. the one exchange of the job: all-gather of the per-image blocks
    blobs = _all_gather_ragged(dist, _pack_images(mine), device) if dist is not None else [_pack_images(mine)]
    table = {}
    for blob in blobs:
        for i, proj, descr, keys in _unpack_images(blob):
            table[i] = (proj, descr, keys)
    assert sorted(table) == list(range(n))
    # 3. all-pairs matching, problems dealt round-robin, match lists all-gathered after each wave
    midx = {}

    def run_wave(pairs):
        local = [((i, j), engine.match_idx(table[i][1], table[j][1])) for (i, j) in pairs[rank::world]]
        got = _all_gather_ragged(dist, _pack_matches(local), device) if dist is not None else [_pack_matches(local)]
        for blob in got:
            midx.update(_unpack_matches(blob))

    if profile == "ex6":
        run_wave(chain_wave(n))
    else:
        run_wave(wave1(n))
        counts = {k: int((v >= 0).sum()) for k, v in midx.items()}
        run_wave(wave2(n, counts))
    counts = {k: int((v >= 0).sum()) for k, v in midx.items()}
    info = dict(nfeat=[len(table[i][2]) for i in range(n)], match_counts=counts, world=world)
    # 4. the sequential part on rank 0
    if rank != 0:
        return None, info
    pano, sinfo = engine.stitch_features([table[i][0] for i in range(n)], [(table[i][1], table[i][2]) for i in range(n)], midx)
    info.update(sinfo)
    return pano, info


# ----------------------------------------------------------------------------------------------------------------------
# The same job with a DEVICE-RESIDENT exchange: descriptor blocks go from the SIFT engine's HBM buffers through NCCL
# into the matcher's tables without touching the host; projected images travel to rank 0 only; match lists come back
# to rank 0 only.  (The host-staged stitch_sharded above stays as the engine-agnostic form the CPU tests drive.)
# ----------------------------------------------------------------------------------------------------------------------
def all_directed(n: int):
    """Every directed problem of an n-image job: a superset of what the reference evaluates (wave1 + wave2 + tree-edge
    directions, ImageProcess.cpp:117-137, 177-178); the stitcher only consults the lists the reference would compute."""
    return [(i, j) for i in range(n) for j in range(n) if i != j]


def deal_problems(problems, nfeat, world):
    """Longest-processing-time-first assignment of directed problems to ranks; cost = NA * NB.  Both directions of an
    image pair go to the same rank (they share the two tables).  Deterministic: every rank computes the same plan."""
    groups = {}
    for (i, j) in problems:
        groups.setdefault((min(i, j), max(i, j)), []).append((i, j))
    items = sorted(groups.items(), key=lambda kv: (-len(kv[1]) * nfeat[kv[0][0]] * nfeat[kv[0][1]], kv[0]))
    load = [0] * world
    plan = [[] for _ in range(world)]
    for key, probs in items:
        r = min(range(world), key=lambda q: (load[q], q))
        plan[r] += sorted(probs)
        load[r] += len(probs) * nfeat[key[0]] * nfeat[key[1]] + 1
    return plan


def _pad_gather(dist, t, cap, dst=None):
    """all-gather (dst None) or gather-to-dst of one 1-D tensor per rank, padded to `cap` elements"""
    import torch
    world = dist.get_world_size()
    buf = torch.zeros(cap, dtype=t.dtype, device=t.device)
    buf[: t.numel()] = t
    if dst is None:
        out = torch.empty(world * cap, dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, buf)
        return out.view(world, cap)
    outs = [torch.empty(cap, dtype=t.dtype, device=t.device) for _ in range(world)] if dist.get_rank() == dst else None
    dist.gather(buf, outs, dst=dst)
    return outs


_PLANE_GROUPS = {}


def plane_group(dist, world):
    """NCCL / gloo sub-group of the ranks that carry the canvas planes (0, 1, 2); created once, by every rank"""
    key = (id(dist), world)
    if key not in _PLANE_GROUPS:
        _PLANE_GROUPS[key] = dist.new_group(ranks=[0, 1, 2])
    return _PLANE_GROUPS[key]


def stitch_sharded_device(engine, images, dist, device, profile="root", staged=None, want_output=True, sync=None, timers=None,
                          planes="auto"):
    """images: list of n planar uint8 arrays (only the entries this rank owns are read), or staged = {i: (device_ptr, w, h)}
    for inputs already resident in HBM.  Returns (panorama | None, info).  `sync()` must wait for the device (torch
    collectives run on torch's stream, the engine on its own; each hand-over is a host synchronisation).  `timers`
    (dict) receives the phase wall times of this rank in ms.  planes: "auto" = with 3 or more ranks (root profile) the
    canvas stages of the sequential stitch loop are split by colour plane over ranks 0, 1, 2 -- the planes of warp /
    shift / blend are independent except for the 16-byte seam statistics of plane 0, broadcast once per edge, and the
    final equalisation, for which rank 0 collects the other two planes; "off" = rank 0 carries all three."""
    import time
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    n = len(images)
    sync = sync or (lambda: None)
    tm = timers if timers is not None else {}
    t0 = time.perf_counter()
    use_planes = planes != "off" and world >= 3 and profile != "ex6"
    canvas_ranks = [0, 1, 2] if use_planes else [0]
    pg = plane_group(dist, world) if use_planes else None

    def lap(name):
        nonlocal t0
        sync()
        t1 = time.perf_counter()
        tm[name] = tm.get(name, 0.0) + (t1 - t0) * 1e3
        t0 = t1

    mine = images_of_rank(n, world, rank)
    engine.shard_begin(n)
    if staged is not None:
        engine.shard_extract(None, mine, staged_ptrs=[staged[i][0] for i in mine], sizes=[staged[i][1:] for i in mine])
    else:
        engine.shard_extract([images[i] for i in mine], mine)
    lap("extract")
    # ---- meta: (w, h, nfeat) of every image, one tiny all-reduce -------------------------------------------------
    meta = torch.zeros((n, 3), dtype=torch.int64, device=device)
    for i in mine:
        w, h = (staged[i][1], staged[i][2]) if staged is not None else (images[i].shape[2], images[i].shape[1])
        meta[i] = torch.tensor([w, h, engine.nfeatures(i)], dtype=torch.int64)
    dist.all_reduce(meta)
    meta = meta.cpu().numpy()
    W, H, NF = [int(x) for x in meta[:, 0]], [int(x) for x in meta[:, 1]], [int(x) for x in meta[:, 2]]
    owner = [i % world for i in range(n)]
    # ---- descriptors + keypoints: all-gather; projections: to the canvas ranks only ---------------------------------
    rows_of = [sum(NF[i] for i in images_of_rank(n, world, r)) for r in range(world)]
    cap_rows = max(max(rows_of), 1)
    send_d = torch.empty((cap_rows, 128), dtype=torch.float32, device=device)
    keys_mine = np.zeros(rows_of[rank], KEY_DTYPE)
    px_of = [sum(3 * W[i] * H[i] for i in images_of_rank(n, world, r)) for r in range(world)]
    cap_px = max(max(px_of), 1)
    send_p = torch.empty(cap_px, dtype=torch.uint8, device=device) if world > 1 else None
    ro = po = 0
    for i in mine:
        engine.shard_export(i, send_d[ro:ro + NF[i]] if NF[i] else None, keys_mine[ro:ro + NF[i]] if NF[i] else None,
                            send_p[po:po + 3 * W[i] * H[i]] if send_p is not None else None)
        ro += NF[i]
        po += 3 * W[i] * H[i]
    all_d = _pad_gather(dist, send_d.view(-1), cap_rows * 128).view(world, cap_rows, 128)
    kbytes = torch.from_numpy(keys_mine.view(np.uint8).copy()).to(device)
    all_k = _pad_gather(dist, kbytes, cap_rows * KEY_DTYPE.itemsize).cpu().numpy()
    all_p = None
    if world > 1:
        all_p = _pad_gather(dist, send_p, cap_px) if use_planes else _pad_gather(dist, send_p, cap_px, dst=0)
    sync()
    ro = [0] * world
    po = [0] * world
    for i in range(n):
        r = owner[i]
        if r != rank:
            k = np.frombuffer(all_k[r].tobytes(), KEY_DTYPE, NF[i], ro[r] * KEY_DTYPE.itemsize)
            proj = all_p[r][po[r]:po[r] + 3 * W[i] * H[i]] if rank in canvas_ranks else None
            engine.shard_import(i, W[i], H[i], NF[i], all_d[r, ro[r]:ro[r] + NF[i]], k, proj)
        ro[r] += NF[i]
        po[r] += 3 * W[i] * H[i]
    lap("exchange")
    # ---- matching: directed problems dealt to the ranks, lists to the canvas ranks --------------------------------------
    problems = chain_wave(n) if profile == "ex6" else all_directed(n)
    plan = deal_problems(problems, NF, world)
    len_of = [sum(NF[j] for (_, j) in plan[r]) for r in range(world)]
    cap_idx = max(max(len_of), 1)
    out_idx = torch.empty(cap_idx, dtype=torch.int32, device=device)
    engine.shard_match([p[0] for p in plan[rank]], [p[1] for p in plan[rank]], out_idx)
    lap("match")
    if world == 1:
        got = [out_idx]
    elif use_planes:
        got = _pad_gather(dist, out_idx, cap_idx)
    else:
        got = _pad_gather(dist, out_idx, cap_idx, dst=0)
    info = dict(nfeat=NF, world=world, plan=[len(p) for p in plan], canvas_ranks=canvas_ranks)
    if rank not in canvas_ranks:
        lap("gather")
        dist.barrier()
        return None, info
    counts = {}
    for r in range(world):
        lists = got[r].cpu().numpy()
        off = 0
        for (i, j) in plan[r]:
            idx = lists[off:off + NF[j]]
            off += NF[j]
            engine.shard_preset(i, j, idx)
            counts[(i, j)] = int((idx >= 0).sum())
    lap("gather")
    info["match_counts"] = counts
    if not use_planes:
        pano, sinfo = engine.shard_stitch(want_output=want_output)
        lap("stitch")
        info.update(sinfo)
        dist.barrier()
        return pano, info

    def exchange(vals, is_source):
        t = torch.tensor(vals, dtype=torch.int32, device=device) if is_source else torch.zeros(4, dtype=torch.int32, device=device)
        dist.broadcast(t, src=0, group=pg)
        return t.cpu().tolist()

    sinfo = engine.shard_stitch_planes(rank, 1, exchange)
    lap("stitch")
    cw, ch = sinfo["size"]
    plane = torch.empty(cw * ch, dtype=torch.uint8, device=device)
    pano = None
    if rank == 0:
        for c in (1, 2):
            dist.recv(plane, src=c)
            sync()
            engine.shard_plane_import(c, plane)
        pano, tinfo = engine.shard_tail(want_output=want_output)
        sinfo.update(tinfo)
    else:
        engine.shard_plane_export(0, plane)
        dist.send(plane, dst=0)
    lap("collect")
    info.update(sinfo)
    dist.barrier()
    return pano, info


# ----------------------------------------------------------------------------------------------------------------------
# Batched image pairs (BASELINE.json configs[4], SURVEY.md 8e last row): "replicas only"
# ----------------------------------------------------------------------------------------------------------------------
PAIR_RECORD = np.dtype([("pair", "<i8"), ("nfeat", "<i4", 2), ("nmatch", "<i4", 2), ("has_h", "<i4", 2), ("H", "<f8", (2, 8))])


def pairs_of_rank(npairs: int, world: int, rank: int):
    return list(range(rank, npairs, world))


def pair_job(engine, img_a, img_b):
    """One independent pair: readFile body of both images (ImageProcess.cpp:12-23), getImgPair in both directions
    (:117-137, 273-351) and RANSAC on every direction that reaches the adjacency threshold (:128, 395-436).  Direction
    0 is getImgPair(a, b): for each feature of b its match in a, RANSAC fits a -> b; direction 1 the reverse.
    Returns one PAIR_RECORD (pair index left at -1)."""
    _, da, ka = engine.extract(img_a)
    _, db, kb = engine.extract(img_b)
    ka = np.ascontiguousarray(ka, KEY_DTYPE)
    kb = np.ascontiguousarray(kb, KEY_DTYPE)
    rec = np.zeros((), PAIR_RECORD)
    rec["pair"] = -1
    rec["nfeat"] = (len(ka), len(kb))
    for d, (dsrc, ksrc, ddst, kdst) in enumerate(((da, ka, db, kb), (db, kb, da, ka))):
        if len(ksrc) < 2 or len(kdst) == 0:  # the 2-NN query needs two database rows (ImageProcess.cpp:327)
            continue
        idx = np.asarray(engine.match_idx(dsrc, ddst))
        sel = idx >= 0
        rec["nmatch"][d] = int(sel.sum())
        if rec["nmatch"][d] >= THRESHOLD:
            rec["H"][d] = engine.ransac(ksrc[idx[sel]].copy(), kdst[sel].copy())
            rec["has_h"][d] = 1
    return rec


def pairs_by_calls(engine, pairs):
    """The same table as engine.pairs, composed from the single-stage calls (pair_job per pair).  The product engine
    (Context.pairs -> pano_b200_pairs) batches the images, the matching problems and the RANSAC problems of a chunk into
    one call instead; this composition is what the tests cross-check it with."""
    rec = np.zeros(len(pairs), PAIR_RECORD)
    for k, (a, b) in enumerate(pairs):
        rec[k] = pair_job(engine, a, b)
        rec[k]["pair"] = k
    return rec


def pairs_batch(engine, pairs, dist=None, device="cpu", gather=True, chunk=8):
    """pairs: list of (img_a, img_b) planar uint8 [3][H][W]; every rank holds the list (or at least its own share at
    the right positions).  Pair p runs on rank p % world, `chunk` pairs per engine.pairs call (2 * chunk images on the
    engine's lanes, 2 * chunk directed matching problems in one launch); there is no data-path collective.  With
    gather=True the small result records (PAIR_RECORD, 160 B per pair) are all-gathered afterwards so that every rank
    returns the full table in pair order; with gather=False each rank returns only its own records."""
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    ids = pairs_of_rank(len(pairs), world, rank)
    mine = np.zeros(len(ids), PAIR_RECORD)
    for k in range(0, len(ids), chunk):
        sub = ids[k:k + chunk]
        rec = engine.pairs([pairs[p] for p in sub])
        rec["pair"] = sub
        mine[k:k + len(sub)] = rec
    if dist is None or not gather:
        return mine
    blobs = _all_gather_ragged(dist, mine.view(np.uint8), device)
    table = np.concatenate([np.frombuffer(b.tobytes(), PAIR_RECORD) for b in blobs])
    table = table[np.argsort(table["pair"], kind="stable")]
    assert np.array_equal(table["pair"], np.arange(len(pairs)))
    return table
