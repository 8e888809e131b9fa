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
