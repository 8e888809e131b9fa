"""GPU tier: the plane-sharded canvas stages (pano_b200_shard_stitch_planes / _plane_export / _plane_import / _tail).
Three stitchers on one GPU stand for ranks 0, 1, 2 of a sharded job: each holds the same images and carries ONE colour
plane through warp / shift / blend; the 16-byte seam statistics of plane 0 are recorded on the first run and replayed to
the other two (what the NCCL broadcast does in dist.stitch_sharded_device); plane 0's stitcher then collects the other
planes and runs the equalisation tail.  The panorama must equal the reference's bit for bit (ImageProcess.cpp:159-271)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_planes(images):
    import torch
    import computervisionimagestich2_b200 as pano
    n = len(images)
    ctxs = [pano.Context(0) for _ in range(3)]
    try:
        for c in ctxs:
            c.set_profile("root", 666666)
            c.shard_begin(n)
            c.shard_extract(images, list(range(n)))
        recorded = []

        def record(vals, is_source):
            assert is_source
            recorded.append(list(vals))
            return vals

        info0 = ctxs[0].shard_stitch_planes(0, 1, record)
        assert len(recorded) == n - 1
        for k in (1, 2):
            it = iter(recorded)

            def replay(vals, is_source):
                assert not is_source and vals is None
                return next(it)

            info = ctxs[k].shard_stitch_planes(k, 1, replay)
            assert info["size"] == info0["size"] and info["log"] == info0["log"]
        w, h = info0["size"]
        for k in (1, 2):
            plane = torch.empty(w * h, dtype=torch.uint8, device="cuda:0")
            ctxs[k].shard_plane_export(0, plane)
            torch.cuda.synchronize()
            ctxs[0].shard_plane_import(k, plane)
        out, _ = ctxs[0].shard_tail()
        return out, info0["log"], recorded
    finally:
        for c in ctxs:
            c.close()


def test_plane_sharded_stitch_equals_reference(ref, input_sets):
    for name in ("Input", "Input2"):
        out, log, stats = _run_planes(input_sets[name])
        want, winfo = ref.stitch_mem(input_sets[name])
        assert log == winfo["log"]
        assert out.shape == want.shape and np.array_equal(out, want), name
        assert all(s[1] > 0 and s[3] > 0 for s in stats)


def test_three_planes_in_one_stitcher_and_errors(ctx, ref, input_sets):
    """count = 3 is the whole job minus the tail; the ex6 profile cannot be plane-sharded (its statistics use all planes)"""
    import computervisionimagestich2_b200 as pano
    imgs = input_sets["Input"]
    ctx.set_profile("root", 666666)
    ctx.shard_begin(4)
    ctx.shard_extract(imgs, [0, 1, 2, 3])
    ctx.shard_stitch_planes(0, 3, None)
    out, _ = ctx.shard_tail()
    want, _ = ref.stitch_mem(imgs)
    assert np.array_equal(out, want)
    ctx.set_profile("ex6", 666666)
    ctx.shard_begin(4)
    ctx.shard_extract(imgs, [0, 1, 2, 3])
    with pytest.raises(pano.PanoError):
        ctx.shard_stitch_planes(1, 1, lambda v, s: [0, 1, 0, 1])
    ctx.set_profile("root", 666666)
