"""Multi-process host logic on CPU (gloo, world_size 2): tile partition + the single film gather.

The per-rank renderer here is the oracle twin standing in for a GPU (test infrastructure); what is
under test is pbrt-v1_b200/tiles.py -- the code bench.py runs with NCCL on the GPU box.  Property:
the frame assembled from independently rendered tiles is bit-identical to a full-frame render.
"""
import importlib.util
import os
import socket
import sys

import numpy as np
import pytest

from parity_util import ROOT, gen_scene


def load_tiles():
    spec = importlib.util.spec_from_file_location("pbrt_b200_tiles", os.path.join(ROOT, "pbrt-v1_b200", "tiles.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_tile_partition_covers_window_exactly():
    tiles = load_tiles()
    for (x0, y0, w, h, t) in [(0, 0, 1024, 1024, 128), (3, 5, 100, 37, 32), (0, 0, 1, 1, 128), (10, 0, 257, 129, 128)]:
        ts = tiles.make_tiles(x0, y0, w, h, t)
        cover = np.zeros((h, w), np.int32)
        for (tx, ty, tw, th) in ts:
            assert tw >= 1 and th >= 1
            cover[ty - y0:ty - y0 + th, tx - x0:tx - x0 + tw] += 1
        assert (cover == 1).all()
        for world in (1, 2, 3, 8):
            got = sorted(sum((tiles.tiles_of_rank(ts, r, world) for r in range(world)), []))
            assert got == sorted(ts)


def _worker(rank, world, port, scene_path, out_path):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity_util import load_ffi
    ffi = load_ffi()
    tiles = load_tiles()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = ffi.Scene(scene_path)

    def render_tiles(mine):
        return [ffi.twin_render(scene, ffi.Window(*t))[0] for t in mine]

    frame = tiles.render_distributed(render_tiles, (scene.x0, scene.y0, scene.width, scene.height), rank, world,
                                     dist=dist, tile=16)
    if rank == 0:
        np.save(out_path, frame)
    else:
        assert frame is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("filt", ["box", "mitchell"])
def test_gloo_world2_tiles_equal_full_frame(ffi, scene_dir, filt):
    import torch.multiprocessing as mp
    sp = gen_scene(os.path.join(scene_dir, "dist_%s.pbrt" % filt), "c3", "--res", "40", "36", "--spp", "2", "2",
                   "--filter", filt, "--filename", os.path.join(scene_dir, "dist.exr"))
    out = os.path.join(scene_dir, "dist_%s.npy" % filt)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, sp, out), nprocs=2, join=True)
    frame = np.load(out)
    scene = ffi.Scene(sp)
    full = ffi.twin_render(scene)[0]
    assert frame.shape == full.shape
    assert np.array_equal(frame.view(np.uint32), full.view(np.uint32))
