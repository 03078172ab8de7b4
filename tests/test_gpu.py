"""GPU parity tests (-m gpu): everything goes through the C ABI of libpbrt_b200.so.

Checkers, in order of authority:
  1. tests/golden/*.npz -- outputs of the compiled reference renderer (committed fixtures);
  2. oracle/_ref run live on this box when present (it travels with the snapshot);
  3. the oracle twin (oracle/twin.cpp), which tests/test_oracle.py pins bit-for-bit to (1).

Tolerances (fp32 film handed to WriteRGBAImage, before half rounding):
  * integer / index work -- sample positions, camera rays, closest-hit primitive ids, t and
    barycentrics, occlusion bits, ray counters, tile assembly: BIT EXACT;
  * radiance: the GPU evaluates sinf/cosf/atan2f/acosf in double precision rounded once, the
    reference calls glibc's float versions, so isolated samples differ in the last bit and a few
    paths flip.  Stated tolerance: |delta| <= max(2^-10 |ref|, 1e-3) for >= 97 % of pixels on these
    2..9 spp fixtures (one flipped sample moves a 4-spp pixel by up to L_light/4), image mean within
    0.5 %; and bit-identical to the twin built with the same double-rounded functions.
"""
import glob
import os

import numpy as np
import pytest

from parity_util import ROOT, gen_scene, image_error_stats, read_sample_dump

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
FRAC_WITHIN_TOL = 0.97
MEAN_REL = 5e-3


def load_case(ffi, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return ffi.Scene(os.path.join(GOLDEN, name + ".pbrt")), g["rgba"], g["samples"], g["stats"]


def gpu_render(ffi, scene, windows=None, keep_samples=True, count=True):
    ds = ffi.DeviceScene(scene)
    ffi.cuda().pbrt_b200_tunable(1, 1 if count else 0)
    ffi.cuda().pbrt_b200_sample_radiance(ds.h, 0, 1 if keep_samples else 0, None)
    st = ds.render(windows)
    ffi.cuda().pbrt_b200_tunable(1, 0)
    return ds, st


@pytest.mark.parametrize("name", CASES)
def test_render_vs_reference_golden(ffi, name):
    scene, rgba_ref, samples_ref, stats_ref = load_case(ffi, name)
    ds, st = gpu_render(ffi, scene)
    rgba = ds.download(0)
    n = samples_ref.shape[0]
    assert st["camera_rays"] == n == stats_ref[0]
    ls = ds.sample_radiance(0, n)
    s = image_error_stats(rgba, rgba_ref)
    assert s["frac_within_tol"] >= FRAC_WITHIN_TOL, s
    assert s["mean_rel"] <= MEAN_REL, s
    d = np.abs(ls[:, :3] - samples_ref[:, 2:5]).max(axis=1)
    assert (d <= 1e-4).mean() >= 0.995, float((d <= 1e-4).mean())
    assert np.array_equal(ls[:, 3], samples_ref[:, 5])            # alpha
    for got, want in ((st["shadow_rays"], stats_ref[1]), (st["points_shaded"], stats_ref[2])):
        assert abs(got - want) <= max(2, 0.01 * want), (got, want)
    assert st["nan_samples"] == 0 and st["infinite_samples"] == 0
    ds.close()


@pytest.mark.parametrize("name", CASES)
def test_render_bit_identical_to_twin(ffi, name):
    """Same arithmetic (double-rounded transcendentals) on CPU and GPU: every sample, every pixel, every counter."""
    scene, _, _, _ = load_case(ffi, name)
    ds, st = gpu_render(ffi, scene)
    rgba, raw = ds.download(0, want_raw=True)
    rgba_t, raw_t, samples_t, st_t = ffi.twin_render(scene, want_samples=True)
    ls = ds.sample_radiance(0, samples_t.shape[0])
    assert np.array_equal(ls[:, :3].view(np.uint32), samples_t[:, 2:5].view(np.uint32))
    assert np.array_equal(raw.view(np.uint32), raw_t.view(np.uint32))
    assert np.array_equal(rgba.view(np.uint32), rgba_t.view(np.uint32))
    for k in ("camera_rays", "intersect_rays", "shadow_rays", "points_shaded", "interior_visits", "leaf_visits",
              "prim_refs_read", "prim_tests", "nan_samples", "negative_samples", "infinite_samples"):
        assert st[k] == st_t[k], (k, st[k], st_t[k])
    assert st["kernel_launches"] > 0
    ds.close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("PBRT_FUZZ_SCENES", "32"))))
def test_fuzz_scenes_bit_identical_to_twin(ffi, tmp_path, seed):
    """Random scenes mixing every supported plugin (tests/fuzz_scenes.py; tests/test_oracle.py pins the twin
    to the compiled reference on the same generator): CUDA path == twin, samples, film and counters."""
    import random
    from fuzz_scenes import scene as fuzz_scene
    sp = str(tmp_path / ("fuzz_%d.pbrt" % seed))
    open(sp, "w").write(fuzz_scene(random.Random(500009 + seed), seed))
    scene = ffi.Scene(sp)
    ds, st = gpu_render(ffi, scene)
    rgba, raw = ds.download(0, want_raw=True)
    rgba_t, raw_t, samples_t, st_t = ffi.twin_render(scene, want_samples=True)
    ls = ds.sample_radiance(0, samples_t.shape[0])
    assert np.array_equal(ls[:, :3].view(np.uint32), samples_t[:, 2:5].view(np.uint32))
    assert np.array_equal(raw.view(np.uint32), raw_t.view(np.uint32))
    assert np.array_equal(rgba.view(np.uint32), rgba_t.view(np.uint32))
    for k in ("camera_rays", "intersect_rays", "shadow_rays", "points_shaded", "interior_visits", "leaf_visits",
              "prim_refs_read", "prim_tests", "nan_samples", "negative_samples", "infinite_samples"):
        assert st[k] == st_t[k], (k, st[k], st_t[k])
    ds.close()


def test_camera_rays_and_traversal_bit_exact(ffi, scene_dir):
    sp = gen_scene(os.path.join(scene_dir, "trav.pbrt"), "c4", "--res", "48", "48", "--spp", "2", "2", "--nphi", "96",
                   "--ntheta", "48", "--lens", "5", "700", "--filename", os.path.join(scene_dir, "trav.exr"))
    scene = ffi.Scene(sp)
    ds = ffi.DeviceScene(scene)
    rays_t, xy_t = ffi.twin_camera_rays(scene)
    rays_g, xy_g = ds.camera_rays(ffi.twin_sample_count(scene, scene.full_window()))
    assert np.array_equal(rays_t.view(np.uint32), rays_g.view(np.uint32))
    assert np.array_equal(xy_t.view(np.uint32), xy_g.view(np.uint32))
    rng = np.random.default_rng(11)
    extra = np.zeros((20000, 8), np.float32)
    extra[:, :3] = rng.uniform(5, 550, (20000, 3))
    d = rng.normal(size=(20000, 3))
    extra[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    extra[:5000, 3:6] *= rng.uniform(10, 400, (5000, 1)).astype(np.float32)   # unnormalised shadow-style segments
    extra[:, 6] = 1e-3
    extra[:, 7] = np.inf
    extra[:5000, 7] = 1 - 1e-3
    # axis-parallel directions exercise the 1/0 = inf slab and split-plane arithmetic
    extra[5000:5300, 3:6] = 0
    extra[5000:5100, 3] = 1
    extra[5100:5200, 4] = -1
    extra[5200:5300, 5] = 1
    allrays = np.concatenate([rays_t, extra])
    pt, tt, b1t, b2t = ffi.twin_intersect(scene, allrays)
    pg, tg, b1g, b2g = ds.intersect(allrays)
    assert np.array_equal(pt, pg)
    assert np.array_equal(tt.view(np.uint32), tg.view(np.uint32))
    assert np.array_equal(b1t.view(np.uint32), b1g.view(np.uint32))
    assert np.array_equal(b2t.view(np.uint32), b2g.view(np.uint32))
    assert np.array_equal(ffi.twin_intersect_p(scene, allrays), ds.intersect_p(allrays))
    assert (pg >= 0).mean() > 0.5
    # empty input
    e = np.zeros((0, 8), np.float32)
    assert ds.intersect(e)[0].shape == (0,)
    ds.close()


@pytest.mark.parametrize("variant", ["plain", "textured", "textured_whitted"])
def test_tiles_are_independent_and_wave_size_invariant(ffi, scene_dir, variant):
    """A window rendered alone == the same pixels of the full frame, bit for bit (Mitchell filter: the
    tile re-traces its apron), and the result does not depend on how samples are split into waves.  The
    textured variants carry per-slot camera-ray differentials and per-hit evaluated materials (path) or
    per-frame child differentials (whitted) through the same waves."""
    extra = []
    if variant != "plain":
        extra += ["--textured", os.path.join(GOLDEN, "tex13x7.exr"), "--mirror"]
    if variant == "textured_whitted":
        extra += ["--integrator", "whitted"]
    sp = gen_scene(os.path.join(scene_dir, "tiles_%s.pbrt" % variant), "c3", "--res", "48", "40", "--spp", "2", "2", "--filter",
                   "mitchell", "--filename", os.path.join(scene_dir, "tiles.exr"), *extra)
    scene = ffi.Scene(sp)
    ds = ffi.DeviceScene(scene)
    ds.render()
    full = ds.download(0)
    wins = [ffi.Window(0, 0, 20, 40), ffi.Window(20, 0, 28, 17), ffi.Window(20, 17, 28, 23)]
    ds.render(wins)
    for i, w in enumerate(wins):
        tile = ds.download(i)
        assert np.array_equal(tile.view(np.uint32), full[w.y0:w.y0 + w.height, w.x0:w.x0 + w.width].view(np.uint32))
    prev = ffi.cuda().pbrt_b200_tunable(0, 1000)        # tiny ragged waves
    try:
        ds2 = ffi.DeviceScene(scene)
        ds2.render()
        assert np.array_equal(ds2.download(0).view(np.uint32), full.view(np.uint32))
        ds2.close()
    finally:
        ffi.cuda().pbrt_b200_tunable(0, prev)
    ds.close()


@pytest.mark.parametrize("name", ["c3", "c4_blob", "c3_spherelight"])
def test_stream_modes_and_tree_order_change_no_bit(ffi, name):
    """How many waves of a frame are in flight (tunable 10), how a wave's kernels are spread over streams (tunable 5: one stream; shadow + probe traversal beside the next
    closest-hit launch; probes alone beside the shadow rays; three streams) and the device-side node order of the
    tree (tunable 9: treelets) are scheduling / layout choices: same image, same samples, same ray counts."""
    scene, _, _, _ = load_case(ffi, name)
    lib = ffi.cuda()
    prev5, prev9, prev10 = lib.pbrt_b200_tunable(5, -1), lib.pbrt_b200_tunable(9, -1), lib.pbrt_b200_tunable(10, -1)
    ref = None
    try:
        # (lanes: 1 = one wave at a time, 3 = two waves in flight on two lanes whatever the frame's size)
        for mode, treelets, lanes in ((0, 0, 1), (1, 0, 1), (2, 0, 1), (3, 0, 1), (3, 1, 1), (0, 1, 1), (3, 0, 3), (0, 0, 3)):
            lib.pbrt_b200_tunable(5, mode)
            lib.pbrt_b200_tunable(9, treelets)
            lib.pbrt_b200_tunable(10, lanes)
            ds, st = gpu_render(ffi, scene, count=False)   # (the counting pass keeps the wave on fewer streams)
            rgba, raw = ds.download(0, want_raw=True)
            ls = ds.sample_radiance(0, st["camera_rays"])
            got = (rgba.copy(), raw.copy(), ls.copy(),
                   tuple(st[k] for k in ("camera_rays", "intersect_rays", "shadow_rays", "points_shaded")))
            ds.close()
            if ref is None:
                ref = got
                continue
            for a, b in zip(got[:3], ref[:3]):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (mode, treelets, lanes)
            assert got[3] == ref[3], (mode, treelets, lanes)
    finally:
        lib.pbrt_b200_tunable(5, prev5)
        lib.pbrt_b200_tunable(9, prev9)
        lib.pbrt_b200_tunable(10, prev10)


def test_bestcandidate_windows_equal_reference_crop_renders(ffi, scene_dir):
    """bestcandidate sampler, several windows in one call (one wave spans them): every window equals the
    twin's (== the reference's) crop render of that window -- the sampler's table placements start at the
    window's own sample extent, exactly as in a reference crop render -- also with tiny ragged waves."""
    sp = gen_scene(os.path.join(scene_dir, "bcwin.pbrt"), "c3", "--res", "40", "36", "--spp", "3", "1", "--bc", "--filter",
                   "mitchell", "--filename", os.path.join(scene_dir, "bcwin.exr"))
    scene = ffi.Scene(sp)
    wins = [ffi.Window(0, 0, 17, 36), ffi.Window(17, 0, 23, 20), ffi.Window(17, 20, 23, 16)]
    for wave in (-1, 5000):
        prev = ffi.cuda().pbrt_b200_tunable(0, wave)
        try:
            ds = ffi.DeviceScene(scene)
            ds.render(wins)
            for i, w in enumerate(wins):
                tile, raw = ds.download(i, want_raw=True)
                rt, rawt, _, _ = ffi.twin_render(scene, w)
                assert np.array_equal(raw.view(np.uint32), rawt.view(np.uint32)), (wave, i)
                assert np.array_equal(tile.view(np.uint32), rt.view(np.uint32)), (wave, i)
            ds.close()
        finally:
            if wave >= 0:
                ffi.cuda().pbrt_b200_tunable(0, prev)


def test_render_host_end_to_end_and_exr(ffi, scene_dir):
    """pbrt_b200_render_host (host buffers both sides) and the full front end: parse -> GPU -> EXR."""
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    exr = os.path.join(scene_dir, "e2e.exr")
    sp = gen_scene(os.path.join(scene_dir, "e2e.pbrt"), "c2", "--res", "40", "40", "--spp", "2", "2", "--filename", exr)
    scene = ffi.Scene(sp)
    outs, raws, st = ffi.render_host(scene, want_raw=True)
    rgba_t, raw_t, _, _ = ffi.twin_render(scene)
    assert np.array_equal(outs[0].view(np.uint32), rgba_t.view(np.uint32))
    assert np.array_equal(raws[0].view(np.uint32), raw_t.view(np.uint32))
    assert st["h2d_bytes"] > 0 and st["d2h_bytes"] == 40 * 40 * 36
    # the `pbrt` code path: WorldEnd -> Scene::Render -> libpbrt_b200.so -> half EXR on disk
    assert ffi.host().pbrt_host_render_file(os.fsencode(sp)) == 0
    assert os.path.exists(exr)
    try:
        img = cv2.imread(exr, cv2.IMREAD_UNCHANGED)
    except Exception:
        img = None
    if img is not None:
        img = img[..., [2, 1, 0, 3]] if img.shape[-1] == 4 else img[..., ::-1]
        want = rgba_t[..., :img.shape[-1]].astype(np.float16).astype(np.float32)
        assert np.array_equal(img.astype(np.float32), want)


def test_live_reference_on_this_box(ffi, scene_dir):
    """Run the compiled reference here (it travels in oracle/_ref) on a scene outside the fixtures."""
    if not ffi.have_reference():
        pytest.skip("oracle/_ref not present in this snapshot")
    sp = gen_scene(os.path.join(scene_dir, "liveg.pbrt"), "c3", "--res", "96", "96", "--spp", "4", "4", "--filename",
                   os.path.join(scene_dir, "liveg.exr"))
    dump = os.path.join(scene_dir, "liveg.samples")
    ffi.run_reference(sp, keyed=True, dump_samples=dump)
    ref, _ = ffi.read_raw(os.path.join(scene_dir, "liveg.exr.raw"))
    rs = read_sample_dump(dump)
    scene = ffi.Scene(sp)
    ds, st = gpu_render(ffi, scene)
    rgba = ds.download(0)
    ls = ds.sample_radiance(0, rs.shape[0])
    s = image_error_stats(rgba, ref)
    d = np.abs(ls[:, :3] - rs[:, 2:5]).max(axis=1)
    print("live reference vs GPU:", s, "samples identical %.5f" % float((d == 0).mean()))
    assert s["frac_within_tol"] >= 0.90 and s["mean_rel"] <= 2e-3, s     # 16 spp glass path: SURVEY 8c tolerance
    assert (d == 0).mean() >= 0.97
    ds.close()


def test_full_size_properties(ffi, scene_dir):
    """BASELINE config 2 at full size (512x512, 16 spp): size-independent properties."""
    sp = gen_scene(os.path.join(scene_dir, "c2full.pbrt"), "c2", "--filename", os.path.join(scene_dir, "c2full.exr"))
    scene = ffi.Scene(sp)
    ds, st = gpu_render(ffi, scene, keep_samples=False, count=False)
    rgba, raw = ds.download(0, want_raw=True)
    assert st["camera_rays"] == 513 * 513 * 16
    assert np.isfinite(rgba).all() and (rgba >= 0).all()
    # box filter 0.5: a pixel receives its own 16 samples (weight 1 each), plus, rarely, a neighbour's
    # sample whose coordinate is exactly an integer (SURVEY App. C hazard 2)
    assert (raw[..., 4] >= 16).all() and (raw[..., 4] <= 18).all() and (raw[..., 4] == 16).mean() > 0.999
    assert ((rgba[..., 3] == 1) | (rgba[..., 3] == 0)).mean() > 0.99
    # the light quad is visible at the top of the frame at its emitted radiance (17, 12, 4)
    assert np.isclose(rgba[..., 0].max(), 17.0, rtol=1e-3)
    # determinism: a second render is bit-identical
    ds.render()
    assert np.array_equal(ds.download(0).view(np.uint32), rgba.view(np.uint32))
    # a vertical strip rendered alone matches the full frame bit for bit
    ds.render([ffi.Window(200, 0, 64, 512)])
    assert np.array_equal(ds.download(0).view(np.uint32), rgba[:, 200:264].view(np.uint32))
    ds.close()


# ---------------------------------------------------------------------------------------------------
# tile mode and the multi-GPU frame (pbrt_b200_render_tiles / _render_frame / _render_host_multi)
@pytest.mark.parametrize("variant", ["box", "mitchell", "ld", "whitted"])
def test_tile_mode_equals_full_frame(ffi, scene_dir, variant):
    """Tiles of the film window are bit-identical to the same pixels of the whole frame, and trace (box filter)
    practically no apron samples."""
    args = ["--res", "200", "152", "--spp", "3", "3"]
    if variant == "mitchell":
        args += ["--filter", "mitchell"]
    elif variant == "ld":
        args += ["--ld"]
    elif variant == "whitted":
        args += ["--integrator", "whitted", "--mirror"]
    sp = gen_scene(os.path.join(scene_dir, "tile_%s.pbrt" % variant), "c3", "--filename",
                   os.path.join(scene_dir, "tile.exr"), *args)
    scene = ffi.Scene(sp)
    ds = ffi.DeviceScene(scene)
    st_full = ds.render()
    full = ds.download(0)
    part = ffi.frame_partition(scene.full_window(), 3, 48)
    tiles = [t for t, _ in part]
    st = ds.render_tiles(tiles)
    for i, t in enumerate(tiles):
        got = ds.download(i)
        want = full[t.y0:t.y0 + t.height, t.x0:t.x0 + t.width]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (variant, i)
    if variant == "box":
        # a 0.5-wide box filter: only a pixel's own samples (and the rare one on an exact integer coordinate) reach
        # it, so tiles trace just those -- fewer than the frame's crop extent, which carries one more row and
        # column of samples that touch nothing (film/image.cpp:148-156)
        assert 200 * 152 * 9 <= st["camera_rays"] <= st_full["camera_rays"]
    else:
        assert st["camera_rays"] >= 200 * 152 * 9
    # single-rank pbrt_b200_render_frame is the plain frame
    ds.render_frame()
    ptr, n = ds.frame_device_ptr()
    assert n == full.size
    ds.close()


def test_tile_mode_full_size_frame(ffi, scene_dir):
    """BASELINE configs[2] at full size: row strips rendered as TILES equal the frame bit for bit -- including the
    handful of pixels a crop render of the same strips gets differently (tests/test_gpu_scale.py)."""
    sp = gen_scene(os.path.join(scene_dir, "c3tiles.pbrt"), "c3", "--filename", os.path.join(scene_dir, "c3tiles.exr"))
    scene = ffi.Scene(sp)
    ds = ffi.DeviceScene(scene)
    st_full = ds.render()
    full = ds.download(0)
    part = ffi.frame_partition(scene.full_window(), 8, 0)
    tiles = [t for t, _ in part]
    assert 150 <= len(tiles) <= 400
    for lo in range(0, len(tiles), 32):
        chunk = tiles[lo:lo + 32]
        ds.render_tiles(chunk)
        for i, t in enumerate(chunk):
            got = ds.download(i)
            assert np.array_equal(got.view(np.uint32), full[t.y0:t.y0 + t.height, t.x0:t.x0 + t.width].view(np.uint32)), t.y0
    ds.close()


def _n_devices(ffi):
    return int(ffi.cuda().pbrt_b200_device_count())


def test_multi_gpu_frame_bit_identical(ffi, scene_dir):
    """pbrt_b200_render_host_multi on 2 (and all) GPUs of this box == the one-GPU frame, bit for bit; PBRT_GPUS drives
    the same path from the scene-file front end."""
    n = _n_devices(ffi)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    exr = os.path.join(scene_dir, "multi.exr")
    sp = gen_scene(os.path.join(scene_dir, "multi.pbrt"), "c3", "--res", "512", "384", "--spp", "4", "4", "--filename", exr)
    scene = ffi.Scene(sp)
    one, _, st1 = ffi.render_host(scene)
    for k in sorted({2, n}):
        frame, st = ffi.render_host_multi(scene, list(range(k)))
        assert np.array_equal(frame.view(np.uint32), one[0].view(np.uint32)), k
        assert 512 * 384 * 16 <= st["camera_rays"] <= st1["camera_rays"]
        assert abs(st["intersect_rays"] + st["shadow_rays"] - st1["intersect_rays"] - st1["shadow_rays"]) <= 1e-2 * st1["intersect_rays"]
    # the `pbrt scene.pbrt` path with PBRT_GPUS=2 writes the same EXR as with one GPU
    outs = []
    for gpus in ("1", "2"):
        os.environ["PBRT_GPUS"] = gpus
        try:
            assert ffi.host().pbrt_host_render_file(os.fsencode(sp)) == 0
        finally:
            del os.environ["PBRT_GPUS"]
        outs.append(open(exr, "rb").read())
    assert outs[0] == outs[1]
