"""CPU suite (-m "not gpu"): pins the oracle against the reference's outputs and checks host logic.

The reference (mmp/pbrt-v1) ships no tests, golden images or example scenes (SURVEY.md section
4), so the anchors are outputs of the reference itself:
  * tests/golden/*.npz -- produced by tests/golden/make_golden.py from the compiled reference
    (oracle/_ref); committed, so they travel to the GPU box;
  * oracle/_ref run live, when it has been built in this checkout (it needs /root/reference to
    build, but runs anywhere).
The oracle twin (oracle/twin.cpp) is checked two ways: built with glibc's float libm it must
reproduce the reference BIT FOR BIT (image and every per-sample radiance); built the way the GPU
computes transcendentals (double precision, rounded once) it must stay within the stated
tolerance.
"""
import ctypes as C
import glob
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from parity_util import ROOT, gen_scene, image_error_stats, read_sample_dump

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))

# Stated tolerance for the double-rounded-libm build (and hence the GPU) vs the reference on the
# fp32 film handed to WriteRGBAImage: a pixel is "within tolerance" when
# |delta| <= max(2^-10 * |ref|, 1e-3) in every channel.  Required fraction of such pixels:
FRAC_WITHIN_TOL = 0.97
# and the image mean (RGB) must agree to this relative error:
MEAN_REL = 5e-3


def twin_with(ffi, lib_name):
    """Returns a render function bound to a specific twin build."""
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", lib_name))
    lib.pbrt_twin_sample_count.restype = C.c_uint64
    lib.pbrt_twin_sample_count.argtypes = [C.c_void_p, C.c_void_p]
    lib.pbrt_twin_render.argtypes = [C.c_void_p] * 6

    def render(scene, window=None):
        w = window or scene.full_window()
        rgba = np.zeros((w.height, w.width, 4), np.float32)
        ns = lib.pbrt_twin_sample_count(scene.ptr, C.byref(w))
        samples = np.zeros((ns, 6), np.float32)
        st = ffi.Stats()
        rc = lib.pbrt_twin_render(scene.ptr, C.byref(w), rgba.ctypes.data_as(C.c_void_p), None,
                                  samples.ctypes.data_as(C.c_void_p), C.byref(st))
        assert rc == 0
        return rgba, samples, st.as_dict()

    return render


def load_case(ffi, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    cwd = os.getcwd()
    os.chdir(GOLDEN)
    try:
        scene = ffi.Scene(os.path.join(GOLDEN, name + ".pbrt"))
    finally:
        os.chdir(cwd)
    return scene, g["rgba"], g["samples"], g["stats"]


def test_golden_cases_present():
    assert len(CASES) >= 15


@pytest.mark.parametrize("name", CASES)
def test_twin_libm_bit_exact_vs_golden(ffi, name):
    """glibc-float-libm twin == compiled reference, bit for bit."""
    scene, rgba_ref, samples_ref, stats_ref = load_case(ffi, name)
    rgba, samples, st = twin_with(ffi, "libpbrt_twin_libm.so")(scene)
    assert samples.shape == samples_ref.shape
    assert np.array_equal(samples.view(np.uint32), samples_ref.view(np.uint32)), "per-sample radiance differs"
    assert np.array_equal(rgba.view(np.uint32), rgba_ref.view(np.uint32)), "image differs"
    assert st["camera_rays"] == stats_ref[0]
    # the reference prints large counters rounded to 3-4 digits ("140.9k")
    for got, want in ((st["shadow_rays"], stats_ref[1]), (st["points_shaded"], stats_ref[2])):
        assert abs(got - want) <= max(1, 0.006 * want)


@pytest.mark.parametrize("name", CASES)
def test_twin_double_libm_within_tolerance(ffi, name):
    """The arithmetic the GPU uses (double-rounded transcendentals) vs the reference."""
    scene, rgba_ref, samples_ref, _ = load_case(ffi, name)
    rgba, samples, _ = twin_with(ffi, "libpbrt_twin.so")(scene)
    s = image_error_stats(rgba, rgba_ref)
    assert s["frac_within_tol"] >= FRAC_WITHIN_TOL, s
    assert s["mean_rel"] <= MEAN_REL, s
    # image coordinates never involve libm: identical
    assert np.array_equal(samples[:, :2], samples_ref[:, :2])
    d = np.abs(samples[:, 2:5] - samples_ref[:, 2:5]).max(axis=1)
    assert (d <= 1e-4).mean() >= 0.995


def test_c1_exact_equals_keyed():
    """C1 (no jitter, point light) does not depend on the RNG: unmodified reference == keyed reference."""
    a = np.load(os.path.join(GOLDEN, "c1_exact.npz"))
    b = np.load(os.path.join(GOLDEN, "c1.npz"))
    assert np.array_equal(a["rgba"], b["rgba"])


def test_survey_appendix_g_anchor(ffi, scene_dir):
    """pixel values of full-size C1 from the unmodified reference."""
    sp = gen_scene(os.path.join(scene_dir, "c1_full.pbrt"), "c1", "--filename", os.path.join(scene_dir, "c1_full.exr"))
    scene = ffi.Scene(sp)
    rgba, samples, st = twin_with(ffi, "libpbrt_twin_libm.so")(scene)
    assert st["camera_rays"] == 66049
    want = {(128, 128): (float.fromhex("0x1.58afdp-3"), float.fromhex("0x1.89ed7ap-5"), float.fromhex("0x1.89ed8cp-5")),
            (100, 90): (float.fromhex("0x1.607754p-2"), float.fromhex("0x1.92d17ap-4"), float.fromhex("0x1.92d19p-4")),
            (150, 110): (float.fromhex("0x1.643e98p-3"), float.fromhex("0x1.9722eep-5"), float.fromhex("0x1.972304p-5")),
            (90, 128): (float.fromhex("0x1.a4882p-3"), float.fromhex("0x1.e09b9ap-5"), float.fromhex("0x1.e09ba2p-5"))}
    for (x, y), rgb in want.items():
        assert tuple(float(v) for v in rgba[y, x, :3]) == rgb, (x, y)
    assert not rgba[60, 128, :3].any() and not rgba[0, 0, :3].any()
    assert int((rgba[..., :3].sum(axis=-1) > 0).sum()) == 10136


def test_live_reference_when_built(ffi, scene_dir):
    """If oracle/_ref is present, render a scene that is NOT among the fixtures and compare live."""
    if not ffi.have_reference():
        pytest.skip("oracle/_ref not built in this checkout")
    sp = gen_scene(os.path.join(scene_dir, "live.pbrt"), "c3", "--res", "40", "28", "--spp", "3", "2", "--filter",
                   "gaussian", "--filename", os.path.join(scene_dir, "live.exr"))
    dump = os.path.join(scene_dir, "live.samples")
    ffi.run_reference(sp, keyed=True, dump_samples=dump)
    ref, _ = ffi.read_raw(os.path.join(scene_dir, "live.exr.raw"))
    rs = read_sample_dump(dump)
    scene = ffi.Scene(sp)
    rgba, samples, _ = twin_with(ffi, "libpbrt_twin_libm.so")(scene)
    assert np.array_equal(samples.view(np.uint32), rs.view(np.uint32))
    assert np.array_equal(rgba.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_scenes_twin_equals_live_reference(ffi, tmp_path, seed):
    """Random scenes mixing every supported plugin (tests/fuzz_scenes.py): compiled reference == libm twin,
    every sample and every pixel, bit for bit.  tools/fuzz_parity.py runs the same check over hundreds."""
    if not ffi.have_reference():
        pytest.skip("oracle/_ref not built in this checkout")
    import random
    from fuzz_scenes import scene as fuzz_scene
    sp = str(tmp_path / ("fuzz_%d.pbrt" % seed))
    open(sp, "w").write(fuzz_scene(random.Random(900001 + seed), seed))
    dump = str(tmp_path / "fuzz.samples")
    ffi.run_reference(sp, keyed=True, dump_samples=dump, cwd=str(tmp_path))
    ref, _ = ffi.read_raw(str(tmp_path / ("fuzz_%d.exr.raw" % seed)))
    rs = read_sample_dump(dump)
    rgba, samples, _ = twin_with(ffi, "libpbrt_twin_libm.so")(ffi.Scene(sp))
    assert np.array_equal(samples.view(np.uint32), rs.view(np.uint32))
    assert np.array_equal(rgba.view(np.uint32), ref.view(np.uint32))


# ---------------------------------------------------------------- keyed RNG known answers
def _mix64(z):
    z &= (1 << 64) - 1
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & ((1 << 64) - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & ((1 << 64) - 1)
    return z ^ (z >> 31)


def _key(x, y, s, stream):
    k = ((x & 0xffffffff) << 32) ^ ((y & 0xffffffff) << 12) ^ ((s & 0xffffffff) << 2) ^ (stream & 0xffffffff)
    return _mix64(k)


def _f32(key, ctr):
    u = _mix64(key + ctr * 0x9E3779B97F4A7C15) >> 32
    return np.float32(u & 0xffffff) / np.float32(1 << 24)


def test_keyed_rng_drives_camera_samples(ffi, scene_dir):
    """Independent Python restatement of the keyed generator predicts the
    stratified image samples the twin produces (samplers/stratified.cpp:98-112)."""
    sp = gen_scene(os.path.join(scene_dir, "rng.pbrt"), "c2", "--res", "8", "8", "--spp", "3", "2",
                   "--filename", os.path.join(scene_dir, "rng.exr"))
    scene = ffi.Scene(sp)
    rays, xy = ffi.twin_camera_rays(scene)
    # extent for the box filter (0.5): [0, 9) x [0, 9)
    i = 0
    for y in range(0, 9):
        for x in range(0, 9):
            key = _key(x, y, 0xFFFF, 0)
            for k in range(6):
                sx, sy = k % 3, k // 3
                ex = (np.float32(sx) + _f32(key, 2 * k + 1)) * (np.float32(1) / np.float32(3)) + np.float32(x)
                ey = (np.float32(sy) + _f32(key, 2 * k + 2)) * (np.float32(1) / np.float32(2)) + np.float32(y)
                assert xy[i, 0] == np.float32(ex) and xy[i, 1] == np.float32(ey), (x, y, k)
                i += 1
    assert i == xy.shape[0]


# ---------------------------------------------------------------- host front end
def test_parser_and_flatten(ffi, scene_dir):
    sp = gen_scene(os.path.join(scene_dir, "h.pbrt"), "c3", "--res", "20", "10", "--spp", "2", "3", "--crop", "0.1",
                   "0.6", "0", "1", "--filename", os.path.join(scene_dir, "h.exr"))
    s = ffi.Scene(sp)
    assert (s.xres, s.yres, s.spp) == (20, 10, 6)
    assert (s.x0, s.width, s.y0, s.height) == (2, 10, 0, 10)
    assert s.n_prims == 33 and s.n_lights == 1 and s.integrator == 2   # 32 triangles + sphere, path
    assert s.n_nodes >= 33


def test_kd_tree_invariants(ffi, scene_dir):
    """Closest-hit and any-hit walks agree with each other on camera rays and random interior rays."""
    sp = gen_scene(os.path.join(scene_dir, "kd.pbrt"), "c4", "--res", "16", "16", "--spp", "1", "1", "--nphi", "40",
                   "--ntheta", "20", "--filename", os.path.join(scene_dir, "kd.exr"))
    scene = ffi.Scene(sp)
    rays, _ = ffi.twin_camera_rays(scene)
    rng = np.random.default_rng(7)
    extra = np.zeros((2000, 8), np.float32)
    extra[:, :3] = rng.uniform(50, 500, (2000, 3))
    d = rng.normal(size=(2000, 3))
    extra[:, 3:6] = d / np.linalg.norm(d, axis=1, keepdims=True)
    extra[:, 6] = 1e-3
    extra[:, 7] = np.inf
    allrays = np.concatenate([rays, extra])
    prim, t, b1, b2 = ffi.twin_intersect(scene, allrays)
    occ = ffi.twin_intersect_p(scene, allrays)
    assert ((prim >= 0) == (occ == 1)).all()
    assert (prim >= 0).mean() > 0.7          # a box open towards the camera: most rays hit
    assert prim.max() < scene.n_prims
    assert (t[prim >= 0] >= allrays[prim >= 0, 6]).all()


def test_c_abi_exports_every_declared_symbol():
    """libpbrt_b200.so must load without a GPU and export exactly the header's entry points."""
    hdr = open(os.path.join(ROOT, "include", "pbrt_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pbrt_b200_[a-z_]+)\s*\(", hdr)))
    assert len(declared) >= 12
    lib_path = os.path.join(ROOT, "pbrt-v1_b200", "csrc", "libpbrt_b200.so")
    if not os.path.exists(lib_path):
        subprocess.run(["make"], cwd=os.path.dirname(lib_path), check=True)
    lib = C.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert lib.pbrt_b200_abi_version() == 4


def test_product_fails_loudly_without_gpu(ffi, scene_dir):
    """No CPU fallback: without a device the render entry points return an error code."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    sp = gen_scene(os.path.join(scene_dir, "nogpu.pbrt"), "c1", "--res", "8", "8", "--filename",
                   os.path.join(scene_dir, "nogpu.exr"))
    scene = ffi.Scene(sp)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        ffi.render_host(scene)
    with pytest.raises(RuntimeError):
        ffi.DeviceScene(scene)


def test_product_does_not_reference_oracle():
    """The product sources never include, link or load anything under oracle/."""
    for path in glob.glob(os.path.join(ROOT, "pbrt-v1_b200", "**", "*"), recursive=True):
        if os.path.isdir(path) or not path.endswith((".h", ".cpp", ".cu", "Makefile")):
            continue
        text = open(path, errors="replace").read()
        for line in text.splitlines():
            code = line.split("//")[0]
            assert "oracle/" not in code or "#include" not in code, (path, line)
            assert "libpbrt_twin" not in code, (path, line)


def test_bestcandidate_table_matches_reference():
    """pbrt-v1_b200/host/bestcandidate_table.bin is the output of tools/gen_bestcandidate_table.cpp (our
    restatement of the generator); it must equal the literals the reference compiles in
    (samplers/sampledata.cpp), bit for bit as float32.  The reference is only present in the build container."""
    import re
    table = np.fromfile(os.path.join(ROOT, "pbrt-v1_b200", "host", "bestcandidate_table.bin"), np.float32)
    assert table.shape == (4096 * 5,)
    assert np.isfinite(table).all() and (table >= 0).all() and (table <= 1).all()
    ref = "/root/reference/samplers/sampledata.cpp"
    if not os.path.exists(ref):
        pytest.skip("reference sources not present")
    lits = re.findall(r"([0-9.]+)f,", open(ref).read())
    want = np.array([np.float32(v) for v in lits], np.float32)
    assert want.shape == table.shape
    assert np.array_equal(want.view(np.uint32), table.view(np.uint32))


def test_unsupported_texture_modes_are_reported_not_rendered(ffi, tmp_path, capfd):
    """Checkerboards that draw random numbers inside the texture (supersampled, 3D) are refused at parse time:
    an error is reported, the texture is not created and the material falls back to its default operand,
    as the reference does for any texture it cannot create (api.cpp:318-331)."""
    sp = str(tmp_path / "unsupported.pbrt")
    open(sp, "w").write("""LookAt 0 0 -5 0 0 0 0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [8] "integer yresolution" [8] "string filename" ["u.exr"]
Sampler "stratified" "integer xsamples" [1] "integer ysamples" [1]
SurfaceIntegrator "path"
WorldBegin
LightSource "point" "color I" [10 10 10] "point from" [0 3 -3]
Texture "a" "color" "checkerboard" "string aamode" ["supersample"]
Texture "b" "color" "checkerboard" "integer dimension" [3]
Texture "c" "color" "checkerboard"
Material "matte" "texture Kd" "c"
Shape "sphere" "float radius" [1]
WorldEnd
""")
    scene = ffi.Scene(sp)
    err = capfd.readouterr().err
    assert "supersample" in err or "random numbers" in err
    assert err.count("Unable to create color texture") == 2
    assert scene.n_prims == 1


def test_noise_permutation_matches_reference():
    """pb_noise_table.h states Perlin's published permutation; where the reference is present, compare it with
    the array the reference's Noise() indexes (core/texture.cpp:31-61)."""
    import re
    ref_src = "/root/reference/core/texture.cpp"
    if not os.path.exists(ref_src):
        pytest.skip("reference sources not present")
    body = re.search(r"NoisePerm\[2 \* NOISE_PERM_SIZE\] = \{(.*?)\};", open(ref_src).read(), re.S).group(1)
    ref = [int(x) for x in re.findall(r"\d+", re.sub(r"//.*", "", body))]
    mine_src = open(os.path.join(ROOT, "pbrt-v1_b200", "csrc", "pb_noise_table.h")).read()
    mine = [int(x) for x in re.findall(r"\d+", re.search(r"PB_NOISE_PERM_HALF\[PB_NOISE_PERM_SIZE\] = \{(.*?)\};", mine_src, re.S).group(1))]
    assert len(mine) == 256 and ref == mine + mine


def test_parallel_kd_build_equals_sequential(ffi, tmp_path):
    """The host kd build forks big nodes' children onto threads and splices the subtrees; the arrays must be
    the sequential build's (== the reference's tree), bit for bit."""
    import hashlib
    import subprocess
    sp = gen_scene(str(tmp_path / "kd.pbrt"), "c4", "--res", "16", "16", "--spp", "1", "1", "--nphi", "200", "--ntheta", "100",
                   "--filename", str(tmp_path / "kd.exr"))
    code = (
        "import sys, hashlib, ctypes as C\n"
        "sys.path.insert(0, %r)\n"
        "from parity_util import load_ffi\n"
        "ffi = load_ffi()\n"
        "sc = ffi.Scene(sys.argv[1])\n"
        "class Head(C.Structure):\n"
        "    _fields_ = [('abi', C.c_uint32), ('nodes', C.c_void_p), ('n_nodes', C.c_uint32), ('n_top', C.c_uint32),\n"
        "                ('refs', C.c_void_p), ('n_refs', C.c_uint32)]\n"
        "h = Head.from_address(sc.ptr)\n"
        "a = C.string_at(h.nodes, 8 * h.n_nodes); b = C.string_at(h.refs, 4 * h.n_refs)\n"
        "print(h.n_nodes, h.n_refs, hashlib.md5(a).hexdigest(), hashlib.md5(b).hexdigest())\n" % os.path.join(ROOT, "tests"))
    outs = []
    for threads in ("1", "8"):
        env = dict(os.environ, PBRT_KD_THREADS=threads)
        r = subprocess.run([sys.executable, "-c", code, sp], env=env, capture_output=True, text=True, check=True)
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1], outs
    assert int(outs[0].split()[0]) > 10000



def test_c4_scale_tree_and_crop_vs_reference(ffi, tmp_path):
    """BASELINE configs[3] geometry at its real size (101 k triangles, 2.6 M kd nodes): the host front end's
    kd-tree equals the one the compiled reference builds word for word (every split, child index and leaf
    list: accelerators/kdtree.cpp:141-312, incl. the maxDepth formula and the badRefines / retry paths a deep
    tree exercises), and a crop rendered as concurrent cropwindow processes of the reference equals the libm
    twin bit for bit, sample by sample."""
    if not ffi.have_reference():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from parity_util import host_kd_arrays, read_kd_dump, row_strips, run_reference_windows
    td = str(tmp_path)
    sp = gen_scene(os.path.join(td, "c4.pbrt"), "c4", "--res", "256", "256", "--spp", "2", "2", "--filename",
                   os.path.join(td, "c4.exr"))
    win = (116, 116, 24, 24)
    kd = os.path.join(td, "kd.bin")
    res, _ = run_reference_windows(ffi, sp, 256, 256, row_strips(win, 3), td, keyed=True, dump_samples=True, dump_kd=kd)
    scene = ffi.Scene(sp)
    b, nodes, refs, ntop = host_kd_arrays(scene)
    rb, rnodes, rrefs, rprims = read_kd_dump(kd)
    assert rprims == scene.n_prims == 101772
    assert res[0]["kd"][0] == nodes.shape[0] == ntop and res[0]["kd"][2] == refs.shape[0]
    assert np.array_equal(b, rb)
    assert np.array_equal(nodes, rnodes)
    assert np.array_equal(refs, rrefs)
    render = twin_with(ffi, "libpbrt_twin_libm.so")
    for r in res:
        rgba, samples, st = render(scene, ffi.Window(*r["window"]))
        assert np.array_equal(rgba.view(np.uint32), r["rgba"].view(np.uint32)), r["window"]
        assert np.array_equal(samples.view(np.uint32), r["samples"].view(np.uint32)), r["window"]
        assert st["intersect_rays"] == r["intersect"] and st["shadow_rays"] == r["intersect_p"]


def test_treelet_node_order_changes_no_hit_and_no_counter(ffi, tmp_path):
    """csrc/pb_treelet.h re-orders the device copy of the top-level tree (sibling pairs in line-sized treelets);
    the walk over it must return the same hits and visit / test counts as over the reference's depth-first array."""
    for name, extra in (("c3", ()), ("c4", ("--nphi", "160", "--ntheta", "80"))):
        sp = gen_scene(str(tmp_path / ("tl_%s.pbrt" % name)), name, "--res", "40", "40", "--spp", "2", "2",
                       "--filename", str(tmp_path / "tl.exr"), *extra)
        scene = ffi.Scene(sp)
        rays, _ = ffi.twin_camera_rays(scene)
        prim, t, _, _ = ffi.twin_intersect(scene, rays)
        ok = prim >= 0
        rng = np.random.default_rng(5)
        p = rays[ok, :3] + rays[ok, 3:6] * t[ok, None]
        d = rng.normal(size=p.shape).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        sec = np.zeros((p.shape[0], 8), np.float32)
        sec[:, :3] = p; sec[:, 3:6] = d; sec[:, 6] = 1e-3; sec[:, 7] = np.inf
        allrays = np.concatenate([rays, sec])
        a = ffi.twin_intersect_treelets(scene, allrays, 0)
        b = ffi.twin_intersect_treelets(scene, allrays, 1)
        assert np.array_equal(a[0], b[0])
        for k in (1, 2, 3):
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32))
        assert np.array_equal(a[4], b[4]) and a[4][0] > 0
        scene.close()


def test_threaded_event_sort_is_std_sort(ffi):
    """host/kdtree.cpp sorts a big node's events with libstdc++'s introsort run on several threads; the result must
    be std::sort's permutation exactly, equal keys included (their order decides which child a primitive lands in)."""
    lib = ffi.host()
    lib.pbrt_host_test_sort_events.restype = C.c_long
    lib.pbrt_host_test_sort_events.argtypes = [C.c_long, C.c_long, C.c_uint, C.c_int]
    for n, distinct, seed, threads in ((200000, 50, 1, 4), (300000, 3000, 2, 8), (150001, 7, 3, 3), (400000, 400000, 4, 16),
                                       (1000, 10, 5, 8)):
        assert lib.pbrt_host_test_sort_events(n, distinct, seed, threads) == 0, (n, distinct, seed, threads)


def test_tokenizer_numbers_round_like_atof(ffi):
    """The scene tokenizer reads a NUMBER as (float)atof(text) (pbrtlex.l:96): the correctly rounded double, then
    rounded to float.  host/parser.cpp gets that double from std::from_chars; Python's float() is the same correctly
    rounded double, so every literal must come out as np.float32(float(literal)) bit for bit -- incl. halfway cases
    of the float grid, denormals, overflow, leading '+', '.5', '5.', exponents."""
    lib = ffi.host()
    lib.pbrt_host_test_scan_numbers.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(7)
    lits = ["0", "-0", "+5", "-.5", "5.", ".25", "1e5", "1E-5", "1e+5", "3.4028235e38", "3.4028236e38", "1e39", "-1e39",
            "1.17549435e-38", "1e-45", "7e-46", "1e-50", "0.1", "0.30000001192092896", "16777217", "16777219",
            "1.00000005960464477539", "1.00000005960464477540", "123456789.987654321", "555.0", "0.000001"]
    for _ in range(4000):
        mant = "%d.%s" % (rng.integers(0, 1000), "".join(str(d) for d in rng.integers(0, 10, rng.integers(1, 18))))
        lits.append(("-" if rng.random() < 0.3 else "") + mant + ("e%d" % rng.integers(-44, 39) if rng.random() < 0.5 else ""))
    for x in rng.random(2000).astype(np.float32):     # halfway points between neighbouring floats, written exactly
        lo = float(x); hi = float(np.nextafter(x, np.float32(2.0)))
        lits.append("%.60g" % ((lo + hi) / 2))
    out = np.zeros(len(lits), np.float32)
    n = lib.pbrt_host_test_scan_numbers(" ".join(lits).encode(), out.ctypes.data, len(lits))
    assert n == len(lits)
    with np.errstate(over="ignore"):
        want = np.array([np.float32(float(s)) for s in lits], np.float32)
    bad = [(lits[i], out[i], want[i]) for i in range(n) if out[i].view(np.uint32) != want[i].view(np.uint32)]
    assert not bad, bad[:5]


@pytest.mark.skipif(not os.path.exists("/root/reference/core/pbrtlex.l"), reason="needs the reference tree")
def test_parser_vocabulary_and_numbers_follow_the_reference_lexer(ffi):
    """The hand-written tokenizer (flex / bison are not in the image) against the TEXT of the reference's lexer:
    every directive keyword of core/pbrtlex.l:113-144 has a handler in host/parser.cpp, and NUMBER tokens are cut
    exactly where pbrtlex.l:96's regular expression (longest match) cuts them."""
    lex = open("/root/reference/core/pbrtlex.l").read()
    keywords = re.findall(r"^([A-Z][A-Za-z]+)\s*\{ return [A-Z]+;", lex, re.M)
    assert len(keywords) >= 32 and "WorldEnd" in keywords and "CoordSysTransform" in keywords
    src = open(os.path.join(ROOT, "pbrt-v1_b200", "host", "parser.cpp")).read()
    missing = [k for k in keywords if ('kw == "%s"' % k) not in src]
    assert not missing, missing
    # the NUMBER macro, with its alternatives ordered so that Python's leftmost alternation is flex's longest match
    m = re.search(r"^NUMBER\s+(\S+)$", lex, re.M)
    assert m and m.group(1) == r"[-+]?([0-9]+|(([0-9]+\.[0-9]*)|(\.[0-9]+)))([eE][-+]?[0-9]+)?"
    number = re.compile(r"[-+]?(?:[0-9]+\.[0-9]*|\.[0-9]+|[0-9]+)(?:[eE][-+]?[0-9]+)?")
    lib = ffi.host()
    lib.pbrt_host_test_scan_numbers.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(11)
    pieces = ["1", ".", "5", "e", "E", "+", "-", "0", "12", " ", "9.", ".25", "e3", "E-2", "+7", "-.5", " "]
    compared = 0
    for _ in range(3000):
        text = "".join(pieces[i] for i in rng.integers(0, len(pieces), rng.integers(1, 10)))
        # what flex would hand over: at every position the longest NUMBER, else skip one character (blank, or an
        # identifier / stray character, which the tokenizer reports differently -- only strings where everything
        # besides blanks is a number are compared)
        pos, want, clean = 0, [], True
        while pos < len(text):
            if text[pos] == " ":
                pos += 1
                continue
            mm = number.match(text, pos)
            if not mm:
                clean = False
                break
            want.append(mm.group(0))
            pos = mm.end()
        if not clean or not want:
            continue
        out = np.zeros(len(want) + 4, np.float32)
        n = lib.pbrt_host_test_scan_numbers(text.encode(), out.ctypes.data, len(out))
        assert n == len(want), (text, want, n)
        with np.errstate(over="ignore"):
            exp = np.array([np.float32(float(w)) for w in want], np.float32)
        assert np.array_equal(out[:n].view(np.uint32), exp.view(np.uint32)), (text, want, out[:n])
        compared += 1
    assert compared >= 500, compared


@pytest.mark.skipif(not os.path.exists("/root/reference/core/pbrtparse.y"), reason="needs the reference tree")
def test_parser_statement_shapes_follow_the_reference_grammar(tmp_path):
    """The shapes of the statements, read off the TEXT of core/pbrtparse.y:294-574: which directives take a quoted
    name + parameter list, which a fixed count of numbers, which a 16-entry array.  A scene that follows the
    grammar loads; the same scene with one number too few in a fixed-count directive does not."""
    import subprocess
    y = open("/root/reference/core/pbrtparse.y").read()
    rules = dict((m.group(1), m.group(2).split()) for m in
                 re.finditer(r"^(?:ri_stmt:|\|)\s*([A-Z]+)((?:[ \t]+(?:STRING|paramlist|num_array|real_num_array|NUM))*)[ \t]*$", y, re.M))
    assert rules["LOOKAT"] == ["NUM"] * 9 and rules["ROTATE"] == ["NUM"] * 4 and rules["SCALE"] == ["NUM"] * 3
    assert rules["TRANSLATE"] == ["NUM"] * 3 and rules["CONCATTRANSFORM"] == ["num_array"] and rules["TRANSFORM"] == ["real_num_array"]
    for d in ("ACCELERATOR", "AREALIGHTSOURCE", "CAMERA", "FILM", "LIGHTSOURCE", "MATERIAL", "PIXELFILTER", "SAMPLER", "SHAPE",
              "SURFACEINTEGRATOR", "VOLUMEINTEGRATOR", "VOLUME"):
        assert rules[d] == ["STRING", "paramlist"], d
    for d in ("COORDINATESYSTEM", "COORDSYSTRANSFORM", "OBJECTBEGIN", "OBJECTINSTANCE", "SEARCHPATH"):
        assert rules[d] == ["STRING"], d
    for d in ("ATTRIBUTEBEGIN", "ATTRIBUTEEND", "IDENTITY", "OBJECTEND", "REVERSEORIENTATION", "TRANSFORMBEGIN", "TRANSFORMEND",
              "WORLDBEGIN", "WORLDEND"):
        assert rules[d] == [], d
    # parameter-list type words (pbrtparse.y:557-565), in the reference's decoding order -- it matters: the test is
    # a prefix match (strncmp), and "float" / "integer" ... share no prefix, so the set is what must agree
    types = re.findall(r'TRY_DECODING_TYPE\("([a-z]+)",', y)
    assert sorted(types) == sorted(["float", "integer", "bool", "point", "vector", "normal", "string", "texture", "color"])
    psrc = open(os.path.join(ROOT, "pbrt-v1_b200", "host", "parser.cpp")).read()
    for tname in types:
        assert ('{"%s", PT_' % tname) in psrc, tname
    names = {"LOOKAT": "LookAt", "ROTATE": "Rotate", "SCALE": "Scale", "TRANSLATE": "Translate",
             "CONCATTRANSFORM": "ConcatTransform", "TRANSFORM": "Transform"}
    ident16 = "1 0 0 0 0 1 0 0 0 0 1 0 0 0 0 1"

    def stmt(d, drop=0):
        if rules[d] in (["num_array"], ["real_num_array"]):
            vals = ident16.split()
            return "%s [%s]" % (names[d], " ".join(vals[:16 - drop]))
        n = len(rules[d]) - drop
        base = {"LOOKAT": "0 0 5 0 0 0 0 1 0", "ROTATE": "30 0 1 0", "SCALE": "1 2 1", "TRANSLATE": "0 .5 0"}[d].split()
        return "%s %s" % (names[d], " ".join(base[:n]))

    def scene(bad=None):
        pre = [stmt("LOOKAT", 1 if bad == "LOOKAT" else 0)]
        body = [stmt(d, 1 if bad == d else 0) for d in ("ROTATE", "SCALE", "TRANSLATE", "CONCATTRANSFORM", "TRANSFORM")]
        return "\n".join(pre + ['Camera "perspective" "float fov" [40]', 'Film "image" "integer xresolution" [4] "integer yresolution" [4] "string filename" ["g.exr"]',
                                 'Sampler "stratified" "integer xsamples" [1] "integer ysamples" [1]', 'SurfaceIntegrator "whitted"', "WorldBegin",
                                 'LightSource "point" "color I" [1 1 1]', "AttributeBegin", 'CoordinateSystem "here"'] + body +
                         ['CoordSysTransform "here"', 'Material "matte"', 'Shape "sphere" "float radius" [1]', "AttributeEnd", "WorldEnd"]) + "\n"

    code = ("import sys; sys.path.insert(0, %r)\nfrom parity_util import load_ffi\nffi = load_ffi()\n"
            "sc = ffi.Scene(sys.argv[1]); print('loaded', sc.n_prims)\n" % os.path.join(ROOT, "tests"))

    def loads(text, tag):
        p = tmp_path / ("g_%s.pbrt" % tag)
        p.write_text(text)
        r = subprocess.run([sys.executable, "-c", code, str(p)], capture_output=True, text=True)
        return r.returncode == 0 and "loaded 1" in r.stdout

    assert loads(scene(), "ok")
    for d in ("LOOKAT", "ROTATE", "SCALE", "TRANSLATE", "CONCATTRANSFORM", "TRANSFORM"):
        assert not loads(scene(bad=d), d), d


def test_todo_entry_tmin_is_the_current_tmax(ffi, tmp_path):
    """The GPU walk keeps 8-byte todo entries (node, tmax) instead of the reference's (node, tmin, tmax): when an
    entry is popped, its tmin is -- bit for bit -- the tmax current at that moment.  Checked on the reference's own
    loop over camera and random bounce rays of two trees."""
    lib = ffi.twin()
    lib.pbrt_twin_stack_invariant.restype = C.c_uint64
    lib.pbrt_twin_stack_invariant.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    for name, extra in (("c3", ()), ("c4", ("--nphi", "160", "--ntheta", "80"))):
        sp = gen_scene(str(tmp_path / ("inv_%s.pbrt" % name)), name, "--res", "48", "48", "--spp", "2", "2",
                       "--filename", str(tmp_path / "inv.exr"), *extra)
        scene = ffi.Scene(sp)
        rays, _ = ffi.twin_camera_rays(scene)
        prim, t, _, _ = ffi.twin_intersect(scene, rays)
        ok = prim >= 0
        rng = np.random.default_rng(9)
        p = rays[ok, :3] + rays[ok, 3:6] * t[ok, None]
        d = rng.normal(size=p.shape).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        sec = np.zeros((p.shape[0], 8), np.float32)
        sec[:, :3] = p; sec[:, 3:6] = d; sec[:, 6] = 1e-3; sec[:, 7] = np.inf
        allrays = np.ascontiguousarray(np.concatenate([rays, sec]), np.float32)
        pops = C.c_uint64(0)
        bad = lib.pbrt_twin_stack_invariant(scene.ptr, allrays.ctypes.data, allrays.shape[0], C.byref(pops))
        assert bad == 0 and pops.value > 1000, (name, bad, pops.value)
        scene.close()
