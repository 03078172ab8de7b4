#!/usr/bin/env python3
"""Generates the golden fixtures in this directory from the COMPILED REFERENCE.

Run in the build container (needs oracle/_ref, i.e. /root/reference + oracle/build_ref.sh):

    python tests/golden/make_golden.py

For every case below it writes the scene file (<case>.pbrt, produced by scenes/gen_scenes.py)
and <case>.npz holding what the reference renderer produced for it:

    rgba      float32 [h, w, 4]  the arrays ImageFilm::WriteImage hands to WriteRGBAImage
    samples   float32 [n, 6]     imageX, imageY, L.rgb (after the scene.cpp:60-74 guards), alpha
                                 for every image sample in the reference's order
    stats     int64 [3]          camera rays, shadow rays, points shaded (the reference's own
                                 post-render statistics)

Cases marked keyed=True come from oracle/_ref/pbrt_ref_keyed (reference + per-sample RNG
re-keying); "c1_exact" comes from the UNMODIFIED reference
(pbrt_ref_exact) and pins the one configuration whose result does not depend on the
random-number stream (Whitted, point light, jitter off).
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from parity_util import gen_scene, load_ffi, read_sample_dump, write_exr  # noqa: E402

CASES = [
    # name, generator scene, args, keyed
    ("c1_exact", "c1", ["--res", "64", "64"], False),
    ("c1", "c1", ["--res", "64", "64"], True),
    ("c1_jitter_mitchell", "c1", ["--res", "32", "32", "--spp", "2", "2", "--jitter", "true", "--filter", "mitchell"], True),
    ("c2", "c2", ["--res", "32", "32", "--spp", "2", "2"], True),
    ("c2_one", "c2", ["--res", "32", "32", "--spp", "2", "2", "--strategy", "one"], True),
    ("c2_nsamples4", "c2", ["--res", "24", "24", "--spp", "2", "2", "--nsamples", "4"], True),
    ("c3", "c3", ["--res", "32", "32", "--spp", "2", "2"], True),
    ("c3_mitchell", "c3", ["--res", "24", "24", "--spp", "2", "2", "--filter", "mitchell"], True),
    ("c3_crop", "c3", ["--res", "48", "48", "--spp", "2", "2", "--crop", "0.25", "0.75", "0.5", "1"], True),
    ("c3_lens", "c3", ["--res", "24", "24", "--spp", "3", "3", "--lens", "20", "900"], True),
    ("c3_spherelight", "c3", ["--res", "24", "24", "--spp", "2", "2", "--spherelight"], True),
    ("c3_whitted_mirror", "c3", ["--res", "32", "32", "--spp", "2", "2", "--integrator", "whitted", "--mirror"], True),
    ("c3_direct_mirror", "c3", ["--res", "32", "32", "--spp", "2", "2", "--integrator", "directlighting", "--mirror"], True),
    ("c4_blob", "c4", ["--res", "32", "32", "--spp", "2", "2", "--nphi", "48", "--ntheta", "24"], True),
    ("c4_blob_normals", "c4", ["--res", "32", "32", "--spp", "2", "2", "--nphi", "48", "--ntheta", "24", "--normals"], True),
    # lowdiscrepancy sampler ((0,2)-sequences, samplers/lowdiscrepancy.cpp)
    ("c3_ld", "c3", ["--res", "24", "24", "--spp", "2", "2", "--ld"], True),
    ("c2_ld_nsamples3", "c2", ["--res", "24", "24", "--spp", "4", "2", "--ld", "--nsamples", "3"], True),
    ("c3_ld_lens_mitchell", "c3", ["--res", "16", "16", "--spp", "4", "4", "--ld", "--lens", "20", "900", "--filter", "mitchell"], True),
    ("c3_ld_whitted_mirror", "c3", ["--res", "16", "16", "--spp", "3", "1", "--ld", "--integrator", "whitted", "--mirror"], True),
]


# hand-written scenes (edge cases the generator does not produce): name -> scene text
INLINE = {
    "edge_empty_world": """LookAt 0 0 -5 0 0 0 0 1 0
Camera "perspective" "float fov" [45]
Film "image" "integer xresolution" [8] "integer yresolution" [8] "string filename" ["edge_empty_world_out.exr"]
Sampler "stratified" "integer xsamples" [1] "integer ysamples" [1]
PixelFilter "box"
SurfaceIntegrator "path"
WorldBegin
WorldEnd
""",
    "edge_no_lights": """LookAt 0 0 -5 0 0 0 0 1 0
Camera "perspective" "float fov" [45]
Film "image" "integer xresolution" [12] "integer yresolution" [8] "string filename" ["edge_no_lights_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [1]
PixelFilter "triangle" "float xwidth" [1.5] "float ywidth" [1]
SurfaceIntegrator "directlighting"
WorldBegin
Shape "sphere" "float radius" [1.2] "float zmin" [-0.5] "float zmax" [0.9] "float phimax" [300]
WorldEnd
""",
    "edge_clipped_sphere_point_lights": """LookAt 0 1 -5 0 0 0 0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [24] "integer yresolution" [20] "string filename" ["edge_clipped_sphere_point_lights_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "sinc" "float xwidth" [2] "float ywidth" [2]
SurfaceIntegrator "whitted" "integer maxdepth" [3]
WorldBegin
LightSource "point" "color I" [30 20 10] "point from" [3 4 -4]
LightSource "point" "color I" [5 10 20] "point from" [-3 2 -3]
AttributeBegin
Material "matte" "color Kd" [.6 .6 .6] "float sigma" [30]
Rotate 35 1 0 0
Shape "sphere" "float radius" [1.2] "float zmin" [-0.7] "float zmax" [0.8] "float phimax" [270]
AttributeEnd
AttributeBegin
Material "mirror"
ReverseOrientation
Translate 0 -1.3 0
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-4 0 -4 4 0 -4 4 0 4 -4 0 4]
AttributeEnd
WorldEnd
""",
    "plastic_spot_distant_path": """LookAt 278 273 -800 278 273 0 0 1 0
Camera "perspective" "float fov" [39.3]
Film "image" "integer xresolution" [32] "integer yresolution" [32] "string filename" ["plastic_spot_distant_path_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator "path" "integer maxdepth" [4]
WorldBegin
LightSource "distant" "color L" [2 2 1.5] "point from" [0 600 -300] "point to" [278 0 278]
LightSource "spot" "color I" [400000 300000 300000] "point from" [100 500 100] "point to" [300 0 300] "float coneangle" [35] "float conedeltaangle" [10]
AttributeBegin
AreaLightSource "area" "color L" [17 12 4]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [213 548.7 227 343 548.7 227 343 548.7 332 213 548.7 332]
AttributeEnd
Material "plastic" "color Kd" [.3 .5 .3] "color Ks" [.6 .6 .6] "float roughness" [.05]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 0 0 0 555 555 0 555 555 0 0]
Material "plastic" "color Kd" [.6 .2 .2] "color Ks" [.3 .3 .3] "float roughness" [.3]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 555 0 555 555 555 555 555 555 0 555]
AttributeBegin
Translate 278 120 300
Material "plastic" "color Kd" [.2 .2 .7] "float roughness" [.002]
Shape "sphere" "float radius" [120]
AttributeEnd
WorldEnd
""",
    "plastic_spot_distant_whitted": """LookAt 278 273 -800 278 273 0 0 1 0
Camera "perspective" "float fov" [39.3]
Film "image" "integer xresolution" [32] "integer yresolution" [32] "string filename" ["plastic_spot_distant_whitted_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator "whitted"
WorldBegin
LightSource "distant" "color L" [2 2 1.5] "point from" [0 600 -300] "point to" [278 0 278]
LightSource "spot" "color I" [400000 300000 300000] "point from" [100 500 100] "point to" [300 0 300] "float coneangle" [35] "float conedeltaangle" [10]
AttributeBegin
AreaLightSource "area" "color L" [17 12 4]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [213 548.7 227 343 548.7 227 343 548.7 332 213 548.7 332]
AttributeEnd
Material "plastic" "color Kd" [.3 .5 .3] "color Ks" [.6 .6 .6] "float roughness" [.05]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 0 0 0 555 555 0 555 555 0 0]
Material "plastic" "color Kd" [.6 .2 .2] "color Ks" [.3 .3 .3] "float roughness" [.3]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 555 0 555 555 555 555 555 555 0 555]
AttributeBegin
Translate 278 120 300
Material "plastic" "color Kd" [.2 .2 .7] "float roughness" [.002]
Shape "sphere" "float radius" [120]
AttributeEnd
WorldEnd
""",
    "plastic_spot_distant_direct": """LookAt 278 273 -800 278 273 0 0 1 0
Camera "perspective" "float fov" [39.3]
Film "image" "integer xresolution" [32] "integer yresolution" [32] "string filename" ["plastic_spot_distant_direct_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator "directlighting"
WorldBegin
LightSource "distant" "color L" [2 2 1.5] "point from" [0 600 -300] "point to" [278 0 278]
LightSource "spot" "color I" [400000 300000 300000] "point from" [100 500 100] "point to" [300 0 300] "float coneangle" [35] "float conedeltaangle" [10]
AttributeBegin
AreaLightSource "area" "color L" [17 12 4]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [213 548.7 227 343 548.7 227 343 548.7 332 213 548.7 332]
AttributeEnd
Material "plastic" "color Kd" [.3 .5 .3] "color Ks" [.6 .6 .6] "float roughness" [.05]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 0 0 0 555 555 0 555 555 0 0]
Material "plastic" "color Kd" [.6 .2 .2] "color Ks" [.3 .3 .3] "float roughness" [.3]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 555 0 555 555 555 555 555 555 0 555]
AttributeBegin
Translate 278 120 300
Material "plastic" "color Kd" [.2 .2 .7] "float roughness" [.002]
Shape "sphere" "float radius" [120]
AttributeEnd
WorldEnd
""",
}


def _uber_shiny(integrator, tag):
    """uber (opacity pass-through, Lambertian, microfacet, specular) and shinymetal (conductor Fresnel)"""
    return """LookAt 278 273 -800 278 273 0 0 1 0
Camera "perspective" "float fov" [39.3]
Film "image" "integer xresolution" [32] "integer yresolution" [32] "string filename" ["uber_shiny_%s_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator %s
WorldBegin
LightSource "point" "color I" [150000 150000 120000] "point from" [400 450 -100]
AttributeBegin
AreaLightSource "area" "color L" [17 12 4]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [213 548.7 227 343 548.7 227 343 548.7 332 213 548.7 332]
AttributeEnd
Material "uber" "color Kd" [.4 .5 .3] "color Ks" [.5 .5 .5] "color Kr" [.3 .3 .3] "float roughness" [.08]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 0 0 0 555 555 0 555 555 0 0]
Material "shinymetal" "color Ks" [.9 .7 .3] "color Kr" [.8 .8 .9] "float roughness" [.15]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 555 0 555 555 555 555 555 555 0 555]
Material "matte" "color Kd" [.65 .05 .05]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 0 0 555 0 0 555 555 0 0 555]
AttributeBegin
Translate 200 110 250
Material "uber" "color Kd" [.2 .3 .8] "color Ks" [.2 .2 .2] "color opacity" [.35 .5 .65] "float roughness" [.2]
Shape "sphere" "float radius" [110]
AttributeEnd
AttributeBegin
Translate 400 90 150
Material "shinymetal" "float roughness" [.01]
Shape "sphere" "float radius" [90]
AttributeEnd
WorldEnd
""" % (tag, integrator)


INLINE["uber_shiny_path"] = _uber_shiny('"path" "integer maxdepth" [4]', "path")
INLINE["uber_shiny_whitted"] = _uber_shiny('"whitted" "integer maxdepth" [3]', "whitted")
INLINE["uber_shiny_direct"] = _uber_shiny('"directlighting" "integer maxdepth" [3]', "direct")


def _quadrics(integrator, tag, lights):
    """every quadric (full and clipped by zmin/zmax/phimax, rotated, one reversed), disk and cylinder area lights"""
    return """LookAt 0 3 -9 0 0.5 0 0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [40] "integer yresolution" [28] "string filename" ["quadrics_%s_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator %s
WorldBegin
%s
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-8 -1 -8 8 -1 -8 8 -1 8 -8 -1 8]
AttributeBegin
Material "matte" "color Kd" [.7 .3 .2]
Translate -3 0 0
Rotate -90 1 0 0
Shape "cylinder" "float radius" [.7] "float zmin" [-1] "float zmax" [1.2] "float phimax" [300]
AttributeEnd
AttributeBegin
Material "plastic" "color Kd" [.2 .6 .3] "color Ks" [.4 .4 .4] "float roughness" [.1]
Translate -1.2 -1 0.5
Rotate -90 1 0 0
Shape "cone" "float radius" [.8] "float height" [2]
AttributeEnd
AttributeBegin
Material "glass" "float index" [1.4]
Translate 0.6 1.3 0
Rotate 90 1 0 0
Rotate 25 0 0 1
Shape "paraboloid" "float radius" [.9] "float zmin" [.2] "float zmax" [2] "float phimax" [330]
AttributeEnd
AttributeBegin
Material "mirror" "color Kr" [.8 .8 .8]
Translate 2.6 -1 1
Rotate -90 1 0 0
Shape "hyperboloid" "point p1" [.9 0 0] "point p2" [.4 .6 2.2]
AttributeEnd
AttributeBegin
Material "matte" "color Kd" [.3 .3 .8] "float sigma" [20]
Translate 0 -0.6 -2.5
Rotate -70 1 0 0
ReverseOrientation
Shape "disk" "float radius" [1] "float innerradius" [.35] "float phimax" [270] "float height" [.1]
AttributeEnd
AttributeBegin
Material "matte" "color Kd" [.5 .5 .1]
Translate 3.5 0.2 -1.5
Scale 1 1.4 1
Shape "cone" "float radius" [.5] "float height" [1] "float phimax" [200]
AttributeEnd
WorldEnd
""" % (tag, integrator, lights)


_QL_AREA = """AttributeBegin
AreaLightSource "area" "color L" [14 13 11] "integer nsamples" [2]
Translate 0 5 0
Rotate 90 1 0 0
Shape "disk" "float radius" [1.2]
AttributeEnd
AttributeBegin
AreaLightSource "area" "color L" [6 8 12]
Translate -4 2.5 2
Rotate 90 0 1 0
Shape "cylinder" "float radius" [.25] "float zmin" [-1] "float zmax" [1]
AttributeEnd"""
_QL_POINT = """LightSource "point" "color I" [60 60 50] "point from" [2 6 -5]
LightSource "point" "color I" [20 25 40] "point from" [-5 3 -2]"""
INLINE["quadrics_path"] = _quadrics('"path" "integer maxdepth" [4]', "path", _QL_AREA)
INLINE["quadrics_direct"] = _quadrics('"directlighting" "integer maxdepth" [3]', "direct", _QL_AREA)
INLINE["quadrics_whitted"] = _quadrics('"whitted" "integer maxdepth" [4]', "whitted", _QL_POINT + "\n" + _QL_AREA)


def _infinite(integrator, tag, extra=""):
    """constant infinite light (plus an area light): escaping camera, specular and MIS-probe rays"""
    return """LookAt 0 3 -9 0 0.5 0 0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [36] "integer yresolution" [28] "string filename" ["infinite_%s_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator %s
WorldBegin
LightSource "infinite" "color L" [.7 .8 1.1] "integer nsamples" [2]
%s
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-6 -1 -6 6 -1 -6 6 -1 6 -6 -1 6]
AttributeBegin
Material "glass" "float index" [1.5]
Translate -1.6 0 0
Shape "sphere" "float radius" [1]
AttributeEnd
AttributeBegin
Material "mirror"
Translate 1.6 0 0.5
Shape "sphere" "float radius" [1]
AttributeEnd
AttributeBegin
Material "plastic" "color Kd" [.7 .2 .2] "color Ks" [.3 .3 .3] "float roughness" [.05]
Translate 0 -0.5 -2.2
Shape "sphere" "float radius" [.5]
AttributeEnd
WorldEnd
""" % (tag, integrator, extra)


_INF_AREA = """AttributeBegin
AreaLightSource "area" "color L" [10 9 7]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 4 -1 1 4 -1 1 4 1 -1 4 1]
AttributeEnd"""
INLINE["infinite_path"] = _infinite('"path" "integer maxdepth" [4]', "path", _INF_AREA)
INLINE["infinite_only_path"] = _infinite('"path" "integer maxdepth" [3]', "only_path")
INLINE["infinite_direct"] = _infinite('"directlighting" "integer maxdepth" [3]', "direct", _INF_AREA)
INLINE["infinite_direct_one"] = _infinite('"directlighting" "integer maxdepth" [3] "string strategy" ["one"]', "direct_one", _INF_AREA)
INLINE["infinite_whitted"] = _infinite('"whitted" "integer maxdepth" [3]', "whitted", _INF_AREA)



def _infinite_map(integrator, tag, extra=""):
    """radiance-mapped infinite light ("mapname", lights/infinite.cpp:63-87) under a rotated, scaled light space"""
    return _infinite(integrator, tag, extra).replace(
        'LightSource "infinite" "color L" [.7 .8 1.1] "integer nsamples" [2]',
        'AttributeBegin\nRotate -90 1 0 0\nRotate 40 0 0 1\nScale 1 1.5 -1\n'
        'LightSource "infinite" "color L" [1.4 1.3 1.2] "integer nsamples" [2] "string mapname" ["tex8x16.exr"]\nAttributeEnd')


def _infinite_is(integrator, tag, extra=""):
    """importance-sampled radiance map (lights/infinitesample.cpp)"""
    return _infinite_map(integrator, tag, extra).replace('LightSource "infinite"', 'LightSource "infinitesample"')


INLINE["infinite_is_path"] = _infinite_is('"path" "integer maxdepth" [4]', "is_path", _INF_AREA)
INLINE["infinite_is_direct"] = _infinite_is('"directlighting" "integer maxdepth" [3]', "is_direct")
INLINE["infinite_is_whitted"] = _infinite_is('"whitted" "integer maxdepth" [3]', "is_whitted", _INF_AREA)
INLINE["infinite_map_path"] = _infinite_map('"path" "integer maxdepth" [4]', "map_path", _INF_AREA)
INLINE["infinite_map_direct"] = _infinite_map('"directlighting" "integer maxdepth" [3]', "map_direct")
INLINE["infinite_map_direct_one"] = _infinite_map('"directlighting" "integer maxdepth" [3] "string strategy" ["one"]', "map_direct_one", _INF_AREA)
INLINE["infinite_map_whitted"] = _infinite_map('"whitted" "integer maxdepth" [3]', "map_whitted", _INF_AREA)


def _latlong_mesh(nphi, ntheta, rx, ry, rz, normals=True):
    """a small closed lat-long mesh (optionally with per-vertex N and uv) as a trianglemesh statement"""
    import math
    P, N, UV, idx = [], [], [], []
    for j in range(ntheta + 1):
        th = math.pi * j / ntheta
        for i in range(nphi + 1):
            ph = 2 * math.pi * i / nphi
            x, y, z = math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)
            P += [rx * x, ry * y, rz * z]
            N += [x / rx, y / ry, z / rz]
            UV += [i / nphi, j / ntheta]
    for j in range(ntheta):
        for i in range(nphi):
            a = j * (nphi + 1) + i
            b = a + nphi + 1
            idx += [a, b, a + 1, a + 1, b, b + 1]
    f = lambda v: " ".join("%.6g" % t for t in v)
    out = 'Shape "trianglemesh" "integer indices" [%s] "point P" [%s]' % (" ".join(map(str, idx)), f(P))
    if normals:
        out += ' "normal N" [%s] "float uv" [%s]' % (f(N), f(UV))
    return out


def _instances(integrator, tag):
    """object instancing: a multi-primitive object (own kd-tree, mesh with shading normals + quadric), a
    single-sphere object (no accelerator) and a two-triangle object, placed with translations,
    rotations, non-uniform scales and a mirroring transform"""
    return """LookAt 0 4 -11 0 0.8 0 0 1 0
Camera "perspective" "float fov" [38]
Film "image" "integer xresolution" [44] "integer yresolution" [30] "string filename" ["instances_%s_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator %s
WorldBegin
LightSource "point" "color I" [90 85 80] "point from" [3 7 -6]
AttributeBegin
AreaLightSource "area" "color L" [9 9 10]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-2 6 -1 2 6 -1 2 6 2 -2 6 2]
AttributeEnd
ObjectBegin "thing"
Material "plastic" "color Kd" [.7 .4 .2] "color Ks" [.3 .3 .3] "float roughness" [.1]
%s
Material "mirror" "color Kr" [.8 .8 .8]
AttributeBegin
Translate 0 1.4 0
Shape "sphere" "float radius" [.45]
AttributeEnd
Material "matte" "color Kd" [.2 .5 .7]
AttributeBegin
Translate 0 -1.1 0
Rotate -90 1 0 0
Shape "cylinder" "float radius" [.35] "float zmin" [0] "float zmax" [.6]
AttributeEnd
ObjectEnd
ObjectBegin "ball"
Material "glass" "float index" [1.45]
Shape "sphere" "float radius" [.6]
ObjectEnd
ObjectBegin "card"
Material "matte" "color Kd" [.8 .8 .3]
%s
ObjectEnd
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9]
AttributeBegin
Translate -3.2 0.6 0
ObjectInstance "thing"
AttributeEnd
AttributeBegin
Translate 0 0.9 1
Rotate 40 0 1 0
Scale 1.5 0.8 1.1
ObjectInstance "thing"
AttributeEnd
AttributeBegin
Translate 3.3 0.7 -0.5
Rotate 25 0 0 1
Scale -1 1 1
ObjectInstance "thing"
AttributeEnd
AttributeBegin
Translate -1.5 -0.4 -2.5
ObjectInstance "ball"
AttributeEnd
AttributeBegin
Translate 1.6 -0.2 -2.8
Scale 1 1.5 0.7
ObjectInstance "ball"
AttributeEnd
AttributeBegin
Translate 0 2.6 3
Rotate 20 1 0 0
ObjectInstance "card"
Translate 2.4 -0.5 0
Rotate 50 0 1 0
ObjectInstance "card"
AttributeEnd
WorldEnd
""" % (tag, integrator, _latlong_mesh(8, 5, .9, .7, .8), 'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 -1 0 1 -1 0 1 1 0 -1 1 0]')


INLINE["instances_path"] = _instances('"path" "integer maxdepth" [4]', "path")
INLINE["instances_whitted"] = _instances('"whitted" "integer maxdepth" [4]', "whitted")
INLINE["instances_direct"] = _instances('"directlighting" "integer maxdepth" [3]', "direct")


def _instanced_emitters(integrator, tag):
    """area-lit shapes inside ObjectBegin: the reference keeps the AreaLight on the primitive (it emits when a
    camera or specular ray hits it) but never adds it to the light vector (api.cpp:383-387)"""
    return """LookAt 0 3 -9 0 0.8 0 0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [40] "integer yresolution" [28] "string filename" ["instances_emitter_%s_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "mitchell"
SurfaceIntegrator %s
WorldBegin
LightSource "point" "color I" [40 40 45] "point from" [-3 6 -5]
ObjectBegin "lamp"
AttributeBegin
AreaLightSource "area" "color L" [6 5 3] "integer nsamples" [2]
Material "matte" "color Kd" [.5 .5 .5]
Shape "sphere" "float radius" [.5]
Translate 0 1.2 0
Rotate 30 1 0 0
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-.6 -.6 0 .6 -.6 0 .6 .6 0 -.6 .6 0]
AttributeEnd
Material "plastic" "color Kd" [.2 .6 .3] "color Ks" [.4 .4 .4] "float roughness" [.05]
AttributeBegin
Translate 0 -0.9 0
Rotate -90 1 0 0
Shape "cone" "float radius" [.5] "float height" [.6]
AttributeEnd
ObjectEnd
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9]
AttributeBegin
Material "mirror" "color Kr" [.9 .9 .9]
Translate 0 0.4 2.5
Shape "sphere" "float radius" [1.4]
AttributeEnd
AttributeBegin
Material "glass" "float index" [1.5]
Translate 0.2 -0.3 -3
Shape "sphere" "float radius" [.6]
AttributeEnd
AttributeBegin
Translate -2.6 0.5 0
ObjectInstance "lamp"
AttributeEnd
AttributeBegin
Translate 2.5 0.7 -0.5
Rotate 35 0 0 1
Scale -1.2 1 0.8
ObjectInstance "lamp"
AttributeEnd
WorldEnd
""" % (tag, integrator)


INLINE["instances_emitter_path"] = _instanced_emitters('"path" "integer maxdepth" [4]', "path")
INLINE["instances_emitter_whitted"] = _instanced_emitters('"whitted" "integer maxdepth" [4]', "whitted")
INLINE["instances_emitter_direct"] = _instanced_emitters('"directlighting" "integer maxdepth" [3]', "direct")


def make_textures():
    """image maps of the tex_* fixtures (committed beside them): non-power-of-two, non-square, tiny"""
    rng = np.random.RandomState(20260117)
    a = rng.rand(7, 13, 3).astype(np.float32)
    a[::2, ::3] *= 0.1
    write_exr(os.path.join(HERE, "tex13x7.exr"), a)
    b = rng.rand(16, 8, 3).astype(np.float32)
    yy, xx = np.mgrid[0:16, 0:8]
    b *= (((xx // 2 + yy // 4) % 2) * 0.8 + 0.2)[..., None].astype(np.float32)
    write_exr(os.path.join(HERE, "tex8x16.exr"), b)
    c = (rng.rand(5, 5, 3) * 1.5).astype(np.float32)
    write_exr(os.path.join(HERE, "tex5x5.exr"), c)


_TEX_HEAD = """LookAt 0 3 -8 0 0.5 0 0 1 0
Camera %s
Film "image" "integer xresolution" [40] "integer yresolution" [28] %s "string filename" ["%s_out.exr"]
%s
PixelFilter %s
SurfaceIntegrator %s
WorldBegin
"""

_TEX_LIGHTS = """LightSource "point" "color I" [50 50 50] "point from" [2 6 -5]
AttributeBegin
AreaLightSource "area" "color L" [8 8 7] "integer nsamples" [1]
Translate -2 5 0
Rotate 90 1 0 0
Shape "disk" "radius" [1.2]
AttributeEnd
"""


def _tex_scene(tag, body, camera='"perspective" "float fov" [40]', film="", sampler='Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]',
               pfilter='"box"', depth=3, integrator=None):
    return _TEX_HEAD % (camera, film, tag, sampler, pfilter, integrator or '"path" "integer maxdepth" [%d]' % depth) + _TEX_LIGHTS.replace('"radius"', '"float radius"') + body + "WorldEnd\n"


# uv mapping with scale/offset on a plane (EWA, strongly anisotropic towards the horizon), a pow2 map on a
# sphere, a glass ball and a mirror so textured surfaces are also reached by rays without differentials
INLINE["tex_uv_ewa_path"] = _tex_scene("tex_uv_ewa_path", """Texture "img" "color" "imagemap" "string filename" ["tex13x7.exr"] "float uscale" [4] "float vscale" [3] "float udelta" [.25] "float vdelta" [-.4]
Material "matte" "texture Kd" "img"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9] "float uv" [0 0 1 0 1 1 0 1]
AttributeBegin
Texture "img2" "color" "imagemap" "string filename" ["tex8x16.exr"]
Material "matte" "texture Kd" "img2"
Translate -1.2 0.5 0.5
Shape "sphere" "float radius" [1.5]
AttributeEnd
AttributeBegin
Material "glass" "float index" [1.5]
Translate 1.6 0 -2
Shape "sphere" "float radius" [.9]
AttributeEnd
AttributeBegin
Material "mirror"
Translate 3 1 2
Rotate -30 0 1 0
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1.5 -2 0 1.5 -2 0 1.5 2 0 -1.5 2 0]
AttributeEnd
""", depth=4)

# the three other mappings (under rotated / scaled texture spaces), trilinear filtering, the three wrap modes
INLINE["tex_mappings_path"] = _tex_scene("tex_mappings_path", """TransformBegin
Rotate 25 1 0 1
Scale 1 1.3 .8
Texture "sph" "color" "imagemap" "string filename" ["tex8x16.exr"] "string mapping" ["spherical"] "string wrap" ["clamp"]
TransformEnd
TransformBegin
Translate .5 0 0
Rotate 90 1 0 0
Texture "cyl" "color" "imagemap" "string filename" ["tex13x7.exr"] "string mapping" ["cylindrical"] "bool trilinear" ["true"]
Texture "cyl2" "color" "imagemap" "string filename" ["tex8x16.exr"] "string mapping" ["cylindrical"] "float maxanisotropy" [16]
TransformEnd
Texture "pla" "color" "imagemap" "string filename" ["tex5x5.exr"] "string mapping" ["planar"] "vector v1" [.3 0 .1] "vector v2" [0 .1 .35] "float udelta" [.1] "float vdelta" [.2] "string wrap" ["black"] "float maxanisotropy" [3]
Material "matte" "texture Kd" "pla"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9]
AttributeBegin
Material "plastic" "texture Kd" "sph" "color Ks" [.2 .2 .2] "float roughness" [.05]
Translate -1.8 0.5 0
Shape "sphere" "float radius" [1.4]
AttributeEnd
AttributeBegin
Material "matte" "texture Kd" "cyl"
Translate 1.8 -1 0.5
Rotate -90 1 0 0
Shape "cylinder" "float radius" [1] "float zmin" [0] "float zmax" [2.5]
AttributeEnd
AttributeBegin
Material "matte" "texture Kd" "cyl2"
Translate 0.2 -0.4 -3
Shape "sphere" "float radius" [.6]
AttributeEnd
""")

# float image maps (sigma, roughness, mix amount), scale / mix expressions over image maps, opacity, and a
# texture on every operand kind
INLINE["tex_float_mix_path"] = _tex_scene("tex_float_mix_path", """Texture "a" "color" "imagemap" "string filename" ["tex13x7.exr"] "float uscale" [2] "float vscale" [2]
Texture "b" "color" "imagemap" "string filename" ["tex8x16.exr"] "string wrap" ["clamp"]
Texture "fa" "float" "imagemap" "string filename" ["tex5x5.exr"]
Texture "fb" "float" "imagemap" "string filename" ["tex8x16.exr"] "bool trilinear" ["true"] "float uscale" [3] "float vscale" [3]
Texture "deg" "float" "scale" "texture tex1" "fa" "float tex2" [60]
Texture "rough" "float" "mix" "float tex1" [.02] "float tex2" [.4] "texture amount" "fb"
Texture "ab" "color" "scale" "texture tex1" "a" "texture tex2" "b"
Texture "mixed" "color" "mix" "texture tex1" "ab" "color tex2" [.9 .2 .1] "texture amount" "fa"
Material "matte" "texture Kd" "mixed" "texture sigma" "deg"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9] "float uv" [0 0 3 0 3 3 0 3]
AttributeBegin
Material "plastic" "texture Kd" "a" "texture Ks" "b" "texture roughness" "rough"
Translate -2.2 0.4 0
Shape "sphere" "float radius" [1.3]
AttributeEnd
AttributeBegin
Material "uber" "texture Kd" "b" "color Ks" [.3 .3 .3] "texture Kr" "a" "texture opacity" "b" "texture roughness" "rough"
Translate 0.6 0.2 -1.5
Shape "sphere" "float radius" [1]
AttributeEnd
AttributeBegin
Material "substrate" "texture Kd" "a" "texture Ks" "ab" "texture uroughness" "rough" "float vroughness" [.2]
Translate 2.8 0.3 1
Shape "sphere" "float radius" [1.2]
AttributeEnd
AttributeBegin
Material "glass" "texture Kr" "b" "texture Kt" "a" "texture index" "fa"
Translate -0.5 -0.3 -3.5
Shape "sphere" "float radius" [.6]
AttributeEnd
""", depth=4)

# a textured mesh with shading normals inside instanced objects, the lowdiscrepancy sampler, depth of field,
# a Mitchell filter and a crop window
INLINE["tex_instances_ld_path"] = _tex_scene("tex_instances_ld_path", """Texture "img" "color" "imagemap" "string filename" ["tex13x7.exr"]
Texture "chk" "color" "imagemap" "string filename" ["tex8x16.exr"] "float uscale" [6] "float vscale" [6]
Material "shinymetal" "texture Ks" "chk" "texture Kr" "img" "float roughness" [.1]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9] "float uv" [0 0 1 0 1 1 0 1]
ObjectBegin "blob"
Material "matte" "texture Kd" "img"
%s
Material "translucent" "texture Kd" "chk" "color Ks" [.2 .2 .2] "texture reflect" "img" "color transmit" [.6 .6 .6]
AttributeBegin
Translate 0 1.5 0
Shape "disk" "float radius" [.8]
AttributeEnd
ObjectEnd
AttributeBegin
Translate -2 0.3 0
ObjectInstance "blob"
AttributeEnd
AttributeBegin
Translate 2 0.5 1
Rotate 40 0 1 1
Scale -1.3 .9 1.1
ObjectInstance "blob"
AttributeEnd
""" % _latlong_mesh(8, 5, 1.0, .8, .9), camera='"perspective" "float fov" [40] "float lensradius" [.15] "float focaldistance" [8]',
    film='"float cropwindow" [.1 .9 .15 .95]', sampler='Sampler "lowdiscrepancy" "integer pixelsamples" [4]', pfilter='"mitchell"')

# bestcandidate sampler + orthographic camera; environment camera + random sampler
_TEX_SMALL = """Texture "img" "color" "imagemap" "string filename" ["tex5x5.exr"] "float uscale" [5] "float vscale" [5] "string wrap" ["clamp"]
Material "matte" "texture Kd" "img"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9] "float uv" [0 0 1 0 1 1 0 1]
AttributeBegin
Texture "s" "color" "imagemap" "string filename" ["tex13x7.exr"] "string mapping" ["spherical"]
Material "plastic" "texture Kd" "s" "color Ks" [.3 .3 .3]
Translate 0 0.6 0
Shape "sphere" "float radius" [1.6]
AttributeEnd
"""
INLINE["tex_bc_ortho_path"] = _tex_scene("tex_bc_ortho_path", _TEX_SMALL, camera='"orthographic" "float screenwindow" [-4 4 -2.8 2.8]',
                                         sampler='Sampler "bestcandidate" "integer pixelsamples" [3]', pfilter='"gaussian"')
INLINE["tex_env_random_path"] = _tex_scene("tex_env_random_path", _TEX_SMALL, camera='"environment"',
                                           sampler='Sampler "random" "integer xsamples" [2] "integer ysamples" [1]')


# Whitted / DirectLighting: ray differentials carried through specular reflection and refraction
# (whitted.cpp:86-131): curved mirrors and glass (dndu, dndv of quadrics), a mirror mesh with shading
# normals, an instanced mirror, nested transmission, all looking at image-mapped surfaces
_TEX_SPECULAR = """Texture "img" "color" "imagemap" "string filename" ["tex13x7.exr"] "float uscale" [6] "float vscale" [6]
Texture "chk" "color" "imagemap" "string filename" ["tex8x16.exr"] "float uscale" [3] "float vscale" [2] %s
Texture "fr" "float" "imagemap" "string filename" ["tex5x5.exr"]
Material "matte" "texture Kd" "img"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9] "float uv" [0 0 1 0 1 1 0 1]
Material "plastic" "texture Kd" "chk" "color Ks" [.1 .1 .1]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 6 9 -1 6 9 8 6 -9 8 6] "float uv" [0 0 1 0 1 1 0 1]
AttributeBegin
Material "mirror" "texture Kr" "chk"
Translate -2.3 0.6 0.5
Shape "sphere" "float radius" [1.5]
AttributeEnd
AttributeBegin
Material "glass" "float index" [1.4] "color Kr" [.7 .7 .7]
Translate 1.2 0.1 -2.2
Scale 1 1.2 .8
Shape "sphere" "float radius" [1]
AttributeEnd
AttributeBegin
Material "uber" "texture Kd" "img" "color Ks" [.1 .1 .1] "color Kr" [.6 .6 .6] "texture roughness" "fr"
Translate 3 -1 1.5
Rotate -90 1 0 0
Shape "cylinder" "float radius" [.9] "float zmin" [0] "float zmax" [2.6]
AttributeEnd
ObjectBegin "m"
Material "mirror"
%s
ObjectEnd
AttributeBegin
Translate 0.3 1.2 3.5
Rotate 20 0 1 0
Scale 1.2 1 -1
ObjectInstance "m"
AttributeEnd
ObjectBegin "c"
AttributeBegin
Material "shinymetal" "color Ks" [.7 .5 .9] "color Kr" [.9 .5 .4] "float roughness" [.1]
Rotate -70 1 0 0
Shape "cone" "float radius" [.8] "float height" [1.2]
AttributeEnd
ObjectEnd
AttributeBegin
Translate 2.2 0.2 -2.6
Rotate 30 0 0 1
Scale -1.1 1.4 .8
ObjectInstance "c"
AttributeEnd
AttributeBegin
Material "glass" "texture Kt" "chk" "float index" [1.6]
Translate -0.6 -0.6 -3.6
Rotate 30 1 1 0
Shape "paraboloid" "float radius" [.5] "float zmax" [.8]
AttributeEnd
"""
INLINE["tex_whitted"] = _tex_scene("tex_whitted", _TEX_SPECULAR % ("", _latlong_mesh(8, 5, 1.1, .9, 1.0)), integrator='"whitted" "integer maxdepth" [4]')
INLINE["tex_direct_all"] = _tex_scene("tex_direct_all", _TEX_SPECULAR % ('"bool trilinear" ["true"]', _latlong_mesh(8, 5, 1.1, .9, 1.0)),
                                      integrator='"directlighting" "integer maxdepth" [3] "string strategy" ["all"]', pfilter='"mitchell"')
INLINE["tex_direct_one_dof"] = _tex_scene("tex_direct_one_dof", _TEX_SPECULAR % ('"string wrap" ["clamp"]', _latlong_mesh(8, 5, 1.1, .9, 1.0)),
                                          camera='"perspective" "float fov" [40] "float lensradius" [.1] "float focaldistance" [8]',
                                          sampler='Sampler "lowdiscrepancy" "integer pixelsamples" [4]',
                                          integrator='"directlighting" "integer maxdepth" [4] "string strategy" ["one"]')


# bump mapping (Material::Bump, core/material.cpp:29-71): image-mapped, expression and constant non-zero
# displacements on a plane, quadrics, a mesh with shading normals and an instanced object, seen directly, in a
# mirror (Whitted: the bumped shading normal drives the reflected differentials) and after diffuse bounces
_TEX_BUMP = """Texture "img" "color" "imagemap" "string filename" ["tex13x7.exr"] "float uscale" [4] "float vscale" [4]
Texture "h1" "float" "imagemap" "string filename" ["tex8x16.exr"] "float uscale" [5] "float vscale" [5]
Texture "h2" "float" "imagemap" "string filename" ["tex5x5.exr"] "string mapping" ["spherical"] "bool trilinear" ["true"]
Texture "h3" "float" "scale" "texture tex1" "h1" "float tex2" [.15]
Texture "h4" "float" "mix" "float tex1" [-.05] "float tex2" [.1] "texture amount" "h2"
Material "matte" "texture Kd" "img" "texture bumpmap" "h3"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9] "float uv" [0 0 1 0 1 1 0 1]
AttributeBegin
Material "plastic" "color Kd" [.7 .3 .2] "color Ks" [.4 .4 .4] "float roughness" [.05] "texture bumpmap" "h4"
Translate -2.2 0.5 0
Shape "sphere" "float radius" [1.4]
AttributeEnd
AttributeBegin
Material "mirror" "texture bumpmap" "h3"
Translate 1.2 0.2 1.5
Rotate -90 1 0 0
Shape "cylinder" "float radius" [1.1] "float zmin" [-1.2] "float zmax" [1.6]
AttributeEnd
AttributeBegin
Material "matte" "color Kd" [.3 .5 .7] "float bumpmap" [.2]
Translate 0.2 -0.4 -3
Shape "sphere" "float radius" [.6]
AttributeEnd
ObjectBegin "b"
Material "uber" "color Kd" [.5 .5 .2] "color Ks" [.2 .2 .2] "color Kr" [.3 .3 .3] "texture bumpmap" "h1"
%s
ObjectEnd
AttributeBegin
Translate 3.2 0.4 -1
Rotate 35 0 1 1
Scale .9 1.2 -.8
ObjectInstance "b"
AttributeEnd
"""
INLINE["tex_bump_path"] = _tex_scene("tex_bump_path", _TEX_BUMP % _latlong_mesh(8, 5, .9, .8, .9), depth=3)
INLINE["tex_bump_whitted"] = _tex_scene("tex_bump_whitted", _TEX_BUMP % _latlong_mesh(8, 5, .9, .8, .9), integrator='"whitted" "integer maxdepth" [3]')
INLINE["tex_bump_direct"] = _tex_scene("tex_bump_direct", _TEX_BUMP % _latlong_mesh(8, 5, .9, .8, .9), integrator='"directlighting" "integer maxdepth" [3]',
                                       sampler='Sampler "bestcandidate" "integer pixelsamples" [4]')


# the procedural textures: uv, bilerp, checkerboard (closed-form and point-sampled, nested), dots, fbm, wrinkled,
# windy, marble -- colour and float versions, every 2D mapping, rotated / scaled texture spaces for the 3D ones
_TEX_PROC = """Texture "chk" "color" "checkerboard" "float uscale" [8] "float vscale" [8] "color tex1" [.9 .9 .9] "color tex2" [.1 .1 .4]
Texture "chk2" "color" "checkerboard" "string aamode" ["none"] "float uscale" [3] "float vscale" [5] "texture tex1" "chk" "color tex2" [.7 .2 .1]
Texture "uvt" "color" "uv" "float uscale" [3] "float vscale" [2]
Texture "bil" "color" "bilerp" "color v00" [.1 .2 .9] "color v01" [.9 .8 .1] "color v10" [.2 .9 .3] "color v11" [.9 .1 .6] "string mapping" ["spherical"]
Texture "fbil" "float" "bilerp" "float v00" [.02] "float v01" [.4] "float v10" [.1] "float v11" [.3]
Texture "dot" "color" "dots" "float uscale" [10] "float vscale" [10] "texture inside" "uvt" "color outside" [.8 .8 .2]
Texture "fdot" "float" "dots" "float uscale" [6] "float vscale" [6] "float inside" [.05] "float outside" [.5]
TransformBegin
Rotate 30 1 1 0
Scale 2 3 1.5
Texture "fb" "color" "fbm" "integer octaves" [5] "float roughness" [.6]
Texture "ffb" "float" "fbm"
Texture "wr" "color" "wrinkled" "integer octaves" [6]
Texture "fwr" "float" "wrinkled" "float roughness" [.4]
Texture "wi" "float" "windy"
Texture "mar" "color" "marble" "float scale" [2] "float variation" [.4]
TransformEnd
Texture "fchk" "float" "checkerboard" "string mapping" ["planar"] "vector v1" [.5 0 0] "vector v2" [0 0 .5] "float tex1" [.02] "texture tex2" "fbil"
Texture "bump" "float" "scale" "texture tex1" "fwr" "float tex2" [.2]
Texture "wibump" "float" "scale" "texture tex1" "wi" "float tex2" [.3]
Material "matte" "texture Kd" "chk2"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9] "float uv" [0 0 1 0 1 1 0 1]
Material "matte" "texture Kd" "mar" "texture bumpmap" "wibump"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 6 9 -1 6 9 8 6 -9 8 6] "float uv" [0 0 1 0 1 1 0 1]
AttributeBegin
Material "plastic" "texture Kd" "bil" "color Ks" [.3 .3 .3] "texture roughness" "fchk"
Translate -2.4 0.5 0
Shape "sphere" "float radius" [1.4]
AttributeEnd
AttributeBegin
Material "matte" "texture Kd" "dot" "texture bumpmap" "bump"
Translate 0.6 0 1
Rotate -90 1 0 0
Shape "cylinder" "float radius" [1] "float zmin" [-1] "float zmax" [1.8]
AttributeEnd
AttributeBegin
Material "uber" "texture Kd" "fb" "texture Ks" "wr" "color Kr" [.2 .2 .2] "texture roughness" "fdot"
Translate 3 0.2 -0.5
Shape "sphere" "float radius" [1.1]
AttributeEnd
AttributeBegin
Material "matte" "texture Kd" "wr" "texture sigma" "ffb"
Translate -0.4 -0.4 -3.2
Shape "disk" "float radius" [.9]
AttributeEnd
"""
INLINE["tex_procedural_path"] = _tex_scene("tex_procedural_path", _TEX_PROC, depth=3)
INLINE["tex_procedural_whitted"] = _tex_scene("tex_procedural_whitted", _TEX_PROC, integrator='"whitted" "integer maxdepth" [3]',
                                              camera='"orthographic" "float screenwindow" [-4 4 -2.8 2.8]')
INLINE["tex_procedural_direct"] = _tex_scene("tex_procedural_direct", _TEX_PROC, integrator='"directlighting" "integer maxdepth" [2] "string strategy" ["one"]',
                                             sampler='Sampler "lowdiscrepancy" "integer pixelsamples" [2]', pfilter='"gaussian"')


def _mapped_lights(integrator, tag):
    """goniometric and projection lights (lights/goniometric.cpp, lights/projection.cpp): image-scaled point lights
    under rotated light spaces, one goniometric light without a map; wide and tall projection images"""
    return """LookAt 0 3 -9 0 0.5 0 0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [36] "integer yresolution" [28] "string filename" ["lights_%s_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator %s
WorldBegin
AttributeBegin
Translate -2 4 -1
Rotate 35 1 0 1
LightSource "goniometric" "color I" [40 38 30] "string mapname" ["tex8x16.exr"]
AttributeEnd
AttributeBegin
Translate 3 2 -3
LightSource "goniometric" "color I" [6 8 10]
AttributeEnd
AttributeBegin
Translate 0 5 -2
Rotate 100 1 0 0
Rotate 20 0 0 1
LightSource "projection" "color I" [160 160 160] "float fov" [50] "string mapname" ["tex13x7.exr"]
AttributeEnd
AttributeBegin
Translate 2.5 3 1
Rotate 60 1 0 0
LightSource "projection" "color I" [60 40 40] "float fov" [70] "string mapname" ["tex8x16.exr"]
AttributeEnd
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-6 -1 -6 6 -1 -6 6 -1 6 -6 -1 6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-6 -1 5 6 -1 5 6 7 5 -6 7 5]
AttributeBegin
Material "glass" "float index" [1.5]
Translate -1.6 0 0
Shape "sphere" "float radius" [1]
AttributeEnd
AttributeBegin
Material "plastic" "color Kd" [.7 .2 .2] "color Ks" [.3 .3 .3] "float roughness" [.05]
Translate 1.4 -0.2 0.5
Shape "sphere" "float radius" [.8]
AttributeEnd
WorldEnd
""" % (tag, integrator)


INLINE["lights_mapped_path"] = _mapped_lights('"path" "integer maxdepth" [3]', "mapped_path")
INLINE["lights_mapped_direct"] = _mapped_lights('"directlighting" "integer maxdepth" [3]', "mapped_direct")
INLINE["lights_mapped_whitted"] = _mapped_lights('"whitted" "integer maxdepth" [3]', "mapped_whitted")


def _measured(integrator, tag):
    """the six measured (Lafortune) materials, one of them bump-mapped"""
    mats = ["bluepaint", "brushedmetal", "clay", "felt", "primer", "skin"]
    body = ""
    for i, m in enumerate(mats):
        body += 'AttributeBegin\nMaterial "%s"%s\nTranslate %g %g %g\nShape "sphere" "float radius" [.8]\nAttributeEnd\n' % (
            m, ' "float bumpmap" [.15]' if m == "clay" else "", -3.2 + 2.1 * (i % 3) + 1.1, 0.1 + 1.9 * (i // 3), (i // 3) * 1.5 - 1)
    return """LookAt 0 3 -9 0 0.8 0 0 1 0
Camera "perspective" "float fov" [40]
Film "image" "integer xresolution" [36] "integer yresolution" [28] "string filename" ["measured_%s_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator %s
WorldBegin
LightSource "point" "color I" [70 70 70] "point from" [3 7 -6]
AttributeBegin
AreaLightSource "area" "color L" [9 9 9]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-2 6 -1 2 6 -1 2 6 2 -2 6 2]
AttributeEnd
Material "brushedmetal"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9]
%sWorldEnd
""" % (tag, integrator, body)


INLINE["measured_path"] = _measured('"path" "integer maxdepth" [4]', "path")
INLINE["measured_direct"] = _measured('"directlighting" "integer maxdepth" [3]', "direct")
INLINE["measured_whitted"] = _measured('"whitted" "integer maxdepth" [3]', "whitted")


def _bestcandidate(integrator, tag, sampler, film='"integer xresolution" [28] "integer yresolution" [20]', lens="", nsamples=1, pfilter='"box"'):
    """bestcandidate sampler (the reference's default when a scene names none)"""
    return """LookAt 278 273 -800 278 273 0 0 1 0
Camera "perspective" "float fov" [39.3] %s
Film "image" %s "string filename" ["bc_%s_out.exr"]
%s
PixelFilter %s
SurfaceIntegrator %s
WorldBegin
AttributeBegin
AreaLightSource "area" "color L" [17 12 4] "integer nsamples" [%d]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [213 548.7 227 343 548.7 227 343 548.7 332 213 548.7 332]
AttributeEnd
Material "matte" "color Kd" [.73 .73 .73]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 0 0 0 555 555 0 555 555 0 0]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 555 0 555 555 555 555 555 555 0 555]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 555 0 555 555 0 555 555 555 0 555 555]
Material "matte" "color Kd" [.12 .45 .15]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 0 0 555 0 0 555 555 0 0 555]
Material "matte" "color Kd" [.65 .05 .05]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [555 0 0 555 0 555 555 555 555 555 555 0]
AttributeBegin
Translate 200 100 250
Material "glass" "float index" [1.5]
Shape "sphere" "float radius" [100]
AttributeEnd
AttributeBegin
Translate 400 80 150
Material "mirror"
Shape "sphere" "float radius" [80]
AttributeEnd
WorldEnd
""" % (lens, film, tag, sampler, pfilter, integrator, nsamples)


INLINE["bc_default_path"] = _bestcandidate('"path" "integer maxdepth" [4]', "default_path", "")   # no Sampler statement at all
INLINE["bc_path_9spp_mitchell"] = _bestcandidate('"path" "integer maxdepth" [3]', "path_9spp_mitchell", 'Sampler "bestcandidate" "integer pixelsamples" [9]',
                                                 pfilter='"mitchell"')
INLINE["bc_direct_nsamples5"] = _bestcandidate('"directlighting"', "direct_nsamples5", 'Sampler "bestcandidate" "integer pixelsamples" [2]', nsamples=5)
INLINE["bc_whitted_lens_crop"] = _bestcandidate('"whitted" "integer maxdepth" [3]', "whitted_lens_crop", 'Sampler "bestcandidate" "integer pixelsamples" [3]',
                                                film='"integer xresolution" [40] "integer yresolution" [32] "float cropwindow" [.2 .85 .3 .9]',
                                                lens='"float lensradius" [15] "float focaldistance" [900]')


def _heightfield():
    import math
    nu, nv = 9, 7
    pz = " ".join("%.6g" % (0.15 * math.sin(1.7 * x) * math.cos(1.3 * y) + 0.05 * ((x * y) % 3)) for y in range(nv) for x in range(nu))
    return """LookAt 0.5 -1.2 1.4 0.5 0.5 0 0 0 1
Camera "perspective" "float fov" [45]
Film "image" "integer xresolution" [36] "integer yresolution" [28] "string filename" ["heightfield_path_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator "path" "integer maxdepth" [3]
WorldBegin
LightSource "point" "color I" [6 6 5] "point from" [0.2 -0.5 2]
AttributeBegin
AreaLightSource "area" "color L" [5 5 6]
Translate 0.5 0.5 1.6
Shape "disk" "float radius" [.4]
AttributeEnd
Material "plastic" "color Kd" [.3 .6 .3] "color Ks" [.2 .2 .2] "float roughness" [.2]
AttributeBegin
Scale 1 1 1.5
Shape "heightfield" "integer nu" [%d] "integer nv" [%d] "float Pz" [%s]
AttributeEnd
Material "matte" "color Kd" [.5 .5 .5]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-3 -3 -0.4 4 -3 -0.4 4 4 -0.4 -3 4 -0.4]
WorldEnd
""" % (nu, nv, pz)


INLINE["heightfield_path"] = _heightfield()


def _camera(cam, tag, extra=""):
    """orthographic (with depth of field) and environment cameras"""
    return """LookAt 278 273 -300 278 273 0 0 1 0
Camera %s
Film "image" "integer xresolution" [32] "integer yresolution" [24] "string filename" ["camera_%s_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator "path" "integer maxdepth" [3]
WorldBegin
AttributeBegin
AreaLightSource "area" "color L" [17 12 4]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [213 548.7 227 343 548.7 227 343 548.7 332 213 548.7 332]
AttributeEnd
Material "matte" "color Kd" [.73 .73 .73]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 0 0 0 555 555 0 555 555 0 0]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 555 0 555 555 555 555 555 555 0 555]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 555 0 555 555 0 555 555 555 0 555 555]
Material "matte" "color Kd" [.12 .45 .15]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 0 0 555 0 0 555 555 0 0 555]
Material "matte" "color Kd" [.65 .05 .05]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [555 0 0 555 0 555 555 555 555 555 555 0]
AttributeBegin
Translate 200 100 250
Material "glass" "float index" [1.5]
Shape "sphere" "float radius" [100]
AttributeEnd
%s
WorldEnd
""" % (cam, tag, extra)


INLINE["camera_orthographic"] = _camera('"orthographic" "float screenwindow" [-300 300 -225 225]', "orthographic")
INLINE["camera_orthographic_lens"] = _camera('"orthographic" "float screenwindow" [-300 300 -225 225] "float lensradius" [12] "float focaldistance" [550]', "orthographic_lens")
INLINE["camera_environment"] = _camera('"environment"', "environment")

INLINE["random_sampler_path_lens"] = _bestcandidate('"path" "integer maxdepth" [4]', "random_path_lens", 'Sampler "random" "integer xsamples" [2] "integer ysamples" [2]',
                                                    lens='"float lensradius" [15] "float focaldistance" [900]').replace("bc_random_path_lens_out", "random_sampler_path_lens_out")
INLINE["random_sampler_direct_nsamples3"] = _bestcandidate('"directlighting"', "random_direct", 'Sampler "random" "integer xsamples" [3] "integer ysamples" [1]',
                                                           nsamples=3).replace("bc_random_direct_out", "random_sampler_direct_nsamples3_out")


def _translucent(integrator, tag):
    """translucent material (BRDFToBTDF lobes) lit from behind and in front; scale / mix textures of constants"""
    return """LookAt 278 273 -800 278 273 0 0 1 0
Camera "perspective" "float fov" [39.3]
Film "image" "integer xresolution" [32] "integer yresolution" [32] "string filename" ["translucent_%s_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator %s
WorldBegin
LightSource "point" "color I" [300000 280000 250000] "point from" [278 300 500]
AttributeBegin
AreaLightSource "area" "color L" [17 12 4]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [213 548.7 227 343 548.7 227 343 548.7 332 213 548.7 332]
AttributeEnd
Texture "base" "color" "constant" "color value" [.8 .7 .5]
Texture "tint" "color" "scale" "texture tex1" "base" "color tex2" [.9 .5 .9]
Texture "blend" "color" "mix" "texture tex1" "base" "color tex2" [.1 .2 .9] "float amount" [.3]
Texture "rough" "float" "mix" "float tex1" [.05] "float tex2" [.4] "float amount" [.25]
Material "matte" "texture Kd" "blend"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 0 0 0 555 555 0 555 555 0 0]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [0 0 555 0 555 555 555 555 555 555 0 555]
Material "translucent" "texture Kd" "tint" "color Ks" [.4 .4 .4] "color reflect" [.4 .4 .4] "color transmit" [.6 .6 .5] "texture roughness" "rough"
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [80 20 300 480 20 300 480 420 330 80 420 330]
AttributeBegin
Translate 400 100 150
Material "translucent" "color Kd" [.2 .6 .3] "color Ks" [0 0 0] "color transmit" [.9 .9 .9] "color reflect" [0 0 0]
Shape "sphere" "float radius" [90]
AttributeEnd
WorldEnd
""" % (tag, integrator)


INLINE["translucent_path"] = _translucent('"path" "integer maxdepth" [4]', "path")
INLINE["translucent_direct"] = _translucent('"directlighting"', "direct")
INLINE["translucent_whitted"] = _translucent('"whitted"', "whitted")

INLINE["substrate_path"] = _uber_shiny('"path" "integer maxdepth" [4]', "path").replace("uber_shiny_path_out", "substrate_path_out") \
    .replace('Material "shinymetal" "color Ks" [.9 .7 .3] "color Kr" [.8 .8 .9] "float roughness" [.15]',
             'Material "substrate" "color Kd" [.3 .4 .7] "color Ks" [.3 .3 .3] "float uroughness" [.05] "float vroughness" [.3]') \
    .replace('Material "shinymetal" "float roughness" [.01]', 'Material "substrate" "float uroughness" [.0005] "float vroughness" [.0005]') \
    .replace('Material "uber" "color Kd" [.4 .5 .3] "color Ks" [.5 .5 .5] "color Kr" [.3 .3 .3] "float roughness" [.08]',
             'Material "substrate" "color Kd" [.6 .5 .3] "color Ks" [.1 .1 .1] "float uroughness" [.2] "float vroughness" [.2]')
INLINE["substrate_direct"] = INLINE["substrate_path"].replace('SurfaceIntegrator "path" "integer maxdepth" [4]', 'SurfaceIntegrator "directlighting"') \
    .replace("substrate_path_out", "substrate_direct_out")


def _loopsubdiv():
    """Loop subdivision surfaces: a closed octahedron (irregular interior vertices), an open bumpy grid
    (boundary vertices of valence 2, 3, 4, regular interior ones) and an open fan (boundary vertex of valence 6)"""
    import math
    grid_p, grid_i = [], []
    n = 4
    for y in range(n):
        for x in range(n):
            grid_p += [x - 1.5, 0.3 * math.sin(2.1 * x + 0.7 * y), y - 1.5]
    for y in range(n - 1):
        for x in range(n - 1):
            a = y * n + x
            grid_i += [a, a + n, a + 1, a + 1, a + n, a + n + 1]
    fan_p = [0, 0, 0]
    for k in range(6):
        t = math.pi * k / 5
        fan_p += [math.cos(t), 0.2 * (k % 2), math.sin(t)]
    fan_i = []
    for k in range(5):
        fan_i += [0, k + 2, k + 1]
    f = lambda v: " ".join("%.6g" % t for t in v)
    return """LookAt 0 4 -9 0 0.3 0 0 1 0
Camera "perspective" "float fov" [38]
Film "image" "integer xresolution" [44] "integer yresolution" [30] "string filename" ["loopsubdiv_path_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator "path" "integer maxdepth" [3]
WorldBegin
LightSource "point" "color I" [70 65 60] "point from" [3 7 -6]
AttributeBegin
AreaLightSource "area" "color L" [8 8 9]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-2 6 -1 2 6 -1 2 6 2 -2 6 2]
AttributeEnd
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1.2 -9 9 -1.2 -9 9 -1.2 9 -9 -1.2 9]
AttributeBegin
Material "plastic" "color Kd" [.7 .3 .2] "color Ks" [.3 .3 .3] "float roughness" [.08]
Translate -3 0.5 0
Shape "loopsubdiv" "integer nlevels" [2] "integer indices" [0 2 4 2 1 4 1 3 4 3 0 4 2 0 5 1 2 5 3 1 5 0 3 5]
  "point P" [1 0 0 -1 0 0 0 0 1 0 0 -1 0 1.3 0 0 -1 0]
AttributeEnd
AttributeBegin
Material "matte" "color Kd" [.2 .5 .7]
Translate 0.3 -0.3 0.5
Rotate 20 1 0 0
Shape "loopsubdiv" "integer nlevels" [2] "integer indices" [%s] "point P" [%s]
AttributeEnd
AttributeBegin
Material "mirror" "color Kr" [.8 .8 .8]
Translate 3.4 -0.2 -0.5
Rotate -60 1 0 0
Scale 1.3 1.3 1.3
Shape "loopsubdiv" "integer nlevels" [3] "integer indices" [%s] "point P" [%s]
AttributeEnd
WorldEnd
""" % (" ".join(map(str, grid_i)), f(grid_p), " ".join(map(str, fan_i)), f(fan_p))


INLINE["loopsubdiv_path"] = _loopsubdiv()


def _nurbs():
    """NURBS patches: a bicubic 5 x 4 patch (non-homogeneous control points, interior knots) and a
    rational quadratic one (Pw, weights: a quarter cylinder swept along a bent profile)"""
    import math
    P = []
    for v in range(4):
        for u in range(5):
            P += [u - 2.0, 0.6 * math.sin(1.3 * u) * math.cos(0.9 * v) + 0.2 * v, v - 1.5]
    w = math.sqrt(0.5)
    prof = [(1, 0, 1), (1, 1, w), (0, 1, 1)]
    Pw = []
    for v, (z, bend) in enumerate([(-1.5, 0.0), (0.0, 0.5), (1.5, 0.0)]):
        for (x, y, wt) in prof:
            Pw += [(x + bend) * wt, y * wt, z * wt, wt]
    f = lambda v: " ".join("%.7g" % t for t in v)
    return """LookAt 0 4 -9 0 0.3 0 0 1 0
Camera "perspective" "float fov" [36]
Film "image" "integer xresolution" [40] "integer yresolution" [28] "string filename" ["nurbs_path_out.exr"]
Sampler "stratified" "integer xsamples" [2] "integer ysamples" [2]
PixelFilter "box"
SurfaceIntegrator "path" "integer maxdepth" [3]
WorldBegin
LightSource "point" "color I" [70 65 60] "point from" [3 7 -6]
AttributeBegin
AreaLightSource "area" "color L" [8 8 9]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-2 6 -1 2 6 -1 2 6 2 -2 6 2]
AttributeEnd
Material "matte" "color Kd" [.6 .6 .6]
Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1.2 -9 9 -1.2 -9 9 -1.2 9 -9 -1.2 9]
AttributeBegin
Material "plastic" "color Kd" [.3 .6 .7] "color Ks" [.3 .3 .3] "float roughness" [.1]
Translate -1.8 0 0.5
Shape "nurbs" "integer nu" [5] "integer uorder" [4] "float uknots" [0 0 0 0 .5 1 1 1 1]
  "integer nv" [4] "integer vorder" [4] "float vknots" [0 0 0 0 1 1 1 1] "point P" [%s]
AttributeEnd
AttributeBegin
Material "shinymetal" "float roughness" [.05]
Translate 2.6 -0.6 0
Rotate -30 0 1 0
Scale 1.4 1.4 1.4
Shape "nurbs" "integer nu" [3] "integer uorder" [3] "float uknots" [0 0 0 1 1 1] "float u0" [.05] "float u1" [.95]
  "integer nv" [3] "integer vorder" [3] "float vknots" [0 0 0 1 1 1] "float Pw" [%s]
AttributeEnd
WorldEnd
""" % (f(P), f(Pw))


INLINE["nurbs_path"] = _nurbs()


def parse_count(text, label):
    m = re.search(re.escape(label) + r"\s+([0-9.]+)([kMG]?)", text)
    if not m:
        return -1
    mult = {"": 1, "k": 1e3, "M": 1e6, "G": 1e9}[m.group(2)]
    return int(float(m.group(1)) * mult)


def main():
    ffi = load_ffi()
    assert ffi.have_reference(), "oracle/_ref missing: run oracle/build_ref.sh (needs /root/reference)"
    cases = list(CASES) + [(n, None, None, True) for n in INLINE]
    make_textures()
    if len(sys.argv) > 1:   # only the named cases
        cases = [c for c in cases if c[0] in sys.argv[1:]]
    for name, scene, args, keyed in cases:
        sp = os.path.join(HERE, name + ".pbrt")
        # relative output name so the committed scene file works from any checkout
        if scene is None:
            open(sp, "w").write(INLINE[name])
        else:
            gen_scene(sp, scene, "--filename", name + "_out.exr", *args)
        dump = os.path.join(HERE, name + ".samples")
        out, _ = ffi.run_reference(sp, keyed=keyed, dump_samples=dump, cwd=HERE)
        rgba, _ = ffi.read_raw(os.path.join(HERE, name + "_out.exr.raw"))
        samples = read_sample_dump(dump)
        stats = np.array([parse_count(out, "Camera Rays Traced"), parse_count(out, "Number of shadow rays traced"),
                          parse_count(out, "Number of points shaded")], np.int64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rgba=rgba, samples=samples, stats=stats)
        os.remove(dump)
        os.remove(os.path.join(HERE, name + "_out.exr.raw"))
        print("%-22s %s  image %s  samples %d  mean %.6f" % (name, "keyed" if keyed else "exact", rgba.shape[:2],
                                                             samples.shape[0], float(rgba[..., :3].mean())))


if __name__ == "__main__":
    main()
