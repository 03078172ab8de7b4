#!/usr/bin/env python3
"""Generate the BASELINE.json benchmark/parity scenes as pbrt-v1 scene files.

The reference ships no example scenes, so every input is
synthetic and produced here (scene parameters: SURVEY.md section 8d):

  c1  Whitted, unit sphere + point light, 256x256, 1 spp, no jitter, box filter
  c2  Cornell box, directlighting (strategy all), 512x512, 4x4 spp
  c3  Cornell box + glass sphere, path maxdepth 5, 1024x1024, 8x8 spp
  c4  Cornell walls + ~100k-triangle procedural blob, path, 1024x1024, 8x8 spp
  c5  same with a ~1M-triangle blob, 1920x1080, 16x16 spp

Resolution, samples and the pixel filter can be overridden so the same scenes
serve as small parity cases:  gen_scenes.py c3 --res 64 64 --spp 2 2 -o x.pbrt
"""
import argparse
import math
import sys

import numpy as np


def fmt(v):
    return "%.9g" % float(np.float32(v))


LENS = None      # (lensradius, focaldistance) set from the command line
LIGHT_NSAMPLES = 1
MIRROR_TALL_BOX = False
SPHERE_LIGHT = False
LD_SAMPLER = False
BC_SAMPLER = False


def header(out, res, spp, integrator, filt, jitter=True, fov=39.3, lookat=None,
           crop=None, filename="pbrt.exr", extra_integrator=""):
    lookat = lookat or (278, 273, -800, 278, 273, 0, 0, 1, 0)
    out.write("LookAt %s\n" % " ".join(fmt(v) for v in lookat))
    cam = 'Camera "perspective" "float fov" [%s]' % fmt(fov)
    if LENS:
        cam += ' "float lensradius" [%s] "float focaldistance" [%s]' % (fmt(LENS[0]), fmt(LENS[1]))
    out.write(cam + "\n")
    film = 'Film "image" "integer xresolution" [%d] "integer yresolution" [%d] "string filename" ["%s"]' % (
        res[0], res[1], filename)
    if crop:
        film += ' "float cropwindow" [%s]' % " ".join(fmt(c) for c in crop)
    out.write(film + "\n")
    if LD_SAMPLER:
        out.write('Sampler "lowdiscrepancy" "integer pixelsamples" [%d]\n' % (spp[0] * spp[1]))
    elif BC_SAMPLER:
        out.write('Sampler "bestcandidate" "integer pixelsamples" [%d]\n' % (spp[0] * spp[1]))
    else:
        out.write('Sampler "stratified" "integer xsamples" [%d] "integer ysamples" [%d] "bool jitter" ["%s"]\n' % (
            spp[0], spp[1], "true" if jitter else "false"))
    out.write('PixelFilter "%s"\n' % filt)
    out.write('SurfaceIntegrator "%s" %s\n' % (integrator, extra_integrator))


def quad(out, pts):
    out.write('Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [%s]\n' % " ".join(
        fmt(c) for p in pts for c in p))


def box(out, hx, hy, hz):
    """Axis-aligned box centred at the origin, 5 faces (no bottom), outward normals."""
    x0, x1, y0, y1, z0, z1 = -hx, hx, -hy, hy, -hz, hz
    quad(out, [(x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0)])  # top
    quad(out, [(x0, y0, z0), (x0, y1, z0), (x1, y1, z0), (x1, y0, z0)])  # front (z0)
    quad(out, [(x1, y0, z1), (x1, y1, z1), (x0, y1, z1), (x0, y0, z1)])  # back (z1)
    quad(out, [(x0, y0, z1), (x0, y1, z1), (x0, y1, z0), (x0, y0, z0)])  # x0
    quad(out, [(x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1)])  # x1


TEXTURE_FILE = None   # --textured FILE: walls get image-map / checkerboard / marble colours and an fbm bump map


def cornell_walls(out):
    white, green, red = (.73, .73, .73), (.12, .45, .15), (.65, .05, .05)
    S = 555.0
    if TEXTURE_FILE:
        out.write('Texture "img" "color" "imagemap" "string filename" ["%s"] "string mapping" ["planar"] '
                  '"vector v1" [.004 0 .001] "vector v2" [0 .004 .003]\n' % TEXTURE_FILE)
        out.write('Texture "chk" "color" "checkerboard" "string mapping" ["planar"] "vector v1" [.012 0 0] '
                  '"vector v2" [0 .012 .012] "texture tex1" "img" "color tex2" [.73 .73 .73]\n')
        out.write('TransformBegin\nScale 40 40 40\nTexture "mar" "color" "marble"\nTexture "fb" "float" "fbm"\nTransformEnd\n')
        out.write('Texture "bump" "float" "scale" "texture tex1" "fb" "float tex2" [3]\n')
        out.write('Texture "greenimg" "color" "scale" "texture tex1" "img" "color tex2" [.3 .9 .35]\n')

    def mat(c):
        if TEXTURE_FILE:
            kd = {white: "chk", green: "greenimg", red: "mar"}.get(c)
            if kd:
                out.write('Material "matte" "texture Kd" "%s" "texture bumpmap" "bump"\n' % kd)
                return
        out.write('Material "matte" "color Kd" [%s]\n' % " ".join(fmt(v) for v in c))

    mat(white)
    quad(out, [(0, 0, 0), (0, 0, S), (S, 0, S), (S, 0, 0)])          # floor
    quad(out, [(0, S, 0), (S, S, 0), (S, S, S), (0, S, S)])          # ceiling
    quad(out, [(0, 0, S), (0, S, S), (S, S, S), (S, 0, S)])          # back wall
    mat(green)
    quad(out, [(S, 0, 0), (S, 0, S), (S, S, S), (S, S, 0)])          # left (from the camera)
    mat(red)
    quad(out, [(0, 0, 0), (0, S, 0), (0, S, S), (0, 0, S)])          # right
    # area light just below the ceiling, facing down
    out.write("AttributeBegin\n")
    out.write('AreaLightSource "area" "color L" [17 12 4] "integer nsamples" [%d]\n' % LIGHT_NSAMPLES)
    mat((.78, .78, .78))
    if SPHERE_LIGHT:
        out.write('Translate 278 480 279\nShape "sphere" "float radius" [40]\n')
    else:
        quad(out, [(213, 548.7, 227), (343, 548.7, 227), (343, 548.7, 332), (213, 548.7, 332)])
    out.write("AttributeEnd\n")
    mat(white)


def cornell_boxes(out, short=True, tall=True):
    if short:
        out.write("AttributeBegin\nTranslate 185 82.5 169\nRotate -17 0 1 0\n")
        box(out, 82.5, 82.5, 82.5)
        out.write("AttributeEnd\n")
    if tall:
        out.write("AttributeBegin\nTranslate 368 165 351\nRotate 17 0 1 0\n")
        if MIRROR_TALL_BOX:
            out.write('Material "mirror" "color Kr" [.9 .9 .9]\n')
        box(out, 82.5, 165, 82.5)
        out.write("AttributeEnd\n")


def blob(out, nphi, ntheta, center=(278, 190, 300), radius=150.0, normals=False):
    """Lat-long blob r = R (1 + .15 sin7p sin5t + .05 sin(23p+3) sin17t); pole fans, no degenerate tris."""
    th = (np.arange(1, ntheta, dtype=np.float64) / ntheta) * math.pi          # interior rings
    ph = (np.arange(nphi, dtype=np.float64) / nphi) * 2 * math.pi
    T, P = np.meshgrid(th, ph, indexing="ij")
    r = radius * (1 + .15 * np.sin(7 * P) * np.sin(5 * T) + .05 * np.sin(23 * P + 3) * np.sin(17 * T))
    x = r * np.sin(T) * np.cos(P)
    y = r * np.cos(T)
    z = r * np.sin(T) * np.sin(P)
    ring = np.stack([x, y, z], -1).reshape(-1, 3)
    verts = np.concatenate([[[0, radius, 0]], ring, [[0, -radius, 0]]]).astype(np.float32)
    nr = ntheta - 1
    idx = []
    south = 1 + nr * nphi
    j = np.arange(nphi)
    jn = (j + 1) % nphi
    top = np.stack([np.zeros(nphi, int), 1 + jn, 1 + j], -1)
    idx.append(top)
    for i in range(nr - 1):
        a = 1 + i * nphi + j
        b = 1 + i * nphi + jn
        c = 1 + (i + 1) * nphi + j
        d = 1 + (i + 1) * nphi + jn
        idx.append(np.stack([a, b, d], -1))
        idx.append(np.stack([a, d, c], -1))
    last = 1 + (nr - 1) * nphi
    idx.append(np.stack([last + j, last + jn, np.full(nphi, south)], -1))
    idx = np.concatenate(idx).astype(np.int64)
    out.write("AttributeBegin\nTranslate %s\n" % " ".join(fmt(c) for c in center))
    out.write('Shape "trianglemesh" "integer indices" [\n')
    flat = idx.reshape(-1)
    for s in range(0, flat.size, 30):
        out.write(" ".join(str(int(v)) for v in flat[s:s + 30]) + "\n")
    out.write('] "point P" [\n')
    for s in range(0, verts.shape[0], 8):
        out.write(" ".join("%.9g" % c for c in verts[s:s + 8].reshape(-1)) + "\n")
    out.write("]")
    if normals:
        n = verts / np.maximum(np.linalg.norm(verts, axis=1, keepdims=True), 1e-20)
        out.write(' "normal N" [\n')
        for s in range(0, n.shape[0], 8):
            out.write(" ".join("%.9g" % c for c in n[s:s + 8].reshape(-1)) + "\n")
        out.write("]")
    out.write("\nAttributeEnd\n")
    return idx.shape[0]


def scene_c1(out, a):
    res = a.res or (256, 256)
    spp = a.spp or (1, 1)
    header(out, res, spp, "whitted", a.filter or "box", jitter=a.jitter if a.jitter is not None else False,
           fov=45, lookat=(0, 0, -5, 0, 0, 0, 0, 1, 0), crop=a.crop, filename=a.filename)
    out.write("WorldBegin\n")
    out.write('LightSource "point" "color I" [40 40 40] "point from" [2 4 -4]\n')
    out.write('Material "matte" "color Kd" [.7 .2 .2]\n')
    out.write('Shape "sphere" "float radius" [1]\n')
    out.write("WorldEnd\n")


def scene_c2(out, a):
    res = a.res or (512, 512)
    spp = a.spp or (4, 4)
    header(out, res, spp, "directlighting", a.filter or "box", jitter=a.jitter if a.jitter is not None else True,
           crop=a.crop, filename=a.filename, extra_integrator='"string strategy" ["%s"]' % (a.strategy or "all"))
    out.write("WorldBegin\n")
    cornell_walls(out)
    cornell_boxes(out)
    out.write("WorldEnd\n")


def scene_c3(out, a):
    res = a.res or (1024, 1024)
    spp = a.spp or (8, 8)
    extra = '"integer maxdepth" [%d]' % a.maxdepth
    if a.strategy:
        extra += ' "string strategy" ["%s"]' % a.strategy
    header(out, res, spp, a.integrator or "path", a.filter or "box",
           jitter=a.jitter if a.jitter is not None else True, crop=a.crop, filename=a.filename,
           extra_integrator=extra)
    out.write("WorldBegin\n")
    cornell_walls(out)
    cornell_boxes(out)
    if not a.noglass:
        out.write('AttributeBegin\nMaterial "glass" "float index" [1.5]\nTranslate 185 245.5 169\n')
        out.write('Shape "sphere" "float radius" [80]\nAttributeEnd\n')
    out.write("WorldEnd\n")


def scene_blob(out, a, nphi, ntheta, res, spp):
    header(out, a.res or res, a.spp or spp, a.integrator or "path", a.filter or "box",
           jitter=a.jitter if a.jitter is not None else True, crop=a.crop, filename=a.filename,
           extra_integrator='"integer maxdepth" [%d]' % a.maxdepth)
    out.write("WorldBegin\n")
    cornell_walls(out)
    n = blob(out, a.nphi or nphi, a.ntheta or ntheta, normals=a.normals)
    out.write("WorldEnd\n")
    sys.stderr.write("blob triangles: %d\n" % n)


def scene_instances(out, a):
    """Cornell walls + a 4x4x... grid of instances of one blob object (ObjectBegin/ObjectInstance):
    27 placements of a 20 k-triangle mesh = 0.55 M instanced triangles from one 20 k-triangle tree."""
    header(out, a.res or (1024, 1024), a.spp or (8, 8), a.integrator or "path", a.filter or "box",
           jitter=a.jitter if a.jitter is not None else True, crop=a.crop, filename=a.filename,
           extra_integrator='"integer maxdepth" [%d]' % a.maxdepth)
    out.write("WorldBegin\n")
    cornell_walls(out)
    out.write('ObjectBegin "blob"\nMaterial "matte" "color Kd" [.6 .55 .4]\n')
    n = blob(out, a.nphi or 144, a.ntheta or 72, center=(0, 0, 0), radius=1.0, normals=a.normals)
    out.write("ObjectEnd\n")
    k = 0
    for iy in range(3):
        for iz in range(3):
            for ix in range(3):
                s = 55 + 10 * ((ix + 2 * iy + 3 * iz) % 3)
                out.write("AttributeBegin\nTranslate %g %g %g\nRotate %g 0 1 0\nRotate %g 1 0 0\nScale %g %g %g\n"
                          'ObjectInstance "blob"\nAttributeEnd\n'
                          % (120 + 158 * ix, 85 + 150 * iy, 140 + 150 * iz, 37 * k, 23 * k, s, s * (1 + .1 * (k % 3)), s))
                k += 1
    out.write("WorldEnd\n")
    sys.stderr.write("instanced blob triangles: %d x %d\n" % (n, k))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene", choices=["c1", "c2", "c3", "c4", "c5", "inst"])
    ap.add_argument("-o", "--output", default="-")
    ap.add_argument("--res", type=int, nargs=2)
    ap.add_argument("--spp", type=int, nargs=2)
    ap.add_argument("--filter")
    ap.add_argument("--crop", type=float, nargs=4)
    ap.add_argument("--jitter", type=lambda s: s.lower() in ("1", "true", "yes"), default=None)
    ap.add_argument("--filename", default="pbrt.exr")
    ap.add_argument("--strategy")
    ap.add_argument("--integrator")
    ap.add_argument("--maxdepth", type=int, default=5)
    ap.add_argument("--noglass", action="store_true")
    ap.add_argument("--textured", metavar="EXR", help="texture the Cornell walls from this image (absolute path)")
    ap.add_argument("--normals", action="store_true")
    ap.add_argument("--nphi", type=int)
    ap.add_argument("--ntheta", type=int)
    ap.add_argument("--lens", type=float, nargs=2, help="lensradius focaldistance")
    ap.add_argument("--nsamples", type=int, default=1, help="area light nsamples")
    ap.add_argument("--mirror", action="store_true", help="tall Cornell box becomes a mirror")
    ap.add_argument("--spherelight", action="store_true", help="spherical area light instead of the quad")
    ap.add_argument("--ld", action="store_true", help="lowdiscrepancy sampler with xsamples*ysamples pixel samples")
    ap.add_argument("--bc", action="store_true", help="bestcandidate sampler with xsamples*ysamples pixel samples")
    a = ap.parse_args()
    global LENS, LIGHT_NSAMPLES, MIRROR_TALL_BOX, SPHERE_LIGHT, LD_SAMPLER, BC_SAMPLER, TEXTURE_FILE
    LD_SAMPLER = a.ld
    BC_SAMPLER = a.bc
    LENS = a.lens
    LIGHT_NSAMPLES = a.nsamples
    MIRROR_TALL_BOX = a.mirror
    SPHERE_LIGHT = a.spherelight
    TEXTURE_FILE = a.textured
    out = sys.stdout if a.output == "-" else open(a.output, "w")
    out.write("# generated by scenes/gen_scenes.py %s\n" % " ".join(sys.argv[1:]))
    if a.scene == "c1":
        scene_c1(out, a)
    elif a.scene == "c2":
        scene_c2(out, a)
    elif a.scene == "c3":
        scene_c3(out, a)
    elif a.scene == "inst":
        scene_instances(out, a)
    elif a.scene == "c4":
        scene_blob(out, a, 320, 160, (1024, 1024), (8, 8))
    else:
        scene_blob(out, a, 1000, 500, (1920, 1080), (16, 16))
    if out is not sys.stdout:
        out.close()


if __name__ == "__main__":
    main()
