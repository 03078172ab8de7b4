"""Seeded random .pbrt scenes mixing every supported plugin (shapes, materials, lights, cameras, samplers,
integrators, filters, instancing, constant textures) at thumbnail size: input for the differential tests
(compiled reference vs oracle twin on the CPU; CUDA path vs twin on the GPU).  scene(random.Random, idx) -> text."""
import math
import os
import re

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def f(*v):
    return " ".join("%.5g" % x for x in v)


def rnd_color(r, lo=0.05, hi=0.95):
    return f(r.uniform(lo, hi), r.uniform(lo, hi), r.uniform(lo, hi))


def material(r):
    k = r.choice(["matte", "matte_on", "glass", "mirror", "plastic", "uber", "shinymetal", "translucent", "substrate", "measured"])
    if k == "measured":
        return 'Material "%s"' % r.choice(["bluepaint", "brushedmetal", "clay", "felt", "primer", "skin"])
    if k == "matte":
        return 'Material "matte" "color Kd" [%s]' % rnd_color(r)
    if k == "matte_on":
        return 'Material "matte" "color Kd" [%s] "float sigma" [%s]' % (rnd_color(r), f(r.uniform(5, 60)))
    if k == "glass":
        return 'Material "glass" "float index" [%s] "color Kr" [%s] "color Kt" [%s]' % (f(r.uniform(1.1, 1.8)), rnd_color(r, .5), rnd_color(r, .5))
    if k == "mirror":
        return 'Material "mirror" "color Kr" [%s]' % rnd_color(r, .4)
    if k == "plastic":
        return 'Material "plastic" "color Kd" [%s] "color Ks" [%s] "float roughness" [%s]' % (rnd_color(r), rnd_color(r, .1, .6), f(r.choice([.002, .05, .2, .6])))
    if k == "uber":
        op = r.choice(["", ' "color opacity" [%s]' % rnd_color(r, .2, 1.0)])
        return 'Material "uber" "color Kd" [%s] "color Ks" [%s] "color Kr" [%s] "float roughness" [%s]%s' % (
            rnd_color(r), rnd_color(r, 0, .5), r.choice(["0 0 0", rnd_color(r, 0, .5)]), f(r.uniform(.01, .5)), op)
    if k == "shinymetal":
        return 'Material "shinymetal" "color Ks" [%s] "color Kr" [%s] "float roughness" [%s]' % (rnd_color(r, .3), rnd_color(r, .3), f(r.uniform(.005, .4)))
    if k == "translucent":
        return 'Material "translucent" "color Kd" [%s] "color Ks" [%s] "color reflect" [%s] "color transmit" [%s] "float roughness" [%s]' % (
            rnd_color(r), r.choice(["0 0 0", rnd_color(r, 0, .5)]), r.choice(["0 0 0", rnd_color(r)]), r.choice(["0 0 0", rnd_color(r)]), f(r.uniform(.02, .5)))
    return 'Material "substrate" "color Kd" [%s] "color Ks" [%s] "float uroughness" [%s] "float vroughness" [%s]' % (
        rnd_color(r), rnd_color(r, .05, .5), f(r.choice([.001, .05, .3])), f(r.choice([.001, .05, .3])))


def textures(r):
    """image-map textures (and scale / mix expressions over them) for a world block; returns the statements and
    the names usable as color / float operands"""
    out, color, flt = [], [], []
    files = ["tex13x7.exr", "tex8x16.exr", "tex5x5.exr"]
    for i in range(r.randint(1, 3)):
        for typ, names in (("color", color), ("float", flt)):
            m = r.choice(["uv", "uv", "spherical", "cylindrical", "planar"])
            p = '"string mapping" ["%s"]' % m
            if m == "uv":
                p += ' "float uscale" [%s] "float vscale" [%s] "float udelta" [%s]' % (f(r.choice([1, 2, 7, 30])), f(r.choice([1, 3, 11])), f(r.uniform(-1, 1)))
            if m == "planar":
                p += ' "vector v1" [%s] "vector v2" [%s]' % (f(r.uniform(-.5, .5), 0, r.uniform(-.5, .5)), f(0, r.uniform(-.5, .5), r.uniform(-.5, .5)))
            p += r.choice(["", ' "string wrap" ["clamp"]', ' "string wrap" ["black"]', ' "bool trilinear" ["true"]', ' "float maxanisotropy" [%s]' % f(r.choice([1, 2, 20]))])
            pre = r.choice(["", "Rotate %s 1 1 0\n" % f(r.uniform(0, 360)), "Scale 1 %s 1\nTranslate .3 .2 .1\n" % f(r.uniform(.5, 2))])
            name = "%s%d" % (typ[0], i)
            out.append('TransformBegin\n%sTexture "%s" "%s" "imagemap" "string filename" ["%s"] %s\nTransformEnd' % (
                pre, name, typ, os.path.join(GOLDEN, r.choice(files)), p))
            names.append(name)
    # procedural textures (textures/{uv,bilerp,checkerboard,dots,fbm,wrinkled,windy,marble}.cpp)
    m2 = lambda: r.choice(['', ' "string mapping" ["spherical"]', ' "string mapping" ["cylindrical"]',
                           ' "float uscale" [%s] "float vscale" [%s]' % (f(r.choice([2, 6, 15])), f(r.choice([3, 9])))])
    pre3 = "TransformBegin\nRotate %s 1 0 1\nScale %s 1 %s\n" % (f(r.uniform(0, 360)), f(r.uniform(.5, 3)), f(r.uniform(.5, 3)))
    out.append('Texture "p_uv" "color" "uv"%s' % m2())
    out.append('Texture "p_bil" "color" "bilerp" "color v00" [%s] "color v11" [%s]%s' % (rnd_color(r), rnd_color(r), m2()))
    out.append('Texture "p_fbil" "float" "bilerp" "float v01" [%s]%s' % (f(r.uniform(0, 1)), m2()))
    out.append('Texture "p_chk" "color" "checkerboard" "texture tex1" "%s" "color tex2" [%s]%s%s' % (
        r.choice(color + ["p_uv", "p_bil"]), rnd_color(r), r.choice(['', ' "string aamode" ["none"]']), m2()))
    out.append('Texture "p_fchk" "float" "checkerboard" "float tex1" [.1] "texture tex2" "p_fbil"%s' % m2())
    out.append('Texture "p_dot" "color" "dots" "texture inside" "%s" "color outside" [%s]%s' % (r.choice(color), rnd_color(r), m2()))
    out.append(pre3 + 'Texture "p_fbm" "color" "fbm" "integer octaves" [%d] "float roughness" [%s]\nTexture "p_ffbm" "float" "fbm"\n'
               'Texture "p_wr" "color" "wrinkled"\nTexture "p_wi" "float" "windy"\nTexture "p_mar" "color" "marble" "float scale" [%s]\nTransformEnd' % (
                   r.randint(1, 8), f(r.uniform(.3, .7)), f(r.uniform(.5, 4))))
    color += ["p_uv", "p_bil", "p_chk", "p_dot", "p_fbm", "p_wr", "p_mar"]
    flt += ["p_fbil", "p_fchk", "p_wi"]
    out.append('Texture "cs" "color" "scale" "texture tex1" "%s" "color tex2" [%s]' % (r.choice(color), rnd_color(r)))
    out.append('Texture "cm" "color" "mix" "texture tex1" "cs" "texture tex2" "%s" "texture amount" "%s"' % (r.choice(color), r.choice(flt)))
    color += ["cs", "cm"]
    out.append('Texture "frough" "float" "mix" "float tex1" [.02] "float tex2" [.5] "texture amount" "%s"' % r.choice(flt))
    out.append('Texture "findex" "float" "mix" "float tex1" [1.2] "float tex2" [1.8] "texture amount" "%s"' % r.choice(flt))
    out.append('Texture "fsigma" "float" "scale" "texture tex1" "%s" "float tex2" [50]' % r.choice(flt))
    out.append('Texture "fbump" "float" "scale" "texture tex1" "%s" "float tex2" [%s]' % (r.choice(flt), f(r.uniform(-.3, .3))))
    return out, color


def retexture(r, mat, color):
    """swaps some constant operands of a Material statement for textures"""
    def csub(m):
        return '"texture %s" "%s"' % (m.group(1), r.choice(color)) if r.random() < .5 else m.group(0)

    def fsub(m):
        name = {"sigma": "fsigma", "index": "findex"}.get(m.group(1), "frough")
        return '"texture %s" "%s"' % (m.group(1), name) if r.random() < .5 else m.group(0)
    if r.random() < .3:
        mat += ' "texture bumpmap" "%s"' % r.choice(["fbump", "frough"])
    mat = re.sub(r'"color (Kd|Ks|Kr|Kt|opacity|reflect|transmit)" \[[^\]]*\]', csub, mat)
    return re.sub(r'"float (sigma|index|roughness|uroughness|vroughness)" \[[^\]]*\]', fsub, mat)


def shape(r):
    k = r.choice(["sphere", "sphere_clip", "cylinder", "disk", "cone", "paraboloid", "hyperboloid", "tri", "heightfield", "loopsubdiv", "nurbs"])
    pm = r.choice([360, 360, r.uniform(60, 350)])
    if k == "sphere":
        return 'Shape "sphere" "float radius" [%s]' % f(r.uniform(.4, 1.2))
    if k == "sphere_clip":
        rad = r.uniform(.5, 1.2)
        return 'Shape "sphere" "float radius" [%s] "float zmin" [%s] "float zmax" [%s] "float phimax" [%s]' % (f(rad), f(-rad * r.uniform(.2, 1)), f(rad * r.uniform(.2, 1)), f(pm))
    if k == "cylinder":
        return 'Shape "cylinder" "float radius" [%s] "float zmin" [%s] "float zmax" [%s] "float phimax" [%s]' % (f(r.uniform(.3, .9)), f(r.uniform(-1, 0)), f(r.uniform(.2, 1.2)), f(pm))
    if k == "disk":
        return 'Shape "disk" "float radius" [%s] "float innerradius" [%s] "float height" [%s] "float phimax" [%s]' % (f(r.uniform(.6, 1.3)), f(r.choice([0, r.uniform(.1, .5)])), f(r.uniform(-.3, .3)), f(pm))
    if k == "cone":
        return 'Shape "cone" "float radius" [%s] "float height" [%s] "float phimax" [%s]' % (f(r.uniform(.4, 1)), f(r.uniform(.6, 1.8)), f(pm))
    if k == "paraboloid":
        return 'Shape "paraboloid" "float radius" [%s] "float zmin" [%s] "float zmax" [%s] "float phimax" [%s]' % (f(r.uniform(.5, 1)), f(r.uniform(0, .4)), f(r.uniform(.8, 1.8)), f(pm))
    if k == "hyperboloid":
        p1 = (r.uniform(.3, 1), r.uniform(-.4, .4), r.choice([0, r.uniform(-.8, 0)]))
        p2 = (r.uniform(.2, .9), r.uniform(-.5, .5), r.choice([0, r.uniform(.6, 1.6)]) if p1[2] != 0 else r.uniform(.6, 1.6))
        return 'Shape "hyperboloid" "point p1" [%s] "point p2" [%s] "float phimax" [%s]' % (f(*p1), f(*p2), f(pm))
    if k == "tri":
        n = r.choice([False, True])
        s = 'Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 -1 0 1 -1 %s 1 1 0 -1 1 %s]' % (f(r.uniform(-.3, .3)), f(r.uniform(-.3, .3)))
        if n:
            s += ' "normal N" [0 .3 -1 .3 0 -1 0 -.3 -1 -.3 0 -1] "float uv" [0 0 1 0 1 1 0 1]'
        return s
    if k == "heightfield":
        nu, nv = r.randint(2, 5), r.randint(2, 5)
        return 'Translate -0.5 -0.5 0 Shape "heightfield" "integer nu" [%d] "integer nv" [%d] "float Pz" [%s]' % (nu, nv, f(*[r.uniform(0, .4) for _ in range(nu * nv)]))
    if k == "loopsubdiv":
        if r.random() < .5:   # octahedron
            return ('Shape "loopsubdiv" "integer nlevels" [%d] "integer indices" [0 2 4 2 1 4 1 3 4 3 0 4 2 0 5 1 2 5 3 1 5 0 3 5] '
                    '"point P" [1 0 0 -1 0 0 0 0 1 0 0 -1 0 %s 0 0 -1 0]' % (r.randint(1, 3), f(r.uniform(.7, 1.5))))
        nfan = r.randint(2, 6)   # open fan
        P = [0, 0, 0]
        for k2 in range(nfan + 1):
            t = math.pi * r.uniform(.9, 1.6) * k2 / nfan
            P += [math.cos(t), math.sin(t), r.uniform(-.2, .2)]
        idx = []
        for k2 in range(nfan):
            idx += [0, k2 + 1, k2 + 2]
        return 'Shape "loopsubdiv" "integer nlevels" [%d] "integer indices" [%s] "point P" [%s]' % (r.randint(1, 2), " ".join(map(str, idx)), f(*P))
    # nurbs
    nu, nv = r.randint(3, 5), r.randint(3, 4)
    uo, vo = r.randint(2, min(4, nu)), r.randint(2, min(4, nv))

    def knots(n, o):
        interior = sorted(r.uniform(.1, .9) for _ in range(n - o))
        return [0.0] * o + interior + [1.0] * o
    P = []
    hom = r.random() < .4
    for v in range(nv):
        for u in range(nu):
            p = [u - (nu - 1) / 2, v - (nv - 1) / 2, r.uniform(-.5, .5)]
            if hom:
                w = r.uniform(.5, 1.5)
                P += [p[0] * w, p[1] * w, p[2] * w, w]
            else:
                P += p
    return 'Scale .5 .5 .5 Shape "nurbs" "integer nu" [%d] "integer uorder" [%d] "float uknots" [%s] "integer nv" [%d] "integer vorder" [%d] "float vknots" [%s] "%s" [%s]' % (
        nu, uo, f(*knots(nu, uo)), nv, vo, f(*knots(nv, vo)), "float Pw" if hom else "point P", f(*P))


def xform(r):
    s = "Translate %s\n" % f(r.uniform(-3, 3), r.uniform(-.5, 1.5), r.uniform(-2, 2))
    if r.random() < .7:
        s += "Rotate %s %s\n" % (f(r.uniform(0, 360)), r.choice(["1 0 0", "0 1 0", "0 0 1", "1 1 0"]))
    if r.random() < .4:
        s += "Scale %s\n" % f(r.uniform(.6, 1.4) * r.choice([1, 1, -1]), r.uniform(.6, 1.4), r.uniform(.6, 1.4))
    if r.random() < .15:
        s += "ReverseOrientation\n"
    return s


def scene(r, idx):
    w, h = r.randint(10, 20), r.randint(8, 16)
    cam = r.choice(['"perspective" "float fov" [%s]' % f(r.uniform(30, 60)),
                    '"perspective" "float fov" [45] "float lensradius" [%s] "float focaldistance" [%s]' % (f(r.uniform(.05, .3)), f(r.uniform(5, 10))),
                    '"orthographic" "float screenwindow" [-4 4 -3 3]', '"environment"'])
    samp = r.choice(['Sampler "stratified" "integer xsamples" [%d] "integer ysamples" [%d] "bool jitter" ["%s"]' % (r.randint(1, 2), r.randint(1, 2), r.choice(["true", "false"])),
                     'Sampler "lowdiscrepancy" "integer pixelsamples" [%d]' % r.choice([1, 2, 4]),
                     'Sampler "bestcandidate" "integer pixelsamples" [%d]' % r.randint(1, 5),
                     'Sampler "random" "integer xsamples" [%d] "integer ysamples" [%d]' % (r.randint(1, 2), r.randint(1, 2)), ""])
    filt = r.choice(['PixelFilter "box"', 'PixelFilter "mitchell"', 'PixelFilter "gaussian"', 'PixelFilter "triangle" "float xwidth" [1.5]', 'PixelFilter "sinc"', ""])
    integ = r.choice(['SurfaceIntegrator "path" "integer maxdepth" [%d]' % r.randint(1, 5), 'SurfaceIntegrator "whitted" "integer maxdepth" [%d]' % r.randint(1, 4),
                      'SurfaceIntegrator "directlighting" "string strategy" ["%s"]' % r.choice(["all", "one"]), ""])
    crop = r.choice(["", "", ' "float cropwindow" [%s]' % f(r.uniform(0, .3), r.uniform(.6, 1), r.uniform(0, .3), r.uniform(.6, 1))])
    out = ["LookAt 0 2.5 -8 0 0.3 0 0 1 0", "Camera " + cam,
           'Film "image" "integer xresolution" [%d] "integer yresolution" [%d]%s "string filename" ["fuzz_%d.exr"]' % (w, h, crop, idx),
           samp, filt, integ, "WorldBegin"]
    nl = 0
    for _ in range(r.randint(0, 2)):
        k = r.choice(["point", "spot", "distant", "infinite", "goniometric", "projection"])
        nl += 1
        if k == "point":
            out.append('LightSource "point" "color I" [%s] "point from" [%s]' % (rnd_color(r, 20, 60), f(r.uniform(-4, 4), r.uniform(3, 7), r.uniform(-6, 0))))
        elif k == "spot":
            out.append('LightSource "spot" "color I" [%s] "point from" [%s] "point to" [0 0 0] "float coneangle" [%s] "float conedeltaangle" [%s]' % (
                rnd_color(r, 40, 120), f(r.uniform(-4, 4), r.uniform(3, 7), r.uniform(-6, 0)), f(r.uniform(20, 50)), f(r.uniform(2, 15))))
        elif k in ("goniometric", "projection"):
            mapn = ' "string mapname" ["%s"]' % os.path.join(GOLDEN, r.choice(["tex13x7.exr", "tex8x16.exr", "tex5x5.exr"]))
            if k == "goniometric" and r.random() < .3:
                mapn = ""
            out.append('AttributeBegin\nTranslate %s\nRotate %s 1 0 .5\nLightSource "%s" "color I" [%s]%s%s\nAttributeEnd' % (
                f(r.uniform(-4, 4), r.uniform(3, 7), r.uniform(-6, 0)), f(r.uniform(40, 140)), k, rnd_color(r, 40, 150), mapn,
                ' "float fov" [%s]' % f(r.uniform(30, 90)) if k == "projection" else ""))
        elif k == "distant":
            out.append('LightSource "distant" "color L" [%s] "point from" [%s] "point to" [0 0 0]' % (rnd_color(r, .5, 2), f(r.uniform(-3, 3), 5, r.uniform(-5, 0))))
        else:
            lmap = ""
            if r.random() < .5:
                lmap = ' "string mapname" ["%s"]' % os.path.join(GOLDEN, r.choice(["tex13x7.exr", "tex8x16.exr", "tex5x5.exr"]))
            out.append('AttributeBegin\nRotate %s 1 .3 0\nLightSource "%s" "color L" [%s] "integer nsamples" [%d]%s\nAttributeEnd' % (
                f(r.uniform(0, 360)), "infinitesample" if lmap and r.random() < .5 else "infinite", rnd_color(r, .2, 1), r.randint(1, 3), lmap))
    for _ in range(r.randint(0 if nl else 1, 2)):
        sh = r.choice(['Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-1 0 -1 1 0 -1 1 0 1 -1 0 1]',
                       'Shape "sphere" "float radius" [.5]', 'Shape "disk" "float radius" [.9]',
                       'Shape "cylinder" "float radius" [.2] "float zmin" [-.8] "float zmax" [.8]'])
        out.append('AttributeBegin\nAreaLightSource "area" "color L" [%s] "integer nsamples" [%d]\nTranslate %s\nRotate %s 1 0 0\n%s\nAttributeEnd' % (
            rnd_color(r, 4, 15), r.randint(1, 3), f(r.uniform(-3, 3), r.uniform(3.5, 6), r.uniform(-2, 2)), f(r.choice([90, 90, 60, 200])), sh))
    color = None
    if r.random() < .4:
        stmts, color = textures(r)
        out += stmts

    def mat():
        m = material(r)
        if color:
            return retexture(r, m, color)
        return m + (' "float bumpmap" [%s]' % f(r.uniform(-.2, .2)) if r.random() < .1 else "")
    out.append('Texture "tc" "color" "constant" "color value" [%s]' % rnd_color(r))
    out.append('Texture "tm" "color" "mix" "texture tex1" "tc" "color tex2" [%s] "float amount" [%s]' % (rnd_color(r), f(r.uniform(0, 1))))
    out.append('Material "matte" "texture Kd" "tm"')
    out.append('Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-9 -1 -9 9 -1 -9 9 -1 9 -9 -1 9]')
    has_obj = r.random() < .5
    if has_obj:
        out.append('ObjectBegin "obj"')
        for _ in range(r.randint(1, 3)):
            lit = 'AreaLightSource "area" "color L" [%s]\n' % rnd_color(r, 1, 6) if r.random() < .2 else ""
            out.append("AttributeBegin\n" + lit + mat() + "\n" + xform(r).replace("Translate", "Translate 0 0 0 Translate", 1) + shape(r) + "\nAttributeEnd")
        out.append("ObjectEnd")
    for _ in range(r.randint(2, 6)):
        if has_obj and r.random() < .35:
            out.append("AttributeBegin\n" + xform(r) + 'ObjectInstance "obj"\nAttributeEnd')
        else:
            out.append("AttributeBegin\n" + mat() + "\n" + xform(r) + shape(r) + "\nAttributeEnd")
    out.append("WorldEnd")
    return "\n".join(out) + "\n"
