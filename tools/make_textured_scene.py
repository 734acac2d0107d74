#!/usr/bin/env python3
"""Writes scenes/textured.xml + scenes/textures/*.pfm + scenes/meshes/tex_*.obj: the textured material-ball scene in
Mitsuba 0.6's XML dialect.  Same data as mitsuba_b200.scene.textured_scene(tex_res=128, n_theta=32, n_phi=64)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mitsuba_b200.scene import Texture, textured_scene

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scenes")
os.makedirs(os.path.join(ROOT, "meshes"), exist_ok=True)
os.makedirs(os.path.join(ROOT, "textures"), exist_ok=True)
d = textured_scene(512, 512, tex_res=128, n_theta=32, n_phi=64)


def write_pfm(fn, img):
    h, w = img.shape[:2]
    with open(os.path.join(ROOT, fn), "wb") as f:
        f.write(b"PF\n" if img.ndim == 3 else b"Pf\n")
        f.write(f"{w} {h}\n-1.0\n".encode())
        f.write(np.ascontiguousarray(img[::-1], "<f4").tobytes())  # PFM stores the bottom row first


def write_obj(fn, m):
    with open(os.path.join(ROOT, fn), "w") as f:
        f.write(f"# {m.name}\n")
        for p in m.P:
            f.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in p))
        if m.UV is not None:
            for t in m.UV:
                f.write("vt %.9g %.9g\n" % (float(t[0]), float(t[1])))
        if m.N is not None:
            for n in m.N:
                f.write("vn %.9g %.9g %.9g\n" % tuple(float(x) for x in n))
        for t in m.idx:
            c = [int(i) + 1 for i in t]
            if m.UV is not None and m.N is not None:
                f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (c[0], c[0], c[0], c[1], c[1], c[1], c[2], c[2], c[2]))
            elif m.UV is not None:
                f.write("f %d/%d %d/%d %d/%d\n" % (c[0], c[0], c[1], c[1], c[2], c[2]))
            else:
                f.write("f %d %d %d\n" % tuple(c))


parts = []
for m in d.meshes:
    fn = f"meshes/tex_{m.name}.obj"
    write_obj(fn, m)
    r = m.bsdf.reflectance
    if isinstance(r, Texture):
        tf = f"textures/{m.name}.pfm"
        write_pfm(tf, np.asarray(r.pixels, np.float32))
        extra = "".join(f'\n\t\t\t\t<float name="{k}" value="{v:.9g}"/>' for k, v in (("uscale", r.uscale), ("vscale", r.vscale), ("uoffset", r.uoffset),
                                                                                 ("voffset", r.voffset)) if v != (1.0 if "scale" in k else 0.0))
        refl = (f'\t\t\t<texture type="bitmap" name="reflectance">\n\t\t\t\t<string name="filename" value="{tf}"/>\n\t\t\t\t<string name="filterType" value="$filter"/>'
                f'\n\t\t\t\t<string name="wrapModeU" value="{r.wrap_u}"/>\n\t\t\t\t<string name="wrapModeV" value="{r.wrap_v}"/>{extra}\n\t\t\t</texture>')
    else:
        refl = '\t\t\t<rgb name="reflectance" value="%s"/>' % " ".join("%.9g" % float(x) for x in r)
    em = ""
    if m.radiance is not None:
        em = '\n\t\t<emitter type="area">\n\t\t\t<rgb name="radiance" value="%s"/>\n\t\t</emitter>' % " ".join("%.9g" % x for x in m.radiance)
    fnorm = "" if m.N is not None else '\n\t\t<boolean name="faceNormals" value="true"/>'
    # the data already holds Mitsuba's v (the file is written unflipped)
    parts.append(f'\t<shape type="obj">\n\t\t<string name="filename" value="{fn}"/>\n\t\t<boolean name="flipTexCoords" value="false"/>{fnorm}\n'
                 f'\t\t<bsdf type="diffuse">\n{refl}\n\t\t</bsdf>{em}\n\t</shape>')
cam = d.camera
xml = f'''<?xml version="1.0" encoding="utf-8"?>
<!-- Textured material ball (SURVEY.md 8f-4): three bitmap textures.  Usage: -D spp=64 -D res=512 -D filter=ewa|trilinear|bilinear|nearest -->
<scene version="0.5.0">
\t<default name="spp" value="16"/>
\t<default name="res" value="256"/>
\t<default name="filter" value="ewa"/>
\t<integrator type="path"/>
\t<sensor type="perspective">
\t\t<float name="fov" value="{cam.fov:.9g}"/>
\t\t<float name="nearClip" value="{cam.near:.9g}"/>
\t\t<float name="farClip" value="{cam.far:.9g}"/>
\t\t<transform name="toWorld">
\t\t\t<matrix value="{" ".join("%.9g" % float(x) for x in cam.to_world.reshape(-1))}"/>
\t\t</transform>
\t\t<sampler type="sobol">
\t\t\t<integer name="sampleCount" value="$spp"/>
\t\t</sampler>
\t\t<film type="hdrfilm">
\t\t\t<integer name="width" value="$res"/>
\t\t\t<integer name="height" value="$res"/>
\t\t\t<rfilter type="box"/>
\t\t</film>
\t</sensor>
{chr(10).join(parts)}
</scene>
'''
open(os.path.join(ROOT, "textured.xml"), "w").write(xml)
print("wrote", os.path.join(ROOT, "textured.xml"))
