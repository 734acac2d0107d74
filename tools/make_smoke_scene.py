#!/usr/bin/env python3
"""Writes scenes/smoke.xml + scenes/volumes/smoke_64.vol + scenes/meshes/smoke_*.obj: the S4 scene of SURVEY.md 8(d) (a density
grid in the unit cube, `heterogeneous` Woodcock medium, isotropic phase, `volpath`) in Mitsuba 0.6's XML dialect.  Same data as
mitsuba_b200.scene.smoke_scene(res=64).  The .vol layout is GridDataSource's (src/volume/gridvolume.cpp:225-296)."""
import os, struct, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from mitsuba_b200.scene import smoke_scene

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scenes")
os.makedirs(os.path.join(ROOT, "meshes"), exist_ok=True)
os.makedirs(os.path.join(ROOT, "volumes"), exist_ok=True)
RES = 64
d = smoke_scene(512, 512, res=RES)
med = next(m.interior for m in d.meshes if m.interior is not None)


def write_vol(path, dens, lo, hi):
    nz, ny, nx = dens.shape
    with open(path, "wb") as f:
        f.write(b"VOL\x03")
        f.write(struct.pack("<iiiii", 1, nx, ny, nz, 1))          # type 1 = float32, resolution, channels
        f.write(struct.pack("<6f", *lo, *hi))                      # data box
        f.write(np.ascontiguousarray(dens, "<f4").tobytes())       # x fastest


write_vol(os.path.join(ROOT, "volumes", f"smoke_{RES}.vol"), med.density, med.aabb_min, med.aabb_max)
shapes = []
for m in d.meshes:
    fn = f"meshes/smoke_{m.name}.obj"
    with open(os.path.join(ROOT, fn), "w") as f:
        f.write(f"# {m.name}: smoke scene part\n")
        for p in m.P:
            f.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in p))
        for t in m.idx:
            f.write("f %d %d %d\n" % tuple(int(i) + 1 for i in t))
    body = ""
    if m.interior is not None:
        body = '\n\t\t<ref name="interior" id="smoke"/>'
    else:
        rgb = " ".join("%.9g" % float(x) for x in m.bsdf.reflectance)
        body = f'\n\t\t<bsdf type="diffuse">\n\t\t\t<rgb name="reflectance" value="{rgb}"/>\n\t\t</bsdf>'
    if m.radiance is not None:
        body += '\n\t\t<emitter type="area">\n\t\t\t<rgb name="radiance" value="%s"/>\n\t\t</emitter>' % " ".join("%.9g" % x for x in m.radiance)
    shapes.append(f'\t<shape type="obj">\n\t\t<string name="filename" value="{fn}"/>\n\t\t<boolean name="faceNormals" value="true"/>{body}\n\t</shape>')
alb = " ".join("%.9g" % x for x in med.albedo)
xml = f'''<?xml version="1.0" encoding="utf-8"?>
<!-- Smoke scene (S4): heterogeneous medium behind an index-matched cube.  Usage: -D spp=256 -D res=512 -->
<scene version="0.5.0">
\t<default name="spp" value="64"/>
\t<default name="res" value="256"/>
\t<integrator type="volpath">
\t\t<integer name="maxDepth" value="-1"/>
\t\t<integer name="rrDepth" value="5"/>
\t</integrator>
\t<medium type="heterogeneous" id="smoke">
\t\t<string name="method" value="woodcock"/>
\t\t<float name="scale" value="{med.scale:.9g}"/>
\t\t<volume name="density" type="gridvolume">
\t\t\t<string name="filename" value="volumes/smoke_{RES}.vol"/>
\t\t</volume>
\t\t<volume name="albedo" type="constvolume">
\t\t\t<spectrum name="value" value="{alb}"/>
\t\t</volume>
\t\t<phase type="isotropic"/>
\t</medium>
\t<sensor type="perspective">
\t\t<float name="fov" value="{d.camera.fov:.9g}"/>
\t\t<string name="fovAxis" value="x"/>
\t\t<float name="nearClip" value="{d.camera.near:.9g}"/>
\t\t<float name="farClip" value="{d.camera.far:.9g}"/>
\t\t<transform name="toWorld">
\t\t\t<lookat origin="0.5, 1.1, -2.6" target="0.5, 0.5, 0.5" up="0, 1, 0"/>
\t\t</transform>
\t\t<sampler type="independent">
\t\t\t<integer name="sampleCount" value="$spp"/>
\t\t</sampler>
\t\t<film type="hdrfilm">
\t\t\t<integer name="width" value="$res"/>
\t\t\t<integer name="height" value="$res"/>
\t\t\t<rfilter type="gaussian"/>
\t\t</film>
\t</sensor>
{chr(10).join(shapes)}
</scene>
'''
open(os.path.join(ROOT, "smoke.xml"), "w").write(xml)
print("wrote", os.path.join(ROOT, "smoke.xml"))
