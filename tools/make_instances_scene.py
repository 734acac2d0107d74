#!/usr/bin/env python3
"""Writes scenes/instances.xml + scenes/meshes/inst_*.obj: a small S3-class scene in Mitsuba 0.6's XML dialect --
one bumpy sphere in a `shapegroup`, placed 9 times with `instance` shapes (scale + translate), a ground plane and an area light.
Same data as mitsuba_b200.scene.stress_scene(9, 32, 32, instanced=True)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mitsuba_b200.scene import stress_scene

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scenes")
os.makedirs(os.path.join(ROOT, "meshes"), exist_ok=True)
d = stress_scene(9, 32, 32, 512, 512, instanced=True)


def write_obj(fn, m):
    with open(os.path.join(ROOT, fn), "w") as f:
        f.write(f"# {m.name}\n")
        for p in m.P:
            f.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in p))
        if m.N is not None:
            for n in m.N:
                f.write("vn %.9g %.9g %.9g\n" % tuple(float(x) for x in n))
        for t in m.idx:
            if m.N is not None:
                f.write("f %d//%d %d//%d %d//%d\n" % tuple(int(i) + 1 for i in (t[0], t[0], t[1], t[1], t[2], t[2])))
            else:
                f.write("f %d %d %d\n" % tuple(int(i) + 1 for i in t))


def bsdf_xml(m, ind):
    rgb = " ".join("%.9g" % float(x) for x in m.bsdf.reflectance)
    return f'{ind}<bsdf type="diffuse">\n{ind}\t<rgb name="reflectance" value="{rgb}"/>\n{ind}</bsdf>'


parts = []
for m in d.meshes:
    fn = f"meshes/inst_{m.name}.obj"
    write_obj(fn, m)
    if m.group >= 0:
        parts.append(f'\t<shape type="shapegroup" id="group{m.group}">\n\t\t<shape type="obj">\n\t\t\t<string name="filename" value="{fn}"/>\n{bsdf_xml(m, chr(9) * 3)}\n\t\t</shape>\n\t</shape>')
    else:
        em = ""
        if m.radiance is not None:
            em = '\n\t\t<emitter type="area">\n\t\t\t<rgb name="radiance" value="%s"/>\n\t\t</emitter>' % " ".join("%.9g" % x for x in m.radiance)
        parts.append(f'\t<shape type="obj">\n\t\t<string name="filename" value="{fn}"/>\n\t\t<boolean name="faceNormals" value="true"/>\n{bsdf_xml(m, chr(9) * 2)}{em}\n\t</shape>')
for inst in d.instances:
    M = " ".join("%.9g" % float(x) for x in inst.to_world.reshape(-1))
    parts.append(f'\t<shape type="instance">\n\t\t<ref id="group{inst.group}"/>\n\t\t<transform name="toWorld">\n\t\t\t<matrix value="{M}"/>\n\t\t</transform>\n\t</shape>')
cam = d.camera
xml = f'''<?xml version="1.0" encoding="utf-8"?>
<!-- Instanced scene (S3 class, small): one shapegroup, nine instances.  Usage: -D spp=64 -D res=512 -->
<scene version="0.5.0">
\t<default name="spp" value="16"/>
\t<default name="res" value="256"/>
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
open(os.path.join(ROOT, "instances.xml"), "w").write(xml)
print("wrote", os.path.join(ROOT, "instances.xml"))
