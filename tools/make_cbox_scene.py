#!/usr/bin/env python3
"""Writes scenes/cbox.xml + scenes/meshes/cbox_*.obj: the S1 Cornell-class scene of SURVEY.md 8(d) in Mitsuba 0.6's
XML dialect (cbox.xml itself is not part of the reference tree).  Same data as mitsuba_b200.scene.cornell_box()."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mitsuba_b200.scene import cornell_box

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scenes")
os.makedirs(os.path.join(ROOT, "meshes"), exist_ok=True)
d = cornell_box(1024, 1024)
shapes = []
for m in d.meshes:
    fn = f"meshes/cbox_{m.name}.obj"
    with open(os.path.join(ROOT, fn), "w") as f:
        f.write(f"# {m.name}: Cornell box part (classic Cornell measurements)\n")
        for p in m.P:
            f.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in p))
        for t in m.idx:
            f.write("f %d %d %d\n" % tuple(int(i) + 1 for i in t))
    rgb = " ".join("%.9g" % float(x) for x in m.bsdf.reflectance)
    em = ""
    if m.radiance is not None:
        em = '\n\t\t<emitter type="area">\n\t\t\t<rgb name="radiance" value="%s"/>\n\t\t</emitter>' % " ".join("%.9g" % x for x in m.radiance)
    shapes.append(f'''\t<shape type="obj">
\t\t<string name="filename" value="{fn}"/>
\t\t<boolean name="faceNormals" value="true"/>
\t\t<bsdf type="diffuse">
\t\t\t<rgb name="reflectance" value="{rgb}"/>
\t\t</bsdf>{em}
\t</shape>''')
xml = f'''<?xml version="1.0" encoding="utf-8"?>
<!-- Cornell-box-class scene (S1).  Usage: -D spp=1024 -D res=1024 -->
<scene version="0.5.0">
\t<default name="spp" value="64"/>
\t<default name="res" value="256"/>
\t<default name="rfilter" value="box"/>
\t<integrator type="path">
\t\t<integer name="maxDepth" value="-1"/>
\t\t<integer name="rrDepth" value="5"/>
\t</integrator>
\t<sensor type="perspective">
\t\t<float name="fov" value="39.3077"/>
\t\t<string name="fovAxis" value="x"/>
\t\t<float name="nearClip" value="10"/>
\t\t<float name="farClip" value="2800"/>
\t\t<transform name="toWorld">
\t\t\t<lookat origin="278, 273, -800" target="278, 273, 0" up="0, 1, 0"/>
\t\t</transform>
\t\t<sampler type="sobol">
\t\t\t<integer name="sampleCount" value="$spp"/>
\t\t</sampler>
\t\t<film type="hdrfilm">
\t\t\t<integer name="width" value="$res"/>
\t\t\t<integer name="height" value="$res"/>
\t\t\t<rfilter type="$rfilter"/>
\t\t</film>
\t</sensor>
{chr(10).join(shapes)}
</scene>
'''
open(os.path.join(ROOT, "cbox.xml"), "w").write(xml)
print("wrote", os.path.join(ROOT, "cbox.xml"))
