"""The scene-file front end (b2_load_xml) against the Python scene description of the same scene."""
import os

import numpy as np
import pytest

from mitsuba_b200 import api
from mitsuba_b200.scene import RenderParams, cornell_box
from oracle import oracle_api as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def test_cbox_xml_matches_python_scene(b2ctx):
    sc, rp = b2ctx.load_xml(os.path.join(ROOT, "scenes", "cbox.xml"), ["spp=16", "res=64"])
    assert rp.spp == 16 and rp.sampler == "sobol" and rp.rfilter == "box" and rp.max_depth == -1 and rp.rr_depth == 5
    film, st = sc.render(rp, parity=True, width=64, height=64)
    fo, so = O.OracleScene(cornell_box(64, 64), sample_to_camera=sc.sample_to_camera()).render(RenderParams(spp=16, sampler="sobol", rfilter="box"))
    assert st["n_triangles"] == 32
    assert rel_l2(api.develop(film), O.develop(fo)) < 2e-4
    assert abs(st["rays"] - so["rays"]) <= 1e-4 * so["rays"]


def test_xml_errors(b2ctx, tmp_path):
    p = tmp_path / "bad.xml"
    p.write_text('<scene version="0.5.0"><integrator type="bdpt"/></scene>')
    with pytest.raises(api.B2Error, match="unsupported integrator"):
        b2ctx.load_xml(str(p))
    p.write_text('<scene version="0.5.0"><sensor type="perspective"><float name="fovv" value="3"/></sensor></scene>')
    with pytest.raises(api.B2Error, match="unreferenced property"):
        b2ctx.load_xml(str(p))
    p.write_text('<scene version="0.5.0"><sensor type="perspective"><float name="fov" value="$undefined"/></sensor></scene>')
    with pytest.raises(api.B2Error, match="undefined parameter"):
        b2ctx.load_xml(str(p))


def test_smoke_xml_matches_python_scene(b2ctx):
    """volpath + heterogeneous medium + gridvolume (.vol) + constvolume + phase + <ref name="interior"> through b2_load_xml."""
    from mitsuba_b200.scene import smoke_scene
    sc, rp = b2ctx.load_xml(os.path.join(ROOT, "scenes", "smoke.xml"), ["spp=16", "res=48"])
    assert rp.integrator == "volpath" and rp.sampler == "independent" and rp.rfilter == "gaussian" and rp.spp == 16
    film, st = sc.render(rp, parity=True, width=48, height=48)
    d = smoke_scene(48, 48, res=64)
    fo, so = O.OracleScene(d, sample_to_camera=sc.sample_to_camera()).render(rp)
    assert st["n_triangles"] == d.n_triangles()
    assert rel_l2(api.develop(film), O.develop(fo)) < 3e-4
    assert abs(st["rays"] - so["rays"]) <= 2e-4 * so["rays"]


def test_medium_xml_errors(b2ctx, tmp_path):
    p = tmp_path / "bad.xml"
    head = '<scene version="0.5.0"><sensor type="perspective"/>'
    p.write_text(head + '<medium type="heterogeneous" id="m"><string name="method" value="simpson"/></medium></scene>')
    with pytest.raises(api.B2Error, match="Unsupported integration method"):
        b2ctx.load_xml(str(p))
    p.write_text(head + '<medium type="heterogeneous" id="m"/></scene>')
    with pytest.raises(api.B2Error, match="No density specified"):
        b2ctx.load_xml(str(p))
    p.write_text(head + '<medium type="homogeneous" id="m"><rgb name="sigmaS" value="1,1,1"/><rgb name="sigmaT" value="2,2,2"/></medium></scene>')
    with pytest.raises(api.B2Error, match="no other combinations"):
        b2ctx.load_xml(str(p))
    p.write_text(head + '<shape type="cube"><ref name="interior" id="nope"/></shape></scene>')
    with pytest.raises(api.B2Error, match="not found"):
        b2ctx.load_xml(str(p))


def test_f3_bsdf_plugins_through_xml(b2ctx, tmp_path):
    """twosided / dielectric / conductor / plastic in a scene file vs the same scene built in Python."""
    import shutil
    from mitsuba_b200.scene import Bsdf
    shutil.copytree(os.path.join(ROOT, "scenes", "meshes"), tmp_path / "meshes")
    xml = open(os.path.join(ROOT, "scenes", "cbox.xml")).read()
    repl = {"cbox_walls": '<bsdf type="twosided"><bsdf type="diffuse"><rgb name="reflectance" value="0.73 0.73 0.73"/></bsdf></bsdf>',
            "cbox_short": '<bsdf type="plastic"><rgb name="diffuseReflectance" value="0.1 0.27 0.36"/><float name="intIOR" value="1.9"/></bsdf>',
            "cbox_tall": '<bsdf type="conductor"><string name="material" value="none"/><rgb name="eta" value="0.2 0.92 1.1"/><rgb name="k" value="3.9 2.45 2.14"/></bsdf>',
            "cbox_right": '<bsdf type="dielectric"><string name="intIOR" value="water"/></bsdf>'}
    import re
    for key, b in repl.items():
        xml = re.sub(r'(<string name="filename" value="meshes/%s.obj"/>\s*<boolean name="faceNormals" value="true"/>\s*)<bsdf type="diffuse">.*?</bsdf>' % key,
                     lambda m: m.group(1) + b, xml, count=1, flags=re.S)
    p = tmp_path / "zoo.xml"
    p.write_text(xml)
    sc, rp = b2ctx.load_xml(str(p), ["spp=16", "res=48"])
    film, st = sc.render(rp, parity=True, width=48, height=48)
    d = cornell_box(48, 48)
    by = {m.name: m for m in d.meshes}
    by["walls"].bsdf = Bsdf("twosided", nested=Bsdf("diffuse", reflectance=(0.73, 0.73, 0.73)))
    by["short"].bsdf = Bsdf("plastic", diffuse_reflectance=(0.1, 0.27, 0.36), int_ior=1.9)
    by["tall"].bsdf = Bsdf("conductor", eta=(0.2, 0.92, 1.1), k=(3.9, 2.45, 2.14))
    by["right"].bsdf = Bsdf("dielectric", int_ior="water")
    fo, so = O.OracleScene(d, sample_to_camera=sc.sample_to_camera()).render(RenderParams(spp=16, sampler="sobol", rfilter="box"))
    assert rel_l2(api.develop(film), O.develop(fo)) < 5e-4


def test_constant_emitter_through_xml(b2ctx, tmp_path):
    import shutil
    shutil.copytree(os.path.join(ROOT, "scenes", "meshes"), tmp_path / "meshes")
    xml = open(os.path.join(ROOT, "scenes", "cbox.xml")).read().replace("</scene>", '\t<emitter type="constant"><rgb name="radiance" value="0.2 0.3 0.5"/><float name="samplingWeight" value="0.5"/></emitter>\n</scene>')
    p = tmp_path / "sky.xml"
    p.write_text(xml)
    sc, rp = b2ctx.load_xml(str(p), ["spp=16", "res=48"])
    film, _ = sc.render(rp, parity=True, width=48, height=48)
    d = cornell_box(48, 48)
    d.env_radiance = (0.2, 0.3, 0.5); d.env_sampling_weight = 0.5
    fo, _ = O.OracleScene(d, sample_to_camera=sc.sample_to_camera()).render(RenderParams(spp=16, sampler="sobol", rfilter="box"))
    assert rel_l2(api.develop(film), O.develop(fo)) < 3e-4


def _write_ply(path, P, N, idx, fmt):
    import struct
    n, m = len(P), len(idx)
    hdr = f"ply\nformat {fmt} 1.0\ncomment generated by tests/test_gpu_xml.py\nelement vertex {n}\nproperty float x\nproperty float y\nproperty float z\n"
    if N is not None:
        hdr += "property float nx\nproperty float ny\nproperty float nz\n"
    hdr += f"property uchar red\nelement face {m}\nproperty list uchar int vertex_indices\nend_header\n"
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if fmt == "ascii":
            for i in range(n):
                row = list(P[i]) + (list(N[i]) if N is not None else []) + [255]
                f.write((" ".join(repr(float(x)) if k < len(row) - 1 else str(int(x)) for k, x in enumerate(row)) + "\n").encode())
            for t in idx:
                f.write((f"{len(t)} " + " ".join(str(int(i)) for i in t) + "\n").encode())
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            for i in range(n):
                f.write(struct.pack(e + "3f", *P[i]))
                if N is not None:
                    f.write(struct.pack(e + "3f", *N[i]))
                f.write(struct.pack("B", 255))
            for t in idx:
                f.write(struct.pack("B", len(t)) + struct.pack(e + f"{len(t)}i", *[int(i) for i in t]))


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_ply_shape_through_xml(b2ctx, tmp_path, fmt):
    """<shape type="ply"> (src/shapes/ply.cpp): ascii / binary, vertex normals, an extra vertex property, a quad face."""
    from mitsuba_b200.scene import Bsdf, Camera, Mesh, SceneDesc, look_at, uv_sphere, _quad
    P, N, _, I = uv_sphere((0, 1, 0), 1.0, 16, 32)
    _write_ply(tmp_path / "ball.ply", P, N, [tuple(t) for t in I], fmt)
    Pq = np.float32([(-4, 0, -4), (-4, 0, 4), (4, 0, 4), (4, 0, -4)])
    _write_ply(tmp_path / "ground.ply", Pq, None, [(0, 1, 2, 3)], fmt)
    Pl, Il = _quad([(-1, 4, -1), (-1, 4, 1), (1, 4, 1), (1, 4, -1)], (0, -1, 0))
    _write_ply(tmp_path / "light.ply", Pl, None, [tuple(t) for t in Il], fmt)
    (tmp_path / "s.xml").write_text('''<scene version="0.5.0"><integrator type="path"/>
      <sensor type="perspective"><float name="fov" value="35"/><transform name="toWorld"><lookat origin="0,2.2,-5" target="0,0.9,0" up="0,1,0"/></transform>
        <sampler type="sobol"><integer name="sampleCount" value="16"/></sampler>
        <film type="hdrfilm"><integer name="width" value="40"/><integer name="height" value="40"/><rfilter type="box"/></film></sensor>
      <shape type="ply"><string name="filename" value="ball.ply"/><bsdf type="diffuse"><rgb name="reflectance" value="0.7 0.4 0.3"/></bsdf></shape>
      <shape type="ply"><string name="filename" value="ground.ply"/><boolean name="faceNormals" value="true"/><bsdf type="diffuse"/></shape>
      <shape type="ply"><string name="filename" value="light.ply"/><boolean name="faceNormals" value="true"/><emitter type="area"><rgb name="radiance" value="20 20 20"/></emitter></shape>
    </scene>''')
    sc, rp = b2ctx.load_xml(str(tmp_path / "s.xml"))
    film, st = sc.render(rp, parity=True, width=40, height=40)
    ground_idx = np.uint32([(0, 1, 2), (3, 0, 2)])  # ply.cpp:276-287
    d = SceneDesc([Mesh(P, I, N=N, bsdf=Bsdf("diffuse", reflectance=(0.7, 0.4, 0.3))), Mesh(Pq, ground_idx, bsdf=Bsdf("diffuse")),
                   Mesh(Pl, Il, bsdf=Bsdf("diffuse", reflectance=(0, 0, 0)), radiance=(20, 20, 20))],
                  Camera(look_at((0, 2.2, -5), (0, 0.9, 0), (0, 1, 0)), fov=35, near=1e-2, far=1e4, width=40, height=40))
    fo, _ = O.OracleScene(d, sample_to_camera=sc.sample_to_camera()).render(RenderParams(spp=16, sampler="sobol", rfilter="box"))
    assert st["n_triangles"] == d.n_triangles()
    assert rel_l2(api.develop(film), O.develop(fo)) < 3e-4


def _serialized_blob(version, P, N, idx, double=False, face_normals=False, name=b"mesh"):
    """One mesh record of src/librender/trimesh.cpp:175-250 (header + zlib stream)."""
    import struct, zlib
    flags = (0x2000 if double else 0x1000) | (0x0001 if N is not None else 0) | (0x0010 if face_normals else 0)
    ft = "<f8" if double else "<f4"
    body = struct.pack("<I", flags) + ((name + b"\0") if version == 4 else b"") + struct.pack("<QQ", len(P), len(idx))
    body += np.ascontiguousarray(P, ft).tobytes()
    if N is not None:
        body += np.ascontiguousarray(N, ft).tobytes()
    body += np.ascontiguousarray(idx, "<u4").tobytes()
    return struct.pack("<HH", 0x041C, version) + zlib.compress(body)


def test_serialized_shape_through_xml(b2ctx, tmp_path):
    """<shape type="serialized"> (src/shapes/serialized.cpp): v4 multi-shape file with the offset dictionary, v3 single shape,
    double precision, the face-normal flag, shapeIndex."""
    import struct
    from mitsuba_b200.scene import Bsdf, Camera, Mesh, SceneDesc, look_at, uv_sphere, _quad
    P, N, _, I = uv_sphere((0, 1, 0), 1.0, 16, 32)
    Pq, Iq = _quad([(-4, 0, -4), (-4, 0, 4), (4, 0, 4), (4, 0, -4)], (0, 1, 0))
    Pl, Il = _quad([(-1, 4, -1), (-1, 4, 1), (1, 4, 1), (1, 4, -1)], (0, -1, 0))
    blobs = [_serialized_blob(4, P, N, I, double=True), _serialized_blob(4, Pq, None, Iq, face_normals=True)]
    offs, data = [], b""
    for b in blobs:
        offs.append(len(data)); data += b
    data += b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<I", len(blobs))
    (tmp_path / "scene.serialized").write_bytes(data)
    (tmp_path / "light.serialized").write_bytes(_serialized_blob(3, Pl, None, Il) + struct.pack("<II", 0, 1))
    (tmp_path / "s.xml").write_text('''<scene version="0.5.0"><integrator type="path"/>
      <sensor type="perspective"><float name="fov" value="35"/><transform name="toWorld"><lookat origin="0,2.2,-5" target="0,0.9,0" up="0,1,0"/></transform>
        <sampler type="sobol"><integer name="sampleCount" value="16"/></sampler>
        <film type="hdrfilm"><integer name="width" value="40"/><integer name="height" value="40"/><rfilter type="box"/></film></sensor>
      <shape type="serialized"><string name="filename" value="scene.serialized"/><bsdf type="diffuse"><rgb name="reflectance" value="0.7 0.4 0.3"/></bsdf></shape>
      <shape type="serialized"><string name="filename" value="scene.serialized"/><integer name="shapeIndex" value="1"/><bsdf type="diffuse"/></shape>
      <shape type="serialized"><string name="filename" value="light.serialized"/><boolean name="faceNormals" value="true"/><emitter type="area"><rgb name="radiance" value="20 20 20"/></emitter></shape>
    </scene>''')
    sc, rp = b2ctx.load_xml(str(tmp_path / "s.xml"))
    film, st = sc.render(rp, parity=True, width=40, height=40)
    d = SceneDesc([Mesh(P, I, N=N, bsdf=Bsdf("diffuse", reflectance=(0.7, 0.4, 0.3))), Mesh(Pq, Iq, bsdf=Bsdf("diffuse")),
                   Mesh(Pl, Il, bsdf=Bsdf("diffuse", reflectance=(0, 0, 0)), radiance=(20, 20, 20))],
                  Camera(look_at((0, 2.2, -5), (0, 0.9, 0), (0, 1, 0)), fov=35, near=1e-2, far=1e4, width=40, height=40))
    fo, _ = O.OracleScene(d, sample_to_camera=sc.sample_to_camera()).render(RenderParams(spp=16, sampler="sobol", rfilter="box"))
    assert st["n_triangles"] == d.n_triangles()
    assert rel_l2(api.develop(film), O.develop(fo)) < 3e-4
    (tmp_path / "bad.xml").write_text((tmp_path / "s.xml").read_text().replace('name="shapeIndex" value="1"', 'name="shapeIndex" value="7"'))
    with pytest.raises(api.B2Error, match="out of range"):
        b2ctx.load_xml(str(tmp_path / "bad.xml"))


def test_instances_xml_matches_python_scene(b2ctx):
    """shapegroup + instance (<ref id>, toWorld matrix) through b2_load_xml (scenes/instances.xml, tools/make_instances_scene.py)."""
    from mitsuba_b200.scene import stress_scene
    sc, rp = b2ctx.load_xml(os.path.join(ROOT, "scenes", "instances.xml"), ["spp=16", "res=64"])
    film, st = sc.render(rp, parity=True)
    ref = api.Scene(b2ctx, stress_scene(9, 32, 32, 64, 64, instanced=True))
    f2, s2 = ref.render(RenderParams(spp=16, sampler="sobol", rfilter="box"), parity=True)
    assert st["n_triangles"] == s2["n_triangles"]
    assert rel_l2(api.develop(film), api.develop(f2)) < 1e-3
    assert abs(st["rays"] - s2["rays"]) <= 1e-3 * s2["rays"]


def test_conductor_material_presets_and_crop_window_through_the_scene_file(b2ctx, tmp_path):
    """<bsdf type="roughconductor"> with its DEFAULT material (the plugin's material="Cu": data/ior/Cu.{eta,k}.spd -> RGB, roughconductor.cpp:174-190)
    and a named preset render like the same scene with explicit RGB eta / k; <film> cropOffsetX/Y + cropWidth/Height (film.cpp:36-47)."""
    from mitsuba_b200.scene import conductor_preset
    tmpl = '''<scene version="0.5.0"><integrator type="path"/>
      <sensor type="perspective"><float name="fov" value="40"/>
        <transform name="toWorld"><lookat origin="0, 1.5, -4" target="0, 0.5, 0" up="0, 1, 0"/></transform>
        <sampler type="sobol"><integer name="sampleCount" value="16"/></sampler>
        <film type="hdrfilm"><integer name="width" value="64"/><integer name="height" value="48"/>%s<rfilter type="box"/></film></sensor>
      <bsdf type="roughconductor" id="metal"><float name="alpha" value="0.2"/><string name="distribution" value="ggx"/>%s</bsdf>
      <shape type="obj"><string name="filename" value="%s"/><ref id="metal"/></shape>
      <shape type="obj"><string name="filename" value="%s"/><bsdf type="diffuse"/><emitter type="area"><rgb name="radiance" value="12, 12, 12"/></emitter></shape>
    </scene>'''
    ball, light = tmp_path / "ball.obj", tmp_path / "light.obj"
    ball.write_text("v -1 0 -1\nv 1 0 -1\nv 1 0 1\nv -1 0 1\nv 0 1.4 0\nf 1 2 3\nf 1 3 4\nf 1 5 2\nf 2 5 3\nf 3 5 4\nf 4 5 1\n")
    light.write_text("v -1 3 -1\nv 1 3 -1\nv 1 3 1\nv -1 3 1\nf 1 2 3\nf 1 3 4\n")   # facing down

    def render(film_extra, bsdf_extra):
        p = tmp_path / "s.xml"
        p.write_text(tmpl % (film_extra, bsdf_extra, ball, light))
        sc, rp = b2ctx.load_xml(str(p))
        return sc.render(rp, parity=True)[0], sc

    for material, explicit in (("", conductor_preset("Cu")), ('<string name="material" value="Au"/>', conductor_preset("Au"))):
        a, _ = render("", material)
        eta, k = explicit
        b, _ = render("", '<string name="material" value="none"/><rgb name="eta" value="%r, %r, %r"/><rgb name="k" value="%r, %r, %r"/>' % (*eta, *k))
        assert a[..., :3].max() > 0.05 and rel_l2(api.develop(a), api.develop(b)) < 1e-5
    with pytest.raises(api.B2Error, match="unknown material preset"):
        render("", '<string name="material" value="Unobtainium"/>')
    crop, sc = render('<integer name="cropOffsetX" value="8"/><integer name="cropOffsetY" value="16"/><integer name="cropWidth" value="40"/><integer name="cropHeight" value="24"/>', "")
    assert crop.shape == (24, 40, 5) and (sc.W, sc.H) == (40, 24)
    with pytest.raises(api.B2Error, match="Invalid crop window"):
        render('<integer name="cropOffsetX" value="60"/><integer name="cropWidth" value="40"/>', "")


def _rgbe_bytes(img, rle):
    """Radiance .hdr encoding of a float RGB image; returns (file bytes, the image a reader decodes from them)."""
    h, w, _ = img.shape
    m = img.max(axis=2)
    e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, 0).astype(np.int32)
    scale = np.where(m > 1e-32, np.ldexp(1.0, 8 - e), 0.0)
    mant = np.clip(np.floor(img * scale[..., None]), 0, 255).astype(np.uint8)
    ebyte = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    rgbe = np.concatenate([mant, ebyte[..., None]], axis=2)
    decoded = np.where(ebyte[..., None] > 0, mant.astype(np.float32) * np.ldexp(np.float32(1.0), ebyte.astype(np.int32) - 136)[..., None], 0).astype(np.float32)
    out = bytearray(b"#?RADIANCE\n# written by tests/test_gpu_xml.py\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n" + f"-Y {h} +X {w}\n".encode())
    if not rle:
        out += rgbe.tobytes()
        return bytes(out), decoded
    for y in range(h):
        out += bytes([2, 2, w >> 8, w & 255])
        for c in range(4):
            row = rgbe[y, :, c]
            x = 0
            while x < w:
                run = 1
                while x + run < w and run < 127 and row[x + run] == row[x]:
                    run += 1
                if run >= 3:
                    out += bytes([128 + run, int(row[x])]); x += run
                else:
                    lit = min(w - x, 5)   # short literal packets
                    out += bytes([lit]) + row[x:x + lit].tobytes(); x += lit
    return bytes(out), decoded


@pytest.mark.parametrize("encoding", ["rle", "flat", "pfm", "exr"])
def test_envmap_emitter_through_the_scene_file(b2ctx, tmp_path, encoding):
    """<emitter type="envmap"> with a Radiance .hdr (run-length coded and flat) or PFM image, scale, toWorld and samplingWeight: the film of
    the loaded scene equals the film of the same scene handed over through the C-ABI, and the reference-pinned oracle's film."""
    import shutil
    import ref_pins
    from mitsuba_b200.scene import EnvMap
    shutil.copytree(os.path.join(ROOT, "scenes", "meshes"), tmp_path / "meshes")
    img = ref_pins.sky_image(32, 16, seed=21, sun=20.0)
    img[:, 8:14] = img[:, 8:9]                      # a few constant stretches so that the scanline coder emits runs
    if encoding == "exr":   # a file written by the OpenEXR library (ZIP, half), 32 x 16 texels: tests/golden/images
        fname = "sky.exr"
        shutil.copy(os.path.join(ROOT, "tests", "golden", "images", "sky_zip_half.exr"), tmp_path / fname)
        decoded = np.load(os.path.join(ROOT, "tests", "golden", "images", "expected.npz"))["sky_zip_half.exr"]
    elif encoding == "pfm":
        fname, decoded = "sky.pfm", img
        with open(tmp_path / fname, "wb") as f:
            f.write(b"PF\n32 16\n-1.0\n" + img[::-1].astype("<f4").tobytes())
    else:
        fname = "sky.hdr"
        data, decoded = _rgbe_bytes(img, encoding == "rle")
        (tmp_path / fname).write_bytes(data)
    emitter = (f'\t<emitter type="envmap"><string name="filename" value="{fname}"/><float name="scale" value="0.8"/><float name="samplingWeight" value="2"/>'
               '<transform name="toWorld"><rotate y="1" angle="35"/></transform></emitter>\n')
    p = tmp_path / "sky.xml"
    p.write_text(open(os.path.join(ROOT, "scenes", "cbox.xml")).read().replace("</scene>", emitter + "</scene>"))
    sc, rp = b2ctx.load_xml(str(p), ["spp=16", "res=48"])
    film, _ = sc.render(rp, parity=True, width=48, height=48)
    a = np.deg2rad(35.0)
    M = np.eye(4); M[0, 0] = M[2, 2] = np.cos(a); M[0, 2] = np.sin(a); M[2, 0] = -np.sin(a)
    d = cornell_box(48, 48)
    d.envmap = EnvMap(pixels=decoded, scale=0.8, to_world=M.astype(np.float32), sampling_weight=2.0)
    sc2 = api.Scene(b2ctx, d)
    film2, _ = sc2.render(RenderParams(spp=16, sampler="sobol", rfilter="box"), parity=True)
    assert rel_l2(api.develop(film), api.develop(film2)) < 1e-5
    fo, _ = O.OracleScene(d, sample_to_camera=sc.sample_to_camera()).render(RenderParams(spp=16, sampler="sobol", rfilter="box"))
    assert rel_l2(api.develop(film), O.develop(fo)) < 1e-3
    sc.close(); sc2.close()


def test_spectra_from_spd_files_and_wavelength_lists_through_the_scene_file(b2ctx, tmp_path):
    """<spectrum filename="x.spd"> and <spectrum value="l:v, ..."> (scenehandler.cpp:557-611): the loaded scene renders like the same scene with the
    RGB colours the host-only conversion (tests/test_spectrum.py: held to the reference's converter) gives for those samples."""
    import shutil
    shutil.copytree(os.path.join(ROOT, "scenes", "meshes"), tmp_path / "meshes")
    wl = np.array([380, 450, 520, 590, 660, 730], np.float32)
    green = np.array([0.05, 0.1, 0.6, 0.35, 0.1, 0.05], np.float32)
    red = np.array([0.04, 0.05, 0.06, 0.3, 0.7, 0.75], np.float32)
    (tmp_path / "red.spd").write_text("# wavelength (nm)  reflectance\n" + "".join(f"{w:g} {v:g}\n" for w, v in zip(wl, red)))
    xml = open(os.path.join(ROOT, "scenes", "cbox.xml")).read()
    xml = xml.replace('<rgb name="reflectance" value="0.14 0.45 0.091"/>', '<spectrum name="reflectance" value="' + ", ".join(f"{w:g}:{v:g}" for w, v in zip(wl, green)) + '"/>')
    xml = xml.replace('<rgb name="reflectance" value="0.63 0.065 0.05"/>', '<spectrum name="reflectance" filename="red.spd"/>')
    p = tmp_path / "spectral.xml"
    p.write_text(xml)
    sc, rp = b2ctx.load_xml(str(p), ["spp=16", "res=48"])
    film, _ = sc.render(rp, parity=True, width=48, height=48)
    d = cornell_box(48, 48)
    by = {m.name: m for m in d.meshes}
    by["right"].bsdf.reflectance = tuple(float(v) for v in api.spectrum_to_rgb(wl, green))
    by["left"].bsdf.reflectance = tuple(float(v) for v in api.spectrum_to_rgb(wl, red))
    assert by["right"].bsdf.reflectance[1] > by["right"].bsdf.reflectance[0] and by["left"].bsdf.reflectance[0] > by["left"].bsdf.reflectance[1]
    fo, _ = O.OracleScene(d, sample_to_camera=sc.sample_to_camera()).render(RenderParams(spp=16, sampler="sobol", rfilter="box"))
    assert rel_l2(api.develop(film), O.develop(fo)) < 3e-4
    sc.close()
