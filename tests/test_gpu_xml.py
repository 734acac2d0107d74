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
