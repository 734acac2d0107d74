"""BSDF fixtures: the subset of data/tests/test_bsdf.xml (reference) that lies on the hot path
(diffuse, roughdielectric beckmann/phong/ggx/as, roughconductor beckmann/as, coating over diffuse /
roughconductor), with the same parameters.  roughconductor's default material "Cu" needs the
.spd -> RGB route (out of scope); copper-/gold-like RGB constants stand in."""
from mitsuba_b200.scene import Bsdf

CU = dict(eta=(0.2004, 0.9240, 1.1022), k=(3.9129, 2.4528, 2.1421))
AU = dict(eta=(0.1431, 0.3749, 1.4424), k=(3.9831, 2.3857, 1.6032))


def configs():
    c = {}
    c["diffuse"] = Bsdf("diffuse")
    for dist in ("beckmann", "phong", "ggx"):
        c[f"roughdielectric_{dist}"] = Bsdf("roughdielectric", distribution=dist, alpha_u=0.3, alpha_v=0.3, int_ior=1.5, ext_ior=1.0)
    c["roughdielectric_as"] = Bsdf("roughdielectric", distribution="as", alpha_u=0.1, alpha_v=0.3, int_ior=1.5, ext_ior=1.0)
    c["roughconductor_beckmann"] = Bsdf("roughconductor", distribution="beckmann", alpha_u=0.3, alpha_v=0.3, **CU)
    c["roughconductor_as"] = Bsdf("roughconductor", distribution="as", alpha_u=0.1, alpha_v=0.3, **AU)
    c["roughconductor_ggx"] = Bsdf("roughconductor", distribution="ggx", alpha_u=0.1, alpha_v=0.1, **CU)
    c["roughconductor_ggx_all"] = Bsdf("roughconductor", distribution="ggx", alpha_u=0.2, alpha_v=0.2, sample_visible=False, **CU)
    c["coating_diffuse"] = Bsdf("coating", int_ior=1.5, ext_ior=1.0, sigma_a=(0.1, 0.2, 0.3), thickness=2.0, nested=Bsdf("diffuse"))
    c["coating_roughconductor"] = Bsdf("coating", int_ior=1.5, ext_ior=1.0, nested=Bsdf("roughconductor", **CU))
    # f-3 plugins, parameters of data/tests/test_bsdf.xml where it has them (dielectric intIOR 1.5/extIOR 1.0, plastic defaults)
    c["dielectric"] = Bsdf("dielectric", int_ior=1.5, ext_ior=1.0)
    c["conductor"] = Bsdf("conductor", **CU)
    c["plastic"] = Bsdf("plastic", int_ior=1.5, ext_ior=1.0, diffuse_reflectance=(0.4, 0.5, 0.2))
    c["plastic_nonlinear"] = Bsdf("plastic", nonlinear=True, diffuse_reflectance=(0.7, 0.3, 0.2), specular_reflectance=(0.8, 0.8, 0.8))
    c["twosided_diffuse"] = Bsdf("twosided", nested=Bsdf("diffuse", reflectance=(0.6, 0.4, 0.3)))
    c["twosided_two"] = Bsdf("twosided", nested=Bsdf("roughconductor", distribution="ggx", alpha_u=0.2, alpha_v=0.2, **CU),
                             nested_back=Bsdf("plastic", diffuse_reflectance=(0.2, 0.3, 0.6)))
    c["twosided_coating"] = Bsdf("twosided", nested=Bsdf("coating", int_ior=1.5, ext_ior=1.0, nested=Bsdf("diffuse")))
    return c


def flatten(b):
    """-> (flat list of dicts, id of b)"""
    from mitsuba_b200.scene import SceneDesc, Mesh
    import numpy as np
    m = Mesh(np.zeros((3, 3), np.float32), np.array([[0, 1, 2]], np.uint32), bsdf=b)
    flat, ids = SceneDesc([m]).flat_bsdfs()
    return flat, ids[0]
