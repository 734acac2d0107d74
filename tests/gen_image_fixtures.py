"""Writes tests/golden/images/*: small OpenEXR / Radiance files produced by the OpenEXR library and OpenCV's RGBE writer (through cv2), and
what those libraries read back from them (expected.npz).  Run where cv2 has OpenEXR support; the test itself needs neither."""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "images")
os.makedirs(OUT, exist_ok=True)
rng = np.random.default_rng(3)
img = (rng.random((21, 19, 3)) ** 3 * 20).astype(np.float32)
img[5:9, 3:17] = 2.5                      # constant stretches: the run-length coders emit runs
img[0, 0] = (0.0, 1e-7, 70000.0)          # half denormal / overflow territory
expected = {}
for comp, cn in ((0, "none"), (1, "rle"), (2, "zips"), (3, "zip")):
    for typ, tn in ((cv2.IMWRITE_EXR_TYPE_FLOAT, "float"), (cv2.IMWRITE_EXR_TYPE_HALF, "half")):
        name = f"rgb_{cn}_{tn}.exr"
        assert cv2.imwrite(os.path.join(OUT, name), img[..., ::-1], [cv2.IMWRITE_EXR_TYPE, typ, cv2.IMWRITE_EXR_COMPRESSION, comp])
        expected[name] = cv2.imread(os.path.join(OUT, name), cv2.IMREAD_UNCHANGED)[..., ::-1].copy()
assert cv2.imwrite(os.path.join(OUT, "gray_zip_half.exr"), img[..., 1].copy(), [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF])
expected["gray_zip_half.exr"] = cv2.imread(os.path.join(OUT, "gray_zip_half.exr"), cv2.IMREAD_UNCHANGED)[..., None].copy()
assert cv2.imwrite(os.path.join(OUT, "rgb_piz.exr"), img[..., ::-1], [cv2.IMWRITE_EXR_COMPRESSION, 4])
assert cv2.imwrite(os.path.join(OUT, "rgb_rle.hdr"), np.clip(img, 0, 60000)[..., ::-1])
expected["rgb_rle.hdr"] = cv2.imread(os.path.join(OUT, "rgb_rle.hdr"), cv2.IMREAD_UNCHANGED)[..., ::-1].copy()
# a small latitude-longitude map for the <emitter type="envmap"> scene-file test (values inside the half range)
sky = (rng.random((16, 32, 3)) ** 3 * 0.8).astype(np.float32)
sky[2:4, 10:13] += 20.0
sky[8:] *= 0.2
assert cv2.imwrite(os.path.join(OUT, "sky_zip_half.exr"), sky[..., ::-1], [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF, cv2.IMWRITE_EXR_COMPRESSION, 3])
expected["sky_zip_half.exr"] = cv2.imread(os.path.join(OUT, "sky_zip_half.exr"), cv2.IMREAD_UNCHANGED)[..., ::-1].copy()
np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)
print(sorted(os.listdir(OUT)), sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)), "bytes")
