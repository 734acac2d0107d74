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
# PNG: files written by libpng (through cv2) hold the decoded integer samples as the expectation; three more are assembled by hand for what
# cv2 does not write: a 4-bit palette, 2-bit grey, and a gAMA chunk, each scan line with a different filter type
import struct, zlib
png_expected = {}
for name, arr in (("rgb8", rng.integers(0, 256, (13, 17, 3), dtype=np.uint8)), ("rgb16", rng.integers(0, 65536, (13, 17, 3), dtype=np.uint16)),
                  ("gray8", rng.integers(0, 256, (13, 17), dtype=np.uint8)), ("gray16", rng.integers(0, 65536, (13, 17), dtype=np.uint16)),
                  ("rgba8", rng.integers(0, 256, (13, 17, 4), dtype=np.uint8))):
    assert cv2.imwrite(os.path.join(OUT, name + ".png"), arr)
    back = cv2.imread(os.path.join(OUT, name + ".png"), cv2.IMREAD_UNCHANGED)
    png_expected[name + ".png"] = (back[..., :3][..., ::-1] if back.ndim == 3 else back[..., None]).copy()
xs, ys = np.linspace(0, 255, 40)[None, :], np.linspace(0, 255, 24)[:, None]
smooth = np.stack([xs + 0 * ys, 0 * xs + ys, (xs + ys) / 2], -1).astype(np.uint8)   # libpng picks sub / up / average / paeth filters on this one
assert cv2.imwrite(os.path.join(OUT, "smooth8.png"), smooth[..., ::-1])
png_expected["smooth8.png"] = smooth


def chunk(kind, data):
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)


def filtered(rows, bpp):
    """PNG scan lines with filter types 0, 1, 2, 3, 4 in turn."""
    out, prev = bytearray(), bytes(len(rows[0]))
    for y, row in enumerate(rows):
        ft = y % 5
        enc = bytearray()
        for i, v in enumerate(row):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            if ft == 0: pred = 0
            elif ft == 1: pred = a
            elif ft == 2: pred = b
            elif ft == 3: pred = (a + b) >> 1
            else:
                pq = a + b - c
                pa, pb, pc = abs(pq - a), abs(pq - b), abs(pq - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            enc.append((v - pred) & 255)
        out += bytes([ft]) + enc
        prev = bytes(row)
    return bytes(out)


def write_png(name, w, h, depth, ctype, rows, bpp, extra=b""):
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)) + extra + \
           chunk(b"IDAT", zlib.compress(filtered(rows, bpp))) + chunk(b"IEND", b"")
    open(os.path.join(OUT, name), "wb").write(data)


pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
idx = rng.integers(0, 16, (9, 11), dtype=np.uint8)
rows = [bytes(int(r[2 * k]) << 4 | (int(r[2 * k + 1]) if 2 * k + 1 < len(r) else 0) for k in range((len(r) + 1) // 2)) for r in idx]
write_png("palette4.png", 11, 9, 4, 3, rows, 1, chunk(b"PLTE", pal.tobytes()))
png_expected["palette4.png"] = pal[idx]
g2 = rng.integers(0, 4, (7, 10), dtype=np.uint8)
rows = [bytes(sum(int(r[4 * k + j]) << (6 - 2 * j) for j in range(4) if 4 * k + j < len(r)) for k in range((len(r) + 3) // 4)) for r in g2]
write_png("gray2.png", 10, 7, 2, 0, rows, 1)
png_expected["gray2.png"] = (g2 * 85)[..., None].astype(np.uint8)
ga = rng.integers(0, 256, (6, 8, 3), dtype=np.uint8)
write_png("rgb8_gama.png", 8, 6, 8, 2, [r.tobytes() for r in ga], 3, chunk(b"gAMA", struct.pack(">I", 45455)))   # file gamma 1 / 2.2
png_expected["rgb8_gama.png"] = ga
np.savez_compressed(os.path.join(OUT, "png_expected.npz"), **png_expected)
np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)
print(sorted(os.listdir(OUT)), sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)), "bytes")
