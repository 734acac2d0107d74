import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_pins
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libb200shim.so"))
cases = {n: (d, r) for n, d, r in ref_pins.image_cases()}
desc, rp = cases["cbox_box_8spp"]
h = ref_pins.reference_scene(lib, desc, rp)
W, H = desc.camera.film_size()
out = np.zeros((H, W, 5), np.float32)
err = C.create_string_buffer(2048)
os.environ["B2_VERBOSE"] = "1"
rc = lib.pathref_render_b200(h, 0, 1, out.ctypes.data_as(C.POINTER(C.c_float)), err, 2048)
print("rc", rc, err.value, "sum", out.sum(0).sum(0), "max", out.max())
ref = np.zeros((H, W, 5), np.float32)
lib.pathref_render(h, ref.ctypes.data_as(C.POINTER(C.c_float)))
print("ref sum", ref.sum(0).sum(0))
