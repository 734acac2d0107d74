#!/usr/bin/env python3
"""Generates tests/golden/* from the reference tree.  Run only in the build container (needs
/root/reference and oracle/_ref/libsobolref.so built by oracle/Makefile):

    make -C oracle && python tests/gen_golden.py

Outputs (committed, so the GPU box and CI never read /root/reference):
  sfmt_kat.json   the 64-bit known-answer words (192) of src/tests/test_random.cpp:436-507 (Random(4321).nextULong())
  sobol_ref.npz   outputs of the REFERENCE's own sobol::sampleSingle / sobol::look_up
                  (src/samplers/sobolseq.h:45-60,104-133, compiled into oracle/_ref) on seeded inputs
"""
import ctypes as C, json, os, re
import numpy as np

REF = os.environ.get("MTS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "golden")


def sfmt_kat():
    src = open(os.path.join(REF, "src/tests/test_random.cpp")).read()
    m = re.search(r"static const uint64_t reference\[\] = \{(.*?)\};\s*ref<Random> rnd = new Random\((\d+)\)", src, re.S)
    words = [int(h, 16) for h in re.findall(r"0x([0-9a-fA-F]+)ULL", m.group(1))]
    assert len(words) >= 190 and int(m.group(2)) == 4321, (len(words), m.group(2))
    json.dump({"seed": 4321, "source": "src/tests/test_random.cpp:436-507", "words": [f"{w:016x}" for w in words]},
              open(os.path.join(OUT, "sfmt_kat.json"), "w"), indent=0)
    print("sfmt_kat.json:", len(words), "words")


def sobol_ref():
    L = C.CDLL(os.path.join(HERE, "..", "oracle", "_ref", "libsobolref.so"))
    L.sobolref_sample.restype = C.c_float
    L.sobolref_sample.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    L.sobolref_look_up.restype = C.c_uint64
    L.sobolref_look_up.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
    rng = np.random.default_rng(20260922)
    n = 4096
    index = np.concatenate([np.arange(64, dtype=np.uint64), rng.integers(0, 1 << 40, n - 64, dtype=np.uint64)])
    dim = np.concatenate([np.arange(64, dtype=np.uint32) % 8, rng.integers(0, 1024, n - 64).astype(np.uint32)])
    scr = np.where(np.arange(n) % 3 == 0, 0, rng.integers(0, 1 << 32, n, dtype=np.uint64)).astype(np.uint32)
    samples = np.array([L.sobolref_sample(int(i), int(d), int(s)) for i, d, s in zip(index, dim, scr)], np.float32)
    m = rng.integers(2, 12, n).astype(np.uint32)
    frame = rng.integers(0, 4096, n).astype(np.uint32)
    px = (rng.integers(0, 1 << 16, n) % (1 << m)).astype(np.uint32)
    py = (rng.integers(0, 1 << 16, n) % (1 << m)).astype(np.uint32)
    scr64 = np.where(np.arange(n) % 2 == 0, 0, rng.integers(0, 1 << 63, n, dtype=np.uint64)).astype(np.uint64)
    lookup = np.array([L.sobolref_look_up(int(a), int(b), int(c), int(d), int(e)) for a, b, c, d, e in zip(m, frame, px, py, scr64)], np.uint64)
    np.savez_compressed(os.path.join(OUT, "sobol_ref.npz"), index=index, dim=dim, scramble=scr, samples=samples,
                        m=m, frame=frame, px=px, py=py, scramble64=scr64, lookup=lookup)
    print("sobol_ref.npz:", n, "samples +", n, "look_ups")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    sfmt_kat()
    sobol_ref()
