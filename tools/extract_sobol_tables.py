#!/usr/bin/env python3
"""Extract the Sobol' direction-number tables the `sobol` sampler is defined by.

The Mitsuba 0.6 `sobol` sampler (src/samplers/sobol.cpp:204-252) is a pure function of
(pixel, sample index, dimension) *given* three third-party data tables shipped in
src/samplers/sobolseq.cpp (Leonhard Gruenschloss 2012, MIT licence; direction numbers from
S. Joe & F. Y. Kuo 2008, "new-joe-kuo-6.21201"):

  matrices32             [1024][52] u32   (sobolseq.cpp:33-53283)
  vdc_sobol_matrices     [26][52]   u64   (sobolseq.cpp:106537-107239)
  vdc_sobol_matrices_inv [26][52]   u64   (sobolseq.cpp:107241-107997)

They are *data*, not code; no other source of these numbers exists offline.  This script parses
the hex literals (no compilation) and writes them as raw little-endian binaries plus a sha256
manifest.  Both the product (mitsuba_b200/data/) and the test fixtures read the binaries; the
generator algorithm (XOR of matrix columns selected by index bits) is written from scratch on
both sides.

Run only where /root/reference exists (the build container):
    python tools/extract_sobol_tables.py
"""
import hashlib, json, os, re, sys
import numpy as np

REF = os.environ.get("MTS_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src/samplers/sobolseq.cpp")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mitsuba_b200", "data")


def grab(text, name, dtype):
    m = re.search(r"Matrices::" + name + r"\b[^=]*=\s*\{", text)
    assert m, name
    start = m.end()
    depth, i = 1, start
    while depth:
        c = text[i]
        depth += (c == "{") - (c == "}")
        i += 1
    body = text[start:i - 1]
    if "{" not in body:
        vals = [int(h, 16) for h in re.findall(r"0x([0-9a-fA-F]+)", body)]
        return np.array(vals, dtype=dtype)
    # 2-D table with partially initialised rows: C zero-fills the remainder of each row
    rows = []
    for row in re.findall(r"\{([^{}]*)\}", body):
        vals = [int(h, 16) for h in re.findall(r"0x([0-9a-fA-F]+)", row)]
        assert len(vals) <= 52
        rows.append(vals + [0] * (52 - len(vals)))
    return np.array(rows, dtype=dtype).reshape(-1)


def main():
    text = open(SRC).read()
    m32 = grab(text, "matrices32", np.uint32)
    vdc = grab(text, "vdc_sobol_matrices", np.uint64)
    inv = grab(text, "vdc_sobol_matrices_inv", np.uint64)
    assert m32.size == 1024 * 52, m32.size
    assert vdc.size == 25 * 52 and inv.size == 26 * 52, (vdc.size, inv.size)  # rows m = 1..25 / 1..26
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name, arr in (("sobol_matrices32.bin", m32), ("sobol_vdc.bin", vdc), ("sobol_vdc_inv.bin", inv)):
        p = os.path.join(OUT, name)
        arr.astype(arr.dtype.newbyteorder("<")).tofile(p)
        manifest[name] = {"sha256": hashlib.sha256(open(p, "rb").read()).hexdigest(),
                          "dtype": str(arr.dtype), "count": int(arr.size)}
    manifest["source"] = "mitsuba 0.6 src/samplers/sobolseq.cpp (Gruenschloss 2012, MIT; Joe-Kuo 2008)"
    json.dump(manifest, open(os.path.join(OUT, "sobol_tables.json"), "w"), indent=1)
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    sys.exit(main())
