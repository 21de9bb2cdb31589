#!/usr/bin/env python
"""Build A/B variants of libsinnerf_b200.so that differ only in field_tc.cu (macros or an older revision of the file).

    python tools/build_variants.py name[:git-rev][:-DMACRO=1,...] ...

Objects of the other sources are compiled once into /tmp/vb; the variant libraries land in variants/ (git-ignored,
shipped to the GPU box by gpurun) and are selected with SNB_LIB_PATH=variants/libsnb_<name>.so.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sinnerf_b200 import build as B  # noqa: E402

OBJ = "/tmp/vb"
os.makedirs(OBJ, exist_ok=True)
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
flags = [f for f in B.NVCC_FLAGS if f not in ("--shared",)]
common = []
for src in B.SOURCES:
    if src == "field_tc.cu":
        continue
    o = os.path.join(OBJ, src + ".o")
    sp = os.path.join(B.CSRC, src)
    if not os.path.exists(o) or os.path.getmtime(o) < os.path.getmtime(sp):
        subprocess.check_call([B._nvcc()] + flags + ["-c", sp, "-o", o])
    common.append(o)
for spec in sys.argv[1:]:
    parts = spec.split(":")
    name, rev, macros = parts[0], (parts[1] if len(parts) > 1 else ""), (parts[2].split(",") if len(parts) > 2 and parts[2] else [])
    src = os.path.join(B.CSRC, "field_tc.cu")
    if rev:
        src = os.path.join(OBJ, f"field_tc_{name}.cu")
        with open(src, "w") as fh:
            fh.write(subprocess.check_output(["git", "-C", ROOT, "show", f"{rev}:sinnerf_b200/csrc/field_tc.cu"], text=True))
    o = os.path.join(OBJ, f"field_tc_{name}.o")
    subprocess.check_call([B._nvcc()] + flags + macros + ["-I", B.CSRC, "-c", src, "-o", o])
    out = os.path.join(ROOT, "variants", f"libsnb_{name}.so")
    subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + common + [o, "-o", out])
    print(out)
