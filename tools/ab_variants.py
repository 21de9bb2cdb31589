#!/usr/bin/env python
"""A/B the field-kernel variants built by tools/build_variants.py: output equality against the first variant and timings.

    python tools/ab_variants.py base defer rot ...        (run on the GPU box; writes to stdout)
"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = sys.argv[1:]
os.makedirs("/tmp/ab", exist_ok=True)


def run(name, *args, timeout=150):
    env = dict(os.environ, SNB_LIB_PATH=os.path.join(ROOT, "variants", f"libsnb_{name}.so"))
    cmd = ["timeout", "-s", "KILL", str(timeout), sys.executable] + list(args)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=ROOT)
    out = (r.stdout + r.stderr).strip().splitlines()
    return r.returncode, out


for prec in ("bf16", "f16x3"):
    for name in names:
        dump = f"/tmp/ab/{name}_{prec}.pt"
        rc, out = run(name, "tools/time_field.py", "--precision", prec, "--rays", "8003", "--iters", "3", "--dump", dump)
        line = out[-1] if out else ""
        cmp_ = ""
        ref = f"/tmp/ab/{names[0]}_{prec}.pt"
        if rc == 0 and os.path.exists(ref) and os.path.exists(dump):
            a, b = torch.load(ref), torch.load(dump)
            d = (a - b).abs()
            cmp_ = f"  | vs {names[0]}: bit-equal={bool(torch.equal(a, b))} max|d|={float(d.max()):.3e} (rgb {float(d[..., :3].max()):.3e}, sigma {float(d[..., 3].max()):.3e}) finite={bool(torch.isfinite(b).all())}"
        print(f"[{name:8s} {prec:5s} rc={rc}] {line}{cmp_}", flush=True)
        if rc != 0:
            print("    " + "\n    ".join(out[-6:]), flush=True)
    for name in names:
        for rays in ("160000", "5292"):
            rc, out = run(name, "tools/time_field.py", "--precision", prec, "--rays", rays, "--iters", "7")
            print(f"[{name:8s} {prec:5s} rc={rc}] {out[-1] if out else ''}", flush=True)
