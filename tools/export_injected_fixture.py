"""Writes the injected-duration fixture of tests/golden/seq_literal_seed123.npz as raw little-endian Float64 files that
tools/patched_reference.jl can read without any Julia package: the RTS-79 unit table, the integer-MW load curve, the
duration lists D[u][k] (k = 0 initial time to failure, then repair, failure, ... -- the order
GeneratingAdequacy/PowerSystemAdequacy.jl:224,243,246 consumes them) and the per-year LOL hours / ENS the CPU oracle and
the CUDA path produce for them.

usage: python tools/export_injected_fixture.py <output directory>"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from powersystemsreliabilityassessment_b200 import rts79

out = sys.argv[1] if len(sys.argv) > 1 else "injected_fixture"
os.makedirs(out, exist_ok=True)
g = np.load(os.path.join(ROOT, "tests", "golden", "seq_literal_seed123.npz"))
cap, mttf, mttr = rts79.units()
load = rts79.load_curve_int().astype(np.float64)
dur = np.ascontiguousarray(g["dur"], dtype=np.float64)          # [U][K]
for name, arr in (("cap", cap), ("mttf", mttf), ("mttr", mttr), ("load", load), ("dur", dur), ("lol", g["lol"]), ("eue", g["eue"])):
    np.ascontiguousarray(arr, dtype="<f8").tofile(os.path.join(out, name + ".f64"))
with open(os.path.join(out, "meta.txt"), "w") as f:
    f.write(f"{dur.shape[0]} {dur.shape[1]} {len(g['lol'])} {len(load)}\n")      # U K years H
print(f"fixture written to {out}: U={dur.shape[0]} K={dur.shape[1]} years={len(g['lol'])} H={len(load)}")
