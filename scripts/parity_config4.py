"""One-off parity run on the 10 034 688-triangle scene (config 4): primary-ray ids / distances at 1920x1080 and
400 k incoherent rays, CUDA vs the oracle's faithful traversal. Too slow for pytest (the oracle builds a boxed
10 M-leaf tree); the result is recorded in profiles/r1_parity_config4.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from voidray_b200 import scenes  # noqa: E402
from voidray_b200.render import Context, RenderTarget  # noqa: E402
from util import random_rays, scene_bounds  # noqa: E402

t0 = time.time()
scene, st, (w, h) = scenes.config4_field()
rs = st.render
print("scene %.1fs, %d triangles" % (time.time() - t0, scene.n_triangles()), flush=True)
t0 = time.time()
osc = O.OracleScene(scene)
print("oracle build %.1fs" % (time.time() - t0), flush=True)
ctx = Context(0)
t0 = time.time()
accel = scene.build_acceleration(ctx)
print("commit %.1fs" % (time.time() - t0), accel.info(), flush=True)
tgt = RenderTarget(accel, (w, h), rs)
out = {"triangles": scene.n_triangles(), "resolution": [w, h]}
_, _, s_ref, p_ref, t_ref, _ = osc.trace_primary(w, h, rs, 0)
s, p, t = tgt.trace_primary(0)
hit = s_ref != 0xFFFFFFFF
out["primary"] = {"rays": int(w * h), "hit_fraction": float(hit.mean()), "id_mismatches": int((s != s_ref).sum() + (p != p_ref).sum()),
                  "t_bit_equal": bool(np.array_equal(t, t_ref)),
                  "max_rel_t_err": float(np.max(np.abs(t[hit] - t_ref[hit]) / t_ref[hit]))}
print(out["primary"], flush=True)
lo, hi = scene_bounds(scene)
o, d = random_rays(400000, lo, hi, seed=9)
s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d)
s, p, t = accel.trace_rays(o, d)
out["incoherent"] = {"rays": 400000, "hit_fraction": float((s_ref != 0xFFFFFFFF).mean()),
                     "id_mismatches": int((s != s_ref).sum() + (p != p_ref).sum()), "t_bit_equal": bool(np.array_equal(t, t_ref))}
print(out["incoherent"], flush=True)
rank_equal = bool(np.array_equal(accel.tie_ranks(0), osc.global_tie_rank(0)))
out["tie_ranks_equal"] = rank_equal
print("tie ranks equal", rank_equal, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_config4.json"), "w"), indent=1)
