"""Is integrator 1 (HDRI importance sampling by one-sample MIS + Russian roulette) worth shipping? Time to equal error.
  python scripts/fast_integrator_eval.py [workload] [width] [height]
Reference image: the reference estimator (integrator 0) without the firefly clamp at 16 384 spp (the clamp biases the two
estimators differently, so the comparison is clamp-free; rendered on the GPU, whose image is gated against the oracle
per pixel). Then both integrators, clamp-free, at 16 / 64 / 256 spp: relMSE against the reference and device time.
efficiency = 1 / (relMSE x time); the ratio of the two efficiencies is the speed-up at equal error."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voidray_b200 import scenes  # noqa: E402
from voidray_b200.render import Context, RenderTarget  # noqa: E402
from voidray_b200.scene import RenderSettings  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2_mossy_ground"
w = int(sys.argv[2]) if len(sys.argv) > 2 else 480
h = int(sys.argv[3]) if len(sys.argv) > 3 else 270
scene, settings, _ = scenes.CONFIGS[name](w, h)
ctx = Context(0)
accel = scene.build_acceleration(ctx)
BIG = 3.0e38


def render(spp, integrator, seed, clamp=BIG):
    rs = RenderSettings(total_samples=spp, max_bounces=settings.render.max_bounces, firefly_clamp=clamp, integrator=integrator,
                        seed=seed)
    t = RenderTarget(accel, (w, h), rs)
    t.accumulate(min(spp, 4))  # warm the kernels of this integrator
    t.clear()
    t.accumulate(spp)
    st = t.stats()
    img = t.read()[..., :3].astype(np.float64)
    t.close()
    return img, st.device_ms, st.ray_segments


def rel_mse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


ref, ref_ms, _ = render(16384, 0, 0xC0FFEE)
rows = []
for spp in (16, 64, 256):
    row = {"spp": spp}
    for integ in (0, 1):
        errs, ms, seg = [], 0.0, 0
        for seed in (11, 22, 33):
            img, m, s = render(spp, integ, seed)
            errs.append(rel_mse(img, ref))
            ms += m / 3
            seg += s / 3
        row[f"relmse_{integ}"] = float(np.mean(errs))
        row[f"ms_{integ}"] = ms
        row[f"segments_per_sample_{integ}"] = seg / (w * h * spp)
    row["speedup_at_equal_error"] = (row["relmse_0"] * row["ms_0"]) / (row["relmse_1"] * row["ms_1"])
    rows.append(row)
print(json.dumps({"workload": name, "width": w, "height": h, "reference": "integrator 0, no clamp, 16384 spp", "reference_ms": ref_ms,
                  "rows": rows}, indent=1))
