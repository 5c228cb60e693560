"""Algorithmic bytes per ray segment, from the oracle:
B_seg = 160 B of path state + 32 B * n_box + 48 B * n_tri, where n_box / n_tri are the per-segment
averages of an EARLY-OUT traversal of the reference-ordered median-split tree over every segment
(all depths) of a reduced-resolution render of the workload. Builder-independent by construction.
Writes profiles/alg_bytes.json.  Usage: python scripts/alg_bytes.py [config ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from voidray_b200 import scenes  # noqa: E402
from voidray_b200.scene import RenderSettings  # noqa: E402

names = sys.argv[1:] or ["config1_mushroom", "config2_mossy_ground", "config3_materials", "config5_combined"]
path = os.path.join(ROOT, "profiles", "alg_bytes.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for name in names:
    scene, settings, (w, h) = scenes.CONFIGS[name]()
    w4, h4, spp = w // 4, h // 4, 4
    rs = RenderSettings(total_samples=spp, max_bounces=settings.render.max_bounces)
    osc = O.OracleScene(scene)
    _, early = osc.render(w4, h4, rs, spp, mode=O.MODE_EARLY_OUT)
    _, faith = osc.render(w4, h4, rs, spp, mode=O.MODE_FAITHFUL)
    n_box = early.box_tests / early.segments
    n_tri = early.tri_tests / early.segments
    out[name] = {
        "n_box": n_box, "n_tri": n_tri, "bytes_per_segment": 160.0 + 32.0 * n_box + 48.0 * n_tri,
        "segments_per_sample": early.segments / (w4 * h4 * spp),
        "reference_traversal": {"n_box": faith.box_tests / faith.segments, "n_tri": faith.tri_tests / faith.segments},
        "measured_on": f"{w4}x{h4} x {spp} spp, all depths, oracle MODE_EARLY_OUT",
    }
    print(name, json.dumps(out[name]))
json.dump(out, open(path, "w"), indent=1)
