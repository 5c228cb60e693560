"""Regenerates tests/golden/oracle_golden.npz from the CPU oracle.

The reference ships no golden vectors for this path and cannot be run here, so
these fixtures are ORACLE outputs (parity unpinned): they anchor the oracle against silent drift and
give the GPU tests a second, committed comparison target. Run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from voidray_b200 import scenes  # noqa: E402
from voidray_b200.scene import RenderSettings  # noqa: E402


def cases():
    """name -> (scene, render settings, (W, H))"""
    out = {}
    s, st, _ = scenes.config1_mushroom(64, 48, 4)
    out["config1"] = (s, st.render, (64, 48))
    s, st, _ = scenes.config3_materials(64, 36, 4)
    out["config3"] = (s, st.render, (64, 36))
    s, st, _ = scenes.config5_combined(64, 36, 4)
    out["config5"] = (s, st.render, (64, 36))
    s, st, _ = scenes.example_cornell()
    out["cornell"] = (s, RenderSettings(total_samples=4, max_bounces=6), (40, 40))
    s, st, _ = scenes.example_spheres()
    out["spheres"] = (s, RenderSettings(total_samples=4, max_bounces=8), (48, 48))
    s, st, _ = scenes.example_material()
    out["material"] = (s, RenderSettings(total_samples=4, max_bounces=6), (40, 40))
    return out


def pairs(n_pixels, n=384, seed=7):
    rng = np.random.default_rng(seed)
    return rng.integers(0, n_pixels, n).astype(np.uint32), rng.integers(0, 4, n).astype(np.uint32)


def main():
    data = {}
    for name, (scene, rs, (w, h)) in cases().items():
        osc = O.OracleScene(scene)
        _, _, surface, prim, t, _ = osc.trace_primary(w, h, rs, 1)
        px, sm = pairs(w * h)
        data[f"{name}/surface"] = surface
        data[f"{name}/prim"] = prim
        data[f"{name}/t"] = t
        data[f"{name}/radiance"] = osc.sample_radiance(w, h, rs, px, sm)
        acc, _ = osc.render(w, h, rs, 4, n_threads=1)
        data[f"{name}/accum"] = acc
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz"), **data)
    print("wrote", len(data), "arrays")


if __name__ == "__main__":
    main()
