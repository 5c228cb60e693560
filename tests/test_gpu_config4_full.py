"""-m gpu, slow: BASELINE.json configs[3] at its full size — 48 x 47 baked mushroom copies = 10 034 688 triangles in one
mesh. Tie ranks (the reference tree's in-order leaf sequence, computed by the library's commit) equal the oracle's,
primary-ray ids and distances at 1920x1080 and 400 k incoherent rays equal the oracle's faithful traversal bit for bit.
Takes a minute or two: the oracle builds the reference's boxed 10 M-leaf tree on the host."""
import numpy as np
import pytest

from voidray_b200 import scenes
from voidray_b200.render import RenderTarget

from util import random_rays, scene_bounds

pytestmark = [pytest.mark.gpu, pytest.mark.slow]
MISS = 0xFFFFFFFF


def test_ten_million_triangle_field_matches_the_oracle(oracle, ctx):
    scene, st, (w, h) = scenes.config4_field()
    assert scene.n_triangles() == 10034688 and (w, h) == (1920, 1080)
    osc = oracle.OracleScene(scene)
    accel = scene.build_acceleration(ctx)
    assert np.array_equal(accel.tie_ranks(0), osc.global_tie_rank(0))
    tgt = RenderTarget(accel, (w, h), st.render)
    _, _, s_ref, p_ref, t_ref, _ = osc.trace_primary(w, h, st.render, 0)
    s, p, t = tgt.trace_primary(0)
    assert np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(t, t_ref)
    assert 0.4 < (s_ref != MISS).mean() < 0.6
    tgt.close()
    o, d = random_rays(400000, *scene_bounds(scene), seed=9)
    s_ref, p_ref, t_ref, _ = osc.trace_rays(o, d)
    s, p, t = accel.trace_rays(o, d)
    assert np.array_equal(s, s_ref) and np.array_equal(p, p_ref) and np.array_equal(t, t_ref)
    assert 0.2 < (s_ref != MISS).mean()
    accel.close()
