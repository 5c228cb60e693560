"""First end-to-end check on a GPU box: CUDA path vs the oracle on config 1 (reduced size)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from voidray_b200 import scenes
from voidray_b200.render import Context, RenderTarget

W, H, SPP = 400, 300, 8
scene, settings, _ = scenes.config1_mushroom(W, H, SPP)
rs = settings.render
t = time.time(); osc = O.OracleScene(scene); print("oracle build %.3fs" % (time.time() - t))
ctx = Context(0)
t = time.time(); acc = scene.build_acceleration(ctx); print("gpu commit %.3fs" % (time.time() - t))
tgt = RenderTarget(acc, (W, H), rs)

# RNG
import ctypes as C
from voidray_b200 import _lib
lib = _lib.load()
out = np.empty(64, np.uint32)
_lib.check(lib.vr_debug_rng_draws(ctx.handle, rs.seed, 5, 7, 64, _lib.uptr(out)))
print("rng equal:", np.array_equal(out, O.rng_draws(rs.seed, 5, 7, 64)))
sp = np.empty((64, 3), np.float32)
_lib.check(lib.vr_debug_unit_sphere(ctx.handle, rs.seed, 5, 7, 64, _lib.fptr(sp)))
print("unit_sphere equal:", np.array_equal(sp, O.unit_sphere(rs.seed, 5, 7, 64)))

# tie ranks
print("tie ranks equal:", np.array_equal(acc.tie_ranks(0), osc.global_tie_rank(0)))

# primary gate
for smp in (0, 3):
    o, d, s_ref, p_ref, t_ref, cnt = osc.trace_primary(W, H, rs, smp)
    s, p, tt = tgt.trace_primary(smp)
    hit = s_ref != 0xFFFFFFFF
    print("sample", smp, "hit frac %.3f" % hit.mean(), "id mismatches", int((s != s_ref).sum() + (p != p_ref).sum()),
          "t bit-equal", bool(np.array_equal(tt[hit], t_ref[hit])),
          "max rel t err", float(np.max(np.abs(tt[hit] - t_ref[hit]) / t_ref[hit])) if hit.any() else 0.0,
          "box/ray %.1f tri/ray %.1f" % (cnt.box_tests / (W * H), cnt.tri_tests / (W * H)))

# per-sample radiance
rng = np.random.default_rng(1)
px = rng.integers(0, W * H, 20000).astype(np.uint32)
sm = rng.integers(0, SPP, 20000).astype(np.uint32)
L_ref = osc.sample_radiance(W, H, rs, px, sm)
L = tgt.sample_radiance(px, sm)
diff = np.abs(L - L_ref)
print("radiance: bit-equal frac %.5f" % np.mean(np.all(L == L_ref, axis=1)), "max abs diff", diff.max(),
      "n > 1e-4:", int((diff.max(axis=1) > 1e-4).sum()))

# accumulate
t = time.time(); tgt.accumulate(SPP); dt = time.time() - t
img = tgt.read()
ref, cnt = osc.render(W, H, rs, SPP)
d = np.abs(img - ref)
st = tgt.stats()
print("accumulate %.4fs  %.1f Msamples/s  %.1f Mrays/s (device %.2f ms, trace %.2f ms)" % (
    dt, W * H * SPP / dt / 1e6, st.ray_segments / dt / 1e6, st.device_ms, st.trace_ms))
print("segments gpu", st.ray_segments, "oracle", cnt.segments)
print("image: max abs diff %.3e, mean abs diff %.3e, bit-equal px frac %.5f" % (d.max(), d.mean(), np.mean(np.all(img == ref, axis=2))))
res = tgt.resolve(1.0, 1.0, 1.0, 1)
res_ref = O.resolve(img, 1.0, 1.0, 1.0, 1)
print("resolve max abs diff %.3e" % np.abs(res - res_ref).max())

# bigger timing run
scene2, settings2, (W2, H2) = scenes.config1_mushroom()
acc2 = scene2.build_acceleration(ctx)
tgt2 = RenderTarget(acc2, (W2, H2), settings2.render)
tgt2.accumulate(4)
tgt2.clear()
t = time.time(); tgt2.accumulate(64); dt = time.time() - t
st = tgt2.stats()
print("config1 800x600x64: %.3fs %.1f Msamples/s %.1f Mrays/s device %.1f ms trace %.1f ms launches %d" % (
    dt, W2 * H2 * 64 / dt / 1e6, st.ray_segments / dt / 1e6, st.device_ms, st.trace_ms, st.kernel_launches))
