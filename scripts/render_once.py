"""One accumulate of a BASELINE config through the host API, for profiler captures:
  python scripts/render_once.py <workload> <spp> [repeats]
prints one JSON line with the ray segments traced (stats().ray_segments) so that per-segment figures can be formed from
the profiler's per-kernel totals. Every repeat clears the target first, so it traces the same segments again."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voidray_b200 import scenes  # noqa: E402
from voidray_b200.render import Context, RenderTarget  # noqa: E402

name, spp = sys.argv[1], int(sys.argv[2])
repeats = int(sys.argv[3]) if len(sys.argv) > 3 else 1
scene, settings, (w, h) = scenes.CONFIGS[name]()
settings.render.total_samples = spp
ctx = Context(0)
accel = scene.build_acceleration(ctx)
target = RenderTarget(accel, (w, h), settings.render)
seg = 0
for _ in range(repeats):
    target.clear()
    target.accumulate(spp)
    seg = target.stats().ray_segments
st = target.stats()
print(json.dumps({"workload": name, "width": w, "height": h, "spp": spp, "repeats": repeats, "segments_per_repeat": seg,
                  "trace_launches_per_repeat": st.trace_launches, "device_ms": st.device_ms, "trace_ms": st.trace_ms}))
