#!/usr/bin/env python
"""bench.py — headline benchmark of the voidray hot path on B200 (contract in the task statement).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU algorithm, host cores)

A "step" is one pass of the hot path over the workload: `iterative_render(target, scene, settings, spp)`
— spp camera samples for every pixel of the frame — on every GPU (rank r owns global sample range
[r*spp, (r+1)*spp), total_samples = N*spp: weak scaling), followed for N > 1 by one NCCL sum-reduce
of the accumulation buffers onto rank 0.

  value  Msamples/s, scene + wavefront state resident in HBM, timed with CUDA events on the stream the
         kernels run on, max over ranks.
  e2e    the same metric through the host-buffer API: vr_scene_commit (flatten + BVH build + H2D of
         geometry / textures / HDRI from host arrays) + clear + accumulate + reduce + D2H of the
         accumulation buffer into pinned host memory, wall clock, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name -> (recipe, description). Default = BASELINE.json configs[1].
    "config2_mossy_ground": "configs[1]: mossy_ground.obj (14 699 tris) with albedo + normal textures, indoor HDRI "
                            "(closed-form substitute), 1920x1080, 256 spp, max depth 8",
    "config1_mushroom": "configs[0]: mushroom.obj (4 448 tris) + studio HDRI (closed-form substitute), 800x600, "
                        "64 spp, max depth 8",
    "config3_materials": "configs[2]: material_testing stand + 4 stand-in meshes (diffuse / metal / dielectric / "
                         "wood-textured), indoor HDRI, 1920x1080, 1024 spp",
    "config4_field": "configs[3]: 48x47 baked mushroom copies = 10 034 688 tris + studio HDRI, 1920x1080, 256 spp",
    "config5_combined": "configs[4]: mossy_ground + mushroom (examples/mushroom.rs), 3840x2160, 4096 spp sample-range "
                        "sharded (512 spp per GPU at 8 GPUs)",
}
DEFAULT_SPP = {"config2_mossy_ground": 256, "config1_mushroom": 64, "config3_materials": 1024, "config4_field": 256,
               "config5_combined": 512}


def load_scene(name: str, spp: int):
    from voidray_b200 import scenes
    scene, settings, dims = scenes.CONFIGS[name]()
    settings.render.total_samples = spp
    return scene, settings, dims


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def alg_bytes(name: str):
    p = os.path.join(ROOT, "profiles", "alg_bytes.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if name in d:
            return d[name]
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpus):
        self.gpus = set(gpus)
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                if int(f[0]) not in self.gpus:
                    continue
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # under load = the upper half of the samples (the sampler also sees the idle gaps between regions)
        load = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


def cpu_oracle_rate(name: str, spp_hint: int, budget_s: float, n_threads: int):
    """The reference algorithm restated on the CPU (oracle, faithful traversal) on a bounded sample of the
    workload: the full frame at k spp, k sized for ~budget_s of work."""
    from oracle import oracle as O
    scene, settings, (w, h) = load_scene(name, spp_hint)
    rs = settings.render
    osc = O.OracleScene(scene)
    t0 = time.perf_counter()
    osc.render(w, h, rs, 1, n_threads=n_threads)
    t1 = time.perf_counter() - t0
    k = int(max(1, min(spp_hint, round(budget_s / max(t1, 1e-3)))))
    t0 = time.perf_counter()
    _, c = osc.render(w, h, rs, k, sample_offset=1, n_threads=n_threads)
    dt = time.perf_counter() - t0
    return {"msamples_per_s": w * h * k / dt / 1e6, "mrays_per_s": c.segments / dt / 1e6, "seconds": dt, "spp": k,
            "width": w, "height": h}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    name = args.workload
    spp = args.spp or DEFAULT_SPP[name]
    cores = os.cpu_count() or 1
    scene, settings, (w, h) = load_scene(name, spp)
    rs = settings.render
    osc = O.OracleScene(scene)
    t0 = time.perf_counter()
    osc.render(w, h, rs, 1, n_threads=cores)
    t1 = time.perf_counter() - t0
    k = int(max(1, min(spp, round(args.ref_step_seconds / max(t1, 1e-3)))))
    for _ in range(args.warmup):
        osc.render(w, h, rs, 1, n_threads=cores)
    seg = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        _, c = osc.render(w, h, rs, k, sample_offset=i * k, n_threads=cores)
        seg += c.segments
    dt = time.perf_counter() - t0
    value = w * h * k * args.steps / dt / 1e6
    sample = f"full {w}x{h} frame at {k} of {spp} spp per step, faithful traversal (boxed median-split tree, both children visited)"
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mrays_per_s": seg / dt / 1e6,
        "config": {"workload": name, "description": WORKLOADS[name], "width": w, "height": h, "spp_per_step": k,
                   "max_bounces": rs.max_bounces,
                   "note": "the Rust reference cannot be built in this image (no cargo/rustc, needs a Vulkan queue); "
                           "this is the oracle port of its algorithm (oracle/voidray_oracle.cpp) on the host cores"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    from voidray_b200.distributed import env_rank, init_process_group
    from voidray_b200.render import Context, RenderTarget
    from voidray_b200.scene import RenderSettings

    rank, world, local = env_rank()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        dist = init_process_group("nccl")
    name = args.workload
    spp = args.spp or DEFAULT_SPP[name]
    scene, settings, (w, h) = load_scene(name, spp)
    stream = torch.cuda.Stream()
    ctx = Context(local, stream.cuda_stream)
    rs = RenderSettings(total_samples=spp * world, max_bounces=settings.render.max_bounces,
                        firefly_clamp=settings.render.firefly_clamp, sample_offset=rank * spp,
                        max_paths_in_flight=args.paths)
    accel = scene.build_acceleration(ctx)
    info = accel.info()
    target = RenderTarget(accel, (w, h), rs)
    acc_t = target.as_torch()
    host_out = torch.empty((h, w, 4), dtype=torch.float32, pin_memory=True).numpy()

    peer_handles = []
    if dist is not None and args.reduce == "peer":
        from voidray_b200.distributed import gather_accum_handles, reduce_accum_peers
        peer_handles = gather_accum_handles(target, 0)

    def reduce_():
        if dist is None:
            return
        if args.reduce == "peer":
            # the root sums the other ranks' accumulation buffers over NVLink peer memory inside its own kernel
            with torch.cuda.stream(stream):
                reduce_accum_peers(target, peer_handles, 0)
        else:
            with torch.cuda.stream(stream):
                dist.reduce(acc_t, dst=0, op=dist.ReduceOp.SUM)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        target.clear()
        target.accumulate(spp)
        reduce_()

    def step_e2e():
        accel.commit()
        target.clear()
        target.accumulate(spp)
        reduce_()
        if rank == 0:
            target.read(host_out)
        else:
            stream.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sampler = ClockSampler(range(world)) if rank == 0 else None
    for _ in range(args.warmup):
        step_device()
    barrier()
    if sampler:
        sampler.start()
    s0 = target.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    stats_acc = {"seg": 0, "trace_ms": 0.0, "trace_launches": 0, "launches": 0}
    for _ in range(args.steps):
        step_device()
        st = target.stats()   # clear() resets the counters each step
        stats_acc["seg"] += st.ray_segments
        stats_acc["trace_ms"] += st.trace_ms
        stats_acc["trace_launches"] += st.trace_launches
        stats_acc["launches"] += st.kernel_launches
    ev1.record(stream)
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))

    # end-to-end through host buffers
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    info = accel.info()

    samples_total = float(w) * h * spp * world * args.steps
    seg_total = sum_over_ranks(float(stats_acc["seg"]))
    launches_total = sum_over_ranks(float(stats_acc["launches"]))
    value = samples_total / (dev_ms * 1e-3) / 1e6
    mrays = seg_total / (dev_ms * 1e-3) / 1e6

    if rank != 0:
        return

    # roofline of the dominant kernel (closest hit), DESIGN.md §5
    peak, peak_src = measured_peak_gbs()
    ab = alg_bytes(name)
    roofline = None
    if ab is not None and stats_acc["trace_ms"] > 0:
        achieved = ab["bytes_per_segment"] * stats_acc["seg"] / (stats_acc["trace_ms"] * 1e-3) / 1e9
        traffic, l2_level = None, None
        tp = os.path.join(ROOT, "profiles", "trace_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get(name)
            l2_level = tj.get(name + "_l2")  # ncu: bytes L2 delivered to the L1s per launch, hit rates
        roofline = {"bound": "hbm", "kernel": "k_trace (closest hit)", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_segment": ab["bytes_per_segment"], "n_box": ab["n_box"], "n_tri": ab["n_tri"],
                    "segments_per_launch": stats_acc["seg"] / max(1, stats_acc["trace_launches"]),
                    "avg_launch_ms": stats_acc["trace_ms"] / max(1, stats_acc["trace_launches"]),
                    "trace_share_of_step": stats_acc["trace_ms"] / dev_ms, "l2_level": l2_level}

    cpu = None
    extra = {}
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        r = cpu_oracle_rate(name, spp, args.cpu_seconds, cores)
        cpu = {"value": r["msamples_per_s"], "unit": "Msamples/s", "cores": cores, "kind": "port",
               "sample": f"full {r['width']}x{r['height']} frame at {r['spp']} of {spp} spp ({r['seconds']:.1f} s), "
                         "oracle port of the reference algorithm, faithful traversal, all host threads",
               "mrays_per_s": r["mrays_per_s"]}
        if not args.no_extra:
            # integrator 1 (HDRI importance sampling + Russian roulette) on the same workload;
            # a different estimator (equal in expectation only without the firefly clamp), so never the headline
            tf = RenderTarget(accel, (w, h), RenderSettings(total_samples=spp, max_bounces=rs.max_bounces,
                                                            firefly_clamp=rs.firefly_clamp, integrator=1))
            tf.accumulate(min(spp, 16))
            tf.clear()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            tf.accumulate(spp)
            f1.record(stream)
            torch.cuda.synchronize()
            fms = f0.elapsed_time(f1)
            extra["integrator_fast"] = {"msamples_per_s": w * h * spp / (fms * 1e-3) / 1e6,
                                        "mrays_per_s": tf.stats().ray_segments / (fms * 1e-3) / 1e6,
                                        "segments_per_sample": tf.stats().ray_segments / (w * h * spp),
                                        "parity_segments_per_sample": stats_acc["seg"] / (w * h * spp * args.steps)}
            tf.close()
        if name != "config1_mushroom" and not args.no_extra:
            # the scene north_star's 100x target is quoted on, measured the same way (device-resident)
            sc1, st1, (w1, h1) = load_scene("config1_mushroom", 64)
            a1 = sc1.build_acceleration(ctx)
            t1 = RenderTarget(a1, (w1, h1), RenderSettings(total_samples=64, max_bounces=8))
            for _ in range(3):
                t1.clear()
                t1.accumulate(64)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(10):
                t1.clear()
                t1.accumulate(64)
            e1.record(stream)
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / 10
            o1 = np.empty((h1, w1, 4), np.float32)
            tt = time.perf_counter()
            for _ in range(5):
                a1.commit()
                t1.clear()
                t1.accumulate(64)
                t1.read(o1)
            e2e1 = (time.perf_counter() - tt) / 5
            c1 = cpu_oracle_rate("config1_mushroom", 64, 8.0, cores)
            extra["config1_mushroom"] = {
                "description": WORKLOADS["config1_mushroom"],
                "msamples_per_s": w1 * h1 * 64 / (ms1 * 1e-3) / 1e6,
                "mrays_per_s": t1.stats().ray_segments / (ms1 * 1e-3) / 1e6,
                "e2e_msamples_per_s": w1 * h1 * 64 / e2e1 / 1e6,
                "cpu_port_msamples_per_s": c1["msamples_per_s"], "cpu_cores": cores,
                "cpu_sample": f"full frame at {c1['spp']} of 64 spp",
            }

    line = {
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (reference OBJ/JPG/TIF assets; closed-form substitutes for the missing EXR HDRIs)",
        "mrays_per_s": mrays,
        "config": {"workload": name, "description": WORKLOADS[name], "width": w, "height": h, "spp_per_gpu": spp,
                   "total_samples": spp * world, "max_bounces": rs.max_bounces, "integrator": "parity (reference estimator)",
                   "parallelism": (f"sample-range x{world} + " + ("peer-memory reduce kernel (CUDA IPC over NVLink)"
                                                                  if args.reduce == "peer" else "ncclReduce"))
                   if world > 1 else "1 GPU",
                   "triangles": info["n_triangles"], "bvh_nodes": info["n_bvh_nodes"],
                   "l2": "inputs larger than L2: each wavefront batch streams ~0.9 GB of path state plus "
                         f"{info['h2d_bytes'] / 1e6:.0f} MB of scene data through the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": samples_total / e2e_s / 1e6, "unit": "Msamples/s",
                "h2d_bytes_per_step": int(info["h2d_bytes"]) * world, "d2h_bytes_per_step": w * h * 16,
                "ms_per_step": e2e_s / args.steps * 1e3,
                "commit_ms": {"flatten_and_bvh": info["flatten_ms"], "upload": info["upload_ms"]},
                "what": "vr_scene_commit (flatten + BVH + H2D) + clear + accumulate + reduce + read_accum to pinned host"},
        "gpu_launches": int(launches_total),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if extra:
        line["other_scenes"] = extra
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2_mossy_ground", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel per GPU per step (default: the config's)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-step-seconds", type=float, default=5.0, help="--impl reference: CPU work per step")
    ap.add_argument("--paths", type=int, default=0, help="wavefront capacity (paths in flight); 0 = library default")
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"],
                    help="N > 1: sum the accumulation buffers with the library's peer-memory kernel (CUDA IPC + NVLink "
                         "loads) or with ncclReduce through torch.distributed")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        try:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier()
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
