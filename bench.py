#!/usr/bin/env python
"""bench.py — headline benchmark of the voidray hot path on B200 (contract in the task statement).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU algorithm, host cores)

Workload (`--workload auto`, the default; `config.workload` / `config.series` in the JSON line say which ran):
  * one GPU visible, --gpus 1  -> config1_mushroom: BASELINE.json configs[0], the scene north_star's ">= 100x the
    reference CPU path" target is quoted on (800x600, 64 spp, depth 8). A step = iterative_render of the 64 spp.
  * several GPUs visible (the 1 -> 8 scaling series, every N including N = 1) -> config5_combined: configs[4], the 4K
    scene north_star's ">= 85 % at 8 GPUs" target is quoted on. STRONG scaling: a step = the whole job, 4096 spp of the
    3840x2160 frame, sample-range sharded over the N ranks (4096 / N spp each, total_samples = 4096), then one sum of
    the accumulation buffers onto rank 0. At N = 1 a step is ~27 s: if K steps do not fit --max-seconds (150 s) the run
    times fewer steps (never fewer spp) and says so ("steps", "steps_requested").
  `--workload NAME [--spp S]` picks any config by hand (weak scaling: S spp per GPU).

One loop serves both numbers. Every step runs, through the host API:
  vr_scene_commit (flatten + BVH build + H2D of geometry / textures / HDRI from host arrays) -> clear ->
  [ev0] accumulate(spp) -> reduce over ranks [ev1] -> read_accum into pinned host memory (rank 0)
  value  Msamples/s over the [ev0, ev1] regions: scene and wavefront state resident in HBM, CUDA events on the
         stream the kernels run on, max over ranks.
  e2e    Msamples/s over the whole loop, wall clock between barriers, max over ranks: what a caller of the plugin sees.
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
    "config1_mushroom": "configs[0]: mushroom.obj (4 448 tris) + studio HDRI (closed-form substitute), 800x600, "
                        "64 spp, max depth 8",
    "config2_mossy_ground": "configs[1]: mossy_ground.obj (14 699 tris) with albedo + normal textures, indoor HDRI "
                            "(closed-form substitute), 1920x1080, 256 spp, max depth 8",
    "config3_materials": "configs[2]: material_testing stand + 4 stand-in meshes (diffuse / metal / dielectric / "
                         "wood-textured), indoor HDRI, 1920x1080, 1024 spp",
    "config4_field": "configs[3]: 48x47 baked mushroom copies = 10 034 688 tris + studio HDRI, 1920x1080, 256 spp",
    "config5_combined": "configs[4]: mossy_ground + mushroom (examples/mushroom.rs), 3840x2160, 4096 spp sample-range "
                        "sharded over the GPUs",
}
DEFAULT_SPP = {"config1_mushroom": 64, "config2_mossy_ground": 256, "config3_materials": 1024, "config4_field": 256,
               "config5_combined": 512}
STRONG_TOTAL_SPP = 4096  # configs[4]


def visible_gpus() -> int:
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        n = sum(1 for line in out.splitlines() if line.startswith("GPU "))
        if n:
            return n
    except Exception:
        pass
    try:
        import torch
        return int(torch.cuda.device_count())
    except Exception:
        return 0


def pick_workload(args):
    """-> (name, series, spp_per_gpu or None for the strong series)."""
    if args.workload != "auto":
        return args.workload, "weak", args.spp or DEFAULT_SPP[args.workload]
    if args.gpus == 1 and visible_gpus() <= 1:
        return "config1_mushroom", "bench", args.spp or DEFAULT_SPP["config1_mushroom"]
    return "config5_combined", "strong", None


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


def profile_json(fname: str, key: str):
    p = os.path.join(ROOT, "profiles", fname)
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(key)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed regions run."""
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
                                          "-lms", "50"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

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
        # under load = the samples drawing more than the idle/active mid-point of the power range seen
        thr = (min(power) + max(power)) / 2.0
        load = [c for c, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm), "samples_under_load": len(load)}


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
    """The reference's own algorithm on the host cores (oracle port: the Rust crate cannot be built here), on the same
    workload the GPU arm picks for this N; a step is a bounded sample of it (the full frame at k spp)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    name, series, spp = pick_workload(args)
    total = STRONG_TOTAL_SPP if series == "strong" else spp
    cores = os.cpu_count() or 1
    scene, settings, (w, h) = load_scene(name, total)
    rs = settings.render
    osc = O.OracleScene(scene)
    t0 = time.perf_counter()
    osc.render(w, h, rs, 1, n_threads=cores)
    t1 = time.perf_counter() - t0
    k = int(max(1, min(total, round(args.ref_step_seconds / max(t1, 1e-3)))))
    for _ in range(args.warmup):
        osc.render(w, h, rs, 1, n_threads=cores)
    seg = 0
    t0 = time.perf_counter()
    for i in range(args.steps):
        _, c = osc.render(w, h, rs, k, sample_offset=i * k, n_threads=cores)
        seg += c.segments
    dt = time.perf_counter() - t0
    value = w * h * k * args.steps / dt / 1e6
    sample = (f"full {w}x{h} frame at {k} of {total} spp per step, faithful traversal (boxed median-split tree, both "
              "children visited), all host threads")
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if series == "strong" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mrays_per_s": seg / dt / 1e6,
        "config": {"workload": name, "series": series, "description": WORKLOADS[name], "width": w, "height": h,
                   "spp_per_step": k, "total_samples": total, "max_bounces": rs.max_bounces,
                   "note": "the Rust reference cannot be built in this image (no cargo/rustc, needs a Vulkan queue); "
                           "this is the oracle port of its algorithm (oracle/voidray_oracle.cpp) on the host cores; "
                           "a rate metric, so the bounded sample (k spp of the same frame) measures the same thing"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


class Job:
    """One workload on this rank's device(s): scene committed, target allocated, step functions."""

    def __init__(self, torch, name, total_samples, my_offset, my_spp, device, stream, paths=0, devices=None):
        from voidray_b200.render import Context, RenderTarget
        from voidray_b200.scene import RenderSettings
        self.torch = torch
        self.name = name
        self.scene, self.settings, (self.w, self.h) = load_scene(name, total_samples)
        self.spp = my_spp
        self.stream = stream
        self.ctx = Context.multi(devices) if devices else Context(device, stream.cuda_stream)
        self.rs = RenderSettings(total_samples=total_samples, max_bounces=self.settings.render.max_bounces,
                                 firefly_clamp=self.settings.render.firefly_clamp, sample_offset=my_offset,
                                 max_paths_in_flight=paths)
        self.accel = self.scene.build_acceleration(self.ctx)
        self.target = RenderTarget(self.accel, (self.w, self.h), self.rs)
        self.host_out = torch.empty((self.h, self.w, 4), dtype=torch.float32, pin_memory=True).numpy()

    def close(self):
        self.target.close()
        self.accel.close()
        self.ctx.close()


def timed_loop(torch, job, steps, reduce_fn, barrier, is_root, timer_stream=None):
    """K e2e steps; returns (sum of the [ev0, ev1] device regions in ms, wall seconds, per-step stats)."""
    stream = timer_stream or job.stream
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    acc = {"seg": 0, "trace_ms": 0.0, "trace_launches": 0, "launches": 0, "accumulate_ms": 0.0}
    barrier()
    t0 = time.perf_counter()
    for e0, e1 in evs:
        job.accel.commit()
        job.target.clear()
        e0.record(stream)
        job.target.accumulate(job.spp)
        reduce_fn()
        e1.record(stream)
        if is_root:
            job.target.read(job.host_out)
        else:
            stream.synchronize()
        st = job.target.stats()  # clear() resets the counters each step
        acc["seg"] += st.ray_segments
        acc["trace_ms"] += st.trace_ms
        acc["trace_launches"] += st.trace_launches
        acc["launches"] += st.kernel_launches
        acc["accumulate_ms"] += st.device_ms
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    return dev_ms, wall, acc


def issue_roofline(name, acc, dev_ms, sm_count, sm_mhz):
    """Roofline of the dominant kernel, k_trace (closest hit). The scene's nodes and triangles live in L1 / L2 for
    every config but config 4, so HBM bytes do not bound it (measured DRAM traffic is a few % of peak): what binds is
    instruction issue. achieved = warp-level instructions per second = (warp instructions per ray segment, from the
    committed ncu count of the same kernel on the same workload, profiles/r2_trace_inst.json) x (segments per second
    of k_trace, measured live with CUDA events around every launch); peak = SMs x 4 schedulers x SM clock."""
    if acc["trace_ms"] <= 0:
        return None
    inst = profile_json("r2_trace_inst.json", name)
    seg_per_s = acc["seg"] / (acc["trace_ms"] * 1e-3)
    peak_hbm, peak_src = measured_peak_gbs()
    out = {"bound": "issue", "kernel": "k_trace (closest hit)", "unit": "Gwarp-inst/s",
           "peak": sm_count * 4 * sm_mhz * 1e6 / 1e9,
           "peak_source": f"{sm_count} SMs x 4 schedulers x {sm_mhz:.0f} MHz (1 warp instruction per scheduler per clock)",
           "segments_per_launch": acc["seg"] / max(1, acc["trace_launches"]),
           "avg_launch_ms": acc["trace_ms"] / max(1, acc["trace_launches"]),
           "trace_share_of_step": acc["trace_ms"] / dev_ms, "gsegments_per_s": seg_per_s / 1e9}
    if inst:
        out["achieved"] = inst["warp_inst_per_segment"] * seg_per_s / 1e9
        out["frac"] = out["achieved"] / out["peak"]
        out["warp_inst_per_segment"] = inst["warp_inst_per_segment"]
        out["lane_efficiency"] = inst["thread_inst_per_segment"] / (32.0 * inst["warp_inst_per_segment"])
        out["inst_source"] = inst.get("source")
        dram = inst.get("dram_bytes_per_segment")
        out["traffic"] = dram * out["segments_per_launch"] if dram is not None else None
        if dram is not None:
            gbs = dram * seg_per_s / 1e9
            out["hbm"] = {"achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                          "peak_source": peak_src, "what": "measured dram__bytes_read + write of k_trace per segment "
                                                           "(ncu) x live segments/s"}
    else:
        out["achieved"] = None
        out["frac"] = None
        out["traffic"] = None
        out["note"] = "no ncu instruction count committed for this workload (profiles/r2_trace_inst.json)"
    return out


def kernel_alone_roofline(torch, name, total, spp, device, stream, paths, sm_count, sm_mhz, steps):
    """The same roofline with ONE wavefront (VOIDRAY_STREAMS=1, read by vr_render_begin): no other kernel shares the GPU
    with a k_trace launch, so launch duration = the kernel's own. A side measurement after the timed loop; the headline
    numbers come from the shipped two-wavefront path, whose k_trace launches overlap the other wavefront's kernels."""
    old = os.environ.get("VOIDRAY_STREAMS")
    os.environ["VOIDRAY_STREAMS"] = "1"
    try:
        job = Job(torch, name, total, 0, spp, device, stream, paths=paths)
    finally:
        if old is None:
            os.environ.pop("VOIDRAY_STREAMS", None)
        else:
            os.environ["VOIDRAY_STREAMS"] = old
    try:
        for _ in range(2):
            job.target.clear()
            job.target.accumulate(spp)
        d_ms, _, acc = timed_loop(torch, job, steps, lambda: None, torch.cuda.synchronize, True)
        out = issue_roofline(name, acc, d_ms, sm_count, sm_mhz)
        if out is not None:
            out = {k: out[k] for k in ("achieved", "frac", "avg_launch_ms", "segments_per_launch", "gsegments_per_s",
                                       "trace_share_of_step") if k in out}
            out["msamples_per_s"] = float(job.w) * job.h * spp * steps / (d_ms * 1e-3) / 1e6
            out["spp_per_step"] = spp
            out["steps"] = steps
        return out
    finally:
        job.close()


def run_ours(args):
    import torch
    from voidray_b200.distributed import env_rank, init_process_group, shard_samples

    rank, world, local = env_rank()
    inproc = args.launcher == "inproc" and args.gpus > 1
    if inproc:
        if rank != 0:
            return  # one process drives every device
        world_eff = args.gpus
    else:
        if world != args.gpus and world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
        world_eff = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1 and not inproc:
        dist = init_process_group("nccl")
    name, series, spp = pick_workload(args)
    if series == "strong":
        total = STRONG_TOTAL_SPP
        if inproc:
            offset, my_spp = 0, total
        else:
            offset, my_spp = shard_samples(total, world_eff, rank)
    else:
        total = spp * world_eff
        offset, my_spp = (0, total) if inproc else (rank * spp, spp)
    stream = torch.cuda.Stream()
    job = Job(torch, name, total, offset, my_spp, local, stream, paths=args.paths,
              devices=list(range(args.gpus)) if inproc else None)
    target, w, h = job.target, job.w, job.h
    props = torch.cuda.get_device_properties(local)

    peer_handles = []
    if dist is not None and args.reduce == "peer":
        from voidray_b200.distributed import gather_accum_handles, reduce_accum_peers
        peer_handles = gather_accum_handles(target, 0)
    acc_t = target.as_torch() if dist is not None else None

    def reduce_():
        if dist is None:
            return
        with torch.cuda.stream(stream):
            if args.reduce == "peer":
                # the root sums the other ranks' accumulation buffers over NVLink peer memory inside its own kernel
                reduce_accum_peers(target, peer_handles, 0)
            else:
                dist.reduce(acc_t, dst=0, op=dist.ReduceOp.SUM)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def over_ranks(x: float, op) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    # nvidia-smi needs ~0.2 s to deliver its first sample and config 1's timed loop is shorter than that: the sampler
    # starts with the warm-up; "under load" keeps the samples at load power only (the warm-up runs the same steps)
    sampler = ClockSampler(range(world_eff)).start() if rank == 0 else None
    # ---- warm-up (untimed), with the reduce checked once against an independent sum of the shards ----
    reduce_check = None
    t_first = None
    for i in range(args.warmup):
        t0 = time.perf_counter()
        job.accel.commit()
        target.clear()
        target.accumulate(my_spp)
        if dist is not None and i == 0:
            own = acc_t.clone()
            reduce_()
            torch.cuda.synchronize()
            dist.reduce(own, dst=0, op=dist.ReduceOp.SUM)  # NCCL's sum of the same shards
            if rank == 0:
                diff = float((acc_t - own).abs().max().item())
                scale = float(own.abs().max().item())
                reduce_check = {"max_abs_diff_vs_nccl_sum_of_shards": diff, "max_abs_value": scale,
                                "ok": bool(diff <= 1e-5 * max(scale, 1.0))}
                assert reduce_check["ok"], f"reduced image differs from the sum of the shards: {reduce_check}"
        else:
            reduce_()
        if rank == 0:
            target.read(job.host_out)
        torch.cuda.synchronize()
        if i == args.warmup - 1:
            t_first = time.perf_counter() - t0
    # config 1's steps are ~8 ms and nvidia-smi delivers a sample every 50 ms: keep the same untimed steps going until the
    # sampler has seen ~1.5 s of this load, then enter the timed loop straight away
    extra_warmup = 0
    if dist is None and t_first is not None and t_first < 0.1:
        t_load = time.perf_counter()
        while time.perf_counter() - t_load < 1.5:
            job.accel.commit()
            target.clear()
            target.accumulate(my_spp)
            reduce_()
            if rank == 0:
                target.read(job.host_out)
            extra_warmup += 1
    # a step of the strong series at N = 1 takes ~20 s: time fewer steps rather than fewer spp
    steps = args.steps
    t_step = over_ranks(t_first, dist.ReduceOp.MAX) if dist is not None else t_first
    if t_step * steps > args.max_seconds:
        steps = max(3, int(args.max_seconds / t_step))

    dev_ms, wall, acc = timed_loop(torch, job, steps, reduce_, barrier, rank == 0)
    if sampler is not None and wall < 1.0:
        time.sleep(0.3)  # nvidia-smi's 50 ms loop needs a moment to flush its last samples
    clocks = sampler.stop() if sampler else None
    max_op = dist.ReduceOp.MAX if dist is not None else None
    sum_op = dist.ReduceOp.SUM if dist is not None else None
    dev_ms_max = over_ranks(dev_ms, max_op)
    e2e_s = over_ranks(wall, max_op)
    seg_total = over_ranks(float(acc["seg"]), sum_op)
    launches_total = over_ranks(float(acc["launches"]), sum_op)
    info = job.accel.info()
    samples_total = float(w) * h * total * steps
    value = samples_total / (dev_ms_max * 1e-3) / 1e6
    mrays = seg_total / (dev_ms_max * 1e-3) / 1e6
    if rank != 0:
        return

    sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    roofline = issue_roofline(name, acc, dev_ms, props.multi_processor_count, sm_mhz)
    if roofline is not None:
        roofline["note_overlap"] = ("timed region of the shipped path: consecutive batches run on two wavefronts / two streams, so a "
                                    "k_trace launch shares the GPU with the other wavefront's kernels; launch time = the union of "
                                    "the launches' [start, end] intervals. kernel_alone = the same measurement with one wavefront")
        if world_eff == 1:
            try:
                alone_spp = my_spp if (t_first or 0.0) < 1.0 else min(my_spp, 16)
                roofline["kernel_alone"] = kernel_alone_roofline(torch, name, total, alone_spp, local, stream, args.paths,
                                                                 props.multi_processor_count, sm_mhz, min(steps, 5))
            except Exception as e:  # a side measurement must not take the headline line with it
                roofline["kernel_alone"] = {"error": repr(e)}

    cpu = None
    extra = {}
    if world_eff == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        r = cpu_oracle_rate(name, total, args.cpu_seconds, cores)
        cpu = {"value": r["msamples_per_s"], "unit": "Msamples/s", "cores": cores, "kind": "port",
               "sample": f"full {r['width']}x{r['height']} frame at {r['spp']} of {total} spp ({r['seconds']:.1f} s), "
                         "oracle port of the reference algorithm, faithful traversal, all host threads",
               "mrays_per_s": r["mrays_per_s"]}
    if world_eff == 1 and series == "bench" and not args.no_extra:
        # the other BASELINE configs on the same GPU, same loop (fewer steps), each with its own clock record
        job.close()
        for other, o_spp, o_steps in (("config2_mossy_ground", 256, 3), ("config3_materials", 1024, 2),
                                      ("config4_field", 256, 2), ("config5_combined", 512, 2)):
            try:
                oj = Job(torch, other, o_spp, 0, o_spp, local, stream, paths=args.paths)
                for _ in range(2):
                    oj.target.clear()
                    oj.target.accumulate(min(o_spp, 64))
                osmp = ClockSampler([local]).start()
                d_ms, o_wall, o_acc = timed_loop(torch, oj, o_steps, lambda: None, torch.cuda.synchronize, True)
                o_clocks = osmp.stop()
                n_s = float(oj.w) * oj.h * o_spp * o_steps
                o_info = oj.accel.info()
                extra[other] = {"description": WORKLOADS[other], "width": oj.w, "height": oj.h, "spp": o_spp, "steps": o_steps,
                                "msamples_per_s": n_s / (d_ms * 1e-3) / 1e6, "mrays_per_s": o_acc["seg"] / (d_ms * 1e-3) / 1e6,
                                "e2e_msamples_per_s": n_s / o_wall / 1e6,
                                "commit_ms": {"flatten_and_bvh": o_info["flatten_ms"], "upload": o_info["upload_ms"]},
                                "h2d_bytes_per_step": int(o_info["h2d_bytes"]),
                                "roofline": issue_roofline(other, o_acc, d_ms, props.multi_processor_count, sm_mhz),
                                "clocks": o_clocks}
                oj.close()
                if extra[other]["roofline"] is not None:
                    extra[other]["roofline"]["kernel_alone"] = kernel_alone_roofline(
                        torch, other, o_spp, min(o_spp, 32), local, stream, args.paths, props.multi_processor_count, sm_mhz, 2)
            except Exception as e:  # a side measurement must not take the headline line with it
                extra[other] = {"error": repr(e)}

    line = {
        "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world_eff, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / steps, "higher_is_better": True,
        "scaling": "strong" if series == "strong" else "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (reference OBJ/JPG/TIF assets; closed-form substitutes for the missing EXR HDRIs)",
        "mrays_per_s": mrays,
        "config": {"workload": name, "series": series, "description": WORKLOADS[name], "width": w, "height": h,
                   "spp_per_gpu": my_spp if not inproc else total / world_eff, "total_samples": total,
                   "max_bounces": job.rs.max_bounces, "integrator": "parity (reference estimator)",
                   "parallelism": ("1 GPU" if world_eff == 1 else
                                   f"sample-range x{world_eff}, " +
                                   ("one process, vr_context_create_multi (peer-memory reduce kernel)" if inproc else
                                    "one process per GPU + " + ("peer-memory reduce kernel (CUDA IPC over NVLink)"
                                                                if args.reduce == "peer" else "ncclReduce"))),
                   "triangles": info["n_triangles"], "bvh_nodes": info["n_bvh_nodes"],
                   "l2": "inputs larger than L2: each wavefront batch streams up to 0.9 GB of path state plus "
                         f"{info['h2d_bytes'] / max(1, world_eff if inproc else 1) / 1e6:.0f} MB of freshly committed "
                         "scene data through the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": samples_total / e2e_s / 1e6, "unit": "Msamples/s",
                "h2d_bytes_per_step": int(info["h2d_bytes"]) * (1 if inproc else world_eff), "d2h_bytes_per_step": w * h * 16,
                "ms_per_step": e2e_s / steps * 1e3,
                "commit_ms": {"flatten_and_bvh": info["flatten_ms"], "upload": info["upload_ms"]},
                "what": "vr_scene_commit (flatten + BVH + H2D) + clear + accumulate + reduce + read_accum to pinned host"},
        "gpu_launches": int(launches_total),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    if extra_warmup:
        line["warmup_extra_steps_for_clock_sampling"] = extra_warmup
    if steps != args.steps:
        line["steps_requested"] = args.steps
        line["steps_note"] = (f"a step is the whole {total}-spp job (~{t_step:.1f} s here): {steps} steps fit --max-seconds "
                              f"{args.max_seconds:.0f}; spp were not reduced")
    if reduce_check is not None:
        line["reduce_check"] = reduce_check
    if extra:
        line["other_scenes"] = extra
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="with --workload NAME: samples per pixel per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-step-seconds", type=float, default=5.0, help="--impl reference: CPU work per step")
    ap.add_argument("--max-seconds", type=float, default=150.0,
                    help="budget of the timed loop; more than this and fewer steps are timed (strong series at N = 1)")
    ap.add_argument("--paths", type=int, default=0, help="wavefront capacity (paths in flight); 0 = library default")
    ap.add_argument("--launcher", default="ranks", choices=["ranks", "inproc"],
                    help="N > 1: one process per GPU (torchrun) or every GPU from this process (vr_context_create_multi)")
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"],
                    help="N > 1, one process per GPU: sum the accumulation buffers with the library's peer-memory kernel "
                         "(CUDA IPC + NVLink loads) or with ncclReduce through torch.distributed")
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
