#!/usr/bin/env python
"""Headline benchmark: Mray/s of the wavefront path-tracing hot path.

  python bench.py --gpus N --steps K --warmup W            (our arm; N>1 under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  (the CPU arm: oracle port on the host cores)

Headline workload = BASELINE.json config C3, the largest single-GPU configuration (kitchen-scale stand-in,
~5 M triangles, 24 OpenPBR materials, 2 rect lights + 1 distant light, NEE on, material sort on, 1920x1080,
spp 1024, max-bounces 8).  The same JSON line carries C2 (1 M random triangles, one diffuse BSDF, spp 256) and C5
(64-material stress, 2 M triangles, spp 512) as `sub_configs`, each measured the same way on fewer steps.

One "step" = one giRender call = the whole frame at the full sample count.  At N>1 the framebuffer is
tile-sharded (strong scaling: the frame is fixed) and the HDR accumulator is sum-reduced to rank 0 by the
library's own ncclReduce inside gtlc_render, before the device->host copy (gtlc_comm_*; torch.distributed only
carries the NCCL id and the timing scalars).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mray/s"
BYTES_RAY_CLOSEST = 32 + 20          # ray in, hit record out
BYTES_RAY_SHADOW = 32 + 4 + 16       # ray in, occlusion out, NEE contribution added to its target
BYTES_NODE, BYTES_TRI = 80, 48       # one BVH8 node visit, one triangle test (canonical sizes; the traversal layout's padding is not "useful")
BYTES_HIT = {0: 124 + 168 + 32 + 16, 1: 124 + 168 + 32 + 64, 2: 124 + 168 + 32 + 192}   # shaded hit by closure: R4 gathers + state in/out + ray out + P_mat
BYTES_MISS = 84 + 16                 # path state in, finished sample out
BYTES_SHADOW_OUT = 52                # shadow-queue entry written by the shade kernel (origin, direction, contribution, seed)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "c5"])
    ap.add_argument("--sub-configs", default=None, help="comma list measured after the headline config (default: the other two of c2,c3,c5; '' = none)")
    ap.add_argument("--sub-steps", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=5, help="steps of the end-to-end leg (host buffers in the timed region); 0 = as many as --steps")
    ap.add_argument("--spp", type=int, default=0, help="override the config's sample count (profiling runs only: the number is then not the headline)")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--tile", type=int, default=64)
    ap.add_argument("--shard", default="tiles", choices=["tiles", "samples"], help="N>1: framebuffer tiles (default; bit-identical image) or sample ranges (SURVEY.md 8e alternative)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the bounded oracle sample")
    ap.add_argument("--cpu-sample-only", action="store_true", help="(internal) reference arm: print the cpu_baseline object of one bounded sample and exit")
    return ap.parse_args()


def make_scene(name, a):
    from gatling_b200 import scenes
    if name == "c2":
        d = scenes.random_triangles(1_000_000, a.width, a.height, a.spp or 256, 8)
        label = "C2: 1000000 random triangles (seed 0x2545F491), single diffuse BSDF, %dx%d, spp %d, max-bounces 8, NEE off" % (a.width, a.height, d.settings.spp)
    elif name == "c3":
        d = scenes.kitchen_standin(width=a.width, height=a.height, spp=a.spp or 1024, max_bounces=8)
        label = ("C3: kitchen-scale stand-in (seed 0x9E3779B9; the Evermotion asset is not redistributable), %d triangles in 401 instances, 24 OpenPBR materials, "
                 "2 rect lights + 1 distant light, NEE on, material sort on, %dx%d, spp %d, max-bounces 8" % (d.triangle_count, a.width, a.height, d.settings.spp))
    else:
        d = scenes.material_stress(width=a.width, height=a.height, spp=a.spp or 512, max_bounces=8)
        label = ("C5: 64-material stress stand-in (GPUOpen matlib is not reachable), %d triangles, 64 OpenPBR parameter sets, white dome + 1 distant light, NEE on, "
                 "material sort on, %dx%d, spp %d, max-bounces 8" % (d.triangle_count, a.width, a.height, d.settings.spp))
    return d, label


def workload_config(label, d, a, n_gpus):
    return {
        "workload": label, "triangles": d.triangle_count, "materials": len(d.materials), "width": d.width, "height": d.height,
        "spp": d.settings.spp, "max_bounces": d.settings.maxBounces, "nee": bool(d.settings.nextEventEstimation),
        "parallelism": "single GPU" if n_gpus <= 1 else
                       ("sample-range sharding (every pixel, samples [spp*rank/n_gpus, spp*(rank+1)/n_gpus)) x%d + one ncclReduce(sum) of the HDR accumulator to rank 0 inside gtlc_render" % n_gpus)
                       if a.shard == "samples" else
                       ("tile-sharded framebuffer (%dx%d tiles, tile %% n_gpus == rank) x%d + one ncclReduce(sum) of the HDR accumulator to rank 0 inside gtlc_render" % (a.tile, a.tile, n_gpus)),
        "l2_policy": "inputs larger than L2: every wave streams ~2.3 GB of path-state / ray / hit queues (L2 = 126 MB); scene data stays resident in HBM as in production",
    }


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                v = float(json.load(f)["hbm_gbs"])
            if v > 0.0:
                return v, "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception as e:                      # an unreadable file must not cost the bench line
            print("bench: MEASURED_PEAKS.json ignored (%s)" % e, file=sys.stderr)
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def profiled_traffic(config):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel from the committed `ncu --set full` capture of this
    config (profiles/r02_traffic.json, written by tools/summarize_profiles.py)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f).get(config)
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------------
# the CPU arm (the reference's compiled integrator, or the oracle port without it): bench.py's only use of oracle/.  Runs in a process of its own so that the OpenMP
# runtime reads its settings at start-up: all host threads, bound to their places (an unbound team was seen parked
# on one core for a whole run: 6.5x spread between boxes in round 1).
# ---------------------------------------------------------------------------------------------------------
def reference_env():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
    env = dict(os.environ)
    env.update({"OMP_NUM_THREADS": str(n), "OMP_PROC_BIND": "spread", "OMP_PLACES": "threads", "OMP_DYNAMIC": "false", "GTLC_REF_ENV": "1"})
    return env


def oracle_sample(a, target_seconds, steps, warmup):
    """Renders a bounded sample of the config with the oracle: the full scene and resolution, a band of image rows around
    the middle at reduced spp, sized by a calibration band to ~target_seconds per step.  Returns (per-step Mray/s list, cores, text)."""
    from oracle import orc
    desc, _ = make_scene(a.config, a)
    o = orc.OracleScene(desc)
    # The reference's own integrator when it is there: oracle/_ref/libglslref.so is rp_main.rgen / .chit / .miss compiled as C++ from the
    # reference's files in the build container (it travels with the repository snapshot); the oracle then only answers its ray queries and
    # mdl_* calls.  Otherwise the oracle's restatement of the same code (bit-identical images, tests/test_reference_glsl_cpu.py).
    ref = None
    if orc.ReferenceIntegrator.available() and not os.environ.get("GTLC_REF_ARM_PORT"):
        try:
            ref = orc.ReferenceIntegrator()
            ref.render(o, rows=(0, 1), spp=1)                   # refuses render settings no pipeline was compiled for
        except Exception as e:                                  # the arm must produce its line: fall back to the port
            print("bench: reference integrator unavailable (%s); timing the oracle port" % e, file=sys.stderr)
            ref = None

    def render_band(rows, spp):
        if ref is not None:
            _, rays, secs = ref.render(o, rows=rows, spp=spp)
            return secs, rays
        _, secs, rays = o.render(rows=rows, spp=spp)
        return secs, rays
    H = desc.height
    mid = H // 2
    cal_rows = min(H, 32)
    render_band((mid, mid + 2), 1)                              # thread start-up
    secs, rays = render_band((mid - cal_rows // 2, mid - cal_rows // 2 + cal_rows), 1)
    rate = (rays[0] + rays[1]) / max(secs, 1e-6)
    spp = 2
    rows = int(min(H, max(2, rate * target_seconds / max((rays[0] + rays[1]) / float(cal_rows) * spp, 1))))
    r0 = max(0, mid - rows // 2)
    for _ in range(warmup):
        render_band((r0, r0 + rows), spp)
    per_step, total_rays, total_secs = [], 0, 0.0
    for _ in range(steps):
        secs, rays = render_band((r0, r0 + rows), spp)
        per_step.append((rays[0] + rays[1]) / secs / 1e6)
        total_rays += rays[0] + rays[1]; total_secs += secs
    cores = orc.lib().orc_num_threads()
    port = None
    if ref is not None and steps > 1:                            # the same band once through the oracle's restatement, for scale (it skips the callback layer)
        _, psecs, prays = o.render(rows=(r0, r0 + rows), spp=spp)
        port = {"value": (prays[0] + prays[1]) / max(psecs, 1e-9) / 1e6, "unit": METRIC, "steps": 1, "note": "the oracle's restatement of the same integrator on the same band (identical rays)"}
    o.close()
    what = ("the reference's own integrator -- rp_main.rgen / .chit / .miss / _shadow.miss compiled as C++ from the reference's files (oracle/_ref/libglslref.so), "
            "OpenMP over image rows; ray queries (software BVH) and the materials' mdl_* entry points are served by the CPU oracle, the reference's Vulkan-RT "
            "pipeline and MDL SDK being neither buildable nor runnable here" if ref is not None else
            "oracle-CPU port (C++/OpenMP restatement whose integrator renders bit-identically to the reference's own shader code compiled as C++, "
            "tests/test_reference_glsl_cpu.py; oracle/_ref/libglslref.so was not found)")
    text = ("%s; %d threads bound to their places: the same scene and resolution, image rows [%d,%d) of %d at spp %d per step, %d rays in %.1f s over %d steps"
            % (what, cores, r0, r0 + rows, H, spp, total_rays, total_secs, steps))
    return per_step, total_rays / max(total_secs, 1e-9) / 1e6, total_secs, cores, text, ("reference" if ref is not None else "port"), port


def spread_of(vals):
    return {"min": min(vals), "median": statistics.median(vals), "max": max(vals), "steps": len(vals)}


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if os.environ.get("GTLC_REF_ENV") != "1":                  # torchrun pins OMP_NUM_THREADS=1; restart with the full, bound team
        os.execve(sys.executable, [sys.executable] + sys.argv, reference_env())
    if a.cpu_sample_only:
        per_step, value, _, cores, text, kind, port = oracle_sample(a, a.cpu_seconds / 3.0, 3, 1)
        print(json.dumps({"value": value, "unit": METRIC, "cores": cores, "kind": kind, "sample": text, "spread": spread_of(per_step), **({"port": port} if port else {})}))
        return
    desc, label = make_scene(a.config, a)
    per_step_target = max(2.0, min(20.0, 120.0 / max(a.steps + a.warmup, 1)))
    per_step, value, secs, cores, text, kind, port = oracle_sample(a, per_step_target, a.steps, a.warmup)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": 1000.0 * secs / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": workload_config(label, desc, a, a.gpus),
           "cpu_baseline": {"value": value, "unit": METRIC, "cores": cores, "kind": kind, "sample": text, "spread": spread_of(per_step), **({"port": port} if port else {})},
           "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def cpu_baseline_subprocess(a):
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--cpu-sample-only", "--config", a.config, "--cpu-seconds", str(a.cpu_seconds),
           "--width", str(a.width), "--height", str(a.height)] + (["--spp", str(a.spp)] if a.spp else [])
    env = reference_env()
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        for line in reversed(res.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"value": None, "unit": METRIC, "cores": 0, "kind": "port", "sample": "oracle sample failed: " + (res.stderr or res.stdout)[-300:]}
    except Exception as e:                                       # the baseline is context, never a reason to lose the GPU line
        return {"value": None, "unit": METRIC, "cores": 0, "kind": "port", "sample": "oracle sample failed: %r" % (e,)}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def measure(name, a, gi, torch, dist, rank, world, local_rank, steps, warmup, sampler=None):
    desc, label = make_scene(name, a)
    t0 = time.perf_counter()
    built = gi.BuiltScene(desc)
    built.color = gi.giCreateRenderBuffer(desc.width, desc.height)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step(readback):
        gi.set_tuning(skipReadback=0 if readback else 1)
        gi.set_shard(built.scene, rank, world, gi.SHARD_SAMPLES if a.shard == "samples" else a.tile)     # resets the accumulation: every step renders samples [0, spp)
        built.render()                                      # at N>1 ends with the library's ncclReduce to rank 0 (and rank 0's D2H copy when readback is on)
        return gi.get_stats()

    # first call builds + uploads the BVH8 (reported separately, outside the timed region like the reference's dirty-BVH path)
    # and counts node / triangle visits of the identical rays with the per-lane kernel (the oracle's walk order)
    gi.set_tuning(countVisits=1)
    s_count = step(False)
    setup_s = time.perf_counter() - t0
    gi.set_tuning(countVisits=0)
    for _ in range(max(warmup, 3)):
        step(False)

    if sampler is not None:
        sampler.start()
    keys = ("closestRays", "shadowRays", "kernelLaunches", "msTraceClosest", "msTotal", "msShade", "msRaygen", "msResolve", "msTraceAny", "msSort")
    # ---- timed: device path (inputs resident, no readback) ----
    barrier()
    t1 = time.perf_counter()
    acc = {k: 0.0 for k in keys}
    for _ in range(steps):
        s = step(False)
        for k in keys:
            acc[k] += s[k]
    barrier()
    wall = time.perf_counter() - t1
    # ---- timed: end to end through the C-ABI with host buffers (render params in, HDR image out every step) ----
    barrier()
    t2 = time.perf_counter()
    e2e_rays = 0
    e2e_steps = min(steps, a.e2e_steps) if a.e2e_steps > 0 else steps
    for _ in range(e2e_steps):
        s = step(True)
        e2e_rays += s["closestRays"] + s["shadowRays"]
    barrier()
    wall_e2e = time.perf_counter() - t2
    clocks = sampler.stop() if sampler is not None else None

    rays_local = acc["closestRays"] + acc["shadowRays"]
    vals = torch.tensor([float(rays_local), float(e2e_rays), wall, wall_e2e, acc["msTotal"]], dtype=torch.float64, device="cuda")
    if dist is not None:
        sums = vals.clone(); dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        maxs = vals.clone(); dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
        total_rays, total_e2e = sums[0].item(), sums[1].item()
        wall, wall_e2e, dev_ms = maxs[2].item(), maxs[3].item(), maxs[4].item()
    else:
        total_rays, total_e2e, dev_ms = float(rays_local), float(e2e_rays), acc["msTotal"]
    built.close()
    if rank != 0:
        return None

    peak, peak_src = measured_peak()
    traffic = profiled_traffic(name) or {}
    max_closure = max(m.closure for m in desc.materials)
    rays_c, rays_s = s_count["closestRays"], s_count["shadowRays"]
    hits = s_count["closestHits"]

    def roof(kernel, alg_bytes, ms_step, launches_step, limiter, extra):
        achieved = alg_bytes / max(ms_step * 1e-3, 1e-12) / 1e9
        t = traffic.get(kernel.split("<")[0] + ("_shadow" if "SHADOW" in kernel else "")) or {}
        r = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
             "traffic": t.get("dram_bytes_per_launch"), "traffic_source": t.get("source"), "dram_pct_of_peak_ncu": t.get("dram_pct_of_peak"),
             "limiter": limiter, "peak_source": peak_src, "algorithmic_bytes_per_step": alg_bytes, "kernel_ms_per_step": ms_step,
             "launches_per_step": launches_step, "algorithmic_bytes_per_launch": alg_bytes / max(launches_step, 1), "avg_launch_ms": ms_step / max(launches_step, 1)}
        r.update(extra)
        return r

    # launch counts as the library's wave loop derives them (gtlc_render): waves x bounces launches of each per-bounce kernel
    px = max(1, desc.width * desc.height // max(world, 1)); max_paths = 1 << 24
    ppw = min(px, max_paths); spw = 1 if ppw < px else max(1, min(desc.settings.spp, max_paths // ppw))
    waves = -(-px // ppw) * -(-desc.settings.spp // spw)
    bounce_launches = waves * desc.settings.maxBounces
    closest_bytes = BYTES_RAY_CLOSEST * rays_c + BYTES_NODE * s_count["nodeVisits"] + BYTES_TRI * s_count["triVisits"]
    shadow_bytes = BYTES_RAY_SHADOW * rays_s + BYTES_NODE * s_count["shadowNodeVisits"] + BYTES_TRI * s_count["shadowTriVisits"]
    shade_bytes = BYTES_HIT[max_closure] * hits + BYTES_MISS * (rays_c - hits) + BYTES_SHADOW_OUT * rays_s
    note = ("algorithmic bytes (SURVEY.md 8d) from the exact device counters of an untimed counting pass over the identical rays (rank 0's shard): "
            "32+20 per closest ray, 32+4+16 per shadow ray, 80 per BVH8 node visit, 48 per triangle test; per shaded hit 124 (R4 gathers) + 168 (path state in+out) + 32 (ray out) + "
            "P_mat (16 / 64 / 192 B), 100 per miss, 52 per shadow-queue entry.  `bound` names the yardstick SURVEY 8d prescribes; `limiter` is what ncu says actually binds the kernel")
    roofline = roof("trace_kernel_batched<CLOSEST>", closest_bytes, acc["msTraceClosest"] / steps, bounce_launches,
                    "instruction issue (67-72 % of the issue slots, ALU pipe 55-64 %) + L1 data pipe (60-71 %): software BVH8 traversal, ~215 SASS instructions per node step at 21-25 of 32 lanes; DRAM far below peak -- see dram_pct_of_peak_ncu (profiles/r02_trace_c3.md, r02_trace_c2.md)",
                    {"nodes_per_ray": s_count["nodeVisits"] / max(rays_c, 1), "tris_per_ray": s_count["triVisits"] / max(rays_c, 1), "rays_per_step": rays_c, "note": note})
    others = []
    if rays_s:
        others.append(roof("trace_kernel_batched<SHADOW>", shadow_bytes, acc["msTraceAny"] / steps, bounce_launches,
                           "issue slots + L1 data pipe, shorter walks (any-hit): per-ray fixed costs weigh more",
                           {"nodes_per_ray": s_count["shadowNodeVisits"] / max(rays_s, 1), "tris_per_ray": s_count["shadowTriVisits"] / max(rays_s, 1), "rays_per_step": rays_s}))
    others.append(roof("shade_kernel", shade_bytes, acc["msShade"] / steps, bounce_launches,
                       "instruction issue (56-65 % of the issue slots at 47 % occupancy: IEEE-exact closure arithmetic, 64 registers with spills) on top of incoherent shading gathers; DRAM at ~38 % of peak over a whole wave (profiles/r02_shade_c3.md, r02_traffic.json)",
                       {"hits_per_step": hits, "misses_per_step": rays_c - hits, "bytes_per_hit": BYTES_HIT[max_closure]}))
    value = total_rays / wall / 1e6
    rec = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": 1000.0 * wall / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(label, desc, a, world),
        "spp_per_s": desc.settings.spp * steps / wall,
        "device_ms_per_step": dev_ms / steps,
        "rays_per_step": total_rays / steps,
        "e2e": {"value": total_e2e / wall_e2e / 1e6, "unit": METRIC, "h2d_bytes_per_step": 240 + 24,
                "d2h_bytes_per_step": desc.width * desc.height * 16, "ms_per_step": 1000.0 * wall_e2e / e2e_steps, "steps": e2e_steps,
                "note": "gtlc_render through the C-ABI: GtlcRenderParams (camera + settings + AOV binding) from host memory in, the full HDR colour AOV copied to pinned host memory "
                        "every step" + (" -- on rank 0, after the in-library ncclReduce of all shards" if world > 1 else "") +
                        "; scene/BVH stay resident as in the reference (rebuilt only when dirty)"},
        "gpu_launches": int(acc["kernelLaunches"]),
        "roofline": roofline, "roofline_kernels": others,
        "kernel_ms_per_step": {"raygen": acc["msRaygen"] / steps, "trace_closest": acc["msTraceClosest"] / steps, "sort": acc["msSort"] / steps, "shade": acc["msShade"] / steps,
                               "trace_shadow": acc["msTraceAny"] / steps, "resolve": acc["msResolve"] / steps},
        "setup_s": setup_s, "bvh_build_ms": s_count["msBvhBuild"], "bvh_nodes": s_count["bvhNodeCount"],
        "clocks": clocks,
    }
    return rec


class StdoutForJsonOnly:
    """Libraries print to fd 1 behind Python's back (NCCL writes its version line there at the first communicator).  The
    bench's stdout carries ONE JSON line: everything else goes to stderr, the line is written to the saved descriptor."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())


def run_ours(a):
    out_fd = StdoutForJsonOnly()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # torchrun pins OMP_NUM_THREADS=1; the host-side BVH8 build (outside the timed region) is OpenMP-parallel
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // world))
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from gatling_b200 import gi
    st = gi.giInitialize(local_rank)
    if st != gi.GiStatus.Ok:
        raise SystemExit("giInitialize failed: " + gi.giLastError())
    torch.cuda.set_device(local_rank)
    if world > 1:                                              # the library's own communicator: torch only carries the 128-byte id
        box = [gi.comm_get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        gi.comm_init(box[0], rank, world)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    out = measure(a.config, a, gi, torch, dist, rank, world, local_rank, a.steps, a.warmup, sampler)
    subs = []
    names = [x for x in (a.sub_configs.split(",") if a.sub_configs is not None else [c for c in ("c2", "c5", "c3") if c != a.config][:2]) if x]
    if a.spp:
        names = []                                             # profiling run
    for name in names:
        r = measure(name, a, gi, torch, dist, rank, world, local_rank, min(a.steps, a.sub_steps), 3)
        if r is not None:
            for k in ("metric", "unit", "higher_is_better", "vs_baseline", "dtype", "data", "scaling", "clocks"):
                r.pop(k, None)
            subs.append(r)
    if rank == 0:
        out["sub_configs"] = subs
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_subprocess(a)
        out_fd.emit(json.dumps(out))
    if world > 1:
        gi.comm_destroy()
    gi.giTerminate()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
