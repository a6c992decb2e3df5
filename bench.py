#!/usr/bin/env python
"""Benchmark of the per-ray hot path (BASELINE.json: ray-steps/s and full-map wall time incl. GRFF).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                       # the reference's CPU algorithm (oracle port)

A "step" is one full ray-traced GR+FF map: BASELINE config 4 (512^2 pixels, 256^3 cube with
n_e/T/B-vector, 8 log-spaced frequencies 75 MHz-1.5 GHz, rays re-traced per frequency with the
publication drivers' per-frequency dt/n_steps/stride presets, cross-sections traced) through the
fused kernel.  For N > 1 (one process per GPU under torchrun, NCCL) the image grows to
512 x (512 N) pixels over the same field of view, rows are dealt round-robin to the ranks (weak
scaling: per-GPU work fixed), the cube is replicated and the image slabs are all-gathered at the
end of every step; `--config c5` runs BASELINE config 5 (2048^2, 512^3, 16 freqs 20-300 MHz) with
its rows sharded instead (strong scaling).

value  = nominal ray-steps (n_rays x sum_f n_steps_f, one ray-step = one RK4 advance of a central
         ray; the two cross-section rays ride along) / device time, inputs resident in HBM.
e2e    = the same through the public Python API with host (pinned) cubes: H2D of the cubes and ray
         starts and D2H of the T_b / V/I maps inside the timed region, every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from raytracinggrff_b200 import synthetic  # noqa: E402

R_SUN_CM = 6.957e10
BYTES_PER_RAY_STEP_CS = 1536      # 3 RK4 x 4 RHS x 8 corners x 16 B
BYTES_PER_RAY_STEP_NOCS = 512


def workload(config: str, n_gpus: int):
    if config == "c4":
        w = dict(name="config4", grid_n=256, extent=3.0, n_pix_x=512, n_pix_y=512 * n_gpus, x_fov=1.44,
                 z_obs=3.0, f0=75e6, f1=1.5e9, n_freq=8, scaling="weak")
    elif config == "c5":
        w = dict(name="config5", grid_n=512, extent=4.0, n_pix_x=2048, n_pix_y=2048, x_fov=2.8, z_obs=4.0,
                 f0=20e6, f1=300e6, n_freq=16, scaling="strong")
    elif config == "c3":
        w = dict(name="config3", grid_n=128, extent=3.0, n_pix_x=64, n_pix_y=64 * n_gpus, x_fov=1.44, z_obs=3.0,
                 f0=75e6, f1=75e6, n_freq=1, scaling="weak")
    else:
        raise SystemExit(f"unknown --config {config}")
    step = np.log10(w["f1"] / w["f0"]) / max(w["n_freq"] - 1, 1)
    freqs = synthetic.log_frequencies(w["f0"], w["n_freq"], step)
    if config == "c3":
        fps = [dict(freq_hz=75e6, dt=6e-3, n_steps=5000, record_stride=10)]
    else:
        fps = [dict(freq_hz=float(f), **synthetic.frequency_scaled_params(float(f))) for f in freqs]
    w["freq_params"] = fps
    w["log_step"] = step
    return w


def rays_of(w, idx=None):
    xs, ys, zs, _ = synthetic.ray_launch_geometry(w["n_pix_x"], w["x_fov"], w["z_obs"], N_pix_y=w["n_pix_y"])
    if idx is not None:
        xs, ys, zs = xs[idx], ys[idx], zs[idx]
    return xs, ys, zs


def pixel_area(w):
    # pixel area of the N=1 map (script/resample_with_ray_tracing.py:360-363); the N>1 map samples
    # the same field of view more finely in y
    dx = 2 * w["x_fov"] / w["n_pix_x"] * R_SUN_CM
    dy = 2 * w["x_fov"] / w["n_pix_y"] * R_SUN_CM
    return dx * dy


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [t.strip() for t in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); smax.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_reference_step(w, sample_stride, n_threads=0):
    """The reference's CPU algorithm (oracle port: ray_trace -> sampler -> GET_MW per pixel) on a
    pixel sub-grid of the same workload.  Returns (nominal ray-steps, seconds, n_rays)."""
    from oracle import oracle
    c = cpu_reference_step.cube
    nx, ny = w["n_pix_x"], w["n_pix_y"]
    sel = (np.arange(0, ny, sample_stride)[:, None] * nx + np.arange(0, nx, sample_stride)[None, :]).ravel()
    xs, ys, zs = rays_of(w, sel)
    kv = np.tile([[0.0, 0.0, -1.0]], (len(xs), 1))
    ray_start = np.column_stack([xs, ys, zs])
    nominal = 0
    t0 = time.perf_counter()
    for p in w["freq_params"]:
        r, cs = oracle.ray_trace(c["omega_pe"], c["x_grid"], c["y_grid"], c["z_grid"], p["freq_hz"], xs, ys, zs, kv,
                                 p["dt"], p["n_steps"], p["record_stride"], True, perturb_ratio=2, n_threads=n_threads)
        smp = oracle.sample_model_with_rays_cpu(c["x_grid"], c["y_grid"], c["z_grid"], c["ne"], c["te"], c["b"], r,
                                                np.array(cs), ray_start, R_SUN_CM)
        n_side = len(range(0, nx, sample_stride))
        oracle.emission_from_samples(smp, n_side, w["x_fov"], p["freq_hz"])
        nominal += len(xs) * p["n_steps"]
        del r, cs, smp
    return nominal, time.perf_counter() - t0, len(xs)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    w = workload(args.config, 1)          # the CPU sample is always taken from the square N=1 image
    cpu_reference_step.cube = synthetic.corona_cube(w["grid_n"], w["extent"], active_region=True)
    stride = args.cpu_sample_stride
    for _ in range(args.warmup):
        cpu_reference_step(w, stride * 4)
    tot_steps, tot_t, n_rays = 0, 0.0, 0
    for _ in range(args.steps):
        n, t, n_rays = cpu_reference_step(w, stride)
        tot_steps += n
        tot_t += t
    value = tot_steps / tot_t
    cores = os.cpu_count()
    sample = (f"{n_rays} rays (every {stride}th pixel in x and y of the {w['n_pix_x']}x{w['n_pix_y']} image) x "
              f"{w['n_freq']} freqs, full n_steps; trace + sampler + GET_MW; theta=90, free-free (reference packing)")
    line = {
        "impl": "reference", "metric": "ray_steps_per_s", "value": value, "unit": "ray-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(w, args),
        "cpu_baseline": {"value": value, "unit": "ray-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def config_dict(w, args):
    return {
        "workload": f"BASELINE {w['name']}: {w['n_pix_x']}x{w['n_pix_y']} px, {w['grid_n']}^3 cube, {w['n_freq']} freqs "
                    f"{w['f0'] / 1e6:g}-{w['f1'] / 1e6:g} MHz, GR+FF, cross-sections on, fused trace+sample+transfer",
        "per_freq": [[p["freq_hz"], p["dt"], p["n_steps"], p["record_stride"]] for p in w["freq_params"]],
        "sharding": "rows interleaved over ranks in groups of 8, cube replicated, NCCL all-gather of the image",
        "thread_order": "4x8-pixel tiles per warp (ray_order), results in row-major ray order",
        "l2": f"cubes (3 x {16 * w['grid_n'] ** 3 / 1e6:.0f} MB float4 + {128 * w['grid_n'] ** 3 / 1e6:.0f} MB cell-major polynomial "
              "cube) exceed the 126 MB L2; no flush between steps",
        "cubes": "built on the GPU from a spherical (phi,lat,r) model" if (w["name"] == "config5" or args.spherical)
                 else "analytic corona, uploaded from pinned host memory",
        "precision": "FP64 ray state and transfer, FP32 cube storage and cell-relative RHS, FP32 angle to B",
        "cross_sections": "pencil rays traced on recorded steps only (the reference computes S at every step but "
                          "outputs only recorded steps, build_rays.py:241-244); RTGRFF_CS_EVERY_STEP=1 restores it",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=["c3", "c4", "c5"])
    ap.add_argument("--cpu-sample-stride", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--spherical", action="store_true", help="build the cubes on the GPU from a spherical model")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    w = workload(args.config, args.gpus)
    if args.impl == "reference":
        return run_reference(args, w)

    # keep stdout clean for the single JSON line: anything libraries print there (NCCL's version banner
    # among others) goes to stderr until the result is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from raytracinggrff_b200 import RaySession, _lib
    from raytracinggrff_b200 import dist as rdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # config 5 (512^3): the cubes are built on the GPU from a spherical (phi, lat, r) model — the
    # reference's own pipeline shape (MAS model -> resample -> trace) — instead of 6 x 1 GB host cubes
    spherical = w["name"] == "config5" or args.spherical
    cube = None if spherical else synthetic.corona_cube(w["grid_n"], w["extent"], active_region=True)
    stream = torch.cuda.current_stream().cuda_stream
    ses = RaySession(context=_lib.Context(local_rank, stream))
    idx, rows = rdist.shard_rays(w["n_pix_x"], w["n_pix_y"], world, rank)
    xs, ys, zs = rays_of(w, idx)
    n_local = len(xs)
    nf = w["n_freq"]
    mr = rdist.max_rows_per_rank(w["n_pix_y"], world)
    area = pixel_area(w)
    fps = w["freq_params"]
    nominal_local = n_local * sum(p["n_steps"] for p in fps)
    nominal_total = w["n_pix_x"] * w["n_pix_y"] * sum(p["n_steps"] for p in fps)

    # pinned host copies of the inputs (what a caller of the public API holds)
    def pinned(a, dtype):
        t = torch.empty(a.shape, dtype=dtype, pin_memory=True)
        t.copy_(torch.from_numpy(np.ascontiguousarray(a)))
        return t.numpy()

    if spherical:
        model = synthetic.spherical_corona(200, 140, 160, r_max=1.8 * w["extent"], active_region=True)
        gline = np.linspace(-w["extent"], w["extent"], w["grid_n"])
        grids = (gline, gline.copy(), gline.copy())
        h2d_cubes = sum(v.data.nbytes + v.phi.nbytes + v.lat.nbytes + v.r.nbytes for v in model.values())

        def upload():
            ses.set_model_from_spherical(model, *grids, phi0_offset=0.0, want_bvec=True)
    else:
        h_w = pinned(cube["omega_pe"], torch.float64)
        h_f = {k: pinned(cube[k], torch.float32) for k in ("ne", "te", "b", "bx", "by", "bz")}
        grids = (cube["x_grid"], cube["y_grid"], cube["z_grid"])
        h2d_cubes = h_w.nbytes + sum(a.nbytes for a in h_f.values())

        def upload():
            ses.set_omega_cube(h_w, *grids)
            ses.set_field_cubes(*grids, h_f["ne"], h_f["te"], h_f["b"], h_f["bx"], h_f["by"], h_f["bz"])

    # device-side image slabs: [2 (tb, vi)][freq][rows_max][n_pix_x]
    slab = torch.zeros((2, nf, mr, w["n_pix_x"]), dtype=torch.float64, device=dev)
    stats_box = {}

    def render():
        # the rank's rows fill the first len(rows) rows of the slab: [freq][ray] with ray row-major
        if len(rows) == mr:
            tb_ptr, vi_ptr = slab[0].data_ptr(), slab[1].data_ptr()
            _, _, st = ses.render_map(xs, ys, zs, fps, trace_crosssections=True, perturb_ratio=2.0,
                                      pixel_area_cm2=area, r_sun_cm=R_SUN_CM, em_flag=4, s_max=30, use_bvec=True,
                                      out_device_ptrs=(tb_ptr, vi_ptr), image_shape=(len(rows), w["n_pix_x"]))
        else:   # ragged share: render into a compact buffer, then place
            tmp = torch.empty((2, nf, len(rows), w["n_pix_x"]), dtype=torch.float64, device=dev)
            _, _, st = ses.render_map(xs, ys, zs, fps, trace_crosssections=True, perturb_ratio=2.0,
                                      pixel_area_cm2=area, r_sun_cm=R_SUN_CM, em_flag=4, s_max=30, use_bvec=True,
                                      out_device_ptrs=(tmp[0].data_ptr(), tmp[1].data_ptr()),
                                      image_shape=(len(rows), w["n_pix_x"]))
            slab[:, :, :len(rows)] = tmp
        stats_box.update(st)
        stats_box["kernel_ms"] = ses.ctx.last_kernel_ms
        if world > 1:
            return rdist.gather_rows(slab, w["n_pix_y"])
        return slab

    upload()
    for _ in range(args.warmup):
        render()
    # ---- timed region: device-resident inputs ----
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    launches0 = ses.ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    active_local = pencil_local = samples_local = 0
    ev0.record()
    for _ in range(args.steps):
        img = render()
        kernel_ms.append(stats_box["kernel_ms"])
        active_local = stats_box["active_ray_steps"]
        pencil_local = stats_box["pencil_steps"]
        samples_local = stats_box["valid_samples"]
    ev1.record()
    barrier()
    clk = clocks.stop()
    launches = ses.ctx.launch_count - launches0
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms, float(active_local), float(np.mean(kernel_ms)), float(pencil_local), float(samples_local)],
                     dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, active_total, kernel_ms_max = float(tmax[0]), float(tsum[1]), float(tmax[2])
        pencil_total, samples_total = float(tsum[3]), float(tsum[4])
    else:
        active_total, kernel_ms_max = float(active_local), float(t[2])
        pencil_total, samples_total = float(pencil_local), float(samples_local)
    ms_per_step = ms / args.steps
    value = nominal_total / (ms_per_step * 1e-3)

    # ---- end to end: host cubes in, host maps out, every step ----
    h2d = h2d_cubes + 3 * 8 * n_local
    d2h = 2 * nf * w["n_pix_x"] * w["n_pix_y"] * 8 if rank == 0 else 0
    for _ in range(1):
        upload(); render()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.e2e_steps):
        upload()
        img = render()
        if rank == 0:
            host_img = img.cpu()
    e1.record()
    barrier()
    e2e_wall = (time.perf_counter() - t0) / args.e2e_steps
    e2e_ms = torch.tensor([e0.elapsed_time(e1) / args.e2e_steps, e2e_wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_t = float(e2e_ms.max()) * 1e-3
    e2e_value = nominal_total / e2e_t

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        traffic = None      # dram__bytes of one launch from the committed ncu capture of exactly this workload
        try:
            tj = json.load(open(ROOT / "profiles" / "roofline_traffic.json"))
            if tj.get("workload") == f"{w['name']}/{world}gpu":
                traffic = tj.get("render_map_kernel_dram_bytes_per_launch")
        except Exception:
            pass
        # algorithmic bytes per launch (one launch per rank per step), SURVEY 8d per-unit figures x the units
        # actually processed: 512 B per central RK4 step, 1024 B per step on which the two pencil rays are
        # traced (recorded steps only: the reference discards the cross-section ratio of the others),
        # 2 x 128 B of field/B-vector corners + 16 B of state per valid sample
        alg_bytes = (active_total * BYTES_PER_RAY_STEP_NOCS + pencil_total * (BYTES_PER_RAY_STEP_CS - BYTES_PER_RAY_STEP_NOCS)
                     + samples_total * (2 * 128 + 16)) / world
        achieved = alg_bytes / (kernel_ms_max * 1e-3) / 1e9
        line = {
            "metric": "ray_steps_per_s", "value": value, "unit": "ray-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(w, args),
            "full_map_wall_s": ms_per_step * 1e-3,
            "nominal_ray_steps_per_step": nominal_total, "active_ray_steps_per_step": active_total,
            "active_ray_steps_per_s": active_total / (ms_per_step * 1e-3),
            "pencil_steps_per_step": pencil_total, "valid_samples_per_step": samples_total,
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "ray-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "full_map_wall_s": e2e_t},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "render_map_kernel", "kernel_ms": kernel_ms_max,
                         "peak_source": peak_src,
                         "note": "achieved = (512 B x active central steps + 1024 B x pencil steps + 272 B x valid "
                                 "samples) / kernel time; the gathers are served by registers/L1 (DRAM traffic ~ cube "
                                 "size), so the algorithmic rate can exceed the HBM copy peak; see DESIGN.md 4.1"},
        }
        if not args.no_cpu_baseline:
            from oracle import oracle
            oracle.build()
            cpu_reference_step.cube = cube if cube is not None else synthetic.corona_cube(w["grid_n"], w["extent"], active_region=True)
            n, tcpu, n_rays = cpu_reference_step(workload(args.config, 1), args.cpu_sample_stride)
            line["cpu_baseline"] = {
                "value": n / tcpu, "unit": "ray-steps/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"{n_rays} rays (every {args.cpu_sample_stride}th pixel in x and y) x {nf} freqs, full n_steps; "
                          f"oracle trace + sampler + GET_MW (theta=90 free-free packing), {tcpu:.1f} s"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
