#!/usr/bin/env python
"""bench.py -- vertex-frames deformed per second on B200, beside the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference]

A step is one pass of the hot path over one batch of synthetic input: assemble + factor the control-point
system, solve all 3F right-hand sides, evaluate every vertex for every frame (BASELINE.json configs[1] = C2:
256 control points, 100k vertices, 240 frames, Gaussian, linear term).  `value` times it with inputs and
outputs resident in HBM; `e2e` times the same step through the host-pointer C ABI (pinned host buffers, H2D
and D2H inside the timed region).  N > 1 (torchrun): weak scaling, every rank owns a 100k-vertex range of an
N x 100k mesh, rank 0 factors and solves, the weights are broadcast with NCCL, results stay sharded.
`--impl reference` times the CPU oracle (the reference's algorithm restated; the reference itself cannot be
compiled here: Houdini HDK, ALGLIB and Eigen are absent) on the host cores on a bounded vertex sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "vertex_frames_per_s"
UNIT = "vertex-frames/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    sm_max=float(p.get("sm_max_mhz", 1965.0)), source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, sm_max=1965.0, source="fallback (B200_PROFILING.md)")


def workload(name: str, world: int):
    from facedeform_b200 import synth
    cfg = dict(synth.CONFIGS[name])
    rig = synth.control_rig(cfg["N"])
    deform = synth.deformed_rig(rig, cfg["F"])
    radius = synth.default_radius(cfg["kernel"], rig.spacing)
    return cfg, rig, deform, radius


class ClockSampler:
    """samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def sample(self):
        if not self.nv:
            return
        nv = self.nv
        try:
            self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                     0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        def loop():
            while not self._stop.is_set():
                self.sample()
                time.sleep(0.01)
        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_sample(cfg, rig, deform, radius, seconds_target=12.0, nthreads=None):
    """the oracle on the host cores over a bounded vertex sample of the same workload (all F frames)."""
    from facedeform_b200 import synth
    from oracle import fd_oracle as o
    nthreads = nthreads or o.num_threads()
    p = o.make_params(model=o.MODEL_ML, term=synth.TERMS[cfg["term"]], kernel=synth.KERNELS[cfg["kernel"]],
                      radius=radius, **{"lambda": 0.0})
    t0 = time.perf_counter()
    st, rad, W = o.fit(p, rig.rest, deform)
    t_fit = time.perf_counter() - t0
    assert st == 1
    # calibrate: pairs/s on a small probe, then size the sample for ~seconds_target of CPU work
    probe = synth.face_mesh(2048, topology=False).P
    t0 = time.perf_counter()
    o.evaluate(p, rig.rest, rad, W, probe, nthreads=nthreads)
    t_probe = max(time.perf_counter() - t0, 1e-4)
    vs = int(min(cfg["V"], max(2048, 2048 * seconds_target / t_probe)))
    mesh = synth.face_mesh(cfg["V"], topology=False)
    idx = np.random.default_rng(5).choice(cfg["V"], vs, replace=False) if vs < cfg["V"] else np.arange(cfg["V"])
    Ps = np.ascontiguousarray(mesh.P[np.sort(idx)])
    t0 = time.perf_counter()
    o.evaluate(p, rig.rest, rad, W, Ps, nthreads=nthreads)
    t_eval = time.perf_counter() - t0
    # whole-step rate of the CPU path on the sample: fit + solve amortised over the full V, eval measured
    t_step_full = t_fit + t_eval * (cfg["V"] / vs)
    value = cfg["V"] * cfg["F"] / t_step_full
    sample = (f"{vs} of {cfg['V']} vertices x {cfg['F']} frames, N={cfg['N']} (fit+solve {t_fit * 1e3:.1f} ms measured in full, "
              f"eval {t_eval:.2f} s on the sample, extrapolated linearly in V)")
    return dict(value=value, unit=UNIT, cores=nthreads, kind="port", sample=sample, eval_s=t_eval, fit_s=t_fit,
                sample_vertices=vs)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg, rig, deform, radius = workload(args.config, 1)
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_sample(cfg, rig, deform, radius, seconds_target=args.cpu_seconds)
        if i >= args.warmup:
            vals.append(r)
    value = float(np.mean([r["value"] for r in vals]))
    ms = cfg["V"] * cfg["F"] / value * 1e3
    last = vals[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(cfg), "note": "CPU oracle (restatement of the reference path; the reference needs HDK/ALGLIB/Eigen and cannot be built)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def factor_table(ctx, sizes, cpu=True):
    """factor ms at N control points (assemble + factor, device resident, median of 3 after one warm-up):
    `gaussian_spd` = Gaussian + uniform radius (symmetric positive definite: fused LU without pivot search),
    `multiquadric_nullspace` = multiquadric + linear term through the null-space transform onto the same fused LU,
    `qnn_pivoted` = Gaussian with per-centre QNN radii (non-symmetric: the general pivoted LU);
    `cpu_oracle` = the FP64 oracle's assemble + LU on the host cores (N <= 2048 only: 8192 takes minutes on a CPU)."""
    import torch
    from facedeform_b200 import make_params, synth
    out = {}
    for n in sizes:
        rig = synth.control_rig(n)
        d_rest = torch.from_numpy(rig.rest).cuda()
        row = {}
        for name, kern, model in (("gaussian_spd", "gaussian", 1), ("multiquadric_nullspace", "multiquadric", 1),
                                  ("qnn_pivoted", "gaussian", 0)):
            p = make_params(model=model, term=0, kernel=synth.KERNELS[kern], radius=synth.default_radius(kern, rig.spacing),
                            **{"lambda": 0.0})
            ts = []
            for i in range(4):
                m = ctx.fit(p, d_rest)
                m.report()
                if i:
                    ts.append(ctx.phase_ms("assemble") + ctx.phase_ms("factor"))
                m.close()
            row[name] = float(np.median(ts))
        if cpu and n <= 2048:
            from oracle import fd_oracle as o
            op = o.make_params(model=1, term=0, kernel=0, radius=synth.default_radius("gaussian", rig.spacing), **{"lambda": 0.0})
            st, rad = o.radii(op, rig.rest)
            t0 = time.perf_counter()
            A = o.assemble(op, rig.rest, rad)
            o.lu_factor(A)
            row["cpu_oracle"] = (time.perf_counter() - t0) * 1e3
        out[str(n)] = row
    return out


def configs_table(ctx):
    """phase times of BASELINE.json's other single-GPU configurations (C1, C3 multiquadric / thin plate, the C3 shape
    with the Gaussian kernel, a slice of C5), device resident -- parity-test cases reported for context, not bench lines."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("configs_probe", os.path.join(ROOT, "profiles", "tools", "configs_probe.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import contextlib
    import io
    rows = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for name, N, V, F, kern in (("C1", 64, 10_000, 1, "gaussian"), ("C3", 2048, 1_000_000, 1, "multiquadric"),
                                    ("C3_thin_plate", 2048, 1_000_000, 1, "thin_plate"),
                                    ("C3_shape_gaussian", 2048, 1_000_000, 1, "gaussian"),
                                    ("C5_slice", 4096, 65_536, 1000, "gaussian")):
            r = mod.run(ctx, name, N, V, F, kern, 3)
            rows[name] = {k: r[k] for k in ("N", "V", "F", "kernel", "assemble_ms", "factor_ms", "solve_ms", "eval_ms",
                                            "eval_vertex_frames_per_s", "eval_alg_tflops")}
    return rows


def describe(cfg):
    return (f"{cfg['N']} control points, {cfg['V']} vertices/GPU, {cfg['F']} frames, {cfg['kernel']} kernel, "
            f"{cfg['term']} term, step = assemble + LU + {3 * cfg['F']}-RHS solve + fused eval")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-path", type=int, default=0)
    ap.add_argument("--weights", default="auto", choices=["auto", "broadcast", "replicated"],
                    help="N > 1: root solves + NCCL broadcast of the weights, or every rank solves the small system itself")
    ap.add_argument("--no-configs-table", action="store_true", help="skip the phase times of C1 / C3 / C5-slice")
    ap.add_argument("--factor-sizes", default="256,1024,2048,4096,8192",
                    help="control-point counts for the factor-ms table (second half of the metric); empty to skip")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from facedeform_b200 import Context, make_params, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    json_fd = None
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version banner to stdout (C level): route fd 1 to stderr for the run, the JSON line goes to the
        # saved descriptor at the end, so stdout carries exactly one line
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    cfg, rig, deform, radius = workload(args.config, world)
    N, V, F = cfg["N"], cfg["V"], cfg["F"]
    # weak scaling: the mesh has world x V vertices, this rank owns the contiguous range of V of them
    mesh = synth.face_mesh(V * world, topology=False)
    b, e = shard.vertex_range(V * world, rank, world)
    P_host = np.ascontiguousarray(mesh.P[b:e])
    params = make_params(model=1, term=synth.TERMS[cfg["term"]], kernel=synth.KERNELS[cfg["kernel"]], radius=radius,
                         eval_path=args.eval_path, **{"lambda": 0.0})

    # a dedicated (non-default) stream shared by torch and the library, so torch's CUDA events bracket the kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = Context(local, stream=stream.cuda_stream)
    d_rest = torch.from_numpy(rig.rest).to(dev)
    d_deform = torch.from_numpy(deform).to(dev)
    d_P = torch.from_numpy(P_host).to(dev)
    d_out = torch.empty((F, V, 3), dtype=torch.float32, device=dev)
    d_fall = torch.empty((V,), dtype=torch.float32, device=dev)

    wmode = shard.weights_mode(N, args.weights) if world > 1 else "single"

    def step_device():
        """one pass, everything resident in HBM: fit + solve (on rank 0 + NCCL broadcast, or replicated on every rank
        for small systems, shard.weights_mode), then the sharded eval."""
        if rank == 0 or wmode == "replicated":
            m = ctx.fit(params, d_rest)
            m.solve(d_deform)
        else:
            m = ctx.receiver(params, d_rest, F)
        if wmode == "broadcast":
            shard.broadcast_model(m, 0, shared_stream=True)
        m.eval(d_P, out=d_out, falloff_out=d_fall)
        return m

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device().close()
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phases = {"assemble": [], "factor": [], "solve": [], "eval": []}
    ev0.record(stream)
    prev = None
    for _ in range(args.steps):
        if prev is not None:
            prev.close()          # stream-ordered free: the next cook reuses the blocks without a device sync
        prev = step_device()
    ev1.record(stream)
    sync_all()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    for k in phases:
        phases[k] = ctx.phase_ms(k)
    prev.close()
    # the dominant kernel alone (the eval launch), timed with CUDA events on its stream, same resident buffers
    m = step_device()
    torch.cuda.synchronize()
    eval_ms = []
    for _ in range(max(5, min(args.steps, 20))):
        m.eval(d_P, out=d_out, falloff_out=d_fall)
        eval_ms.append(ctx.phase_ms("eval"))
    eval_ms_mean = float(np.mean(eval_ms))
    m.close()

    # ---- end to end through the host-pointer C ABI ------------------------------------------------------------
    h_rest = torch.from_numpy(rig.rest).pin_memory().numpy()
    h_deform = torch.from_numpy(deform).pin_memory().numpy()
    h_P = torch.from_numpy(P_host).pin_memory().numpy()
    h_out = torch.empty((F, V, 3), dtype=torch.float32).pin_memory().numpy()
    h_fall = torch.empty((V,), dtype=torch.float32).pin_memory().numpy()

    def step_e2e():
        if rank == 0 or wmode == "replicated":
            m = ctx.fit(params, h_rest)
            m.solve(h_deform)
        else:
            m = ctx.receiver(params, h_rest, F)
        if wmode == "broadcast":
            shard.broadcast_model(m, 0, shared_stream=True)
        m.eval(h_P, out=h_out, falloff_out=h_fall)
        m.close()

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    sync_all()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    clocks = sampler.stop()

    # max over ranks
    t = torch.tensor([ms_total, e2e_s, eval_ms_mean], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s, eval_ms_mean = (float(x) for x in t.tolist())
    ms_step = ms_total / args.steps
    units = float(V) * world * F  # vertex-frames all ranks processed per step
    value = units / (ms_step * 1e-3)
    e2e_value = units / e2e_s

    if rank == 0:
        peaks = load_peaks()
        pairs = float(V) * N
        # algorithmic work of one eval launch = the contraction Phi[V x N] . W[N x 3F] (2 flop per MAC)
        # plus 8 + 2 FP32 flop per (vertex, centre) pair for the distance and the kernel
        alg_flops = pairs * (2.0 * 3 * F + 8.0 + 2.0)
        achieved_tf = alg_flops / (eval_ms_mean * 1e-3) / 1e12
        alg_bytes = V * 12.0 + V * F * 12.0 + V * 4.0
        tensor_path = args.eval_path != 1 and 3 * F >= 48
        if tensor_path:
            roofline = {
                "kernel": "tc::k_eval_tc (tcgen05.mma kind::f16, FP16 hi/lo splits: 3 MMAs per algorithmic MAC)",
                "bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["bf16"], "peak_source": peaks["source"] + " bf16_tflops (burst; FP16 = BF16 rate)",
                "note": "algorithmic flops; the split-precision scheme issues 3x that on the tensor pipe, so 1/3 is the ceiling of this fraction",
                "traffic": None, "launch_ms": eval_ms_mean,
            }
        else:
            fp32_peak_tf = 148 * 128 * 2 * peaks["sm_max"] * 1e6 / 1e12
            roofline = {
                "kernel": "k_eval_simt", "bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / fp32_peak_tf,
                "peak_source": "derived: 148 SM x 128 FP32 lanes x 2 flop x clocks.max.sm (FP32 FMA issue is not in MEASURED_PEAKS.json)",
                "traffic": None, "launch_ms": eval_ms_mean,
            }
        # DRAM traffic of one launch from the committed `ncu --set full` capture of this command (profiles/), C2 only
        try:
            if args.config == "C2" and tensor_path:
                with open(os.path.join(ROOT, "profiles", "r1e_eval_tc_ncu_summary.json")) as f:
                    nc = json.load(f)
                roofline["traffic"] = (float(nc["dram__bytes_read.sum"][0]) + float(nc["dram__bytes_write.sum"][0])) * 1e6
                roofline["traffic_source"] = "profiles/r1e_eval_tc_ncu_summary.json (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch)"
                roofline["algorithmic_bytes"] = alg_bytes
        except Exception:
            pass
        roofline["hbm"] = {"achieved": alg_bytes / (eval_ms_mean * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                           "frac": alg_bytes / (eval_ms_mean * 1e-3) / 1e9 / peaks["hbm"], "peak_source": peaks["source"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": describe(cfg), "name": args.config,
                       "l2": "per-step output (F x V x 12 B = %.0f MB) exceeds the 126 MB L2; no explicit flush" % (F * V * 12 / 1e6),
                       "precision": "FP64 assemble/factor/solve, FP32 evaluation", "parallelism": f"vertex-range x{world}" + ("" if world == 1 else f", weights {wmode}")},
            "phase_ms_last_step": phases, "factor_ms": {"n_ctrl": N, "assemble": phases["assemble"], "factor": phases["factor"],
                                                        "solve": phases["solve"]},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(h_rest.nbytes + h_deform.nbytes + h_P.nbytes),
                    "d2h_bytes_per_step": int(h_out.nbytes + h_fall.nbytes)},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and args.factor_sizes:
            line["factor_ms_by_n"] = factor_table(ctx, [int(x) for x in args.factor_sizes.split(",") if x],
                                                  cpu=not args.no_cpu_baseline)
        if world == 1 and not args.no_configs_table:
            line["other_configs"] = configs_table(ctx)
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_sample(cfg, rig, deform, radius, seconds_target=args.cpu_seconds)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if json_fd is None:
            print(json.dumps(line), flush=True)
        else:
            os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
