#!/usr/bin/env python
"""bench.py -- vertex-frames deformed per second on B200, beside the reference CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference]

A step is one pass of the hot path over one batch of synthetic input: assemble + factor the control-point
system, solve all 3F right-hand sides, evaluate every vertex for every frame (BASELINE.json configs[1] = C2:
256 control points, 100k vertices, 240 frames, Gaussian, linear term).  `value` times it with inputs and
outputs resident in HBM; `e2e` times the same step through the host-pointer C ABI (pinned host buffers, H2D
and D2H inside the timed region).  N > 1 (torchrun): weak scaling, every rank owns a 100k-vertex range of an
N x 100k mesh, rank 0 factors and solves, the weights are broadcast with NCCL INSIDE the timed region (`comm` in the
JSON line carries its bytes and device time), results stay sharded.  `--config C5` (BASELINE configs[4]: 16M vertices
x 4096 control points x 1000 frames) is STRONG scaling: the 16M vertices are split over the ranks, the frames go
through in chunks (`--frame-chunk`, the 192 GB result cannot be resident at once), every chunk = solve on rank 0 +
NCCL broadcast + sharded eval.
`--impl reference` times the CPU oracle (the reference's algorithm restated; the reference itself cannot be
compiled here: Houdini HDK, ALGLIB and Eigen are absent) on the host cores on a bounded vertex sample: all cores
(the line's value; the thread count comes from sched_getaffinity, never from OMP_NUM_THREADS, which torchrun
overrides) and 1 thread (what the reference does: NO_RBF_THREADS, SOP_FaceDeform.hpp:11).
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


def host_cores() -> int:
    """the cores this process may run on -- NOT OMP_NUM_THREADS: torch.distributed.run exports OMP_NUM_THREADS=1"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(cfg, rig, deform, radius, seconds_target=12.0, nthreads=None):
    """the oracle on the host cores over a bounded vertex sample of the same workload (all F frames)."""
    from facedeform_b200 import synth
    from oracle import fd_oracle as o
    nthreads = nthreads or host_cores()
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
    if cfg["F"] > 240:  # C5: the CPU sample covers a 240-frame chunk (the oracle's cost is linear in F)
        cfg = dict(cfg, F=240)
        deform = deform[:240]
    cores = host_cores()
    vals, vals1 = [], []
    for i in range(args.warmup + args.steps):
        r = cpu_sample(cfg, rig, deform, radius, seconds_target=args.cpu_seconds * 0.6, nthreads=cores)
        r1 = cpu_sample(cfg, rig, deform, radius, seconds_target=args.cpu_seconds * 0.4, nthreads=1)
        if i >= args.warmup:
            vals.append(r)
            vals1.append(r1)
    value = float(np.mean([r["value"] for r in vals]))
    value1 = float(np.mean([r["value"] for r in vals1]))
    ms = cfg["V"] * cfg["F"] / value * 1e3
    last, last1 = vals[-1], vals1[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(cfg), "note": "CPU oracle (restatement of the reference path; the reference needs HDK/ALGLIB/Eigen and cannot be built)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
        "cpu_baseline_1thread": {"value": value1, "unit": UNIT, "cores": 1, "kind": "port", "sample": last1["sample"],
                                 "note": "what the reference does: NO_RBF_THREADS (SOP_FaceDeform.hpp:11)"},
        "host_cores": cores, "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_subprocess(args):
    """the CPU leg in its own process: the product process never loads oracle/libfd_oracle.so"""
    import subprocess
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMP_NUM_THREADS"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0", "--config",
           args.config, "--cpu-seconds", str(args.cpu_seconds)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            j = json.loads(ln)
            return j["cpu_baseline"], j.get("cpu_baseline_1thread")
    raise RuntimeError("cpu baseline subprocess failed: " + r.stderr[-500:])


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
        out[str(n)] = row
    if cpu:  # the oracle's assemble + LU, in a subprocess (the product process never loads the oracle)
        import subprocess
        code = ("import sys, time, json; sys.path.insert(0, %r)\n"
                "from oracle import fd_oracle as o\nfrom facedeform_b200 import synth\nres = {}\n"
                "for n in %r:\n"
                "    rig = synth.control_rig(n)\n"
                "    op = o.make_params(model=1, term=0, kernel=0, radius=synth.default_radius('gaussian', rig.spacing), **{'lambda': 0.0})\n"
                "    st, rad = o.radii(op, rig.rest)\n"
                "    t0 = time.perf_counter(); A = o.assemble(op, rig.rest, rad); o.lu_factor(A)\n"
                "    res[str(n)] = (time.perf_counter() - t0) * 1e3\n"
                "print(json.dumps(res))\n") % (ROOT, [n for n in sizes if n <= 2048])
        env = dict(os.environ)
        env.pop("OMP_NUM_THREADS", None)
        try:
            r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
            for k, v in json.loads(r.stdout.strip().splitlines()[-1]).items():
                out[k]["cpu_oracle"] = v
        except Exception:
            pass
    return out


def _load_tool(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "profiles", "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def configs_table(ctx):
    """phase times of BASELINE.json's other single-GPU configurations (C1, C3 multiquadric / thin plate, the C3 shape
    with the Gaussian kernel, a slice of C5), C2 with FP32 forced (tensor cores), per-cook solves against a cached
    factorisation and the DirectBSEdit passes -- device resident, reported for context, not bench lines."""
    import contextlib
    import io
    mod = _load_tool("configs_probe")
    rows = {}
    keys = ("N", "V", "F", "kernel", "assemble_ms", "factor_ms", "solve_ms", "eval_ms", "eval_vertex_frames_per_s", "eval_alg_tflops")
    with contextlib.redirect_stdout(io.StringIO()):
        for name, N, V, F, kern, extra in (("C1", 64, 10_000, 1, "gaussian", {}), ("C3", 2048, 1_000_000, 1, "multiquadric", {}),
                                           ("C3_thin_plate", 2048, 1_000_000, 1, "thin_plate", {}),
                                           ("C3_shape_gaussian", 2048, 1_000_000, 1, "gaussian", {}),
                                           ("C5_slice", 4096, 65_536, 1000, "gaussian", {}),
                                           ("C5_slice_fp32_tensor", 4096, 65_536, 1000, "gaussian", {"eval_precision": 1}),
                                           ("C2_fp32_tensor", 256, 100_000, 240, "gaussian", {"eval_precision": 1}),
                                           ("C2_fp64", 256, 100_000, 240, "gaussian", {"eval_precision": 2})):
            r = mod.run(ctx, name, N, V, F, kern, 3, **extra)
            rows[name] = {k: r[k] for k in keys}
            if extra:
                rows[name]["eval_precision"] = ["AUTO", "FP32", "FP64"][extra["eval_precision"]]
        try:
            rows["dbse"] = _load_tool("dbse_dev_probe").run(ctx, 1_000_000, 32)
        except Exception as ex:
            rows["dbse"] = {"failed": str(ex)}
        try:
            rows["per_cook"] = per_cook_rows(ctx)
        except Exception as ex:
            rows["per_cook"] = {"failed": str(ex)}
    return rows


def per_cook_rows(ctx):
    """one frame per cook against a cached factorisation (the reference's usage): solve ms of cook 1 (sweeps), cook 2
    (builds the explicit inverse) and the median of the following ones"""
    import torch
    from facedeform_b200 import make_params, synth
    out = {}
    for N, kern in ((2048, "gaussian"), (2048, "multiquadric"), (8192, "gaussian")):
        rig = synth.control_rig(N)
        d_rest = torch.from_numpy(rig.rest).cuda()
        p = make_params(model=1, term=0, kernel=synth.KERNELS[kern], radius=synth.default_radius(kern, rig.spacing), **{"lambda": 0.0})
        m = ctx.fit(p, d_rest)
        ts = []
        for i in range(6):
            m.solve(torch.from_numpy(synth.deformed_rig(rig, 1, seed=10 + i)).cuda())
            ctx.synchronize()
            ts.append(ctx.phase_ms("solve"))
        m.close()
        out[f"{kern}_{N}"] = {"solve_ms_first": round(ts[0], 4), "solve_ms_builds_inverse": round(ts[1], 4),
                              "solve_ms_steady": round(float(np.median(ts[2:])), 4)}
    return out


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
    ap.add_argument("--eval-precision", type=int, default=0, help="0 AUTO (FP32 while the error bound allows, else FP64), 1 FP32, 2 FP64")
    ap.add_argument("--weights", default="auto", choices=["auto", "broadcast", "replicated"],
                    help="N > 1: root solves + NCCL broadcast of the weights (default), or every rank solves the small system itself")
    ap.add_argument("--frame-chunk", type=int, default=0, help="frames per solve + broadcast + eval pass (0: all; C5 defaults to 250)")
    ap.add_argument("--no-configs-table", action="store_true", help="skip the phase times of C1 / C3 / C5-slice")
    ap.add_argument("--no-e2e", action="store_true")
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
    N, F = cfg["N"], cfg["F"]
    strong = args.config == "C5"  # C5 names a total size (16M vertices): strong scaling; the others: V vertices per GPU
    V_total = cfg["V"] if strong else cfg["V"] * world
    mesh = synth.face_mesh(V_total, topology=False)
    b, e = shard.vertex_range(V_total, rank, world)
    V = e - b
    P_host = np.ascontiguousarray(mesh.P[b:e])
    del mesh
    Fc = args.frame_chunk if args.frame_chunk > 0 else (250 if strong else F)
    Fc = min(Fc, F)
    chunks = [(f0, min(f0 + Fc, F)) for f0 in range(0, F, Fc)]
    params = make_params(model=1, term=synth.TERMS[cfg["term"]], kernel=synth.KERNELS[cfg["kernel"]], radius=radius,
                         eval_path=args.eval_path, eval_precision=args.eval_precision, **{"lambda": 0.0})

    # a dedicated (non-default) stream shared by torch and the library, so torch's CUDA events bracket the kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = Context(local, stream=stream.cuda_stream)
    d_rest = torch.from_numpy(rig.rest).to(dev)
    d_deform = torch.from_numpy(deform).to(dev)
    d_P = torch.from_numpy(P_host).to(dev)
    d_out = torch.empty((Fc, V, 3), dtype=torch.float32, device=dev)   # one frame chunk; C2: all frames
    d_fall = torch.empty((V,), dtype=torch.float32, device=dev)

    # N > 1: the north_star's exchange step -- rank 0 factors and solves, NCCL broadcast of the weights -- unless asked otherwise
    wmode = (args.weights if args.weights != "auto" else "broadcast") if world > 1 else "single"
    receivers = {}

    def step_device():
        """one pass, everything resident in HBM: fit on rank 0 (or on every rank: --weights replicated), then per frame
        chunk solve (+ NCCL broadcast of the weights) + the sharded eval."""
        root = rank == 0 or wmode == "replicated"
        m = ctx.fit(params, d_rest) if root else None
        for (f0, f1) in chunks:
            if root:
                m.solve(d_deform[f0:f1])
                mm = m
            else:
                mm = receivers.get(f1 - f0)
                if mm is None:
                    mm = receivers[f1 - f0] = ctx.receiver(params, d_rest, f1 - f0)
            if wmode == "broadcast":
                shard.broadcast_model(mm, 0, shared_stream=True)
            mm.eval(d_P, out=d_out[: f1 - f0], falloff_out=d_fall)
        return m

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        m = step_device()
        if m is not None:
            m.close()
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
    if prev is not None:
        prev.close()
    # the dominant kernel alone (the eval launch of one frame chunk), timed with CUDA events on its stream, same resident buffers
    m = step_device()
    mm = m if m is not None else receivers[chunks[-1][1] - chunks[-1][0]]
    rep = mm.report()
    torch.cuda.synchronize()
    eval_ms = []
    for _ in range(max(5, min(args.steps, 20))):
        mm.eval(d_P, out=d_out[: mm.frames], falloff_out=d_fall)
        eval_ms.append(ctx.phase_ms("eval"))
    eval_ms_mean = float(np.mean(eval_ms))
    F_eval = mm.frames
    eval_kernel = int(rep.eval_kernel)
    # the exchange step alone: the NCCL broadcast of the weight block + radii (+ the receivers' table build), CUDA events
    comm_ms = 0.0
    comm_bytes = 0
    if wmode == "broadcast":
        wp, wb = mm.weights_dev()
        rp, rb = mm.radii_dev()
        comm_bytes = int(wb + rb)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        c0.record(stream)
        for _ in range(10):
            shard.broadcast_model(mm, 0, shared_stream=True)
        c1.record(stream)
        sync_all()
        comm_ms = c0.elapsed_time(c1) / 10
    if m is not None:
        m.close()

    # ---- end to end through the host-pointer C ABI ------------------------------------------------------------
    e2e_s = copy_s = float("nan")
    h2d_bytes = d2h_bytes = 0
    if not args.no_e2e:
        h_rest = torch.from_numpy(rig.rest).pin_memory().numpy()
        h_deform = torch.from_numpy(deform).pin_memory().numpy()
        h_P = torch.from_numpy(P_host).pin_memory().numpy()
        h_out_t = torch.empty((Fc, V, 3), dtype=torch.float32).pin_memory()
        h_out = h_out_t.numpy()
        h_fall = torch.empty((V,), dtype=torch.float32).pin_memory().numpy()
        d_def_stage = torch.empty((Fc, N, 3), dtype=torch.float32, device=dev)

        def step_e2e():
            root = rank == 0 or wmode == "replicated"
            m = ctx.fit(params, h_rest) if root else None
            for (f0, f1) in chunks:
                if root:
                    m.solve(h_deform[f0:f1])
                    mm = m
                else:
                    mm = receivers[f1 - f0]
                if wmode == "broadcast":
                    shard.broadcast_model(mm, 0, shared_stream=True)
                mm.eval(h_P, out=h_out[: f1 - f0], falloff_out=h_fall)
            if m is not None:
                m.close()

        def step_copy_only():
            """the step's PCIe traffic without any kernel: what bounds e2e (inputs H2D, results D2H, pinned memory)"""
            for (f0, f1) in chunks:
                d_def_stage[: f1 - f0].copy_(torch.from_numpy(h_deform[f0:f1]), non_blocking=True)
                d_P.copy_(torch.from_numpy(h_P), non_blocking=True)
                h_out_t[: f1 - f0].copy_(d_out[: f1 - f0], non_blocking=True)
                torch.cuda.current_stream().synchronize()

        e2e_steps = max(3, min(args.steps, 10)) if not strong else max(1, min(args.steps, 2))
        for _ in range(2 if not strong else 1):
            step_e2e()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        sync_all()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        step_copy_only()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_copy_only()
        sync_all()
        copy_s = (time.perf_counter() - t0) / e2e_steps
        h2d_bytes = int(h_rest.nbytes + h_deform.nbytes + h_P.nbytes * len(chunks))
        d2h_bytes = int(V * F * 12 + h_fall.nbytes * len(chunks))
    clocks = sampler.stop()

    # max over ranks
    t = torch.tensor([ms_total, e2e_s, eval_ms_mean, comm_ms, copy_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s, eval_ms_mean, comm_ms, copy_s = (float(x) for x in t.tolist())
    ms_step = ms_total / args.steps
    units = float(V_total) * F  # vertex-frames all ranks processed per step
    value = units / (ms_step * 1e-3)
    e2e_value = units / e2e_s if e2e_s == e2e_s else None

    if rank == 0:
        peaks = load_peaks()
        pairs = float(V) * N
        # algorithmic work of one eval launch = the contraction Phi[V x N] . W[N x 3F] (2 flop per MAC)
        # plus 8 + 2 flop per (vertex, centre) pair for the distance and the kernel
        alg_flops = pairs * (2.0 * 3 * F_eval + 8.0 + 2.0)
        achieved_tf = alg_flops / (eval_ms_mean * 1e-3) / 1e12
        alg_bytes = V * 12.0 + V * F_eval * 12.0 + V * 4.0
        if eval_kernel == 2:
            roofline = {
                "kernel": "tc::k_eval_tc (tcgen05.mma kind::f16, FP16 hi/lo splits: 3 MMAs per algorithmic MAC)",
                "bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["bf16"], "peak_source": peaks["source"] + " bf16_tflops (burst; FP16 = BF16 rate)",
                "note": "algorithmic flops; the split-precision scheme issues 3x that on the tensor pipe, so 1/3 is the ceiling of this fraction",
                "traffic": None, "launch_ms": eval_ms_mean,
            }
        elif eval_kernel == 4:
            roofline = {
                "kernel": "tcx::k_eval_tcx (tcgen05.mma kind::f16, exact integer leading digit + FP16 mid/lo: 6 or 8 MMAs per "
                          "algorithmic MAC, Phi generated in FP64)",
                "bound": "tensor", "achieved": achieved_tf, "peak": peaks["bf16"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["bf16"], "peak_source": peaks["source"] + " bf16_tflops (burst; FP16 = BF16 rate)",
                "note": "algorithmic flops; the scheme issues 6x (digit width >= 9) or 8x that on the tensor pipe (one N = 240 MMA "
                        "serves two 120-column blocks), so 1/6 or 1/8 is the ceiling of this fraction; the kernel is bound by the "
                        "FP64 generation of Phi and the MMAs' shared-memory operand reads (DESIGN.md section 4)",
                "traffic": None, "launch_ms": eval_ms_mean,
            }
        elif eval_kernel == 3:
            fp64_peak_tf = 148 * 64 * 2 * peaks["sm_max"] * 1e6 / 1e12
            roofline = {
                "kernel": "k_eval64_mma (mma.sync.m8n8k4.f64: FD_EVAL_AUTO chose FP64, the FP32 error bound exceeded the tolerance)"
                          if 3 * F_eval >= 48 else "k_eval_f64 / k_eval_simt<double>",
                "bound": "tensor", "achieved": achieved_tf, "peak": fp64_peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak_tf,
                "peak_source": "derived: 148 SM x 64 FP64 lanes x 2 flop x clocks.max.sm (the FP64 pipe, DMMA and DFMA alike; "
                               "MEASURED_PEAKS.json holds no FP64 figure)",
                "traffic": None, "launch_ms": eval_ms_mean,
            }
        else:
            fp32_peak_tf = 148 * 128 * 2 * peaks["sm_max"] * 1e6 / 1e12
            roofline = {
                "kernel": "k_eval_f32c (FP32 FMA/SFU, packed FFMA2 along the columns)" if F_eval >= 4 else "k_eval_f32x2",
                "bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / fp32_peak_tf,
                "peak_source": "derived: 148 SM x 128 FP32 lanes x 2 flop x clocks.max.sm (FP32 FMA issue is not in MEASURED_PEAKS.json)",
                "traffic": None, "launch_ms": eval_ms_mean,
            }
        # DRAM traffic of one launch from the committed `ncu --set full` capture of this command (profiles/), C2 only
        try:
            if args.config == "C2" and eval_kernel == 1 and os.path.exists(os.path.join(ROOT, "profiles", "r2_eval_f32c_ncu_summary.json")):
                src = "r2_eval_f32c_ncu_summary.json"
                with open(os.path.join(ROOT, "profiles", src)) as f:
                    nc = json.load(f)
                roofline["traffic"] = (float(nc["dram__bytes_read.sum"][0]) + float(nc["dram__bytes_write.sum"][0])) * 1e6
                roofline["traffic_source"] = f"profiles/{src} (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch)"
            elif args.config == "C2" and eval_kernel == 4 and os.path.exists(os.path.join(ROOT, "profiles", "r2l_eval_tcx_ncu_summary.json")):
                src = "r2l_eval_tcx_ncu_summary.json"
                with open(os.path.join(ROOT, "profiles", src)) as f:
                    nc = json.load(f)
                roofline["traffic"] = (float(nc["dram__bytes_read.sum"][0]) + float(nc["dram__bytes_write.sum"][0])) * 1e6
                roofline["traffic_source"] = f"profiles/{src} (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch)"
            elif args.config == "C2" and eval_kernel == 2:
                src = "r2_eval_tc_ncu_summary.json"
                if not os.path.exists(os.path.join(ROOT, "profiles", src)):
                    src = "r1e_eval_tc_ncu_summary.json"
                with open(os.path.join(ROOT, "profiles", src)) as f:
                    nc = json.load(f)
                roofline["traffic"] = (float(nc["dram__bytes_read.sum"][0]) + float(nc["dram__bytes_write.sum"][0])) * 1e6
                roofline["traffic_source"] = f"profiles/{src} (dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch)"
                roofline["algorithmic_bytes"] = alg_bytes
        except Exception:
            pass
        roofline["hbm"] = {"achieved": alg_bytes / (eval_ms_mean * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                           "frac": alg_bytes / (eval_ms_mean * 1e-3) / 1e9 / peaks["hbm"], "peak_source": peaks["source"]}
        dtype = "f64" if eval_kernel == 3 else "f32"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": describe(cfg) if not strong else describe_c5(cfg, V_total, world, Fc), "name": args.config,
                       "l2": "per-pass output (%d frames x V x 12 B = %.0f MB) exceeds the 126 MB L2; no explicit flush" % (Fc, Fc * V * 12 / 1e6),
                       "precision": "FP64 assemble/factor/solve, %s evaluation (eval_precision %s)" % (
                           "FP64" if eval_kernel == 3 else "FP32", ["AUTO", "FP32", "FP64"][args.eval_precision]),
                       "eval_kernel": {1: "FMA/SFU FP32", 2: "tensor cores (FP16 hi/lo)", 3: "FP64",
                                       4: "tensor cores (exact leading digit, FP64-class)"}.get(eval_kernel),
                       "cancellation": float(rep.cancellation),
                       "parallelism": f"vertex-range x{world}" + ("" if world == 1 else f", weights {wmode}")},
            "phase_ms_last_step": phases, "factor_ms": {"n_ctrl": N, "assemble": phases["assemble"], "factor": phases["factor"],
                                                        "solve": phases["solve"]},
            "roofline": roofline,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world > 1:
            line["comm"] = {"collective": "ncclBroadcast (weights FP64 block + radii), root = rank 0" if wmode == "broadcast" else None,
                            "in_timed_region": wmode == "broadcast", "bytes_per_pass": comm_bytes, "passes_per_step": len(chunks),
                            "ms_per_pass": comm_ms, "share_of_step": comm_ms * len(chunks) / ms_step if ms_step > 0 else None,
                            "note": "ms_per_pass = broadcast + the receivers' table build, CUDA events, max over ranks, measured alone after the timed steps"}
        if e2e_value is not None:
            line["e2e"] = {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d_bytes,
                           "d2h_bytes_per_step": d2h_bytes, "copy_only_ms_per_step": copy_s * 1e3,
                           "frac_of_copy_only": copy_s / e2e_s if e2e_s > 0 else None,
                           "note": "copy_only = the same H2D + D2H traffic (pinned) with no kernel: the PCIe / host-memory ceiling of e2e, max over ranks"}
        if world == 1 and args.factor_sizes:
            line["factor_ms_by_n"] = factor_table(ctx, [int(x) for x in args.factor_sizes.split(",") if x],
                                                  cpu=not args.no_cpu_baseline)
        if world == 1 and not args.no_configs_table:
            line["other_configs"] = configs_table(ctx)
        if not args.no_cpu_baseline and world == 1:
            try:
                cb, cb1 = cpu_baseline_subprocess(args)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
                if cb1:
                    line["cpu_baseline_1thread"] = cb1
            except Exception as ex:  # the bench line must not die with the CPU leg
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": f"failed: {ex}"}
        if json_fd is None:
            print(json.dumps(line), flush=True)
        else:
            os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for r in receivers.values():
        r.close()
    ctx.close()
    return 0


def describe_c5(cfg, V_total, world, Fc):
    return (f"{cfg['N']} control points, {V_total} vertices split over {world} GPU(s), {cfg['F']} frames in chunks of {Fc}, "
            f"{cfg['kernel']} kernel, {cfg['term']} term, step = assemble + LU, then per chunk {3 * Fc}-RHS solve + weight broadcast + fused eval")


if __name__ == "__main__":
    sys.exit(main())
