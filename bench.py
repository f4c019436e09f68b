#!/usr/bin/env python
"""Benchmark of the MAP-deconvolution hot path (BASELINE.json metric: MAP iterations/sec, fwd+bwd+Adam).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

One "step" = one MAP iteration with the reference's semantics (jolideco/core.py:214-229): NPred forward
of one dataset, Poisson loss, the full GMM patch prior, their gradients and one Adam update.
Default workload = BASELINE.json configs[1] (256x256, oversample 2, GMM K=256, one dataset).
With N>1 ranks the default workload runs one independent deconvolution per GPU (BASELINE configs[4]:
independent bootstrap/restart runs of the same shape; weak scaling, no data-path collective);
`--workload joint1024|cfg3` runs the dataset-sharded joint deconvolution with one NCCL all-reduce
of the flux gradient per iteration (strong scaling).

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "map_iterations_per_sec"
UNIT = "iter/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--backend", type=int, default=None, help="prior kernel: 0 CUDA cores, 1 tcgen05")
    ap.add_argument("--marginalize", action="store_true", help="logsumexp over components instead of max")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end MAPDeconvolver.run leg (profiling runs)")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed iterations")
    ap.add_argument("--breakdown", action="store_true",
                    help="add per-entry-point device times (CUDA events around every C-ABI call of eager steps)")
    ap.add_argument("--collective", default="nccl", choices=["nccl", "peer"],
                    help="joint multi-GPU step: NCCL all-reduce + Adam, or the fused peer-memory reduce+Adam kernel")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU baseline: the torch port of the reference loop (oracle/torch_port.py), on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference(workload, steps, warmup, budget_s, marginalize, joint=False):
    import torch

    from oracle import jolideco_oracle as O
    from oracle import torch_port as T

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = workload["cfg"]
    datasets = [T.Dataset(d, workload["f"]) for d in workload["datasets"].values()]
    import torch.nn.functional as F

    flux_up = F.interpolate(torch.from_numpy(workload["flux_init"][None, None].astype(np.float32)),
                            scale_factor=workload["f"], mode="bilinear").numpy()[0, 0]
    gmm = None
    if workload["gmm_arrays"] is not None:
        gmm = T.GMM(*workload["gmm_arrays"], O.get_pixel_weights(8, 4))
    loop = T.MapLoop(flux_up, datasets, gmm, marginalize=marginalize)
    rng = np.random.default_rng(0)
    D = len(datasets)

    def one(i):
        sh = rng.integers(-2, 3, size=2)
        if joint:
            loop.joint_step(sh)
        else:
            loop.step(i % D, sh)

    t0 = time.perf_counter()
    for i in range(max(1, min(warmup, 2))):
        one(i)
    t_est = (time.perf_counter() - t0) / max(1, min(warmup, 2))
    n = int(max(2, min(steps, budget_s / max(t_est, 1e-6))))
    t0 = time.perf_counter()
    for i in range(n):
        one(i)
    dt = time.perf_counter() - t0
    return dict(value=n / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"{n} full-size {'joint ' if joint else ''}steps of {workload['name']} "
                       f"(torch {torch.__version__} CPU port of the reference loop, {cores} threads)",
                steps=n, ms_per_step=1e3 * dt / n)


def config_of(workload, args, extra=None):
    cfg = workload["cfg"]
    fH = cfg["H"] * cfg["f"]
    out = {"workload": f"{workload['name']}: {cfg['desc']}", "flux_grid": [fH, fH], "counts_grid": [cfg["H"], cfg["H"]],
           "upsampling": cfg["f"], "psf": [cfg["psf"] * cfg["f"]] * 2, "n_datasets": cfg["D"], "gmm_components": cfg["K"],
           "patches": ((fH - 8) // 4 + 1) ** 2 if cfg["K"] else 0, "marginalize": bool(args.marginalize),
           "l2_flush_between_iterations": not args.no_flush,
           "iteration_semantics": ("joint: all D datasets + one prior + one Adam step (TotalLoss.__call__)"
                                   if workload["name"] in ("joint1024", "cfg3", "cfg4") else
                                   "reference step: one dataset + full prior + Adam (core.py:214-229)")}
    if extra:
        out.update(extra)
    return out


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_run(J, workload, args, device, n_epochs, seed=0, mode="sequential"):
    """(MAPDeconvolver, components) for the workload through the public API."""
    import torch

    prior = J.UniformPrior()
    if workload["gmm_arrays"] is not None:
        gmm = J.GaussianMixtureModel.from_numpy(*workload["gmm_arrays"], meta=J.GaussianMixtureModelMeta(stride=4))
        prior = J.GMMPatchPrior(gmm=gmm, stride=4, generator=torch.Generator().manual_seed(seed),
                                marginalize=args.marginalize, backend=args.backend)
    comps = J.FluxComponents()
    comps["flux"] = J.SpatialFluxComponent.from_numpy(flux=workload["flux_init"], upsampling_factor=workload["f"],
                                                      prior=prior)
    deco = J.MAPDeconvolver(n_epochs=n_epochs, learning_rate=0.1, display_progress=False, device=device,
                            use_cuda_graph=not args.no_graph, mode=mode, collective=args.collective)
    return deco, comps


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from jolideco_b200 import synthetic

    if args.workload == "cfg5":
        return bench_batched(args, rank, local_rank, world)
    joint = args.workload in ("joint1024", "cfg3", "cfg4")
    workload = synthetic.make_workload(args.workload, seed=0 if joint else rank)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        res = cpu_reference(workload, args.steps, args.warmup, 90.0, args.marginalize, joint=joint)
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": res["steps"], "warmup": min(args.warmup, 2), "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "strong" if joint else "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config_of(workload, args),
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (B200)
    import torch
    import torch.distributed as dist

    import jolideco_b200 as J
    from jolideco_b200 import engine as E
    from jolideco_b200 import ops

    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    ops.require_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    pg = dist.group.WORLD if (world > 1 and joint) else None

    # build the engine through the public classes (same path MAPDeconvolver.run takes)
    D_total = workload["cfg"]["D"]
    deco, comps = build_run(J, workload, args, device, n_epochs=1, seed=rank)
    comps = comps.to(device)
    datasets = workload["datasets"]
    if joint and world > 1:  # dataset d -> rank d mod world
        datasets = {k: v for i, (k, v) in enumerate(datasets.items()) if i % world == rank}
    total_loss = J.TotalLoss.from_datasets_and_components(datasets=datasets, components=comps, beta=1.0, device=device)
    n_draws = args.steps * 3 + args.warmup + 64
    eng = deco._build_engine(total_loss, comps, n_draws)
    if pg is not None:
        eng = rebuild_with_group(E, eng, pg, args.collective)
    eng.warmup(joint=joint)
    D_local = len(eng.datasets)

    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def run_step(i):
        if joint:
            eng.joint_step()
        else:
            eng.step(i % D_local)

    for i in range(args.warmup):
        run_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    launches0 = E._STATS["launches"]
    evs = []
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_step(i)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = E._STATS["launches"] - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    iters_per_step = 1
    total_iters = args.steps * (1 if joint else world)
    value = total_iters / (dev_ms / 1e3)

    # ------------------------------------------------------------------ roofline of the dominant kernel
    roofline = None
    if rank == 0 or joint:  # joint steps contain a collective: every rank has to take part
        roofline = measure_roofline(E, eng, run_step, args, workload, flush)

    breakdown = None
    if args.breakdown and rank == 0 and not (joint and world > 1):
        breakdown = measure_breakdown(E, eng, run_step, flush)

    # ------------------------------------------------------------------ e2e through MAPDeconvolver.run (host buffers)
    e2e = None
    if not args.no_e2e:
        e2e_local = measure_e2e(J, workload, args, device, rank, joint)
        t = torch.tensor([e2e_local["seconds"]], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_runs = world if not joint else 1
        e2e = {"value": n_runs * e2e_local["iters"] / float(t.item()), "unit": UNIT,
               "h2d_bytes_per_step": e2e_local["h2d"], "d2h_bytes_per_step": e2e_local["d2h"],
               "what": e2e_local["what"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(workload, args.steps, args.warmup, args.cpu_budget_s, args.marginalize, joint=joint)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if joint else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(workload, args, {"parallelism": (f"datasets sharded over {world} ranks + prior row "
                                                                      f"blocks, NCCL all-reduce of the flux gradient")
                                                     if (joint and world > 1) else
                                                     (f"{world} independent runs, one per GPU" if world > 1 else "1 GPU"),
                                                     "prior_backend": eng.backend if eng.prior else None,
                                                     "cuda_graph": eng.use_graph,
                                                     "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps}),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if breakdown:
            line["breakdown_us_per_step"] = breakdown
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_breakdown(E, eng, run_step, flush, n=10):
    """Device time per C-ABI entry point and step: CUDA events around every call of eager steps (L2 flushed before
    each step like the timed loop; kernels run back to back, so inter-kernel gaps are not included)."""
    import torch

    graph = eng.use_graph
    eng.use_graph = False
    E._STATS["timed"], E._STATS["events"] = "*", []
    for i in range(n):
        if flush is not None:
            flush.fill_(i & 0xFF)
        run_step(i)
    torch.cuda.synchronize()
    out = {}
    for name, a, b in E._STATS["events"]:
        out[name] = out.get(name, 0.0) + a.elapsed_time(b) * 1e3 / n
    E._STATS["timed"], E._STATS["events"] = None, []
    eng.use_graph = graph
    return {k: round(v, 2) for k, v in sorted(out.items(), key=lambda kv: -kv[1])}


def bench_batched(args, rank, local_rank, world, n_runs=64):
    """BASELINE configs[4]: 64 independent 256x256 GMM-prior runs, dealt to the ranks and interleaved on CUDA
    streams on each GPU (strong scaling over runs, no collective)."""
    import torch
    import torch.distributed as dist

    import jolideco_b200 as J
    from jolideco_b200 import engine as E
    from jolideco_b200 import ops, synthetic

    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    ops.require_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))
    base = synthetic.make_workload("cfg5", seed=0)
    gmm = J.GaussianMixtureModel.from_numpy(*base["gmm_arrays"], meta=J.GaussianMixtureModelMeta(stride=4))
    rng = np.random.default_rng(5)
    jobs = []
    for r in range(n_runs):  # bootstrap: resampled counts of the same observation, own seed per run
        ds = {k: dict(v, counts=rng.poisson(np.clip(v["counts"], 0, None)).astype(np.float32))
              for k, v in base["datasets"].items()}
        prior = J.GMMPatchPrior(gmm=gmm, stride=4, generator=torch.Generator().manual_seed(r),
                                marginalize=args.marginalize, backend=args.backend)
        comp = J.SpatialFluxComponent.from_numpy(flux=base["flux_init"], upsampling_factor=base["f"], prior=prior)
        jobs.append(dict(datasets=ds, components=comp))
    t_setup0 = time.perf_counter()
    batch = J.BatchedRuns(jobs, n_epochs=args.steps + args.warmup, n_streams=16, rank=rank, world=world, device=device,
                          use_cuda_graph=not args.no_graph)
    t_setup = time.perf_counter() - t_setup0
    batch.run_epochs(args.warmup, trace=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = E._STATS["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for s_ in batch.streams:
        s_.wait_event(e0)
    batch.run_epochs(args.steps, trace=False)
    batch.join()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1), wall * 1e3 + t_setup * 1e3], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    total_iters = n_runs * args.steps  # one dataset per run: one MAP iteration per run and epoch
    if rank == 0:
        cfg = config_of(base, args, {"runs": n_runs, "runs_per_gpu": len(batch.runs), "streams": len(batch.streams),
                                     "parallelism": f"{n_runs} independent runs dealt to {world} GPU(s), interleaved on "
                                                    f"{len(batch.streams)} CUDA streams each",
                                     "l2_flush_between_iterations": False,
                                     "working_set": "64 runs x (flux, Adam state, dataset, spectra) stream through L2"})
        line = {"metric": METRIC, "value": total_iters / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "roofline": None,
                "cpu_baseline": None,
                "e2e": {"value": total_iters / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": None,
                        "d2h_bytes_per_step": None, "what": "BatchedRuns setup (H2D of every run) + timed epochs"},
                "gpu_launches": E._STATS["launches"] - launches0, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def rebuild_with_group(E, eng, pg, collective="nccl"):
    """Same engine with the dataset shard of this rank and a process group for the gradient all-reduce."""
    prior = None
    if eng.prior is not None:
        prior = dict(eng.prior)
    new = E.MapEngine(eng.theta, eng.datasets, prior=prior, mask=eng.mask, use_log_flux=eng.use_log_flux, beta=eng.beta,
                      lr=eng.lr, betas=(eng.b1, eng.b2), eps=eng.eps,
                      shift_table=eng.shift_table.cpu().numpy() if eng.shift_table is not None else None,
                      use_graph=eng.use_graph, process_group=pg, collective=collective)
    return new


def measure_roofline(E, eng, run_step, args, workload, flush):
    """Average duration of the dominant kernel (GMM prior forward, or the conv for a uniform prior),
    CUDA events around that launch inside real steps (eager mode), against its algorithmic work."""
    import torch

    cfg = workload["cfg"]
    fH = cfg["H"] * cfg["f"]
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    graph = eng.use_graph
    eng.use_graph = False
    extra = {}
    if eng.prior is not None:
        name = {1: "jd_gmm_prior_forward_tc", 2: "jd_gmm_prior_forward_tc16"}.get(eng.backend, "jd_gmm_prior_forward")
        if eng.backend == 1 and getattr(eng, "sk_ws", None) is not None:
            name = "jd_gmm_prior_forward_tc_sk"
        P = eng.P
        work = 2.0 * P * 64 * 64 * eng.packed.K  # useful flops, counted once
        bf16 = peaks.get("bf16_tflops_sustained")
        # split-TF32 runs on the TF32 pipe (half the bf16 rate); the split-FP16 kernel on the f16 pipe (= bf16 rate)
        div, what = (1.0, "sustained cuBLAS bf16") if eng.backend == 2 else (2.0, "1/2 x sustained cuBLAS bf16")
        peak, src = (bf16 / div, f"measured ({what}, MEASURED_PEAKS.json)") if bf16 else (
            1590.0 / div, f"fallback ({what} = 1.59 PFLOP/s)")
        bound, unit = "tensor", "TFLOP/s"
        # the 3-term split issues 3 products per useful one; the triangular trim keeps 320/512 of each
        issued = 3.0 * ((320.0 / 512.0) if eng.packed.upper_tri else 1.0)
        extra = {"issued_over_useful_flops": issued,
                 "ncu": ncu_reference(name, workload["name"])}
    else:
        name = "jd_conv_forward_direct"
        k = cfg["psf"] * cfg["f"]
        work = 2.0 * fH * fH * k * k
        peak, src = 148 * 128 * 2 * 1.965e9 / 1e12, "nominal FP32 pipe (no measured figure)"
        bound, unit = "tensor", "TFLOP/s"
    E._STATS["timed"], E._STATS["events"] = name, []
    n = max(5, min(args.steps, 20))
    for i in range(n):
        if flush is not None:
            flush.fill_(i & 0xFF)
        run_step(i)
    torch.cuda.synchronize()
    durs = [a.elapsed_time(b) for a, b in E._STATS["events"]]
    E._STATS["timed"], E._STATS["events"] = None, []
    eng.use_graph = graph
    avg_ms = float(np.mean(durs))
    achieved = work / (avg_ms * 1e-3) / 1e12
    out = {"bound": bound, "kernel": name, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
           "traffic": None, "avg_launch_ms": avg_ms, "launches_timed": len(durs), "algorithmic_work": work,
           "peak_source": src}
    if extra:
        out["issued_frac"] = achieved * extra["issued_over_useful_flops"] / peak
        out["issued_over_useful_flops"] = extra["issued_over_useful_flops"]
        ncu = extra["ncu"]
        if ncu:  # one committed `ncu --set full` capture of this kernel on this workload (profiles/)
            out["traffic"] = ncu.get("dram_bytes")
            out["ncu"] = ncu
    return out


def ncu_reference(kernel, workload):
    """DRAM traffic / pipe utilisation of the kernel from the committed ncu capture (profiles/ncu_reference.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_reference.json")) as fh:
            return json.load(fh).get(f"{kernel}:{workload}")
    except Exception:
        return None


def measure_e2e(J, workload, args, device, rank, joint=False):
    """MAPDeconvolver(n_epochs=E).run(datasets, components) with host (numpy) datasets: includes the
    host->device copies of every dataset, GMM constants and flux init, the per-epoch trace read-back
    and the final flux device->host copy."""
    import torch

    D = workload["cfg"]["D"]
    mode = "joint" if joint else "sequential"
    epochs = max(1, args.steps if joint else args.steps // D)
    seed = 0 if joint else rank  # joint: every rank must draw the same shifts
    # one short call first so that context / module load is not billed to the timed call
    deco, comps = build_run(J, workload, args, device, n_epochs=2, seed=seed, mode=mode)
    deco.run(datasets=workload["datasets"], components=comps)
    deco, comps = build_run(J, workload, args, device, n_epochs=epochs, seed=seed, mode=mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = deco.run(datasets=workload["datasets"], components=comps)
    flux = res.flux_upsampled_total
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    iters = epochs if joint else epochs * D
    h2d = sum(a.nbytes for d in workload["datasets"].values() for a in d.values()) + comps["flux"]._flux_upsampled.numel() * 4
    if workload["gmm_arrays"] is not None:
        K = workload["cfg"]["K"]
        h2d += K * (2 * 64 * 64 + 2 * 64 + 1) * 4
    d2h = epochs * 8 * (D + 1) + flux.nbytes
    return {"seconds": dt, "iters": iters, "h2d": h2d / iters, "d2h": d2h / iters,
            "what": f"MAPDeconvolver(n_epochs={epochs}, mode={mode!r}).run(numpy datasets) incl. setup, H2D of all inputs, "
                    "trace D2H and final flux D2H"}


if __name__ == "__main__":
    main()
