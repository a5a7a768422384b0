#!/usr/bin/env python
"""bench.py -- throughput of the FEP perturbed-pair path on synthetic solvated systems.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C5] [--impl ours|reference]

A "step" is one force step with free-energy output of the reference
(dispatchFreeEnergyKernel, src/gromacs/nbnxm/freeenergydispatch.cpp:147-308): the pass at the current
lambda (forces, shift forces, Vc/Vvdw per energy-group pair, dV/dlambda) followed by the L+1
energy-only foreign-lambda passes.  Metric (BASELINE.json): perturbed pair-interactions/s, one
pair-interaction = one (i,j) list entry evaluated at one lambda point, so a step is
P * (1 + (L+1)) pair-interactions -- the reference's own count of kernel work.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, device time (CUDA events, max over
ranks), L2 flushed between steps.  `e2e`: the same step through the public call with HOST buffers
(fepb200_compute: pinned staging, H2D of the touched coordinates, kernels, the result block written
into pinned host memory by the last kernel, scatter-add into the caller's force array; at N > 1
upload / kernels + reduction / download of the result block), wall clock.  `roofline`: the dominant kernel
(fep_foreign_kernel) against the FP32 pipe.  `cpu_baseline` / --impl reference: the reference's
own CPU SIMD kernel (oracle/_ref, compiled from the reference's sources) on this box's host cores.
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
for _p in (os.path.join(ROOT, "gromacs-fep-gpu_b200", "python"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "FEP perturbed pair-interactions/s"
UNIT = "pair-interactions/s"
FLOP_PER_PAIR, FLOP_PER_ENTRY = 150, 12  # the reference's own count, nb_free_energy.cpp:1181-1186
L2_FLUSH_BYTES = 512 << 20


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=float(d.get("hbm_gbs", 6650.0)), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)),
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


def _make_problem(name, args):
    from fepb200.synth import SPECS, make_system

    if args.n_foreign is None:
        return make_system(name)
    import dataclasses

    return make_system(dataclasses.replace(SPECS[name], n_foreign=args.n_foreign))


def _workload(problem):
    nb = problem.nblist
    passes = 1 + (problem.n_foreign + 1)
    return dict(pairs=nb.nrj, entries=nb.nri, passes=passes, units_per_step=nb.nrj * passes)


def _config(problem, name, world, extra=None):
    nb = problem.nblist
    cfg = dict(workload=f"{name}: {problem.natoms}-atom synthetic water box, {len(problem.perturbed)} perturbed atoms, "
                        f"{nb.nrj} perturbed pairs in {nb.nri} i-entries, {problem.n_foreign} foreign lambda",
               natoms=problem.natoms, pairs=nb.nrj, entries=nb.nri, n_foreign=problem.n_foreign,
               energy_group_pairs=problem.nenergrp_pairs,
               passes_per_step=1 + problem.n_foreign + 1,
               flags="FORCE|SHIFTFORCE|POTENTIAL|FOREIGNLAMBDA",
               parallelism=f"pair-list shards x{world}", l2="flushed between timed steps (512 MiB write)")
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------
# the reference's CPU kernel on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference(problem, flags, steps, warmup, budget_s=25.0):
    """Times oracle/_ref (the reference's nb_free_energy.cpp compiled in place, mixed precision,
    widest SIMD this host has, OpenMP over all host cores) or, if it did not travel, the C port."""
    from oracle import oracle

    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    nb = problem.nblist
    kind = "reference" if oracle.have_ref("sp") else "port"

    def run(prob):
        if kind == "reference":
            return oracle.run_ref(prob, flags, precision="sp", nthreads=cores, use_simd=True, repeats=1)
        return oracle.run_port(prob, flags, nthreads=cores, repeats=1)

    # bounded sample: all i-entries if one step fits the budget, else a leading range of entries
    import copy

    sample, frac = problem, 1.0
    t0 = time.perf_counter()
    r = run(sample)
    first = sum(r["seconds"])
    wall_first = time.perf_counter() - t0
    if wall_first * (steps + warmup) > budget_s and nb.nri > 64:
        frac = max(budget_s / (wall_first * (steps + warmup)), 0.02)
        e1 = max(64, int(nb.nri * frac))
        sample = copy.copy(problem)
        sample.nblist = nb.slice_entries(0, e1)
        frac = sample.nblist.nrj / nb.nrj
    passes = 1 + problem.n_foreign + 1
    for _ in range(warmup):
        run(sample)
    times = []
    for _ in range(steps):
        r = run(sample)
        times.append(sum(r["seconds"]))
    t = sum(times) / len(times)
    value = sample.nblist.nrj * passes / t
    simd = r.get("simd", "scalar C")
    return dict(value=value, unit=UNIT, cores=cores, kind=kind,
                sample=f"{sample.nblist.nrj} of {nb.nrj} pairs ({frac:.0%} of the i-entries' pairs) x {passes} passes "
                       f"per step, {steps} steps, {simd} mixed precision, OpenMP {cores} threads; first call {first:.3f}s",
                ms_per_step=t * 1e3)


def run_reference_gpu_arm(args, name):
    """Extra arm (not part of the driver contract): the reference fork's OWN CUDA FEP kernels, compiled in
    place for sm_100a (oracle/_ref/libfepfork_cuda.so), on the same problem and the same GPU.  Device
    time of its two kernels (CUDA events inside the harness, best of --steps, warm caches, no L2 flush)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fepb200 import params as P
    from oracle import oracle

    if not oracle.have_fork_cuda():
        print(json.dumps(dict(impl="reference-gpu", unavailable="oracle/_ref/libfepfork_cuda.so not built")), flush=True)
        return
    import dataclasses

    from fepb200.synth import SPECS, make_system

    spec = dataclasses.replace(SPECS[name], n_energy_groups=1)  # the fork has no energy groups
    if args.n_foreign is not None:
        spec = dataclasses.replace(spec, n_foreign=args.n_foreign)
    problem = make_system(spec)
    why = oracle.fork_cuda_unsupported(problem)
    if why:
        print(json.dumps(dict(impl="reference-gpu", unavailable="the fork's GPU FEP kernels do not cover: " + why)), flush=True)
        return
    flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    r = oracle.run_fork_cuda(problem, flags, repeats=max(args.steps, 1) + max(args.warmup, 3))
    t = sum(r["seconds"])
    wl = _workload(problem)
    line = dict(impl="reference-gpu", metric=METRIC, value=wl["units_per_step"] / t, unit=UNIT, n_gpus=1, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=t * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=_config(problem, name, 1, dict(parallelism="1 GPU", l2="warm (best of the repeats)",
                                                      energy_group_pairs=1)),
                kernel_ms=dict(current_lambda_kernel=r["seconds"][0] * 1e3, foreign_kernel=r["seconds"][1] * 1e3),
                note="k_calc_nb_fep + k_calc_nb_fep_foreign of the fork (nbnxm/cuda/nbnxm_cuda.cu:755-851), kernels only")
    print(json.dumps(line), flush=True)


def run_reference_arm(args, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fepb200 import params as P
    from fepb200.synth import make_system

    problem = _make_problem(name, args)
    flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    cpu = cpu_reference(problem, flags, args.steps, args.warmup, budget_s=120.0)
    line = dict(impl="reference", metric=METRIC, value=cpu["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=cpu.pop("ms_per_step"), higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="f32", data="synthetic", config=_config(problem, name, 1, dict(
                    parallelism=f"OpenMP x{cpu['cores']} host threads", l2="n/a (CPU)")),
                cpu_baseline=cpu, e2e=dict(value=cpu["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def _fork_gpu_beside(args, name):
    """`bench.py --impl reference-gpu` in a subprocess; a dict for the JSON line (never raises)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-gpu", "--config", name, "--steps", "20"]
    if args.n_foreign is not None:
        cmd += ["--n-foreign", str(args.n_foreign)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return dict(unavailable=f"exit code {r.returncode}: {(r.stderr or '').strip()[-200:]}")
        d = json.loads(lines[-1])
        if "unavailable" in d:
            return dict(unavailable=d["unavailable"])
        return dict(value=d["value"], unit=d["unit"], ms_per_step=d["ms_per_step"], kernel_ms=d["kernel_ms"],
                    what="the reference fork's own CUDA FEP kernels (nbnxm_fep_cuda_kernel.cuh + nbnxm_foreign_fep_cuda_kernel.cuh, "
                         "compiled in place for sm_100a, oracle/_ref/libfepfork_cuda.so) with the fork's launch configuration, "
                         "same problem with one energy group, kernels only, warm caches, best of 20 -- ours (`value`) is the "
                         "whole step incl. the reduction epilogue with L2 flushed between steps")
    except Exception as exc:  # noqa: BLE001 -- a reported baseline must not take the bench down
        return dict(unavailable=f"{type(exc).__name__}: {exc}"[:300])


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, name):
    import torch
    import torch.distributed as dist

    from fepb200 import params as P
    from fepb200.distributed import ShardedFep
    from fepb200.synth import make_system

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: whatever libraries print while we run (NCCL's version
    # banner at communicator creation) goes to stderr; the descriptor is restored for the result
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    problem = _make_problem(name, args)
    wl = _workload(problem)
    flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    sh = ShardedFep(problem, local, rank, world)
    ctx = sh.ctx
    lay = ctx.layout()
    x_host = torch.from_numpy(np.ascontiguousarray(problem.x)).pin_memory()
    x_np = x_host.numpy()
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device="cuda")

    # everything below is issued on the context's stream
    torch.cuda.set_stream(sh.stream)
    # clocks are sampled from before the warm-up to after the last timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- value: inputs resident, device time --------------------------------------------
    ctx.upload_x(x_np, problem.shiftvec)
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        sh.launch(flags)
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    n0 = ctx.launch_count()
    barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        sh.launch(flags)
        stops[i].record()
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - n0
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    dev_ms = max_over_ranks(dev_ms)
    ms_per_step = dev_ms / args.steps
    value = wl["units_per_step"] / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel: per-kernel CUDA events, same launches -----------
    ctx.set_profiling(True)
    kms = []
    for i in range(args.steps):
        flush.zero_()
        sh.launch(flags)
        torch.cuda.synchronize()
        kms.append(ctx.kernel_ms())
    ctx.set_profiling(False)
    k_pass, k_foreign, k_epi = (sum(k[j] for k in kms) / len(kms) for j in range(3))
    my_pairs, my_entries, my_atoms = int(lay.nrj), int(lay.nri), int(lay.ntouched)
    if sh.reduction == "fused":
        # every rank holds the full layout, evaluates its share of the pairs and owns a range of atoms
        # (its share is an equal number of trips, i.e. of pairs up to the padding of the trips)
        p0, p1, a0, a1 = ctx.peer_ranges()
        my_entries = int(round(my_entries / world))
        my_pairs, my_atoms = int(round(my_pairs / world)), a1 - a0
    points = problem.n_foreign + 1
    peaks = _peaks()
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    fp32_peak = sms * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
    # The dominant kernel is the one that evaluates the L+1 foreign-lambda passes.  On small lists the
    # library fuses the current-lambda pass into the same launch (then k_pass ~ 0 and the launch does
    # L+2 passes); algorithmic flop = the reference's own count per pass (nb_free_energy.cpp:1181-1186).
    fused = k_pass < 0.25 * k_foreign and k_pass < 0.004
    passes_in_launch = points + (1 if fused else 0)
    alg_flop = (FLOP_PER_PAIR * my_pairs + FLOP_PER_ENTRY * my_entries) * passes_in_launch
    achieved = alg_flop / (k_foreign * 1e-3) / 1e12 if k_foreign > 0 else 0.0
    # algorithmic bytes of the step (pair records + touched-atom data + result), for the HBM view
    alg_bytes = my_pairs * 16 + int(lay.ntouched) * (12 + 16 + 12)
    dominant = "fep_beutler_kernel" if problem.params.softcoreType == 0 and problem.params.alphaVdw != 0 \
        and problem.params.vdw_modifier != 3 else "fep_foreign_kernel"
    roofline = dict(bound="fp32", kernel=dominant + (" (current-lambda pass fused in)" if fused else " (foreign-lambda passes)"),
                    achieved=achieved, peak=fp32_peak, unit="TFLOP/s",
                    frac=achieved / fp32_peak,
                    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from the
                    # ncu --set full capture committed as profiles/r01_ncu_full_final_raw.csv (C5, 1 GPU)
                    traffic=17102080 if (name == "C5" and world == 1) else None,
                    note="achieved = ALGORITHMIC flop (150/pair + 12/i-entry per lambda pass, the reference's count) / kernel time; "
                         "the kernel hoists everything lambda-independent out of the pass loop, so the executed "
                         "FP32 instruction count per pass is far below 150 and frac can exceed 1; "
                         "see profiles/ for executed-instruction pipe utilisation",
                    peak_source=f"{sms} SMs x 128 lanes x 2 x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} sm_max_mhz)",
                    algorithmic_flop_per_launch=alg_flop, passes_in_launch=passes_in_launch,
                    kernel_ms=dict(pass_kernel=k_pass, foreign_kernel=k_foreign, epilogue_kernel=k_epi),
                    pair_points_per_s=my_pairs * passes_in_launch / (k_foreign * 1e-3) if k_foreign > 0 else 0.0,
                    # the whole step against the same roof: all passes the reference would run
                    # (1 + (L+1)) at its own flop count, over the step's device time (`value`'s clock)
                    step=dict(algorithmic_flop=(FLOP_PER_PAIR * wl["pairs"] + FLOP_PER_ENTRY * wl["entries"]) * wl["passes"],
                              achieved=(FLOP_PER_PAIR * wl["pairs"] + FLOP_PER_ENTRY * wl["entries"]) * wl["passes"]
                              / (ms_per_step * 1e-3) / 1e12,
                              frac=(FLOP_PER_PAIR * wl["pairs"] + FLOP_PER_ENTRY * wl["entries"]) * wl["passes"]
                              / (ms_per_step * 1e-3) / 1e12 / (fp32_peak * world)),
                    # executed work of the same launch: warp instructions from the ncu capture of this kernel on
                    # this workload (profiles/README.md: 11.4 M for the 21-point foreign launch on C5) over the live
                    # kernel time, against the issue rate of the chip (4 schedulers x 1 warp instruction per clock per SM)
                    issue=(dict(warp_instructions=11.4e6, source="profiles/r01_ncu_full_final_raw.csv (smsp__inst_executed.sum)",
                                peak_per_s=sms * 4 * peaks["sm_max_mhz"] * 1e6,
                                frac=11.4e6 / (k_foreign * 1e-3) / (sms * 4 * peaks["sm_max_mhz"] * 1e6) if k_foreign > 0 else 0.0)
                           if (name == "C5" and world == 1 and not fused and args.n_foreign is None) else None),
                    hbm=dict(algorithmic_bytes_per_step=alg_bytes,
                             achieved_gbs=alg_bytes / ((k_pass + k_foreign + k_epi) * 1e-3) / 1e9,
                             peak_gbs=peaks["hbm_gbs"]))

    # ---- e2e: host buffers through the public call ------------------------------------------
    out = ctx.new_outputs()
    for _ in range(3):
        sh.step(x_np, problem.shiftvec, flags | P.CLEAR_OUTPUTS, out)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lap = [0.0, 0.0, 0.0] if os.environ.get("FEPB200_E2E_PHASES") else None
    for _ in range(args.steps):
        flush.zero_()
        if lap is None:
            sh.step(x_np, problem.shiftvec, flags | P.CLEAR_OUTPUTS, out)
        else:  # the same three calls step() makes, with the host clock between them (diagnosis)
            ta = time.perf_counter()
            ctx.upload_x(x_np, problem.shiftvec)
            tb = time.perf_counter()
            sh.launch(flags | P.CLEAR_OUTPUTS)
            tc = time.perf_counter()
            ctx.download(flags | P.CLEAR_OUTPUTS, out)
            td = time.perf_counter()
            lap = [lap[0] + tb - ta, lap[1] + tc - tb, lap[2] + td - tc]
    if lap is not None:
        print(f"[e2e phases] rank {rank}: upload_x {lap[0] / args.steps * 1e6:.1f} us, launch {lap[1] / args.steps * 1e6:.1f} us, "
              f"download {lap[2] / args.steps * 1e6:.1f} us", file=sys.stderr, flush=True)
    torch.cuda.synchronize()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    # the L2 flush is not part of the step: measure and subtract its device time
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        flush.zero_()
    f1.record()
    torch.cuda.synchronize()
    e2e_s = max(e2e_s - f0.elapsed_time(f1) * 1e-3, 1e-9)
    e2e_value = wl["units_per_step"] / (e2e_s / args.steps)
    h2d = 816 + 12 * int(lay.ntouched)  # DynHead (45 shift vectors + current-lambda block) + packed xyz per touched atom
    d2h = (3 * my_atoms + 135) * 4 + int(lay.f64_words) * 8

    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference(problem, flags, 3, 1)
            cpu.pop("ms_per_step", None)
        # the fork's own CUDA FEP kernels on this GPU and this problem, beside ours (reported, like
        # cpu_baseline).  After all of our timing, in a process of its own: foreign kernels stay out of
        # this one, and whatever happens there cannot touch the numbers above.
        fork_gpu = None
        if world == 1 and not args.no_fork_gpu:
            fork_gpu = _fork_gpu_beside(args, name)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                    data="synthetic", config=_config(problem, name, world, dict(
                        reduction=sh.reduction if world > 1 else "none",
                        outputs="forces reduce-scattered by atom range, scalars on every rank" if sh.reduction == "fused"
                        else "full result on every rank")),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             ms_per_step=e2e_s / args.steps * 1e3),
                    gpu_launches=int(launches) * world, clocks=clocks, roofline=roofline, cpu_baseline=cpu,
                    fork_gpu_baseline=fork_gpu,
                    wall_ms_per_step_incl_flush=t_wall / args.steps * 1e3, device=ctx.describe())
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    sh.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="C5", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fork-gpu", action="store_true", help="skip the fork's CUDA kernels beside ours (fork_gpu_baseline)")
    ap.add_argument("--n-foreign", type=int, default=None, help="override the number of foreign lambda points")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference-gpu":
        return run_reference_gpu_arm(args, args.config)
    if args.impl == "ours" and world != args.gpus:
        if world == 1 and args.gpus > 1:
            # re-launch ourselves one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.impl == "reference":
        run_reference_arm(args, args.config)
    else:
        if world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
            # torchrun defaults every rank to ONE OpenMP thread; the library's host-side gather /
            # scatter of the touched atoms (fepb200_compute) is threaded: give each rank its share
            # of the host cores (must happen before libgomp is loaded)
            try:
                cores = len(os.sched_getaffinity(0))
            except (AttributeError, OSError):
                cores = os.cpu_count() or 1
            os.environ["OMP_NUM_THREADS"] = str(max(1, min(16, cores // world)))
        run_ours(args, args.config)


if __name__ == "__main__":
    main()
