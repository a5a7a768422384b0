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
ranks), L2 flushed between steps; at N > 1 the line also carries the same measurement with the ranks' streams aligned
on the device after the flush and before the start event of every timed step (`run.ms_per_step_aligned`; --align makes
it the headline).  `e2e`: the same step through the public call with HOST buffers
(fepb200_compute: pinned staging, H2D of the touched coordinates, kernels, the result block written
into pinned host memory by the last kernel, scatter-add into the caller's force array; at N > 1
upload / kernels + reduction / download of the result block), wall clock.  `roofline`: the kernel that took
longest in THIS run (per-kernel CUDA events) against the FP32 pipe by the reference's own flop count, every kernel
of the step beside it, each with the executed-instruction view from the committed ncu capture of the same sources
(profiles/r02_ncu_counters.json, tools/ncu_counters.py; ignored when the kernel sources have changed since).
`every_step`: the force-only step (L = 0), what a production run pays on every MD step.  `configs`: C1-C4 on the
same GPU (N = 1).  `cpu_baseline` / --impl reference: the reference's own CPU SIMD kernel (oracle/_ref, compiled
from the reference's sources) on this box's host cores.
"""
from __future__ import annotations

import argparse
import hashlib
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


def _config(problem, name, extra=None):
    """What is computed: the same for every arm and every N (how it is run is under `run`)."""
    nb = problem.nblist
    cfg = dict(workload=f"{name}: {problem.natoms}-atom synthetic water box, {len(problem.perturbed)} perturbed atoms, "
                        f"{nb.nrj} perturbed pairs in {nb.nri} i-entries, {problem.n_foreign} foreign lambda",
               natoms=problem.natoms, pairs=nb.nrj, entries=nb.nri, n_foreign=problem.n_foreign,
               energy_group_pairs=problem.nenergrp_pairs,
               passes_per_step=1 + problem.n_foreign + 1,
               flags="FORCE|SHIFTFORCE|POTENTIAL|FOREIGNLAMBDA")
    if extra:
        cfg.update(extra)
    return cfg


def _source_hash():
    """sha256 over the kernel sources: ties profiles/r02_ncu_counters.json to the code it was captured from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "gromacs-fep-gpu_b200", "csrc")
    for fn in sorted(os.listdir(d)):
        if fn.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(d, fn), "rb") as fh:
                h.update(fn.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def _ncu_counters(name):
    """Per-kernel counters of the committed ncu capture for this workload, or (None, why)."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_counters.json")
    if not os.path.exists(path):
        return None, "profiles/r02_ncu_counters.json not present"
    with open(path) as fh:
        d = json.load(fh)
    if d.get("source_hash") != _source_hash():
        return None, f"profiles/r02_ncu_counters.json was captured from other kernel sources ({d.get('source_hash')} != {_source_hash()})"
    if name not in d.get("workloads", {}):
        return None, f"no ncu capture of {name} in profiles/r02_ncu_counters.json"
    return d["workloads"][name], d.get("captured_with", "")


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
                config=_config(problem, name, dict(energy_group_pairs=1)),
                run=dict(parallelism="1 GPU", l2="warm (best of the repeats)"),
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
                vs_baseline=None, dtype="f32", data="synthetic", config=_config(problem, name),
                run=dict(parallelism=f"OpenMP x{cpu['cores']} host threads", l2="n/a (CPU)"),
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


def _nb_beside(args):
    """the non-perturbed cluster-pair kernel (include/fepb200_nb.h) timed on C3's atoms in the reference's
    GPU cluster layout, device-resident, L2 flushed between launches; tools/nb_bench.py in a subprocess (never raises).
    Roofline by the reference's own flop accounting for these kernels (nrnb.cpp:90-95 x listed atom pairs)."""
    cmd = [sys.executable, os.path.join(ROOT, "tools", "nb_bench.py"), "C3", "--steps", "20"]
    try:
        out = {}
        for key, extra in (("force", ["--cpu-baseline", "--fork-gpu"]), ("force_energy_virial", ["--energy", "--fork-gpu"])):
            r = subprocess.run(cmd + extra, capture_output=True, text=True, timeout=600)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode != 0 or not lines:
                return dict(unavailable=f"exit code {r.returncode}: {(r.stderr or '').strip()[-200:]}")
            d = json.loads(lines[-1])
            peak = 148 * 128 * 2 * _peaks()["sm_max_mhz"] * 1e-6  # TFLOP/s, like the FEP kernels' roofline
            d["frac_of_fp32_peak"] = d["algorithmic_tflops"] / peak
            out[key] = d
        return out
    except Exception as exc:  # noqa: BLE001 -- a reported side line must not take the bench down
        return dict(unavailable=f"{type(exc).__name__}: {exc}"[:300])


def _side_config(name, device, flush, steps):
    """One more BASELINE.json configuration on this GPU: device time per step (L2 flushed between steps, inputs resident),
    the same through fepb200_compute with host buffers, per-kernel times."""
    import torch

    from fepb200 import params as P
    from fepb200.lib import FepContext
    from fepb200.synth import make_system

    prob = make_system(name)
    wl = _workload(prob)
    flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    with FepContext(device) as ctx:
        ctx.set_problem(prob)
        ctx.upload_x(prob.x, prob.shiftvec)
        for _ in range(3):
            ctx.launch(flags)
        ctx.wait()
        ms = []
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            ctx.launch(flags)
            ctx.wait()
            ms.append(ctx.last_launch_ms())
        fo = []
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            ctx.launch(P.DO_FORCE)
            ctx.wait()
            fo.append(ctx.last_launch_ms())
        ctx.set_profiling(True)
        rows = []
        for _ in range(min(steps, 10)):
            flush.zero_()
            torch.cuda.synchronize()
            ctx.launch(flags)
            ctx.wait()
            rows.append(ctx.kernel_ms())
        ctx.set_profiling(False)
        out = ctx.new_outputs()
        x = np.ascontiguousarray(prob.x)
        for _ in range(3):
            ctx.compute(x, prob.shiftvec, flags | P.CLEAR_OUTPUTS, out)
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.compute(x, prob.shiftvec, flags | P.CLEAR_OUTPUTS, out)
        e2e = (time.perf_counter() - t0) / steps
    t = sum(ms) / len(ms)
    return dict(config=_config(prob, name), ms_per_step=t, value=wl["units_per_step"] / (t * 1e-3), unit=UNIT,
                force_only_ms_per_step=sum(fo) / len(fo),
                kernel_ms=dict(zip(("pass_kernel", "foreign_kernel", "epilogue_kernel"),
                                   (sum(r[j] for r in rows) / len(rows) for j in range(3)))),
                e2e=dict(ms_per_step=e2e * 1e3, value=wl["units_per_step"] / e2e, note="host buffers, L2 not flushed"),
                steps=steps)


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
    # N > 1, align = True: the ranks' streams are aligned ON THE DEVICE between the L2 flush and the start event of every
    # timed step (ShardedFep.align: the symmetric-memory barrier kernel, outside the timed region), so that the ranks'
    # different flush times are not waited for inside the step's own cross-GPU barrier.  The headline is measured
    # WITHOUT (ranks run free, as in round 1 and 2's records) unless --align is given; the line carries both
    # (run.ms_per_step_aligned / _unaligned).  On 2 B200 the two agree to 1 us (profiles/r02_multi_gpu_bench_lines.json).
    def timed_steps(fl, steps, align):
        ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        torch.cuda.synchronize()
        t_w0 = time.perf_counter()
        for i in range(steps):
            flush.zero_()
            if align:
                sh.align()
            ev0[i].record()
            sh.launch(fl)
            ev1[i].record()
        torch.cuda.synchronize()
        barrier()
        t_w = time.perf_counter() - t_w0
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))) / steps, t_w

    can_align = world > 1 and sh.align()
    aligned = can_align and args.align
    torch.cuda.synchronize()
    n0 = ctx.launch_count()
    ms_per_step, t_wall = timed_steps(flags, args.steps, aligned)
    launches = ctx.launch_count() - n0
    ms_other = timed_steps(flags, args.steps, not aligned)[0] if can_align else None
    value = wl["units_per_step"] / (ms_per_step * 1e-3)

    # ---- per-kernel times of the same launches: CUDA events of the library around every kernel ------
    def kernel_times(fl, steps):
        ctx.set_profiling(True)
        rows = []
        for _ in range(steps):
            flush.zero_()
            sh.launch(fl)
            torch.cuda.synchronize()
            rows.append(ctx.kernel_ms())
        ctx.set_profiling(False)
        return [sum(k[j] for k in rows) / len(rows) for j in range(3)]

    def device_ms(fl, steps):
        for _ in range(3):
            flush.zero_()
            sh.launch(fl)
        return timed_steps(fl, steps, aligned)[0]

    k_pass, k_foreign, k_epi = kernel_times(flags, args.steps)
    my_pairs, my_entries, my_atoms = int(lay.nrj), int(lay.nri), int(lay.ntouched)
    if sh.reduction == "fused":
        # every rank holds the full layout, evaluates its share of the pairs and owns a range of atoms
        # (its share is an equal number of trips, i.e. of pairs up to the padding of the trips)
        p0, p1, a0, a1 = ctx.peer_ranges()
        my_entries = int(round(my_entries / world))
        my_pairs, my_atoms = int(round(my_pairs / world)), a1 - a0
    elif sh.reduction in ("p2p", "p2p-push"):
        p0, p1, a0, a1 = ctx.peer_ranges()  # the atoms this rank owns after the reduce-scatter
        my_atoms = a1 - a0
    points = problem.n_foreign + 1
    peaks = _peaks()
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    fp32_peak = sms * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12
    issue_peak = sms * 4 * peaks["sm_max_mhz"] * 1e6  # warp instructions per second: 4 schedulers per SM, 1 per clock
    # On small lists the library fuses the current-lambda pass into the launch of the foreign passes (then the
    # "pass" events bracket nothing and the launch does L+2 passes).
    fused = k_pass < 0.25 * k_foreign and k_pass < 0.004
    flop_pass = FLOP_PER_PAIR * my_pairs + FLOP_PER_ENTRY * my_entries  # the reference's count, one lambda pass
    counters, counters_note = (None, "multi-GPU run") if world > 1 or args.n_foreign is not None else _ncu_counters(name)

    def kernel_view(key, label, ms, passes):
        """One kernel of the step: algorithmic flop (the reference's count x passes) over the live time, and -- when the
        committed ncu capture belongs to these sources -- executed warp instructions over the live time."""
        v = dict(kernel=label, ms=ms, passes=passes)
        if passes > 0 and ms > 0:
            v["algorithmic_flop"] = flop_pass * passes
            v["achieved_tflops"] = flop_pass * passes / (ms * 1e-3) / 1e12
            v["flop_frac_of_fp32_peak"] = v["achieved_tflops"] / fp32_peak
        c = (counters or {}).get(key)
        if c and ms > 0:
            v["ncu"] = c
            v["issue_frac_live"] = c["inst_executed"] / (ms * 1e-3) / issue_peak
        return v

    views = []
    if not fused:
        views.append(kernel_view("pass", "current-lambda force pass (fep_beutler_kernel<C=0,FORCE> / fep_pass_kernel)", k_pass, 1))
    views.append(kernel_view("foreign", "foreign-lambda passes" + (" + current-lambda pass, one launch" if fused else ""),
                             k_foreign, points + (1 if fused else 0)))
    views.append(kernel_view("epilogue", "fep_epilogue_kernel (per-atom force sums, shift forces, scalars)", k_epi, 0))
    dominant = max(views, key=lambda v: v["ms"])
    # epilogue: a streaming sum, bounded by bytes -- contributions read + forces written
    epi_bytes = 16 * (my_pairs + my_pairs // 32 + 1) + 12 * my_atoms
    views[-1]["algorithmic_bytes"] = epi_bytes
    views[-1]["achieved_gbs"] = epi_bytes / (k_epi * 1e-3) / 1e9 if k_epi > 0 else 0.0
    views[-1]["frac_of_hbm_peak"] = views[-1]["achieved_gbs"] / peaks["hbm_gbs"]
    step_flop = (FLOP_PER_PAIR * wl["pairs"] + FLOP_PER_ENTRY * wl["entries"]) * wl["passes"]
    if "achieved_tflops" in dominant:
        d_ach, d_frac = dominant["achieved_tflops"], dominant["flop_frac_of_fp32_peak"]
    else:
        d_ach, d_frac = views[-1]["achieved_gbs"], views[-1]["frac_of_hbm_peak"]
    frac_kind = "algorithmic flop / FP32 peak"
    if d_frac > 1.0:
        # The foreign-lambda launch hoists everything lambda-independent out of its point loop, so it EXECUTES far
        # fewer than 150 flop per pair-point and the flop count says nothing about hardware use: report what the
        # hardware did instead (issue slots used while the kernel ran), and keep the flop figure beside it.
        if "issue_frac_live" in dominant:
            d_frac, frac_kind = dominant["issue_frac_live"], "executed warp instructions (ncu) / live time / issue peak"
        else:
            d_frac, frac_kind = None, "algorithmic flop exceeds the FP32 peak (hoisted lambda loop); no ncu counters for these sources"
    roofline = dict(bound="fp32" if "achieved_tflops" in dominant else "hbm", kernel=dominant["kernel"],
                    achieved=d_ach, peak=fp32_peak if "achieved_tflops" in dominant else peaks["hbm_gbs"],
                    unit="TFLOP/s" if "achieved_tflops" in dominant else "GB/s", frac=d_frac, frac_is=frac_kind,
                    algorithmic_frac=dominant.get("flop_frac_of_fp32_peak"),
                    traffic=(dominant.get("ncu") or {}).get("dram_bytes"),
                    kernels=views, counters_note=counters_note,
                    note="achieved = ALGORITHMIC flop (150 per pair + 12 per i-entry per lambda pass, the reference's own count, "
                         "nb_free_energy.cpp:1181-1186) / live kernel time (CUDA events of this run); `kernel` is the kernel that took "
                         "longest in this run",
                    peak_source=f"{sms} SMs x 128 lanes x 2 x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} sm_max_mhz; "
                                "MEASURED_PEAKS.json has no FP32 entry)",
                    step=dict(algorithmic_flop=step_flop, achieved=step_flop / (ms_per_step * 1e-3) / 1e12,
                              frac=step_flop / (ms_per_step * 1e-3) / 1e12 / (fp32_peak * world)),
                    hbm=dict(algorithmic_bytes_per_step=my_pairs * 22 + int(lay.ntouched) * (12 + 12) + epi_bytes,
                             peak_gbs=peaks["hbm_gbs"]))

    # ---- the every-step path: forces only, no foreign lambda (L = 0) ---------------------------------------------
    f_only = P.DO_FORCE
    fo_ms = device_ms(f_only, args.steps)
    fo_k = kernel_times(f_only, min(args.steps, 20))
    every_step = dict(flags="FORCE", ms_per_step=fo_ms, pairs_per_s=wl["pairs"] / (fo_ms * 1e-3),
                      kernel_ms=dict(pass_kernel=fo_k[0], epilogue_kernel=fo_k[2]),
                      achieved_tflops=(FLOP_PER_PAIR * wl["pairs"] + FLOP_PER_ENTRY * wl["entries"]) / (fo_ms * 1e-3) / 1e12,
                      frac_of_fp32_peak=(FLOP_PER_PAIR * wl["pairs"] + FLOP_PER_ENTRY * wl["entries"]) / (fo_ms * 1e-3) / 1e12
                      / (fp32_peak * world),
                      note="what a production FEP run pays on every MD step (foreign energies only every nstdhdl steps): device "
                           "time of pass + epilogue, L2 flushed between steps")
    fv_ms = device_ms(P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL, min(args.steps, 50))
    every_step["with_energies_and_virial_ms"] = fv_ms

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
        # C1-C4 on the same GPU (N = 1 only): their own ms_per_step, value and e2e
        side = None
        if world == 1 and not args.no_side_configs and args.n_foreign is None:
            side = {}
            for other in ("C1", "C2", "C3", "C4", "C5"):
                if other != name:
                    try:
                        side[other] = _side_config(other, local, flush, min(args.steps, 30))
                    except Exception as exc:  # noqa: BLE001 -- a side line must not take the bench down
                        side[other] = dict(unavailable=f"{type(exc).__name__}: {exc}"[:300])
        nb_line = None
        if world == 1 and not args.no_side_configs and args.n_foreign is None:
            nb_line = _nb_beside(args)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                    data="synthetic", config=_config(problem, name),
                    run=dict(parallelism=f"pair-list shards x{world}", l2="flushed between timed steps (512 MiB write)",
                             reduction=sh.reduction if world > 1 else "none",
                             rank_alignment=("device-side barrier over the ranks between the L2 flush and the start event of "
                                             "every timed step (outside the timed region)") if aligned else "none",
                             ms_per_step_aligned=ms_per_step if aligned else ms_other,
                             ms_per_step_unaligned=ms_other if aligned else (ms_per_step if world > 1 else None),
                             outputs="forces reduce-scattered by atom range, scalars on every rank"
                             if sh.reduction in ("fused", "p2p", "p2p-push") else "full result on every rank"),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             ms_per_step=e2e_s / args.steps * 1e3,
                             note="wall clock of the public call with host buffers; the device time of the L2 flush between "
                                  "steps (measured separately) is subtracted"),
                    gpu_launches=int(launches) * world, clocks=clocks, roofline=roofline, every_step=every_step,
                    kernel_ms=dict(pass_kernel=k_pass, foreign_kernel=k_foreign, epilogue_kernel=k_epi),
                    configs=side, cluster_pair_kernel=nb_line, cpu_baseline=cpu, fork_gpu_baseline=fork_gpu,
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
    ap.add_argument("--no-side-configs", action="store_true", help="skip the C1-C4 sub-lines (configs)")
    ap.add_argument("--n-foreign", type=int, default=None, help="override the number of foreign lambda points")
    ap.add_argument("--align", action="store_true",
                    help="N > 1: headline measured with the ranks aligned on the device before each timed step (see run.rank_alignment)")
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
