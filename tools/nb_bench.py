"""Times the non-perturbed cluster-pair kernel (include/fepb200_nb.h) on one of the synthetic systems, device-resident.
Prints one JSON line: cluster pairs, atom-pair evaluations per second, algorithmic-flop rate.  Used by bench.py (`nb`
sub-line) and by the ncu recipe (tools/gpu_call_*.sh).  Usage: python tools/nb_bench.py [C3] [--steps 20] [--energy]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gromacs-fep-gpu_b200", "python"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

# Algorithmic flop, the reference's own accounting: every atom pair of every LISTED cluster pair (nci_tot x 8 x 8,
# src/gromacs/nbnxm/pairlist.cpp:4231-4257 -> kerneldispatch.cpp:419) times the per-pair figure of
# src/gromacs/gmxlib/nrnb.cpp:90-95: "NxN RF Elec. + LJ" 38 [F] / 54 [V&F], "NxN Ewald Elec. + LJ" 66 [F] / 107 [V&F].
FLOP_PER_PAIR = {("ewald", False): 66, ("ewald", True): 107, ("rf", False): 38, ("rf", True): 54}


def run(name="C3", steps=20, warmup=3, energy=False, cpu_baseline=False, fork_gpu=False):
    import torch

    from fepb200 import params as P
    from fepb200 import synth_nb
    from fepb200.nb import NbContext, NB_Q_FROM_XQ
    from fepb200.synth import make_system

    t0 = time.time()
    pr = make_system(name)
    cs = synth_nb.build_cluster_system(pr, rlist=1.1)
    t_build = time.time() - t0
    nb = NbContext(0)
    nb.setup(cs, pr.params)
    d_xq = torch.from_numpy(cs.xq).cuda()
    d_f = torch.zeros((cs.natoms, 3), dtype=torch.float32, device="cuda")
    d_fs = torch.zeros(135, dtype=torch.float32, device="cuda")
    d_e = torch.zeros(2, dtype=torch.float64, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    flags = P.DO_FORCE | NB_Q_FROM_XQ | ((P.DO_POTENTIAL | P.DO_SHIFTFORCE) if energy else 0)
    ms = []
    for it in range(warmup + steps):
        flush.fill_(it & 0xFF)
        d_f.zero_()
        torch.cuda.synchronize()
        nb.launch_device(d_xq.data_ptr(), cs.shiftvec, flags, d_f.data_ptr(), d_fs.data_ptr(), d_e.data_ptr())
        nb.wait()
        if it >= warmup:
            ms.append(nb.last_kernel_ms())
    ms = np.array(ms)
    cp = nb.cluster_pairs
    evals = cp * 64
    elec = "ewald" if pr.params.elec_ewald else "rf"
    out = dict(workload=f"{name}: {cs.natoms} atoms in grid order ({int((cs.atom_index >= 0).sum())} real), "
                        f"{cs.sci.shape[0]} sci entries, {cs.cj.shape[0]} packed j entries, {cp} cluster pairs",
               kernel=f"fep_nb_kernel<{elec},{'energy' if energy else 'force'}>", ms=float(ms.mean()), ms_min=float(ms.min()),
               atom_pair_evals_per_s=evals / (ms.mean() * 1e-3), cluster_pairs=int(cp), list_build_s=t_build,
               pairs_in_cutoff=int(cs.pairs_in_cutoff), pairs_in_cutoff_per_s=cs.pairs_in_cutoff / (ms.mean() * 1e-3),
               flop_per_listed_atom_pair=FLOP_PER_PAIR[(elec, energy)],
               algorithmic_tflops=evals * FLOP_PER_PAIR[(elec, energy)] / (ms.mean() * 1e-3) / 1e12)
    nb.close()
    if cpu_baseline:
        # the reference's own kernel for these lists (kernel_gpu_ref.cpp compiled in place, mixed precision, plain C, one
        # thread -- it has no threaded or SIMD variant), timed on this host on the same list
        from oracle import nb_oracle

        if nb_oracle.have_ref("sp"):
            r = nb_oracle.run_ref(cs, pr.params, energy=energy, precision="sp", repeats=2)
            out["cpu_baseline"] = dict(kind="reference", cores=1, ms=r["seconds"] * 1e3,
                                       atom_pair_evals_per_s=evals / r["seconds"],
                                       sample=f"the whole {name} list, best of 2",
                                       what="nbnxn_kernel_gpu_ref (oracle/_ref/libnbref_sp.so)")
        else:
            out["cpu_baseline"] = dict(unavailable="oracle/_ref/libnbref_sp.so not built")
    if fork_gpu:
        # the reference's own CUDA cluster-pair kernel on this GPU and this list (compiled in place for sm_100a); its lists
        # are split on the host for load balance (nbnxm/pairlist.cpp: sci entries of at most a few packed j entries), so it
        # gets the list cut into the same chunk sizes our work items have
        from oracle import nb_oracle

        if nb_oracle.have_fork_cuda():
            cs_split = synth_nb.build_cluster_system(pr, rlist=1.1, max_cj_groups_per_sci=16)
            best = None
            for label, c in (("list as searched", cs), ("entries split at 16 packed j entries", cs_split)):
                r = nb_oracle.run_fork_cuda(c, pr.params, energy=energy, repeats=10)
                if best is None or r["ms"] < best["ms"]:
                    best = dict(ms=r["ms"], list=label, atom_pair_evals_per_s=evals / (r["ms"] * 1e-3))
            best["what"] = ("nbnxn_kernel_Elec%s_VdwLJ_%s_cuda of the reference (nbnxm_cuda_kernel.cuh compiled in place, "
                            "oracle/_ref/libnbfork_cuda.so), kernel only, warm caches, best of 10 -- ours above has the L2 "
                            "flushed before every launch" % ("Ew" if elec == "ewald" else "RF", "VF" if energy else "F"))
            out["reference_gpu_kernel"] = best
        else:
            out["reference_gpu_kernel"] = dict(unavailable="oracle/_ref/libnbfork_cuda.so not built")
    return out, cs, pr


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("name", nargs="?", default="C3")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--energy", action="store_true")
    ap.add_argument("--cpu-baseline", action="store_true")
    ap.add_argument("--fork-gpu", action="store_true")
    a = ap.parse_args()
    out, _, _ = run(a.name, a.steps, a.warmup, a.energy, a.cpu_baseline, a.fork_gpu)
    print(json.dumps(out))
