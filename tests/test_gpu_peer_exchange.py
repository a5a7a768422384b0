"""The fused multi-GPU exchange (fepb200_set_peer_exchange) exercised on ONE device: N contexts of one
process play the ranks, their exchange buffers are plain device allocations, so the cross-rank
barrier of the epilogue and its "peer" reads run exactly the code they run over NVLink.  Forces are owned by atom range: the sum of the ranks' force arrays must be BIT-IDENTICAL to
the single-context result (same scatter slots, same summation order), as must the shift forces and
Vc/Vv (same reduction jobs); dV/dlambda and the foreign energies are sums of per-CTA partials whose
tiling differs with the split, so they agree to rounding; everything must match the oracle at the
tolerances of north_star.  The real multi-GPU run of the same path is in test_multi_gpu.py."""
import numpy as np
import pytest

from fepb200 import params as P

pytestmark = pytest.mark.gpu

ALL = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA


def _problem(kind):
    from fepb200.synth import make_system, scaled_spec

    if kind == "beutler_ewald":
        return make_system(scaled_spec("C2", 3.2, 1, 20, n_foreign=4))
    if kind == "beutler_rf_groups":
        return make_system(scaled_spec("C4", 4.2, 2, 25, n_foreign=5))
    if kind == "gapsys":
        return make_system(scaled_spec("C3", 3.6, 2, 20, n_foreign=3))
    raise ValueError(kind)


def _run_ranks(prob, nranks, flags, steps=1):
    """All ranks' outputs of `steps` lockstep steps (the last step's), plus their ranges."""
    import torch

    from fepb200.lib import FepContext

    ctxs = [FepContext(0) for _ in range(nranks)]
    try:
        for c in ctxs:
            c.set_problem(prob)
        nbytes = (ctxs[0].exchange_bytes(nranks) + 4095) // 4096 * 4096
        assert nbytes > 0
        bufs = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(nranks)]
        torch.cuda.synchronize()
        ptrs = [int(b.data_ptr()) for b in bufs]
        for r, c in enumerate(ctxs):
            c.set_peer_exchange(nranks, r, ptrs, nbytes)
        ranges = [c.peer_ranges() for c in ctxs]
        outs = None
        for _ in range(steps):
            for c in ctxs:
                c.upload_x(prob.x, prob.shiftvec)
            for c in ctxs:  # every rank's kernels are queued before anybody waits for a result
                c.launch(flags)
            outs = [c.download(flags) for c in ctxs]
        return outs, ranges, [c.touched_atoms() for c in ctxs]
    finally:
        for c in ctxs:
            c.close()


def _single(prob, flags):
    from fepb200.lib import FepContext

    with FepContext(0) as c:
        c.set_problem(prob)
        return c.compute(prob.x, prob.shiftvec, flags)


@pytest.mark.parametrize("kind", ["beutler_ewald", "beutler_rf_groups", "gapsys"])
@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_ranks_on_one_device_reproduce_the_single_context_result(kind, nranks):
    from oracle import oracle

    prob = _problem(kind)
    one = _single(prob, ALL)
    outs, ranges, touched = _run_ranks(prob, nranks, ALL, steps=3)
    # the ranges are partitions of the trips (the mirror of the device layout tells how many) and of the touched atoms
    from fepb200.shard import trip_layout

    assert ranges[0][0] == 0 and ranges[-1][1] == trip_layout(prob.nblist, prob.nenergrp_pairs)["n_trips"]
    assert ranges[0][2] == 0 and ranges[-1][3] == len(touched[0])
    for a, b in zip(ranges, ranges[1:]):
        assert a[1] == b[0] and a[3] == b[2]
    # forces: every rank holds exactly the atoms it owns, the sum over ranks is the single-GPU result
    f_sum = np.zeros_like(one["f"])
    for out, (_, _, a0, a1), t in zip(outs, ranges, touched):
        mask = np.zeros(prob.natoms, bool)
        mask[t[a0:a1]] = True
        assert not np.any(out["f"][~mask])
        f_sum += out["f"]
    assert np.array_equal(f_sum, one["f"])
    want = oracle.run_best(prob, ALL)
    rms = np.sqrt(np.mean((f_sum - want["f"]) ** 2) / np.mean(want["f"] ** 2))
    assert rms < 1e-5
    # sums over sorted slots (shift forces; Vc/Vv per energy-group pair when there are several) are bit-identical to the
    # single-GPU result; sums over per-CTA partials (dV/dlambda, foreign energies, Vc/Vv of a single energy-group
    # pair) follow the launch geometry, which differs between one GPU and a share of the list
    exact = ("fshift", "Vc", "Vv") if prob.nenergrp_pairs > 1 else ("fshift",)
    for out in outs:
        for k in exact:
            assert np.array_equal(out[k], one[k]), k
        for k in ("dvdl", "foreign_energy", "foreign_dvdl") + (() if prob.nenergrp_pairs > 1 else ("Vc", "Vv")):
            scale = np.maximum(np.abs(want[k]), 1e-2 * np.max(np.abs(want[k])))
            assert np.all(np.abs(out[k] - want[k]) <= 1e-4 * scale), k
            assert np.allclose(out[k], one[k], rtol=1e-6, atol=1e-6 * np.max(np.abs(one[k]))), k
            assert np.array_equal(out[k], outs[0][k]), k  # every rank sums the same partials in the same order


def test_flag_subsets_and_switching_the_exchange_off():
    from fepb200.lib import FepContext

    prob = _problem("beutler_ewald")
    for flags in (P.DO_FORCE, P.DO_POTENTIAL, P.DO_FORCE | P.DO_POTENTIAL, P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA):
        one = _single(prob, flags)
        outs, _, _ = _run_ranks(prob, 2, flags)
        assert np.array_equal(outs[0]["f"] + outs[1]["f"], one["f"])
        assert np.array_equal(outs[1]["fshift"], one["fshift"]), flags
        for k in ("Vc", "Vv", "dvdl"):
            assert np.allclose(outs[1][k], one[k], rtol=1e-6, atol=1e-6 * np.max(np.abs(one[k])) + 1e-30), (flags, k)
    # nranks = 1 restores the plain single-GPU path on the same context
    import torch

    with FepContext(0) as c:
        c.set_problem(prob)
        ref = c.compute(prob.x, prob.shiftvec, ALL)
        nbytes = c.exchange_bytes(1)
        buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        c.set_peer_exchange(1, 0, [int(buf.data_ptr())], nbytes)
        again = c.compute(prob.x, prob.shiftvec, ALL)
        for k in ref:
            assert np.array_equal(ref[k], again[k]), k


def test_bad_arguments_are_refused():
    from fepb200.lib import FepContext, FepError

    prob = _problem("beutler_ewald")
    with FepContext(0) as c:
        with pytest.raises(FepError):
            c.set_peer_exchange(2, 0, [256, 512], 1 << 20)  # no list yet
        c.set_problem(prob, rank=0, nranks=2)
        with pytest.raises(FepError):
            c.set_peer_exchange(2, 0, [256, 512], 1 << 20)  # holds a shard, not the full list
        c.set_problem(prob)
        with pytest.raises(FepError):
            c.set_peer_exchange(9, 0, [256] * 9, 1 << 20)
        with pytest.raises(FepError):
            c.set_peer_exchange(2, 2, [256, 512], 1 << 20)
        with pytest.raises(FepError):
            c.set_peer_exchange(2, 0, [256, 513], 1 << 30)  # misaligned
        with pytest.raises(FepError):
            c.set_peer_exchange(2, 0, [256, 512], 1024)  # too small
        # a refused set-up leaves the context usable on its own
        out = c.compute(prob.x, prob.shiftvec, ALL)
        assert np.isfinite(out["f"]).all()
