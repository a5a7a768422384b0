"""The reference fork's own CUDA FEP kernels, compiled in place for sm_100a (oracle/ref_build/fork_cuda),
run beside ours on the B200: they must agree with our path and with the fp64 oracle to the precision
the fork's float kernels have, and their device time is printed next to ours (tests/fork_cuda_compare.py).

Runs in a subprocess (the fork's kernels are foreign code in this process otherwise)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libfepfork_cuda.so")


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libfepfork_cuda.so not built (make -C oracle fork_cuda)")
def test_fork_gpu_kernels_beside_ours():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fork_cuda_compare.py"), "small"], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    broken = [ln for ln in lines if "fork_error" in ln]
    if broken:  # the fork's kernels failed in our harness: the baseline is missing, the product was not involved
        pytest.xfail("the fork's CUDA kernels did not run in the harness: " + json.dumps(broken)[:1500])
    assert len(lines) == 2 and not any("unavailable" in ln for ln in lines), lines
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "fork_cuda_compare_small.jsonl"), "w") as fh:
            fh.write(r.stdout)
    except OSError:
        pass
    outside = []
    for ln in lines:
        print(json.dumps(ln))
        # ours against the oracle: the tolerances of BASELINE.json
        assert ln["ours_vs_oracle"]["force_rel_rms"] < 1e-5
        assert max(ln["ours_vs_oracle"][k] for k in ("Vc", "Vv", "dvdl", "foreign_energy")) < 1e-4
        # the fork's float kernels (atomics, no cut-off test on the soft-core radius, erff, other clamps;
        # SURVEY 2e): a sanity band that shows the harness feeds them the same problem, not a parity claim
        assert ln["ours_us"]["step"] > 0
        fork_ok = (ln["fork_vs_oracle"]["force_rel_rms"] < 1e-2
                   and max(ln["fork_vs_oracle"][k] for k in ("Vc", "Vv", "foreign_energy")) < 5e-2 and ln["fork_us"]["both"] > 0)
        if not fork_ok:
            outside.append(ln)
    if outside:  # the baseline harness, not the product: reported, and the tests that sort after this one still run under -x
        pytest.xfail("the fork's kernels in our harness are outside the sanity band: " + json.dumps(outside)[:1500])
