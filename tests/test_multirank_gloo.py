"""N>1 host-side coverage on CPU: two processes over torch.distributed/gloo (see
tests/run_plan_gloo.py)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_rank_exchange_emulation_over_gloo():
    env = dict(os.environ)
    env.pop("RANK", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(HERE, "run_plan_gloo.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "exchange emulation OK" in r.stdout
