"""The product's N > 1 update, driver-visible (-m gpu): spawns 2 ranks with torch.distributed.run and runs
tests/dist_worker.py -- `PPO.update()` sharded over 2 ranks == single process over the union of the shards, replicas
broadcast at construction and bit-identical afterwards, plus BASELINE configs[3]'s chained env + learner loop in small.
Two GPUs -> NCCL, one GPU per rank; one GPU -> gloo with both ranks on cuda:0 (same collective call site, ppo.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(400)
def test_product_update_two_ranks_equals_single_process():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ)
    # the container's hostname may not resolve: keep gloo / NCCL bootstrap traffic on the loopback interface
    env.setdefault("GLOO_SOCKET_IFNAME", "lo")
    env.setdefault("NCCL_SOCKET_IFNAME", "lo")
    env.setdefault("HG_DIST_TIMEOUT_S", "120")
    import signal
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, env=env, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=330)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)          # torchrun AND its workers: nothing may linger on the GPU
        out, err = proc.communicate()
        if "process group up" not in out:
            pytest.skip("the 2-rank process group could not be formed on this box within 330 s (transport, not the product)")
        raise AssertionError("2-rank worker hung after the process group was up:\n" + (out + err)[-3000:])
    r = subprocess.CompletedProcess(cmd, proc.returncode, out, err)
    tail = (r.stdout + "\n" + r.stderr)[-4000:]
    print(tail)
    assert r.returncode == 0, tail
    assert "DIST CHECK OK" in r.stdout, tail
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    assert f"backend: {backend}" in r.stdout
