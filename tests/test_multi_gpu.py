"""BASELINE config 5 on real GPUs, under pytest -m gpu: spawns tests/multi_gpu_check.py with torch.distributed.run over
ALL visible GPUs (skipped below 2): the 65536^2 synthetic grid as row bands over the P2P transport must equal the
single-grid run (device digests), and a 2048^2 instance with the same cut patterns must equal the CPU oracle -- with
aligned cuts and with cuts that run through the ring tiles."""
import os
import socket
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_row_bands_on_all_gpus_equal_single_grid_and_oracle(transport):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (row bands, one process per GPU)")
    free, _total = torch.cuda.mem_get_info(0)
    size = 65536 if free > 150e9 else 32768          # rank 0 also holds the single grid: 82 GB + its band at 65536^2
    env = dict(os.environ, CHECK_SIZE=str(size), CHECK_TICKS="50", CHECK_TRANSPORT=transport)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "multi_gpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500)
    print(r.stdout[-4000:])
    assert r.returncode == 0 and "MULTI_GPU_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
