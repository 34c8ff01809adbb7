"""Global-batch graph option of the data-parallel path (`ddp.attach(model, global_graph=True)`): two ranks equal one
process on the concatenated batch.  Runs both ranks on cuda:0 over gloo so that it needs a single GPU (the NCCL variant of the
same script is `torchrun ... tests/ddp_global_graph_check.py` on a multi-GPU box)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_with_global_graph_equal_one_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, DDP_ONE_GPU="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "ddp_global_graph_check.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode(errors="replace")
    print(out[-3000:])
    assert r.returncode == 0 and "ddp_global_graph_check ok" in out
