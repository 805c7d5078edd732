"""The doc-sharded search across PROCESSES (one rank per GPU, NCCL + the peer-memory exchange kernel): skipped on a
one-GPU box.  tools/check_sharded.py does the work under torchrun; this test launches it and reads its verdict."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("k", [10, 100])
def test_sharded_search_across_processes_all_queries_vs_oracle(k):
    world = min(4, torch.cuda.device_count())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29533 + k % 7), os.path.join(ROOT, "tools", "check_sharded.py"), "--k", str(k)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert p.returncode == 0 and lines, p.stdout[-2000:] + p.stderr[-2000:]
    rec = json.loads(lines[-1])
    assert rec["world"] == world and rec["queries_checked"] == 512
    for key in ("n_docs_global", "global_idf_equals_reference_idf", "peer_exchange", "nccl_all_gather", "pipeline_lanes"):
        assert rec[key] is True, rec
