"""-m gpu, needs >= 2 GPUs (skipped on the 1-GPU driver box; run with `gpurun --gpus 2`): N ranks over NCCL against the
single-process global-batch oracle + exact cross-rank replica equality (tools/ddp_check.py holds the assertions)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_matches_global_batch_oracle_and_replicas_stay_identical(tmp_path):
    out = tmp_path / "ddp.json"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", str(ROOT / "tools" / "ddp_check.py"), str(out)],
                       capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert out.exists() and '"replicas_bit_identical_after_5_steps": true' in out.read_text()
