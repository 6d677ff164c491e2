"""Multi-GPU parity (NCCL): a 2-rank sharded run equals the un-sharded run.  Needs >= 2 GPUs on the box
(`gpurun --gpus 2`); skipped on a single-GPU box."""
import json

import pytest

from .test_dist_gloo import launch

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_equals_unsharded(tmp_path, nproc):
    if _n_gpus() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    out = str(tmp_path / "nccl.json")
    r = launch("nccl", nproc, out, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    assert res["ok"], res
