"""world_size-2 `gloo` coverage (CPU) of the host-side logic of the sharded N > 1 path
(SURVEY section 8e): index-range sharding, id broadcast, all-reduce of per-shard deposits."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(mode, nproc, out, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), mode, out]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_shard_range_tiles_the_particles(gp):
    for n, g in ((10, 3), (7, 8), (1_000_000_000, 8), (0, 2), (5, 1)):
        spans = [gp.shard_range(n, g, r) for r in range(g)]
        assert sum(c for _, c in spans) == n
        pos = 0
        for first, count in spans:
            assert count >= 0 and (count == 0 or first == pos)
            pos += count
    with pytest.raises(ValueError):
        gp.shard_range(10, 2, 2)


def test_two_rank_gloo(tmp_path):
    out = str(tmp_path / "gloo.json")
    r = launch("gloo", 2, out)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    assert res["ok"] and res["world_size"] == 2 and res["err"] < 1e-13
