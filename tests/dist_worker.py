"""Worker launched by tests/test_dist_gloo.py (CPU, gloo) and tests/test_gpu_multi.py (GPU, NCCL) under
`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ...`.

mode gloo : host-side logic of the sharded path -- index-range sharding, unique-id broadcast,
            all-reduce of per-shard deposits (oracle deposits stand in for the device kernels),
            max-over-ranks timing.  No GPU needed.
mode nccl : the real thing -- every rank owns a shard of the particles on its GPU, libgempic_b200
            all-reduces the grid moments with NCCL; the result must equal the 1-rank run of the
            full set computed by rank 0 on its own GPU with a second, un-sharded group.  Covers
            HamiltonianSplitting{1,2} (fused / un-fused), HamiltonianSplittingBoris (fused step / separate
            pushes) and HamiltonianSplitting{2,3} (fused / un-fused) -- BASELINE configs 3, 4, 5.
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402
from tests.helpers import landau_state, rel_err  # noqa: E402


def run_gloo(out_path):
    gp = ge.load_package()
    from oracle import oracle as orc

    orc.build()
    dc = gp.DistributedContext(backend="gloo")
    assert dc.world_size >= 2
    n, nx, L = 20_001, 32, 4 * math.pi
    state = landau_state(n, L, seed=99)
    first, count = dc.shard(n)
    # shards tile the index range
    counts = dc.allreduce_sum(np.array([count], dtype=np.float64))
    assert int(counts[0]) == n
    # id broadcast: rank 0's 128-byte blob reaches everyone
    blob = bytes(range(128)) if dc.rank == 0 else None
    got = dc.broadcast_bytes(blob, 128, src=0)
    assert got == bytes(range(128))
    # per-shard deposit + all-reduce == full deposit
    mesh = orc.OneDGrid(0.0, L, nx)
    ks0 = orc.ParticleMeshCoupling1D(mesh, n, 3, "galerkin")

    def deposit(cols):
        pg = orc.ParticleGroup(1, 2, cols.shape[1], common_weight=1.0 / n)
        pg.array[:, :] = cols
        rho = np.zeros(nx)
        for i in range(cols.shape[1]):
            ks0.add_charge(rho, pg.array[0, i], pg.get_charge(i))
        return rho

    rho_local = deposit(state[:, first:first + count])
    rho_sum = dc.allreduce_sum(rho_local)
    rho_full = deposit(state)
    err = rel_err(rho_sum, rho_full)
    assert err < 1e-13, err
    t = dc.max_over_ranks(float(dc.rank + 1))
    assert t == float(dc.world_size)
    dc.barrier()
    if dc.rank == 0:
        json.dump({"ok": True, "world_size": dc.world_size, "err": err}, open(out_path, "w"))
    dc.finalize()


def run_nccl(out_path):
    import torch

    gp = ge.load_package()
    dc = gp.DistributedContext(backend="nccl")
    assert dc.world_size >= 2 and torch.cuda.device_count() >= dc.world_size
    dc.init_library_comm()
    errs = gp.sharded_parity(dc)
    dc.barrier()
    dc.finalize()
    if dc.rank == 0:
        ok = all(v < 1e-11 for v in errs.values())
        json.dump({"ok": ok, "world_size": dc.world_size, "errs": errs}, open(out_path, "w"))
        assert ok, errs


if __name__ == "__main__":
    mode, out = sys.argv[1], sys.argv[2]
    (run_gloo if mode == "gloo" else run_nccl)(out)
