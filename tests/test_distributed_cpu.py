"""World-size-2 gloo test of the multi-GPU host logic: static block partition of the batch, independent
per-rank solves (the oracle stands in for the GPU here), one all-gather of objectives / iterations / statuses."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from interiorpointmethod_b200.batch import gather_results, shard_range
    from oracle import ipm_oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_range(B, rank, world)
    obj, its, st = [], [], []
    for i in range(first, first + count):
        A, b, c = orc.synthetic_dense_lp(12, 30, i)
        r = orc.solve(A, b, c, tol=1e-8, max_iter=50000, y0_is_one=False, linear="normal")
        obj.append(r["obj"]); its.append(r["k"]); st.append(r["status"])
    o, k, s = gather_results(torch.tensor(obj, dtype=torch.float64), torch.tensor(its, dtype=torch.int32),
                             torch.tensor(st, dtype=torch.int32))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), obj=o.numpy(), its=k.numpy(), st=s.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 8])
def test_sharded_batch_gathers_in_lp_order(tmp_path, B):
    import torch.multiprocessing as mp
    from oracle import ipm_oracle as orc

    port = 29500 + (os.getpid() % 2000) + B
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["obj"], r1["obj"]) and len(r0["obj"]) == B
    for i in range(B):
        A, b, c = orc.synthetic_dense_lp(12, 30, i)
        r = orc.solve(A, b, c, tol=1e-8, max_iter=50000, y0_is_one=False, linear="normal")
        assert r0["obj"][i] == r["obj"] and r0["its"][i] == r["k"] and r0["st"][i] == 0
