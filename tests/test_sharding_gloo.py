"""N > 1 host logic on CPU: two gloo ranks shard a MultiNODE problem by series, evaluate their shards (the oracle stands in
for the GPU kernel here) and sum-allreduce [shared process-model gradient | loss] -- the exchange libude_cuda performs
with one NCCL call.  The result must equal the unsharded evaluation."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import cases
    import ude_b200 as ude
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if kind == "series":
        prob, theta = cases.make_case("node_d4x1", "multi_shoot", 0, lengths=(9, 4, 7, 1, 12, 6, 3), hidden=6)
    elif kind == "series_param":
        prob, theta = cases.make_case("series_growth_d1", "mse", 1, lengths=(9, 4, 7, 5, 6))
    else:
        prob, theta = cases.make_case("node_time_d3", "cond_lik", 0, lengths=(9, 8, 7, 6, 5, 4), n_models=6)
    sh = ude.sharding.shard_problem(prob, theta, rank, world)
    L, g = oracle.loss_grad(sh.problem, sh.theta)
    if prob.n_models == 1:
        n_u = sh.uhat_slices[0][1] - sh.uhat_slices[0][0]
        buf = torch.from_numpy(np.concatenate([g[n_u:n_u + sh.pm_shared], L]))
        dist.all_reduce(buf)                       # the one collective of the path
        red = buf.numpy()
        g[n_u:n_u + sh.pm_shared] = red[:-1]
        L = red[-1:]
    q.put((rank, sh, L, g))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["series", "series_param", "models"])
def test_two_rank_sharding_matches_unsharded(kind):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    import ude_b200 as ude
    from oracle import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if kind == "series":
        prob, theta = cases.make_case("node_d4x1", "multi_shoot", 0, lengths=(9, 4, 7, 1, 12, 6, 3), hidden=6)
    elif kind == "series_param":
        prob, theta = cases.make_case("series_growth_d1", "mse", 1, lengths=(9, 4, 7, 5, 6))
    else:
        prob, theta = cases.make_case("node_time_d3", "cond_lik", 0, lengths=(9, 8, 7, 6, 5, 4), n_models=6)
    L_ref, g_ref = oracle.loss_grad(prob, theta)
    shards = [r[1] for r in res]
    if prob.n_models == 1:
        for r in res:
            assert abs(r[2][0] - L_ref[0]) < 1e-12 * abs(L_ref[0])        # every rank holds the full loss after the allreduce
        n_u0 = shards[0].uhat_slices[0][1] - shards[0].uhat_slices[0][0]
        g = ude.sharding.gather_gradient(prob, shards, [r[3] for r in res], reduced_pm_grad=res[0][3][n_u0:n_u0 + shards[0].pm_shared])
    else:
        L = np.concatenate([r[2] for r in res])
        assert np.allclose(L, L_ref, rtol=1e-13)
        g = ude.sharding.gather_gradient(prob, shards, [r[3] for r in res])
    assert cases.relerr(g, g_ref) < 1e-12
