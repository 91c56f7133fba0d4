"""N > 1 on real GPUs (skipped on a one-GPU box): fused finalize + peer-memory exchange and the NCCL path, both against the
oracle on the unsharded problem.  The world-size-2 gloo test (test_sharding_gloo.py) covers the host logic on CPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_or_more_gpus_peer_exchange_and_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    env = dict(os.environ, UDE_PEER_TIMEOUT_MS="15000")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_OK" in out.stdout


@pytest.mark.gpu
def test_peer_exchange_between_threads_of_one_process():
    """ude_peer_attach with ranks that are host threads of ONE process driving one GPU each (one handle per device
    from Julia tasks): buffers attach by pointer after cudaDeviceEnablePeerAccess instead of through CUDA IPC."""
    import threading
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    import ude_b200 as ude
    from oracle import oracle
    prob, theta = cases.make_case("node_d4x1", "multi_shoot", 0, lengths=(9, 4, 7, 1, 12, 6, 3, 8, 5, 11), hidden=32, reg=1e-3)
    L_ref, g_ref = oracle.loss_grad(prob, theta)
    world = 2
    shards = [ude.sharding.shard_problem(prob, theta, r, world) for r in range(world)]
    engines = [ude.Engine(shards[r].problem, device=r) for r in range(world)]
    blobs = [engines[r].peer_export(world) for r in range(world)]
    for r in range(world):
        engines[r].peer_attach(blobs, world, r)
    out, errs = [None] * world, []

    def work(r):
        try:
            for _ in range(3):
                out[r] = engines[r].loss_grad(shards[r].theta)
        except Exception as ex:      # noqa: BLE001
            errs.append(ex)

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    for e in engines:
        e.close()
    assert not errs, errs
    n_u0 = shards[0].uhat_slices[0][1] - shards[0].uhat_slices[0][0]
    n_u1 = shards[1].uhat_slices[0][1] - shards[1].uhat_slices[0][0]
    pm0, pm1 = out[0][1][n_u0:n_u0 + shards[0].pm_shared], out[1][1][n_u1:n_u1 + shards[1].pm_shared]
    assert np.array_equal(pm0, pm1) and out[0][0][0] == out[1][0][0]
    assert abs(out[0][0][0] - L_ref[0]) < 1e-10 * abs(L_ref[0])
    base = prob.d * prob.n_cols
    assert cases.relerr(pm0, g_ref[base:base + shards[0].pm_shared]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("loss_function", ["multiple shooting", "conditional likelihood", "derivative matching"])
def test_train_on_two_gpus_equals_one_gpu(loss_function):
    """train_(model, devices=(0, 1)): data-parallel device-resident Adam from one process (threads + peer exchange by pointer)
    gives the single-GPU result (the sums are taken in a different order: 1e-9 after 25 iterations)."""
    import numpy as np
    import pandas as pd
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, ROOT)
    import ude_b200 as ude
    Y, ts = ude.synthetic.lotka_volterra_data(np.random.default_rng(0), 7, 12, random_u0=True)
    dfm = pd.concat([pd.DataFrame({"time": ts, "series": s + 1, "x1": Y[s, 0], "x2": Y[s, 1]}) for s in range(7)], ignore_index=True)
    Xc = pd.concat([pd.DataFrame({"time": np.linspace(-0.5, 4.0, 6), "series": s + 1, "X": np.cos(s + np.linspace(0, 2, 6))}) for s in range(7)],
                   ignore_index=True)
    a = ude.MultiNODE(dfm, Xc, hidden_units=6, seed=1, reg_weight=1e-4)
    b = ude.MultiNODE(dfm, Xc, hidden_units=6, seed=1, reg_weight=1e-4)
    kw = dict(verbose=False, regularization_weight=1e-3, optim_options=dict(maxiter=25, step_size=0.01))
    ude.train_(a, loss_function, **kw)
    ude.train_(b, loss_function, devices=(0, 1), **kw)
    assert np.max(np.abs(a.extra["loss_trace"].ravel() - b.extra["loss_trace"].ravel()) / np.abs(a.extra["loss_trace"].ravel())) < 1e-9
    assert np.max(np.abs(a.parameters - b.parameters)) < 1e-9 * max(1.0, np.max(np.abs(a.parameters)))


@pytest.mark.gpu
def test_batched_cross_validation_on_two_gpus():
    """leave_future_out_batched(devices=(0, 1)): folds x members are whole models per GPU; identical to one GPU (no exchange,
    same arithmetic per model)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    sys.path.insert(0, ROOT)
    import ude_b200 as ude
    df = ude.LotkaVolterra(plot=False, datasize=18)
    m = ude.NODE(df, hidden_units=5, seed=3)
    kw = dict(k=3, loss_function="conditional likelihood", optim_options=dict(maxiter=10), member_seeds=[1, 2], do_forecast=False)
    one = ude.leave_future_out_batched(m, **kw)
    two = ude.leave_future_out_batched(m, devices=(0, 1), **kw)
    assert len(one[0]) == len(two[0]) == 6
    for a, b in zip(one[0], two[0]):
        assert np.max(np.abs(a.parameters - b.parameters)) < 1e-12
    assert np.max(np.abs(one[4] - two[4])) < 1e-10 * np.max(np.abs(one[4]))
