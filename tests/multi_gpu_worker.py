"""Worker of tests/test_gpu_multi.py: run under torchrun with one rank per GPU.  Every rank shards the same seeded problem
by series, evaluates its shard through libude_cuda with (a) the fused finalize + peer-memory exchange kernel and (b) the
NCCL allreduce, and checks both against the CPU oracle on the UNSHARDED problem (1e-10) and against each other."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import cases
    import ude_b200 as ude
    from oracle import oracle
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    worst = 0.0
    for kind in ("series", "series_param", "wide"):
        if kind == "series":
            prob, theta = cases.make_case("node_d4x1", "multi_shoot", 0, lengths=(9, 4, 7, 1, 12, 6, 3, 8, 5, 11, 2, 10, 6, 7, 3, 9, 4), hidden=32, reg=1e-3)
        elif kind == "series_param":
            prob, theta = cases.make_case("series_growth_d1", "mse", 1, lengths=(9, 4, 7, 5, 6, 3, 8, 4, 5, 6, 7, 8, 9, 3, 4, 5, 6, 7))
        else:
            prob, theta = cases.make_case("node_d8", "cond_lik", 0, lengths=(6, 5, 9, 4, 7, 3, 8, 2, 6, 5, 4, 3, 7, 8, 9, 2, 5), hidden=128)
        L_ref, g_ref = oracle.loss_grad(prob, theta)
        sh = ude.sharding.shard_problem(prob, theta, rank, world)
        n_u = sh.uhat_slices[0][1] - sh.uhat_slices[0][0]
        results = {}
        for mode in ("peer", "nccl"):
            with ude.Engine(sh.problem, device=local) as eng:
                if mode == "peer":
                    mine = torch.frombuffer(bytearray(eng.peer_export(world)), dtype=torch.uint8).to(dev)
                    allb = [torch.zeros(128, dtype=torch.uint8, device=dev) for _ in range(world)]
                    dist.all_gather(allb, mine)
                    eng.peer_attach([b.cpu().numpy().tobytes() for b in allb], world, rank)
                    assert eng.launches_per_eval() == 2
                else:
                    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
                    if rank == 0:
                        idt.copy_(torch.frombuffer(bytearray(ude.get_unique_id()), dtype=torch.uint8))
                    dist.broadcast(idt, 0)
                    eng.comm_init(idt.cpu().numpy().tobytes(), world, rank)
                for rep in range(3):                      # several epochs through the double-buffered exchange
                    L, g = eng.loss_grad(sh.theta)
                results[mode] = (L.copy(), g.copy())
                if mode == "peer" and kind == "series":
                    th2, trace = eng.adam(sh.theta, 1e-3, 6)      # graph-replayed Adam with the fused exchange inside
                    assert eng.adam_used_graph()
                    results["adam"] = (th2[n_u:n_u + sh.pm_shared].copy(), trace.copy())
        for mode in ("peer", "nccl"):
            L, g = results[mode]
            eL = abs(L[0] - L_ref[0]) / abs(L_ref[0])
            # shared process-model gradient is complete on every rank; uhat block is this rank's slice of the global one
            lo, hi, llo = sh.uhat_slices[0]
            e1 = cases.relerr(g[n_u:n_u + sh.pm_shared], g_ref[prob.d * prob.n_cols: prob.d * prob.n_cols + sh.pm_shared])
            e2 = cases.relerr(g[llo:llo + (hi - lo)], g_ref[lo:hi])
            worst = max(worst, eL, e1, e2)
            assert eL < 1e-10 and e1 < 1e-10 and e2 < 1e-10, (kind, mode, eL, e1, e2)
        # the fused exchange sums in rank order: every rank must hold the SAME BITS
        Lp, gp = results["peer"]
        t = torch.from_numpy(np.concatenate([Lp, gp[n_u:n_u + sh.pm_shared]])).to(dev)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        for other in gathered:
            assert torch.equal(other.view(torch.int64), gathered[0].view(torch.int64)), (kind, "ranks disagree bitwise")
        if "adam" in results:
            pm, trace = results["adam"]
            t = torch.from_numpy(np.concatenate([pm, trace.ravel()])).to(dev)
            gathered = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(gathered, t)
            for other in gathered:
                assert torch.equal(other.view(torch.int64), gathered[0].view(torch.int64)), "Adam iterates differ between ranks"
            assert trace.ravel()[-1] < trace.ravel()[0]
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK worst_rel_err=%.3e world=%d" % (worst, world))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
