"""Randomised differential test of the data-parallel path: every rank shards the same random single-model problem by series
(sharding.shard_problem), evaluates its shard with the fused finalize + peer-memory exchange, and the summed loss / shared
process-model gradient / own uhat-gradient slice are compared with the C oracle on the UNSHARDED problem (1e-10); all ranks
must hold the same bits.  usage (under gpurun --gpus 2):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/fuzz_multi_gpu.py --cases 80"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=80)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import cases
    import ude_b200 as ude
    from oracle import oracle
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(a.seed)                  # same stream on every rank
    names = list(cases.RHS_SPECS)
    fails, ran = 0, 0
    for it in range(a.cases):
        rhs = names[int(rng.integers(len(names)))]
        loss = cases.LOSSES[int(rng.integers(len(cases.LOSSES)))]
        solver = int(rng.integers(2))
        hidden = int(rng.choice([3, 5, 8, 10, 12, 32 if cases.RHS_SPECS[rhs][0] == 0 else 10]))
        n_series = int(rng.integers(world, 4 * world + 3))
        lengths = tuple(int(x) for x in rng.choice([1, 2, 3, 4, 6, 9, 14], size=n_series))
        kw = dict(hidden=hidden, lengths=lengths, seed=int(rng.integers(1 << 30)), substeps=int(rng.choice([1, 2, 4])),
                  skip=bool(rng.integers(2)) and loss in ("cond_lik", "mse"))
        prob, theta = cases.make_case(rhs, loss, solver, **kw)
        L_ref, g_ref = oracle.loss_grad(prob, theta)
        sh = ude.sharding.shard_problem(prob, theta, rank, world)
        n_u = sh.uhat_slices[0][1] - sh.uhat_slices[0][0]
        ok = True
        try:
            with ude.Engine(sh.problem, device=local) as eng:
                mine = torch.frombuffer(bytearray(eng.peer_export(world)), dtype=torch.uint8).to(dev)
                allb = [torch.zeros(128, dtype=torch.uint8, device=dev) for _ in range(world)]
                dist.all_gather(allb, mine)
                eng.peer_attach([b.cpu().numpy().tobytes() for b in allb], world, rank)
                for rep in range(2):
                    L, g = eng.loss_grad(sh.theta)
            lo, hi, llo = sh.uhat_slices[0]
            pmb = prob.d * prob.n_cols
            eL = abs(L[0] - L_ref[0]) / abs(L_ref[0])
            e1 = cases.relerr(g[n_u:n_u + sh.pm_shared], g_ref[pmb:pmb + sh.pm_shared]) if sh.pm_shared else 0.0
            e2 = cases.relerr(g[llo:llo + (hi - lo)], g_ref[lo:hi]) if hi > lo and np.any(g_ref[lo:hi]) else float(np.max(np.abs(g[llo:llo + (hi - lo)]))) if hi > lo else 0.0
            ok = eL < 1e-10 and e1 < 1e-10 and e2 < 1e-10
            t = torch.from_numpy(np.concatenate([L, g[n_u:n_u + sh.pm_shared]])).to(dev)
            gathered = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(gathered, t)
            ok = ok and all(torch.equal(o.view(torch.int64), gathered[0].view(torch.int64)) for o in gathered)
            if not ok:
                print("FAIL rank", rank, it, rhs, loss, solver, kw, eL, e1, e2, flush=True)
        except Exception as ex:      # every rank sees the same refusals (shape), so the collectives stay aligned
            ok = False
            print("EXC rank", rank, it, rhs, loss, solver, kw, repr(ex)[:300], flush=True)
        flag = torch.tensor([0 if ok else 1], device=dev)
        dist.all_reduce(flag)
        fails += int(flag.item() > 0)
        ran += 1
    dist.barrier()
    if rank == 0:
        print("multi-gpu cases", ran, "world", world, "failures", fails)
    dist.destroy_process_group()
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
