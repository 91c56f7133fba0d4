import sys, os, argparse; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cases
import ude_b200 as ude
prob, theta = ude.synthetic.config_c5(n_segments=200000)
res = {}
for name, dt, kern in (("f64", 0, ""), ("tf32", 1, "tf32"), ("f32seg", 1, "g32h4")):
    os.environ["UDE_KERNEL"] = kern
    with ude.Engine(prob, dtype=dt) as eng:
        L, g = eng.loss_grad(theta)
        res[name] = (L.copy(), g.copy(), eng.kernel_name())
g0 = res["f64"][1]
for name in ("tf32", "f32seg"):
    L, g, kn = res[name]
    print(name, kn, "loss rel", abs(L[0]-res["f64"][0][0])/abs(res["f64"][0][0]))
    for leaf, lo, hi in cases.leaf_groups(prob, 0):
        a, b = g[lo:hi], g0[lo:hi]
        print("   %-6s max|b| %.3e  max abs err %.3e  rel-to-max %.2e  l1 rel %.2e  mean signed err/mean|b| %.2e" % (leaf, np.max(np.abs(b)), np.max(np.abs(a-b)), np.max(np.abs(a-b))/np.max(np.abs(b)), abs(np.sum(np.abs(a))-np.sum(np.abs(b)))/np.sum(np.abs(b)), np.mean(a-b)/np.mean(np.abs(b))))
    n_u = prob.d*prob.n_cols
    d = (g[:n_u]-g0[:n_u]).reshape(-1, 8)
    print("   uhat err by component:", np.max(np.abs(d),axis=0))
    print("   uhat err even cols (series starts) vs odd:", np.max(np.abs(d[0::2])), np.max(np.abs(d[1::2])))
