import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cases
import ude_b200 as ude
from oracle import ukf_oracle_np as ukf
for rhs, hidden, lengths in (("node_d2",5,(8,5)), ("node_d4x1",6,(6,4)), ("lv_ude_abs",4,(7,))):
    prob, theta = cases.make_case(rhs, "cond_lik", 0, hidden=hidden, lengths=lengths, seed=0)
    d=prob.d; rng=np.random.default_rng(1)
    F = 0.3*np.eye(d)+0.05*rng.normal(size=(d,d)); Peta = 0.1*np.eye(d)+0.02*np.ones((d,d))
    for alpha in (1.0, 1e-1, 1e-2, 1e-3):
        a = ukf.loss(prob, theta, F, Peta, alpha=alpha)
        b = float(ukf.loss(prob, theta.astype(np.longdouble), F, Peta, alpha=alpha))
        g_pm, g_F = ukf.complex_step_grad(prob, theta, F, Peta, alpha=alpha)
        with ude.Engine(prob) as eng:
            L, g, gF = eng.ukf_loss_grad(theta, F, Peta, alpha=alpha)
        n_u = d*prob.n_cols
        print(rhs, alpha, "f64-oracle vs longdouble %.2e | cuda vs longdouble %.2e | cuda vs f64-oracle %.2e | grad pm %.2e  grad F %.2e" % (
            abs(a-b)/abs(b), abs(L-b)/abs(b), abs(L-a)/abs(a), cases.relerr(g[n_u:], g_pm), cases.relerr(gF.reshape(-1,order="F"), g_F)))
