"""Generate tests/golden/*.npz by running the LIVE reference (read-only at
/root/reference) on seeded inputs.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the reference does not travel to the GPU box):
    python oracle/gen_golden.py
The fixtures are committed; tests read only the .npz files.

The reference package imports `nengo` unconditionally
(ssp_bayes_opt/__init__.py:15-16); nengo is not installed, so the module is
stubbed in sys.modules before import.
"""
import contextlib
import io
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

REF = os.environ.get("SSPBO_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    for name in ("nengo", "nengo.params", "nengo.utils", "nengo.utils.builder", "nengo.dists",
                 "nengo.processes", "nengo.solvers", "nengo.neurons"):
        sys.modules.setdefault(name, MagicMock())
    sys.modules["nengo"].Network = type("Network", (object,), {})
    sys.path.insert(0, REF)
    import ssp_bayes_opt  # noqa: F401
    from ssp_bayes_opt import sspspace, blr, agents
    return sspspace, blr, agents


def himmelblau(x):
    return -((x[:, 0] ** 2 + x[:, 1] - 11) ** 2 + (x[:, 0] + x[:, 1] ** 2 - 7) ** 2
             + x[:, 0] + x[:, 1]) / 100


def gen_acq_ascent(sspspace, agents):
    """8. phi-space acquisition search: `acquisition_func` (ssp_agent.py:156-185) and the L-BFGS-B restarts of
    bayesian_optimization.py:270-280 from `initial_guess()` draws, on an MI agent after 20 sequential updates."""
    from scipy.optimize import minimize
    quiet = contextlib.redirect_stdout(io.StringIO())
    b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    xs = np.random.default_rng(21).uniform(b2[:, 0], b2[:, 1], size=(30, 2))
    ys = himmelblau(xs).reshape(-1, 1)
    with quiet:
        sp = sspspace.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=b2, length_scale=1.0)
        agt = agents.SSPAgent(xs[:10], ys[:10], sp, decoder_method="from-set", length_scale=1.0)
        for i in range(10, 30):
            _, var_t, _ = agt.eval(xs[i:i + 1])
            agt.update(xs[i:i + 1], ys[i:i + 1], var_t, step_num=i)
    min_func, gradient = agt.acquisition_func()
    np.random.seed(3)
    phi0 = np.stack([agt.initial_guess().flatten() for _ in range(5)])
    probe = np.random.default_rng(22).normal(size=(4, sp.ssp_dim)) * np.array([[0.05], [0.2], [0.5], [2.0]])
    sol_x, sol_fun, sol_nit, dec = [], [], [], []
    for p0 in phi0:
        soln = minimize(min_func, p0, jac=gradient, method="L-BFGS-B")
        sol_x.append(soln.x); sol_fun.append(float(np.ravel(soln.fun)[0])); sol_nit.append(soln.nit)
        dec.append(np.ravel(agt.decode(np.copy(np.atleast_2d(soln.x)))))
    np.savez_compressed(
        os.path.join(OUT, "acq_ascent.npz"), A=sp.phase_matrix, ls=np.ravel(sp.length_scale), bounds=b2, xs=xs, ys=ys,
        m=agt.blr.m, S=agt.blr.S, beta=np.float64(np.ravel(agt.blr.beta)[0]), var_weight=np.float64(agt.var_weight),
        gamma_t=np.float64(np.ravel(agt.gamma_t)[0]), phi0=phi0, probe=probe,
        probe_f=np.array([np.ravel(min_func(p))[0] for p in probe]), probe_g=np.stack([gradient(p) for p in probe]),
        phi0_f=np.array([np.ravel(min_func(p))[0] for p in phi0]),
        lbfgs_x=np.stack(sol_x), lbfgs_fun=np.array(sol_fun), lbfgs_nit=np.array(sol_nit), lbfgs_decoded=np.stack(dec),
        init_pts=agt.init_samples[1])


def gen_multi_agent(agents):
    """9. multi-agent trajectory agent (agents/ssp_multi_agent.py): encode (L2-normalised), eval, decode."""
    quiet = contextlib.redirect_stdout(io.StringIO())
    bm = np.array([[-2.0, 2.0], [-2.0, 2.0]])
    out = {}
    for tag, (n_agents, L, ssp_dim, seed) in {"a3": (3, 4, 51, 0), "a2": (2, 5, 120, 4)}.items():
        xd = 2
        rng = np.random.default_rng(30 + n_agents)
        trajs = rng.uniform(-2, 2, size=(12, L * n_agents * xd))
        tys = -np.sum(trajs ** 2, axis=1, keepdims=True) / (L * n_agents)
        with quiet:
            agt = agents.SSPMultiAgent(trajs, tys, n_agents=n_agents, x_dim=xd, traj_len=L, ssp_dim=ssp_dim,
                                       domain_bounds=bm, length_scale=0.8, time_length_scale=1.2, seed=seed,
                                       decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
        xc = rng.uniform(-2, 2, size=(40, L * n_agents * xd))
        mu, var, acq = agt.eval(xc)
        out.update({f"{tag}_Ax": agt.ssp_x_spaces[0].phase_matrix, f"{tag}_At": agt.ssp_t_space.phase_matrix,
                    f"{tag}_ls_x": np.ravel(agt.ssp_x_spaces[0].length_scale), f"{tag}_ls_t": np.ravel(agt.ssp_t_space.length_scale),
                    f"{tag}_agent_sps": agt.agent_sps, f"{tag}_timestep_ssps": agt.timestep_ssps,
                    f"{tag}_trajs": trajs, f"{tag}_tys": tys, f"{tag}_xc": xc, f"{tag}_phi": agt.encode(xc),
                    f"{tag}_mu": mu, f"{tag}_var": var, f"{tag}_acq": acq, f"{tag}_beta": np.float64(np.ravel(agt.blr.beta)[0]),
                    f"{tag}_m": agt.blr.m, f"{tag}_S": agt.blr.S, f"{tag}_init_pts": agt.init_samples[0][1],
                    f"{tag}_decoded": agt.decode(agt.encode(xc[:1])), f"{tag}_dims": np.array([n_agents, L, xd, ssp_dim, seed])})
    out["bounds"] = bm
    np.savez_compressed(os.path.join(OUT, "multi_agent.npz"), **out)


def gen_multi_agent_sep(agents):
    """9b. multi-agent trajectory agent with ONE RandomSSPSpace PER AGENT (same_agt_x_space=False, ssp_multi_agent.py:58-66),
    each on its own rows of domain_bounds: encode, eval, decode; and the K-fold length-scale search of the multi-agent agent
    (ssp_multi_agent.py:131-164), which the reference can only run with one agent (its x0 has n_agents + 1 entries against two
    bounds) -- every objective evaluation recorded as in gen_lengthscale_cv."""
    from ssp_bayes_opt.agents import ssp_multi_agent as ref_mod
    quiet = contextlib.redirect_stdout(io.StringIO())
    out = {}
    for tag, (n_agents, L, ssp_dim, seed) in {"s3": (3, 4, 51, 2), "s2": (2, 5, 121, 7)}.items():
        xd = 2
        bm = np.array([[-2.0, 2.0], [-2.0, 2.0], [-1.0, 3.0], [0.0, 2.5], [-3.0, 1.0], [-2.0, 2.0]])[:n_agents * xd]
        rng = np.random.default_rng(60 + n_agents)
        lo, hi = np.tile(bm[:, 0], L), np.tile(bm[:, 1], L)                # candidate layout (traj_len, n_agents, x_dim)
        trajs = rng.uniform(lo, hi, size=(12, L * n_agents * xd))
        tys = -np.sum(trajs ** 2, axis=1, keepdims=True) / (L * n_agents)
        with quiet:
            agt = agents.SSPMultiAgent(trajs, tys, n_agents=n_agents, x_dim=xd, traj_len=L, ssp_dim=ssp_dim,
                                       domain_bounds=bm, length_scale=0.8, time_length_scale=1.2, seed=seed,
                                       same_agt_x_space=False, decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
        assert agt.ssp_x_spaces[0] is not agt.ssp_x_spaces[1]
        xc = rng.uniform(lo, hi, size=(40, L * n_agents * xd))
        mu, var, acq = agt.eval(xc)
        out.update({f"{tag}_Ax": np.stack([sp.phase_matrix for sp in agt.ssp_x_spaces]), f"{tag}_At": agt.ssp_t_space.phase_matrix,
                    f"{tag}_ls_x": np.stack([np.ravel(sp.length_scale) for sp in agt.ssp_x_spaces]),
                    f"{tag}_ls_t": np.ravel(agt.ssp_t_space.length_scale),
                    f"{tag}_agent_sps": agt.agent_sps, f"{tag}_timestep_ssps": agt.timestep_ssps,
                    f"{tag}_trajs": trajs, f"{tag}_tys": tys, f"{tag}_xc": xc, f"{tag}_phi": agt.encode(xc),
                    f"{tag}_mu": mu, f"{tag}_var": var, f"{tag}_acq": acq, f"{tag}_beta": np.float64(np.ravel(agt.blr.beta)[0]),
                    f"{tag}_m": agt.blr.m, f"{tag}_S": agt.blr.S,
                    f"{tag}_init_pts": np.stack([agt.init_samples[i][1] for i in range(n_agents)]),
                    f"{tag}_decoded": agt.decode(agt.encode(xc[:1])), f"{tag}_bounds": bm,
                    f"{tag}_dims": np.array([n_agents, L, xd, ssp_dim, seed])})
    # length-scale search, one agent
    L, xd = 4, 2
    bt = np.array([[-2.0, 2.0], [-2.0, 2.0]])
    rng = np.random.default_rng(41)
    trajs = rng.uniform(-2, 2, size=(14, L * xd))
    tys = (np.sin(trajs[:, ::2]).sum(axis=1, keepdims=True) - 0.3 * np.sum(trajs[:, 1::2] ** 2, axis=1, keepdims=True)) / L
    log = []
    real_minimize = ref_mod.minimize

    def recording_minimize(fun, *a, **kw):
        def wrapped(x):
            v = fun(x)
            log.append((np.array(x, dtype=np.float64).ravel().copy(), float(v)))
            return v
        kw["x0"] = np.ravel(kw["x0"])                  # (2, 1) -> (2,): see gen_lengthscale_cv
        return real_minimize(wrapped, *a, **kw)

    ref_mod.minimize = recording_minimize
    try:
        with quiet:
            agt = agents.SSPMultiAgent(trajs, tys, n_agents=1, x_dim=xd, traj_len=L, ssp_dim=51, domain_bounds=bt, seed=3,
                                       decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    finally:
        ref_mod.minimize = real_minimize
    out.update(cv_Ax=agt.ssp_x_spaces[0].phase_matrix, cv_At=agt.ssp_t_space.phase_matrix, cv_agent_sps=agt.agent_sps,
               cv_trajs=trajs, cv_tys=tys, cv_bounds=bt, cv_eval_x=np.stack([x for x, _ in log]),
               cv_eval_f=np.array([f for _, f in log]),
               cv_result=np.concatenate([np.ravel(agt.ssp_x_spaces[0].length_scale)[:1], np.ravel(agt.ssp_t_space.length_scale)[:1]]),
               cv_beta=np.float64(np.ravel(agt.blr.beta)[0]), cv_m=agt.blr.m)
    np.savez_compressed(os.path.join(OUT, "multi_agent_sep.npz"), **out)
    print("multi_agent_sep.npz written")


def gen_lengthscale_cv(agents):
    """10. K-fold length-scale search of the trajectory agent (ssp_traj_agent.py:113-143): every objective evaluation
    scipy's L-BFGS-B makes inside the live reference is recorded, together with the result."""
    from ssp_bayes_opt.agents import ssp_traj_agent as ref_mod
    quiet = contextlib.redirect_stdout(io.StringIO())
    bt = np.array([[-2.0, 2.0], [-2.0, 2.0]])
    L, xd = 4, 2
    rng = np.random.default_rng(40)
    trajs = rng.uniform(-2, 2, size=(14, L * xd))
    tys = (np.sin(trajs[:, ::2]).sum(axis=1, keepdims=True) - 0.3 * np.sum(trajs[:, 1::2] ** 2, axis=1, keepdims=True)) / L
    log = []
    real_minimize = ref_mod.minimize

    def recording_minimize(fun, *a, **kw):
        def wrapped(x):
            v = fun(x)
            log.append((np.array(x, dtype=np.float64).ravel().copy(), float(v)))
            return v
        # the reference passes x0 of shape (2, 1); scipy >= 1.11 rejects that ("'x0' must only have one dimension"), older
        # releases squeezed it.  Squeeze here so the reference's search runs under the installed scipy.
        kw["x0"] = np.ravel(kw["x0"])
        return real_minimize(wrapped, *a, **kw)

    ref_mod.minimize = recording_minimize
    try:
        with quiet:
            agt = agents.SSPTrajectoryAgent(trajs, tys, x_dim=xd, traj_len=L, ssp_dim=51, domain_bounds=bt,
                                            decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    finally:
        ref_mod.minimize = real_minimize
    np.savez_compressed(os.path.join(OUT, "lengthscale_cv.npz"), Ax=agt.ssp_x_space.phase_matrix,
                        At=agt.ssp_t_space.phase_matrix, trajs=trajs, tys=tys, bounds=bt,
                        eval_x=np.stack([x for x, _ in log]), eval_f=np.array([f for _, f in log]),
                        result=np.concatenate([np.ravel(agt.ssp_x_space.length_scale)[:1], np.ravel(agt.ssp_t_space.length_scale)[:1]]),
                        beta=np.float64(np.ravel(agt.blr.beta)[0]), m=agt.blr.m)


def gen_lengthscale_gp(sspspace, agents):
    """11. length scale of SSPAgent when none is given: sklearn GP with the reference's sinc kernel
    (agents/ssp_agent.py:96-109, agents/kernels/sinc_kernel.py), on two small designs."""
    from ssp_bayes_opt.agents.kernels.sinc_kernel import SincKernel
    quiet = contextlib.redirect_stdout(io.StringIO())
    b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    out = {"bounds": b2}
    for tag, n, seed in (("a", 12, 50), ("b", 25, 51)):
        xs = np.random.default_rng(seed).uniform(b2[:, 0], b2[:, 1], size=(n, 2))
        ys = himmelblau(xs).reshape(-1, 1)
        with quiet:
            sp = sspspace.HexagonalSSPSpace(2, ssp_dim=51, domain_bounds=b2, length_scale=1.0)
            agt = agents.SSPAgent(xs, ys, sp, decoder_method="from-set")
        out[f"{tag}_xs"], out[f"{tag}_ys"] = xs, ys
        out[f"{tag}_length_scale"] = np.ravel(agt.ssp_space.length_scale)
        out[f"{tag}_m"] = agt.blr.m
    # kernel values and the gradient the reference hands to sklearn, at two length scales
    X = np.random.default_rng(52).uniform(-2, 2, size=(6, 2))
    for i, ls in enumerate((1.0, 0.37)):
        K, dK = SincKernel(length_scale=ls)(X, eval_gradient=True)
        out[f"k{i}_ls"], out[f"k{i}_K"], out[f"k{i}_dK"] = np.float64(ls), K, dK
    out["k_X"] = X
    np.savez_compressed(os.path.join(OUT, "lengthscale_gp.npz"), **out)


def gen_rff(agents):
    """9. Random-Fourier-feature baseline agent (agents/rff_agent.py): features drawn from numpy's global generator after
    np.random.seed(7), both kernel types; encode, eval (MI acquisition), five sequential updates, the x-space objective."""
    quiet = contextlib.redirect_stdout(io.StringIO())
    b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    rng = np.random.default_rng(31)
    xs = rng.uniform(b2[:, 0], b2[:, 1], size=(17, 2))
    ys = himmelblau(xs).reshape(-1, 1)
    xq = rng.uniform(b2[:, 0], b2[:, 1], size=(40, 2))
    out = dict(xs=xs, ys=ys, xq=xq)
    for kt in ("rbf", "matern"):
        np.random.seed(7)
        with quiet:
            agt = agents.RFFAgent(xs[:12], ys[:12], ssp_dim=96, kernel_type=kt, gamma_c=1.0, beta_ucb=5.0, var_decay=0.25)
        out[f"{kt}_W"], out[f"{kt}_B"], out[f"{kt}_scale"] = agt.W, agt.B, agt.sqrt_2_alpha_over_m
        out[f"{kt}_phi"] = agt.encode(xq)
        mu, var, acq = agt.eval(xq)
        out[f"{kt}_mu0"], out[f"{kt}_var0"], out[f"{kt}_acq0"] = mu, var, acq
        for i in range(12, 17):
            _, var_t, _ = agt.eval(xs[i:i + 1])
            agt.update(xs[i:i + 1], ys[i:i + 1], var_t, step_num=i)
        mu, var, acq = agt.eval(xq)
        out[f"{kt}_mu1"], out[f"{kt}_var1"], out[f"{kt}_acq1"] = mu, var, acq
        out[f"{kt}_m"], out[f"{kt}_S"], out[f"{kt}_beta"] = agt.blr.m, agt.blr.S, agt.blr.beta
        out[f"{kt}_gamma"], out[f"{kt}_lam"] = np.ravel(agt.gamma_t), agt.var_weight
        f, g = agt.acquisition_func()
        out[f"{kt}_obj"] = np.array([f(x) for x in xq[:8]]).ravel()
        out[f"{kt}_grad"] = np.array([g(x) for x in xq[:8]])
    np.savez_compressed(os.path.join(OUT, "rff.npz"), **out)
    print("rff.npz written")


def main():
    sspspace, blr, agents = import_reference()
    os.makedirs(OUT, exist_ok=True)
    quiet = contextlib.redirect_stdout(io.StringIO())
    if "--only-acq-ascent" in sys.argv[1:]:
        gen_acq_ascent(sspspace, agents)
        return
    if "--only-multi-agent-sep" in sys.argv[1:]:
        gen_multi_agent_sep(agents)
        return
    if "--only-multi-agent" in sys.argv[1:]:
        gen_multi_agent(agents)
        return
    if "--only-lengthscale-gp" in sys.argv[1:]:
        gen_lengthscale_gp(sspspace, agents)
        return
    if "--only-lengthscale-cv" in sys.argv[1:]:
        gen_lengthscale_cv(agents)
        return

    # ---- 1. phase matrices + encode, the spaces of the reference's own tests ----
    b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    rng = np.random.default_rng(42)
    x2 = rng.uniform(b2[:, 0], b2[:, 1], size=(24, 2))
    spaces = {
        "hex51": sspspace.HexagonalSSPSpace(2, ssp_dim=51, domain_bounds=b2, length_scale=1.0, rng=0),
        "rand51": sspspace.RandomSSPSpace(2, ssp_dim=51, domain_bounds=b2, length_scale=1.0, rng=0),
        "hex151": sspspace.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=b2, length_scale=0.5),
        "rand512": sspspace.RandomSSPSpace(2, ssp_dim=512, domain_bounds=b2, length_scale=1.0, rng=3),
    }
    enc = {}
    for name, sp in spaces.items():
        enc[f"{name}_A"] = sp.phase_matrix
        enc[f"{name}_ls"] = np.asarray(sp.length_scale, dtype=np.float64).reshape(-1)
        enc[f"{name}_phi"] = sp.encode(x2)
    enc["x2"] = x2
    # small length scale / wide bounds (tests/test_ssps.py:92-98): |theta| ~ 1e3 rad
    b15 = 15 * np.array([[-1.0, 1.0], [-1.0, 1.0]])
    sp = sspspace.HexagonalSSPSpace(2, ssp_dim=151, scale_min=2 * np.pi / np.sqrt(6) - 0.5,
                                    scale_max=2 * np.pi / np.sqrt(6) + 0.5, domain_bounds=b15,
                                    length_scale=0.05)
    x15 = np.random.default_rng(7).uniform(-15, 15, size=(24, 2))
    enc["hexsmall_A"] = sp.phase_matrix
    enc["hexsmall_ls"] = np.asarray(sp.length_scale).reshape(-1)
    enc["hexsmall_phi"] = sp.encode(x15)
    enc["x15"] = x15
    # C3 space: Random n=6 ssp_dim=1024 -> d=1023
    b6 = np.array([[0.0, 1.0]] * 6)
    sp6 = sspspace.RandomSSPSpace(6, ssp_dim=1024, domain_bounds=b6, length_scale=0.25, rng=0)
    x6 = np.random.default_rng(11).uniform(0, 1, size=(8, 6))
    enc["rand1024_A"] = sp6.phase_matrix
    enc["rand1024_ls"] = np.asarray(sp6.length_scale).reshape(-1)
    enc["rand1024_phi"] = sp6.encode(x6)
    enc["x6"] = x6
    np.savez_compressed(os.path.join(OUT, "encode.npz"), **enc)

    # ---- 2. grid sample points + from-set decode ----
    sp = spaces["hex51"]
    ssps, pts = sp.get_sample_pts_and_ssps(20, "grid")
    q = sp.encode(np.array([[1.3, -3.4], [0.2, 4.1], [-4.9, -4.9]]))
    dec = sp.decode(q, method="from-set", samples=(ssps, pts))
    b3 = np.array([[0.0, 1.0], [-1.0, 2.0], [3.0, 4.0]])
    sp3 = sspspace.RandomSSPSpace(3, ssp_dim=31, domain_bounds=b3, length_scale=1.0, rng=1)
    pts3 = sp3.get_sample_points(4, "grid")
    np.savez_compressed(os.path.join(OUT, "decode.npz"), A=sp.phase_matrix, bounds=b2, pts=pts,
                        ssps=ssps, queries=q, decoded=dec, bounds3=b3, pts3=pts3,
                        bind=sp.bind(q[0], q[1]), invert=sp.invert(q))

    # ---- 3. BLR: batch init update, sequential updates, predict ----
    out = {}
    xs = np.random.default_rng(42).uniform(b2[:, 0], b2[:, 1], size=(8, 2))
    ys = himmelblau(xs).reshape(-1, 1)
    for name in ("hex51", "rand51"):
        sp = spaces[name]
        model = blr.BayesianLinearRegression(sp.ssp_dim)
        model.update(sp.encode(xs), ys)
        out[f"{name}_beta"] = np.float64(model.beta)
        out[f"{name}_m0"], out[f"{name}_S0"] = model.m.copy(), model.S.copy()
        xt = np.random.default_rng(5).uniform(b2[:, 0], b2[:, 1], size=(40, 2))
        yt = himmelblau(xt).reshape(-1, 1)
        for i in range(40):
            model.update(sp.encode(xt[i:i + 1]), yt[i:i + 1])
        out[f"{name}_m40"], out[f"{name}_S40"] = model.m.copy(), model.S.copy()
        xc = np.random.default_rng(6).uniform(b2[:, 0], b2[:, 1], size=(64, 2))
        mu, var = model.predict(sp.encode(xc))
        out[f"{name}_mu"], out[f"{name}_var"] = mu, var
        out[f"{name}_xt"], out[f"{name}_yt"], out[f"{name}_xc"] = xt, yt, xc
    out["xs"], out["ys"] = xs, ys
    # single-target beta rule (blr.py:57-58)
    model = blr.BayesianLinearRegression(spaces["hex51"].ssp_dim)
    model.update(spaces["hex51"].encode(xs[:1]), ys[:1])
    out["single_beta"] = np.float64(np.ravel(model.beta)[0])
    out["single_m"], out["single_S"] = model.m.copy(), model.S.copy()
    np.savez_compressed(os.path.join(OUT, "blr.npz"), **out)

    # ---- 4. posterior after 1000 sequential updates (north-star 1e-6 bar) ----
    sp = sspspace.RandomSSPSpace(2, ssp_dim=128, domain_bounds=b2, length_scale=1.0, rng=2)  # d=127
    xt = np.random.default_rng(8).uniform(b2[:, 0], b2[:, 1], size=(1000, 2))
    yt = himmelblau(xt).reshape(-1, 1)
    model = blr.BayesianLinearRegression(sp.ssp_dim)
    model.update(sp.encode(xt[:10]), yt[:10])
    for i in range(10, 1000):
        model.update(sp.encode(xt[i:i + 1]), yt[i:i + 1])
    np.savez_compressed(os.path.join(OUT, "blr1000.npz"), A=sp.phase_matrix, xt=xt, yt=yt,
                        beta=np.float64(model.beta), m=model.m, S=model.S)

    # ---- 5. SSPAgent.eval / update scalars (MI and pure UCB) ----
    out = {}
    for tag, gamma_c in (("mi", 1.0), ("ucb", 0.0)):
        sp = sspspace.HexagonalSSPSpace(2, ssp_dim=51, domain_bounds=b2, length_scale=1.0, rng=0)
        with quiet:
            agt = agents.SSPAgent(xs, ys, sp, decoder_method="from-set", gamma_c=gamma_c,
                                  beta_ucb=10.0, var_decay=0.25, length_scale=1.0)
        xt = np.random.default_rng(9).uniform(b2[:, 0], b2[:, 1], size=(6, 2))
        for i in range(6):
            x_t = xt[i:i + 1]
            y_t = np.atleast_2d(himmelblau(x_t))
            _, var_t, _ = agt.eval(x_t)
            agt.update(x_t, y_t, var_t, step_num=i)
        xc = np.random.default_rng(10).uniform(b2[:, 0], b2[:, 1], size=(256, 2))
        mu, var, acq = agt.eval(xc)
        out[f"{tag}_mu"], out[f"{tag}_var"], out[f"{tag}_acq"] = mu, var, acq
        out[f"{tag}_gamma_t"] = np.float64(np.ravel(agt.gamma_t)[0])
        out[f"{tag}_var_weight"] = np.float64(agt.var_weight)
        out[f"{tag}_m"], out[f"{tag}_S"] = agt.blr.m, agt.blr.S
        out[f"{tag}_beta"] = np.float64(agt.blr.beta)
        out[f"{tag}_init_pts"] = agt.init_samples[1]
        out[f"{tag}_decoded"] = agt.decode(agt.encode(xc[:3]))
        out["xt"], out["xc"] = xt, xc
    np.savez_compressed(os.path.join(OUT, "agent.npz"), **out)

    # ---- 6. trajectory agent encode / decode ----
    bt = np.array([[-2.0, 2.0], [-2.0, 2.0]])
    L, xd = 4, 2
    trajs = np.random.default_rng(12).uniform(-2, 2, size=(10, L * xd))
    tys = -np.sum(trajs ** 2, axis=1, keepdims=True)
    with quiet:
        agt = agents.SSPTrajectoryAgent(trajs, tys, x_dim=xd, traj_len=L, ssp_dim=51, domain_bounds=bt,
                                        length_scale=0.7, time_length_scale=1.3,
                                        decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    xc = np.random.default_rng(13).uniform(-2, 2, size=(32, L * xd))
    mu, var, acq = agt.eval(xc)
    np.savez_compressed(
        os.path.join(OUT, "traj.npz"), Ax=agt.ssp_x_space.phase_matrix, At=agt.ssp_t_space.phase_matrix,
        ls_x=np.ravel(agt.ssp_x_space.length_scale), ls_t=np.ravel(agt.ssp_t_space.length_scale),
        timestep_ssps=agt.timestep_ssps, trajs=trajs, tys=tys, xc=xc, phi=agt.encode(xc), mu=mu, var=var,
        acq=acq, beta=np.float64(agt.blr.beta), m=agt.blr.m, S=agt.blr.S, init_pts=agt.init_samples[1],
        decoded=agt.decode(agt.encode(xc[:1])), bounds=bt)
    # ---- 7. direct-optim decode (sspspace.py:358-375) with an explicit decoding set ----
    bo = np.array([[-4.0, 4.0], [-4.0, 4.0], [-4.0, 4.0]])
    spo = sspspace.RandomSSPSpace(3, ssp_dim=201, domain_bounds=bo, length_scale=1.5, rng=1)
    rngo = np.random.default_rng(9)
    xo = rngo.uniform(-3.5, 3.5, (8, 3))
    xo[0] = [4.0, -4.0, 3.9]
    ssps_o = spo.encode(xo) + 0.02 * rngo.normal(size=(8, spo.ssp_dim))
    set_ssps, set_pts = spo.get_sample_pts_and_ssps(12, "grid")
    dec = spo.decode(ssps_o, method="direct-optim", samples=(set_ssps, set_pts))
    dec0 = spo.decode(ssps_o, method="from-set", samples=(set_ssps, set_pts))
    np.savez_compressed(os.path.join(OUT, "decode_optim.npz"), A=spo.phase_matrix, ls=np.ravel(spo.length_scale), bounds=bo,
                        ssps=ssps_o, set_pts=set_pts, x0=dec0, decoded=dec)
    gen_acq_ascent(sspspace, agents)
    gen_multi_agent(agents)
    gen_multi_agent_sep(agents)
    gen_lengthscale_cv(agents)
    gen_lengthscale_gp(sspspace, agents)
    gen_rff(agents)
    print("golden fixtures written to", os.path.abspath(OUT))
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
