"""Pin the numpy oracle against outputs of the live reference (tests/golden/*.npz,
written by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import ssp_oracle as orc

B2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])


def test_phase_matrices_bit_exact(golden):
    g = golden("encode.npz")
    assert np.array_equal(orc.random_phase_matrix(2, 51, rng=0), g["rand51_A"])
    assert np.array_equal(orc.random_phase_matrix(2, 512, rng=3), g["rand512_A"])
    assert np.array_equal(orc.random_phase_matrix(6, 1024, rng=0), g["rand1024_A"])
    assert g["rand1024_A"].shape == (1023, 6)          # ssp_dim is a request (sspspace.py:821-827)
    assert np.array_equal(orc.hex_phase_matrix(2, 51), g["hex51_A"])
    assert g["hex51_A"].shape[0] == 25
    assert np.array_equal(orc.hex_phase_matrix(2, 151), g["hex151_A"])
    s = 2 * np.pi / np.sqrt(6)
    assert np.array_equal(orc.hex_phase_matrix(2, 151, scale_min=s - 0.5, scale_max=s + 0.5), g["hexsmall_A"])


def test_encode_matches_reference(golden):
    g = golden("encode.npz")
    for name, xk in (("hex51", "x2"), ("rand51", "x2"), ("hex151", "x2"), ("rand512", "x2"),
                     ("hexsmall", "x15"), ("rand1024", "x6")):
        phi = orc.encode(g[f"{name}_A"], g[f"{name}_ls"], g[xk])
        assert phi.shape == g[f"{name}_phi"].shape
        np.testing.assert_allclose(phi, g[f"{name}_phi"], rtol=0, atol=1e-15)
        # unit norm (tests/test_ssps.py:31-37) and the closed form of SURVEY section 7
        np.testing.assert_allclose(np.linalg.norm(phi, axis=1), 1.0, atol=1e-12)
        np.testing.assert_allclose(orc.encode_closed_form(g[f"{name}_A"], g[f"{name}_ls"], g[xk]), phi,
                                   atol=5e-13)


def test_grid_points_decode_bind(golden):
    g = golden("decode.npz")
    pts = orc.grid_sample_points(g["bounds"], 20)
    assert np.array_equal(pts, g["pts"])
    assert np.array_equal(orc.grid_sample_points(g["bounds3"], 4), g["pts3"])   # n>=3 meshgrid('xy') order
    ssps = orc.encode(g["A"], np.ones(2), pts)
    np.testing.assert_allclose(ssps, g["ssps"], atol=1e-15)
    assert np.array_equal(orc.decode_from_set(g["queries"], g["ssps"], g["pts"]), g["decoded"])
    np.testing.assert_allclose(orc.bind(g["queries"][0], g["queries"][1]), g["bind"], atol=1e-15)
    assert np.array_equal(orc.invert(g["queries"]), g["invert"])


def test_blr_matches_reference(golden):
    g = golden("blr.npz")
    e = golden("encode.npz")
    for name in ("hex51", "rand51"):
        A, ls = e[f"{name}_A"], e[f"{name}_ls"]
        model = orc.BLR(A.shape[0])
        model.update(orc.encode(A, ls, g["xs"]), g["ys"])
        assert model.beta == g[f"{name}_beta"]
        np.testing.assert_allclose(model.m, g[f"{name}_m0"], rtol=0, atol=1e-11 * np.abs(model.m).max())
        np.testing.assert_allclose(model.S, g[f"{name}_S0"], rtol=0, atol=1e-11 * np.abs(model.S).max())
        for i in range(40):
            model.update(orc.encode(A, ls, g[f"{name}_xt"][i:i + 1]), g[f"{name}_yt"][i:i + 1])
        np.testing.assert_allclose(model.m, g[f"{name}_m40"], rtol=0, atol=1e-10 * np.abs(model.m).max())
        np.testing.assert_allclose(model.S, g[f"{name}_S40"], rtol=0, atol=1e-10 * np.abs(model.S).max())
        mu, var = model.predict(orc.encode(A, ls, g[f"{name}_xc"]))
        assert mu.shape == (1, 64) and var.shape == (64,)          # tests/test_blr.py:28-39
        np.testing.assert_allclose(mu, g[f"{name}_mu"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(var, g[f"{name}_var"], rtol=1e-9)
    model = orc.BLR(e["hex51_A"].shape[0])
    model.update(orc.encode(e["hex51_A"], e["hex51_ls"], g["xs"][:1]), g["ys"][:1])
    assert np.ravel(model.beta)[0] == g["single_beta"]
    np.testing.assert_allclose(model.S, g["single_S"], rtol=1e-12, atol=1e-13)


def test_blr_1000_updates(golden):
    g = golden("blr1000.npz")
    A = g["A"]
    model = orc.BLR(A.shape[0])
    phis = orc.encode(A, np.ones(2), g["xt"])
    model.update(phis[:10], g["yt"][:10])
    for i in range(10, 1000):
        model.update(phis[i:i + 1], g["yt"][i:i + 1])
    assert model.beta == g["beta"]
    np.testing.assert_allclose(model.S, g["S"], rtol=0, atol=1e-9 * np.abs(g["S"]).max())
    np.testing.assert_allclose(model.m, g["m"], rtol=0, atol=1e-9 * np.abs(g["m"]).max())


def test_blr_sequential_equals_batch():
    """The reference's own invariant, tests/test_blr.py:72-92."""
    rng = np.random.default_rng(4)
    xs, ys = rng.normal(size=(6, 10)), rng.normal(size=(6, 1))
    a, b = orc.BLR(10, beta=1.0), orc.BLR(10, beta=1.0)
    a.update(xs, ys)
    for i in range(6):
        b.update(xs[i:i + 1], ys[i:i + 1])
    assert np.allclose(a.m, b.m, atol=1e-6) and np.allclose(a.S, b.S, atol=1e-6)


def test_agent_eval_matches_reference(golden):
    g = golden("agent.npz")
    A = golden("encode.npz")["hex51_A"]
    for tag in ("mi", "ucb"):
        model = orc.BLR(A.shape[0])
        model.m, model.S, model.beta = g[f"{tag}_m"], g[f"{tag}_S"], float(g[f"{tag}_beta"])
        phis = orc.encode(A, np.ones(2), g["xc"])
        mu, var, acq = orc.agent_eval(phis, model, float(g[f"{tag}_var_weight"]), float(g[f"{tag}_gamma_t"]))
        np.testing.assert_allclose(mu, g[f"{tag}_mu"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(var, g[f"{tag}_var"], rtol=1e-10)
        np.testing.assert_allclose(acq, g[f"{tag}_acq"], rtol=1e-9, atol=1e-12)
        score, idx = orc.score_and_argmax(phis, model, float(g[f"{tag}_var_weight"]), float(g[f"{tag}_gamma_t"]))
        assert idx == int(np.argmax(g[f"{tag}_mu"].ravel() + g[f"{tag}_acq"]))
    assert g["ucb_gamma_t"] == 0.0 and g["mi_gamma_t"] > 0.0
    assert g["mi_var_weight"] == 10.0 + 6 * 0.25                    # var_decay is ADDED (ssp_agent.py:208-211)


def test_agent_replay_from_scratch(golden):
    """Replay SSPAgent construction + 6 updates with oracle primitives only."""
    g = golden("agent.npz")
    b = golden("blr.npz")
    A = golden("encode.npz")["hex51_A"]
    ls = np.ones(2)
    for tag, gamma_c in (("mi", 1.0), ("ucb", 0.0)):
        model = orc.BLR(A.shape[0])
        model.update(orc.encode(A, ls, b["xs"]), b["ys"])
        lam, gam = 10.0, 0.0
        from tests.conftest import himmelblau
        for i in range(6):
            x_t = g["xt"][i:i + 1]
            phi = orc.encode(A, ls, x_t)
            _, var_t, _ = orc.agent_eval(phi, model, lam, gam)
            model.update(phi, np.atleast_2d(himmelblau(x_t)))
            lam, gam = orc.agent_update_scalars(lam, gam, 0.25, gamma_c, var_t)
        np.testing.assert_allclose(model.m, g[f"{tag}_m"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(np.ravel(gam)[0], g[f"{tag}_gamma_t"], rtol=1e-12)
        pts = orc.grid_sample_points(B2, 100)                       # ssp_agent.py:117-120
        assert np.array_equal(pts, g[f"{tag}_init_pts"])


def test_trajectory_matches_reference(golden):
    g = golden("traj.npz")
    L, xd = 4, 2
    T = orc.traj_timestep_ssps(g["At"], g["ls_t"], L)
    np.testing.assert_allclose(T, g["timestep_ssps"], atol=1e-15)
    phi = orc.traj_encode(g["Ax"], g["ls_x"], T, g["xc"], L, xd)
    np.testing.assert_allclose(phi, g["phi"], atol=1e-14)
    model = orc.BLR(g["Ax"].shape[0])
    model.m, model.S, model.beta = g["m"], g["S"], float(g["beta"])
    mu, var, acq = orc.agent_eval(phi, model, 5.0, 0.0)
    np.testing.assert_allclose(mu, g["mu"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(acq, g["acq"], rtol=1e-10)
    pts = g["init_pts"]
    ssps = orc.encode(g["Ax"], g["ls_x"], pts)
    dec = orc.traj_decode_from_set(phi[:1], T, ssps, pts)
    assert np.array_equal(dec, g["decoded"])


@pytest.mark.parametrize("tag", ["a3", "a2"])
def test_multi_agent_matches_reference(golden, tag):
    """oracle.multi_agent_encode / decode / sp_vectors against the live reference's SSPMultiAgent
    (agents/ssp_multi_agent.py:166-208); the tag vectors include DC coefficients of both signs."""
    g = golden("multi_agent.npz")
    n_agents, L, xd, ssp_dim, seed = (int(v) for v in g[f"{tag}_dims"])
    d = g[f"{tag}_Ax"].shape[0]
    tags = orc.sp_vectors(n_agents, d, seed)
    np.testing.assert_allclose(tags, g[f"{tag}_agent_sps"], atol=1e-15)
    T = orc.traj_timestep_ssps(g[f"{tag}_At"], g[f"{tag}_ls_t"], L)
    np.testing.assert_allclose(T, g[f"{tag}_timestep_ssps"], atol=1e-15)
    phi = orc.multi_agent_encode(g[f"{tag}_Ax"], g[f"{tag}_ls_x"], T, tags, g[f"{tag}_xc"], L, n_agents, xd)
    np.testing.assert_allclose(phi, g[f"{tag}_phi"], atol=1e-14)
    model = orc.BLR(d)
    model.update(orc.multi_agent_encode(g[f"{tag}_Ax"], g[f"{tag}_ls_x"], T, tags, g[f"{tag}_trajs"], L, n_agents, xd), g[f"{tag}_tys"])
    np.testing.assert_allclose(model.m, g[f"{tag}_m"], rtol=1e-9, atol=1e-12)
    mu, var, acq = orc.agent_eval(phi, model, 5.0, 0.0)
    np.testing.assert_allclose(mu, g[f"{tag}_mu"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(var, g[f"{tag}_var"], rtol=1e-9)
    np.testing.assert_allclose(acq, g[f"{tag}_acq"], rtol=1e-9)
    pts = g[f"{tag}_init_pts"]
    ssps = orc.encode(g[f"{tag}_Ax"], g[f"{tag}_ls_x"], pts)
    dec = orc.multi_agent_decode_from_set(phi[:1], T, tags, ssps, pts)
    assert np.array_equal(dec, g[f"{tag}_decoded"])


@pytest.mark.parametrize("tag", ["s3", "s2"])
def test_multi_agent_per_agent_spaces_match_reference(golden, tag):
    """The same with one RandomSSPSpace per agent on its own rows of domain_bounds (same_agt_x_space=False,
    agents/ssp_multi_agent.py:58-66): phase matrices drawn with rng = seed + i, encode, BLR, eval, per-agent decode."""
    g = golden("multi_agent_sep.npz")
    n_agents, L, xd, ssp_dim, seed = (int(v) for v in g[f"{tag}_dims"])
    As, lss = list(g[f"{tag}_Ax"]), list(g[f"{tag}_ls_x"])
    d = As[0].shape[0]
    for i in range(n_agents):
        assert np.array_equal(orc.random_phase_matrix(xd, ssp_dim, rng=seed + i), As[i])
    assert not np.array_equal(As[0], As[1])
    tags = orc.sp_vectors(n_agents, d, seed)
    np.testing.assert_allclose(tags, g[f"{tag}_agent_sps"], atol=1e-15)
    T = orc.traj_timestep_ssps(g[f"{tag}_At"], g[f"{tag}_ls_t"], L)
    phi = orc.multi_agent_encode(As, lss, T, tags, g[f"{tag}_xc"], L, n_agents, xd)
    np.testing.assert_allclose(phi, g[f"{tag}_phi"], atol=1e-14)
    model = orc.BLR(d)
    model.update(orc.multi_agent_encode(As, lss, T, tags, g[f"{tag}_trajs"], L, n_agents, xd), g[f"{tag}_tys"])
    np.testing.assert_allclose(model.m, g[f"{tag}_m"], rtol=1e-9, atol=1e-12)
    mu, var, acq = orc.agent_eval(phi, model, 5.0, 0.0)
    np.testing.assert_allclose(mu, g[f"{tag}_mu"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(var, g[f"{tag}_var"], rtol=1e-9)
    np.testing.assert_allclose(acq, g[f"{tag}_acq"], rtol=1e-9)
    pts = list(g[f"{tag}_init_pts"])
    ssps = [orc.encode(As[i], lss[i], pts[i]) for i in range(n_agents)]
    dec = orc.multi_agent_decode_from_set(phi[:1], T, tags, ssps, pts)
    assert np.array_equal(dec, g[f"{tag}_decoded"])


def test_multi_agent_lengthscale_cv_objective_matches_reference(golden):
    """oracle.multi_agent_lengthscale_cv_loss against every objective evaluation of the live reference's
    SSPMultiAgent._optimize_lengthscale (ssp_multi_agent.py:131-164; one agent -- the reference's x0 / bounds only agree
    then), recorded by oracle/gen_golden.py."""
    g = golden("multi_agent_sep.npz")
    assert len(g["cv_eval_f"]) > 20
    for x, f in list(zip(g["cv_eval_x"], g["cv_eval_f"]))[::5]:
        v = orc.multi_agent_lengthscale_cv_loss(g["cv_Ax"], g["cv_At"], g["cv_agent_sps"], x, g["cv_trajs"], g["cv_tys"], 4, 1, 2)
        np.testing.assert_allclose(v, f, rtol=1e-10)


def test_lengthscale_cv_objective_matches_reference(golden):
    """oracle.traj_lengthscale_cv_loss against every objective evaluation scipy's L-BFGS-B made inside the live reference's
    SSPTrajectoryAgent._optimize_lengthscale (ssp_traj_agent.py:113-143), recorded by oracle/gen_golden.py."""
    g = golden("lengthscale_cv.npz")
    assert len(g["eval_f"]) > 50
    for x, f in list(zip(g["eval_x"], g["eval_f"]))[::7]:
        v = orc.traj_lengthscale_cv_loss(g["Ax"], g["At"], x, g["trajs"], g["tys"], 4, 2)
        np.testing.assert_allclose(v, f, rtol=1e-11)
    folds = list(orc.kfold_indices(14, 5))
    assert [len(te) for _, te in folds] == [3, 3, 3, 3, 2] and np.array_equal(folds[1][1], [3, 4, 5])


def test_counter_uniform_is_range_consistent():
    a = orc.counter_uniform(0, 0, 1000, 6)
    b = orc.counter_uniform(0, 400, 100, 6)
    assert a.dtype == np.float32 and a.min() >= 0.0 and a.max() < 1.0
    assert np.array_equal(a[400:500], b)
    assert abs(a.mean() - 0.5) < 0.02


def test_decode_direct_optim_matches_reference(golden):
    """oracle.decode_direct_optim (scipy L-BFGS-B on -phi(x).unit(ssp), sspspace.py:358-375) against the live reference's
    decode(method='direct-optim') on noisy SSPs, one of them with an active bound."""
    g = golden("decode_optim.npz")
    from oracle import ssp_oracle as orc
    set_ssps = orc.encode(g["A"], g["ls"], g["set_pts"])
    x0 = orc.decode_from_set(g["ssps"], set_ssps, g["set_pts"])
    assert np.array_equal(x0, g["x0"])
    got = orc.decode_direct_optim(g["A"], g["ls"], g["ssps"], x0, bounds=g["bounds"])
    np.testing.assert_allclose(got, g["decoded"], rtol=0, atol=1e-6)   # finite-difference gradients amplify 1e-16 in phi


def test_acquisition_func_and_lbfgs_route_match_reference(golden):
    """oracle.acquisition_func against the live reference's SSPAgent.acquisition_func (ssp_agent.py:156-185) at probe
    points inside and outside the clamp radius, and the L-BFGS-B restart loop of bayesian_optimization.py:270-287 from the
    reference's own `initial_guess()` draws (which sit far outside the ball, so the reference stops after 0 steps)."""
    g = golden("acq_ascent.npz")
    from oracle import ssp_oracle as orc
    f, grad = orc.acquisition_func(g["m"], g["S"], g["beta"], float(g["var_weight"]), float(g["gamma_t"]))
    for p, pf, pg in zip(g["probe"], g["probe_f"], g["probe_g"]):
        np.testing.assert_allclose(np.ravel(f(p))[0], pf, rtol=1e-12)
        np.testing.assert_allclose(grad(p), pg, rtol=1e-12, atol=1e-12)
    for p, pf in zip(g["phi0"], g["phi0_f"]):
        np.testing.assert_allclose(np.ravel(f(p))[0], pf, rtol=1e-12)
    xs, vals, nit = orc.maximize_acquisition_lbfgs(f, grad, g["phi0"])
    np.testing.assert_allclose(vals, -g["lbfgs_fun"], rtol=1e-9)
    np.testing.assert_array_equal(nit, g["lbfgs_nit"])
    np.testing.assert_allclose(xs, g["lbfgs_x"], rtol=1e-9, atol=1e-9)
    set_ssps = orc.encode(g["A"], g["ls"], g["init_pts"])
    np.testing.assert_array_equal(orc.decode_from_set(xs, set_ssps, g["init_pts"]), g["lbfgs_decoded"])


def test_rff_features_and_blr_golden(golden):
    """agents/rff_agent.py: the feature map and the BLR over it (MI acquisition, additive var_decay), live-reference fixture."""
    g = golden("rff.npz")
    for kt in ("rbf", "matern"):
        W, B, c = g[f"{kt}_W"], g[f"{kt}_B"], float(g[f"{kt}_scale"])
        np.testing.assert_allclose(orc.rff_encode(W, B, c, g["xq"]), g[f"{kt}_phi"], rtol=0, atol=1e-15)
        blr = orc.BLR(W.shape[0])
        blr.update(orc.rff_encode(W, B, c, g["xs"][:12]), g["ys"][:12])
        mu, var, acq = orc.agent_eval(orc.rff_encode(W, B, c, g["xq"]), blr, 5.0, 0.0)
        np.testing.assert_allclose(mu, g[f"{kt}_mu0"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(var, g[f"{kt}_var0"], rtol=1e-9)
        np.testing.assert_allclose(acq, g[f"{kt}_acq0"], rtol=1e-9)
        lam, gamma = 5.0, 0.0
        for i in range(12, 17):
            phi = orc.rff_encode(W, B, c, g["xs"][i:i + 1])
            _, var_t, _ = orc.agent_eval(phi, blr, lam, gamma)
            blr.update(phi, g["ys"][i:i + 1])
            lam, gamma = orc.agent_update_scalars(lam, gamma, 0.25, 1.0, var_t)
        assert np.isclose(lam, float(g[f"{kt}_lam"])) and np.allclose(gamma, g[f"{kt}_gamma"], rtol=1e-9)
        mu, var, acq = orc.agent_eval(orc.rff_encode(W, B, c, g["xq"]), blr, lam, float(np.ravel(gamma)[0]))
        np.testing.assert_allclose(mu, g[f"{kt}_mu1"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(var, g[f"{kt}_var1"], rtol=1e-8)
        np.testing.assert_allclose(acq, g[f"{kt}_acq1"], rtol=1e-8)
        assert np.max(np.abs(blr.m - g[f"{kt}_m"])) <= 1e-9 * np.max(np.abs(g[f"{kt}_m"]))
