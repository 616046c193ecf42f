"""GPU parity tests added in round 2: the BASELINE configs at their real shapes (C5: d = 1945, L = 10; C4: d = 511 / 487), the
NaN-candidate policy of the fused scorer, the CTA-pair (cta_group::2) variant of the fp16 + e4m3 scorer, agents sharing one
SSPSpace, float64 eval, the lock-step rerun protocol and the kernel-timing entry points.  Oracle = oracle/ssp_oracle.py (numpy
float64, pinned to the live reference by tests/golden); the unmodified reference from oracle/_ref is used as well where present.
Tolerances as everywhere: scores / variances within 1e-4 relative, argmax identical where the top-2 gap exceeds it."""
import numpy as np
import pytest

from oracle import ssp_oracle as orc
from tests.test_gpu_parity import RTOL_SCORE, _c3_setup, assert_scores_close, pkg, rel  # noqa: F401  (pkg is a fixture)
from tests.test_gpu_f8 import _mixed_candidates

pytestmark = pytest.mark.gpu


def himmelblau(x):
    return -((x[:, 0] ** 2 + x[:, 1] - 11) ** 2 + (x[:, 0] + x[:, 1] ** 2 - 7) ** 2) / 100.0


# ------------------------------------------------------------------------------------------ C5 at its real shape
def test_c5_trajectory_real_shape(pkg):
    """BASELINE config C5: SSPTrajectoryAgent, 10 waypoints, x_dim = 2, ssp_dim = 2048 -> d = 1945 (KC_pad = 1984, 31
    k-blocks), 1e4 uniform trajectories, every engine that takes trajectories (ssp_traj_agent.py:145-176)."""
    from ssp_bayesopt_b200 import _lib
    L, xd, N = 10, 2, 10000
    bounds = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    rng = np.random.default_rng(5)
    trajs = rng.uniform(-5, 5, (14, L * xd))
    tys = -np.sum(trajs ** 2, axis=1, keepdims=True) / 100.0
    agt = pkg.SSPTrajectoryAgent(trajs, tys, x_dim=xd, traj_len=L, ssp_dim=2048, domain_bounds=bounds, length_scale=1.0,
                                 time_length_scale=1.0, decoder_method="from-set", gamma_c=0.0, beta_ucb=10.0)
    Ax, At = agt.ssp_x_space.phase_matrix, agt.ssp_t_space.phase_matrix
    assert Ax.shape[0] == 1945
    T = orc.traj_timestep_ssps(At, np.array([1.0]), L)
    ref = orc.BLR(Ax.shape[0])
    ref.update(orc.traj_encode(Ax, np.ones(xd), T, trajs, L, xd), tys)
    assert rel(agt.blr.m, ref.m) < 1e-8
    xc = rng.uniform(-5, 5, (N, L * xd)).astype(np.float32)
    xc[:6] = trajs[:6].astype(np.float32)                       # on top of observations
    xc[6:12] = (trajs[:6] + 0.02 * rng.normal(size=(6, L * xd))).astype(np.float32)
    phis = orc.traj_encode(Ax, np.ones(xd), T, xc.astype(np.float64), L, xd)
    # K1 at this size: encode of the first rows
    np.testing.assert_allclose(agt.encode(xc[:64].astype(np.float64)), phis[:64], atol=2e-12)
    ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, 10.0, 0.0)
    ref_score = ref_mu.ravel() + ref_acq
    srt = np.sort(ref_score)[::-1]
    eng = agt._scoring_engine()
    for code, name, fmt in ((_lib.ENGINE_UMMA_LR, "umma-lowrank", 0), (_lib.ENGINE_UMMA_LR, "umma-lowrank", 1),
                            (_lib.ENGINE_UMMA_F8, "umma-factor-f8", None), (_lib.ENGINE_UMMA, "umma", None),
                            (_lib.ENGINE_SIMT, "simt", None), (_lib.ENGINE_AUTO, "umma-lowrank", None)):
        eng.set_operand_format(fmt)
        _, (mu, var, acq) = eng.score(xc, 10.0, 0.0, want_outputs=True, engine=code)
        assert eng.last_engine() == name, (name, eng.last_engine())
        assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
        np.testing.assert_allclose(var.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)
        s, idx = eng.best()
        if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
            assert idx == int(np.argmax(ref_score)), (name, idx)
        assert abs(s - ref_score.max()) <= RTOL_SCORE * abs(ref_score.max())
    eng.set_operand_format(None)
    # decode of the winner: unbind every timestep, from-set (ssp_traj_agent.py:165-176)
    dec = agt.decode(phis[int(np.argmax(ref_score))][None, :])
    assert dec.shape == (L * xd,) or dec.shape == (1, L * xd)


# ------------------------------------------------------------------------------------------ C4 at its real dimensions
@pytest.mark.parametrize("kind,d_expect", [("rand", 511), ("hex", 487)])
def test_c4_batched_runs_real_dimension(pkg, kind, d_expect):
    """BASELINE config C4: independent Himmelblau runs at ssp_dim = 512 (-> d = 511 random / 487 hexagonal) over the shared
    100 x 100 grid, stepped in lock-step; every run's pick equals the oracle's arg-max for that run, step by step."""
    bounds = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    R, steps = 5, 3
    agents, refs, spaces = [], [], []
    for r in range(R):
        if kind == "rand":
            sp = pkg.RandomSSPSpace(2, ssp_dim=512, domain_bounds=bounds, length_scale=1.0, rng=r)
        else:
            sp = pkg.HexagonalSSPSpace(2, ssp_dim=512, domain_bounds=bounds, length_scale=1.0 + 0.1 * r)
        assert sp.ssp_dim == d_expect
        xs0 = np.random.default_rng(200 + r).uniform(-5, 5, (10, 2))
        ys0 = himmelblau(xs0).reshape(-1, 1)
        agents.append(pkg.SSPAgent(xs0, ys0, sp, gamma_c=0.0, beta_ucb=10.0, length_scale=np.asarray(sp.length_scale).ravel()[0]))
        ref = orc.BLR(sp.ssp_dim)
        ref.update(orc.encode(sp.phase_matrix, np.asarray(sp.length_scale).ravel(), xs0), ys0)
        refs.append(ref)
        spaces.append(sp)
    grid = orc.grid_sample_points(bounds, 100)
    runs = pkg.BatchedRuns(agents, grid, n_streams=4)
    for _ in range(steps):
        pts, ys, scores = runs.step(himmelblau)
        for r in range(R):
            A, ls = spaces[r].phase_matrix, np.asarray(spaces[r].length_scale).ravel()
            score, idx = orc.score_and_argmax(orc.encode(A, ls, grid), refs[r], 10.0, 0.0)
            srt = np.sort(score)[::-1]
            if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
                assert np.array_equal(pts[r], grid[idx]), (r, pts[r], grid[idx])
            assert abs(scores[r] - score.max()) <= RTOL_SCORE * abs(score.max())
            refs[r].update(orc.encode(A, ls, pts[r:r + 1]), np.atleast_2d(ys[r]))
    for r in range(R):
        assert rel(agents[r].blr.m, refs[r].m) < 1e-8


def test_batched_runs_rerun_protocol(pkg):
    """A lock-step whose candidates leave the declared |x| bound: the per-run status words (sspbo_score_batch_status) make
    BatchedRuns recompute those runs with the any-range phase path, as DeviceEngine.best() does for single runs."""
    bounds = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    R = 3
    agents, refs, spaces = [], [], []
    for r in range(R):
        sp = pkg.RandomSSPSpace(2, ssp_dim=128, domain_bounds=bounds, length_scale=1.0, rng=r)
        xs0 = np.random.default_rng(300 + r).uniform(-5, 5, (8, 2))
        ys0 = himmelblau(xs0).reshape(-1, 1)
        agents.append(pkg.SSPAgent(xs0, ys0, sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0))
        ref = orc.BLR(sp.ssp_dim)
        ref.update(orc.encode(sp.phase_matrix, np.ones(2), xs0), ys0)
        refs.append(ref)
        spaces.append(sp)
    cand = np.random.default_rng(9).uniform(-5, 5, (3000, 2))
    cand[17] = [40.0, -33.0]                                    # far outside the declared bound
    runs = pkg.BatchedRuns(agents, cand, n_streams=2)
    scores, idx, pts = runs.select()
    for r in range(R):
        score, i = orc.score_and_argmax(orc.encode(spaces[r].phase_matrix, np.ones(2), cand), refs[r], 10.0, 0.0)
        assert abs(scores[r] - score.max()) <= RTOL_SCORE * abs(score.max())
        srt = np.sort(score)[::-1]
        if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
            assert idx[r] == i
        assert not agents[r]._scoring_engine().lowrank_info()["phase32"]     # the policy switched to the any-range path


# ------------------------------------------------------------------------------------------ NaN candidates
@pytest.mark.parametrize("engine", ["lowrank", "lowrank-f8", "factor-f8", "exact", "simt", "auto"])
def test_score_nan_candidate_counts_as_maximum(pkg, engine):
    """np.argmax returns the first NaN (sspspace.py:356): a candidate with a NaN coordinate encodes to a
    NaN row in the reference, its score is NaN and it wins.  The integer phase paths of the generators would turn NaN into
    0, so the generators poison psi of such rows; the key orders NaN above every number."""
    from ssp_bayesopt_b200 import _lib
    sp, model, ref, X = _c3_setup(pkg, 40)
    eng = sp.engine
    xc = orc.counter_uniform(0, 0, 3000, 6)
    xc[1234, 3] = np.nan
    xc[2500, 0] = np.nan
    code, fmt, ph = {"lowrank": (_lib.ENGINE_UMMA_LR, 0, 1), "lowrank-f8": (_lib.ENGINE_UMMA_LR, 1, 2),
                     "factor-f8": (_lib.ENGINE_UMMA_F8, None, 1), "exact": (_lib.ENGINE_EXACT, None, 1),
                     "simt": (_lib.ENGINE_SIMT, None, 1), "auto": (_lib.ENGINE_AUTO, None, 1)}[engine]
    eng.set_operand_format(fmt)
    eng.set_policy(phase32=ph)
    eng.score_stats(reset=True)
    _, (mu, var, acq) = eng.score(xc, 10.0, 0.0, want_outputs=True, engine=code)
    s, idx = eng.best()
    assert np.isnan(s) and idx == 1234, (s, idx)
    out = (mu + acq).double().cpu().numpy()
    assert np.isnan(out[1234]) and np.isnan(out[2500])
    fin = np.delete(np.arange(3000), [1234, 2500])
    phis = orc.encode(sp.phase_matrix, 0.25 * np.ones(6), xc[fin].astype(np.float64))
    ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, 10.0, 0.0)
    assert_scores_close(mu.double().cpu().numpy()[fin], acq.double().cpu().numpy()[fin], ref_mu, ref_acq)
    assert eng.score_stats(reset=True)["out_of_range"] == 0      # NaN is not an out-of-bound candidate: no recomputation
    eng.set_operand_format(None)
    eng.set_policy(phase32=1)


# ------------------------------------------------------------------------------------------ CTA pairs
@pytest.mark.parametrize("T,form", [(200, "lr"), (255, "lr"), (400, "lr"), (60, "factor")])
def test_cta_pair_variant_matches_single(pkg, T, form):
    """tcgen05 cta_group::2 variant of the fp16 + e4m3 scorer (operand format 2: tiles of 256 candidates, each CTA of a
    cluster of two holds half of the operand rows): against the oracle, and bit-identical to the single-CTA kernel."""
    from ssp_bayesopt_b200 import _lib
    sp, model, ref, X = _c3_setup(pkg, T)
    eng = sp.engine
    xc = _mixed_candidates(X, 5000 + 77)                       # ragged: the last pair tile is partly empty
    code, label = (_lib.ENGINE_UMMA_LR, "umma-lowrank") if form == "lr" else (_lib.ENGINE_UMMA_F8, "umma-factor-f8")
    phis = orc.encode(sp.phase_matrix, 0.25 * np.ones(6), xc.astype(np.float64))
    ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, 10.0, 0.0)
    outs = {}
    for fmt in (1, 2):
        eng.set_operand_format(fmt)
        eng.score_stats(reset=True)
        _, (mu, var, acq) = eng.score(xc, 10.0, 0.0, want_outputs=True, engine=code)
        assert eng.last_engine() == label
        outs[fmt] = [t.double().cpu().numpy() for t in (mu, var, acq)] + [eng.best()]
        assert_scores_close(outs[fmt][0], outs[fmt][2], ref_mu, ref_acq)
        np.testing.assert_allclose(outs[fmt][1], ref_var, rtol=RTOL_SCORE)
    for a, b in zip(outs[1][:3], outs[2][:3]):
        assert np.array_equal(a, b)
    assert outs[1][3] == outs[2][3]
    eng.set_operand_format(None)


# ------------------------------------------------------------------------------------------ host-layer fixes
def test_agents_sharing_a_space_keep_independent_posteriors(pkg):
    """In the reference every agent owns its BLR and sharing an SSPSpace is legal ('ssp-custom', lists of agents).  Here a
    device context holds one posterior: the second agent over a space gets its own context, and a second raw BLR on an
    engine that hosts a live one is refused."""
    from ssp_bayesopt_b200 import _lib
    bounds = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    sp = pkg.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=bounds, length_scale=1.0)
    rng = np.random.default_rng(0)
    xa, xb = rng.uniform(-5, 5, (8, 2)), rng.uniform(-5, 5, (9, 2))
    ya, yb = himmelblau(xa).reshape(-1, 1), (himmelblau(xb) * 3 + 1).reshape(-1, 1)
    a = pkg.SSPAgent(xa, ya, sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0)
    b = pkg.SSPAgent(xb, yb, sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0)
    assert a._scoring_engine() is not b._scoring_engine()
    refs = []
    for xs, ys in ((xa, ya), (xb, yb)):
        ref = orc.BLR(sp.ssp_dim)
        ref.update(orc.encode(sp.phase_matrix, np.ones(2), xs), ys)
        refs.append(ref)
    x_t = np.array([[1.0, -2.0]])
    b.update(x_t, np.array([[0.5]]), 0.0)
    refs[1].update(orc.encode(sp.phase_matrix, np.ones(2), x_t), np.array([[0.5]]))
    assert rel(a.blr.m, refs[0].m) < 1e-9 and rel(b.blr.m, refs[1].m) < 1e-9
    assert rel(a.blr.S, refs[0].S) < 1e-9 and rel(b.blr.S, refs[1].S) < 1e-9
    grid = orc.grid_sample_points(bounds, 40)
    for agt, ref in ((a, refs[0]), (b, refs[1])):
        mu, var, acq = agt.eval(grid)
        ref_mu, ref_var, ref_acq = orc.agent_eval(orc.encode(sp.phase_matrix, np.ones(2), grid), ref, 10.0, 0.0)
        assert_scores_close(mu.ravel(), acq, ref_mu, ref_acq)
    with pytest.raises(_lib.SspboError):
        pkg.BayesianLinearRegression(sp.ssp_dim, engine=a._scoring_engine())


def test_eval_small_sets_is_float64(pkg, golden):
    """SSPAgent.eval on a handful of points (x_t of a BO step) returns float64-accurate mu / var / acq like the reference, so
    eval -> update(x_t, y_t, var_t) accumulates an exact gamma_t (ssp_agent.py:125-137, 217-218)."""
    sp, model, ref, X = _c3_setup(pkg, 30)
    eng = sp.engine
    xs = np.random.default_rng(4).uniform(0, 1, (7, 6))
    mu, var, acq = eng.eval_host(xs, 10.0, 2.5)
    ref_mu, ref_var, ref_acq = orc.agent_eval(orc.encode(sp.phase_matrix, 0.25 * np.ones(6), xs), ref, 10.0, 2.5)
    np.testing.assert_allclose(mu, ref_mu.ravel(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, ref_var, rtol=1e-9)
    np.testing.assert_allclose(acq, ref_acq, rtol=1e-8)


def test_kernel_timing_entry_points(pkg):
    """sspbo_set_timing / sspbo_timing_read: CUDA events around the scorer kernel and the flagged pass of every score call."""
    sp, model, ref, X = _c3_setup(pkg, 20)
    eng = sp.engine
    xc = eng.fill_uniform(0, 0, 1 << 16)
    eng.score(xc, 10.0, 0.0)
    eng.set_timing(True)
    for _ in range(3):
        eng.score(xc, 10.0, 0.0)
    t = eng.timing_read(reset=True)
    assert t["calls"] == 3 and t["scorer_ms"] > 0 and t["flagged_ms"] >= 0
    assert eng.timing_read()["calls"] == 0
    eng.set_timing(False)
    eng.score(xc, 10.0, 0.0)
    assert eng.timing_read()["calls"] == 0


def test_reference_copy_agrees_with_gpu_when_present(pkg):
    """oracle/_ref (the unmodified reference, copied by oracle/build_ref.py): SSPAgent.eval of the live code against the fused
    scorer on config C3's space at rank 25.  Skipped where the copy is absent."""
    from oracle import ref_loader
    refpkg = ref_loader.load()
    if refpkg is None:
        pytest.skip("oracle/_ref not built")
    import contextlib, io
    b6 = np.array([[0.0, 1.0]] * 6)
    X = np.random.default_rng(1).uniform(0, 1, (25, 6))
    y = orc.hartmann6(X).reshape(-1, 1)
    with contextlib.redirect_stdout(io.StringIO()):
        rsp = refpkg.sspspace.RandomSSPSpace(6, ssp_dim=1024, domain_bounds=b6, length_scale=0.25, rng=0)
        ragt = refpkg.agents.SSPAgent(X[:10], y[:10], rsp, decoder_method="from-set", gamma_c=0.0, beta_ucb=10.0, length_scale=0.25)
        for i in range(10, 25):
            ragt.update(X[i:i + 1], y[i:i + 1], 0.0, step_num=i)
    sp = pkg.RandomSSPSpace(6, ssp_dim=1024, domain_bounds=b6, length_scale=0.25, rng=0)
    assert np.array_equal(sp.phase_matrix, rsp.phase_matrix)
    agt = pkg.SSPAgent(X[:10], y[:10], sp, gamma_c=0.0, beta_ucb=10.0, length_scale=0.25)
    agt.update(X[10:], y[10:], 0.0)
    assert rel(agt.blr.m, ragt.blr.m) < 1e-7 and rel(agt.blr.S, ragt.blr.S) < 1e-7
    xc = orc.counter_uniform(0, 0, 4000, 6)
    rmu, rvar, racq = ragt.eval(xc.astype(np.float64))
    mu, var, acq = agt.eval(xc)
    assert_scores_close(mu.ravel(), acq, rmu, racq)
    np.testing.assert_allclose(var, rvar, rtol=RTOL_SCORE)


# ------------------------------------------------------------------------------------------ encode_and_deriv
@pytest.mark.parametrize("kind,ssp_dim,n", [("hex", 151, 2), ("rand", 1024, 6), ("rand", 64, 3)])
def test_encode_and_deriv(pkg, kind, ssp_dim, n):
    """SSPSpace.encode_and_deriv (sspspace.py:278-306) on the device against the oracle's float64 statement."""
    bounds = np.array([[-4.0, 4.0]] * n)
    ls = np.linspace(0.6, 1.4, n)
    sp = (pkg.HexagonalSSPSpace(n, ssp_dim=ssp_dim, domain_bounds=bounds, length_scale=1.0) if kind == "hex"
          else pkg.RandomSSPSpace(n, ssp_dim=ssp_dim, domain_bounds=bounds, length_scale=1.0, rng=3))
    sp.update_lengthscale(ls)
    for N in (1, 7, 300):
        x = np.random.default_rng(N).uniform(-4, 4, (N, n))
        phi, grad = sp.encode_and_deriv(x)
        rphi, rgrad = orc.encode_and_deriv(sp.phase_matrix, ls, x)
        assert phi.shape == (N, sp.ssp_dim) and grad.shape == (N, sp.ssp_dim, n)
        np.testing.assert_allclose(phi, rphi, atol=2e-13)
        np.testing.assert_allclose(grad, rgrad, atol=2e-12 * max(1.0, np.max(np.abs(rgrad))))
    with pytest.raises(AssertionError):
        sp.encode_and_deriv(np.zeros((2, n + 1)))


# ------------------------------------------------------------------------------------------ RFF baseline agent
@pytest.mark.parametrize("kt", ["rbf", "matern"])
def test_rff_agent_golden(pkg, golden, kt):
    """agents/rff_agent.py on the GPU (feature map + float64 BLR) against the live-reference fixture: the constructor draws
    the reference's features (same sklearn fit, same numpy draws), and with the fixture's features adopted exactly, encode /
    eval / five updates / the x-space objective and gradient agree."""
    g = golden("rff.npz")
    xs, ys, xq = g["xs"], g["ys"], g["xq"]
    np.random.seed(7)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        a0 = pkg.RFFAgent(xs[:12], ys[:12], ssp_dim=96, kernel_type=kt, gamma_c=1.0, beta_ucb=5.0, var_decay=0.25)
    np.testing.assert_allclose(a0.W, g[f"{kt}_W"], rtol=1e-4, atol=1e-7)       # the GP fit may differ in the last digits across hosts
    np.testing.assert_allclose(a0.B, g[f"{kt}_B"], rtol=0, atol=0)
    np.testing.assert_allclose(a0.sqrt_2_alpha_over_m, float(g[f"{kt}_scale"]), rtol=1e-4)
    agt = pkg.RFFAgent(xs[:12], ys[:12], ssp_dim=96, kernel_type=kt, gamma_c=1.0, beta_ucb=5.0, var_decay=0.25,
                       features=(g[f"{kt}_W"], g[f"{kt}_B"], float(g[f"{kt}_scale"])))
    np.testing.assert_allclose(agt.encode(xq), g[f"{kt}_phi"], rtol=0, atol=2e-14)
    mu, var, acq = agt.eval(xq)
    assert mu.shape == (1, 40) and var.shape == (40,) and acq.shape == (40,)
    np.testing.assert_allclose(mu, g[f"{kt}_mu0"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(var, g[f"{kt}_var0"], rtol=1e-8)
    np.testing.assert_allclose(acq, g[f"{kt}_acq0"], rtol=1e-8)
    for i in range(12, 17):
        _, var_t, _ = agt.eval(xs[i:i + 1])
        agt.update(xs[i:i + 1], ys[i:i + 1], var_t, step_num=i)
    mu, var, acq = agt.eval(xq)
    np.testing.assert_allclose(mu, g[f"{kt}_mu1"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(var, g[f"{kt}_var1"], rtol=1e-7)
    np.testing.assert_allclose(acq, g[f"{kt}_acq1"], rtol=1e-7)
    assert rel(agt.blr.m, g[f"{kt}_m"]) < 1e-8 and rel(agt.blr.S, g[f"{kt}_S"]) < 1e-8
    assert np.isclose(agt.var_weight, float(g[f"{kt}_lam"])) and np.allclose(np.ravel(agt.gamma_t), g[f"{kt}_gamma"], rtol=1e-8)
    f, grad = agt.acquisition_func()
    np.testing.assert_allclose(np.array([f(x) for x in xq[:8]]).ravel(), g[f"{kt}_obj"], rtol=1e-7)
    np.testing.assert_allclose(np.array([grad(x) for x in xq[:8]]), g[f"{kt}_grad"], rtol=1e-6, atol=1e-8)


def test_rff_maximize_route(pkg):
    """maximize(agent_type='rff-rbf'): the reference's x-space L-BFGS-B route (bayesian_optimization.py:252-262) with the
    agent's features and posterior on the device."""
    import contextlib, io
    bounds = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    f = lambda x: float(himmelblau(np.atleast_2d(x))[0])
    xs0 = np.random.default_rng(2).uniform(-5, 5, (10, 2))
    ys0 = himmelblau(xs0).reshape(-1, 1)
    opt = pkg.BayesianOptimization(f=f, bounds=bounds, random_state=0)
    with contextlib.redirect_stdout(io.StringIO()):
        opt.maximize(init_points=(xs0, ys0), n_iter=4, num_restarts=3, agent_type="rff-rbf", ssp_dim=64, gamma_c=1.0)
    assert opt.xs.shape == (14, 2) and opt.ys.shape == (14,)
    assert np.all(opt.xs[10:] >= -5) and np.all(opt.xs[10:] <= 5)
    assert isinstance(opt.agt, pkg.RFFAgent) and np.isfinite(opt.max["target"])


def test_batched_runs_use_grouped_kernels(pkg):
    """Config C4: a lock-step of R runs that share a layout is three launches for the selection (grouped scorer over (run,
    tile), grouped float64 pass, status words) and one for all updates -- not several launches per run -- and gives what the
    per-run path gives (posterior and picks compared run by run with independently stepped agents)."""
    bounds = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    R, steps = 7, 4
    def make():
        out = []
        for r in range(R):
            sp = pkg.RandomSSPSpace(2, ssp_dim=256, domain_bounds=bounds, length_scale=1.0, rng=r)
            xs0 = np.random.default_rng(400 + r).uniform(-5, 5, (6, 2))
            out.append(pkg.SSPAgent(xs0, himmelblau(xs0).reshape(-1, 1), sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0))
        return out
    grid = orc.grid_sample_points(bounds, 50)
    a_batch, a_single = make(), make()
    runs = pkg.BatchedRuns(a_batch, grid, n_streams=3)
    engines = [a._scoring_engine() for a in a_batch]
    for _ in range(steps):
        before = sum(e.launch_count() for e in engines)
        pts, ys, scores = runs.step(himmelblau)
        launched = sum(e.launch_count() for e in engines) - before
        assert launched <= 5, launched                          # 4 grouped launches (+ slack), not ~7 per run
        for r, agt in enumerate(a_single):                      # the per-run route on independent agents
            agt.best_candidate(grid)
            s, idx = agt._scoring_engine().best()
            assert abs(s - scores[r]) <= 1e-6 * abs(s)
            assert np.array_equal(grid[idx], pts[r])
            agt.update(grid[idx:idx + 1], np.atleast_2d(himmelblau(grid[idx:idx + 1])), 0.0)
    for a, b in zip(a_batch, a_single):
        assert rel(a.blr.m, b.blr.m) < 1e-10 and rel(a.blr.S, b.blr.S) < 1e-10
        assert a._scoring_engine().lowrank_info() == b._scoring_engine().lowrank_info()


# ------------------------------------------------------------------------------------------ pruning of the float64 pass
@pytest.mark.parametrize("lam,gamma", [(0.01, 0.0), (0.3, 0.0), (10.0, 0.0), (2.0, 50.0)])
@pytest.mark.parametrize("fmt,T", [(0, 60), (1, 200)])
def test_argmax_when_the_winner_is_a_flagged_candidate(pkg, lam, gamma, fmt, T):
    """The float64 pass skips flagged candidates that cannot beat the best key.  With a small variance weight the arg-max IS a
    flagged candidate (right next to the best observation: high mean, almost no variance left), so the pruning must keep it:
    key-only calls (the pruning path) against the oracle's arg-max and against the call with outputs (no pruning)."""
    from ssp_bayesopt_b200 import _lib
    sp, model, ref, X = _c3_setup(pkg, T)
    eng = sp.engine
    eng.set_operand_format(fmt)
    xc = _mixed_candidates(X, 8000, seed=11)
    phis = orc.encode(sp.phase_matrix, 0.25 * np.ones(6), xc.astype(np.float64))
    ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, lam, gamma)
    ref_score = ref_mu.ravel() + ref_acq
    srt = np.argsort(ref_score)[::-1]
    eng.score_stats(reset=True)
    eng.score(xc, lam, gamma, engine=_lib.ENGINE_UMMA_LR)                     # key only: pruned float64 pass
    st = eng.score_stats(reset=True)                                          # (best() below reads and resets the counters itself)
    eng.score(xc, lam, gamma, engine=_lib.ENGINE_UMMA_LR)
    s, idx = eng.best()
    _, (mu, var, acq) = eng.score(xc, lam, gamma, want_outputs=True, engine=_lib.ENGINE_UMMA_LR)
    s2, idx2 = eng.best()
    assert (s, idx) == (s2, idx2)                                             # pruning changes nothing
    if (ref_score[srt[0]] - ref_score[srt[1]]) > RTOL_SCORE * abs(ref_score[srt[0]]):
        assert idx == int(srt[0]), (idx, int(srt[0]))
    assert abs(s - ref_score.max()) <= RTOL_SCORE * abs(ref_score.max())
    assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
    assert st["flagged"] >= 60 and st["overflow"] == 0
    if lam <= 0.01:                                                           # the winner sits among the flagged ones here
        assert ref_var[srt[0]] < 0.3 / 0.01
    eng.set_operand_format(None)


def test_eval_between_chunked_selection_keeps_the_running_key(pkg):
    """SSPAgent.eval on a large set between best_candidate(..., reset_best=False) chunks must not disturb the running key."""
    sp, model, ref, X = _c3_setup(pkg, 20)
    agt_eng = sp.engine
    xc = orc.counter_uniform(0, 0, 6000, 6)
    agt_eng.score(xc[:3000], 10.0, 0.0, index0=0, reset_best=True)
    mu, var, acq = agt_eng.eval_host(xc[100:400], 10.0, 0.0)                 # 300 rows: the tensor route
    agt_eng.score(xc[3000:], 10.0, 0.0, index0=3000, reset_best=False)
    s, idx = agt_eng.best()
    agt_eng.score(xc, 10.0, 0.0)
    s_all, idx_all = agt_eng.best()
    assert (s, idx) == (s_all, idx_all)
    ref_mu, ref_var, ref_acq = orc.agent_eval(orc.encode(sp.phase_matrix, 0.25 * np.ones(6), xc[100:400].astype(np.float64)), ref, 10.0, 0.0)
    assert_scores_close(mu, acq, ref_mu, ref_acq)


# ------------------------------------------------------------------ multi-agent: one x space per agent, length-scale search
@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["s3", "s2"])
def test_multi_agent_per_agent_spaces_golden(pkg, golden, tag):
    """SSPMultiAgent(same_agt_x_space=False): one RandomSSPSpace per agent on its own rows of domain_bounds
    (ssp_multi_agent.py:58-66), one frequency table per agent on the device (sspbo_space_set_classes).  Against the live
    reference (tests/golden/multi_agent_sep.npz): phase matrices, normalised encode, posterior, eval through every engine
    that forms |phi| and both operand formats, per-agent decode."""
    from ssp_bayesopt_b200 import _lib
    g = golden("multi_agent_sep.npz")
    n_agents, L, xd, ssp_dim, seed = (int(v) for v in g[f"{tag}_dims"])
    agt = pkg.SSPMultiAgent(g[f"{tag}_trajs"], g[f"{tag}_tys"], n_agents=n_agents, x_dim=xd, traj_len=L, ssp_dim=ssp_dim,
                            domain_bounds=g[f"{tag}_bounds"], length_scale=0.8, time_length_scale=1.2, seed=seed,
                            same_agt_x_space=False, decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    for i in range(n_agents):
        assert np.array_equal(agt.ssp_x_spaces[i].phase_matrix, g[f"{tag}_Ax"][i])
        assert np.array_equal(np.asarray(agt.ssp_x_spaces[i].domain_bounds), g[f"{tag}_bounds"][i * xd:(i + 1) * xd])
    np.testing.assert_allclose(agt.agent_sps, g[f"{tag}_agent_sps"], atol=1e-15)
    phi = agt.encode(g[f"{tag}_xc"])
    np.testing.assert_allclose(phi, g[f"{tag}_phi"], atol=5e-13)
    assert abs(float(np.ravel(agt.blr.beta)[0]) - float(g[f"{tag}_beta"])) <= 1e-12 * float(g[f"{tag}_beta"])
    assert rel(agt.blr.m, g[f"{tag}_m"]) < 1e-9 and rel(agt.blr.S, g[f"{tag}_S"]) < 1e-9
    mu, var, acq = agt.eval(g[f"{tag}_xc"])                                   # 40 candidates: exact float64 engine
    assert_scores_close(mu, acq, g[f"{tag}_mu"], g[f"{tag}_acq"])
    np.testing.assert_allclose(var, g[f"{tag}_var"], rtol=RTOL_SCORE)
    eng = agt._scoring_engine()
    for code, name, fmt in ((_lib.ENGINE_SIMT, "simt", -1), (_lib.ENGINE_UMMA_LR, "umma-lowrank", 0),
                            (_lib.ENGINE_UMMA_LR, "umma-lowrank", 1)):
        for ph in (1, 0):                                                     # float32 / Q31.32 phase arithmetic
            eng.set_policy(phase32=ph)
            eng.set_operand_format(fmt if fmt >= 0 else None)
            _, (mu, var, acq) = eng.score(g[f"{tag}_xc"], 5.0, 0.0, want_outputs=True, engine=code)
            assert eng.last_engine() == name
            assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), g[f"{tag}_mu"], g[f"{tag}_acq"])
            np.testing.assert_allclose(var.double().cpu().numpy(), g[f"{tag}_var"], rtol=RTOL_SCORE)
    eng.set_policy(phase32=1)
    eng.set_operand_format(None)
    with pytest.raises(_lib.SspboError):                                      # the dense-factor kernel: one table, no |phi|
        eng.score(g[f"{tag}_xc"], 5.0, 0.0, engine=_lib.ENGINE_UMMA)
    assert np.array_equal(agt.decode(phi[:1]), g[f"{tag}_decoded"])


@pytest.mark.gpu
def test_multi_agent_per_agent_spaces_tcgen05_and_driver(pkg):
    """A larger candidate set over per-agent spaces (d = 301, 3 agents x 6 timesteps): AUTO takes the low-rank tcgen05 engine
    in trajectory mode; scores and the arg-max against the oracle restatement (ssp_multi_agent.py:166-186), distinct spaces
    passed as a list; a context re-programmed from shared to per-agent tables; then the BO driver."""
    from ssp_bayesopt_b200 import _lib
    n_agents, L, xd, ssp_dim, N = 3, 6, 2, 301, 3000
    bounds = np.array([[-2.0, 2.0], [-2.0, 2.0], [-1.0, 3.0], [0.0, 2.5], [-3.0, 1.0], [-2.0, 2.0]])
    lo, hi = np.tile(bounds[:, 0], L), np.tile(bounds[:, 1], L)
    rng = np.random.default_rng(23)
    trajs = rng.uniform(lo, hi, (14, L * n_agents * xd))
    tys = -np.sum(trajs ** 2, axis=1, keepdims=True) / (L * n_agents)
    spaces = [pkg.RandomSSPSpace(xd, ssp_dim=ssp_dim, domain_bounds=bounds[i * xd:(i + 1) * xd], length_scale=1, rng=40 + i)
              for i in range(n_agents)]
    agt = pkg.SSPMultiAgent(trajs, tys, n_agents=n_agents, x_dim=xd, traj_len=L, ssp_x_spaces=spaces, domain_bounds=bounds,
                            length_scale=0.7, time_length_scale=1.2, seed=1, decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    As = [sp.phase_matrix for sp in spaces]
    T = orc.traj_timestep_ssps(agt.ssp_t_space.phase_matrix, np.array([1.2]), L)
    tags = orc.sp_vectors(n_agents, As[0].shape[0], 1)
    enc = lambda x: orc.multi_agent_encode(As, [0.7 * np.ones(2)] * n_agents, T, tags, x, L, n_agents, xd)
    np.testing.assert_allclose(agt.encode(trajs[:5]), enc(trajs[:5]), atol=5e-13)
    ref = orc.BLR(As[0].shape[0])
    ref.update(enc(trajs), tys)
    xc = rng.uniform(lo, hi, (N, L * n_agents * xd)).astype(np.float32)
    xc[:4] = trajs[:4].astype(np.float32)
    ref_mu, ref_var, ref_acq = orc.agent_eval(enc(xc.astype(np.float64)), ref, 5.0, 0.0)
    ref_score = ref_mu.ravel() + ref_acq
    eng = agt._scoring_engine()
    for ph in (1, 0):
        for fmt in (0, 1):
            eng.set_policy(phase32=ph)
            eng.set_operand_format(fmt)
            _, (mu, var, acq) = eng.score(xc, 5.0, 0.0, want_outputs=True)
            assert eng.last_engine() == "umma-lowrank"
            assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
            np.testing.assert_allclose(var.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)
            srt = np.sort(ref_score)[::-1]
            if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
                assert eng.best()[1] == int(np.argmax(ref_score))
    eng.set_policy(phase32=1)
    eng.set_operand_format(None)
    # the rows of a candidate are (timestep, agent): swapping two agents' coordinates changes the encoding
    sw = trajs[:3].reshape(3, L, n_agents, xd)[:, :, [1, 0, 2], :].reshape(3, -1)
    assert np.max(np.abs(agt.encode(sw) - agt.encode(trajs[:3]))) > 1e-3
    f = lambda x: float(-np.sum(np.asarray(x) ** 2) / (L * n_agents))
    opt = pkg.BayesianOptimization(f=lambda x, info=None: f(x), bounds=bounds)
    opt.maximize(init_points=(trajs, tys), n_iter=3, agent_type="ssp-multi", n_agents=n_agents, x_dim=xd, traj_len=L,
                 ssp_dim=ssp_dim, length_scale=0.8, time_length_scale=1.2, seed=1, same_agt_x_space=False,
                 decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0, n_candidates=20000)
    assert opt.xs.shape == (17, L * n_agents * xd)
    assert np.all(opt.xs >= lo - 1e-9) and np.all(opt.xs <= hi + 1e-9)


@pytest.mark.gpu
def test_multi_agent_length_scale_search(pkg, golden):
    """SSPMultiAgent without `length_scale`: the K-fold search of ssp_multi_agent.py:131-164 with encodings and Gram matrix
    from the device, against every objective evaluation recorded inside the live reference (one agent, the only case the
    reference's own scipy call accepts); then three agents with their own spaces, objective against the oracle."""
    g = golden("multi_agent_sep.npz")
    agt = pkg.SSPMultiAgent(g["cv_trajs"], g["cv_tys"], n_agents=1, x_dim=2, traj_len=4, ssp_dim=51, domain_bounds=g["cv_bounds"],
                            seed=3, decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    assert np.array_equal(agt.ssp_x_spaces[0].phase_matrix, g["cv_Ax"]) and np.array_equal(agt.ssp_t_space.phase_matrix, g["cv_At"])
    found = agt.length_scale()
    for x, f in list(zip(g["cv_eval_x"], g["cv_eval_f"]))[::5]:
        np.testing.assert_allclose(agt._cv_loss(x, g["cv_trajs"], g["cv_tys"]), f, rtol=1e-9)
    f_found = agt._cv_loss(found, g["cv_trajs"], g["cv_tys"])
    f_ref = agt._cv_loss(g["cv_result"], g["cv_trajs"], g["cv_tys"])
    np.testing.assert_allclose(f_ref, g["cv_eval_f"].min(), rtol=1e-9)
    assert f_found <= f_ref + 2e-3 * abs(f_ref)
    # three agents, own spaces: the objective at a few trial points against the oracle; the agent works afterwards
    n_agents, L, xd = 3, 3, 2
    bounds = np.array([[-2.0, 2.0]] * (n_agents * xd))
    rng = np.random.default_rng(8)
    trajs = rng.uniform(-2, 2, (12, L * n_agents * xd))
    tys = -np.sum(trajs ** 2, axis=1, keepdims=True) / (L * n_agents)
    agt = pkg.SSPMultiAgent(trajs, tys, n_agents=n_agents, x_dim=xd, traj_len=L, ssp_dim=65, domain_bounds=bounds, seed=5,
                            same_agt_x_space=False, decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    ls = agt.length_scale()
    assert ls.shape == (n_agents + 1,) and np.all(ls[:n_agents] == ls[0]) and np.all(ls >= 1 / np.sqrt(12) - 1e-12)
    As = [sp.phase_matrix for sp in agt.ssp_x_spaces]
    for trial in ((4.0, 4.0), (0.9, 2.5), (ls[0], ls[-1])):
        want = orc.multi_agent_lengthscale_cv_loss(As, agt.ssp_t_space.phase_matrix, agt.agent_sps, trial, trajs, tys, L, n_agents, xd)
        np.testing.assert_allclose(agt._cv_loss(trial, trajs, tys), want, rtol=1e-9)
    assert agt._cv_loss((ls[0], ls[-1]), trajs, tys) <= agt._cv_loss((4.0, 4.0), trajs, tys) + 1e-9
    # _cv_loss left the context at the trial scales / timesteps 0 .. L: re-point it as the constructor does
    for sp in agt.ssp_x_spaces:
        sp.update_lengthscale(ls[0])
    agt.ssp_t_space.update_lengthscale(ls[-1])
    agt._point_engine(np.linspace(1, L + 1, L).reshape(-1, 1))
    T = orc.traj_timestep_ssps(agt.ssp_t_space.phase_matrix, np.array([ls[-1]]), L)
    want = orc.multi_agent_encode(As, [ls[0] * np.ones(xd)] * n_agents, T, agt.agent_sps, trajs[:4], L, n_agents, xd)
    np.testing.assert_allclose(agt.encode(trajs[:4]), want, atol=5e-13)


# ------------------------------------------------------------------ low-rank posterior updates: the dense state on demand
@pytest.mark.gpu
def test_low_rank_updates_keep_the_dense_state_on_demand(pkg):
    """Updates in low-rank form (k_blr_update_lr) touch only the rows of V, b_psi and m'; S, S_inv, b, m and Sp are folded in
    when somebody reads them.  Sequences that exercise every hand-over -- more pending observations than the S_inv ring
    holds, reads between updates (predict: S and m; tiny eval: Sp; attribute reads), the unfused kernels in between, a
    multi-point update, export / import -- against the oracle's BLR (blr.py:39-97) at every checkpoint."""
    from ssp_bayesopt_b200 import _lib
    rng = np.random.default_rng(12)
    bounds = np.array([[-2.0, 2.0]] * 3)
    sp = pkg.RandomSSPSpace(3, ssp_dim=129, domain_bounds=bounds, length_scale=0.7, rng=2)
    d = sp.ssp_dim
    X = rng.uniform(-2, 2, (120, 3))
    y = (np.sin(X[:, :1]) - 0.4 * X[:, 1:2] ** 2 + 0.1 * X[:, 2:3])
    phis = orc.encode(sp.phase_matrix, 0.7 * np.ones(3), X)
    agt = pkg.SSPAgent(X[:6], y[:6], sp, length_scale=0.7, decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    eng = agt._scoring_engine()
    ref = orc.BLR(d)
    ref.update(phis[:6], y[:6])

    def check(tag):
        model = agt.blr
        assert rel(model.m, ref.m) < 1e-8, (tag, rel(model.m, ref.m))
        assert rel(model.S, ref.S) < 1e-8, (tag, rel(model.S, ref.S))
        assert rel(model.S_inv, ref.S_inv) < 1e-10, (tag, rel(model.S_inv, ref.S_inv))
        xq = rng.uniform(-2, 2, (5, 3))
        mu, var, _ = agt.eval(xq)                                           # float64 engine over Sp
        rmu, rvar = ref.predict(orc.encode(sp.phase_matrix, 0.7 * np.ones(3), xq))
        np.testing.assert_allclose(np.ravel(mu), np.ravel(rmu), rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(np.ravel(var), np.ravel(rvar), rtol=1e-7)
        pmu, pvar = model.predict(phis[:4])                                  # k_predict over S, m
        rmu, rvar = ref.predict(phis[:4])
        np.testing.assert_allclose(np.ravel(pmu), np.ravel(rmu), rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(np.ravel(pvar), np.ravel(rvar), rtol=1e-7)

    check("initial design")
    i = 6
    for _ in range(40):                                                      # more than the ring of 32, nothing read in between
        agt.update(X[i:i + 1], y[i:i + 1], 0.0); ref.update(phis[i:i + 1], y[i:i + 1]); i += 1
    check("40 single observations")
    for _ in range(3):                                                       # reads between single observations
        mu, var = eng.observe(X[i], float(y[i, 0]))
        rmu, rvar = ref.predict(phis[i:i + 1])
        np.testing.assert_allclose([mu, var], [float(np.ravel(rmu)[0]), float(np.ravel(rvar)[0])], rtol=1e-7, atol=1e-9)
        ref.update(phis[i:i + 1], y[i:i + 1]); i += 1
        check("observe + read")
    assert eng.lib.sspbo_set_update_path(eng._ctx, 0) == 0                   # the separate kernels, which update everything in place
    for _ in range(3):
        agt.update(X[i:i + 1], y[i:i + 1], 0.0); ref.update(phis[i:i + 1], y[i:i + 1]); i += 1
    assert eng.lib.sspbo_set_update_path(eng._ctx, 1) == 0
    for _ in range(5):                                                       # back to the low-rank form: b_psi rebuilt from b
        agt.update(X[i:i + 1], y[i:i + 1], 0.0); ref.update(phis[i:i + 1], y[i:i + 1]); i += 1
    check("unfused kernels in between")
    agt.update(X[i:i + 7], y[i:i + 7], 0.0); ref.update(phis[i:i + 7], y[i:i + 7]); i += 7     # several points in one call
    check("seven points in one call")
    packed = eng.blr_export()                                                # complete state leaves; low-rank rows travel with it
    for _ in range(4):
        agt.update(X[i:i + 1], y[i:i + 1], 0.0)
    eng.blr_import(packed, None)
    for _ in range(4):
        agt.update(X[i:i + 1], y[i:i + 1], 0.0); ref.update(phis[i:i + 1], y[i:i + 1]); i += 1
    check("after import")
    # scoring with every engine still agrees with the reference posterior
    xc = rng.uniform(-2, 2, (3000, 3)).astype(np.float32)
    ref_mu, ref_var, ref_acq = orc.agent_eval(orc.encode(sp.phase_matrix, 0.7 * np.ones(3), xc.astype(np.float64)), ref, 5.0, 0.0)
    for code in (_lib.ENGINE_UMMA_LR, _lib.ENGINE_UMMA, _lib.ENGINE_SIMT, _lib.ENGINE_EXACT):
        _, (mu, var, acq) = eng.score(xc, 5.0, 0.0, want_outputs=True, engine=code)
        assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
        np.testing.assert_allclose(var.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)


@pytest.mark.gpu
def test_batched_runs_grouped_kernels_beyond_rank_127(pkg):
    """The grouped scorer of config C4 keeps serving a lock-step once the runs hold more than 127 observations (one chunk of
    up to 256 split-fp16 operand rows): still a handful of launches per lock-step, same picks as independently stepped agents
    pinned to the same operand format."""
    bounds = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    R, steps, n0 = 5, 3, 150
    def make():
        out = []
        for r in range(R):
            sp = pkg.RandomSSPSpace(2, ssp_dim=512, domain_bounds=bounds, length_scale=1.0, rng=r)
            xs0 = np.random.default_rng(700 + r).uniform(-5, 5, (n0, 2))
            out.append(pkg.SSPAgent(xs0, himmelblau(xs0).reshape(-1, 1), sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0))
        return out
    grid = orc.grid_sample_points(bounds, 60)
    a_batch, a_single = make(), make()
    for a in a_single:
        a._scoring_engine().set_operand_format(0)
    runs = pkg.BatchedRuns(a_batch, grid, n_streams=2)
    engines = [a._scoring_engine() for a in a_batch]
    assert engines[0].lowrank_info()["rank"] == n0
    for _ in range(steps):
        before = sum(e.launch_count() for e in engines)
        pts, ys, scores = runs.step(himmelblau)
        launched = sum(e.launch_count() for e in engines) - before
        assert launched <= 5, launched
        for r, agt in enumerate(a_single):
            agt.best_candidate(grid)
            s, idx = agt._scoring_engine().best()
            assert abs(s - scores[r]) <= 1e-6 * abs(s)
            assert np.array_equal(grid[idx], pts[r])
            agt.update(grid[idx:idx + 1], np.atleast_2d(himmelblau(grid[idx:idx + 1])), 0.0)
    for a, b in zip(a_batch, a_single):
        assert rel(a.blr.m, b.blr.m) < 1e-10 and rel(a.blr.S, b.blr.S) < 1e-10


# ------------------------------------------------------------------ multi-GPU key exchange over peer memory (sspbo_dist.cu)
@pytest.mark.gpu
def test_key_exchange_between_contexts(pkg):
    """The peer-memory winner selection with three 'ranks' living in this process on one GPU (slot arrays handed over as
    pointers, each rank on its own stream so the three exchange kernels are resident together): every rank ends up with the
    maximum key, over several exchanges (both slot buffers), and a rank whose scoring pass asks for a rerun makes every
    rank's read-back report it."""
    import ctypes as C
    import torch
    from ssp_bayesopt_b200 import _lib
    world = 3
    bounds = np.array([[-2.0, 2.0], [-2.0, 2.0]])
    sp = pkg.HexagonalSSPSpace(2, ssp_dim=51, domain_bounds=bounds, length_scale=0.7)
    rng = np.random.default_rng(5)
    X = rng.uniform(-2, 2, (12, 2))
    y = -np.sum(X ** 2, axis=1, keepdims=True)
    agents = [pkg.SSPAgent(X, y, sp, length_scale=0.7, gamma_c=0.0, beta_ucb=5.0) for _ in range(world)]   # same posterior
    engs = [a._scoring_engine() for a in agents]
    lib = engs[0].lib
    streams = [torch.cuda.Stream() for _ in range(world)]
    slots = (C.c_void_p * world)()
    for r, e in enumerate(engs):
        buf = (C.c_ubyte * 64)()
        assert lib.sspbo_dist_init(e._ctx, r, world, buf) == 0
        p = C.c_void_p()
        assert lib.sspbo_dist_slots(e._ctx, C.byref(p)) == 0
        slots[r] = p.value
    for e in engs:
        assert lib.sspbo_dist_connect_ptrs(e._ctx, slots) == 0 and lib.sspbo_dist_ready(e._ctx) == 1
    cand = rng.uniform(-1.9, 1.9, (6000, 2)).astype(np.float32)               # (stays inside the declared |x| <= 2 when shifted below)
    shard = [cand[r * 2000:(r + 1) * 2000] for r in range(world)]
    for step in range(3):                                                      # exchanges 0, 1, 2: both slot buffers
        want = None
        for r, e in enumerate(engs):
            e.score(shard[r] + np.float32(0.01 * step), 5.0, 0.0, index0=r * 2000)
            torch.cuda.synchronize()
            k = int(e._best.item()) & 0xFFFFFFFFFFFFFFFF
            want = k if want is None else max(want, k)
        for r, e in enumerate(engs):                                           # all three kernels in flight together
            with torch.cuda.stream(streams[r]):
                assert lib.sspbo_dist_exchange(e._ctx, C.c_void_p(e._best.data_ptr()), C.c_void_p(streams[r].cuda_stream)) == 0
        torch.cuda.synchronize()
        for e in engs:
            st, key = e._finish()
            assert key == want and not st["rerun"]
    # rank 1 sees candidates outside its declared |x| bound on the fast phase path: everybody is told to rerun
    far = shard[1].copy(); far[0] = [7.0, -9.0]
    for r, e in enumerate(engs):
        e.score(far if r == 1 else shard[r], 5.0, 0.0, index0=r * 2000)
    torch.cuda.synchronize()
    for r, e in enumerate(engs):
        with torch.cuda.stream(streams[r]):
            assert lib.sspbo_dist_exchange(e._ctx, C.c_void_p(e._best.data_ptr()), C.c_void_p(streams[r].cuda_stream)) == 0
    torch.cuda.synchronize()
    flags = [e._finish()[0] for e in engs]
    assert all(f["rerun"] for f in flags) and flags[1]["out_of_range"] >= 1 and flags[0]["out_of_range"] == 0


# ------------------------------------------------------------------ checkpoint / rollback of the low-rank posterior
@pytest.mark.gpu
def test_posterior_checkpoint_and_rollback(pkg):
    """`blr_checkpoint` / `blr_rollback`: observations made after a checkpoint are undone by one small kernel while the
    posterior is in low-rank form -- scores, picks and (after a later read) m, S, S_inv equal those of an agent that never
    saw them -- and rollback declines once the dense state has absorbed one of them."""
    rng = np.random.default_rng(21)
    bounds = np.array([[-2.0, 2.0]] * 3)
    def make():
        sp = pkg.RandomSSPSpace(3, ssp_dim=257, domain_bounds=bounds, length_scale=0.6, rng=4)
        return pkg.SSPAgent(X[:20], y[:20], sp, length_scale=0.6, gamma_c=0.0, beta_ucb=5.0)
    X = rng.uniform(-2, 2, (60, 3))
    y = np.cos(X[:, :1]) - 0.3 * X[:, 1:2] * X[:, 2:3]
    a, ref_agent = make(), make()
    ea, er = a._scoring_engine(), ref_agent._scoring_engine()
    cand = rng.uniform(-2, 2, (4000, 3)).astype(np.float32)
    ea.blr_checkpoint()
    for rounds in range(3):                                                    # fantasise five observations, undo them, repeat
        for i in range(20, 25):
            a.update(X[i:i + 1], y[i:i + 1] + rounds, 0.0)
        assert ea.blr_rank() == 25
        a.best_candidate(cand); s_f, _ = ea.best()
        assert ea.blr_rollback() and ea.blr_rank() == 20
        a.best_candidate(cand); ref_agent.best_candidate(cand)
        assert ea.best() == er.best()                                          # same kernel, same operands: identical key
    # both agents now take the same real observations
    for i in range(25, 30):
        a.update(X[i:i + 1], y[i:i + 1], 0.0); ref_agent.update(X[i:i + 1], y[i:i + 1], 0.0)
    assert rel(a.blr.m, ref_agent.blr.m) < 1e-12 and rel(a.blr.S, ref_agent.blr.S) < 1e-12
    assert rel(a.blr.S_inv, ref_agent.blr.S_inv) < 1e-12
    # the attribute reads above folded observations 20..24 (the real ones) into S: the checkpoint at rank 20 is out of reach
    assert ea.blr_rollback() is False and ea.blr_rank() == 25
    ea.blr_checkpoint()                                                        # a new one at rank 25 works again
    a.update(X[30:31], y[30:31], 0.0)
    assert ea.blr_rollback() and ea.blr_rank() == 25
    a.best_candidate(cand); ref_agent.best_candidate(cand)
    assert ea.best() == er.best()


# ------------------------------------------------------------------ config C2 over its whole run
@pytest.mark.gpu
def test_c2_full_run_stays_on_the_low_rank_form(pkg):
    """BASELINE config C2 as specified -- Branin, hex d = 151, the 1000 x 1000 grid, 250 iterations -- through
    `BayesianOptimization.maximize`: the posterior reaches rank 260 > d, and the engine choice (estimated time of both forms,
    including the factor refresh a changed posterior costs) keeps every step on the low-rank form; the run finds Branin's
    optimum on the grid, and late steps cost about what early ones do (no step pays a Cholesky refresh)."""
    bounds = np.array([[-5.0, 10.0], [0.0, 15.0]])
    a, b, c, r, s, t = 1, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6., 10., 1 / (8 * np.pi)
    branin = lambda x: -(a * (x[:, 1] - b * x[:, 0] ** 2 + c * x[:, 0] - r) ** 2 + s * (1 - t) * np.cos(x[:, 0]) + s)
    xs0 = np.random.default_rng(0).uniform(bounds[:, 0], bounds[:, 1], (10, 2))
    opt = pkg.BayesianOptimization(f=branin, bounds=bounds)
    opt.maximize(init_points=(xs0, branin(xs0).reshape(-1, 1)), n_iter=250, agent_type="ssp-hex", ssp_dim=151,
                 length_scale=1.0, beta_ucb=10.0, gamma_c=0.0, candidates_per_dim=1000)
    eng = opt.agt._scoring_engine()
    info = eng.lowrank_info()
    assert info["rank"] == 260 and info["valid"] and not info["disabled"]
    assert eng.last_engine() == "umma-lowrank"
    assert opt.max["target"] > -0.41                                         # Branin's maxima of -f are -0.3979
    ft = opt.full_times * 1e-6                                               # ms per step
    assert np.median(ft[200:]) < 3.0 * np.median(ft[20:60])
    # the posterior still equals the oracle's after the 260 observations
    ref = orc.BLR(151)
    phis = orc.encode(opt.agt.ssp_space.phase_matrix, np.ones(2), opt.xs)
    ref.update(phis[:10], opt.ys[:10].reshape(-1, 1))
    for i in range(10, len(opt.xs)):
        ref.update(phis[i:i + 1], opt.ys[i:i + 1].reshape(-1, 1))
    assert rel(opt.agt.blr.m, ref.m) < 1e-6 and rel(opt.agt.blr.S, ref.S) < 1e-6
