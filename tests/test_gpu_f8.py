"""GPU parity tests of the tensor-memory scorer with fp16 + e4m3 operands (csrc/sspbo_score_f8.cu) and of candidate
sets generated inside the scorers: low-rank form up to rank 511 (two chunks), factor form at every rank including the
full-rank posterior after 1000 updates at d = 1023, both against the numpy oracle.
Tolerance as everywhere: mu / var / acquisition within 1e-4 relative, argmax identical where the top-2 gap exceeds it.
"""
import numpy as np
import pytest

from oracle import ssp_oracle as orc
from tests.test_gpu_parity import RTOL_SCORE, _c3_setup, assert_scores_close, pkg  # noqa: F401  (pkg is a fixture)

pytestmark = pytest.mark.gpu


def _mixed_candidates(X, N, seed=5, n_near=60):
    """uniform candidates plus rows at every distance from the observations (remaining prior variance from ~0 to ~1)"""
    rng = np.random.default_rng(seed)
    n = X.shape[1]
    xc = orc.counter_uniform(0, 0, N, n)
    m = min(n_near, X.shape[0])
    for j, scale in enumerate((0.0, 1e-4, 1e-3, 3e-3, 1e-2, 2e-2, 4e-2, 8e-2, 0.15)):
        xc[64 * j:64 * j + m] = np.clip(X[:m] + scale * rng.normal(size=(m, n)), 0, 1).astype(np.float32)
    return xc


def _check_against_oracle(eng, sp, ref, xc, ls, engine, label, pairs=((10.0, 0.0), (4.0, 900.0))):
    phis = orc.encode(sp.phase_matrix, ls, xc.astype(np.float64))
    for lam, gam in pairs:
        ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, lam, gam)
        ref_score = ref_mu.ravel() + ref_acq
        _, (mu, var, acq) = eng.score(xc, lam, gam, want_outputs=True, engine=engine)
        assert eng.last_engine() == label, eng.last_engine()
        mu, var, acq = (t.double().cpu().numpy() for t in (mu, var, acq))
        assert_scores_close(mu, acq, ref_mu, ref_acq)
        np.testing.assert_allclose(var, ref_var, rtol=RTOL_SCORE)
        s, idx = eng.best()
        srt = np.sort(ref_score)[::-1]
        if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
            assert idx == int(np.argmax(ref_score))
        assert abs(s - ref_score.max()) <= RTOL_SCORE * abs(ref_score.max())
    return phis


@pytest.mark.parametrize("T,phase32", [(60, 1), (127, 1), (127, 0), (250, 2), (255, 1)])
def test_lowrank_f8_c3_one_chunk(pkg, T, phase32):
    """Low-rank form with fp16 + e4m3 operands forced at ranks 60 .. 255 (64- to 256-row chunks, both accumulator
    layouts, the three phase paths) on config C3, candidates at every distance from the observations."""
    from ssp_bayesopt_b200 import _lib
    sp, model, ref, X = _c3_setup(pkg, T)
    eng = sp.engine
    assert eng.lowrank_info()["rank"] == T and eng.lowrank_info()["valid"]
    eng.set_policy(phase32=phase32)
    eng.set_operand_format(1)
    xc = _mixed_candidates(X, 6000)
    eng.score_stats(reset=True)
    _check_against_oracle(eng, sp, ref, xc, 0.25 * np.ones(6), _lib.ENGINE_UMMA_LR, "umma-lowrank")
    eng.score(xc, 10.0, 0.0, engine=_lib.ENGINE_UMMA_LR)
    st = eng.score_stats(reset=True)
    assert st["scored"] == 6000 and 60 <= st["flagged"] < 6000 // 4 and st["overflow"] == 0 and st["out_of_range"] == 0, st
    # the two operand formats agree far below the tolerance on the candidates neither flags
    eng.set_operand_format(0)
    _, (mu0, var0, acq0) = eng.score(xc, 10.0, 0.0, want_outputs=True, engine=_lib.ENGINE_UMMA_LR)
    eng.set_operand_format(1)
    _, (mu1, var1, acq1) = eng.score(xc, 10.0, 0.0, want_outputs=True, engine=_lib.ENGINE_UMMA_LR)
    np.testing.assert_allclose(var1.cpu().numpy(), var0.cpu().numpy(), rtol=5e-5)
    eng.set_operand_format(None)
    eng.set_policy(phase32=1)


@pytest.mark.parametrize("T", [256, 300, 511])
def test_lowrank_two_chunks_c3(pkg, T):
    """Ranks 256 .. 511 at d = 1023: two chunks of 144 .. 256 operand rows, psi generated once per chunk; AUTO keeps the
    low-rank form there (fewer operand rows x k-blocks than the triangular factor)."""
    from ssp_bayesopt_b200 import _lib
    sp, model, ref, X = _c3_setup(pkg, T)
    eng = sp.engine
    info = eng.lowrank_info()
    assert info["rank"] == T and info["valid"]
    xc = _mixed_candidates(X, 5000)
    eng.score_stats(reset=True)
    _check_against_oracle(eng, sp, ref, xc, 0.25 * np.ones(6), _lib.ENGINE_UMMA_LR, "umma-lowrank", pairs=((10.0, 0.0),))
    # AUTO: low rank by the cost model; a tenth of this candidate set sits on the observations, which trips the 1 % flag
    # policy at the first read-back, after which AUTO uses the factor form
    eng.set_policy(lowrank=True)
    eng.score_stats(reset=True)
    _, (mu, var, acq) = eng.score(xc, 10.0, 0.0, want_outputs=True)
    assert eng.last_engine() == "umma-lowrank"
    eng.best()
    assert eng.lowrank_info()["disabled"]
    _, (mu2, var2, acq2) = eng.score(xc, 10.0, 0.0, want_outputs=True)
    assert eng.last_engine() == "umma-factor-f8"
    np.testing.assert_allclose(var2.cpu().numpy(), var.cpu().numpy(), rtol=1e-4)
    eng.set_policy(lowrank=True)


@pytest.mark.parametrize("T", [60, 300])
def test_factor_f8_c3(pkg, T):
    """Factor form (var = 1/beta + |R psi|^2, operand = [m'; L^T], four chunks with triangular k-block skipping) at
    d = 1023 against the oracle, including candidates on top of the observations (no float64 pass in this form)."""
    from ssp_bayesopt_b200 import _lib
    sp, model, ref, X = _c3_setup(pkg, T)
    eng = sp.engine
    xc = _mixed_candidates(X, 5000)
    _check_against_oracle(eng, sp, ref, xc, 0.25 * np.ones(6), _lib.ENGINE_UMMA_F8, "umma-factor-f8")
    # the posterior changes: the factor images follow
    x_t = np.random.default_rng(3).uniform(0, 1, (1, 6))
    y_t = orc.hartmann6(x_t).reshape(-1, 1)
    model.update(sp.encode(x_t), y_t)
    phi = orc.encode(sp.phase_matrix, 0.25 * np.ones(6), x_t)[0]
    v = ref.S @ phi
    coef = ref.beta / (1 + ref.beta * (phi @ v))
    bvec = np.linalg.solve(ref.S, ref.m) + ref.beta * phi[:, None] * y_t[0]
    ref.S = ref.S - np.outer(v, v) * coef
    ref.m = ref.S @ bvec
    _check_against_oracle(eng, sp, ref, xc[:2000], 0.25 * np.ones(6), _lib.ENGINE_UMMA_F8, "umma-factor-f8", pairs=((10.0, 0.0),))


def test_full_rank_after_1000_updates_c3(pkg):
    """SURVEY.md section 8d, C3 full-rank case: 1000 observations at d = 1023.  The low-rank rows stop at 511, AUTO hands
    over to the factor form; scores against the oracle's O(d^2) float64 posterior."""
    from ssp_bayesopt_b200 import _lib
    sp, model, ref, X = _c3_setup(pkg, 1000)
    eng = sp.engine
    info = eng.lowrank_info()
    assert not info["valid"]
    rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
    assert rel(model.S, ref.S) < 1e-6 and rel(model.m, ref.m) < 1e-6       # posterior within 1e-6 after 1000 updates
    xc = _mixed_candidates(X, 4000)
    phis = orc.encode(sp.phase_matrix, 0.25 * np.ones(6), xc.astype(np.float64))
    ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, 10.0, 0.0)
    _, (mu, var, acq) = eng.score(xc, 10.0, 0.0, want_outputs=True)              # AUTO
    assert eng.last_engine() == "umma-factor-f8"
    mu, var, acq = (t.double().cpu().numpy() for t in (mu, var, acq))
    assert_scores_close(mu, acq, ref_mu, ref_acq)
    np.testing.assert_allclose(var, ref_var, rtol=RTOL_SCORE)
    ref_score = ref_mu.ravel() + ref_acq
    srt = np.sort(ref_score)[::-1]
    if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
        assert eng.best()[1] == int(np.argmax(ref_score))
    # the older dense-factor kernel (psi staged in shared memory) gives the same numbers
    _, (mu2, var2, acq2) = eng.score(xc, 10.0, 0.0, want_outputs=True, engine=_lib.ENGINE_UMMA)
    np.testing.assert_allclose(var2.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)


@pytest.mark.parametrize("kind,ssp_dim,n_dim,N,T", [("hex", 151, 2, 1000, 25), ("hex", 51, 2, 129, 25), ("rand", 512, 2, 4097, 130),
                                                    ("rand", 64, 3, 128, 25), ("rand", 300, 6, 77, 40), ("rand", 600, 12, 900, 120)])
def test_f8_engine_shapes(pkg, kind, ssp_dim, n_dim, N, T):
    """Other shapes: d = 151 / 25 / 511 / 63 / 299 / 599, 1 .. 10 k-blocks, ragged N, float64 candidates, 12 dimensions, MI
    gamma; both forms of the fp16 + e4m3 kernel against the oracle."""
    from ssp_bayesopt_b200 import _lib
    rng = np.random.default_rng(ssp_dim + N)
    bounds = np.array([[-3.0, 4.0]] * n_dim)
    if kind == "hex":
        sp = pkg.HexagonalSSPSpace(n_dim, ssp_dim=ssp_dim, domain_bounds=bounds, length_scale=0.8, rng=0)
    else:
        sp = pkg.RandomSSPSpace(n_dim, ssp_dim=ssp_dim, domain_bounds=bounds, length_scale=0.8, rng=0)
    ls = 0.8 * np.ones(n_dim)
    X = rng.uniform(-3, 4, (T, n_dim))
    y = np.sin(X.sum(axis=1, keepdims=True))
    model = pkg.BayesianLinearRegression(sp.ssp_dim, engine=sp.engine)
    ref = orc.BLR(sp.ssp_dim)
    for i in range(0, T, 5):
        model.update(sp.encode(X[i:i + 5]), y[i:i + 5])
        ref.update(orc.encode(sp.phase_matrix, ls, X[i:i + 5]), y[i:i + 5])
    xc = rng.uniform(-3, 4, (N, n_dim))                                   # float64 candidates
    xc[:5] = X[:5]                                                         # observed points: smallest variance
    eng = sp.engine
    eng.set_operand_format(1)
    _check_against_oracle(eng, sp, ref, xc, ls, _lib.ENGINE_UMMA_LR, "umma-lowrank", pairs=((10.0, 0.0), (3.0, 250.0)))
    _check_against_oracle(eng, sp, ref, xc, ls, _lib.ENGINE_UMMA_F8, "umma-factor-f8", pairs=((10.0, 0.0), (3.0, 250.0)))
    eng.set_operand_format(None)


@pytest.mark.parametrize("T", [40, 200])
def test_generated_candidates_match_explicit(pkg, T):
    """Candidate sets generated inside the scorer (counter-based uniform stream; n-D grid in the reference's meshgrid
    order) give the same keys and outputs as the same rows passed explicitly - both tensor-memory kernels, the float64
    pass over flagged candidates included, and an engine that needs the rows written out (SIMT)."""
    from ssp_bayesopt_b200 import _lib
    sp, model, ref, X = _c3_setup(pkg, T)
    eng = sp.engine
    N, start = 20000, 123457
    x_exp = eng.fill_uniform(7, start, N)
    for engine in (_lib.ENGINE_UMMA_LR, _lib.ENGINE_UMMA_F8, _lib.ENGINE_SIMT):
        _, (mu, var, acq) = eng.score(x_exp, 10.0, 0.0, index0=start, want_outputs=True, engine=engine)
        key = eng.best()
        _, (mu2, var2, acq2) = eng.score_generated("uniform", N, 10.0, 0.0, index0=start, seed=7, want_outputs=True, engine=engine)
        key2 = eng.best()
        assert key == key2
        for a, b in ((mu, mu2), (var, var2), (acq, acq2)):
            assert (a == b).all()
    # 6-D grid, counts chosen so every axis matters; rows [start, start + N) of the flattened meshgrid
    axes = [np.linspace(0.0, 1.0, c) for c in (7, 5, 6, 4, 9, 8)]
    total = int(np.prod([len(a) for a in axes]))
    gstart, gN = 1000, total - 1000
    pts = eng.grid_points(axes, start=gstart, count=gN)
    mesh = np.array([m.reshape(-1) for m in np.meshgrid(*axes)]).T         # sspspace.py:488-495
    assert (pts.cpu().numpy() == mesh[gstart:]).all()
    for engine in (_lib.ENGINE_UMMA_LR, _lib.ENGINE_UMMA_F8, _lib.ENGINE_AUTO):
        _, (mu, var, acq) = eng.score(pts, 10.0, 0.0, index0=gstart, want_outputs=True, engine=engine)
        key = eng.best()
        _, (mu2, var2, acq2) = eng.score_generated("grid", gN, 10.0, 0.0, index0=gstart, axes=axes, want_outputs=True, engine=engine)
        key2 = eng.best()
        assert key == key2
        np.testing.assert_allclose(var2.cpu().numpy(), var.cpu().numpy(), rtol=2e-6)   # float32 vs float64 candidate rounding
        np.testing.assert_allclose(mu2.cpu().numpy(), mu.cpu().numpy(), rtol=1e-5, atol=1e-5)
    # 2-D grid (x fastest within y, the C2 layout) against the oracle directly
    b2 = np.array([[-5.0, 10.0], [0.0, 15.0]])
    sp2 = pkg.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=b2, length_scale=1.0)
    rng = np.random.default_rng(0)
    X2 = rng.uniform(b2[:, 0], b2[:, 1], (30, 2))
    y2 = np.sin(0.4 * X2[:, :1]) + np.cos(0.3 * X2[:, 1:])
    agt = pkg.SSPAgent(X2, y2, sp2, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0)
    ref2 = orc.BLR(sp2.ssp_dim)
    ref2.update(orc.encode(sp2.phase_matrix, np.ones(2), X2), y2)
    ax2 = [np.linspace(b2[0, 0], b2[0, 1], 101), np.linspace(b2[1, 0], b2[1, 1], 77)]
    pts2 = orc.grid_sample_points(b2, [101, 77])
    ref_mu, ref_var, ref_acq = orc.agent_eval(orc.encode(sp2.phase_matrix, np.ones(2), pts2), ref2, 10.0, 0.0)
    e2 = agt._scoring_engine()
    _, (mu, var, acq) = e2.score_generated("grid", 101 * 77, 10.0, 0.0, axes=ax2, want_outputs=True)
    assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
    assert e2.best()[1] == int(np.argmax(ref_mu.ravel() + ref_acq))


@pytest.mark.parametrize("L,ssp_dim,N,T,ph", [(4, 51, 300, 20, 1), (10, 600, 3000, 150, 1), (10, 600, 3000, 150, 0)])
def test_trajectory_f8(pkg, L, ssp_dim, N, T, ph):
    """Trajectory contexts (sum of L phasors per frequency, |phi|^2 formed by the generators) through both forms of the
    fp16 + e4m3 kernel against the oracle's bind-based encoder (agents/ssp_traj_agent.py:145-163)."""
    from ssp_bayesopt_b200 import _lib
    bounds = np.array([[-2.0, 2.0], [-2.0, 2.0]])
    rng = np.random.default_rng(L + ssp_dim)
    trajs = rng.uniform(-2, 2, (T, 2 * L))
    tys = -np.sum(trajs ** 2, axis=1, keepdims=True) / L
    agt = pkg.SSPTrajectoryAgent(trajs[:12], tys[:12], x_dim=2, traj_len=L, ssp_dim=ssp_dim, domain_bounds=bounds, length_scale=0.7,
                                 time_length_scale=1.3, decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    Ax, At = agt.ssp_x_space.phase_matrix, agt.ssp_t_space.phase_matrix
    Tm = orc.traj_timestep_ssps(At, np.array([1.3]), L)
    enc = lambda x: orc.traj_encode(Ax, 0.7 * np.ones(2), Tm, x, L, 2)
    ref = orc.BLR(Ax.shape[0])
    ref.update(enc(trajs[:12]), tys[:12])
    for i in range(12, T):
        agt.update(trajs[i:i + 1], tys[i:i + 1], 0.0)
        ref.update(enc(trajs[i:i + 1]), tys[i:i + 1])
    xc = rng.uniform(-2, 2, (N, 2 * L)).astype(np.float32)
    xc[:4] = trajs[:4].astype(np.float32)
    phis = enc(xc.astype(np.float64))
    ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, 5.0, 0.0)
    eng = agt._scoring_engine()
    eng.set_operand_format(1)
    eng.set_policy(phase32=ph)
    for engine, label in ((_lib.ENGINE_UMMA_LR, "umma-lowrank"), (_lib.ENGINE_UMMA_F8, "umma-factor-f8")):
        _, (mu, var, acq) = eng.score(xc, 5.0, 0.0, want_outputs=True, engine=engine)
        assert eng.last_engine() == label
        assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
        np.testing.assert_allclose(var.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)
    eng.set_operand_format(None)
    eng.set_policy(phase32=1)
