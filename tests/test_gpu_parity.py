"""GPU parity tests: the CUDA path (through the C ABI, via the package's reference-shaped API) against the
numpy oracle and the committed golden fixtures of the live reference.

Tolerances (BASELINE.json north_star): phi and acquisition scores within 1e-4 relative (phi: normwise per
row); argmax identical wherever the oracle's top-2 relative gap exceeds 1e-4; posterior m, S within 1e-6
relative (normwise) after 1,000 updates.  Index / integer results (grid order, from-set decode indices,
the counter-based generator) are bit-exact.
"""
import os

import numpy as np
import pytest

from oracle import ssp_oracle as orc
from tests.conftest import himmelblau

pytestmark = pytest.mark.gpu

B2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
RTOL_SCORE = 1e-4


@pytest.fixture(scope="module")
def pkg():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ssp_bayesopt_b200 as p
    return p


def space_from(pkg, A, ls, bounds=None):
    return pkg.SSPSpace(A.shape[1], A.shape[0], phase_matrix=A, domain_bounds=bounds, length_scale=np.asarray(ls))


def assert_scores_close(mu, acq, ref_mu, ref_acq, rtol=RTOL_SCORE):
    """score = mu + acq can cancel to ~0 (MI acquisition), so 'within 1e-4 relative' is taken relative to the
    magnitude of the two terms; acq (always > 0) is additionally checked element-wise."""
    mu, acq, ref_mu, ref_acq = (np.ravel(a) for a in (mu, acq, ref_mu, ref_acq))
    scale = np.abs(ref_mu) + np.abs(ref_acq)
    err = np.abs((mu + acq) - (ref_mu + ref_acq)) / scale
    assert err.max() < rtol, err.max()
    np.testing.assert_allclose(acq, ref_acq, rtol=rtol)
    assert np.max(np.abs(mu - ref_mu) / scale) < rtol


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


# ------------------------------------------------------------------------------------------ K1 encode
@pytest.mark.parametrize("name,xk", [("hex51", "x2"), ("rand51", "x2"), ("hex151", "x2"), ("rand512", "x2"),
                                     ("hexsmall", "x15"), ("rand1024", "x6")])
def test_encode_golden(pkg, golden, name, xk):
    g = golden("encode.npz")
    sp = space_from(pkg, g[f"{name}_A"], g[f"{name}_ls"])
    phi = sp.encode(g[xk])
    assert phi.dtype == np.float64 and phi.shape == g[f"{name}_phi"].shape
    np.testing.assert_allclose(phi, g[f"{name}_phi"], rtol=0, atol=2e-13)
    np.testing.assert_allclose(np.linalg.norm(phi, axis=1), 1.0, atol=1e-12)      # tests/test_ssps.py:31-37


def test_encode_edge_cases(pkg, golden):
    g = golden("encode.npz")
    sp = space_from(pkg, g["rand51_A"], g["rand51_ls"])
    assert sp.encode(np.zeros((0, 2))).shape == (0, 51)                            # empty input
    one = sp.encode(np.array([1.3, -3.4]))                                          # 1-D input -> (1, d)
    assert one.shape == (1, 51)
    np.testing.assert_allclose(one, orc.encode(g["rand51_A"], g["rand51_ls"], [[1.3, -3.4]]), atol=2e-13)
    np.testing.assert_allclose(sp.encode(np.zeros((3, 2)))[:, 0], 1.0, atol=1e-15)  # phi(0) = identity
    x32 = g["x2"].astype(np.float32)                                                # float32 candidates accepted
    np.testing.assert_allclose(sp.encode(x32), orc.encode(g["rand51_A"], g["rand51_ls"], x32.astype(np.float64)),
                               atol=2e-13)
    with pytest.raises(AssertionError):
        sp.encode(np.zeros((2, 3)))
    # ragged batch sizes around the kernel's candidates-per-block
    for n in (1, 3, 4, 5, 1023):
        x = np.random.default_rng(n).uniform(-5, 5, (n, 2))
        np.testing.assert_allclose(sp.encode(x), orc.encode(g["rand51_A"], g["rand51_ls"], x), atol=2e-13)


def test_encode_large_random(pkg, golden):
    g = golden("encode.npz")
    A, ls = g["rand1024_A"], g["rand1024_ls"]
    sp = space_from(pkg, A, ls)
    x = orc.counter_uniform(3, 0, 4096, 6).astype(np.float64)
    phi, ref = sp.encode(x), orc.encode(A, ls, x)
    err = np.max(np.abs(phi - ref), axis=1) / np.max(np.abs(ref), axis=1)
    assert err.max() < 1e-4 and err.max() < 1e-10, err.max()


@pytest.mark.parametrize("name,xk,N", [("rand1024", "x6", 5000), ("hex151", "x2", 1000), ("rand51", "x2", 129)])
def test_encode_tensor_core_f32(pkg, golden, name, xk, N):
    """encode(x, dtype=float32): phi = psi B^T on the tensor cores (split-fp16 basis), within the 1e-4 the path is
    specified to (normwise per row) - in fact ~1e-6 - with and without a BLR on the context, float32 / float64 input."""
    g = golden("encode.npz")
    A, ls = g[f"{name}_A"], g[f"{name}_ls"]
    n = A.shape[1]
    bounds = np.array([[-6.0, 6.0]] * n)
    sp = space_from(pkg, A, ls, bounds)
    rng = np.random.default_rng(N)
    x = rng.uniform(-5, 5, (N, n)) if n == 2 else orc.counter_uniform(4, 0, N, n).astype(np.float64)
    ref = orc.encode(A, ls, x)
    for xin in (x, x.astype(np.float32)):
        phi = sp.encode(xin, dtype=np.float32)
        assert phi.dtype == np.float32 and phi.shape == ref.shape
        r = orc.encode(A, ls, xin.astype(np.float64))
        err = np.max(np.abs(phi - r), axis=1) / np.max(np.abs(r), axis=1)
        assert err.max() < 1e-4 and err.max() < 2e-5, err.max()
    np.testing.assert_allclose(np.linalg.norm(phi, axis=1), 1.0, atol=1e-5)
    assert sp.encode(np.zeros((0, n)), dtype=np.float32).shape == (0, A.shape[0])
    # the same context keeps scoring correctly afterwards (the encoder shares the kernel and its operand geometry)
    model = pkg.BayesianLinearRegression(sp.ssp_dim, engine=sp.engine)
    model.update(sp.encode(x[:8]), np.sin(x[:8].sum(axis=1, keepdims=True)))
    phi2 = sp.encode(x[:64], dtype=np.float32)
    assert np.max(np.abs(phi2 - ref[:64])) < 2e-5 * np.max(np.abs(ref[:64]))


# ------------------------------------------------------------------------------------------ K3 BLR
@pytest.mark.parametrize("name", ["hex51", "rand51"])
def test_blr_golden(pkg, golden, name):
    g, e = golden("blr.npz"), golden("encode.npz")
    sp = space_from(pkg, e[f"{name}_A"], e[f"{name}_ls"])
    model = pkg.BayesianLinearRegression(sp.ssp_dim, engine=sp.engine)
    model.update(sp.encode(g["xs"]), g["ys"])
    assert model.beta == g[f"{name}_beta"]
    assert model.m.shape == (sp.ssp_dim, 1) and model.S.shape == (sp.ssp_dim, sp.ssp_dim)
    assert rel(model.m, g[f"{name}_m0"]) < 1e-11 and rel(model.S, g[f"{name}_S0"]) < 1e-11
    for i in range(40):
        model.update(sp.encode(g[f"{name}_xt"][i:i + 1]), g[f"{name}_yt"][i:i + 1])
    assert rel(model.m, g[f"{name}_m40"]) < 1e-9 and rel(model.S, g[f"{name}_S40"]) < 1e-9
    mu, var = model.predict(sp.encode(g[f"{name}_xc"]))
    assert mu.shape == (1, 64) and var.shape == (64,)                               # tests/test_blr.py:28-39
    np.testing.assert_allclose(mu, g[f"{name}_mu"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(var, g[f"{name}_var"], rtol=1e-8)
    assert np.all(var > 0)                                                          # tests/test_blr.py:42-52
    Sinv = model.S_inv
    np.testing.assert_allclose(Sinv @ model.S, np.eye(sp.ssp_dim), atol=1e-7)


def test_blr_single_target_beta_rule(pkg, golden):
    g, e = golden("blr.npz"), golden("encode.npz")
    sp = space_from(pkg, e["hex51_A"], e["hex51_ls"])
    model = pkg.BayesianLinearRegression(sp.ssp_dim, engine=sp.engine)
    model.update(sp.encode(g["xs"][:1]), g["ys"][:1])
    assert np.ravel(model.beta)[0] == g["single_beta"]                              # 1/|t| (blr.py:57-58)
    assert rel(model.S, g["single_S"]) < 1e-11 and rel(model.m, g["single_m"]) < 1e-11


def test_blr_1000_updates(pkg, golden):
    """north star: posterior m and S within 1e-6 relative after 1,000 updates."""
    g = golden("blr1000.npz")
    sp = space_from(pkg, g["A"], np.ones(2))
    model = pkg.BayesianLinearRegression(sp.ssp_dim, engine=sp.engine)
    phis = sp.encode(g["xt"])
    model.update(phis[:10], g["yt"][:10])
    for i in range(10, 1000):
        model.update(phis[i:i + 1], g["yt"][i:i + 1])
    assert model.beta == g["beta"]
    assert rel(model.S, g["S"]) < 1e-6 and rel(model.m, g["m"]) < 1e-6, (rel(model.S, g["S"]), rel(model.m, g["m"]))


def test_blr_raw_features_sequential_equals_batch(pkg):
    """The reference's own invariant (tests/test_blr.py:72-92) on a BLR without an SSP space (even dim)."""
    rng = np.random.default_rng(4)
    xs, ys = rng.normal(size=(6, 10)), rng.normal(size=(6, 1))
    a = pkg.BayesianLinearRegression(10, beta=1.0)
    b = pkg.BayesianLinearRegression(10, beta=1.0)
    a.update(xs, ys)
    for i in range(6):
        b.update(xs[i:i + 1], ys[i:i + 1])
    assert np.allclose(a.m, b.m, atol=1e-6) and np.allclose(a.S, b.S, atol=1e-6)
    ref = orc.BLR(10, beta=1.0)
    ref.update(xs, ys)
    assert rel(a.m, ref.m) < 1e-10 and rel(a.S, ref.S) < 1e-10
    mu, var = a.predict(xs)
    rmu, rvar = ref.predict(xs)
    np.testing.assert_allclose(mu, rmu, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(var, rvar, rtol=1e-9)
    with pytest.raises(AssertionError):
        a.update(np.zeros((2, 9)), np.zeros((2, 1)))
    with pytest.raises(AssertionError):
        a.update(np.zeros((2, 10)), np.zeros(2))


# ------------------------------------------------------------------------------------------ K2 fused scorer
def _replay_agent(pkg, golden, tag, gamma_c):
    g, b = golden("agent.npz"), golden("blr.npz")
    sp = pkg.HexagonalSSPSpace(2, ssp_dim=51, domain_bounds=B2, length_scale=1.0, rng=0)
    agt = pkg.SSPAgent(b["xs"], b["ys"], sp, decoder_method="from-set", gamma_c=gamma_c, beta_ucb=10.0,
                       var_decay=0.25, length_scale=1.0)
    for i in range(6):
        x_t = g["xt"][i:i + 1]
        _, var_t, _ = agt.eval(x_t)
        agt.update(x_t, np.atleast_2d(himmelblau(x_t)), var_t, step_num=i)
    return agt, g


@pytest.mark.parametrize("tag,gamma_c", [("mi", 1.0), ("ucb", 0.0)])
def test_agent_eval_golden(pkg, golden, tag, gamma_c):
    agt, g = _replay_agent(pkg, golden, tag, gamma_c)
    assert rel(agt.blr.m, g[f"{tag}_m"]) < 1e-7
    np.testing.assert_allclose(np.ravel(agt.gamma_t)[0], g[f"{tag}_gamma_t"], rtol=1e-5)
    assert agt.var_weight == g[f"{tag}_var_weight"]
    mu, var, acq = agt.eval(g["xc"])
    assert mu.shape == (1, 256) and var.shape == (256,) and acq.shape == (256,)
    ref = g[f"{tag}_mu"].ravel() + g[f"{tag}_acq"]
    assert_scores_close(mu, acq, g[f"{tag}_mu"], g[f"{tag}_acq"])
    np.testing.assert_allclose(var, g[f"{tag}_var"], rtol=RTOL_SCORE)
    key = agt.best_candidate(g["xc"])
    s, idx = agt._scoring_engine().best()
    srt = np.sort(ref)[::-1]
    if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
        assert idx == int(np.argmax(ref))
    assert abs(s - ref.max()) <= RTOL_SCORE * np.max(np.abs(g[f"{tag}_mu"]) + np.abs(g[f"{tag}_acq"]))
    assert np.array_equal(agt.init_samples[1], g[f"{tag}_init_pts"])
    assert np.array_equal(agt.decode(agt.encode(g["xc"][:3])), g[f"{tag}_decoded"])


def _c3_setup(pkg, T, seed=1):
    """BASELINE config C3 space (Random n=6, ssp_dim=1024 -> d=1023, l=0.25) with T Hartmann-6 observations."""
    b6 = np.array([[0.0, 1.0]] * 6)
    sp = pkg.RandomSSPSpace(6, ssp_dim=1024, domain_bounds=b6, length_scale=0.25, rng=0)
    X = np.random.default_rng(seed).uniform(0, 1, (T, 6))
    y = orc.hartmann6(X).reshape(-1, 1)
    model = pkg.BayesianLinearRegression(sp.ssp_dim, engine=sp.engine)
    phis = sp.encode(X)
    model.update(phis[:10], y[:10])
    ref = orc.BLR(sp.ssp_dim)
    ref.update(orc.encode(sp.phase_matrix, 0.25 * np.ones(6), X[:10]), y[:10])
    beta = ref.beta
    S, bvec = ref.S.copy(), beta * (phis[:10].T @ y[:10])
    for i in range(10, T):                         # O(d^2) float64 restatement of blr.py:63-75 (survey section 7)
        model.update(phis[i:i + 1], y[i:i + 1])
        phi = phis[i]
        v = S @ phi
        S = S - np.outer(v, v) * (beta / (1 + beta * (phi @ v)))
        bvec = bvec + beta * phi[:, None] * y[i]
    ref.S, ref.m = S, S @ bvec
    return sp, model, ref, X


ENGINE_LABEL = {"simt": "simt", "umma": "umma", "lowrank": "umma-lowrank", "auto": "umma-lowrank", "exact": "exact"}


@pytest.mark.parametrize("engine", ["simt", "umma", "lowrank", "auto", "exact"])
def test_score_c3_parity(pkg, engine):
    from ssp_bayesopt_b200 import _lib
    sp, model, ref, X = _c3_setup(pkg, 60)
    assert rel(model.S, ref.S) < 1e-9 and rel(model.m, ref.m) < 1e-7
    N = 8192
    xc = orc.counter_uniform(0, 0, N, 6)
    xc[:32] = (X[:32] + 1e-3 * np.random.default_rng(2).normal(size=(32, 6))).astype(np.float32)  # low-variance rows
    lam, gam = 10.0, 0.0
    phis = orc.encode(sp.phase_matrix, 0.25 * np.ones(6), xc.astype(np.float64))
    ref_score, ref_idx = orc.score_and_argmax(phis, ref, lam, gam)
    eng = sp.engine
    code = {"simt": _lib.ENGINE_SIMT, "umma": _lib.ENGINE_UMMA, "auto": _lib.ENGINE_AUTO, "lowrank": _lib.ENGINE_UMMA_LR,
            "exact": _lib.ENGINE_EXACT}[engine]
    _, (mu, var, acq) = eng.score(xc, lam, gam, want_outputs=True, engine=code)
    assert eng.last_engine() == ENGINE_LABEL[engine]
    score = (mu + acq).double().cpu().numpy()
    ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, lam, gam)
    assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
    np.testing.assert_allclose(var.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)
    err = np.abs(score - ref_score) / np.abs(ref_score)
    assert err.max() < RTOL_SCORE, err.max()
    srt = np.sort(ref_score)[::-1]
    gap = (srt[0] - srt[1]) / abs(srt[0])
    s, idx = eng.best()
    if gap > RTOL_SCORE:
        assert idx == ref_idx, (idx, ref_idx, gap)
    assert abs(s - ref_score.max()) <= RTOL_SCORE * abs(ref_score.max())
    # index0 offsets the packed index; shards combine by max
    eng.score(xc[:4096], lam, gam, index0=1000, engine=code)
    k1 = int(eng._best.item())
    eng.score(xc[4096:], lam, gam, index0=1000 + 4096, reset_best=False, engine=code)
    s2, idx2 = eng.best()
    assert idx2 == idx + 1000 and s2 == s
    assert (int(eng._best.item()) & 0xFFFFFFFFFFFFFFFF) >= (k1 & 0xFFFFFFFFFFFFFFFF)


@pytest.mark.parametrize("kind,ssp_dim,n_dim,N", [("hex", 151, 2, 1000), ("hex", 51, 2, 129), ("rand", 512, 2, 4097),
                                                  ("rand", 64, 3, 128), ("rand", 300, 6, 77)])
def test_score_umma_shapes(pkg, kind, ssp_dim, n_dim, N):
    """tcgen05 engine on the other BASELINE shapes (d=151, d=511), odd d, ragged N, float64 candidates, MI gamma."""
    from ssp_bayesopt_b200 import _lib
    rng = np.random.default_rng(ssp_dim + N)
    bounds = np.array([[-3.0, 4.0]] * n_dim)
    if kind == "hex":
        sp = pkg.HexagonalSSPSpace(n_dim, ssp_dim=ssp_dim, domain_bounds=bounds, length_scale=0.8, rng=0)
    else:
        sp = pkg.RandomSSPSpace(n_dim, ssp_dim=ssp_dim, domain_bounds=bounds, length_scale=0.8, rng=0)
    ls = 0.8 * np.ones(n_dim)
    X = rng.uniform(-3, 4, (25, n_dim))
    y = np.sin(X.sum(axis=1, keepdims=True))
    model = pkg.BayesianLinearRegression(sp.ssp_dim, engine=sp.engine)
    ref = orc.BLR(sp.ssp_dim)
    for i in range(0, 25, 5):
        model.update(sp.encode(X[i:i + 5]), y[i:i + 5])
        ref.update(orc.encode(sp.phase_matrix, ls, X[i:i + 5]), y[i:i + 5])
    xc = rng.uniform(-3, 4, (N, n_dim))                                   # float64 candidates
    xc[:5] = X[:5]                                                         # observed points: smallest variance
    phis = orc.encode(sp.phase_matrix, ls, xc)
    for lam, gam in ((10.0, 0.0), (3.0, 250.0)):
        ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, lam, gam)
        ref_score = ref_mu.ravel() + ref_acq
        eng = sp.engine
        _, (mu, var, acq) = eng.score(xc, lam, gam, want_outputs=True, engine=_lib.ENGINE_UMMA)
        assert eng.last_engine() == "umma"
        assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
        np.testing.assert_allclose(var.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)
        s, idx = eng.best()
        srt = np.sort(ref_score)[::-1]
        if N > 1 and (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
            assert idx == int(np.argmax(ref_score))
        # both engines agree far below the tolerance
        _, (mu2, var2, acq2) = eng.score(xc, lam, gam, want_outputs=True, engine=_lib.ENGINE_SIMT)
        np.testing.assert_allclose(var.cpu().numpy(), var2.cpu().numpy(), rtol=2e-5)
        # low-rank engine (rank 25, one to sixteen k-blocks, both table layouts) + its float64 pass on the observed rows
        assert eng.lowrank_info()["rank"] == 25 and eng.lowrank_info()["valid"]
        _, (mu3, var3, acq3) = eng.score(xc, lam, gam, want_outputs=True, engine=_lib.ENGINE_UMMA_LR)
        assert eng.last_engine() == "umma-lowrank"
        assert_scores_close(mu3.double().cpu().numpy(), acq3.double().cpu().numpy(), ref_mu, ref_acq)
        np.testing.assert_allclose(var3.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)
        s3, idx3 = eng.best()
        if N > 1 and (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
            assert idx3 == int(np.argmax(ref_score))


@pytest.mark.parametrize("T,phase32", [(60, 1), (60, 0), (60, 2), (127, 1), (250, 1)])
def test_score_lowrank_c3(pkg, T, phase32):
    """Low-rank tcgen05 engine (S = I/alpha - V^T V, one row per observation) on config C3 at ranks 60 / 127 / 250
    (64-, 128- and 256-column accumulators), with candidates placed at every distance from the observations so that the
    remaining prior variance spans (0, 1): the ones below the flag threshold must come back from the exact float64
    pass, all within the 1e-4 tolerance; all three phase paths (float32 chain, 32-bit fixed point, Q31.32)."""
    from ssp_bayesopt_b200 import _lib
    sp, model, ref, X = _c3_setup(pkg, T)
    eng = sp.engine
    info = eng.lowrank_info()
    assert info["rank"] == T and info["valid"] and info["phase32"]
    eng.set_policy(phase32=phase32)
    N = 6000
    rng = np.random.default_rng(5)
    xc = orc.counter_uniform(0, 0, N, 6)
    for j, scale in enumerate((0.0, 1e-4, 1e-3, 3e-3, 1e-2, 2e-2, 4e-2, 8e-2, 0.15)):
        xc[64 * j:64 * j + 60] = np.clip(X[:60] + scale * rng.normal(size=(60, 6)), 0, 1).astype(np.float32)
    phis = orc.encode(sp.phase_matrix, 0.25 * np.ones(6), xc.astype(np.float64))
    eng.score_stats(reset=True)
    for lam, gam in ((10.0, 0.0), (4.0, 900.0)):
        ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, lam, gam)
        ref_score = ref_mu.ravel() + ref_acq
        _, (mu, var, acq) = eng.score(xc, lam, gam, want_outputs=True, engine=_lib.ENGINE_UMMA_LR)
        assert eng.last_engine() == "umma-lowrank"
        mu, var, acq = (t.double().cpu().numpy() for t in (mu, var, acq))
        assert_scores_close(mu, acq, ref_mu, ref_acq)
        np.testing.assert_allclose(var, ref_var, rtol=RTOL_SCORE)
        s, idx = eng.best()
        srt = np.sort(ref_score)[::-1]
        if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
            assert idx == int(np.argmax(ref_score))
    # the exact pass was exercised (observed points have var << prior) but stayed a small minority
    eng.score(xc, 10.0, 0.0, engine=_lib.ENGINE_UMMA_LR)
    st = eng.score_stats(reset=True)
    assert st["scored"] == N and 60 <= st["flagged"] < N // 4 and st["overflow"] == 0 and st["out_of_range"] == 0, st
    frac_prior = (ref_var - 1.0 / ref.beta) * 0.01                           # remaining fraction of the prior variance
    assert frac_prior.min() < 1e-3 and frac_prior.max() > 0.5
    eng.set_policy(phase32=1)


@pytest.mark.parametrize("T", [100, 140, 210])
def test_score_lowrank_beyond_half_rank_small_space(pkg, T):
    """AUTO keeps the low-rank engine up to rank 1.5 d for d >= 128: hex d = 151 (3 k-blocks per tile) at ranks above d / 2,
    between d / 2 and d, and above d (rows of V then are linearly dependent, the form stays exact) - accumulator layouts
    of 112, 144 and 224 rows.  Scores against the oracle on random candidates and on candidates near the observations."""
    bounds = np.array([[-5.0, 10.0], [0.0, 15.0]])
    sp = pkg.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=bounds, length_scale=1.0)
    rng = np.random.default_rng(T)
    X = rng.uniform(bounds[:, 0], bounds[:, 1], (T, 2))
    y = np.sin(0.4 * X[:, :1]) + np.cos(0.3 * X[:, 1:])
    agt = pkg.SSPAgent(X[:10], y[:10], sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0)
    ref = orc.BLR(sp.ssp_dim)
    ref.update(orc.encode(sp.phase_matrix, np.ones(2), X[:10]), y[:10])
    for i in range(10, T):
        agt.update(X[i:i + 1], y[i:i + 1], 0.0)
        ref.update(orc.encode(sp.phase_matrix, np.ones(2), X[i:i + 1]), y[i:i + 1])
    eng = agt._scoring_engine()
    assert eng.lowrank_info()["rank"] == T and eng.lowrank_info()["valid"]
    N = 5000
    xc = rng.uniform(bounds[:, 0], bounds[:, 1], (N, 2))
    xc[:40] = X[:40] + 0.02 * rng.normal(size=(40, 2))
    xc = np.clip(xc, bounds[:, 0], bounds[:, 1]).astype(np.float32)
    phis = orc.encode(sp.phase_matrix, np.ones(2), xc.astype(np.float64))
    ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, 10.0, 0.0)
    eng.score_stats(reset=True)
    _, (mu, var, acq) = eng.score(xc, 10.0, 0.0, want_outputs=True)              # AUTO
    # AUTO estimates the time of both forms: at d = 151 the factor form would have to refresh its factor for this posterior
    # (fold, Cholesky, operand images: ~0.1 ms), far more than the few extra operand rows of the low-rank form cost here
    assert eng.last_engine() == "umma-lowrank"
    mu, var, acq = (t.double().cpu().numpy() for t in (mu, var, acq))
    assert_scores_close(mu, acq, ref_mu, ref_acq)
    np.testing.assert_allclose(var, ref_var, rtol=RTOL_SCORE)
    ref_score = ref_mu.ravel() + ref_acq
    srt = np.sort(ref_score)[::-1]
    if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
        assert eng.best()[1] == int(np.argmax(ref_score))
    if T > 160:                                                                   # more rows than the factor has: the factor form on request
        from ssp_bayesopt_b200 import _lib
        _, (mu, var, acq) = eng.score(xc, 10.0, 0.0, want_outputs=True, engine=_lib.ENGINE_UMMA_F8)
        assert eng.last_engine() == "umma-factor-f8"
        mu, var, acq = (t.double().cpu().numpy() for t in (mu, var, acq))
        assert_scores_close(mu, acq, ref_mu, ref_acq)
        np.testing.assert_allclose(var, ref_var, rtol=RTOL_SCORE)
        # ... and AUTO takes it once the factor is fresh and the set is large enough for the rows to matter
        big = np.tile(xc, (600, 1))
        eng.score(big, 10.0, 0.0)
        assert eng.last_engine() == "umma-factor-f8"


def test_score_lowrank_12d(pkg):
    """Domains with 9..16 dimensions: the low-rank tcgen05 engine covers them (the dense tcgen05 kernel stops at 8)."""
    from ssp_bayesopt_b200 import _lib
    n = 12
    bounds = np.array([[0.0, 1.0]] * n)
    sp = pkg.RandomSSPSpace(n, ssp_dim=600, domain_bounds=bounds, length_scale=0.5, rng=2)
    ls = 0.5 * np.ones(n)
    rng = np.random.default_rng(21)
    X = rng.uniform(0, 1, (40, n))
    y = np.sin(X.sum(axis=1, keepdims=True))
    model = pkg.BayesianLinearRegression(sp.ssp_dim, engine=sp.engine)
    ref = orc.BLR(sp.ssp_dim)
    model.update(sp.encode(X), y)
    ref.update(orc.encode(sp.phase_matrix, ls, X), y)
    xc = rng.uniform(0, 1, (3000, n)).astype(np.float32)
    xc[:8] = X[:8].astype(np.float32)
    phis = orc.encode(sp.phase_matrix, ls, xc.astype(np.float64))
    ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, 10.0, 0.0)
    eng = sp.engine
    for ph in (1, 0):                                   # float32 chain, Q31.32
        eng.set_policy(phase32=ph)
        _, (mu, var, acq) = eng.score(xc, 10.0, 0.0, want_outputs=True)
        assert eng.last_engine() == "umma-lowrank"
        assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
        np.testing.assert_allclose(var.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)
        assert eng.best()[1] == int(np.argmax(ref_mu.ravel() + ref_acq))


def test_score_lowrank_policy_and_rerun(pkg):
    """(a) candidates outside the declared bound: the 32-bit phase path reports them and the step is recomputed with
    the Q31.32 path; (b) more flagged candidates than the list holds: recomputed with the full factor; the AUTO
    engine then stays on the full factor; (c) a packed posterior (export / import) carries the low-rank rows."""
    from ssp_bayesopt_b200 import _lib
    bounds = np.array([[-2.0, 2.0]] * 2)
    sp = pkg.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=bounds, length_scale=0.7, rng=0)
    ls = 0.7 * np.ones(2)
    rng = np.random.default_rng(11)
    X = rng.uniform(-2, 2, (30, 2))
    y = np.cos(X.sum(axis=1, keepdims=True))
    model = pkg.BayesianLinearRegression(sp.ssp_dim, engine=sp.engine)
    ref = orc.BLR(sp.ssp_dim)
    model.update(sp.encode(X), y)
    ref.update(orc.encode(sp.phase_matrix, ls, X), y)
    eng = sp.engine
    assert eng.lowrank_info() == dict(rank=30, valid=True, disabled=False, phase32=True)
    # (a)
    xc = rng.uniform(-2, 2, (5000, 2))
    xc[100] = [7500.0, -9.25]                                               # far outside the declared |x| <= 2
    ref_score, ref_idx = orc.score_and_argmax(orc.encode(sp.phase_matrix, ls, xc), ref, 10.0, 0.0)
    eng.score_stats(reset=True)
    _, (mu, var, acq) = eng.score(xc, 10.0, 0.0, want_outputs=True)
    score = (mu + acq).double().cpu().numpy()
    assert np.max(np.abs(score - ref_score) / np.abs(ref_score)) < RTOL_SCORE
    assert eng.best()[1] == ref_idx
    assert not eng.lowrank_info()["phase32"]                                 # switched to the 64-bit phase path
    eng.set_policy(phase32=True)
    # (b)
    N = (1 << 17) + 5000
    xn = np.repeat(X, N // 30 + 1, axis=0)[:N] + 1e-3 * rng.normal(size=(N, 2))
    ref_mu, ref_var, ref_acq = orc.agent_eval(orc.encode(sp.phase_matrix, ls, xn[:3000]), ref, 10.0, 0.0)
    _, (mu, var, acq) = eng.score(xn, 10.0, 0.0, want_outputs=True)
    # every row sits on an observation: the low-rank form's flag list overflows, then the factor form's (fp16 + e4m3
    # operands) as well, and the step ends on the split-fp16 dense kernel, which never flags
    assert eng.last_engine() == "umma" and eng.lowrank_info()["disabled"]
    assert_scores_close(mu[:3000].double().cpu().numpy(), acq[:3000].double().cpu().numpy(), ref_mu, ref_acq)
    eng.score(xc[:4096], 10.0, 0.0)
    assert eng.last_engine() == "umma"
    eng.set_policy(lowrank=True)
    eng.score(xc[:4096], 10.0, 0.0)
    assert eng.last_engine() == "umma-lowrank"
    # (c) a packed posterior carries the low-rank rows: a second context adopts it and scores identically
    eng.score(xc[:4096], 10.0, 0.0, engine=_lib.ENGINE_UMMA_LR)
    want_key = int(eng._best.item())
    sp2 = pkg.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=bounds, length_scale=0.7, rng=0)
    model2 = pkg.BayesianLinearRegression(sp2.ssp_dim, engine=sp2.engine)
    sp2.engine.blr_import(eng.blr_export(), model.beta)
    model2.beta = model.beta                             # the host mirror of the adopting side (blr.py:54-59 would re-estimate it)
    assert sp2.engine.lowrank_info()["valid"] and sp2.engine.lowrank_info()["rank"] == 30
    sp2.engine.score(xc[:4096], 10.0, 0.0, engine=_lib.ENGINE_UMMA_LR)
    assert int(sp2.engine._best.item()) == want_key
    assert rel(model2.S, model.S) == 0 and rel(model2.m, model.m) == 0
    # ... and keeps updating it: one more observation on both, same operand row appended
    x_new = rng.uniform(-2, 2, (1, 2))
    for m_, s_ in ((model, sp), (model2, sp2)):
        m_.update(s_.encode(x_new), np.array([[0.3]]))
    eng.score(xc[:4096], 10.0, 0.0, engine=_lib.ENGINE_UMMA_LR)
    sp2.engine.score(xc[:4096], 10.0, 0.0, engine=_lib.ENGINE_UMMA_LR)
    assert int(sp2.engine._best.item()) == int(eng._best.item()) and sp2.engine.lowrank_info()["rank"] == 31


def test_score_host_streaming_matches_resident(pkg, golden):
    """Host candidates streamed in chunks (H2D overlapped) give the same packed key as device-resident scoring."""
    import torch
    agt, g = _replay_agent(pkg, golden, "ucb", 0.0)
    eng = agt._scoring_engine()
    x = np.random.default_rng(3).uniform(-5, 5, (50001, 2)).astype(np.float32)
    eng.score(x, 10.0, 0.0, index0=7)
    want = int(eng._best.item())
    for host in (x, torch.from_numpy(x).pin_memory()):
        eng.score_host(host, 10.0, 0.0, index0=7, chunk=4096)
        assert int(eng._best.item()) == want
    eng.score_host(x[:0], 10.0, 0.0)
    assert int(eng._best.item()) == 0


def test_score_edge_cases(pkg, golden):
    agt, g = _replay_agent(pkg, golden, "ucb", 0.0)
    eng = agt._scoring_engine()
    # ragged sizes around the 16-candidate tile, and the tie-break: duplicated rows -> lowest index
    x = np.repeat(g["xc"][:1], 37, axis=0)
    eng.score(x, 10.0, 0.0)
    assert eng.best()[1] == 0
    for n in (1, 15, 16, 17, 255):
        mu, var, acq = agt.eval(g["xc"][:n])
        assert mu.shape == (1, n)
        assert_scores_close(mu, acq, g["ucb_mu"].ravel()[:n], g["ucb_acq"][:n])
    # many single-k-block tiles per CTA (d = 25): generator, MMA and epilogue are several tiles apart; every engine and
    # every repetition must give the same winner (regression: mu buffers of tiles in flight)
    from ssp_bayesopt_b200 import _lib
    xs = np.random.default_rng(3).uniform(-5, 5, (50001, 2)).astype(np.float32)
    eng.score(xs, 10.0, 0.0, engine=_lib.ENGINE_SIMT)
    want_idx = eng.best()[1]
    for rep in range(12):
        for code in (_lib.ENGINE_UMMA, _lib.ENGINE_UMMA_LR):
            eng.score(xs, 10.0, 0.0, engine=code)
            assert _lib.key_unpack(int(eng._best.item()))[1] == want_idx, (rep, code)
    key, outs = eng.score(np.zeros((0, 2)), 10.0, 0.0, want_outputs=True)          # empty candidate set
    assert int(key.item()) == 0 and outs[0].numel() == 0
    with pytest.raises(Exception):
        eng.score(g["xc"], 10.0, -1.0)                                             # gamma_t < 0


# ------------------------------------------------------------------------------------------ K4 + grids
def test_grid_points_and_from_set_decode(pkg, golden):
    g = golden("decode.npz")
    sp = space_from(pkg, g["A"], np.ones(2), g["bounds"])
    ssps, pts = sp.get_sample_pts_and_ssps(20, "grid")
    assert np.array_equal(pts, g["pts"])                                           # bit-exact, meshgrid('xy') order
    np.testing.assert_allclose(ssps, g["ssps"], atol=2e-13)
    assert np.array_equal(sp.decode(g["queries"], method="from-set", samples=(ssps, pts)), g["decoded"])
    assert np.array_equal(sp.decode(g["queries"], method="from-set", samples=(g["ssps"], g["pts"])), g["decoded"])
    sp3 = pkg.RandomSSPSpace(3, ssp_dim=31, domain_bounds=g["bounds3"], length_scale=1.0, rng=1)
    assert np.array_equal(sp3.get_sample_points(4, "grid"), g["pts3"])             # n >= 3 axis order
    # shard of a grid == slice of the full grid
    axes = sp.grid_axes(20, "grid")
    part = sp.engine.grid_points(axes, 137, 100).cpu().numpy()
    assert np.array_equal(part, g["pts"][137:237])
    # ties -> first index; zero query (norm floor 1e-6, sspspace.py:352) -> index 0
    dup = np.vstack([g["ssps"][5:6]] * 4 + [g["ssps"][:3]])
    idx = sp.engine.from_set_argmax(dup, g["ssps"][5:6]).cpu().numpy()
    assert idx[0] == 0
    assert sp.engine.from_set_argmax(g["ssps"], np.zeros((1, sp.ssp_dim))).cpu().numpy()[0] == 0
    with pytest.raises(NotImplementedError):
        sp.decode(g["queries"], method="nope")
    # K5 binding (sspspace.py:551-557): golden value, broadcasting, identity and inverse
    np.testing.assert_allclose(sp.bind(g["queries"][0], g["queries"][1]), g["bind"], atol=1e-14)
    np.testing.assert_allclose(sp.bind(g["queries"], g["queries"][1]), orc.bind(g["queries"], g["queries"][1]), atol=1e-14)
    np.testing.assert_allclose(sp.bind(g["queries"], sp.identity()), g["queries"], atol=1e-15)
    np.testing.assert_allclose(sp.bind(g["queries"][:1], sp.invert(g["queries"][:1])), sp.identity(), atol=1e-12)
    x12 = np.array([[1.0, -2.0], [0.5, 3.0]])
    np.testing.assert_allclose(sp.bind(sp.encode(x12[:1]), sp.encode(x12[1:])), sp.encode(x12.sum(0, keepdims=True)), atol=1e-12)


def test_decode_round_trip_reference_tests(pkg):
    """tests/test_ssps.py:39-54, :75-80, :92-98 re-run against the GPU path."""
    b15 = 15 * np.array([[-1.0, 1.0], [-1.0, 1.0]])
    s = 2 * np.pi / np.sqrt(6)
    x = np.atleast_2d(np.array([1.3, -3.4]))
    hexs = pkg.HexagonalSSPSpace(2, ssp_dim=151, scale_min=s - 0.5, scale_max=s + 0.5, domain_bounds=b15, length_scale=1)
    assert np.isclose(np.inner(hexs.encode(x), hexs.encode(x)), 1)
    assert np.all(np.isclose(x, hexs.decode(hexs.encode(x), method="from-set", num_samples=300), atol=1e-1))
    rnd = pkg.RandomSSPSpace(2, ssp_dim=151, domain_bounds=b15, length_scale=1, rng=0)
    assert np.all(np.isclose(x, rnd.decode(rnd.encode(x), method="from-set", num_samples=300), atol=1e-1))
    small = pkg.HexagonalSSPSpace(2, ssp_dim=151, scale_min=s - 0.5, scale_max=s + 0.5, domain_bounds=b15,
                                  length_scale=0.05)
    assert np.all(np.isclose(x, small.decode(small.encode(x), method="from-set", num_samples=500), atol=1e-2))
    assert np.all(np.isclose(x, hexs.decode(hexs.encode(x), method="direct-optim"), atol=1e-4))


def test_decode_direct_optim_device(pkg):
    """'direct-optim' decode on the device (analytic projected Newton) against the reference's flow restated with
    scipy's L-BFGS-B (oracle.decode_direct_optim, sspspace.py:358-375) from the same from-set start, and the
    reference's own round-trip tests (tests/test_ssps.py:75-90: np.isclose at default tolerance)."""
    b15 = np.array([[-15.0, 15.0], [-15.0, 15.0]])
    s = 2 * np.pi / np.sqrt(6)
    x = np.atleast_2d(np.array([1.3, -3.4]))
    hexs = pkg.HexagonalSSPSpace(2, ssp_dim=151, scale_min=s - 0.5, scale_max=s + 0.5, domain_bounds=b15, length_scale=1)
    rnd = pkg.RandomSSPSpace(2, ssp_dim=151, domain_bounds=b15, length_scale=1, rng=0)
    for sp in (hexs, rnd):
        assert np.all(np.isclose(x, sp.decode(sp.encode(x), method="direct-optim")))          # rtol 1e-5, atol 1e-8
    # several queries at once, noisy SSPs (not exactly on the manifold), explicit decoding set
    rng = np.random.default_rng(9)
    b4 = np.array([[-4.0, 4.0], [-4.0, 4.0], [-4.0, 4.0]])
    sp3 = pkg.RandomSSPSpace(3, ssp_dim=201, domain_bounds=b4, length_scale=1.5, rng=1)
    ls = 1.5 * np.ones(3)
    xs = rng.uniform(-3.5, 3.5, (12, 3))
    xs[0] = [4.0, -4.0, 3.9]                                                                    # on / next to the box
    ssps = orc.encode(sp3.phase_matrix, ls, xs) + 0.02 * rng.normal(size=(12, sp3.ssp_dim))
    samples = sp3.get_sample_pts_and_ssps(12, "grid")
    got = sp3.decode(ssps, method="direct-optim", samples=samples)
    x0 = orc.decode_from_set(ssps, samples[0], samples[1])
    want = orc.decode_direct_optim(sp3.phase_matrix, ls, ssps, x0, bounds=b4)
    assert got.shape == (12, 3)
    # same local maximum: L-BFGS-B with numerical gradients stops within ~1e-5 (1e-4 with an active bound) of it, the
    # Newton iteration lands on it
    np.testing.assert_allclose(got, want, atol=2e-4)
    assert np.median(np.abs(got - want)) < 5e-6
    unit = ssps / np.linalg.norm(ssps, axis=1, keepdims=True)
    f_got = np.sum(orc.encode(sp3.phase_matrix, ls, got) * unit, axis=1)
    f_want = np.sum(orc.encode(sp3.phase_matrix, ls, want) * unit, axis=1)
    assert np.all(f_got >= f_want - 1e-12)                                                      # never a worse optimum
    assert np.all(got >= b4[:, 0] - 1e-15) and np.all(got <= b4[:, 1] + 1e-15)
    assert sp3.decode(np.zeros((0, sp3.ssp_dim)), method="direct-optim", samples=samples).shape == (0, 3)
    # and against the live reference's own output (tests/golden/decode_optim.npz)
    gd = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode_optim.npz"))
    spg = space_from(pkg, gd["A"], gd["ls"], gd["bounds"])
    gs = (orc.encode(gd["A"], gd["ls"], gd["set_pts"]), gd["set_pts"])
    np.testing.assert_allclose(spg.decode(gd["ssps"], method="direct-optim", samples=gs), gd["decoded"], atol=2e-4)


# ------------------------------------------------------------------------------------------ trajectory agent
def test_trajectory_agent_golden(pkg, golden):
    g = golden("traj.npz")
    agt = pkg.SSPTrajectoryAgent(g["trajs"], g["tys"], x_dim=2, traj_len=4, ssp_dim=51, domain_bounds=g["bounds"],
                                 length_scale=0.7, time_length_scale=1.3, decoder_method="from-set", gamma_c=0.0,
                                 beta_ucb=5.0)
    assert np.array_equal(agt.ssp_x_space.phase_matrix, g["Ax"]) and np.array_equal(agt.ssp_t_space.phase_matrix, g["At"])
    np.testing.assert_allclose(agt.timestep_ssps, g["timestep_ssps"], atol=2e-13)
    phi = agt.encode(g["xc"])
    np.testing.assert_allclose(phi, g["phi"], atol=5e-13)
    assert agt.blr.beta == g["beta"]
    assert rel(agt.blr.m, g["m"]) < 1e-9 and rel(agt.blr.S, g["S"]) < 1e-9
    mu, var, acq = agt.eval(g["xc"])
    assert_scores_close(mu, acq, g["mu"], g["acq"])
    np.testing.assert_allclose(var, g["var"], rtol=RTOL_SCORE)
    assert np.array_equal(agt.decode(phi[:1]), g["decoded"])


@pytest.mark.parametrize("L,ssp_dim,N", [(4, 51, 300), (10, 200, 1500), (10, 600, 5000)])
def test_trajectory_tcgen05_engine(pkg, L, ssp_dim, N):
    """BASELINE config C5 shape in miniature: the tcgen05 engine sums L phasors per frequency (trajectory mode)."""
    from ssp_bayesopt_b200 import _lib
    bounds = np.array([[-2.0, 2.0], [-2.0, 2.0]])
    rng = np.random.default_rng(L)
    trajs = rng.uniform(-2, 2, (12, 2 * L))
    tys = -np.sum(trajs ** 2, axis=1, keepdims=True) / L
    agt = pkg.SSPTrajectoryAgent(trajs, tys, x_dim=2, traj_len=L, ssp_dim=ssp_dim, domain_bounds=bounds, length_scale=0.7,
                                 time_length_scale=1.3, decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    Ax, At = agt.ssp_x_space.phase_matrix, agt.ssp_t_space.phase_matrix
    T = orc.traj_timestep_ssps(At, np.array([1.3]), L)
    ref = orc.BLR(Ax.shape[0])
    ref.update(orc.traj_encode(Ax, 0.7 * np.ones(2), T, trajs, L, 2), tys)
    xc = rng.uniform(-2, 2, (N, 2 * L)).astype(np.float32)
    xc[:4] = trajs[:4].astype(np.float32)
    phis = orc.traj_encode(Ax, 0.7 * np.ones(2), T, xc.astype(np.float64), L, 2)
    ref_mu, ref_var, ref_acq = orc.agent_eval(phis, ref, 5.0, 0.0)
    eng = agt._scoring_engine()
    for code, name, ph in ((_lib.ENGINE_UMMA, "umma", 1), (_lib.ENGINE_SIMT, "simt", 1),
                           (_lib.ENGINE_UMMA_LR, "umma-lowrank", 1), (_lib.ENGINE_UMMA_LR, "umma-lowrank", 0)):
        eng.set_policy(phase32=ph)                    # low-rank trajectory mode: float32 chain and Q31.32 phases
        _, (mu, var, acq) = eng.score(xc, 5.0, 0.0, want_outputs=True, engine=code)
        assert eng.last_engine() == name
        assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
        np.testing.assert_allclose(var.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)
        ref_score = ref_mu.ravel() + ref_acq
        srt = np.sort(ref_score)[::-1]
        if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
            assert eng.best()[1] == int(np.argmax(ref_score))


@pytest.mark.parametrize("tag", ["a3", "a2"])
def test_multi_agent_golden(pkg, golden, tag):
    """SSPMultiAgent against the live reference's output (tests/golden/multi_agent.npz):
    tag vectors, normalised encode, posterior, eval through every engine that forms |phi|, decode."""
    from ssp_bayesopt_b200 import _lib
    g = golden("multi_agent.npz")
    n_agents, L, xd, ssp_dim, seed = (int(v) for v in g[f"{tag}_dims"])
    agt = pkg.SSPMultiAgent(g[f"{tag}_trajs"], g[f"{tag}_tys"], n_agents=n_agents, x_dim=xd, traj_len=L, ssp_dim=ssp_dim,
                            domain_bounds=g["bounds"], length_scale=0.8, time_length_scale=1.2, seed=seed,
                            decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    assert np.array_equal(agt.ssp_x_spaces[0].phase_matrix, g[f"{tag}_Ax"])
    assert np.array_equal(agt.ssp_t_space.phase_matrix, g[f"{tag}_At"])
    np.testing.assert_allclose(agt.agent_sps, g[f"{tag}_agent_sps"], atol=1e-15)
    np.testing.assert_allclose(agt.timestep_ssps, g[f"{tag}_timestep_ssps"], atol=2e-13)
    phi = agt.encode(g[f"{tag}_xc"])
    np.testing.assert_allclose(phi, g[f"{tag}_phi"], atol=5e-13)
    np.testing.assert_allclose(np.linalg.norm(phi, axis=1), 1.0, atol=1e-13)
    assert abs(float(np.ravel(agt.blr.beta)[0]) - float(g[f"{tag}_beta"])) <= 1e-12 * float(g[f"{tag}_beta"])
    assert rel(agt.blr.m, g[f"{tag}_m"]) < 1e-9 and rel(agt.blr.S, g[f"{tag}_S"]) < 1e-9
    mu, var, acq = agt.eval(g[f"{tag}_xc"])                                   # 40 candidates: exact float64 engine
    assert_scores_close(mu, acq, g[f"{tag}_mu"], g[f"{tag}_acq"])
    np.testing.assert_allclose(var, g[f"{tag}_var"], rtol=RTOL_SCORE)
    eng = agt._scoring_engine()
    for code, name in ((_lib.ENGINE_SIMT, "simt"), (_lib.ENGINE_UMMA_LR, "umma-lowrank")):
        _, (mu, var, acq) = eng.score(g[f"{tag}_xc"], 5.0, 0.0, want_outputs=True, engine=code)
        assert eng.last_engine() == name
        assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), g[f"{tag}_mu"], g[f"{tag}_acq"])
        np.testing.assert_allclose(var.double().cpu().numpy(), g[f"{tag}_var"], rtol=RTOL_SCORE)
    with pytest.raises(_lib.SspboError):                                      # the dense-factor kernel does not form |phi|
        eng.score(g[f"{tag}_xc"], 5.0, 0.0, engine=_lib.ENGINE_UMMA)
    assert np.array_equal(agt.decode(phi[:1]), g[f"{tag}_decoded"])


def test_multi_agent_tcgen05_and_driver(pkg):
    """Larger multi-agent candidate sets: AUTO picks the low-rank tcgen05 engine in trajectory mode with normalisation
    (both phase paths), checked against the oracle restatement of ssp_multi_agent.py:166-186; then the BO driver."""
    from ssp_bayesopt_b200 import _lib
    bounds = np.array([[-2.0, 2.0], [-2.0, 2.0]])
    n_agents, L, xd, ssp_dim, N = 3, 6, 2, 300, 3000
    rng = np.random.default_rng(17)
    trajs = rng.uniform(-2, 2, (14, L * n_agents * xd))
    tys = -np.sum(trajs ** 2, axis=1, keepdims=True) / (L * n_agents)
    agt = pkg.SSPMultiAgent(trajs, tys, n_agents=n_agents, x_dim=xd, traj_len=L, ssp_dim=ssp_dim, domain_bounds=bounds,
                            length_scale=0.8, time_length_scale=1.2, seed=1, decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    Ax, At = agt.ssp_x_spaces[0].phase_matrix, agt.ssp_t_space.phase_matrix
    T = orc.traj_timestep_ssps(At, np.array([1.2]), L)
    tags = orc.sp_vectors(n_agents, Ax.shape[0], 1)
    enc = lambda x: orc.multi_agent_encode(Ax, 0.8 * np.ones(2), T, tags, x, L, n_agents, xd)
    ref = orc.BLR(Ax.shape[0])
    ref.update(enc(trajs), tys)
    xc = rng.uniform(-2, 2, (N, L * n_agents * xd)).astype(np.float32)
    xc[:4] = trajs[:4].astype(np.float32)
    ref_mu, ref_var, ref_acq = orc.agent_eval(enc(xc.astype(np.float64)), ref, 5.0, 0.0)
    ref_score = ref_mu.ravel() + ref_acq
    eng = agt._scoring_engine()
    for ph in (1, 0):
        eng.set_policy(phase32=ph)
        _, (mu, var, acq) = eng.score(xc, 5.0, 0.0, want_outputs=True)
        assert eng.last_engine() == "umma-lowrank"
        assert_scores_close(mu.double().cpu().numpy(), acq.double().cpu().numpy(), ref_mu, ref_acq)
        np.testing.assert_allclose(var.double().cpu().numpy(), ref_var, rtol=RTOL_SCORE)
        srt = np.sort(ref_score)[::-1]
        if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
            assert eng.best()[1] == int(np.argmax(ref_score))
    f = lambda x: float(-np.sum(np.asarray(x) ** 2) / (L * n_agents))
    opt = pkg.BayesianOptimization(f=lambda x, info=None: f(x), bounds=bounds)
    opt.maximize(init_points=(trajs, tys), n_iter=3, agent_type="ssp-multi", n_agents=n_agents, x_dim=xd, traj_len=L,
                 ssp_dim=ssp_dim, length_scale=0.8, time_length_scale=1.2, seed=1, decoder_method="from-set", gamma_c=0.0,
                 beta_ucb=5.0, n_candidates=20000)
    assert opt.xs.shape == (17, L * n_agents * xd) and np.all(np.abs(opt.xs) <= 2)


def test_agent_without_length_scale(pkg, golden):
    """SSPAgent with no `length_scale`: the sinc-kernel GP fit on the host picks it (ssp_agent.py:96-109), then the usual
    device path; length scale and posterior mean against the live reference."""
    g = golden("lengthscale_gp.npz")
    for tag in ("a", "b"):
        sp = pkg.HexagonalSSPSpace(2, ssp_dim=51, domain_bounds=g["bounds"], length_scale=1.0)
        agt = pkg.SSPAgent(g[f"{tag}_xs"], g[f"{tag}_ys"], sp, decoder_method="from-set")
        np.testing.assert_allclose(np.ravel(agt.ssp_space.length_scale), g[f"{tag}_length_scale"], rtol=1e-6)
        assert rel(agt.blr.m, g[f"{tag}_m"]) < 1e-5


def test_trajectory_lengthscale_search(pkg, golden):
    """Omitted `length_scale`: the K-fold search of ssp_traj_agent.py:113-143 with the encode and
    the Gram matrix on the device.  Objective against every evaluation recorded inside the live reference; result against
    the reference's (the valley is flat along the time length scale, so the optimum is compared through the objective)."""
    g = golden("lengthscale_cv.npz")
    rng = np.random.default_rng(3)
    eng = pkg.DeviceEngine()
    sp = pkg.RandomSSPSpace(2, ssp_dim=64, domain_bounds=g["bounds"], length_scale=1.0, rng=0)
    phi = sp.encode(rng.uniform(-2, 2, (9, 2)))
    np.testing.assert_allclose(sp.engine.gram(phi).cpu().numpy(), phi @ phi.T, rtol=0, atol=1e-14)
    agt = pkg.SSPTrajectoryAgent(g["trajs"], g["tys"], x_dim=2, traj_len=4, ssp_dim=51, domain_bounds=g["bounds"],
                                 decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    found = np.array([np.ravel(agt.ssp_x_space.length_scale)[0], np.ravel(agt.ssp_t_space.length_scale)[0]])
    assert np.array_equal(agt.ssp_x_space.phase_matrix, g["Ax"]) and np.array_equal(agt.ssp_t_space.phase_matrix, g["At"])
    for x, f in list(zip(g["eval_x"], g["eval_f"]))[::5]:
        np.testing.assert_allclose(agt._cv_loss(x, g["trajs"], g["tys"]), f, rtol=1e-9)
    f_found = agt._cv_loss(found, g["trajs"], g["tys"])
    f_ref = agt._cv_loss(g["result"], g["trajs"], g["tys"])
    np.testing.assert_allclose(f_ref, g["eval_f"].min(), rtol=1e-9)
    assert f_found <= f_ref + 2e-3 * abs(f_ref)
    np.testing.assert_allclose(found[0], g["result"][0], rtol=0.1)
    assert found[1] > 50.0                                             # flat direction: any large time length scale
    # the agent is usable afterwards with the length scales it found (engine re-pointed at timesteps 1 .. L+1)
    agt.ssp_x_space.update_lengthscale(found[0]); agt.ssp_t_space.update_lengthscale(found[1])
    agt._point_engine(np.linspace(1, 5, 4))
    T = orc.traj_timestep_ssps(g["At"], np.array([found[1]]), 4)
    want = orc.traj_encode(g["Ax"], found[0] * np.ones(2), T, g["trajs"][:3], 4, 2)
    np.testing.assert_allclose(agt.encode(g["trajs"][:3]), want, atol=5e-13)


# ------------------------------------------------------------------------------------------ generators, driver
def test_fill_uniform_bit_exact(pkg):
    eng = pkg.DeviceEngine()
    for seed, start, count, n in ((0, 0, 1000, 6), (7, 123456789, 257, 2), (1, 2 ** 33, 64, 20)):
        got = eng.fill_uniform(seed, start, count, n).cpu().numpy()
        assert np.array_equal(got, orc.counter_uniform(seed, start, count, n))


def test_grid_row_lookup_matches_device_grid(pkg):
    """`maximize` names the winning grid row on the host; it must be the row the device grid holds."""
    for n, per_dim in ((2, 7), (3, 5)):
        bounds = np.array([[-1.0, 2.0]] * n)
        opt = pkg.BayesianOptimization(f=lambda x: -np.sum(x ** 2, axis=1), bounds=bounds)
        xs0 = np.random.default_rng(n).uniform(-1, 2, (6, n))
        agt, _, _ = opt.initialize_agent((xs0, -np.sum(xs0 ** 2, axis=1, keepdims=True)), 'ssp-rand', domain_bounds=bounds,
                                         ssp_dim=33, length_scale=0.9, rng=0)
        cand = opt._build_candidates(agt, 'grid', per_dim, None, 0)
        dev = cand['x'].cpu().numpy()
        assert dev.shape == (per_dim ** n, n)
        for i in (0, 1, per_dim, per_dim ** n - 1, per_dim ** n // 2 + 1):
            assert np.array_equal(cand['lookup'](i), dev[i]), (n, i)


def _ascent_numpy(m, S, beta, lam, gamma, phi0, margin, max_iter, tol):
    """float64 numpy statement of the iteration sspbo_acq_ascent runs (phi <- margin grad / |grad|), per restart."""
    m = np.ravel(m)
    c = gamma + 1.0 / beta
    out, vals, its = [], [], []
    for phi in np.atleast_2d(phi0).astype(np.float64):
        nrm = np.linalg.norm(phi)
        phi = margin * phi / nrm if nrm > margin else phi.copy()
        moved, k = np.inf, 0
        while True:
            y = S @ phi
            s = np.sqrt(c + phi @ y)
            val = m @ phi + lam * s - np.sqrt(gamma)
            if moved <= tol * margin or k == max_iter:
                break
            g = m + (lam / s) * y
            new = margin * g / np.linalg.norm(g)
            moved, phi, k = np.linalg.norm(new - phi), new, k + 1
        out.append(phi); vals.append(val); its.append(k)
    return np.stack(out), np.array(vals), np.array(its)


def test_acq_ascent_device(pkg, golden):
    """K6: phi-space acquisition search on the device against the reference's route
    (oracle.acquisition_func = ssp_agent.py:156-185, L-BFGS-B restarts = bayesian_optimization.py:270-287, pinned by
    tests/golden/acq_ascent.npz from the live reference)."""
    g = golden("acq_ascent.npz")
    sp = space_from(pkg, g["A"], g["ls"], g["bounds"])
    agt = pkg.SSPAgent(g["xs"][:10], g["ys"][:10], sp, decoder_method="from-set", length_scale=1.0)
    for i in range(10, 30):
        _, var_t, _ = agt.eval(g["xs"][i:i + 1])
        agt.update(g["xs"][i:i + 1], g["ys"][i:i + 1], var_t, step_num=i)
    lam, gamma = float(agt.var_weight), agt._gamma()
    assert lam == float(g["var_weight"]) and abs(gamma - float(g["gamma_t"])) <= 1e-4 * float(g["gamma_t"])
    m, S, beta = agt.blr.m, agt.blr.S, float(np.ravel(agt.blr.beta)[0])
    assert rel(m, g["m"]) < 1e-7 and rel(S, g["S"]) < 1e-7
    f_dev, _ = orc.acquisition_func(m, S, beta, lam, gamma)          # the reference's objective on the device's own posterior
    eng = agt._scoring_engine()

    # (1) from the reference's own starting points: same objective value at the returned phi, never below the reference's
    #     L-BFGS-B result (which stops at the clamped start), and the iterates of the float64 numpy statement
    for max_iter, tol in ((0, 0.0), (25, 0.0), (400, 1e-9)):
        phi, val, its = (t.cpu().numpy() for t in eng.acq_ascent(g["phi0"], lam, gamma, max_iter=max_iter, tol=tol))
        want_phi, want_val, want_its = _ascent_numpy(m, S, beta, lam, gamma, g["phi0"], 4.0, max_iter, tol)
        assert phi.shape == g["phi0"].shape and val.shape == (5,)
        np.testing.assert_allclose(val, [-np.ravel(f_dev(p))[0] for p in phi], rtol=1e-12)
        np.testing.assert_allclose(val, want_val, rtol=1e-10)
        np.testing.assert_allclose(phi, want_phi, atol=1e-7 * 4.0)
        assert np.all(np.linalg.norm(phi, axis=1) <= 4.0 * (1 + 1e-12))
        if max_iter == 0:
            np.testing.assert_allclose(val, -g["lbfgs_fun"], rtol=1e-6)     # the reference's result: the clamped start
            assert np.all(its == 0)
        else:
            assert np.all(val >= -g["lbfgs_fun"] * (1 - 1e-9))
        if tol > 0:
            assert np.all(np.abs(its - want_its) <= 2) and np.all(its < max_iter)
    # (2) stationarity of the converged points: phi* = 4 grad / |grad| with the reference's gradient
    _, grad = orc.acquisition_func(m, S, beta, lam, gamma)
    for p in phi:
        gr = -grad(p)
        np.testing.assert_allclose(p, 4.0 * gr / np.linalg.norm(gr), atol=1e-7)
    # (3) monotone: more steps never lower the value
    vals = [eng.acq_ascent(g["phi0"], lam, gamma, max_iter=k, tol=0.0)[1].cpu().numpy() for k in (0, 1, 2, 5, 10, 50)]
    assert np.all(np.diff(np.stack(vals), axis=0) >= -1e-9 * np.abs(vals[0]))
    # (4) more restarts than one tile, starting points inside the ball, a single 1-D start
    rng = np.random.default_rng(5)
    p0 = rng.normal(size=(19, sp.ssp_dim)) * rng.uniform(0.01, 1.0, size=(19, 1))
    phi, val, its = (t.cpu().numpy() for t in eng.acq_ascent(p0, lam, gamma, max_iter=60, tol=0.0))
    want_phi, want_val, _ = _ascent_numpy(m, S, beta, lam, gamma, p0, 4.0, 60, 0.0)
    np.testing.assert_allclose(val, want_val, rtol=1e-10)
    np.testing.assert_allclose(phi, want_phi, atol=4e-7)
    one = eng.acq_ascent(p0[3], lam, gamma, max_iter=60, tol=0.0)
    np.testing.assert_array_equal(one[0].cpu().numpy()[0], phi[3])
    # (5) agent / driver surface
    phis, vals = agt.maximize_acquisition(num_restarts=5, rng=0)
    assert phis.shape == (5, sp.ssp_dim) and vals.shape == (5,)
    assert np.all(vals >= -g["lbfgs_fun"].max() - 1e-6)
    x_star = agt.decode(phis[np.argmax(vals):np.argmax(vals) + 1])
    assert x_star.shape == (1, 2) and np.all(np.abs(x_star) <= 5)
    bounds = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    f = lambda x: himmelblau(np.atleast_2d(x))
    runs = []
    for _ in range(2):
        opt = pkg.BayesianOptimization(f=f, bounds=bounds, random_state=4)
        opt.maximize(init_points=(g["xs"][:10], g["ys"][:10]), n_iter=4, num_restarts=3, agent_type="ssp-hex", ssp_dim=151,
                     length_scale=1.0, decoder_method="from-set", acq_optimizer="phi-ascent")
        assert opt.xs.shape == (14, 2) and np.all(np.abs(opt.xs) <= 5)
        runs.append(opt.xs.copy())
    assert np.array_equal(runs[0], runs[1])                               # seeded starts, deterministic kernel


def test_acq_ascent_raw_feature_posterior(pkg):
    """The ascent only needs (m, S, beta): a posterior over raw features (sspbo_blr_init_dim), d not a multiple of the
    warp size, rank-deficient data."""
    rng = np.random.default_rng(2)
    d = 77
    model = pkg.BayesianLinearRegression(d)
    X = rng.normal(size=(12, d)) / np.sqrt(d)
    model.update(X, (X[:, :3].sum(axis=1, keepdims=True)))
    m, S, beta = model.m, model.S, float(np.ravel(model.beta)[0])
    p0 = rng.normal(size=(3, d))
    phi, val, its = (t.cpu().numpy() for t in model.engine.acq_ascent(p0, 2.5, 0.3, margin=1.5, max_iter=300, tol=1e-10))
    want_phi, want_val, want_its = _ascent_numpy(m, S, beta, 2.5, 0.3, p0, 1.5, 300, 1e-10)
    np.testing.assert_allclose(val, want_val, rtol=1e-10)
    np.testing.assert_allclose(phi, want_phi, atol=1e-7)
    f, _ = orc.acquisition_func(m, S, beta, 2.5, 0.3, margin=1.5)
    np.testing.assert_allclose(val, [-np.ravel(f(p))[0] for p in phi], rtol=1e-12)


def test_observe_equals_eval_then_update(pkg, golden):
    """`SSPAgent.observe` (predictive scalars taken from the fused update launch) against `eval` followed by `update` on a
    twin agent, MI acquisition so gamma_t depends on the returned variance; also the trajectory and multi-agent contexts."""
    b = golden("blr.npz")
    rng = np.random.default_rng(6)

    def twin():
        sp = pkg.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=B2, length_scale=1.0)
        return pkg.SSPAgent(b["xs"], b["ys"], sp, decoder_method="from-set", gamma_c=1.0, beta_ucb=10.0, var_decay=0.25,
                            length_scale=1.0)
    a1, a2 = twin(), twin()
    for i in range(8):
        x_t = rng.uniform(-5, 5, (1, 2))
        y_t = np.atleast_2d(himmelblau(x_t))
        mu1, var1, acq1 = a1.eval(x_t)
        a1.update(x_t, y_t, var1, step_num=i)
        mu2, var2, acq2 = a2.observe(x_t, y_t, step_num=i)
        assert mu2.shape == (1, 1) and var2.shape == (1,) and acq2.shape == (1,)
        np.testing.assert_allclose(mu2, mu1, rtol=1e-5, atol=1e-6 * np.abs(b["ys"]).max())   # eval returns float32-rounded values
        np.testing.assert_allclose(var2, var1, rtol=1e-5)
        np.testing.assert_allclose(acq2, acq1, rtol=1e-5)
    assert a1.var_weight == a2.var_weight
    np.testing.assert_allclose(np.ravel(a2.gamma_t), np.ravel(a1.gamma_t), rtol=1e-5)
    assert np.array_equal(a1.blr.m, a2.blr.m) and np.array_equal(a1.blr.S, a2.blr.S)         # same kernels, same order
    xc = rng.uniform(-5, 5, (2000, 2))
    k1, k2 = a1.best_candidate(xc), a2.best_candidate(xc)
    assert a1._scoring_engine().best()[1] == a2._scoring_engine().best()[1]
    # normalised multi-agent context: mean / variance of the unit-norm phi
    g = golden("multi_agent.npz")
    n_agents, L, xd, ssp_dim, seed = (int(v) for v in g["a2_dims"])
    agt = pkg.SSPMultiAgent(g["a2_trajs"], g["a2_tys"], n_agents=n_agents, x_dim=xd, traj_len=L, ssp_dim=ssp_dim,
                            domain_bounds=g["bounds"], length_scale=0.8, time_length_scale=1.2, seed=seed,
                            decoder_method="from-set", gamma_c=0.0, beta_ucb=5.0)
    mu, var, acq = agt.observe(g["a2_xc"][:1], np.array([[0.3]]))
    np.testing.assert_allclose(mu.ravel(), g["a2_mu"].ravel()[:1], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, g["a2_var"][:1], rtol=1e-8)
    np.testing.assert_allclose(acq, g["a2_acq"][:1], rtol=1e-8)


def test_update_paths_agree(pkg, golden):
    """The single-launch posterior update against the separate-kernel path it replaces (kept as the fallback), on an SSP
    context (psi-space covariance, low-rank rows, scoring afterwards) and on a raw-feature context."""
    b = golden("blr.npz")
    rng = np.random.default_rng(12)
    agents = []
    for fused in (1, 0):
        sp = pkg.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=B2, length_scale=1.0)
        assert sp.engine.lib.sspbo_set_update_path(sp.engine._ctx, fused) == 0
        agents.append(pkg.SSPAgent(b["xs"], b["ys"], sp, decoder_method="from-set", gamma_c=0.0, beta_ucb=10.0, length_scale=1.0))
    xs, ys = rng.uniform(-5, 5, (30, 2)), rng.normal(size=(30, 1))
    for a in agents:
        for i in range(30):
            a.update(xs[i:i + 1], ys[i:i + 1], 0.0)
    a1, a0 = agents
    assert rel(a1.blr.m, a0.blr.m) < 1e-11 and rel(a1.blr.S, a0.blr.S) < 1e-11 and rel(a1.blr.S_inv, a0.blr.S_inv) < 1e-12
    assert a1._scoring_engine().lowrank_info() == a0._scoring_engine().lowrank_info()
    xc = rng.uniform(-5, 5, (3000, 2)).astype(np.float32)
    outs = []
    for a in agents:
        _, (mu, var, acq) = a._scoring_engine().score(xc, 10.0, 0.0, want_outputs=True)
        assert a._scoring_engine().last_engine() == "umma-lowrank"
        outs.append((mu.cpu().numpy(), var.cpu().numpy()))
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(outs[0][1], outs[1][1], rtol=1e-5)
    models = []
    for fused in (1, 0):
        m = pkg.BayesianLinearRegression(77)
        m.engine.lib.sspbo_set_update_path(m.engine._ctx, fused)
        X = np.random.default_rng(2).normal(size=(12, 77))
        m.update(X, X[:, :1])
        models.append(m)
    assert rel(models[0].m, models[1].m) < 1e-11 and rel(models[0].S, models[1].S) < 1e-11


def test_host_point_entry_points(pkg, golden):
    """sspbo_eval_host / sspbo_blr_update_x / sspbo_score_finish against the tensor-based calls they shortcut."""
    import ctypes as C
    from ssp_bayesopt_b200 import _lib
    b = golden("blr.npz")
    rng = np.random.default_rng(8)

    def agent():
        sp = pkg.RandomSSPSpace(2, ssp_dim=128, domain_bounds=B2, length_scale=1.0, rng=2)
        return pkg.SSPAgent(b["xs"], b["ys"], sp, decoder_method="from-set", gamma_c=0.0, beta_ucb=10.0, length_scale=1.0)
    a1, a2 = agent(), agent()
    e1, e2 = a1._scoring_engine(), a2._scoring_engine()
    # eval: 64 host points (one C call) == the same points as a device tensor (sspbo_score with outputs), 65 points take the old path
    x64, x65 = rng.uniform(-5, 5, (64, 2)), rng.uniform(-5, 5, (65, 2))
    for xs in (x64, x65, x64[:1]):
        got = e1.eval_host(xs, 10.0, 0.5, engine=_lib.ENGINE_EXACT)
        _, (mu, var, acq) = e1.score(e1.to_device(xs), 10.0, 0.5, want_outputs=True, engine=_lib.ENGINE_EXACT)
        for g_, w_ in zip(got, (mu, var, acq)):
            assert g_.dtype == np.float64 and g_.shape == (len(xs),)
            if len(xs) <= 64:      # float64 straight from the exact engine; the tensor route rounds the same values to float32
                np.testing.assert_array_equal(g_.astype(np.float32), w_.cpu().numpy())
            else:
                np.testing.assert_array_equal(g_, w_.double().cpu().numpy())
    with pytest.raises(AssertionError):
        e1.eval_host(np.zeros((3, 5)), 10.0, 0.5)
    # update from host points (130 of them: three library calls) == update(encode(x), y)
    xs, ys = rng.uniform(-5, 5, (130, 2)), rng.normal(size=(130, 1))
    a1.blr.update_points(xs, ys)
    a2.blr.update(a2.encode(xs), ys)
    # (the two routes encode with different batch shapes, i.e. different float64 summation orders in K1)
    assert rel(a1.blr.m, a2.blr.m) < 1e-11 and rel(a1.blr.S, a2.blr.S) < 1e-11
    assert e1.lowrank_info() == e2.lowrank_info()
    with pytest.raises(AssertionError):
        a1.blr.update_points(np.zeros((2, 3)), np.zeros((2, 1)))
    rc = e1.lib.sspbo_blr_update_x(e1._ctx, xs.ctypes.data_as(_lib._pd), ys.ctypes.data_as(_lib._pd), 65, None)
    assert rc != 0 and b"64" in e1.lib.sspbo_last_error(e1._ctx)
    # selection read-back: key and counters in one call
    xc = rng.uniform(-5, 5, (5000, 2)).astype(np.float32)
    key = a1.best_candidate(xc)
    want = int(key.item()) & 0xFFFFFFFFFFFFFFFF
    st, got = e1._finish()
    assert got == want and not st["rerun"] and st["scored"] in (0, 5000)      # counters belong to the low-rank engine only
    assert e1.best() == _lib.key_unpack(want)
    # empty ascent
    out = e1.acq_ascent(np.zeros((0, e1.d)), 10.0, 0.5)
    assert out[0].shape == (0, e1.d) and out[1].shape == (0,)


def _build_c_abi_smoke(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "ssp_bayesopt_b200", "lib")
    exe = str(tmp_path / "c_abi_smoke")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-std=c99", "-Wall", "-I" + os.path.join(root, "include"), "-I" + os.path.join(cuda, "include"),
           os.path.join(root, "tests", "c_abi_smoke.c"), "-L" + libdir, "-lsspbo", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm",
           "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe]
    done = subprocess.run(cmd, capture_output=True, text=True)
    assert done.returncode == 0, done.stderr
    return exe


def test_c_abi_from_plain_c(pkg, tmp_path):
    """The drop-in boundary used from a plain C program (tests/c_abi_smoke.c, gcc, no Python / torch in the process): one BO
    step through sspbo.h; every scoring engine must name the same winner and observe == eval_host."""
    import subprocess
    exe = _build_c_abi_smoke(tmp_path)
    done = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert done.returncode == 0, done.stdout + done.stderr
    lines = [l for l in done.stdout.splitlines() if l.startswith("engine ")]
    assert len(lines) == 4 and done.stdout.strip().endswith("c abi ok"), done.stdout
    idx = {l.split("index ")[1].split()[0] for l in lines}
    scores = [float(l.split("score ")[1].split()[0]) for l in lines]
    assert len(idx) == 1, done.stdout
    assert max(scores) - min(scores) <= RTOL_SCORE * abs(scores[0]), done.stdout


def test_agents_shape_contract(pkg):
    """The reference's per-agent smoke tests (tests/test_agents.py:40-70 SSPAgent, :162-200 trajectory, :207-246 multi-agent),
    same constructor calls and assertions, against the GPU agents."""
    rng = np.random.default_rng(0)
    xs = rng.uniform(B2[:, 0], B2[:, 1], size=(8, 2))
    ys = (-(xs[:, 0] ** 2 + xs[:, 1] - 11) ** 2 - (xs[:, 0] + xs[:, 1] ** 2 - 7) ** 2).reshape(-1, 1)
    space = pkg.HexagonalSSPSpace(2, ssp_dim=51, domain_bounds=B2, length_scale=1.0, rng=0)
    agt = pkg.SSPAgent(xs, ys, space, decoder_method="from-set", length_scale=1.0)
    mu, var, phi = agt.eval(xs)
    assert mu.shape[-1] == 8 and var.shape[-1] == 8
    min_func, jac = agt.acquisition_func()
    assert callable(min_func) and callable(jac)
    assert np.ravel(min_func(agt.encode(xs[:1]).ravel())).shape == (1,) and jac(agt.encode(xs[:1]).ravel()).shape == (agt.ssp_dim,)
    x_new, y_new = xs[:1] + 0.1, ys[:1]
    mu, var, _ = agt.eval(x_new)
    agt.update(x_new, y_new, float(var.flat[0]), step_num=1)
    x = np.array([[1.0, -2.0]])
    assert agt.decode(space.encode(x)).shape == x.shape
    assert agt.encode(xs).shape == (8, agt.ssp_dim)

    rng = np.random.default_rng(1)
    txs = rng.uniform(-5, 5, size=(5, 3, 2)).reshape(5, -1)
    tys = rng.uniform(-1, 0, size=(5, 1))
    tagt = pkg.SSPTrajectoryAgent(txs, tys, x_dim=2, traj_len=3, domain_bounds=B2, decoder_method="from-set", length_scale=1.0,
                                  ssp_dim=51)
    assert tagt.encode(txs).shape == (5, tagt.ssp_dim)
    assert tagt.eval(txs)[0].shape[-1] == 5
    assert callable(tagt.acquisition_func()[0])
    rng = np.random.default_rng(99)
    x_new, y_new = rng.uniform(-5, 5, size=(1, 6)), rng.uniform(-1, 0, size=(1, 1))
    mu, var, _ = tagt.eval(x_new)
    tagt.update(x_new, y_new, float(var.flat[0]), step_num=1)
    assert tagt.decode(tagt.encode(x_new)).shape == (6,)

    rng = np.random.default_rng(2)
    mxs = rng.uniform(-5, 5, size=(5, 3 * 2 * 2))
    mys = rng.uniform(-1, 0, size=(5, 1))
    magt = pkg.SSPMultiAgent(mxs, mys, n_agents=2, x_dim=2, traj_len=3, domain_bounds=B2, decoder_method="from-set",
                             length_scale=1.0, ssp_dim=51)
    assert magt.encode(mxs).shape == (5, magt.ssp_dim)
    assert magt.eval(mxs)[0].shape[-1] == 5
    assert callable(magt.acquisition_func()[0])
    rng = np.random.default_rng(99)
    x_new, y_new = rng.uniform(-5, 5, size=(1, 12)), rng.uniform(-1, 0, size=(1, 1))
    mu, var, _ = magt.eval(x_new)
    magt.update(x_new, y_new, float(var.flat[0]), step_num=1)
    assert magt.decode(magt.encode(x_new)).shape == (12,)


def test_maximize_api(pkg):
    """tests/test_bayesian_optimization.py:14-47 shape contract + README-style run (config C1, reduced)."""
    bounds = np.array([[-2.0, 2.0], [-2.0, 2.0]])
    f = lambda x: -(x[:, 0] ** 2 + x[:, 1] ** 2)
    xs0 = np.random.default_rng(0).uniform(-2, 2, (5, 2))
    for n_iter in (0, 1, 6):
        opt = pkg.BayesianOptimization(f=f, bounds=bounds)
        opt.maximize(init_points=(xs0, f(xs0).reshape(-1, 1)), n_iter=n_iter, agent_type="ssp-hex", ssp_dim=151,
                     length_scale=0.5, decoder_method="from-set", beta_ucb=10.0, gamma_c=0.0)
        assert opt.xs.shape == (5 + n_iter, 2) and opt.ys.shape == (5 + n_iter,)
        assert set(opt.max) == {"target", "params"} and len(opt.res) == 5 + n_iter
        assert opt.times.shape == (n_iter,)
    # the chosen points are the oracle's argmax over the same 100 x 100 grid, step by step
    pts = orc.grid_sample_points(bounds, 100)
    A = opt.agt.ssp_space.phase_matrix
    ref = orc.BLR(A.shape[0])
    phis0 = orc.encode(A, 0.5 * np.ones(2), xs0)
    ref.update(phis0, f(xs0).reshape(-1, 1))
    grid_phis = orc.encode(A, 0.5 * np.ones(2), pts)
    for t in range(6):
        score, idx = orc.score_and_argmax(grid_phis, ref, 10.0, 0.0)
        srt = np.sort(score)[::-1]
        if (srt[0] - srt[1]) / abs(srt[0]) > RTOL_SCORE:
            assert opt.acq_indices[t] == idx
            assert np.array_equal(opt.xs[5 + t], pts[idx])
        x_t = opt.xs[5 + t:6 + t]
        ref.update(orc.encode(A, 0.5 * np.ones(2), x_t), np.atleast_2d(f(x_t)))
    two_arg = lambda x, info: float(-(x ** 2).sum())
    opt = pkg.BayesianOptimization(f=two_arg, bounds=bounds)
    opt.maximize(init_points=3, n_iter=1, agent_type="ssp-rand", ssp_dim=51, length_scale=0.5, rng=0,
                 acq_candidates="uniform", n_candidates=5000)
    assert opt.xs.shape == (4, 2) and np.all(np.abs(opt.xs) <= 2)


def test_full_size_grid_properties(pkg):
    """BASELINE config C2 size (1000 x 1000 Branin grid, Hex d=151): size-independent properties of the fused
    path - sharding invariance of the argmax and agreement of the winner with the oracle on its neighbourhood."""
    bounds = np.array([[-5.0, 10.0], [0.0, 15.0]])
    sp = pkg.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=bounds, length_scale=1.0)
    xs0 = np.random.default_rng(0).uniform(bounds[:, 0], bounds[:, 1], (10, 2))
    a, b, c, r, s, t = 1, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6., 10., 1 / (8 * np.pi)
    branin = lambda x: -(a * (x[:, 1] - b * x[:, 0] ** 2 + c * x[:, 0] - r) ** 2 + s * (1 - t) * np.cos(x[:, 0]) + s)
    agt = pkg.SSPAgent(xs0, branin(xs0).reshape(-1, 1), sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0)
    eng = sp.engine
    axes = sp.grid_axes(1000, "grid")
    grid = eng.grid_points(axes)
    assert grid.shape == (1000000, 2)
    agt.best_candidate(grid)
    s_all, i_all = eng.best()
    first = True
    for k in range(4):                                           # 4 ragged shards
        lo, hi = k * 250007, min(1000000, (k + 1) * 250007)
        agt.best_candidate(grid[lo:hi], index0=lo, reset_best=first)
        first = False
    assert eng.best() == (s_all, i_all)
    lo = max(0, i_all - 2000)
    pts = grid[lo:lo + 4000].cpu().numpy()
    ref = orc.BLR(sp.ssp_dim)
    ref.update(orc.encode(sp.phase_matrix, np.ones(2), xs0), branin(xs0).reshape(-1, 1))
    score, _ = orc.score_and_argmax(orc.encode(sp.phase_matrix, np.ones(2), pts), ref, 10.0, 0.0)
    assert abs(score[i_all - lo] - s_all) <= RTOL_SCORE * abs(s_all)
    assert score.max() <= s_all * (1 + np.sign(s_all) * RTOL_SCORE) + 1e-12


def test_batched_independent_runs(pkg):
    """BASELINE config C4 in miniature: independent runs (own space seed, own posterior) stepped in lock-step over a
    shared Himmelblau grid; every run's picks must equal the oracle's arg-max for that run, step by step."""
    bounds = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    R, steps, per_dim = 6, 3, 40
    agents, refs, spaces = [], [], []
    for r in range(R):
        sp = pkg.RandomSSPSpace(2, ssp_dim=128, domain_bounds=bounds, length_scale=1.0, rng=r)   # d = 127
        xs0 = np.random.default_rng(100 + r).uniform(-5, 5, (6, 2))
        ys0 = himmelblau(xs0).reshape(-1, 1)
        agents.append(pkg.SSPAgent(xs0, ys0, sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0))
        ref = orc.BLR(sp.ssp_dim)
        ref.update(orc.encode(sp.phase_matrix, np.ones(2), xs0), ys0)
        refs.append(ref)
        spaces.append(sp)
    grid = orc.grid_sample_points(bounds, per_dim)
    runs = pkg.BatchedRuns(agents, grid, n_streams=3)
    for _ in range(steps):
        pts, ys, scores = runs.step(himmelblau)
        for r in range(R):
            A = spaces[r].phase_matrix
            score, idx = orc.score_and_argmax(orc.encode(A, np.ones(2), grid), refs[r], 10.0, 0.0)
            srt = np.sort(score)[::-1]
            if (srt[0] - srt[1]) > RTOL_SCORE * abs(srt[0]):
                assert np.array_equal(pts[r], grid[idx]), (r, pts[r], grid[idx])
            assert abs(scores[r] - score.max()) <= RTOL_SCORE * abs(score.max())
            refs[r].update(orc.encode(A, np.ones(2), pts[r:r + 1]), np.atleast_2d(ys[r]))
    for r in range(R):
        assert rel(agents[r].blr.m, refs[r].m) < 1e-8
