"""Size-independent properties of the oracle (the checker of the CUDA path), hypothesis-driven on small random problems.
CPU only.  They are the invariants the GPU tests rely on at sizes where no reference output exists."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import ssp_oracle as orc

SETTINGS = dict(max_examples=25, deadline=None, derandomize=True)      # same examples on every run


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), n=st.integers(1, 4), half=st.integers(2, 20))
def test_encode_unit_norm_and_closed_form(seed, n, half):
    """phi has unit norm (tests/test_ssps.py:31-37) and the IFFT statement equals the closed form used by the kernels
    (phi = B psi, psi = [1, cos theta, sin theta])."""
    rng = np.random.default_rng(seed)
    A = orc.conjugate_symmetric_rows(rng.uniform(-np.pi, np.pi, (half, n)))
    ls = rng.uniform(0.2, 2.0, n)
    x = rng.uniform(-3, 3, (5, n))
    phi = orc.encode(A, ls, x)
    assert phi.shape == (5, 2 * half + 1)
    np.testing.assert_allclose(np.linalg.norm(phi, axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(phi, orc.encode_closed_form(A, ls, x), atol=1e-12)


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), d=st.integers(3, 41))
def test_bind_invert_identities(seed, d):
    """a (*) inv(a) = identity for unitary a; binding commutes and is associative (sspspace.py:551-560)."""
    rng = np.random.default_rng(seed)
    a, b, c = (orc.sp_vectors(2, d, rng)[0] for _ in range(3))
    ident = np.zeros(d)
    ident[0] = 1.0
    np.testing.assert_allclose(orc.bind(a, orc.invert(a))[0], ident, atol=1e-12)
    np.testing.assert_allclose(orc.bind(a, b), orc.bind(b, a), atol=1e-12)
    np.testing.assert_allclose(orc.bind(orc.bind(a, b), c), orc.bind(a, orc.bind(b, c)), atol=1e-12)


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), d=st.integers(4, 30), k=st.integers(1, 12))
def test_blr_sequential_equals_batch_and_low_rank_form(seed, d, k):
    """Sequential rank-one updates equal the batch update (tests/test_blr.py:72-92), and the posterior covariance has the
    low-rank form the tensor-core scorer uses: S = I/alpha - V^T V with one row of V per observation."""
    rng = np.random.default_rng(seed)
    phis = rng.normal(size=(k, d)) / np.sqrt(d)
    ts = rng.normal(size=(k, 1))
    batch, seq = orc.BLR(d, beta=2.0), orc.BLR(d, beta=2.0)
    batch.update(phis, ts)
    rows = []
    for i in range(k):
        y = seq.S @ phis[i]
        rows.append(y * np.sqrt(seq.beta / (1.0 + seq.beta * phis[i] @ y)))
        seq.update(phis[i:i + 1], ts[i:i + 1])
    np.testing.assert_allclose(seq.S, batch.S, atol=1e-8)
    np.testing.assert_allclose(seq.m, batch.m, atol=1e-8)
    V = np.stack(rows)
    np.testing.assert_allclose(seq.S, np.eye(d) / 0.01 - V.T @ V, atol=1e-8)      # alpha = 0.01 (blr.py:27)


@settings(**SETTINGS)
@given(n=st.integers(2, 120), splits=st.integers(2, 50))
def test_kfold_indices_partition(n, splits):
    splits = min(splits, n)
    folds = list(orc.kfold_indices(n, splits))
    assert len(folds) == splits
    tests = np.concatenate([te for _, te in folds])
    assert np.array_equal(tests, np.arange(n))                         # contiguous, unshuffled, each point tested once
    sizes = [len(te) for _, te in folds]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    for tr, te in folds:
        assert len(tr) + len(te) == n and not set(tr) & set(te)


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), count=st.integers(1, 200), start=st.integers(0, 2 ** 40), n=st.integers(1, 8))
def test_counter_uniform_is_a_pure_function_of_the_index(seed, count, start, n):
    """Any index range of the candidate stream can be regenerated anywhere (sharding over GPUs relies on it)."""
    whole = orc.counter_uniform(seed, start, count, n)
    cut = count // 2
    assert np.array_equal(whole[cut:], orc.counter_uniform(seed, start + cut, count - cut, n))
    assert whole.dtype == np.float32 and whole.min() >= 0.0 and whole.max() < 1.0


def test_argmax_conventions():
    """First index on ties and NaN counts as the maximum (np.argmax; sspspace.py:356): the packed-key order of the device."""
    phis = np.eye(3)
    model = orc.BLR(3, beta=1.0)
    score, idx = orc.score_and_argmax(np.vstack([phis, phis]), model, 1.0, 0.0)
    assert idx == 0 and np.allclose(score, score[0])
    assert int(np.argmax(np.array([1.0, np.nan, 5.0]))) == 1


def test_encode_and_deriv_is_the_derivative_of_encode():
    """oracle.encode_and_deriv (sspspace.py:278-306, the derivative its docstring describes) against central differences of
    `encode`, hexagonal and random spaces, per-axis length scales."""
    rng = np.random.default_rng(3)
    for A, n in ((orc.hex_phase_matrix(2, 51), 2), (orc.random_phase_matrix(3, 64, rng=1), 3)):
        x = rng.uniform(-3, 3, (6, n))
        ls = rng.uniform(0.5, 2.0, n)
        phi, grad = orc.encode_and_deriv(A, ls, x)
        assert grad.shape == (6, A.shape[0], n)
        np.testing.assert_allclose(phi, orc.encode(A, ls, x), atol=1e-15)
        h = 1e-6
        for a in range(n):
            e = np.zeros(n); e[a] = h
            fd = (orc.encode(A, ls, x + e) - orc.encode(A, ls, x - e)) / (2 * h)
            np.testing.assert_allclose(grad[:, :, a], fd, atol=5e-9)


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), n_agents=st.integers(1, 3), L=st.integers(1, 4))
def test_multi_agent_encode_structure(seed, n_agents, L):
    """Multi-agent encoder (ssp_multi_agent.py:166-186): unit norm; identical per-agent spaces give the shared-space encoding;
    and the spectrum of the un-normalised sum is the sum over the (timestep, agent) waypoints of unit phasors
    exp(i (A_i x_ji / l_i + phase(T_j) + phase(a_i))) -- the identity the device encoder evaluates (one frequency table per
    agent, one phase-offset row per waypoint)."""
    rng = np.random.default_rng(seed)
    xd, ssp_dim = 2, 31
    As = [orc.random_phase_matrix(xd, ssp_dim, rng=seed + i) for i in range(n_agents)]
    lss = [rng.uniform(0.5, 2.0) * np.ones(xd) for _ in range(n_agents)]
    At = orc.random_phase_matrix(1, ssp_dim, rng=seed + 17)
    T = orc.traj_timestep_ssps(At, np.array([1.3]), L)
    tags = orc.sp_vectors(n_agents, ssp_dim, seed)
    x = rng.uniform(-2, 2, (5, L * n_agents * xd))
    phi = orc.multi_agent_encode(As, lss, T, tags, x, L, n_agents, xd)
    np.testing.assert_allclose(np.linalg.norm(phi, axis=1), 1.0, atol=1e-12)
    same = orc.multi_agent_encode([As[0]] * n_agents, [lss[0]] * n_agents, T, tags, x, L, n_agents, xd)
    shared = orc.multi_agent_encode(As[0], lss[0], T, tags, x, L, n_agents, xd)
    np.testing.assert_allclose(same, shared, atol=1e-14)
    K = (ssp_dim - 1) // 2
    pts = x.reshape(-1, L, n_agents, xd)
    tag_ph = np.angle(np.fft.fft(tags, axis=-1))[:, 1:K + 1]
    t_ph = np.angle(np.fft.fft(T, axis=-1))[:, 1:K + 1]
    spec = np.zeros((x.shape[0], K), dtype=complex)
    for j in range(L):
        for i in range(n_agents):
            theta = pts[:, j, i, :] @ (As[i][1:K + 1] / lss[i]).T
            spec += np.exp(1j * (theta + t_ph[j] + tag_ph[i]))
    un = phi * 0
    for s in range(x.shape[0]):                                           # the same sum through explicit binds, un-normalised
        acc = np.zeros(ssp_dim)
        for i in range(n_agents):
            si = np.zeros(ssp_dim)
            for j in range(L):
                si = si + orc.bind(T[j], orc.encode(As[i], lss[i], pts[s:s + 1, j, i, :])[0])
            acc = acc + orc.bind(tags[i], si)
        un[s] = acc
    np.testing.assert_allclose(np.fft.fft(un, axis=-1)[:, 1:K + 1], spec, atol=1e-10)
