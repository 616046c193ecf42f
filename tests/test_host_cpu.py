"""CPU-only tests: C-ABI library exports, host-side mirrors of the reference constructors, sharding and the
packed-key reduction over gloo (world_size 2).  No compute call touches a GPU here."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ssp_bayesopt_b200 import _lib
    header = open(os.path.join(ROOT, "include", "sspbo.h")).read()
    declared = set(re.findall(r"\b(sspbo_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sspbo.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.sspbo_abi_version() == 1


def test_key_pack_orders_like_argmax():
    from ssp_bayesopt_b200 import _lib
    k = _lib.key_pack
    assert k(1.0, 5) > k(0.5, 0)                       # higher score wins
    assert k(1.0, 3) > k(1.0, 4)                       # ties: lowest index wins (np.argmax)
    assert k(-1.0, 0) > k(-2.0, 0) and k(0.0, 9) > k(-0.0, 9) or k(0.0, 9) >= k(-0.0, 9)
    assert k(float("nan"), 7) > k(float("inf"), 0)     # NaN counts as the maximum
    assert k(-1e30, 1) > 0
    s, i = _lib.key_unpack(k(-3.25, 4000000000))
    assert s == -3.25 and i == 4000000000


def test_phase_matrices_bit_identical_to_reference(golden):
    from ssp_bayesopt_b200 import sspspace
    g = golden("encode.npz")
    b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    assert np.array_equal(sspspace.RandomSSPSpace(2, ssp_dim=51, domain_bounds=b2, rng=0).phase_matrix, g["rand51_A"])
    assert np.array_equal(sspspace.RandomSSPSpace(6, ssp_dim=1024, rng=0).phase_matrix, g["rand1024_A"])
    hx = sspspace.HexagonalSSPSpace(2, ssp_dim=51, domain_bounds=b2, rng=0)
    assert np.array_equal(hx.phase_matrix, g["hex51_A"]) and hx.ssp_dim == 25
    assert np.array_equal(sspspace.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=b2).phase_matrix, g["hex151_A"])
    s = 2 * np.pi / np.sqrt(6)
    hs = sspspace.HexagonalSSPSpace(2, ssp_dim=151, scale_min=s - 0.5, scale_max=s + 0.5, length_scale=0.05)
    assert np.array_equal(hs.phase_matrix, g["hexsmall_A"])
    assert hs.length_scale.shape == (2, 1)
    hs.update_lengthscale(0.1)
    assert hs.length_scale.shape == (2,)               # sspspace.py:243-252 quirk kept
    with pytest.raises(AssertionError):
        sspspace.SSPSpace(2, 7, phase_matrix=np.zeros((7, 3)))


def test_host_algebra_matches_reference(golden):
    from ssp_bayesopt_b200 import sspspace
    g = golden("decode.npz")
    sp = sspspace.SSPSpace(2, g["A"].shape[0], phase_matrix=g["A"], domain_bounds=g["bounds"])
    assert np.array_equal(sp.invert(g["queries"]), g["invert"])
    assert sp.identity().shape == (1, sp.ssp_dim)
    np.testing.assert_allclose(np.linalg.norm(sp.normalize(3 * g["queries"]), axis=1), 1.0)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ssp_bayesopt_b200 import _lib, sspspace
    sp = sspspace.RandomSSPSpace(2, ssp_dim=51, rng=0)
    with pytest.raises(_lib.SspboError):
        sp.encode(np.zeros((1, 2)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ssp_bayesopt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_shard_range_partitions():
    from ssp_bayesopt_b200.dist import shard_range
    for total in (0, 1, 7, 1000, 10 ** 8):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as td
from ssp_bayesopt_b200 import _lib, dist
td.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
r = td.get_rank()
assert dist.rank_world() == (r, 2) and dist.is_distributed()
# rank 0 holds a positive score (top bit of the key set), rank 1 a higher one and a tie at a lower index
cases = [((0.5, 10), (0.75, 2000000)), ((1.0, 9), (1.0, 3)), ((-2.0, 1), (-1.0, 4000000000)), ((float("nan"), 8), (9e9, 0))]
for c in cases:
    s, i = c[r]
    key = _lib.key_pack(s, i)
    t = torch.tensor([key - (1 << 64) if key >= (1 << 63) else key], dtype=torch.int64)
    dist.allreduce_best(t)
    got = _lib.key_unpack(int(t.item()))
    want = max(_lib.key_pack(*c[0]), _lib.key_pack(*c[1]))
    assert (int(t.item()) & 0xFFFFFFFFFFFFFFFF) == want, (c, got)
# sharded argmax == global argmax on a synthetic score vector
g = torch.Generator().manual_seed(0)
scores = torch.randn(1001, generator=g)
s0, s1 = dist.shard_range(1001, r, 2)
loc = int(torch.argmax(scores[s0:s1])) + s0
key = _lib.key_pack(float(scores[loc]), loc)
t = torch.tensor([key - (1 << 64) if key >= (1 << 63) else key], dtype=torch.int64)
dist.allreduce_best(t)
assert _lib.key_unpack(int(t.item()))[1] == int(torch.argmax(scores))
# more than 2^32 candidates: shard-local indices + (key, shard start) pairs (SURVEY.md 8e)
starts = [0, 5_000_000_000]
for c, want in [(((0.5, 10), (0.75, 7)), (0.75, 5_000_000_007)), (((1.0, 4_000_000_000), (1.0, 3)), (1.0, 4_000_000_000)),
                (((float("nan"), 8), (9e9, 0)), None)]:
    s, i = c[r]
    key = _lib.key_pack(s, i)
    t = torch.tensor([key - (1 << 64) if key >= (1 << 63) else key], dtype=torch.int64)
    got = dist.allgather_best(t, starts[r])
    if want is None:
        assert got[0] != got[0] and got[1] == 8, got          # NaN wins, as in np.argmax
    else:
        assert got == want, (got, want)
# maximize(init_points=<int>) under torch.distributed: every rank must build its agent from rank 0's initial design (the
# Sobol samplers are unseeded) -- agent and space construction stubbed out, this is the host logic only
import numpy as np
from ssp_bayesopt_b200 import bayesian_optimization as bo
seen = {}
class _Agent:
    def __init__(self, xs, ys, space, **kw): seen["xs"], seen["ys"] = np.array(xs), np.array(ys)
    def update(self, xs, ys, s, n): seen["perimeter"] = np.array(xs)
bo.agents.SSPAgent = _Agent
bo.sspspace.HexagonalSSPSpace = lambda *a, **k: None
opt = bo.BayesianOptimization(f=lambda x: float(-np.sum(x ** 2)) + r, bounds=np.array([[-2.0, 2.0], [-1.0, 3.0]]))
opt.initialize_agent(init_points=6, agent_type="ssp-hex", num_perimeter_samples=3)
mine = torch.from_numpy(np.concatenate([seen["xs"].ravel(), seen["ys"].ravel(), seen["perimeter"].ravel()]))
both = [torch.empty_like(mine) for _ in range(2)]
td.all_gather(both, mine)
assert torch.equal(both[0], both[1]) and seen["xs"].shape == (6, 2) and seen["perimeter"].shape == (3, 2)
# dist.reduce_best with an engine that cannot use the peer-memory exchange (here: a stand-in without a GPU): the key is made
# final locally, max-reduced (the NCCL path of the product, gloo here) and read back; every rank gets the global winner
class _Eng:
    _exchange_on = False
    def __init__(self, score, idx):
        key = _lib.key_pack(score, idx)
        self._best = torch.tensor([key - (1 << 64) if key >= (1 << 63) else key], dtype=torch.int64)
        self.settled = False
    def settle(self): self.settled = True
    def best(self): return _lib.key_unpack(int(self._best.item()))
for c in [((0.5, 10), (0.75, 2000000)), ((1.0, 9), (1.0, 3)), ((-2.0, 1), (-1.0, 4000000000))]:
    e = _Eng(*c[r])
    got = dist.reduce_best(e)
    want = _lib.key_unpack(max(_lib.key_pack(*c[0]), _lib.key_pack(*c[1])))
    assert e.settled and got == want, (got, want)
td.destroy_process_group()
print("ok", r)
'''


def test_gloo_world2_key_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out


def test_cv_loss_dual_form_matches_reference_evaluations():
    """Host half of the trajectory agent's length-scale search: the dual-form K-fold loss computed from a Gram matrix (here
    formed from the oracle's phi on the CPU; the product forms it on the device) against the objective values scipy
    evaluated inside the live reference (tests/golden/lengthscale_cv.npz)."""
    import os
    from oracle import ssp_oracle as orc
    from ssp_bayesopt_b200.agents.ssp_traj_agent import cv_loss_from_gram
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lengthscale_cv.npz"))
    L, xd = 4, 2
    for x, f in list(zip(g["eval_x"], g["eval_f"]))[::6]:
        T = orc.encode(g["At"], np.array([x[1]]), np.linspace(0, L, L).reshape(-1, 1))
        phi = orc.traj_encode(g["Ax"], x[0] * np.ones(xd), T, g["trajs"], L, xd)
        np.testing.assert_allclose(cv_loss_from_gram(phi @ phi.T, g["tys"]), f, rtol=1e-9)
    # fold sizes of KFold(n_splits=min(n, 50)) for n > 50: 57 points -> 7 folds of 2, 43 of 1
    rng = np.random.default_rng(0)
    phi = rng.normal(size=(57, 9))
    assert np.isfinite(cv_loss_from_gram(phi @ phi.T, rng.normal(size=57)))


@pytest.mark.parametrize("tag", ["a3", "a2"])
def test_multi_agent_phase_table_reproduces_reference_encoding(tag):
    """Host half of the multi-agent encoder: tag vectors (SPSpace mirror) and the waypoint phase table handed to the device.
    Summing unit phasors with those offsets and synthesising phi on the CPU must give the live reference's encode
    (tests/golden/multi_agent.npz) - the device kernels evaluate exactly this sum."""
    import os
    import ssp_bayesopt_b200 as pkg
    from ssp_bayesopt_b200.agents.ssp_multi_agent import waypoint_phase_table
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multi_agent.npz"))
    n_agents, L, xd, ssp_dim, seed = (int(v) for v in g[f"{tag}_dims"])
    Ax, d = g[f"{tag}_Ax"], g[f"{tag}_Ax"].shape[0]
    K = (d - 1) // 2
    tags = pkg.SPSpace(n_agents, d, rng=seed)
    np.testing.assert_allclose(tags.vectors, g[f"{tag}_agent_sps"], atol=1e-15)
    np.testing.assert_allclose(tags.bind(tags.vectors, tags.inverse_vectors), np.tile(tags.identity(), (n_agents, 1)), atol=1e-12)
    tau, dc = waypoint_phase_table(np.linspace(1, L + 1, L), g[f"{tag}_At"], g[f"{tag}_ls_t"], tags.vectors)
    assert tau.shape == (L * n_agents, K) and set(np.unique(dc)) <= {-1, 1}
    x = g[f"{tag}_xc"].reshape(-1, L * n_agents, xd)                      # waypoints in (timestep, agent) order
    theta = np.einsum("kn,cwn->cwk", Ax[1:K + 1] / g[f"{tag}_ls_x"], x) + tau[None]
    spec = np.exp(1j * theta).sum(axis=1)                                 # (candidates, K)
    j = np.arange(d)[:, None] * np.arange(1, K + 1)[None, :] * (2 * np.pi / d)
    phi = (dc.sum() + 2 * (spec.real @ np.cos(j).T - spec.imag @ np.sin(j).T)) / d
    phi /= np.linalg.norm(phi, axis=1, keepdims=True)
    np.testing.assert_allclose(phi, g[f"{tag}_phi"], atol=1e-13)


def test_sinc_kernel_and_gp_length_scale_match_reference():
    """SincKernel (values and the hyper-parameter gradient handed to scikit-learn) and the GP length-scale fit of
    SSPAgent._optimize_lengthscale against the live reference (tests/golden/lengthscale_gp.npz).  Host-only."""
    import os
    from ssp_bayesopt_b200.agents.kernels import SincKernel
    from ssp_bayesopt_b200.agents.ssp_agent import SSPAgent
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lengthscale_gp.npz"))
    for i in (0, 1):
        K, dK = SincKernel(length_scale=float(g[f"k{i}_ls"]))(g["k_X"], eval_gradient=True)
        np.testing.assert_allclose(K, g[f"k{i}_K"], atol=1e-15)
        np.testing.assert_allclose(dK, g[f"k{i}_dK"], atol=1e-14)
    Ka, dKa = SincKernel(length_scale=np.array([1.0, 0.5]))(g["k_X"], eval_gradient=True)
    assert dKa.shape == (6, 6, 2) and np.allclose(np.diag(Ka), 1.0)
    for tag in ("a", "b"):
        ls = SSPAgent._optimize_lengthscale(None, g[f"{tag}_xs"], g[f"{tag}_ys"])
        np.testing.assert_allclose(ls, g[f"{tag}_length_scale"][:1], rtol=1e-6)


def test_c_abi_links_from_plain_c(tmp_path):
    """tests/c_abi_smoke.c compiles as C99 against include/sspbo.h and links against the built library with gcc (running it
    needs a GPU: tests/test_gpu_parity.py::test_c_abi_from_plain_c)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "ssp_bayesopt_b200", "lib", "libsspbo.so")
    if not os.path.exists(lib):
        pytest.skip("libsspbo.so not built")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(root, "include"), "-I" + os.path.join(cuda, "include"),
           os.path.join(root, "tests", "c_abi_smoke.c"), "-L" + os.path.dirname(lib), "-lsspbo", "-L" + os.path.join(cuda, "lib64"),
           "-lcudart", "-lm", "-o", str(tmp_path / "c_abi_smoke")]
    done = subprocess.run(cmd, capture_output=True, text=True)
    assert done.returncode == 0, done.stderr


def test_bayesian_optimization_surface_without_device():
    """Constructor contract of bayesian_optimization.py:22-66 and the `res` / `max` views (:324-333); no device needed."""
    import ssp_bayesopt_b200 as pkg
    bounds = np.array([[-1.0, 1.0], [0.0, 2.0]])
    with pytest.raises(AssertionError):
        pkg.BayesianOptimization(f=None, bounds=bounds)
    with pytest.raises(AssertionError):
        pkg.BayesianOptimization(f=lambda x: 0.0, bounds=None)
    with pytest.raises(AssertionError):
        pkg.BayesianOptimization(f=lambda x: 0.0, bounds=np.zeros((2, 3)))
    one = pkg.BayesianOptimization(f=lambda x: float(np.sum(x)), bounds=bounds)
    two = pkg.BayesianOptimization(f=lambda x, info: float(np.sum(x)) + len(info), bounds=bounds)
    assert one.target(np.ones((1, 2)), "ignored") == 2.0          # one positional argument: f(x)
    assert two.target(np.ones((1, 2)), "ab") == 4.0               # otherwise f(x, info_str)
    assert one.data_dim == 2 and one.xs is None and one.ys is None
    one.xs, one.ys = np.array([[0.0, 1.0], [0.5, 0.5]]), np.array([1.0, 3.0])
    assert one.max["target"] == 3.0 and np.array_equal(one.max["params"], one.xs[1])
    assert [r["target"] for r in one.res] == [1.0, 3.0]
    with pytest.raises(NotImplementedError):
        one.initialize_agent(init_points=(one.xs, one.ys.reshape(-1, 1)), agent_type="gp")


def test_domains_and_tag_space_shapes():
    """Initial-design samplers (agents/domains.py) and SPSpace corner cases, host only."""
    import ssp_bayesopt_b200 as pkg
    from ssp_bayesopt_b200.agents import domains
    b = np.array([[-2.0, 3.0], [0.0, 1.0]])
    pts = domains.BoundedDomain(b).sample(8)
    assert pts.shape == (8, 2) and np.all(pts >= b[:, 0]) and np.all(pts <= b[:, 1])
    tr = domains.TrajectoryDomain(5, 2, b).sample(4)
    assert tr.shape == (4, 10) and tr.min() >= -2.0 and tr.max() <= 3.0          # square box: [min low, max high]
    np.random.seed(0)
    goals = [np.array([1.0, 1.0]), np.array([-1.0, 0.5])]
    mt = domains.MultiTrajectoryDomain(3, 4, 2, b, goals).sample(8)
    assert mt.shape == (8, 4 * 2 * 3)
    last = mt.reshape(8, 4, 2, 3)[:, -1]                                           # last waypoints sit within 0.1 of a goal
    for i in range(3):
        assert np.all(np.abs(last[:, :, i] - goals[(2 * i + 1) % 2]) <= 0.1 + 1e-12)
    with pytest.raises(NotImplementedError):
        domains.BoundedDomain(b).sample(4, method="grid")
    one = pkg.SPSpace(1, 9)
    assert np.array_equal(one.vectors, one.identity()) and np.array_equal(one.inverse_vectors, one.identity())
    given = pkg.SPSpace(2, 9, vectors=np.eye(2, 9))
    assert np.array_equal(given.vectors, np.eye(2, 9))
    tags = pkg.SPSpace(4, 33, rng=np.random.default_rng(3))
    assert np.allclose(np.abs(np.fft.fft(tags.vectors, axis=-1)), 1.0)             # unitary


def test_sinc_kernel_reference_properties():
    """The reference's own kernel tests (tests/test_sinc_kernel.py:8-68) against this package's SincKernel."""
    from ssp_bayesopt_b200.agents.kernels import SincKernel
    from ssp_bayesopt_b200.agents.kernels.sinc_kernel import _sinc_prime
    xs = np.linspace(-1, 1, 10000)
    assert np.allclose(np.diff(np.sinc(xs)) / np.diff(xs), _sinc_prime(xs)[1:], atol=1e-3)      # :8-11
    k = SincKernel()
    a, b = np.random.default_rng(0).random((100, 2)), np.random.default_rng(1).random((40, 2))
    assert k(a, b).shape == (100, 40)                                                           # :14-19
    K = k(a, a)
    assert np.allclose(K, K.T) and np.allclose(np.diag(K), 1.0)                                 # :22-34
    pts = np.array([[0.0, 0.0], [0.5, 0.5]])
    K, dK = k(pts, eval_gradient=True)
    assert dK.shape == (2, 2, 1)                                                                # :37-41
    u = pts[:, None, :] - pts[None, :, :]
    expected = (_sinc_prime(u)[:, :, 0] * np.sinc(u)[:, :, 1] * (-u[:, :, 0])
                + np.sinc(u)[:, :, 0] * _sinc_prime(u)[:, :, 1] * (-u[:, :, 1]))
    assert np.allclose(expected.flatten(), dK.flatten())                                        # :44-53
    with pytest.raises(ValueError):
        k(a, b, eval_gradient=True)
    from sklearn.datasets import load_iris                                                      # :56-68
    from sklearn.gaussian_process import GaussianProcessClassifier
    from sklearn.gaussian_process.kernels import RBF
    X, y = load_iris(return_X_y=True)
    sinc_score = GaussianProcessClassifier(kernel=SincKernel(), random_state=0).fit(X, y).score(X, y)
    rbf_score = GaussianProcessClassifier(kernel=1.0 * RBF(1.0), random_state=0).fit(X, y).score(X, y)
    assert np.allclose(sinc_score, rbf_score, atol=2e-2), (rbf_score, sinc_score)


def test_vectorised_key_unpack_matches_c():
    """BatchedRuns unpacks thousands of keys per lock-step with numpy; same results as sspbo_key_unpack."""
    from ssp_bayesopt_b200 import _lib
    from ssp_bayesopt_b200.batched import _unpack_keys
    vals = [(1.5, 3), (-2.25, 4000000000), (0.0, 7), (-0.0, 9), (1e30, 0), (-1e-30, 12), (float("inf"), 5), (float("nan"), 8)]
    keys = np.array([_lib.key_pack(s, i) for s, i in vals], dtype=np.uint64).astype(np.int64)
    sc, ix = _unpack_keys(keys)
    for k, a, b in zip(keys, sc, ix):
        s, i = _lib.key_unpack(int(k))
        assert (a == s or (np.isnan(a) and np.isnan(s))) and b == i
    sc, ix = _unpack_keys(np.zeros(2, dtype=np.int64))
    assert np.all(np.isneginf(sc))


def test_host_streaming_piece_schedule():
    """engine.stream_pieces: the pieces of a host candidate set cover it exactly, never exceed the staging buffer, ramp up from
    a short first piece and down to a short last one, and no sliver is left where the ramps meet."""
    from ssp_bayesopt_b200.engine import stream_pieces
    chunk = 1 << 21
    for N in (1, 100, 40_000, (1 << 19), (1 << 19) + 1, 3_000_001, 12_500_000, 100_000_003):
        p = stream_pieces(N, chunk)
        assert p[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(p, p[1:])) and p[-1][0] + p[-1][1] == N
        assert all(0 < n <= chunk for _, n in p)
        if N >= 1 << 20:
            assert all(n >= 4096 for _, n in p)
    sizes = [n for _, n in stream_pieces(12_500_000, chunk)]
    assert sizes[0] == 1 << 19 and sizes[-1] == 1 << 19 and max(sizes) == chunk
    k = sizes.index(max(sizes))
    assert sizes[:k + 1] == sorted(sizes[:k + 1]) and sizes[-4:] == sorted(sizes[-4:], reverse=True)
    assert [n for _, n in stream_pieces(5000, chunk, first_piece=1000, growth=2.0)] == [1000, 3000, 1000]
