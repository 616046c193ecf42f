"""BayesianOptimization.maximize with the candidate set sharded over the ranks of torch.distributed (NCCL, one process per
GPU) against the same run in a single process: the chosen points must be identical step by step.
    python tools/dist_maximize.py single out.npy
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_maximize.py dist out.npy"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
from oracle import ssp_oracle as orc   # only for the synthetic objective (Hartmann-6), shared with the tests


def main():
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "dist":
        import torch.distributed as td
        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    b6 = np.array([[0.0, 1.0]] * 6)
    xs0 = np.random.default_rng(0).uniform(0, 1, (12, 6))
    f = lambda x: orc.hartmann6(np.atleast_2d(x))
    out = {}
    for cands in ("uniform", "explicit"):
        kw = dict(acq_candidates="uniform", n_candidates=3_000_001, candidate_seed=3)
        if cands == "explicit":
            kw = dict(acq_candidates=np.random.default_rng(5).uniform(0, 1, (200_003, 6)))
        opt = pkg.BayesianOptimization(f=f, bounds=b6)
        opt.maximize(init_points=(xs0, f(xs0).reshape(-1, 1)), n_iter=6, agent_type="ssp-rand", ssp_dim=1024, rng=0,
                     length_scale=0.25, gamma_c=1.0, beta_ucb=10.0, **kw)
        out[cands] = np.concatenate([opt.xs[12:], opt.acq_indices.reshape(-1, 1).astype(np.float64)], axis=1)
    got = np.stack([out["uniform"], out["explicit"]])
    if mode == "single":
        np.save(path, got)
        print("single-process run saved", got.shape)
        return
    import torch.distributed as td
    rank, world = td.get_rank(), td.get_world_size()
    t = torch.as_tensor(got, device="cuda")
    gathered = [torch.empty_like(t) for _ in range(world)]
    td.all_gather(gathered, t)
    same_across = all(torch.equal(g, gathered[0]) for g in gathered)
    want = np.load(path)
    if rank == 0:
        print(f"world {world}: ranks agree {same_across}; equals single-process run: points {np.array_equal(got[..., :6], want[..., :6])}"
              f" indices {np.array_equal(got[..., 6], want[..., 6])}")
        assert same_across and np.array_equal(got, want)
    td.destroy_process_group()


if __name__ == "__main__":
    main()
