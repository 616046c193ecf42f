#!/usr/bin/env python
"""Benchmark of the SSP-BO hot path: fused encode + BLR-UCB + argmax candidates/s at d=1024 (-> 1023).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[2] ("C3"): RandomSSPSpace(n=6, ssp_dim=1024 -> d=1023, K=511), bounds [0,1]^6,
length_scale 0.25, rng=0; posterior after 10 + 50 Hartmann-6 observations (rank r = 10 + t, t = 50); 1e8
counter-based uniform float32 candidates per BO step, sharded over 8 GPUs = 1.25e7 candidates per GPU per step (weak
scaling: the per-GPU shard is fixed, so N GPUs score N * 1.25e7 candidates per step).  One step = score every local
candidate with the fused kernel, max-reduce the packed (score, index) key over ranks, read the winner, evaluate Hartmann-6
on it, apply the float64 rank-one posterior update -- and then put the rank-60 posterior back (a device-to-device import of
the saved state, inside the timed step), so that the posterior rank, which the kernel's rate depends on, is the same in
every step whatever --steps is.  `by_rank` gives the kernel-only rate at ranks 10 ... 1000 (full rank).

`value`   candidates/s, whole job, candidates resident in HBM, device-timed (CUDA events, max over ranks).
`e2e`     same metric through the public API (`SSPAgent.best_candidate`) from pinned HOST candidates: the H2D copy of the
          shard and the D2H read of the winner are inside the timed region every step (all --steps of them).
`roofline` the dominant kernel (the tcgen05 scorer): `achieved` = tensor flops it really issues per launch (3 split-fp16
          products x 2 x 1024 x BN per candidate; BN = operand rows) / its average launch duration, measured with CUDA events on
          the launching stream inside the timed region (sspbo_set_timing); `peak` = MEASURED_PEAKS.json sustained bf16.  The
          dense-form algorithmic figure of SURVEY.md 8d (2 d^2 + 2 d + 2 K n flops per candidate) is kept under
          `algorithmic_dense`: the low-rank form does less work than that, so it is not a utilisation figure.  `traffic` =
          dram bytes per launch from the committed ncu capture (profiles/r2_roofline_traffic.json), null when absent.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference (oracle/_ref, copied by oracle/build_ref.py) --
          SSPAgent.eval + argmax over chunks -- on the host cores; the numpy oracle port when that copy is absent.
"""
import os
import sys

if "reference" in sys.argv:
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count())

import argparse
import contextlib
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_REQ, N_DIM, LS, LAM, GAMMA = 1024, 6, 0.25, 10.0, 0.0
PER_GPU = 12_500_000
N_OBS = 60
METRIC = "fused encode+UCB+argmax candidates/sec at d=1024"
WORKLOAD = ("C3: RandomSSPSpace n=6 ssp_dim=1024 (d=1023,K=511) ls=0.25 rng=0, Hartmann-6 posterior after 10+50 "
            "observations (rank 60, restored every step), counter-based uniform f32 candidates")


def f_alg(d, K, n):
    return 2.0 * d * d + 2.0 * d + 2.0 * K * n


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(tflops=float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1590.0))),
                    tflops_burst=float(j.get("bf16_tflops", 1590.0)), hbm=float(j.get("hbm_gbs", 6650.0)),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region: NVML (a few microseconds per sample) when available,
    nvidia-smi otherwise."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False       # samples: (sm_mhz, max_mhz, [4 flags])
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        flags = [bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4)]   # HwSlowdown, HwThermal, SwThermal, SwPowerCap
        self.samples.append((float(sm), float(mx), flags))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        if len(parts) >= 6:
            self.samples.append((float(parts[0]), float(parts[1]), [p.lower().startswith("active") for p in parts[2:6]]))

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.01)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2][i] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def problem_setup():
    """Seeded C3 posterior inputs, identical on every rank and for both arms."""
    from oracle import ssp_oracle as orc          # only for the objective (Hartmann-6) and the shared generator
    X = np.random.default_rng(1).uniform(0, 1, (N_OBS, N_DIM))
    y = orc.hartmann6(X).reshape(-1, 1)
    return X, y, orc


# ------------------------------------------------------------------------------------------ CPU arm
def blas_info():
    """threadpoolctl view of the BLAS / OpenMP pools numpy runs on (BASELINE.md section 3 asks for it next to the number)."""
    try:
        from threadpoolctl import threadpool_info
        return [{k: i.get(k) for k in ("user_api", "internal_api", "version", "num_threads", "threading_layer")}
                for i in threadpool_info()]
    except Exception as e:                                   # pragma: no cover
        return [{"error": str(e)}]


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return None


class CpuArm:
    """The reference path on the host: C3 posterior after N_OBS observations, then per step `eval(X_chunk)` -> score ->
    running argmax over chunks of the candidate sample.  kind "reference": the unmodified package from oracle/_ref
    (SSPAgent.eval, agents/ssp_agent.py:125-137; posterior built by its own BLR.update); kind "port": oracle/ssp_oracle.py."""

    def __init__(self):
        from oracle import ref_loader
        X, y, orc = problem_setup()
        self.orc = orc
        self.ref = ref_loader.load()
        b6 = np.array([[0.0, 1.0]] * N_DIM)
        if self.ref is not None:
            self.kind = "reference"
            with contextlib.redirect_stdout(sys.stderr):
                sp = self.ref.sspspace.RandomSSPSpace(N_DIM, ssp_dim=D_REQ, domain_bounds=b6, length_scale=LS, rng=0)
                self.agt = self.ref.agents.SSPAgent(X[:10], y[:10], sp, decoder_method="from-set", gamma_c=0.0, beta_ucb=LAM,
                                                    length_scale=LS)
                # the remaining observations through the reference's own update (agents/ssp_agent.py:187-224 -> blr.py:39-82,
                # O(d^3) per observation as written there: ~30 ms each at d = 1023)
                for i in range(10, N_OBS):
                    self.agt.update(X[i:i + 1], y[i:i + 1], 0.0, step_num=i)
            self.d = sp.ssp_dim
        else:
            self.kind = "port"
            A = orc.random_phase_matrix(N_DIM, D_REQ, rng=0)
            ls = LS * np.ones(N_DIM)
            blr = orc.BLR(A.shape[0])
            blr.update(orc.encode(A, ls, X[:10]), y[:10])
            S, beta = blr.S.copy(), blr.beta
            bvec = beta * (orc.encode(A, ls, X[:10]).T @ y[:10])
            for i in range(10, N_OBS):                     # O(d^2) form of blr.py:63-75 to keep the setup short
                phi = orc.encode(A, ls, X[i:i + 1])[0]
                v = S @ phi
                S = S - np.outer(v, v) * (beta / (1 + beta * (phi @ v)))
                bvec = bvec + beta * phi[:, None] * y[i]
            blr.S, blr.m = S, S @ bvec
            self.A, self.ls, self.blr, self.d = A, ls, blr, A.shape[0]

    def scores(self, xc):
        """score = mu + acq of rows xc (float64), the quantity the fused kernel arg-maxes"""
        if self.kind == "reference":
            mu, var, acq = self.agt.eval(xc)
            return np.ravel(mu) + np.ravel(acq)
        score, _ = self.orc.score_and_argmax(self.orc.encode(self.A, self.ls, xc), self.blr, LAM, GAMMA)
        return score

    def step(self, start, sample, chunk):
        """one pass over rows [start, start + sample) of the counter-based candidate stream; returns (seconds, score, index)"""
        xc = self.orc.counter_uniform(0, start, sample, N_DIM).astype(np.float64)
        t0 = time.perf_counter()
        best, best_i = -np.inf, -1
        for c0 in range(0, sample, chunk):
            sc = self.scores(xc[c0:c0 + chunk])
            i = int(np.argmax(sc))
            if sc[i] > best:
                best, best_i = float(sc[i]), start + c0 + i
        return time.perf_counter() - t0, best, best_i

    def rate(self, sample, chunk, steps=1, warmup=0):
        times = []
        for it in range(warmup + steps):
            dt, _, _ = self.step(it * sample, sample, chunk)
            if it >= warmup:
                times.append(dt)
        per_step = float(np.mean(times))
        return sample / per_step, per_step

    def describe(self, sample, chunk, threads):
        what = ("the unmodified reference package (oracle/_ref: SSPAgent.eval + np.argmax, agents/ssp_agent.py:125-137, "
                "sspspace.py:254-276, blr.py:84-97)" if self.kind == "reference" else
                "oracle/ssp_oracle.py, the numpy float64 restatement pinned to the live reference by tests/golden "
                "(oracle/_ref is absent)")
        return f"{sample} candidates/step of the C3 shard in chunks of {chunk}, {what}, numpy/OpenBLAS with {threads} threads"


def one_thread_rate(arm, sample, chunk):
    """the same pass with the BLAS pool limited to one thread (encode is mostly single-threaded)"""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            return arm.rate(sample, chunk, steps=1, warmup=0)[0]
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)                     # torchrun's OMP_NUM_THREADS=1 must not shrink the BLAS pool
    except Exception:
        pass
    arm = CpuArm()
    sample, chunk = args.cpu_sample, 10000
    rate, per_step = arm.rate(sample, chunk, steps=args.steps, warmup=args.warmup)
    threads = max([i.get("num_threads") or 1 for i in blas_info()] + [1])
    one = one_thread_rate(arm, min(sample, 20000), chunk)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "candidates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "candidates_per_step": sample, "posterior_rank": N_OBS,
                   "note": "bounded sample of the 1.25e7-per-GPU shard, same seeds / phase matrix / posterior / candidate stream"},
        "cpu_baseline": {"value": rate, "unit": "candidates/s", "cores": threads, "kind": arm.kind,
                         "sample": arm.describe(sample, chunk, threads), "host_cores": cores, "cpu_model": cpu_model(),
                         "threadpool_info": blas_info(), "one_thread_value": one},
        "e2e": {"value": rate, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "roofline": None,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def executed_flops_per_candidate(engine, rank, d):
    """Tensor flops the scorer really issues per candidate, in fp16-MMA equivalents (an e4m3 MMA of twice the K counts as one
    fp16 MMA: same tensor-pipe time).  Returns (flops, description)."""
    kc = ((d + 63) // 64) * 64
    n_kb = kc // 64
    if engine == "umma-lowrank":
        bn16 = -(-(rank + 1) // 16) * 16
        if bn16 <= 128:                                      # split fp16, merged hi|lo operand: N = 2 BN and N = BN MMAs per K step
            return 3 * 2.0 * kc * bn16, f"low rank, split fp16: 3 x 2 x {kc} x {bn16}"
        nc = -(-rank // 256)
        bn = -(-(-(-rank // nc)) // 16) * 16
        return 2 * 2.0 * kc * bn * nc, f"low rank, fp16 + e4m3: (1 + 2 x 1/2) x 2 x {kc} x {nc}x{bn}"
    if engine == "umma-factor-f8":
        nc = -(-d // 256)
        bn = -(-(-(-d // nc)) // 16) * 16
        blocks = sum(n_kb - (c * bn) // 64 for c in range(nc))     # chunk c starts at factor column ~c*bn (upper bound on the skip)
        return 2 * 2.0 * 64 * bn * blocks, f"triangular factor, fp16 + e4m3: 2 x 2 x 64 x {bn} x {blocks} (chunk, k-block) pairs"
    if engine == "umma":
        return 3 * 2.0 * kc * 1024 * 0.56, "dense triangular factor, split fp16: 3 x 0.56 x 2 x 1024^2"
    return None, None


def committed_traffic():
    """dram bytes per scorer launch from the committed ncu --set full capture of this round (written by
    profiles/raw_metrics.py --json), or None"""
    p = os.path.join(ROOT, "profiles", "r2_roofline_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def run_native(args):
    import torch
    import torch.distributed as td
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ssp_bayesopt_b200 as pkg
    from ssp_bayesopt_b200 import _lib, dist

    X, y, orc = problem_setup()
    b6 = np.array([[0.0, 1.0]] * N_DIM)
    sp = pkg.RandomSSPSpace(N_DIM, ssp_dim=D_REQ, domain_bounds=b6, length_scale=LS, rng=0, device=local)
    # 10 initial observations (they fix beta, blr.py:54-59), then 50 sequential ones: the same construction as the CPU arm
    agt = pkg.SSPAgent(X[:10], y[:10], sp, decoder_method="from-set", gamma_c=0.0, beta_ucb=LAM, length_scale=LS,
                       score_engine={"auto": _lib.ENGINE_AUTO, "simt": _lib.ENGINE_SIMT, "umma": _lib.ENGINE_UMMA}[args.engine])
    agt.update(X[10:], y[10:], 0.0)
    eng = agt._scoring_engine()
    d, K = sp.ssp_dim, (sp.ssp_dim - 1) // 2
    n_local = args.per_gpu
    total = n_local * world
    start = rank * n_local
    cand = eng.fill_uniform(0, start, n_local)                    # resident shard, 300 MB > L2 (126 MB)
    chunk = args.chunk
    saved, beta0 = eng.blr_export(), eng.blr_beta()               # the rank-N_OBS posterior every step starts from
    eng.blr_checkpoint()                                          # ... and the cheap way back to it (low-rank form: one small kernel)
    torch.cuda.synchronize()

    def select(x_dev):
        first = True
        for c0 in range(0, n_local, chunk):
            agt.best_candidate(x_dev[c0:c0 + chunk], index0=start + c0, reset_best=first)
            first = False
        # one rank: counters + the 8-byte key in one synchronising read-back; several: first the key exchange over peer memory
        # (one kernel writing into every peer's slots; NCCL all-reduce of the key if peer access is not available)
        return dist.reduce_best(eng)

    def winner_row(gidx):
        u = orc.counter_uniform(0, gidx, 1, N_DIM)                # any rank can regenerate any candidate
        return u.astype(np.float64)

    def observe_and_restore(gidx):
        x_t = winner_row(gidx)
        agt.update(x_t, np.atleast_2d(orc.hartmann6(x_t)), np.array([0.0]))   # encode x_t + float64 rank-one update on device
        if not eng.blr_rollback():                                # back to rank N_OBS: every step scores the same posterior
            eng.blr_import(saved, beta0)                          # (the dense state absorbed the observation: full restore)
            eng.blr_checkpoint()

    def bo_step(x_dev):
        score, gidx = select(x_dev)
        observe_and_restore(gidx)
        return score, gidx

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        bo_step(cand)
    barrier()
    eng.score_stats(reset=True)
    eng.set_timing(True)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    last = None
    for i in range(args.steps):
        last = bo_step(cand)
    t_end.record()
    barrier()
    sampler.stop_flag = True
    launches = eng.launch_count() - launches0
    ms_total = t_start.elapsed_time(t_end)
    ktime = eng.timing_read(reset=True)                           # CUDA events around every scorer launch of the timed region
    eng.set_timing(False)
    engine_used = eng.last_engine()
    info = eng.lowrank_info()

    # ---- end-to-end timing: pinned host candidates -> H2D -> fused score -> D2H key, through the public API ----
    host = torch.empty((n_local, N_DIM), dtype=torch.float32).pin_memory()
    host.copy_(cand)
    e2e_steps = args.steps
    def e2e_step():
        # public API on HOST candidates: SSPAgent.best_candidate streams the pinned shard to the GPU in chunks
        # (H2D inside the timed region, overlapped with the scoring), then the 8-byte key is read back
        agt.best_candidate(host, index0=start)
        score, gidx = dist.reduce_best(eng)
        observe_and_restore(gidx)
        return score, gidx
    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        last_e2e = e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    # ---- the same step with the candidate set described instead of shipped (counter-based stream generated in the kernel) ----
    def gen_step():
        eng.score_generated("uniform", n_local, float(agt.var_weight), agt._gamma(), index0=start, seed=0)
        score, gidx = dist.reduce_best(eng)
        observe_and_restore(gidx)
        return score, gidx
    gen_step()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        last_gen = gen_step()
    g1.record()
    barrier()
    gen_ms = g0.elapsed_time(g1)

    # max over ranks
    stats = torch.tensor([ms_total, ktime["scorer_ms"], ktime["flagged_ms"], e2e_ms, gen_ms], dtype=torch.float64, device=cand.device)
    if world > 1:
        td.all_reduce(stats, op=td.ReduceOp.MAX)
    ms_total, scorer_ms, flagged_ms, e2e_ms, gen_ms = [float(v) for v in stats.cpu()]
    sampler.join(timeout=2)

    solo = world == 1 and rank == 0
    sweep = rank_sweep(pkg, sp, cand, orc, X, y, world, td if world > 1 else None) if not args.no_sweep else None
    enc = encode_only(eng, cand, load_peaks()) if (solo and not args.no_sweep) else None
    upd = update_cost(agt, sp, orc) if (solo and not args.no_sweep) else None
    asc = acq_ascent_cost(agt, sp, orc) if (solo and not args.no_sweep) else None
    c2 = bo_step_c2(pkg) if (solo and not args.no_c2) else None
    side = side_configs() if (solo and not args.no_side) else None
    if rank == 0:
        peaks = load_peaks()
        ms_per_step = ms_total / args.steps
        value = total / (ms_per_step * 1e-3)
        calls = max(int(ktime["calls"]), 1)
        launch_ms = scorer_ms / calls                             # average duration of one scorer launch (CUDA events)
        launch_cands = n_local * args.steps / calls
        ex_flops, ex_what = executed_flops_per_candidate(engine_used, info["rank"], d)
        achieved = launch_cands * ex_flops / (launch_ms * 1e-3) / 1e12 if ex_flops else None
        alg = f_alg(d, K, N_DIM)
        alg_tf = launch_cands * alg / (launch_ms * 1e-3) / 1e12
        traffic = committed_traffic()
        # parity of the winners of the last steps with the CPU arm (outside every timed region)
        check, cpu = None, None
        if world == 1:
            arm = CpuArm()
            cpu_rate, cpu_step = arm.rate(args.cpu_sample, 10000, steps=1, warmup=0)
            threads = max([i.get("num_threads") or 1 for i in blas_info()] + [1])
            cpu = {"value": cpu_rate, "unit": "candidates/s", "cores": threads, "kind": arm.kind,
                   "sample": arm.describe(args.cpu_sample, 10000, threads), "host_cores": os.cpu_count(), "cpu_model": cpu_model(),
                   "one_thread_value": one_thread_rate(arm, 20000, 10000)}
            check = winner_check(arm, orc, [("resident", last), ("e2e", last_e2e), ("generated", last_gen)], n_local)
        line = {
            "metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": ("f16 hi/lo split (+ e4m3 corrections above 128 operand rows) mma / f32 accumulate / fixed-point phase / "
                      "f64 posterior" if engine_used.startswith("umma") else "f32 contraction / f64 phase+posterior"),
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "candidates_per_gpu_per_step": n_local, "candidates_per_step": total, "chunk": chunk,
                       "engine": engine_used, "posterior_rank": info["rank"], "phase32": info["phase32"],
                       "parallelism": f"candidate-shard x{world}",
                       "key_reduction": ("none (one rank)" if world == 1 else
                                         "peer-memory exchange kernel (sspbo_dist.cu)" if eng._exchange_on else "NCCL all_reduce(MAX)"),
                       "l2": "inputs (300 MB candidate shard) larger than L2; posterior operands rewritten every step",
                       "last_winner": {"score": last[0], "index": last[1]}},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["tflops"]) if achieved else None,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                         "traffic_source": (traffic or {}).get("source"),
                         "kernel": {"umma-lowrank": "k_score_lr<6,2,0> / k_score_f8<6,2,0,0>", "umma-factor-f8": "k_score_f8<6,2,0,0>",
                                    "umma": "k_score_umma"}.get(engine_used, engine_used),
                         "launch_candidates": launch_cands, "launch_ms": launch_ms, "launches_timed": calls,
                         "flagged_pass_ms_per_launch": flagged_ms / calls,
                         "executed_flops_per_candidate": ex_flops, "executed_what": ex_what,
                         "frac_of_burst_peak": (achieved / peaks["tflops_burst"]) if achieved else None,
                         "algorithmic_dense": {"flops_per_candidate": alg, "tflops": alg_tf, "frac": alg_tf / peaks["tflops"],
                                               "note": "SURVEY 8d dense-form work / launch time; above 1 because the low-rank form "
                                                       "does less work, not a utilisation figure"},
                         "binding": (traffic or {}).get("binding"),
                         "note": f"achieved = tensor flops issued per launch / average launch duration over {calls} launches "
                                 f"(CUDA events on the launching stream); peak {peaks['source']}"},
            "by_rank": sweep,
            "encode_only": enc,
            "posterior_update": upd,
            "acq_ascent": asc,
            "cpu_baseline": cpu,
            "e2e": {"value": total * e2e_steps / (e2e_ms * 1e-3), "unit": "candidates/s",
                    "h2d_bytes_per_step": int(n_local * N_DIM * 4), "d2h_bytes_per_step": 8,
                    "steps": e2e_steps},
            "e2e_generated": {"value": total * args.steps / (gen_ms * 1e-3), "unit": "candidates/s", "steps": args.steps,
                              "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
                              "note": "candidate descriptor (counter-based uniform stream, seed, start) instead of candidate rows: "
                                      "x is generated inside the scorer, bit-identical to the resident rows"},
            "winner_check": check,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "bo_step": c2,
            "side": side,
        }
        print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


def winner_check(arm, orc, winners, n_local):
    """Each winner (score, global index) re-scored by the CPU arm: the score within 1e-4 relative, and no row of a CPU-scored
    sample of the same shard beats it by more than that.  Outside every timed region."""
    out = {"checker": arm.kind, "tolerance": 1e-4, "ok": True, "winners": []}
    sample = orc.counter_uniform(0, 0, 20000, N_DIM).astype(np.float64)
    smax = float(np.max(arm.scores(sample)))
    for name, (score, gidx) in winners:
        ref = float(arm.scores(orc.counter_uniform(0, gidx, 1, N_DIM).astype(np.float64))[0])
        rel = abs(score - ref) / abs(ref)
        ok = rel <= 1e-4 and ref >= smax - 1e-4 * abs(smax) and 0 <= gidx < n_local
        out["winners"].append({"run": name, "index": int(gidx), "score": float(score), "cpu_score": ref, "rel_err": rel,
                               "cpu_sample_max": smax, "ok": bool(ok)})
        out["ok"] = bool(out["ok"] and ok)
    return out


def side_configs():
    """Driver-visible lines for the other BASELINE configs (C4 batched runs, C5 trajectories) and the multi-agent encoder."""
    out = {}
    with contextlib.redirect_stdout(sys.stderr):
        from tools import configs_bench as cb
        for name, fn, kw in (("c5", cb.c5, {}), ("multi_agent", cb.multi, {}), ("c4", cb.c4, {"runs": 4096, "steps": 3})):
            try:
                out[name] = fn(**kw)
            except Exception as e:                               # a side line must not take the headline down
                out[name] = {"error": f"{type(e).__name__}: {e}"}
    return out


def rank_sweep(pkg, sp_unused, cand, orc, X, y, world, td, n=1 << 22, reps=3,
               ranks=(10, 60, 127, 128, 255, 256, 512, 1000)):
    """Kernel-only candidates/s per GPU (CUDA events, device-resident candidates, one launch of n candidates, float64 pass over
    the flagged candidates included) of the engine ENGINE_AUTO picks at fixed posterior ranks, up to the full-rank posterior
    after 1000 observations.  With several ranks of a job every rank runs the sweep at the same time and the slowest counts."""
    import torch
    from ssp_bayesopt_b200 import _lib
    local = cand.device.index
    b6 = np.array([[0.0, 1.0]] * N_DIM)
    sp = pkg.RandomSSPSpace(N_DIM, ssp_dim=D_REQ, domain_bounds=b6, length_scale=LS, rng=0, device=local)
    rng = np.random.default_rng(7)
    Xs = np.concatenate([X, rng.uniform(0, 1, (max(ranks) - X.shape[0], N_DIM))])
    ys = np.concatenate([y, orc.hartmann6(Xs[X.shape[0]:]).reshape(-1, 1)])
    agt = pkg.SSPAgent(Xs[:10], ys[:10], sp, decoder_method="from-set", gamma_c=0.0, beta_ucb=LAM, length_scale=LS)
    eng = agt._scoring_engine()
    x = cand[:n]
    d = sp.ssp_dim
    peaks = load_peaks()
    out, done = [], 10
    for r in ranks:
        if r > done:
            agt.update(Xs[done:r], ys[done:r], 0.0)
            done = r
        eng.set_policy(lowrank=True)
        eng.score_stats(reset=True)
        eng.score(x, LAM, GAMMA)                                  # warm-up (operand images, tensor maps)
        eng.best()
        if td is not None:
            td.barrier()
        torch.cuda.synchronize()
        eng.set_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.score(x, LAM, GAMMA)
        e1.record()
        torch.cuda.synchronize()
        kt = eng.timing_read(reset=True)
        eng.set_timing(False)
        st = eng.score_stats(reset=True)
        rate = x.shape[0] * reps / (e0.elapsed_time(e1) * 1e-3)
        if td is not None:
            t = torch.tensor([rate], dtype=torch.float64, device=cand.device)
            td.all_reduce(t, op=td.ReduceOp.MIN)
            rate = float(t.item())
        name = eng.last_engine()
        lr_rank = eng.lowrank_info()["rank"] if eng.lowrank_info()["valid"] else None
        fl, what = executed_flops_per_candidate(name, r if lr_rank is None else lr_rank, d)
        tf = x.shape[0] * reps * fl / (kt["scorer_ms"] * 1e-3) / 1e12 if (fl and kt["scorer_ms"] > 0) else None
        out.append({"rank": r, "engine": name, "cand_per_s_per_gpu": rate, "cand_per_s_job": rate * world,
                    "flagged_fraction": st["flagged"] / max(st["scored"], 1),
                    "scorer_ms_per_launch": kt["scorer_ms"] / max(kt["calls"], 1),
                    "flagged_pass_ms_per_launch": kt["flagged_ms"] / max(kt["calls"], 1),
                    "executed_tflops": tf, "frac_of_tensor_peak": (tf / peaks["tflops_burst"]) if tf else None,
                    "tensor_peak": "burst bf16 (kernel timed alone)", "executed_what": what})
    return out


def encode_only(eng, cand, peaks, n=1 << 16, reps=3):
    """The unfused path: API `encode` (K1) writes phi (N x d float64) to HBM.  Achieved write bandwidth against the
    measured HBM peak; the kernel is float64 throughout (the API promises 1e-13), so it is FP64-pipe-bound, not HBM-bound."""
    import torch
    x = cand[:n]
    eng.encode_device(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.encode_device(x)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    gbs = n * eng.d * 8 / sec / 1e9
    out = {"f64": {"kernel": "k_encode (float64 phi to HBM)", "candidates": n, "candidates_per_s": n / sec,
                   "hbm_write_gbs": gbs, "hbm_peak_gbs": peaks["hbm"], "frac_of_hbm_peak": gbs / peaks["hbm"],
                   "note": "d K float64 FMA per candidate (output columns j and d - j share their products): bound by the FP64 pipe; the fused scorer never materialises phi"}}
    n32 = min(cand.shape[0], 1 << 20)
    x32 = cand[:n32]
    eng.encode_f32_device(x32)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        eng.encode_f32_device(x32)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    pitch = eng.lib.sspbo_encode_pitch(eng._ctx)
    gbs = n32 * pitch * 4 / sec / 1e9
    tf = n32 * 3 * 2.0 * 1024 * pitch / sec / 1e12
    out["f32_tensor_core"] = {"kernel": "k_score_umma in encode mode (split-fp16 inverse-DFT basis, float32 phi to HBM)",
                              "candidates": n32, "candidates_per_s": n32 / sec, "hbm_write_gbs": gbs,
                              "hbm_peak_gbs": peaks["hbm"], "frac_of_hbm_peak": gbs / peaks["hbm"],
                              "executed_tflops": tf, "frac_of_tensor_peak": tf / peaks["tflops"],
                              "note": "tensor-pipe-bound: 3 x 2 x 1024 x 1024 executed flops per candidate"}
    return out


def update_cost(agt, sp, orc, reps=20):
    """One posterior observation: device float64 rank-one update (x_t encoded inside K3's single launch) against the
    reference's expression restated in numpy (blr.py:63-75, O(d^3) as written there), same d."""
    import torch
    rng = np.random.default_rng(11)
    xs = rng.uniform(0, 1, (reps, N_DIM))
    ys = orc.hartmann6(xs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        agt.update(xs[i:i + 1], np.atleast_2d(ys[i]), np.array([0.0]))
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    # the low-rank updates leave S, S_inv, b, m for later: what reading them afterwards costs (all `reps` observations folded in)
    t0 = time.perf_counter()
    agt._scoring_engine().blr_get(m=True, S=True, Sinv=True)
    torch.cuda.synchronize()
    fold_ms = (time.perf_counter() - t0) * 1e3
    ref = orc.BLR(sp.ssp_dim, beta=1.0)
    phis = orc.encode(sp.phase_matrix, LS * np.ones(N_DIM), xs[:3])
    ref.update(phis[:1], np.atleast_2d(ys[0]))
    t0 = time.perf_counter()
    for i in range(1, 3):
        ref.update(phis[i:i + 1], np.atleast_2d(ys[i]))
    cpu_ms = (time.perf_counter() - t0) / 2 * 1e3
    return {"gpu_ms_per_observation": gpu_ms, "cpu_port_ms_per_observation": cpu_ms, "d": int(sp.ssp_dim),
            "dense_state_on_demand_ms": fold_ms, "dense_state_observations_folded": reps,
            "note": "wall clock incl. host calls; the device update is float64 in low-rank form, O(rank d) from the rows of V "
                    "(k_blr_update_lr; S, S_inv, psi-space covariance follow on demand), the reference expression O(d^3)"}


def acq_ascent_cost(agt, sp, orc, restarts=8, steps=100, reps=5):
    """K6: phi-space acquisition search at the C3 posterior.  Device: `restarts` climbs of `steps`
    fixed-point steps in one cooperative launch (CUDA events).  CPU: the reference's route restated in the oracle --
    scipy L-BFGS-B on `acquisition_func` (ssp_agent.py:156-185, bayesian_optimization.py:270-287) per restart, from the
    same starting points (inside the ball, so L-BFGS-B does iterate)."""
    import torch
    eng = agt._scoring_engine()
    d = int(sp.ssp_dim)
    rng = np.random.default_rng(12)
    phi0 = orc.encode(sp.phase_matrix, LS * np.ones(N_DIM), rng.uniform(0, 1, (restarts, N_DIM)))
    lam, gamma = float(agt.var_weight), agt._gamma()
    p0 = eng.to_device(phi0, torch.float64)
    eng.acq_ascent(p0, lam, gamma, max_iter=steps, tol=0.0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        _, val, _ = eng.acq_ascent(p0, lam, gamma, max_iter=steps, tol=0.0)
    ev[1].record()
    torch.cuda.synchronize()
    gpu_ms = ev[0].elapsed_time(ev[1]) / reps
    _, val_c, its_c = eng.acq_ascent(p0, lam, gamma, max_iter=2000, tol=1e-7)
    m, S, beta = agt.blr.m, agt.blr.S, float(np.ravel(agt.blr.beta)[0])
    f, g = orc.acquisition_func(m, S, beta, lam, gamma)
    t0 = time.perf_counter()
    _, vals_ref, nit = orc.maximize_acquisition_lbfgs(f, g, phi0[:2])
    cpu_ms = (time.perf_counter() - t0) / 2 * 1e3
    return {"d": d, "restarts": restarts, "steps": steps, "gpu_ms_per_call": gpu_ms, "gpu_us_per_step": gpu_ms * 1e3 / steps,
            "S_bytes_per_step": 8 * d * d, "S_GBps": 8 * d * d * steps / (gpu_ms * 1e-3) / 1e9,
            "f64_gflops": 2.0 * d * d * restarts * steps / (gpu_ms * 1e-3) / 1e9,
            "value_after_steps": float(val.max().item()), "value_converged": float(val_c.max().item()),
            "steps_to_converge_tol1e-7": [int(v) for v in its_c.cpu()],
            "cpu_port_ms_per_restart": cpu_ms, "cpu_port_value": float(np.max(vals_ref)), "cpu_port_lbfgs_nit": [int(v) for v in nit],
            "note": "latency-bound: one grid barrier per step; S (float64, d x d) is re-read from L2 every step by all CTAs "
                    "together; CPU = scipy L-BFGS-B on the oracle's acquisition_func per restart"}


def bo_step_c2(pkg, n_iter=250):
    """'ms per BO step' of BASELINE config C2 through the public driver: Branin, ssp-hex ssp_dim=151, 1000 x 1000
    candidate grid, `BayesianOptimization.maximize` (fused scoring of the 1e6 grid + winner read-back + target
    evaluation + eval(x_t) + float64 posterior update per step)."""
    bounds = np.array([[-5.0, 10.0], [0.0, 15.0]])
    a, b, c, r, s, t = 1, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6., 10., 1 / (8 * np.pi)
    branin = lambda x: -(a * (x[:, 1] - b * x[:, 0] ** 2 + c * x[:, 0] - r) ** 2 + s * (1 - t) * np.cos(x[:, 0]) + s)
    xs0 = np.random.default_rng(0).uniform(bounds[:, 0], bounds[:, 1], (10, 2))
    opt = pkg.BayesianOptimization(f=branin, bounds=bounds)
    opt.maximize(init_points=(xs0, branin(xs0).reshape(-1, 1)), n_iter=n_iter, agent_type="ssp-hex", ssp_dim=151,
                 length_scale=1.0, beta_ucb=10.0, gamma_c=0.0, candidates_per_dim=1000)
    ft = opt.full_times[5:] * 1e-6                                  # ns -> ms, skip the first (warm-up) steps
    out = {"config": f"C2: Branin 2-D ssp-hex d=151, 1e6-point candidate grid, BayesianOptimization.maximize, {n_iter} iterations",
           "ms_per_bo_step": float(np.mean(ft)), "ms_median": float(np.median(ft)), "ms_min": float(np.min(ft)), "steps": int(len(ft)),
           "candidates_per_s": float(1e6 / (np.mean(opt.times[5:]) * 1e-9)), "best_target": float(opt.max["target"]),
           # the posterior gains one observation per step (rank 10 + step): the step time by stretch of the run
           # (medians: the first launch of a kernel the run has not used yet -- the fp16 + e4m3 scorer from rank 128 on --
           # loads its module lazily, some 20 ms once, which the mean above includes)
           "ms_by_iteration": {f"{a}-{b}": float(np.median(opt.full_times[a:b]) * 1e-6)
                               for a, b in ((5, 50), (50, 118), (118, 200), (200, n_iter)) if b <= n_iter and b > a},
           "ms_max": float(np.max(ft)),
           "engine_at_end": opt.agt._scoring_engine().last_engine()}
    # the same step through the reference's arithmetic (numpy oracle port) on the host: eval + argmax over a bounded sample of
    # the grid (extrapolated to 1e6 points) plus the reference's posterior update expression
    from oracle import ssp_oracle as orc
    A = opt.agt.ssp_space.phase_matrix
    ref = orc.BLR(A.shape[0])
    ref.update(orc.encode(A, np.ones(2), xs0), branin(xs0).reshape(-1, 1))
    sample = np.random.default_rng(1).uniform(bounds[:, 0], bounds[:, 1], (20000, 2))
    t0 = time.perf_counter()
    score, idx = orc.score_and_argmax(orc.encode(A, np.ones(2), sample), ref, 10.0, 0.0)
    t_score = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref.update(orc.encode(A, np.ones(2), sample[idx:idx + 1]), np.atleast_2d(branin(sample[idx:idx + 1])))
    t_upd = time.perf_counter() - t0
    out["cpu_port_ms_per_bo_step"] = float((t_score * (1e6 / len(sample)) + t_upd) * 1e3)
    out["cpu_port_note"] = ("oracle/ssp_oracle.py on the host cores: eval + argmax timed on 20000 of the 1e6 grid points and "
                            "scaled by 50, plus one posterior update")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "umma"])
    ap.add_argument("--per-gpu", type=int, default=PER_GPU)
    ap.add_argument("--chunk", type=int, default=1 << 22)
    ap.add_argument("--no-c2", action="store_true", help="skip the config-C2 'ms per BO step' side measurement")
    ap.add_argument("--no-sweep", action="store_true", help="skip the kernel-only rank sweep")
    ap.add_argument("--no-side", action="store_true", help="skip the side lines of configs C4 / C5 / multi-agent")
    ap.add_argument("--cpu-sample", type=int, default=100000, help="candidates per step of the CPU arm (BASELINE.md 3: >= 1e5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
