#!/usr/bin/env python
"""Benchmark of the SSP-BO hot path: fused encode + BLR-UCB + argmax candidates/s at d=1024 (-> 1023).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[2] ("C3"): RandomSSPSpace(n=6, ssp_dim=1024 -> d=1023, K=511), bounds [0,1]^6,
length_scale 0.25, rng=0; posterior after 10 + 50 Hartmann-6 observations; 1e8 counter-based uniform float32
candidates per BO step, sharded over 8 GPUs = 1.25e7 candidates per GPU per step (weak scaling: the per-GPU shard
is fixed, so N GPUs score N * 1.25e7 candidates per step).  One step = score every local candidate with the
fused kernel, max-reduce the packed (score, index) key over ranks, read the winner, evaluate Hartmann-6 on
it and apply the float64 rank-one posterior update (so every step sees a new posterior).

`value`   candidates/s, whole job, candidates resident in HBM, device-timed (CUDA events, max over ranks).
`e2e`     same metric through the public API (`SSPAgent.best_candidate`) from pinned HOST candidates: the
          H2D copy of the shard and the D2H read of the winner are inside the timed region every step.
`roofline` tensor-pipe: achieved = candidates * F_alg / kernel time, F_alg = 2 d^2 + 2 d + 2 K n flops
          (SURVEY.md section 8d), peak = MEASURED_PEAKS.json bf16 dense (sustained, the kernel runs inside a long step).
          While the posterior has rank r <= 512 the scorer uses the low-rank form (2 d r flops instead of 2 d^2), so
          the algorithmic figure can exceed the dense-form peak; `executed` gives the tensor flops really issued, and
          `by_rank` the kernel-only rate at other ranks and for the dense (full-factor) engine.
`cpu_baseline` / `--impl reference`: the numpy oracle port of the reference path on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_REQ, N_DIM, LS, LAM, GAMMA = 1024, 6, 0.25, 10.0, 0.0
PER_GPU = 12_500_000
N_OBS = 60
METRIC = "fused encode+UCB+argmax candidates/sec at d=1024"


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
def cpu_oracle_rate(sample, chunk, steps=1, warmup=0):
    """Reference path restated in numpy (oracle/ssp_oracle.py): encode -> predict -> acq -> argmax, chunked so the
    (d x chunk) complex128 temporary stays small.  Returns (candidates/s, seconds per step, cores)."""
    X, y, orc = problem_setup()
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
    times = []
    for it in range(warmup + steps):
        xc = orc.counter_uniform(0, it * sample, sample, N_DIM).astype(np.float64)
        t0 = time.perf_counter()
        best, best_i = -np.inf, -1
        for c0 in range(0, sample, chunk):
            score, i = orc.score_and_argmax(orc.encode(A, ls, xc[c0:c0 + chunk]), blr, LAM, GAMMA)
            if score[i] > best:
                best, best_i = score[i], c0 + i
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    per_step = float(np.mean(times))
    return sample / per_step, per_step, os.cpu_count()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample, chunk = 20000, 10000
    rate, per_step, cores = cpu_oracle_rate(sample, chunk, steps=args.steps, warmup=args.warmup)
    d = 1023
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "candidates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3: RandomSSPSpace n=6 ssp_dim=1024 (d=1023) Hartmann-6 posterior, uniform candidates",
                   "candidates_per_step": sample, "note": "bounded sample of the 1.25e7-per-GPU shard"},
        "cpu_baseline": {"value": rate, "unit": "candidates/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} candidates/step in chunks of {chunk}, numpy/OpenBLAS on all host cores "
                                   "(the reference is pure Python and cannot travel to the GPU box; this is its numpy "
                                   "restatement, oracle/ssp_oracle.py, pinned to the live reference by tests/golden)"},
        "e2e": {"value": rate, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "roofline": None,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
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
    agt = pkg.SSPAgent(X, y, sp, decoder_method="from-set", gamma_c=0.0, beta_ucb=LAM, length_scale=LS,
                       score_engine={"auto": _lib.ENGINE_AUTO, "simt": _lib.ENGINE_SIMT, "umma": _lib.ENGINE_UMMA}[args.engine])
    eng = agt._scoring_engine()
    d, K = sp.ssp_dim, (sp.ssp_dim - 1) // 2
    n_local = args.per_gpu
    total = n_local * world
    start = rank * n_local
    cand = eng.fill_uniform(0, start, n_local)                    # resident shard, 300 MB > L2 (126 MB)
    chunk = args.chunk
    torch.cuda.synchronize()

    def select(x_dev):
        first = True
        for c0 in range(0, n_local, chunk):
            agt.best_candidate(x_dev[c0:c0 + chunk], index0=start + c0, reset_best=first)
            first = False
        eng.settle()                                              # low-rank pass counters (synchronises; replays if asked)
        dist.allreduce_best(eng._best)
        return eng.best()                                         # D2H of the 8-byte key (synchronises)

    def winner_row(gidx):
        u = orc.counter_uniform(0, gidx, 1, N_DIM)                # any rank can regenerate any candidate
        return u.astype(np.float64)

    def bo_step(x_dev, kernel_events=None):
        if kernel_events is not None:
            kernel_events[0].record()
        score, gidx = select(x_dev)
        if kernel_events is not None:
            kernel_events[1].record()
        x_t = winner_row(gidx)
        y_t = np.atleast_2d(orc.hartmann6(x_t))
        var_t = np.array([0.0])
        agt.update(x_t, y_t, var_t)                               # encode x_t + float64 rank-one update on device
        return score, gidx

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    for _ in range(args.warmup):
        bo_step(cand)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    last = None
    for i in range(args.steps):
        last = bo_step(cand, ev[i])
    t_end.record()
    barrier()
    sampler.stop_flag = True
    launches = eng.launch_count() - launches0
    ms_total = t_start.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))   # scoring launches + key reduction, per step
    engine_used = eng.last_engine()

    # ---- end-to-end timing: pinned host candidates -> H2D -> fused score -> D2H key, through the public API ----
    host = torch.empty((n_local, N_DIM), dtype=torch.float32).pin_memory()
    host.copy_(cand)
    e2e_steps = max(1, min(args.steps, 3))
    def e2e_step():
        # public API on HOST candidates: SSPAgent.best_candidate streams the pinned shard to the GPU in chunks
        # (H2D inside the timed region, overlapped with the scoring), then the 8-byte key is read back
        agt.best_candidate(host, index0=start)
        eng.settle()
        dist.allreduce_best(eng._best)
        score, gidx = eng.best()
        x_t = winner_row(gidx)
        agt.update(x_t, np.atleast_2d(orc.hartmann6(x_t)), np.array([0.0]))
        return score, gidx
    e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    # max over ranks
    stats = torch.tensor([ms_total, kernel_ms, e2e_ms], dtype=torch.float64, device=cand.device)
    if world > 1:
        td.all_reduce(stats, op=td.ReduceOp.MAX)
    ms_total, kernel_ms, e2e_ms = [float(v) for v in stats.cpu()]
    sampler.join(timeout=2)

    info = eng.lowrank_info()
    sweep = rank_sweep(pkg, eng, agt, cand, orc) if (rank == 0 and world == 1 and not args.no_sweep) else None
    enc = encode_only(eng, cand, load_peaks()) if (rank == 0 and world == 1 and not args.no_sweep) else None
    upd = update_cost(agt, sp, orc) if (rank == 0 and world == 1 and not args.no_sweep) else None
    asc = acq_ascent_cost(agt, sp, orc) if (rank == 0 and world == 1 and not args.no_sweep) else None
    c2 = bo_step_c2(pkg) if (rank == 0 and world == 1 and not args.no_c2) else None
    if rank == 0:
        peaks = load_peaks()
        ms_per_step = ms_total / args.steps
        value = total / (ms_per_step * 1e-3)
        flops = f_alg(d, K, N_DIM)
        achieved = n_local * flops / (kernel_ms * 1e-3) / 1e12      # per GPU, TFLOP/s algorithmic
        cpu_rate, cpu_step, cores = cpu_oracle_rate(20000, 10000, steps=1, warmup=0) if world == 1 else (None, None, None)
        line = {
            "metric": METRIC, "value": value, "unit": "candidates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": ("f16x3-split mma / f32 accumulate / fixed-point phase / f64 posterior" if engine_used.startswith("umma")
                      else "f32 contraction / f64 phase+posterior"),
            "data": "synthetic",
            "config": {"workload": "C3: RandomSSPSpace n=6 ssp_dim=1024 (d=1023,K=511) ls=0.25 rng=0, Hartmann-6 posterior "
                                   f"after {N_OBS}+step observations, counter-based uniform f32 candidates",
                       "candidates_per_gpu_per_step": n_local, "candidates_per_step": total, "chunk": chunk,
                       "engine": engine_used, "posterior_rank": info["rank"], "phase32": info["phase32"],
                       "parallelism": f"candidate-shard x{world}",
                       "l2": "inputs (300 MB candidate shard) larger than L2; posterior operands refreshed every step",
                       "last_winner": {"score": last[0], "index": last[1]}},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"],
                         # dram__bytes_read+write of one scorer launch (chunk of 4,194,304 candidates) from the ncu --set full
                         # captures summarised in profiles/ (algorithmic bytes: 4*n per candidate = 100.7 MB)
                         "traffic": (106.1e6 if engine_used == "umma-lowrank" else 108.6e6) if engine_used.startswith("umma") else None,
                         "launch_candidates": min(chunk, n_local),
                         "executed": executed_note(engine_used, info["rank"], d, n_local, kernel_ms, peaks),
                         "note": f"algorithmic (dense form, SURVEY 8d) {flops / 1e6:.3f} MFLOP/candidate x {n_local} candidates / "
                                 f"{kernel_ms:.2f} ms scoring time per step per GPU (all launches of the step, incl. the key "
                                 "read-back); frac > 1 means the low-rank form beats the dense-form roofline by doing less "
                                 f"work, not that the tensor pipe runs above peak; peak {peaks['source']}"},
            "by_rank": sweep,
            "encode_only": enc,
            "posterior_update": upd,
            "acq_ascent": asc,
            "cpu_baseline": ({"value": cpu_rate, "unit": "candidates/s", "cores": cores, "kind": "port",
                              "sample": "20000 candidates in chunks of 10000, oracle/ssp_oracle.py (numpy float64, OpenBLAS)"}
                             if cpu_rate else None),
            "e2e": {"value": total * e2e_steps / (e2e_ms * 1e-3), "unit": "candidates/s",
                    "h2d_bytes_per_step": int(n_local * N_DIM * 4), "d2h_bytes_per_step": 8,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "bo_step": c2,
        }
        print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


def executed_note(engine, rank, d, n_local, kernel_ms, peaks):
    """Tensor flops the kernel really issues per candidate (3 split-fp16 MMAs per product)."""
    kc = 1024
    if engine == "umma-lowrank":
        r_eff = rank + 1                                   # row 0 of the operand is m'
        nc = 1 if r_eff <= 256 else 2
        rows = r_eff if nc == 1 else rank
        bn = -(-(-(-rows // nc)) // 16) * 16
        per_cand = 3 * 2.0 * kc * bn * nc
        what = f"low-rank: 3 x 2 x {kc} x {nc}x{bn} per candidate"
    elif engine == "umma":
        per_cand = 3 * 2.0 * kc * 1024 * 0.56
        what = "dense triangular factor: 3 x 0.56 x 2 x 1024^2 per candidate"
    else:
        return None
    tf = n_local * per_cand / (kernel_ms * 1e-3) / 1e12
    return {"tflops": tf, "frac_of_peak": tf / peaks["tflops"], "what": what}


def rank_sweep(pkg, eng, agt, cand, orc, n=1 << 22, reps=3):
    """Kernel-only candidates/s (CUDA events, device-resident candidates, one launch of n candidates) of the low-rank
    engine at the current and at higher posterior ranks, and of the dense full-factor engine."""
    import torch
    from ssp_bayesopt_b200 import _lib
    x = cand[:n]
    out = []

    def timed(code):
        eng.score(x, LAM, GAMMA, engine=code)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.score(x, LAM, GAMMA, engine=code)
        e1.record()
        torch.cuda.synchronize()
        return x.shape[0] * reps / (e0.elapsed_time(e1) * 1e-3)

    rng = np.random.default_rng(7)
    for target in (None, 128, 250):
        while target is not None and eng.lowrank_info()["rank"] < target:
            x_t = rng.uniform(0, 1, (1, N_DIM))
            agt.update(x_t, np.atleast_2d(orc.hartmann6(x_t)), np.array([0.0]))
        r = eng.lowrank_info()["rank"]
        row = {"rank": r, "lowrank_cand_per_s": timed(_lib.ENGINE_UMMA_LR)}
        if target in (None, 250):
            row["dense_cand_per_s"] = timed(_lib.ENGINE_UMMA)
        eng.score_stats(reset=True)
        out.append(row)
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
                   "note": "2 d K float64 FMA per candidate: bound by the FP64 pipe; the fused scorer never materialises phi"}}
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
    """One posterior observation: device float64 rank-one update (encode x_t + K3, about 15 launches) against the
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
    ref = orc.BLR(sp.ssp_dim, beta=1.0)
    phis = orc.encode(sp.phase_matrix, LS * np.ones(N_DIM), xs[:3])
    ref.update(phis[:1], np.atleast_2d(ys[0]))
    t0 = time.perf_counter()
    for i in range(1, 3):
        ref.update(phis[i:i + 1], np.atleast_2d(ys[i]))
    cpu_ms = (time.perf_counter() - t0) / 2 * 1e3
    return {"gpu_ms_per_observation": gpu_ms, "cpu_port_ms_per_observation": cpu_ms, "d": int(sp.ssp_dim),
            "note": "wall clock incl. host calls; the device update is O(d^2) float64 (S, S_inv, psi-space covariance, one "
                    "operand row), the reference expression O(d^3)"}


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


def bo_step_c2(pkg, n_iter=30):
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
    out = {"config": "C2: Branin 2-D ssp-hex d=151, 1e6-point candidate grid, BayesianOptimization.maximize",
           "ms_per_bo_step": float(np.mean(ft)), "ms_min": float(np.min(ft)), "steps": int(len(ft)),
           "candidates_per_s": float(1e6 / (np.mean(opt.times[5:]) * 1e-9)), "best_target": float(opt.max["target"])}
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
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "umma"])
    ap.add_argument("--per-gpu", type=int, default=PER_GPU)
    ap.add_argument("--chunk", type=int, default=1 << 22)
    ap.add_argument("--no-c2", action="store_true", help="skip the config-C2 'ms per BO step' side measurement")
    ap.add_argument("--no-sweep", action="store_true", help="skip the kernel-only rank sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
