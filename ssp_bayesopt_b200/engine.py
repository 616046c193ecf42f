"""One device context of the SSP-BO hot path (a `sspbo_ctx` of libsspbo.so).

PyTorch is used for plumbing only: device buffers, the current CUDA stream and
(in `dist.py`) the NCCL process group.  All arithmetic is in the hand-written
kernels behind the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


# engines whose fast phase path may ask for a recomputation (candidates outside the declared |x| bound, flag-list overflow)
_RERUN_ENGINES = (_lib.ENGINE_UMMA_LR, _lib.ENGINE_UMMA_F8)


def _torch():
    import torch
    return torch


def require_cuda(device=None):
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.SspboError("ssp_bayesopt_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    return int(torch.device("cuda", device).index if not isinstance(device, int) else device)


def stream_pieces(N, chunk, first_piece=None, growth=1.5):
    """Piece schedule [(start, count), ...] of a host candidate set streamed through the staging buffers: a short first piece
    (its copy is the only one no scoring overlaps), pieces growing by `growth` up to `chunk`, and the mirror image at the end
    (the scoring of the last piece is the only one no copy overlaps).  A remainder below 4096 rows joins a neighbour."""
    front, back, remaining, n = [], [], int(N), int(first_piece or max(chunk >> 2, 1 << 15))
    while remaining > 0:
        n = min(n, chunk)
        take = min(n, remaining)
        front.append(take); remaining -= take
        if remaining > 0:
            take = min(n, remaining)
            back.append(take); remaining -= take
        n = int(n * growth)
    sizes = front + back[::-1]
    i = len(front) - 1 if len(front) > len(back) else len(front)   # the piece taken last: where the two ramps meet
    if len(sizes) > 1 and sizes[i] < 4096:                       # a small remainder joins a neighbour
        for j in (i - 1, i + 1):
            if 0 <= j < len(sizes) and sizes[j] + sizes[i] <= chunk:
                sizes[j] += sizes[i]; sizes[i] = 0
                break
    pieces, c0 = [], 0
    for take in sizes:
        if take:
            pieces.append((c0, take))
            c0 += take
    return pieces


class DeviceEngine:
    """Owns a device context: SSP space tables, the float64 BLR posterior and its scoring factor."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        self.device = require_cuda(device)
        self.torch = _torch()
        self._ctx = C.c_void_p()
        rc = self.lib.sspbo_ctx_create(self.device, C.byref(self._ctx))
        if rc != 0:
            raise _lib.SspboError(f"sspbo_ctx_create({self.device}) failed ({rc})")
        self.d = 0
        self.K = 0
        self.n = 0
        self.L = 1
        self._best = self.torch.zeros(1, dtype=self.torch.int64, device=self._dev())
        self._blr_owner = None       # weakref to the BayesianLinearRegression whose posterior lives in this context
        self._pending = []           # score calls since the last key reset (replayed if the device asks for a rerun)
        self._exchange_on = None     # peer-memory key exchange across ranks: None = not set up yet (dist_connect)
        self._lr_used = False        # ... and whether any of them ran the low-rank engine

    # ------------------------------------------------------------------ plumbing
    def _dev(self):
        return self.torch.device("cuda", self.device)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc, what):
        _lib.check(self.lib, self._ctx, rc, what)

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.sspbo_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def to_device(self, a, dtype=None):
        """numpy / torch -> contiguous device tensor (float32 or float64 kept, everything else -> float64)."""
        torch = self.torch
        if isinstance(a, torch.Tensor):
            t = a
            if dtype is not None:
                t = t.to(dtype)
            elif t.dtype not in (torch.float32, torch.float64):
                t = t.to(torch.float64)
            return t.to(self._dev()).contiguous()
        a = np.asarray(a)
        if dtype is None:
            a = a if a.dtype in (np.float32, np.float64) else a.astype(np.float64)
            return torch.from_numpy(np.ascontiguousarray(a)).to(self._dev())
        return torch.from_numpy(np.ascontiguousarray(a)).to(self._dev(), dtype)

    @staticmethod
    def _dtype_code(t):
        import torch
        return _lib.F32 if t.dtype == torch.float32 else _lib.F64

    # ------------------------------------------------------------------ space
    def set_space(self, phase_matrix: np.ndarray, length_scale) -> None:
        """phase_matrix (d, n) as built by the reference constructors; rows 1..K are A+."""
        A = np.asarray(phase_matrix, dtype=np.float64)
        d, n = A.shape
        if d % 2 != 1 or d < 3:
            raise ValueError(f"phase matrix must have 2K+1 rows (conjugate symmetric), got {d}")
        K = (d - 1) // 2
        a_plus = np.ascontiguousarray(A[1:K + 1])
        inv_ls = np.ascontiguousarray(1.0 / np.asarray(length_scale, dtype=np.float64).reshape(-1) * np.ones(n))
        rc = self.lib.sspbo_space_set(self._ctx, a_plus.ctypes.data_as(_lib._pd), inv_ls.ctypes.data_as(_lib._pd), K, n)
        self._check(rc, "sspbo_space_set")
        self.d, self.K, self.n, self.L = d, K, n, 1

    def set_space_classes(self, phase_matrices, length_scales) -> None:
        """One phase matrix per agent (all (d, n)): waypoint j of a row is encoded with matrix j % len(phase_matrices)
        (SSPMultiAgent with per-agent x spaces, ssp_multi_agent.py:58-66).  Followed by set_waypoints."""
        As = [np.asarray(A, dtype=np.float64) for A in phase_matrices]
        d, n = As[0].shape
        if d % 2 != 1 or d < 3 or any(A.shape != (d, n) for A in As):
            raise ValueError("phase matrices must share one (2K+1, n) shape")
        K = (d - 1) // 2
        a_plus = np.ascontiguousarray(np.stack([A[1:K + 1] for A in As]))
        inv_ls = np.ascontiguousarray(np.stack([1.0 / np.asarray(ls, dtype=np.float64).reshape(-1) * np.ones(n)
                                                for ls in length_scales]))
        assert inv_ls.shape == (len(As), n)
        rc = self.lib.sspbo_space_set_classes(self._ctx, a_plus.ctypes.data_as(_lib._pd), inv_ls.ctypes.data_as(_lib._pd),
                                              K, n, len(As))
        self._check(rc, "sspbo_space_set_classes")
        self.d, self.K, self.n, self.L = d, K, n, (len(As) if len(As) > 1 else 1)

    def set_xbound(self, x_absmax: float) -> None:
        """Declared bound on |x| of every candidate coordinate (max |domain_bounds|); 0 = unknown."""
        self._check(self.lib.sspbo_space_set_xbound(self._ctx, float(x_absmax)), "sspbo_space_set_xbound")

    def set_traj(self, tau: np.ndarray | None, L: int) -> None:
        """tau (L, K): phase in radians of timestep j at frequency k."""
        if tau is None:
            rc = self.lib.sspbo_space_set_traj(self._ctx, None, 1)
            L = 1
        else:
            tau = np.ascontiguousarray(tau, dtype=np.float64)
            assert tau.shape == (L, self.K), f"tau must be ({L}, {self.K}), got {tau.shape}"
            rc = self.lib.sspbo_space_set_traj(self._ctx, tau.ctypes.data_as(_lib._pd), L)
        self._check(rc, "sspbo_space_set_traj")
        self.L = L

    def set_waypoints(self, tau: np.ndarray, dc_sign=None, normalize=False) -> None:
        """Generalised trajectory mode (multi-agent trajectories): tau (L, K) radians per waypoint and frequency,
        dc_sign (L,) the +-1 DC coefficient of each waypoint's tag, normalize: consumers use phi / |phi|."""
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        L = tau.shape[0]
        assert tau.shape == (L, self.K), f"tau must be (L, {self.K}), got {tau.shape}"
        sp = None
        if dc_sign is not None:
            sign = np.ascontiguousarray(dc_sign, dtype=np.int32)
            assert sign.shape == (L,)
            sp = sign.ctypes.data_as(_lib._pi)
        rc = self.lib.sspbo_space_set_waypoints(self._ctx, tau.ctypes.data_as(_lib._pd), sp, L, 1 if normalize else 0)
        self._check(rc, "sspbo_space_set_waypoints")
        self.L = L

    # ------------------------------------------------------------------ K1
    def encode_device(self, x):
        """x (N, n*L) -> device tensor phi (N, d) float64."""
        xt = self.to_device(x)
        if xt.dim() == 1:
            xt = xt.reshape(1, -1)
        if xt.shape[1] != self.n * self.L:
            raise AssertionError(f"Expected {self.n * self.L}-d data, got {tuple(xt.shape)}")
        N = xt.shape[0]
        phi = self.torch.empty((N, self.d), dtype=self.torch.float64, device=self._dev())
        rc = self.lib.sspbo_encode(self._ctx, C.c_void_p(xt.data_ptr()), self._dtype_code(xt), N,
                                   C.c_void_p(phi.data_ptr()), self._stream())
        self._check(rc, "sspbo_encode")
        return phi

    def encode(self, x) -> np.ndarray:
        return self.encode_device(x).cpu().numpy()

    def encode_and_deriv_device(self, x):
        """x (N, n) -> (phi (N, d), grad (N, d, n)) float64 device tensors: grad[s, j, a] = d phi_j(x_s) / d x_a."""
        xt = self.to_device(x)
        if xt.dim() == 1:
            xt = xt.reshape(1, -1)
        if xt.shape[1] != self.n * self.L:
            raise AssertionError(f"Expected {self.n * self.L}-d data, got {tuple(xt.shape)}")
        N = xt.shape[0]
        phi = self.torch.empty((N, self.d), dtype=self.torch.float64, device=self._dev())
        grad = self.torch.empty((self.n, N, self.d), dtype=self.torch.float64, device=self._dev())
        rc = self.lib.sspbo_encode_deriv(self._ctx, C.c_void_p(xt.data_ptr()), self._dtype_code(xt), N,
                                         C.c_void_p(phi.data_ptr()), C.c_void_p(grad.data_ptr()), self._stream())
        self._check(rc, "sspbo_encode_deriv")
        return phi, grad.permute(1, 2, 0)

    def rff_encode_device(self, W, B, scale, x):
        """scale * cos(x W^T + B): W (m, n), B (m,) device float64 tensors, x (N, n) -> (N, m) float64 device tensor."""
        xt = self.to_device(x)
        if xt.dim() == 1:
            xt = xt.reshape(1, -1)
        m, n = W.shape
        if xt.shape[1] != n:
            raise AssertionError(f"Expected {n}-d data, got {tuple(xt.shape)}")
        N = xt.shape[0]
        phi = self.torch.empty((N, m), dtype=self.torch.float64, device=self._dev())
        rc = self.lib.sspbo_rff_encode(self._ctx, C.c_void_p(W.data_ptr()), C.c_void_p(B.data_ptr()), int(m), int(n), float(scale),
                                       C.c_void_p(xt.data_ptr()), self._dtype_code(xt), N, C.c_void_p(phi.data_ptr()), self._stream())
        self._check(rc, "sspbo_rff_encode")
        return phi

    def encode_f32_device(self, x):
        """x (N, n*L) -> device tensor phi (N, d) float32 (a view of an (N, pitch) buffer) from the tensor-core encoder."""
        xt = self.to_device(x)
        if xt.dim() == 1:
            xt = xt.reshape(1, -1)
        if xt.shape[1] != self.n * self.L:
            raise AssertionError(f"Expected {self.n * self.L}-d data, got {tuple(xt.shape)}")
        N = xt.shape[0]
        pitch = self.lib.sspbo_encode_pitch(self._ctx)
        buf = self.torch.empty((N, pitch), dtype=self.torch.float32, device=self._dev())
        rc = self.lib.sspbo_encode_f32(self._ctx, C.c_void_p(xt.data_ptr()), self._dtype_code(xt), N,
                                       C.c_void_p(buf.data_ptr()), self._stream())
        self._check(rc, "sspbo_encode_f32")
        return buf[:, :self.d]

    # ------------------------------------------------------------------ K3
    def blr_init(self, alpha: float, dim: int | None = None) -> None:
        if self.K:
            rc = self.lib.sspbo_blr_init(self._ctx, float(alpha))
        else:
            rc = self.lib.sspbo_blr_init_dim(self._ctx, int(dim), float(alpha))
            self.d = int(dim)
        self._check(rc, "sspbo_blr_init")

    def blr_set_beta(self, beta: float) -> None:
        self._check(self.lib.sspbo_blr_set_beta(self._ctx, float(beta)), "sspbo_blr_set_beta")

    def blr_beta(self):
        b, s = C.c_double(), C.c_int()
        self._check(self.lib.sspbo_blr_get_beta(self._ctx, C.byref(b), C.byref(s)), "sspbo_blr_get_beta")
        return b.value if s.value else None

    def blr_update(self, phis, ts) -> None:
        pt = self.to_device(phis, self.torch.float64)
        ts = np.ascontiguousarray(np.asarray(ts, dtype=np.float64).reshape(-1))
        assert pt.shape[0] == ts.shape[0]
        rc = self.lib.sspbo_blr_update(self._ctx, C.c_void_p(pt.data_ptr()), ts.ctypes.data_as(_lib._pd), pt.shape[0],
                                       self._stream())
        self._check(rc, "sspbo_blr_update")

    def blr_update_points(self, xs, ts) -> None:
        """Encode host points on the device and apply them as observations, one C call, no synchronisation."""
        xs = np.ascontiguousarray(np.atleast_2d(xs), dtype=np.float64)
        ts = np.ascontiguousarray(np.asarray(ts, dtype=np.float64).reshape(-1))
        assert xs.shape[0] == ts.shape[0]
        if xs.shape[1] != self.n * self.L:
            raise AssertionError(f"Expected {self.n * self.L}-d data, got {tuple(xs.shape)}")
        for i in range(0, xs.shape[0], 64):
            rc = self.lib.sspbo_blr_update_x(self._ctx, xs[i:i + 64].ctypes.data_as(_lib._pd), ts[i:i + 64].ctypes.data_as(_lib._pd),
                                             min(64, xs.shape[0] - i), self._stream())
            self._check(rc, "sspbo_blr_update_x")

    def observe(self, x, t):
        """(mu, var) of the BLR at phi(x) under the current posterior, then the posterior update with target t: one C call,
        one synchronisation (eval(x_t) + update(x_t, y_t) of a BO step)."""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
        if x.shape[0] != self.n * self.L:
            raise AssertionError(f"Expected {self.n * self.L}-d data, got {tuple(x.shape)}")
        mu, var = C.c_double(), C.c_double()
        rc = self.lib.sspbo_observe(self._ctx, x.ctypes.data_as(_lib._pd), float(t), C.byref(mu), C.byref(var), self._stream())
        self._check(rc, "sspbo_observe")
        return mu.value, var.value

    def blr_get(self, m=True, S=True, Sinv=False):
        torch = self.torch
        d = self.d
        mt = torch.empty(d, dtype=torch.float64, device=self._dev()) if m else None
        St = torch.empty((d, d), dtype=torch.float64, device=self._dev()) if S else None
        Si = torch.empty((d, d), dtype=torch.float64, device=self._dev()) if Sinv else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self._check(self.lib.sspbo_blr_get(self._ctx, ptr(mt), ptr(St), ptr(Si), self._stream()), "sspbo_blr_get")
        return mt, St, Si

    def blr_export(self):
        n = self.lib.sspbo_blr_export_size(self._ctx)
        buf = self.torch.empty(n, dtype=self.torch.float64, device=self._dev())
        self._check(self.lib.sspbo_blr_export(self._ctx, C.c_void_p(buf.data_ptr()), self._stream()), "sspbo_blr_export")
        buf.sspbo_lr_rank = self.blr_rank()         # travels with the tensor inside this process: lets blr_import skip a read-back
        return buf

    def blr_checkpoint(self) -> None:
        """Remember the posterior as it is now (low-rank form); `blr_rollback` returns to it."""
        self._check(self.lib.sspbo_blr_checkpoint(self._ctx, self._stream()), "sspbo_blr_checkpoint")

    def blr_rollback(self) -> bool:
        """Undo the observations made since `blr_checkpoint` (one small kernel).  False -- nothing changed -- when the dense
        state has absorbed one of them in the meantime; restore an exported state (`blr_import`) then."""
        rc = self.lib.sspbo_blr_rollback(self._ctx, self._stream())
        if rc == -4:                                  # SSPBO_EUNSUPPORTED
            return False
        self._check(rc, "sspbo_blr_rollback")
        return True

    def blr_rank(self) -> int:
        """Number of valid low-rank rows of the posterior (= observations so far), -1 when they are not valid."""
        r = C.c_int(-1)
        self._check(self.lib.sspbo_blr_rank(self._ctx, C.byref(r)), "sspbo_blr_rank")
        return int(r.value)

    def blr_import(self, buf, beta) -> None:
        assert buf.numel() == self.lib.sspbo_blr_export_size(self._ctx)
        rank = getattr(buf, "sspbo_lr_rank", None)
        if rank is None:                            # (a buffer received from another process: the rank word is read back)
            rc = self.lib.sspbo_blr_import(self._ctx, C.c_void_p(buf.data_ptr()), float(beta or 0.0), self._stream())
        else:
            rc = self.lib.sspbo_blr_import_ranked(self._ctx, C.c_void_p(buf.data_ptr()), float(beta or 0.0), int(rank), self._stream())
        self._check(rc, "sspbo_blr_import")

    def blr_predict(self, phi):
        pt = self.to_device(phi, self.torch.float64)
        N = pt.shape[0]
        mu = self.torch.empty(N, dtype=self.torch.float64, device=self._dev())
        var = self.torch.empty(N, dtype=self.torch.float64, device=self._dev())
        rc = self.lib.sspbo_blr_predict(self._ctx, C.c_void_p(pt.data_ptr()), N, C.c_void_p(mu.data_ptr()),
                                        C.c_void_p(var.data_ptr()), self._stream())
        self._check(rc, "sspbo_blr_predict")
        return mu, var

    # ------------------------------------------------------------------ K2
    def score(self, x, lam, gamma_t, index0=0, want_outputs=False, engine=_lib.ENGINE_AUTO, reset_best=True):
        """Fused encode + acquisition + argmax over candidates x (N, n*L), host or device.
        Returns (best_key_tensor, (mu, var, acq) | None); all device tensors, nothing is synchronised."""
        torch = self.torch
        xt = self.to_device(x)
        if xt.dim() == 1:
            xt = xt.reshape(1, -1)
        if xt.shape[1] != self.n * self.L:
            raise AssertionError(f"Expected {self.n * self.L}-d data, got {tuple(xt.shape)}")
        N = xt.shape[0]
        outs = None
        ptrs = [None, None, None]
        if want_outputs:
            outs = tuple(torch.empty(N, dtype=torch.float32, device=self._dev()) for _ in range(3))
            ptrs = [C.c_void_p(t.data_ptr()) for t in outs]
        if reset_best:
            self._best.zero_()
            self._pending = []
            self._lr_used = False

        def call():
            rc = self.lib.sspbo_score(self._ctx, C.c_void_p(xt.data_ptr()), self._dtype_code(xt), N, int(index0),
                                      float(lam), float(gamma_t), ptrs[0], ptrs[1], ptrs[2],
                                      C.c_void_p(self._best.data_ptr()), int(engine), self._stream())
            self._check(rc, "sspbo_score")
        call()
        self._pending.append(call)
        self._lr_used = self._lr_used or self.lib.sspbo_last_engine(self._ctx) in _RERUN_ENGINES
        if want_outputs and self._lr_used and self._needs_rerun():   # outputs are about to be read: settle them now
            self._replay()
        return self._best, outs

    def score_generated(self, kind, N, lam, gamma_t, index0=0, seed=0, axes=None, want_outputs=False,
                        engine=_lib.ENGINE_AUTO, reset_best=True):
        """Fused scoring of rows [index0, index0 + N) of a candidate set generated inside the kernel: kind 'uniform'
        (counter-based uniform(0, 1) stream, `seed`) or 'grid' (`axes`: per-axis numpy arrays, reference meshgrid order).
        No candidate byte is read from HBM or the host.  Returns (best_key_tensor, (mu, var, acq) | None)."""
        torch = self.torch
        outs, ptrs = None, [None, None, None]
        if want_outputs:
            outs = tuple(torch.empty(N, dtype=torch.float32, device=self._dev()) for _ in range(3))
            ptrs = [C.c_void_p(t.data_ptr()) for t in outs]
        if kind == "uniform":
            code, ax_ptr, cnt_ptr, keep = _lib.CAND_UNIFORM, None, None, None
        elif kind == "grid":
            counts = np.array([len(a) for a in axes], dtype=np.int32)
            assert len(axes) == self.n, f"Expected {self.n} axes, got {len(axes)}"
            flat = self._grid_axes_device(axes)
            code, ax_ptr, cnt_ptr, keep = _lib.CAND_GRID, C.c_void_p(flat.data_ptr()), counts.ctypes.data_as(_lib._pi), (flat, counts)
        else:
            raise ValueError(f"unknown generated candidate kind {kind!r}")
        if reset_best:
            self._best.zero_()
            self._pending = []
            self._lr_used = False

        def call():
            rc = self.lib.sspbo_score_generated(self._ctx, code, int(seed), ax_ptr, cnt_ptr, int(N), int(index0), float(lam),
                                                float(gamma_t), ptrs[0], ptrs[1], ptrs[2], C.c_void_p(self._best.data_ptr()),
                                                int(engine), self._stream())
            self._check(rc, "sspbo_score_generated")
        call()
        self._pending.append(call)
        self._gen_keep = keep
        self._lr_used = self._lr_used or self.lib.sspbo_last_engine(self._ctx) in _RERUN_ENGINES
        if want_outputs and self._lr_used and self._needs_rerun():
            self._replay()
        return self._best, outs

    def _grid_axes_device(self, axes):
        """per-axis values concatenated on the device; cached while the same arrays are passed"""
        key = tuple((id(a), len(a)) for a in axes)
        cached = getattr(self, "_axes_cache", None)
        if cached is None or cached[0] != key:
            flat = self.to_device(np.concatenate([np.asarray(a, dtype=np.float64) for a in axes]), self.torch.float64)
            self._axes_cache = (key, flat, list(axes))
        return self._axes_cache[1]

    def set_operand_format(self, fmt):
        """Operand format of the low-rank tensor-memory engine: None / -1 by rank, 0 split fp16, 1 fp16 + e4m3."""
        self._check(self.lib.sspbo_set_operand_format(self._ctx, -1 if fmt is None else int(fmt)), "sspbo_set_operand_format")

    def eval_host(self, x, lam, gamma_t, engine=_lib.ENGINE_AUTO):
        """(mu, var, acq) as float64 numpy arrays with a single device->host copy (SSPAgent.eval)."""
        torch = self.torch
        if isinstance(x, np.ndarray) and engine in (_lib.ENGINE_AUTO, _lib.ENGINE_EXACT):
            xh = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
            if 0 < xh.shape[0] <= 64:                     # x_t of a BO step: one C call, one synchronisation
                if xh.shape[1] != self.n * self.L:
                    raise AssertionError(f"Expected {self.n * self.L}-d data, got {tuple(xh.shape)}")
                out = np.empty((3, xh.shape[0]), dtype=np.float64)
                rc = self.lib.sspbo_eval_host(self._ctx, xh.ctypes.data_as(_lib._pd), xh.shape[0], float(lam), float(gamma_t),
                                              out[0].ctypes.data_as(_lib._pd), out[1].ctypes.data_as(_lib._pd),
                                              out[2].ctypes.data_as(_lib._pd), self._stream())
                self._check(rc, "sspbo_eval_host")
                return out[0], out[1], out[2]
        xt = self.to_device(x)
        if xt.dim() == 1:
            xt = xt.reshape(1, -1)
        if xt.shape[1] != self.n * self.L:
            raise AssertionError(f"Expected {self.n * self.L}-d data, got {tuple(xt.shape)}")
        N = xt.shape[0]
        out = torch.empty((3, max(N, 1)), dtype=torch.float32, device=self._dev())
        ptrs = [C.c_void_p(out[i].data_ptr()) for i in range(3)]
        # eval() may sit between chunked best_candidate(..., reset_best=False) calls: the running key and its pending calls are
        # settled first and left alone; this call scores into a key of its own
        self.settle()
        key = torch.zeros(1, dtype=torch.int64, device=self._dev())

        def call():
            rc = self.lib.sspbo_score(self._ctx, C.c_void_p(xt.data_ptr()), self._dtype_code(xt), N, 0, float(lam),
                                      float(gamma_t), ptrs[0], ptrs[1], ptrs[2], C.c_void_p(key.data_ptr()),
                                      int(engine), self._stream())
            self._check(rc, "sspbo_score")
        call()
        if self.lib.sspbo_last_engine(self._ctx) in _RERUN_ENGINES:
            for _ in range(3):                                       # flag-list overflow / out-of-bound rows: the context switched engines
                if not self._needs_rerun():
                    break
                call()
        host = out[:, :N].cpu().numpy().astype(np.float64)           # one D2H; the widening happens on the host
        return host[0], host[1], host[2]

    def score_host(self, x_host, lam, gamma_t, index0=0, engine=_lib.ENGINE_AUTO, chunk=1 << 21, reset_best=True,
                   first_piece=None, growth=1.5):
        """Fused scoring of candidates that live in HOST memory (numpy or CPU tensor, ideally pinned): the set is
        streamed through three device staging buffers so the H2D copies of the next chunks overlap the kernel of chunk i.
        Returns the device key tensor (nothing is synchronised)."""
        torch = self.torch
        xt = x_host if isinstance(x_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x_host))
        if xt.dtype not in (torch.float32, torch.float64):
            xt = xt.to(torch.float64)
        xt = xt.contiguous()
        if xt.dim() == 1:
            xt = xt.reshape(1, -1)
        cols = self.n * self.L
        if xt.shape[1] != cols:
            raise AssertionError(f"Expected {cols}-d data, got {tuple(xt.shape)}")
        N = xt.shape[0]
        key = (xt.dtype, cols, chunk)
        if getattr(self, "_stage_key", None) != key:
            self._stage = [torch.empty((chunk, cols), dtype=xt.dtype, device=self._dev()) for _ in range(3)]
            self._stage_key = key
            self._copy_stream = torch.cuda.Stream(device=self._dev())
        compute = torch.cuda.current_stream(self.device)
        if reset_best:
            self._best.zero_()
            self._pending = []
            self._lr_used = False
        call = lambda: self._score_host_stream(xt, N, lam, gamma_t, index0, engine, chunk, first_piece, growth)
        call()
        self._pending.append(call)
        self._lr_used = self._lr_used or self.lib.sspbo_last_engine(self._ctx) in _RERUN_ENGINES
        return self._best

    def _score_host_stream(self, xt, N, lam, gamma_t, index0, engine, chunk, first_piece=None, growth=1.5):
        torch = self.torch
        compute = torch.cuda.current_stream(self.device)
        nbuf = len(self._stage)                                  # three staging buffers: the copy engine may run two pieces ahead
        copied = [torch.cuda.Event() for _ in range(nbuf)]
        done = [torch.cuda.Event() for _ in range(nbuf)]
        self._copy_stream.wait_stream(compute)                   # staging buffers may still be read by earlier work
        # short first piece: its copy is the only one no scoring overlaps; the following ones grow by 1.5x, about the ratio
        # of link to scoring rate at d = 1023, so the copy of piece i + 1 fits under the scoring of piece i.  The schedule
        # ramps down again at the end: the scoring of the LAST piece is the only one no copy overlaps
        # (12.5e6 candidates: 0.5, 0.8, 1.2, 1.8, 2.1, ..., 1.8, 1.2, 0.8, 0.5 million).
        pieces = stream_pieces(N, chunk, first_piece, growth)
        for i, (c0, n) in enumerate(pieces):
            b = i % nbuf
            with torch.cuda.stream(self._copy_stream):
                if i >= nbuf:
                    self._copy_stream.wait_event(done[b])
                self._stage[b][:n].copy_(xt[c0:c0 + n], non_blocking=True)
                copied[b].record(self._copy_stream)
            compute.wait_event(copied[b])
            rc = self.lib.sspbo_score(self._ctx, C.c_void_p(self._stage[b].data_ptr()), self._dtype_code(xt), n,
                                      int(index0 + c0), float(lam), float(gamma_t), None, None, None,
                                      C.c_void_p(self._best.data_ptr()), int(engine), C.c_void_p(compute.cuda_stream))
            self._check(rc, "sspbo_score")
            done[b].record(compute)
        return self._best

    def score_stats(self, reset=True):
        """Counters of the low-rank engine since the last reset (synchronises): dict(scored, flagged, overflow,
        out_of_range, rerun).  `rerun` means results since the last reset must be recomputed (policy already switched)."""
        v = [C.c_int64() for _ in range(4)]
        rr = C.c_int()
        rc = self.lib.sspbo_score_stats(self._ctx, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(v[3]), C.byref(rr),
                                        1 if reset else 0, self._stream())
        self._check(rc, "sspbo_score_stats")
        return dict(scored=v[0].value, flagged=v[1].value, overflow=v[2].value, out_of_range=v[3].value, rerun=bool(rr.value))

    def lowrank_info(self):
        v = [C.c_int() for _ in range(4)]
        self._check(self.lib.sspbo_lowrank_info(self._ctx, *[C.byref(x) for x in v]), "sspbo_lowrank_info")
        return dict(rank=v[0].value, valid=bool(v[1].value), disabled=bool(v[2].value), phase32=bool(v[3].value))

    def set_policy(self, lowrank=None, phase32=None):
        enc = lambda f: -1 if f is None else int(f)         # phase32: False/0 Q31.32 only, True/1 fastest, 2 fixed point only
        self._check(self.lib.sspbo_set_policy(self._ctx, enc(lowrank), enc(phase32)), "sspbo_set_policy")

    def _needs_rerun(self):
        return self.score_stats(reset=True)["rerun"]

    def settle(self):
        """Make the key / outputs of the score calls since the last reset final (synchronises): if the low-rank pass
        overflowed its flag list or saw candidates outside the declared bound, the calls are replayed with the safe
        engines.  Call before combining keys across ranks; `best()` calls it."""
        if self._pending:
            if self._lr_used and self._needs_rerun():    # only the low-rank engine can ask for a rerun
                self._replay()
            self._pending = []

    def _replay(self):
        """Recompute every score call since the last key reset (the context has switched to a safer engine: low-rank form ->
        factor form -> the engines that never flag; fast phase path -> Q31.32)."""
        calls, self._pending = self._pending, []
        st = None
        for _ in range(3):
            self._best.zero_()
            for c in calls:
                c()
            st = self.score_stats(reset=True)
            if not st["rerun"]:
                break
        self._pending = calls
        if st["rerun"]:
            raise _lib.SspboError(f"scoring still asks for a rerun after the policy switches: {st}")

    def best(self):
        """(score, index) of the packed key; synchronises.  If the low-rank pass overflowed its flag list or saw
        candidates outside the declared bound, the step is recomputed with the safe engines first."""
        st, key = self._finish()                          # counters of the low-rank pass and the key in one synchronisation
        if self._pending and self._lr_used and st["rerun"]:
            self._replay()
            _, key = self._finish()
        self._pending = []
        return _lib.key_unpack(key)

    # ------------------------------------------------------------------ multi-GPU: key exchange over peer memory
    def dist_connect(self) -> bool:
        """Collective over torch.distributed (call on every rank): sets up the peer-memory key exchange of sspbo_dist.cu --
        every rank exports its slot array, the 64-byte handles are all-gathered, every rank maps its peers' arrays.  Returns
        whether the exchange is usable on ALL ranks (otherwise the NCCL all-reduce of dist.allreduce_best stays in use)."""
        import os
        import torch.distributed as td
        rank, world = td.get_rank(), td.get_world_size()
        ok, mine = 1, bytes(64)
        if os.environ.get("SSPBO_DIST_EXCHANGE", "1") == "0" or world > 16:
            ok = 0
        else:
            buf = (C.c_ubyte * 64)()
            if self.lib.sspbo_dist_init(self._ctx, rank, world, buf) == 0:
                mine = bytes(buf)
            else:
                ok = 0
        gathered = [None] * world
        td.all_gather_object(gathered, (ok, mine))
        if all(g[0] for g in gathered):
            blob = b"".join(g[1] for g in gathered)
            arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            ok = 1 if self.lib.sspbo_dist_connect(self._ctx, arr) == 0 else 0
        else:
            ok = 0
        flag = self.torch.tensor([ok], dtype=self.torch.int32, device=self._dev())
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        self._exchange_on = bool(int(flag.item()))
        return self._exchange_on

    def best_global(self):
        """(score, index) of the best key over ALL ranks (call on every rank; synchronises once): one kernel writes this
        rank's key into every peer's slot array and takes the maximum of what the peers wrote (sspbo_dist_exchange), then the
        usual read-back.  When any rank's scoring pass asks for a rerun, every rank recomputes its shard and exchanges again."""
        best_ptr = C.c_void_p(self._best.data_ptr())
        self._check(self.lib.sspbo_dist_exchange(self._ctx, best_ptr, self._stream()), "sspbo_dist_exchange")
        st, key = self._finish()
        if st["rerun"]:
            if self._pending:
                self._replay()                            # (zeroes the key, recomputes this rank's calls with the safer engines)
            else:
                self._best.zero_()
            self._check(self.lib.sspbo_dist_exchange(self._ctx, best_ptr, self._stream()), "sspbo_dist_exchange")
            st, key = self._finish()
            if st["rerun"]:
                raise _lib.SspboError(f"scoring still asks for a rerun after every rank recomputed its shard: {st}")
        self._pending = []
        return _lib.key_unpack(key)

    def _finish(self):
        v = [C.c_int64() for _ in range(4)]
        rr, key = C.c_int(), C.c_uint64()
        rc = self.lib.sspbo_score_finish(self._ctx, C.c_void_p(self._best.data_ptr()), C.byref(key), C.byref(v[0]),
                                         C.byref(v[1]), C.byref(v[2]), C.byref(v[3]), C.byref(rr), 1, self._stream())
        self._check(rc, "sspbo_score_finish")
        return dict(scored=v[0].value, flagged=v[1].value, overflow=v[2].value, out_of_range=v[3].value,
                    rerun=bool(rr.value)), int(key.value)

    def last_engine(self) -> str:
        return _lib.ENGINE_NAMES.get(self.lib.sspbo_last_engine(self._ctx), "?")

    def launch_count(self) -> int:
        return int(self.lib.sspbo_launch_count(self._ctx))

    def set_timing(self, enable: bool) -> None:
        """CUDA events around the scorer kernel and the float64 pass over flagged candidates of every score call."""
        self._check(self.lib.sspbo_set_timing(self._ctx, 1 if enable else 0), "sspbo_set_timing")

    def timing_read(self, reset=True):
        """dict(scorer_ms, flagged_ms, calls) summed over the score calls since the last reset (synchronises)."""
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        rc = self.lib.sspbo_timing_read(self._ctx, C.byref(a), C.byref(b), C.byref(n), 1 if reset else 0, self._stream())
        self._check(rc, "sspbo_timing_read")
        return dict(scorer_ms=a.value, flagged_ms=b.value, calls=n.value)

    # ------------------------------------------------------------------ K4
    def from_set_argmax(self, sample_ssps, queries):
        st = self.to_device(sample_ssps, self.torch.float64)
        qt = self.to_device(queries, self.torch.float64)
        if qt.dim() == 1:
            qt = qt.reshape(1, -1)
        assert st.shape[1] == qt.shape[1], f"Expected {tuple(st.shape)} dim, got {tuple(qt.shape)}"
        if self.d == 0:
            raise _lib.SspboError("from_set_argmax before set_space")
        assert st.shape[1] == self.d
        idx = self.torch.empty(qt.shape[0], dtype=self.torch.int64, device=self._dev())
        rc = self.lib.sspbo_from_set_argmax(self._ctx, C.c_void_p(st.data_ptr()), st.shape[0], C.c_void_p(qt.data_ptr()),
                                            qt.shape[0], C.c_void_p(idx.data_ptr()), self._stream())
        self._check(rc, "sspbo_from_set_argmax")
        return idx

    def decode_optim(self, queries, x0, bounds=None):
        """Direct-optim decode on the device: argmax_x phi(x) . unit(query) from x0 inside `bounds` (n, 2).
        Returns a device tensor (q, n) float64."""
        qt = self.to_device(queries, self.torch.float64)
        if qt.dim() == 1:
            qt = qt.reshape(1, -1)
        assert qt.shape[1] == self.d, f"Expected {self.d}-d SSPs, got {tuple(qt.shape)}"
        xt = self.to_device(np.atleast_2d(x0) if not hasattr(x0, "is_cuda") else x0, self.torch.float64)
        assert tuple(xt.shape) == (qt.shape[0], self.n), (tuple(xt.shape), qt.shape[0], self.n)
        out = self.torch.empty_like(xt)
        bp = None
        if bounds is not None:
            b = np.ascontiguousarray(np.asarray(bounds, dtype=np.float64).reshape(self.n, 2))
            bp = b.ctypes.data_as(_lib._pd)
        rc = self.lib.sspbo_decode_optim(self._ctx, C.c_void_p(qt.data_ptr()), qt.shape[0], C.c_void_p(xt.data_ptr()), bp,
                                         C.c_void_p(out.data_ptr()), self._stream())
        self._check(rc, "sspbo_decode_optim")
        return out

    def gram(self, phi):
        """G = phi phi^T (n, n) float64 device tensor for n encoded rows (n, d)."""
        pt = self.to_device(phi, self.torch.float64)
        assert pt.dim() == 2 and pt.shape[1] == self.d, f"Expected (n, {self.d}) rows, got {tuple(pt.shape)}"
        out = self.torch.empty((pt.shape[0], pt.shape[0]), dtype=self.torch.float64, device=self._dev())
        rc = self.lib.sspbo_gram(self._ctx, C.c_void_p(pt.data_ptr()), pt.shape[0], C.c_void_p(out.data_ptr()), self._stream())
        self._check(rc, "sspbo_gram")
        return out

    # ------------------------------------------------------------------ K6
    def acq_ascent(self, phi0, lam, gamma_t, margin=4.0, max_iter=200, tol=1e-7):
        """Maximise the acquisition in phi-space from the rows of phi0 (R, d), all restarts in one cooperative kernel.
        Returns device tensors (phi (R, d) float64, values (R,) float64, steps (R,) int32)."""
        torch = self.torch
        pt = self.to_device(phi0, torch.float64)
        if pt.dim() == 1:
            pt = pt.reshape(1, -1)
        assert pt.shape[1] == self.d, f"Expected {self.d}-d starting points, got {tuple(pt.shape)}"
        R = pt.shape[0]
        out = torch.empty_like(pt)
        val = torch.empty(R, dtype=torch.float64, device=self._dev())
        its = torch.empty(R, dtype=torch.int32, device=self._dev())
        rc = self.lib.sspbo_acq_ascent(self._ctx, C.c_void_p(pt.data_ptr()), R, float(lam), float(gamma_t), float(margin),
                                       int(max_iter), float(tol), C.c_void_p(out.data_ptr()), C.c_void_p(val.data_ptr()),
                                       C.c_void_p(its.data_ptr()), self._stream())
        self._check(rc, "sspbo_acq_ascent")
        return out, val, its

    # ------------------------------------------------------------------ K5
    def bind(self, a, b):
        """Circular convolution of the rows of a (qa, d) and b (qb, d) with numpy broadcasting over rows."""
        at = self.to_device(np.atleast_2d(a) if not hasattr(a, "is_cuda") else a, self.torch.float64)
        bt = self.to_device(np.atleast_2d(b) if not hasattr(b, "is_cuda") else b, self.torch.float64)
        assert at.shape[1] == self.d and bt.shape[1] == self.d, (tuple(at.shape), tuple(bt.shape), self.d)
        q = max(at.shape[0], bt.shape[0])
        out = self.torch.empty((q, self.d), dtype=self.torch.float64, device=self._dev())
        rc = self.lib.sspbo_bind(self._ctx, C.c_void_p(at.data_ptr()), at.shape[0], C.c_void_p(bt.data_ptr()), bt.shape[0],
                                 C.c_void_p(out.data_ptr()), self._stream())
        self._check(rc, "sspbo_bind")
        return out

    # ------------------------------------------------------------------ generators
    def fill_uniform(self, seed: int, start: int, count: int, n: int | None = None):
        n = n or self.n * self.L
        out = self.torch.empty((count, n), dtype=self.torch.float32, device=self._dev())
        rc = self.lib.sspbo_fill_uniform(self._ctx, int(seed), int(start), int(count), int(n), C.c_void_p(out.data_ptr()),
                                         self._stream())
        self._check(rc, "sspbo_fill_uniform")
        return out

    def grid_points(self, axes, start=0, count=None):
        """axes: list of per-axis numpy linspaces.  Rows [start, start+count) in the reference's meshgrid order."""
        counts = np.array([len(a) for a in axes], dtype=np.int32)
        total = int(np.prod(counts.astype(np.int64)))
        count = total - start if count is None else count
        flat = self.to_device(np.concatenate([np.asarray(a, dtype=np.float64) for a in axes]), self.torch.float64)
        out = self.torch.empty((count, len(axes)), dtype=self.torch.float64, device=self._dev())
        rc = self.lib.sspbo_grid_points(self._ctx, C.c_void_p(flat.data_ptr()), counts.ctypes.data_as(_lib._pi), len(axes),
                                        int(start), int(count), C.c_void_p(out.data_ptr()), self._stream())
        self._check(rc, "sspbo_grid_points")
        return out
