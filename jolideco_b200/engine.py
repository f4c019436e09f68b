"""Fused MAP engine: the body of the reference's inner loop (jolideco/core.py:214-229) as a fixed
sequence of C-ABI kernel launches on one CUDA stream, replayed from a CUDA graph.

No autograd, no host synchronisation inside a step: the two `torch.randint` draws of the cycle spin
(utils/torch.py:108-116) are pre-drawn on the host into a device table, the Adam step counter and
bias corrections live on the device (`jd_step_begin`), and losses accumulate into device doubles
that are only read for the per-epoch trace (loss.py:212-250).

Two step semantics (SURVEY.md §7 "semantics of iteration"):
  * `step(d)`      — the reference's: one dataset's likelihood + the full prior scaled 1/D + Adam;
  * `joint_step()` — one pass over all (local) datasets + one prior + one Adam on
                     sum_d L_d - beta * prior (`TotalLoss.__call__`, loss.py:257-261); with
                     `torch.distributed` the datasets are sharded over ranks, the prior is
                     row-block sharded and the flux gradient is all-reduced (NCCL / NVLink).
"""
import os

import numpy as np
import torch

from . import _lib, dist, ops

_p = ops._ptr

LIK_DTYPE = ops.LIK_DTYPE
# datasets on the FFT path go out together (grid.y = dataset, Poisson statistic fused into the inverse row pass);
# JD_FFT_BATCHED=0 keeps the per-dataset convolution / Poisson / adjoint launches
FFT_BATCHED = os.environ.get("JD_FFT_BATCHED", "1") != "0"

# launch accounting / per-kernel timing hooks (bench.py): every C-ABI call below is exactly one
# kernel launch on the current stream
_STATS = {"launches": 0, "timed": None, "events": [], "dry": False}


# kernels launched by entry points that launch more than one
_KERNELS_PER_CALL = {"jd_conv_forward_fft": 3, "jd_conv_backward_fft": 3, "jd_likelihood_forward_fft": 3,
                     "jd_likelihood_backward_fft": 3}


_SPLIT_CACHE = {}  # tuned `split_clusters` per engine shape (see MapEngine._tune_split)
_PEER_POOL = {}  # free symmetric-memory buffer sets per (device, process group, pixels) (see MapEngine._enable_peer)


def _call(name, *args):
    if _STATS["dry"]:  # argument-building pass before a graph capture (lazy device tables are created here)
        return
    n = _KERNELS_PER_CALL.get(name, 1)
    if name == "jd_gmm_prior_backward" and args[-2] is not None:
        n = 4  # histogram, scan, scatter, bucketed GEMV
    _STATS["launches"] += n
    if _STATS["timed"] == name or _STATS["timed"] == "*":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call(name, *args)
        e1.record()
        _STATS["events"].append((e0, e1) if _STATS["timed"] == name else (name, e0, e1))
        return
    _lib.call(name, *args)


class DatasetBuffers:
    """Device-resident arrays of one dataset (loss.py:104-118, npred.py:281-295)."""

    def __init__(self, counts, exposure, psf, background, f, bkg_log_norm=None, name="", train_bkg_norm=False,
                 shift_xy=None, train_shift=False):
        self.counts, self.exposure, self.psf, self.background = counts, exposure, psf, background
        self.f = int(f) if f else 1
        # log background norm (NPredCalibration._background_norm, a (1,) CUDA float tensor sharing the
        # parameter's storage) and, when it is trained, its private Adam state
        self.bkg_log_norm = bkg_log_norm
        self.train_bkg_norm = bool(train_bkg_norm) and bkg_log_norm is not None
        if self.train_bkg_norm:
            self.cal_m = torch.zeros_like(bkg_log_norm)
            self.cal_v = torch.zeros_like(bkg_log_norm)
            self.cal_t = torch.zeros(1, dtype=torch.int32, device=bkg_log_norm.device)
        # non-zero sub-pixel shift (NPredCalibration.shift_xy, a (2,) CUDA float view of the parameter's storage:
        # shift_x, shift_y) and, when it is trained, its private Adam state (one step counter for the pair, as
        # torch.optim.Adam keeps one per parameter tensor).  None = no shift operator (the reference skips it, and
        # never trains the shift, when it is ~0: utils/torch.py:211)
        self.shift_xy = shift_xy
        self.train_shift = bool(train_shift) and shift_xy is not None
        if self.train_shift:
            self.shift_m = torch.zeros_like(shift_xy)
            self.shift_v = torch.zeros_like(shift_xy)
            self.shift_t = torch.zeros(1, dtype=torch.int32, device=shift_xy.device)
        self.name = name
        self.H, self.W = int(counts.shape[-2]), int(counts.shape[-1])
        self.fH, self.fW = int(exposure.shape[-2]), int(exposure.shape[-1])
        self.kh, self.kw = int(psf.shape[-2]), int(psf.shape[-1])
        for t in (counts, exposure, psf, background):
            ops._check(t, "dataset buffer")
        # PSF rows of up to ~37 taps can go through the batched direct kernels with the Poisson statistic fused into the
        # convolution epilogue (jd_likelihood_forward / _backward) - the engine takes them when a launch fills the SMs
        # (MapEngine._use_batched); JD_LIK_BATCHED=0 keeps the separate conv / Poisson / conv launches.  Larger PSFs:
        # cached PSF spectrum + scratch for the shared-memory FFT path (built on first use).
        self.lik_ok = (os.environ.get("JD_LIK_BATCHED", "1") != "0"
                       and self.H * self.f == self.fH and self.W * self.f == self.fW
                       and _lib.load().jd_likelihood_supported(self.kh, self.kw, self.f) == 1)
        self._fft = None
        # counts-only Stirling term of nn.PoissonNLLLoss(full=True) (loss.py:35-37): a constant of the dataset
        self.loss_const = ops.stirling_constant(counts)
        self.geom = (self.fH, self.fW, self.kh, self.kw, self.f, self.H, self.W)
        self.n_tiles = ((self.fH + 63) // 64) * ((self.fW + 63) // 64)
        self.flux_s = self.dflux_s = self.dpool = None  # per-dataset scratch, allocated by the engine

    @property
    def fft(self):
        """FFT plan (cached PSF spectrum + workspace) for PSFs of at least ops.FFT_MIN_PSF_AREA taps, else None."""
        if self._fft is None and self.kh * self.kw >= ops.FFT_MIN_PSF_AREA:
            self._fft = ops.FFTConvPlan(self.psf, self.fH, self.fW)
        return self._fft


# The batched likelihood kernels give every thread an 8 x 8 output block (64 x 64 outputs per 64-thread CTA): a launch
# needs about a CTA per SM before they beat the FFT / split-row direct kernels, which spread a small image over more
# threads (cfg2's single 512^2 dataset is 64 CTAs: 209 us against 29 us for the FFT path).  JD_LIK_MIN_CTAS overrides.
LIK_MIN_CTAS = int(os.environ.get("JD_LIK_MIN_CTAS", "148"))


class MapEngine:
    def __init__(self, theta, datasets, prior=None, mask=None, use_log_flux=True, beta=1.0, lr=0.1, betas=(0.9, 0.999),
                 eps=1e-8, shift_table=None, datasets_validation=(), use_graph=True, process_group=None,
                 prior_weight=None, dataset_index=None, n_datasets_global=None, validation_index=None,
                 n_validation_global=None, counts_shape=None, collective="nccl", stream_k=None, overlap=None):
        """theta: 2-D CUDA fp32 tensor updated in place (the component's parameter storage).
        prior: None (uniform) or dict(packed=GMMPacked, stride, marginalize, backend).
        shift_table: (N, 2) int array of pre-drawn (row, col) cycle-spin shifts in consumption order.
        stream_k: stream-K decomposition of the tcgen05 prior forward (None = ops.use_stream_k: only when the
        patch tiles do not fill the SMs; it buys latency of ONE run, not throughput of concurrent runs).
        overlap: run the likelihood chain (convolutions, Poisson) and the prior chain (tensor-core forward, backward)
        of a step on two streams - both only depend on the flux and meet at the Adam kernel; the prior kernels hold one
        CTA per SM with ~27 KB of shared memory to spare, so the FP32 / FFT kernels co-reside.  None = on unless
        JD_OVERLAP=0 (parity suite green and cfg2 step 172 -> 153 us with it, profiles/r02_summary.md)."""
        ops.require_device(theta.device)
        self.theta = ops._check(theta, "theta")
        assert theta.ndim == 2
        self.dev = theta.device
        self.fH, self.fW = (int(s) for s in theta.shape)
        self.n = self.fH * self.fW
        self.mask = ops._check(mask, "mask", torch.uint8)
        self.use_log_flux = bool(use_log_flux)
        self.datasets = list(datasets)
        self.datasets_validation = list(datasets_validation)
        self.D = len(self.datasets)
        # multi-GPU: this rank holds datasets `dataset_index` of `n_datasets_global` (dist.shard_indices)
        self.dataset_index = list(range(self.D)) if dataset_index is None else list(dataset_index)
        self.Dg = self.D if n_datasets_global is None else int(n_datasets_global)
        self.validation_index = (list(range(len(self.datasets_validation))) if validation_index is None
                                 else list(validation_index))
        self.Vg = len(self.datasets_validation) if n_validation_global is None else int(n_validation_global)
        if counts_shape is None:
            ref = (self.datasets + self.datasets_validation)[0]
            counts_shape = (ref.H, ref.W)
        self.counts_shape = tuple(counts_shape)
        self.prior_weight = self.Dg if prior_weight is None else prior_weight
        self.beta, self.lr, self.b1, self.b2, self.eps = float(beta), float(lr), float(betas[0]), float(betas[1]), float(eps)
        self.prior = prior
        self.pg = process_group
        self.world = 1 if process_group is None else torch.distributed.get_world_size(process_group)
        self.rank = 0 if process_group is None else torch.distributed.get_rank(process_group)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.m = torch.zeros_like(theta)
        self.v = torch.zeros_like(theta)
        self.flux = torch.empty_like(theta)
        self.conv = torch.empty_like(theta)
        # per-dataset likelihood gradients d L_d / d flux of a (joint) step: summed, in a fixed order, by the Adam /
        # local-reduce kernel instead of read-modify-write accumulation by every adjoint launch
        self.parts = torch.zeros((max(self.D, 1), self.fH, self.fW), **f32)
        self.dflux_l = self.parts[0]
        self.counters = torch.zeros(2, dtype=torch.int32, device=self.dev)
        self.cur_shift = torch.zeros(2, dtype=torch.int32, device=self.dev)
        self.adam_scalars = torch.zeros(2, **f32)
        # acc[0] = Poisson loss sum of the last reference step, acc[1] = sum_p v_p of the last prior evaluation, then three
        # slots per local dataset j: acc[2+3j] = d loss / d log(background norm), acc[3+3j : 5+3j] = d loss / d (shift_x,
        # shift_y), then one slot per local dataset: its Poisson loss sum in the last JOINT step (kept apart so that the
        # per-epoch trace can reuse them, `_trace_body`)
        Dm = max(self.D, 1)
        self._loss0 = 2 + 3 * Dm
        self.acc = torch.zeros(self._loss0 + Dm, dtype=torch.float64, device=self.dev)
        self._last_joint = False
        for d in self.datasets + self.datasets_validation:
            if d.shift_xy is not None and d.flux_s is None:
                d.flux_s = torch.empty_like(theta)   # this dataset's shifted flux
                d.dflux_s = torch.empty_like(theta)  # gradient w.r.t. it
        for d in self.datasets:
            if d.dpool is None:
                d.dpool = torch.empty((d.H, d.W), **f32)
        self._tables = {}
        self._trace_dst = None
        self.n_trace = self.Dg + 1 + self.Vg
        self.acc_trace = torch.zeros(self.n_trace, dtype=torch.float64, device=self.dev)
        self.shift_table = None
        self.n_shifts = 0
        if prior is not None:
            self.stride = int(prior["stride"])
            self.marginalize = bool(prior["marginalize"])
            self.backend = int(prior.get("backend", 0))
            self.packed = prior["packed"]
            self.sk_ws = None
            if self.backend == 1:
                self.packed.Bt  # pack the tensor-core operand before any graph capture
            elif self.backend == 2:
                ops._bt16(self.packed)
            elif self.backend in (3, 4):
                ops._btm(self.packed)
            elif self.backend == 5:
                ops._bt16(self.packed)
            self.ny, self.nx = ops.patch_grid(self.fH, self.fW, self.stride)
            self.c = self.stride**2 / ops.PD / self.n
            # row-block shard of the prior (whole grid on one GPU)
            self.rows = self._row_block(self.rank, self.world)
            P = (self.rows[1] - self.rows[0]) * self.nx
            self.P = P
            if self.backend == 1 and P > 0 and (ops.use_stream_k(P, self.dev) if stream_k is None else stream_k):
                self.sk_ws = ops.tc_sk_workspace(P, self.packed.K, self.dev)
            if self.backend in (3, 4, 5) and P > 0:
                self.sk_ws = ops.tcm_workspace(P, self.packed.K, self.dev, self.backend)
            self.value = torch.empty(max(P, 1), **f32)
            self.argmax = torch.empty(max(P, 1), dtype=torch.int32, device=self.dev)
            # logsumexp mode keeps logp for the backward; the tensor-core kernels use a component-major layout
            self.logp = torch.empty((max(P, 1), self.packed.K), **f32) if self.marginalize else None
            if self.marginalize and self.backend in (1, 2, 3, 4, 5):
                ops._bt_lam(self.packed)
            self.G = torch.empty((max(P, 1), ops.PD), **f32)
            # bucketed max-mode backward (patches grouped by winning component, Lam_k staged once per 64 patches): the
            # faster path at large patch counts (ops.use_bwd_bucketed; JD_BWD_BUCKETED = 0 | 1 forces it)
            self.bwd_ws = None
            if P > 0 and ops.use_bwd_bucketed(P, self.marginalize):
                self.bwd_ws = ops.gmm_backward_workspace(P, self.packed.K, self.dev)
            self.dflux_p = torch.zeros_like(theta)
            if shift_table is None:
                shift_table = np.zeros((1, 2), dtype=np.int32)
            tab = np.ascontiguousarray(np.asarray(shift_table, dtype=np.int32).reshape(-1, 2))
            self.shift_table = torch.from_numpy(tab).to(self.dev)
            self.n_shifts = int(tab.shape[0])
        # multi-rank joint steps replay two graphs (before / after the gradient all-reduce) with the NCCL call
        # launched eagerly in between: capturing the collective inside the graph hung with uneven dataset shards
        self.use_graph = bool(use_graph)
        # "nccl": ncclAllReduce of the flux gradient + replicated Adam; "peer": one fused kernel that reduces the
        # gradient slices over NVLink peer memory, runs Adam on the owned slice and broadcasts theta (jd_peer.cu)
        if collective not in ("nccl", "peer"):
            raise ValueError(f"Unknown collective: {collective}, must be 'nccl' or 'peer'")
        self.collective = collective if self.world > 1 else "nccl"
        self._theta_param = None
        if self.collective == "peer":
            self._enable_peer()
        self._graphs = {}
        self._graph_nodes = {}
        self._capture_stream = None
        self.overlap = (os.environ.get("JD_OVERLAP", "1") == "1") if overlap is None else bool(overlap)
        # steps with ONE dataset: the two-tile prior forward on `split_clusters` CTA pairs, enqueued first, and the
        # likelihood chain on the remaining SMs from the side stream (0: in order on one stream).  "auto": chosen by
        # `warmup` from timings of this engine's own gradient pass (JD_SPLIT_CLUSTERS = auto | n)
        split = os.environ.get("JD_SPLIT_CLUSTERS", "auto")
        self.split_auto = split == "auto"
        self.split_clusters = 0 if self.split_auto else int(split)
        # created outside any capture
        self._side = torch.cuda.Stream(device=self.dev) if (self.overlap and self.dev.type == "cuda") else None

    # ------------------------------------------------------------------------------------------
    def _enable_peer(self):
        """Move theta and the partial-gradient buffer into symmetric memory (peer-addressable over NVLink), plus the
        flag block of the in-kernel cross-rank barriers (jd_adam_allreduce_peer_sync).  Allocation + rendezvous of
        the three buffers costs several hundred ms (handle exchange, mapping every peer): a finished run hands its
        set back (`release_peer`) and the next engine of the same size on the same group takes it over.  Acquire and
        release happen in program order on every rank, so the ranks always pair up the same set."""
        import torch.distributed._symmetric_memory as symm

        if self.n % 4:
            raise _lib.JolidecoB200Error("collective='peer' needs a flux pixel count divisible by 4")
        if self.world > 32:
            raise _lib.JolidecoB200Error("collective='peer' supports up to 32 ranks")
        key = (self.dev.index, getattr(self.pg, "group_name", None) or id(self.pg), self.n)
        free = _PEER_POOL.setdefault(key, [])
        with torch.cuda.device(self.dev):
            if free:
                bufs = free.pop()
            else:
                bufs = {"grad": symm.empty(self.n, dtype=torch.float32, device=self.dev),
                        "theta": symm.empty(self.n, dtype=torch.float32, device=self.dev),
                        "sig": symm.empty(64, dtype=torch.int32, device=self.dev)}
                for name in ("grad", "theta", "sig"):
                    bufs["h_" + name] = symm.rendezvous(bufs[name], self.pg)
            self._peer_key, self._peer_bufs = key, bufs
            self.sym_grad, self.sym_theta, self.sym_sig = bufs["grad"], bufs["theta"], bufs["sig"]
            self.h_grad, self.h_theta, self.h_sig = bufs["h_grad"], bufs["h_theta"], bufs["h_sig"]
            self.sym_grad.zero_()
            self.sym_sig.zero_()
            self.sym_theta.copy_(self.theta.reshape(-1))
            # epoch, finished CTAs (never restored), 2 pad, 4 x int64 time stamps of the last launch (jd_peer.cu)
            self.sync_state = torch.zeros(12, dtype=torch.int32, device=self.dev)
            torch.cuda.synchronize(self.dev)
        self.h_sig.barrier(channel=0)  # every flag block is zeroed before any rank can signal
        self._theta_param = self.theta  # the component's parameter storage: refreshed by sync_theta()
        self.theta = self.sym_theta.view(self.fH, self.fW)
        self.dflux_l = self.sym_grad.view(self.fH, self.fW)  # output of the local gradient reduce, read by the peers

    def release_peer(self):
        """End of a run (every rank, same program point): theta goes back to the component's parameter storage, the
        symmetric buffers to the pool.  The engine cannot step afterwards."""
        if getattr(self, "_peer_bufs", None) is None:
            return
        self.sync_theta()
        torch.cuda.synchronize(self.dev)
        self.theta = self._theta_param
        self._theta_param = None
        self.dflux_l = None
        self._graphs, self._graph_nodes = {}, {}  # they hold the symmetric addresses
        _PEER_POOL[self._peer_key].append(self._peer_bufs)
        self._peer_bufs = None
        self.sym_grad = self.sym_theta = self.sym_sig = self.h_grad = self.h_theta = self.h_sig = None

    def sync_theta(self):
        """Copy the working theta back into the component's parameter storage (peer mode only)."""
        if self._theta_param is not None:
            self._theta_param.copy_(self.theta)

    def _row_block(self, rank, world):
        return dist.row_block(self.ny, rank, world)

    def _s(self):
        return torch.cuda.current_stream().cuda_stream

    def _begin(self, advance_adam, zero_acc, with_shift):
        tab = self.shift_table if (with_shift and self.prior is not None) else None
        _call("jd_step_begin", _p(self.counters), _p(tab), self.n_shifts, _p(self.cur_shift), int(advance_adam),
              self.lr, self.b1, self.b2, _p(self.adam_scalars), _p(zero_acc), int(zero_acc.numel()), self._s())

    def _begin_flux(self, advance_adam, zero_acc):
        """step bookkeeping + flux = exp(theta) in one launch"""
        tab = self.shift_table if self.prior is not None else None
        _call("jd_step_begin_flux", _p(self.counters), _p(tab), self.n_shifts, _p(self.cur_shift), int(advance_adam),
              self.lr, self.b1, self.b2, _p(self.adam_scalars), _p(zero_acc), int(zero_acc.numel()), _p(self.theta),
              _p(self.mask), _p(self.flux), self.n, int(self.use_log_flux), self._s())

    def _adam_joint(self, n_parts, with_G, scale_b):
        """sum of the likelihood gradient parts + fold of the patch gradients G + chain rule + Adam, one launch"""
        G = self.G if with_G else None
        rows = self.rows if self.prior is not None else (0, 0)
        stride = self.stride if self.prior is not None else 1
        _call("jd_adam_joint_step_dev", _p(self.theta), _p(self.m), _p(self.v), _p(self.flux), _p(self.mask),
              _p(self.parts), int(n_parts), self.n, _p(G), float(scale_b), int(self.use_log_flux), self.fH, self.fW,
              _p(self.cur_shift), stride, rows[0], rows[1], _p(self.adam_scalars), self.b1, self.b2, self.eps, self._s())

    def _grad_reduce(self, n_parts, with_G, scale_b):
        """this rank's partial gradient (likelihood parts + fold of its patch rows) -> dflux_l, no update"""
        G = self.G if with_G else None
        rows = self.rows if self.prior is not None else (0, 0)
        stride = self.stride if self.prior is not None else 1
        _call("jd_grad_reduce_local", _p(self.parts), int(n_parts), self.n, _p(G), float(scale_b), self.fH, self.fW,
              _p(self.cur_shift), stride, rows[0], rows[1], _p(self.dflux_l), self._s())

    def _flux(self):
        _call("jd_flux_forward", _p(self.theta), _p(self.mask), _p(self.flux), self.n, int(self.use_log_flux), self._s())

    def _slot(self, j):
        """device address of the calibration-gradient accumulators (dlogb, dshift_x, dshift_y) of local dataset j"""
        return self.acc.data_ptr() + 8 * (2 + 3 * j)

    def _loss_slot(self, j):
        """device address of local dataset j's loss accumulator of a joint step"""
        return self.acc.data_ptr() + 8 * (self._loss0 + j)

    def poisson_sum(self, vals):
        """Poisson loss sum of the last step from a host copy of `acc` (reference step: acc[0]; joint step: the slots)"""
        return float(vals[0] + vals[self._loss0:].sum())

    def _table(self, entries, want_grad, fft=False):
        """Device table of jd_lik_dataset (fft: jd_fftlik_dataset) records for one batched launch pair (cached: every
        pointer is static)."""
        key = (tuple((id(d), j) for d, _, j in entries), bool(want_grad), tuple(lp for _, lp, _ in entries), bool(fft))
        tab = self._tables.get(key)
        if tab is None:
            rec = np.zeros(len(entries), dtype=ops.FFTLIK_DTYPE if fft else LIK_DTYPE)
            for r, (d, loss_ptr, j) in zip(rec, entries):
                if fft:
                    r["workspace"], r["psf_hat"] = _p(d.fft.workspace), _p(d.fft.psf_hat)
                shifted = d.shift_xy is not None
                r["flux"] = _p(d.flux_s if shifted else self.flux)
                r["exposure"], r["psf"] = _p(d.exposure), _p(d.psf)
                r["background"], r["counts"] = _p(d.background), _p(d.counts)
                r["bkg_log_norm"] = _p(d.bkg_log_norm) or 0
                r["loss_sum"] = loss_ptr or 0
                r["loss_const"] = d.loss_const
                if want_grad:
                    r["dpool"] = _p(d.dpool)
                    r["dlogb"] = self._slot(j) if d.train_bkg_norm else 0
                    r["dflux"] = _p(d.dflux_s) if shifted else self.parts.data_ptr() + 4 * self.n * j
            host = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy())
            tab = host.to(self.dev) if self.dev.type == "cuda" else host
            self._tables[key] = tab
        return tab

    def _likelihoods(self, entries, want_grad):
        """Likelihood terms of `entries` = [(dataset, device address of its loss accumulator, gradient part index)]:
        NPred forward + Poisson statistic (+ gradient into parts[j] when want_grad).  Datasets the batched direct
        kernels cover go out in one launch per direction and geometry; the rest (FFT path) one by one."""
        s = self._s()
        groups = {}
        for e in entries:
            if e[0].lik_ok:
                groups.setdefault(e[0].geom, []).append(e)
        groups = {g: es for g, es in groups.items() if self._use_batched(es)}
        batched = [e for es in groups.values() for e in es]
        rest = [e for e in entries if not any(e is b for b in batched)]
        # large PSFs: the FFT path, all datasets of a geometry per launch
        fft_groups = {}
        if FFT_BATCHED:
            for e in rest:
                d = e[0]
                if d.fft is not None and d.f in (1, 2) and d.H * d.f == d.fH and d.W * d.f == d.fW:
                    fft_groups.setdefault(d.geom, []).append(e)
        fft_batched = [e for es in fft_groups.values() for e in es]
        single = [e for e in rest if not any(e is b for b in fft_batched)]
        batched = batched + fft_batched
        for d, _, _ in batched:
            if d.shift_xy is not None:  # calibration shift: the NPred model sees the shifted flux (npred.py:226-230)
                _call("jd_shift_forward", _p(self.flux), _p(d.shift_xy), d.f, d.fH, d.fW, _p(d.flux_s), s)
        for (fH, fW, kh, kw, f, H, W), es in groups.items():
            _call("jd_likelihood_forward", _p(self._table(es, want_grad)), len(es), fH, fW, kh, kw, f, H, W, 1e-25,
                  1.0 / (H * W), s)
        for (fH, fW, kh, kw, f, H, W), es in fft_groups.items():
            _call("jd_likelihood_forward_fft", _p(self._table(es, want_grad, fft=True)), len(es), fH, fW, kh, kw, f, H, W,
                  1e-25, 1.0 / (H * W), s)
        if want_grad:
            for d, _, j in batched:
                if d.train_bkg_norm:  # Adam on log(background norm) with the parameter's own step counter
                    _call("jd_adam_scalar_step_dev", _p(d.bkg_log_norm), _p(d.cal_m), _p(d.cal_v), self._slot(j),
                          _p(d.cal_t), 1, self.lr, self.b1, self.b2, self.eps, s)
            for (fH, fW, kh, kw, f, H, W), es in groups.items():
                _call("jd_likelihood_backward", _p(self._table(es, want_grad)), len(es), fH, fW, kh, kw, f, H, W, s)
            for (fH, fW, kh, kw, f, H, W), es in fft_groups.items():
                _call("jd_likelihood_backward_fft", _p(self._table(es, want_grad, fft=True)), len(es), fH, fW, kh, kw, f,
                      H, W, s)
            for d, _, j in batched:
                if d.shift_xy is not None:
                    self._shift_backward(d, j)
        for d, loss_ptr, j in single:
            self._likelihood_single(d, loss_ptr, j, want_grad)

    @staticmethod
    def _use_batched(es):
        """One launch over the datasets `es` (same geometry) fills the machine"""
        return sum(d.n_tiles for d, _, _ in es) >= LIK_MIN_CTAS

    def _shift_backward(self, d, j):
        s = self._s()
        dshift = self._slot(j) + 8 if d.train_shift else None
        _call("jd_shift_backward", _p(d.dflux_s), _p(self.flux), _p(d.shift_xy), d.f, d.fH, d.fW, _p(self.parts[j]), 0,
              dshift, s)
        if d.train_shift:  # Adam on (shift_x, shift_y): one parameter tensor, one step counter
            _call("jd_adam_scalar_step_dev", _p(d.shift_xy), _p(d.shift_m), _p(d.shift_v), dshift, _p(d.shift_t), 2,
                  self.lr, self.b1, self.b2, self.eps, s)

    def _likelihood_single(self, d, loss_acc, j, want_grad):
        """One dataset through the separate convolution (FFT or direct) / Poisson / adjoint launches."""
        s = self._s()
        src = self.flux
        if d.shift_xy is not None:
            _call("jd_shift_forward", _p(self.flux), _p(d.shift_xy), d.f, d.fH, d.fW, _p(d.flux_s), s)
            src = d.flux_s
        if d.fft is not None:
            _call("jd_conv_forward_fft", _p(src), _p(d.exposure), _p(d.fft.psf_hat), _p(d.fft.workspace),
                  _p(self.conv), d.fH, d.fW, d.kh, d.kw, s)
        else:
            _call("jd_conv_forward_direct", _p(src), _p(d.exposure), _p(d.psf), _p(self.conv), d.fH, d.fW, d.kh,
                  d.kw, s)
        train_cal = want_grad and d.train_bkg_norm
        _call("jd_poisson_forward_backward", _p(self.conv), _p(d.background), _p(d.bkg_log_norm), _p(d.counts), None,
              _p(d.dpool) if want_grad else None, loss_acc, self._slot(j) if train_cal else None, d.H, d.W,
              d.f, d.fW, 1e-25, 1.0 / (d.H * d.W), s)
        if train_cal:
            _call("jd_adam_scalar_step_dev", _p(d.bkg_log_norm), _p(d.cal_m), _p(d.cal_v), self._slot(j),
                  _p(d.cal_t), 1, self.lr, self.b1, self.b2, self.eps, s)
        if not want_grad:
            return
        out = self.parts[j] if d.shift_xy is None else d.dflux_s
        if d.fft is not None:
            _call("jd_conv_backward_fft", _p(d.dpool), _p(d.exposure), _p(d.fft.psf_hat), _p(d.fft.workspace),
                  _p(out), 0, d.fH, d.fW, d.kh, d.kw, d.f, d.H, d.W, s)
        else:
            _call("jd_conv_backward_direct", _p(d.dpool), _p(d.exposure), _p(d.psf), _p(out),
                  0, d.fH, d.fW, d.kh, d.kw, d.f, d.H, d.W, s)
        if d.shift_xy is not None:
            self._shift_backward(d, j)

    def _prior_forward(self, sum_acc, clusters=0):
        if self.P <= 0:
            return
        if clusters and self.backend in (4, 5):  # on part of the SMs (see _gradients)
            bt, binv = ops._bt16(self.packed) if self.backend == 5 else ops._btm(self.packed)
            _call("jd_gmm_prior_forward_tcx2_on", self.backend - 4, int(clusters), _p(self.flux), self.fH, self.fW,
                  _p(self.cur_shift), self.stride, self.rows[0], self.rows[1], _p(bt), _p(binv), _p(self.packed.mw),
                  _p(self.packed.ck), self.packed.K, int(self.packed.upper_tri), int(self.packed.zero_mean),
                  int(self.marginalize), _p(self.sk_ws), _p(self.value), _p(self.argmax), _p(self.logp), sum_acc,
                  self._s())
            return
        if self.backend in (3, 4, 5):
            bt, binv = ops._bt16(self.packed) if self.backend == 5 else ops._btm(self.packed)
            _call(ops.TCM_ENTRY[self.backend], _p(self.flux), self.fH, self.fW, _p(self.cur_shift), self.stride,
                  self.rows[0], self.rows[1], _p(bt), _p(binv), _p(self.packed.mw), _p(self.packed.ck), self.packed.K,
                  int(self.packed.upper_tri), int(self.packed.zero_mean), int(self.marginalize), _p(self.sk_ws),
                  _p(self.value), _p(self.argmax), _p(self.logp), sum_acc, self._s())
            return
        if self.backend == 2:
            bt, binv = ops._bt16(self.packed)
            _call("jd_gmm_prior_forward_tc16", _p(self.flux), self.fH, self.fW, _p(self.cur_shift), self.stride,
                  self.rows[0], self.rows[1], _p(bt), _p(binv), _p(self.packed.mw), _p(self.packed.ck), self.packed.K,
                  int(self.packed.upper_tri), int(self.packed.zero_mean), int(self.marginalize), _p(self.value),
                  _p(self.argmax), _p(self.logp), sum_acc, self._s())
            return
        if self.backend == 1 and self.sk_ws is not None:
            _call("jd_gmm_prior_forward_tc_sk", _p(self.flux), self.fH, self.fW, _p(self.cur_shift), self.stride,
                  self.rows[0], self.rows[1], _p(self.packed.Bt), _p(self.packed.mw), _p(self.packed.ck), self.packed.K,
                  int(self.packed.upper_tri), int(self.packed.zero_mean), int(self.marginalize), _p(self.sk_ws),
                  _p(self.value), _p(self.argmax), _p(self.logp), sum_acc, self._s())
            return
        if self.backend == 1:
            _call("jd_gmm_prior_forward_tc", _p(self.flux), self.fH, self.fW, _p(self.cur_shift), self.stride,
                  self.rows[0], self.rows[1], _p(self.packed.Bt), _p(self.packed.mw), _p(self.packed.ck), self.packed.K,
                  int(self.packed.upper_tri), int(self.packed.zero_mean), int(self.marginalize), _p(self.value), _p(self.argmax), _p(self.logp),
                  sum_acc, self._s())
            return
        _call("jd_gmm_prior_forward", _p(self.flux), self.fH, self.fW, _p(self.cur_shift), self.stride, self.rows[0],
              self.rows[1], _p(self.packed.Lw), _p(self.packed.mw), _p(self.packed.ck), self.packed.K,
              int(self.marginalize), _p(self.value), _p(self.argmax), _p(self.logp), sum_acc, self.backend, self._s())

    def _prior_gradient(self, scale):
        """per-patch gradient rows G (consumed by _adam_fold)"""
        if self.marginalize and self.backend in (1, 2, 3, 4, 5):
            _call("jd_gmm_prior_backward_lse_tc", _p(self.flux), self.fH, self.fW, _p(self.cur_shift), self.stride,
                  self.rows[0], self.rows[1], _p(ops._bt_lam(self.packed)), _p(self.packed.bk), self.packed.K,
                  _p(self.logp), _p(self.value), float(scale), _p(self.G), self._s())
            return
        if self.bwd_ws is None and ops.use_bwd_tri(self.packed, self.P, self.marginalize):
            _call("jd_gmm_prior_backward_max_tri", _p(self.flux), self.fH, self.fW, _p(self.cur_shift), self.stride,
                  self.rows[0], self.rows[1], _p(self.packed.Lw), _p(self.packed.mw), self.packed.K, _p(self.argmax),
                  float(scale), _p(self.G), self._s())
            return
        _call("jd_gmm_prior_backward", _p(self.flux), self.fH, self.fW, _p(self.cur_shift), self.stride, self.rows[0],
              self.rows[1], _p(self.packed.Lam), _p(self.packed.bk), self.packed.K, int(self.marginalize),
              _p(self.argmax), _p(self.logp), _p(self.value), float(scale), _p(self.G), _p(self.bwd_ws), self._s())

    # ------------------------------------------------------------------------------------------
    def _fork(self):
        """Side stream that has waited for everything enqueued so far on the current stream (event fork; inside a CUDA
        graph capture this pulls the side stream into the capture)."""
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        self._side.wait_stream(torch.cuda.current_stream(self.dev))
        return self._side

    def _join(self, side):
        torch.cuda.current_stream(self.dev).wait_stream(side)

    def _gradients(self, entries, prior_scale):
        """Likelihood gradients of `entries` into their parts and the prior's per-patch gradient rows G.  The two
        chains only depend on the flux and write disjoint buffers (dpool, parts, acc[0], acc[2:] | value, argmax, logp,
        G, acc[1]): with `overlap` the likelihood chain runs on a side stream beside the tensor-core prior kernels
        (which leave shared memory and registers for co-resident FP32 CTAs) and joins before the update."""
        has_prior = self.prior is not None and self.P > 0
        # (one dataset: a handful of launches beside a prior kernel that owns every SM is slower than in order; beside
        # a prior kernel on part of the SMs it can be faster - `split_clusters`)
        if self.overlap and has_prior and len(entries) > 1:
            side = self._fork()
            with torch.cuda.stream(side):
                self._likelihoods(entries, want_grad=True)
            self._prior_forward(self.acc.data_ptr() + 8)
            self._prior_gradient(prior_scale)
            self._join(side)
            return has_prior
        if self.overlap and has_prior and self.split_clusters and self.backend in (4, 5):
            side = self._fork()
            self._prior_forward(self.acc.data_ptr() + 8, self.split_clusters)  # first: its CTA pairs claim their SMs
            with torch.cuda.stream(side):
                self._likelihoods(entries, want_grad=True)
            self._prior_gradient(prior_scale)
            self._join(side)
            return has_prior
        self._likelihoods(entries, want_grad=True)
        if has_prior:
            self._prior_forward(self.acc.data_ptr() + 8)
            self._prior_gradient(prior_scale)
        return has_prior

    def _step_body(self, i):
        """Reference step for dataset i: total = L_i - beta * prior / D  (core.py:214-229)."""
        self._begin_flux(advance_adam=1, zero_acc=self.acc)
        has_prior = self._gradients([(self.datasets[i], self.acc.data_ptr(), 0)], -self.c if self.prior else 0.0)
        self._adam_joint(1, has_prior, -self.beta / self.prior_weight)

    def _joint_pre(self):
        """Joint step up to the local gradient: d L_d / d flux of the local datasets in parts[j], the gradient rows G
        of the local patch rows scaled by beta c (loss.py:257-261)."""
        self._begin_flux(advance_adam=1, zero_acc=self.acc)
        entries = [(d, self._loss_slot(j), j) for j, d in enumerate(self.datasets)]
        return self._gradients(entries, self.c * self.beta if self.prior else 0.0)

    def _joint_body(self):
        """Joint step on sum_d L_d - beta * prior.  One GPU: a single update kernel.  Several ranks: local reduce into
        the (symmetric) partial-gradient buffer, then the fused peer-memory reduce + Adam + theta broadcast with both
        cross-rank barriers inside the kernel, or ncclAllReduce + replicated Adam."""
        has_prior = self._joint_pre()
        if self.world == 1:
            self._adam_joint(self.D, has_prior, 1.0)
            return
        self._grad_reduce(self.D, has_prior, 1.0)
        if self.collective == "peer":
            self._peer_update()
        else:
            torch.distributed.all_reduce(self.dflux_l, group=self.pg)
            self._joint_post()

    def _peer_update(self):
        if self.h_grad is None:
            raise _lib.JolidecoB200Error("this engine's peer buffers were released at the end of its run")
        _call("jd_adam_allreduce_peer_sync", self.h_grad.buffer_ptrs_dev, self.h_theta.buffer_ptrs_dev,
              self.h_sig.buffer_ptrs_dev, _p(self.sync_state), self.rank, self.world, _p(self.m), _p(self.v),
              _p(self.flux), _p(self.mask), int(self.use_log_flux), self.n, _p(self.adam_scalars), self.b1, self.b2,
              self.eps, self._s())

    def _joint_local(self):
        has_prior = self._joint_pre()
        self._grad_reduce(self.D, has_prior, 1.0)

    def _joint_post(self):
        _call("jd_adam_step_dev", _p(self.theta), _p(self.m), _p(self.v), _p(self.flux), _p(self.mask), _p(self.dflux_l),
              None, 0.0, int(self.use_log_flux), self.n, _p(self.adam_scalars), self.b1, self.b2, self.eps, self._s())

    def _run(self, key, body):
        if not self.use_graph:
            body()
            return
        g = self._graphs.get(key)
        if g is not None:
            _STATS["launches"] += self._graph_nodes[key]
        if g is None:
            # one eager pass would advance the state; capture directly instead (kernels are not
            # executed during capture) after making sure lazy one-time setup has happened
            # dry pass: builds every lazily created device table (host -> device copies are illegal inside a capture)
            _STATS["dry"] = True
            try:
                body()
            finally:
                _STATS["dry"] = False
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            before = _STATS["launches"]
            # low-level capture: the torch.cuda.graph() context manager also runs gc.collect() and
            # empty_cache() on entry (~100 ms per capture), which would dominate short runs.
            # thread_local: the NCCL watchdog thread may query events while this thread captures.
            if self._capture_stream is None:
                self._capture_stream = torch.cuda.Stream(device=self.dev)
            with torch.cuda.stream(self._capture_stream):
                g.capture_begin(capture_error_mode="thread_local" if self.pg is not None else "global")
                try:
                    body()
                finally:
                    g.capture_end()
            self._graph_nodes[key] = _STATS["launches"] - before
            self._graphs[key] = g
        g.replay()

    def warmup(self, joint=False):
        """Run every kernel once outside graph capture (function attributes, module load, NCCL
        communicator), then restore the optimiser state so that training starts from step 0."""
        tensors = [self.theta, self.m, self.v, self.counters, self.acc]
        for d in self.datasets:
            if d.train_bkg_norm:
                tensors += [d.bkg_log_norm, d.cal_m, d.cal_v, d.cal_t]
            if d.train_shift:
                tensors += [d.shift_xy, d.shift_m, d.shift_v, d.shift_t]
        state = [t.clone() for t in tensors]
        graph = self.use_graph
        self.use_graph = False
        if joint:
            self._joint_body()  # collective: every rank calls warmup(joint=True)
        else:
            for i in range(min(self.D, 1)):
                self._step_body(i)
        self.use_graph = graph
        torch.cuda.synchronize(self.dev)
        self._tune_split(joint)
        if joint and self.world > 1:
            torch.distributed.barrier(group=self.pg)  # no peer is still writing theta slices into this replica
        for t, s in zip(tensors, state):
            t.copy_(s)
        torch.cuda.synchronize(self.dev)
        if joint and self.world > 1:
            torch.distributed.barrier(group=self.pg)

    def _tune_split(self, joint):
        """One-dataset steps: time this engine's gradient pass (local kernels only, scratch outputs, state restored by
        `warmup`) in order and with the prior forward on a fraction of the SM pairs beside the likelihood chain; keep
        the split only if it is clearly faster.  Which split wins depends on the balance of the two chains (P, K, PSF
        size, direct or FFT path), so it is measured rather than modelled: ~2 ms once per run.  Ranks of a joint run
        decide together (MAX of the timings over ranks): one rank left in order would be the step time of all."""
        pairs = (torch.cuda.get_device_properties(self.dev).multi_processor_count // 2) if self.dev.type == "cuda" else 0
        cands = [0] + sorted({max(1, int(round(pairs * f))) for f in (0.25, 0.35, 0.5, 0.65, 0.8)})
        collective = bool(joint and self.world > 1 and self.split_auto and self.dev.type == "cuda")
        single = (joint and self.D == 1) or (not joint and self.D >= 1)
        eligible = bool(self.split_auto and self.overlap and single and self.prior is not None and self.P > 0
                        and self.backend in (4, 5) and self.dev.type == "cuda")
        if not eligible and not collective:
            return
        key = None
        if eligible:
            d0 = self.datasets[0]
            key = (self.dev.index, self.fH, self.fW, self.P, self.packed.K, self.backend, bool(self.marginalize),
                   self.stride, d0.geom, d0.shift_xy is not None, bool(joint), self.world)
            if not collective and key in _SPLIT_CACHE:  # batched independent runs build many engines of one shape
                self.split_clusters = _SPLIT_CACHE[key]
                return
        timings = {c: 0.0 for c in cands}
        if eligible:
            entries = [(d0, self.acc.data_ptr() if not joint else self._loss_slot(0), 0)]
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            graph, self.use_graph = self.use_graph, False
            try:
                for c in cands:
                    self.split_clusters = c
                    self._gradients(entries, self.c)
                    start.record()
                    for _ in range(3):
                        self._gradients(entries, self.c)
                    end.record()
                    end.synchronize()
                    timings[c] = start.elapsed_time(end)
            finally:
                self.use_graph = graph
                self.split_clusters = 0
        if collective:  # every rank of the group takes part, eligible or not (ineligible ranks contribute zeros)
            t = torch.tensor([timings[c] for c in cands], dtype=torch.float64, device=self.dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=self.pg)
            timings = dict(zip(cands, t.tolist()))
        if not eligible:
            return
        best = min(timings, key=timings.get)
        self.split_clusters = best if timings[best] < 0.95 * timings[0] else 0
        self.split_timings_ms = {c: t / 3 for c, t in timings.items()}
        _SPLIT_CACHE[key] = self.split_clusters

    def step(self, i):
        self._last_joint = False
        self._run(("step", i), lambda: self._step_body(i))

    def joint_step(self):
        self._last_joint = True
        if self.world == 1 or self.collective == "peer":
            self._run(("joint",), self._joint_body)  # one graph: the peer kernel carries its own cross-rank barriers
            return
        self._run(("joint-local",), self._joint_local)
        torch.distributed.all_reduce(self.dflux_l, group=self.pg)  # eager NCCL between the two graphs
        self._run(("joint-post",), self._joint_post)

    # ------------------------------------------------------------------------------------------
    def trace_enqueue(self, out_row, refresh_flux=False):
        """Same evaluation as `trace_losses` but asynchronous: the raw accumulators are copied on the
        stream into `out_row` (a device double tensor of n_trace entries); decode with `trace_decode`."""
        reuse = self._trace_reuse(refresh_flux)
        self._run(("trace", bool(refresh_flux), reuse), lambda: self._trace_body(refresh_flux, reuse))
        out_row.copy_(self.acc_trace, non_blocking=True)

    def trace_decode(self, vals):
        """Raw accumulator row(s) (host numpy, already summed over ranks) -> (datasets, prior, validation)."""
        npix = self.counts_shape[0] * self.counts_shape[1]
        ld = [float(vals[j] / npix) for j in range(self.Dg)]
        lp = float(vals[self.Dg] * self.c) if self.prior is not None else 0.0
        lv = [float(vals[self.Dg + 1 + j] / npix) for j in range(self.Vg)]
        return ld, lp, lv

    def _trace_reuse(self, refresh_flux):
        """After a joint step the training datasets' losses of the trace are the ones that step just accumulated: same
        flux (the trace is evaluated at the flux of the step's start), same data - unless a calibration parameter is
        trained (the reference evaluates the trace with the parameters AFTER the step, core.py:229/245)."""
        return bool(self._last_joint and not refresh_flux and self.D > 0
                    and not any(d.train_bkg_norm or d.train_shift for d in self.datasets))

    def _trace_body(self, refresh_flux, reuse=False):
        self._begin(advance_adam=0, zero_acc=self.acc_trace, with_shift=True)
        if refresh_flux:
            self._flux()
        base = self.acc_trace.data_ptr()
        entries = []
        if reuse:
            if self._trace_dst is None:
                self._trace_dst = torch.tensor(self.dataset_index, dtype=torch.int64, device=self.dev)
            if not _STATS["dry"]:
                self.acc_trace.index_copy_(0, self._trace_dst, self.acc[self._loss0:self._loss0 + self.D])
        else:
            entries = [(d, base + 8 * j, None) for j, d in zip(self.dataset_index, self.datasets)]
        entries += [(d, base + 8 * (self.Dg + 1 + j), None) for j, d in zip(self.validation_index, self.datasets_validation)]
        has_prior = self.prior is not None and self.P > 0
        if self.overlap and has_prior and len(entries) > 1:
            side = self._fork()
            with torch.cuda.stream(side):
                self._likelihoods(entries, want_grad=False)
            self._prior_forward(base + 8 * self.Dg)
            self._join(side)
            return
        self._likelihoods(entries, want_grad=False)
        if has_prior:
            self._prior_forward(base + 8 * self.Dg)  # this rank's patch-row block

    def trace_losses(self, refresh_flux=False):
        """Per-epoch trace (loss.py:212-250): every dataset's Poisson loss and one more prior draw,
        evaluated at the flux of the LAST step's start (the reference hands the stale `fluxes` tuple
        to append_trace, core.py:217/245).  One host sync.  Returns (datasets, prior, validation)."""

        reuse = self._trace_reuse(refresh_flux)
        self._run(("trace", bool(refresh_flux), reuse), lambda: self._trace_body(refresh_flux, reuse))
        if self.world > 1:  # every slot is written by exactly one rank (prior: partial sums): one all-reduce
            torch.distributed.all_reduce(self.acc_trace, group=self.pg)
        return self.trace_decode(self.acc_trace.cpu().numpy())

    def last_step_losses(self):
        """(Poisson mean loss, prior value) of the most recent step; syncs."""
        vals = self.acc.clone()
        if self.world > 1:
            torch.distributed.all_reduce(vals, group=self.pg)
        vals = vals.cpu().numpy()
        npix = self.counts_shape[0] * self.counts_shape[1]
        return self.poisson_sum(vals) / npix, float(vals[1] * self.c) if self.prior is not None else 0.0

    def flux_numpy(self):
        self._flux()
        return self.flux.cpu().numpy()
