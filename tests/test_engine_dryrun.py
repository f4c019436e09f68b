"""Dry run of the fused engine's HOST logic on CPU: the C-ABI calls are recorded instead of executed (no kernel runs,
no result is checked here - parity is the GPU tests' job).  Guards the launch sequence of a step / joint step / trace
evaluation and the Python paths around it (calibration routing, accumulator offsets) against host-side regressions."""
import contextlib
import types

import numpy as np
import pytest
import torch

import jolideco_b200 as J
from jolideco_b200 import engine as E
from jolideco_b200 import ops
from jolideco_b200.core import MAPDeconvolver


@pytest.fixture
def recorder(monkeypatch):
    calls = []

    def fake_call(name, *args):
        calls.append((name, args))

    monkeypatch.setattr(E._lib, "call", fake_call)
    monkeypatch.setenv("JD_OVERLAP", "0")  # one stream unless a test asks for the fork / join structure
    monkeypatch.setattr(E, "LIK_MIN_CTAS", 0)  # batched likelihood kernels whatever the image size
    monkeypatch.setattr(ops, "require_device", lambda *a, **k: None)
    monkeypatch.setattr(ops, "_check", lambda t, name, dtype=torch.float32: t)
    monkeypatch.setattr(ops, "use_stream_k", lambda P, device: False)
    monkeypatch.setattr(E.MapEngine, "_s", lambda self: 0)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", lambda *a, **k: contextlib.nullcontext())
    return calls


def fake_packed(K=4, upper_tri=True, zero_mean=False):
    z = torch.zeros
    return types.SimpleNamespace(K=K, D=64, upper_tri=upper_tri, zero_mean=zero_mean, Lw=z(K, 64, 64), mw=z(K, 64),
                                 ck=z(K), Lam=z(K, 64, 64), bk=z(K, 64), Bt=z(K * 32768, dtype=torch.uint8),
                                 _Bt_lam=z(K * 32768, dtype=torch.uint8), _Bt16=None, device=torch.device("cpu"))


def dataset(n=32, k=5, f=1, **kw):
    H = n // f
    return E.DatasetBuffers(torch.ones(H, H), torch.ones(n, n), torch.ones(k, k), torch.ones(H, H), f, **kw)


def names(calls):
    return [c[0] for c in calls]


def test_reference_step_launch_sequence(recorder):
    eng = E.MapEngine(torch.zeros(32, 32), [dataset()], prior=dict(packed=fake_packed(), stride=4, marginalize=False,
                                                                   backend=1), use_graph=False,
                      shift_table=np.zeros((4, 2), dtype=np.int32))
    eng.step(0)
    assert names(recorder) == ["jd_step_begin_flux", "jd_likelihood_forward", "jd_likelihood_backward",
                               "jd_gmm_prior_forward_tc", "jd_gmm_prior_backward_max_tri", "jd_adam_joint_step_dev"]
    assert E._STATS["launches"] >= 6
    fwd, bwd, adam = recorder[1][1], recorder[2][1], recorder[5][1]
    assert fwd[1] == 1 and fwd[2:9] == (32, 32, 5, 5, 1, 32, 32) and bwd[0] == fwd[0]   # one dataset, same table
    assert adam[5] == eng.parts.data_ptr() and adam[6] == 1 and adam[8] == eng.G.data_ptr()
    del recorder[:]
    eng.trace_enqueue(torch.zeros(eng.n_trace, dtype=torch.float64))
    assert names(recorder) == ["jd_step_begin", "jd_likelihood_forward", "jd_gmm_prior_forward_tc"]
    assert recorder[1][1][0] != fwd[0]                                                  # loss-only table: no dpool


def test_logsumexp_and_uniform_prior_sequences(recorder):
    eng = E.MapEngine(torch.zeros(32, 32), [dataset()], prior=dict(packed=fake_packed(), stride=4, marginalize=True,
                                                                   backend=1), use_graph=False)
    eng.step(0)
    assert names(recorder)[-3:] == ["jd_gmm_prior_forward_tc", "jd_gmm_prior_backward_lse_tc", "jd_adam_joint_step_dev"]
    del recorder[:]
    eng = E.MapEngine(torch.zeros(32, 32), [dataset()], prior=None, use_graph=False)
    eng.step(0)
    assert names(recorder) == ["jd_step_begin_flux", "jd_likelihood_forward", "jd_likelihood_backward",
                               "jd_adam_joint_step_dev"]
    assert recorder[3][1][8] is None                                                    # uniform prior: no fold


def test_separate_launches_when_the_batched_kernels_are_switched_off(recorder, monkeypatch):
    monkeypatch.setenv("JD_LIK_BATCHED", "0")
    eng = E.MapEngine(torch.zeros(32, 32), [dataset()], prior=None, use_graph=False)
    eng.step(0)
    assert names(recorder) == ["jd_step_begin_flux", "jd_conv_forward_direct", "jd_poisson_forward_backward",
                               "jd_conv_backward_direct", "jd_adam_joint_step_dev"]


def test_small_launches_keep_the_separate_kernels(recorder, monkeypatch):
    """One 64 x 64 tile per dataset does not fill the SMs: conv / Poisson / conv per dataset; 2 x 80 tiles do."""
    monkeypatch.setattr(E, "LIK_MIN_CTAS", 148)
    eng = E.MapEngine(torch.zeros(32, 32), [dataset(), dataset()], prior=None, use_graph=False)
    eng.joint_step()
    assert names(recorder) == ["jd_step_begin_flux"] + ["jd_conv_forward_direct", "jd_poisson_forward_backward",
                                                        "jd_conv_backward_direct"] * 2 + ["jd_adam_joint_step_dev"]
    del recorder[:]
    big = [dataset(n=576), dataset(n=576)]
    assert big[0].n_tiles == 81
    eng = E.MapEngine(torch.zeros(576, 576), big, prior=None, use_graph=False)
    eng.joint_step()
    assert names(recorder) == ["jd_step_begin_flux", "jd_likelihood_forward", "jd_likelihood_backward",
                               "jd_adam_joint_step_dev"]
    del recorder[:]
    eng.step(0)  # the reference step handles one dataset: 81 CTAs, separate kernels
    assert "jd_conv_forward_direct" in names(recorder)


def test_fft_path_is_taken_for_large_psfs(recorder, monkeypatch):
    """PSFs of >= ops.FFT_MIN_PSF_AREA taps: the shared-memory FFT path, all datasets of a geometry per launch with the
    Poisson statistic fused into the last pass (jd_likelihood_*_fft); JD_FFT_BATCHED=0: one dataset at a time."""
    plan = types.SimpleNamespace(psf_hat=torch.zeros(4), workspace=torch.zeros(4))
    monkeypatch.setattr(ops, "FFTConvPlan", lambda psf, fH, fW: types.SimpleNamespace(psf_hat=torch.zeros(4),
                                                                                       workspace=torch.zeros(4)))
    ds = [dataset(n=64, k=41), dataset(n=64, k=41)]
    eng = E.MapEngine(torch.zeros(64, 64), ds, prior=None, use_graph=False)
    eng.joint_step()
    assert names(recorder) == ["jd_step_begin_flux", "jd_likelihood_forward_fft", "jd_likelihood_backward_fft",
                               "jd_adam_joint_step_dev"]
    fwd, bwd = recorder[1][1], recorder[2][1]
    assert fwd[1] == 2 and fwd[2:9] == (64, 64, 41, 41, 1, 64, 64) and bwd[0] == fwd[0]
    table = eng._table([(d, eng.acc.data_ptr(), j) for j, d in enumerate(ds)], True, fft=True)
    rec = table.numpy().view(ops.FFTLIK_DTYPE)
    assert rec.shape == (2,) and rec.itemsize == 112
    assert [int(r["workspace"]) for r in rec] == [d.fft.workspace.data_ptr() for d in ds]
    assert [int(r["psf_hat"]) for r in rec] == [d.fft.psf_hat.data_ptr() for d in ds]
    assert int(rec[1]["dflux"]) == eng.parts.data_ptr() + 4 * eng.n
    del recorder[:]
    monkeypatch.setattr(E, "FFT_BATCHED", False)
    eng.step(0)
    assert "jd_conv_forward_fft" in names(recorder) and "jd_conv_backward_fft" in names(recorder)
    assert "jd_conv_forward_direct" not in names(recorder) and "jd_likelihood_forward" not in names(recorder)
    assert plan is not None


def test_calibration_accumulators_and_shift_sequence(recorder):
    logb = torch.zeros(1)
    shift = torch.tensor([0.4, -0.7])
    ds = [dataset(bkg_log_norm=logb, train_bkg_norm=True, shift_xy=shift, train_shift=True),
          dataset(bkg_log_norm=torch.zeros(1), train_bkg_norm=True)]
    eng = E.MapEngine(torch.zeros(32, 32), ds, prior=None, use_graph=False)
    assert ds[0].flux_s is not None and ds[1].flux_s is None and eng.acc.numel() == 2 + 3 * 2 + 2
    eng.step(0)
    seq = names(recorder)
    assert seq == ["jd_step_begin_flux", "jd_shift_forward", "jd_likelihood_forward", "jd_adam_scalar_step_dev",
                   "jd_likelihood_backward", "jd_shift_backward", "jd_adam_scalar_step_dev", "jd_adam_joint_step_dev"]
    base = eng.acc.data_ptr()
    rec = np.frombuffer(eng._tables[next(iter(eng._tables))].numpy().tobytes(), dtype=E.LIK_DTYPE)[0]
    assert rec["loss_sum"] == base and rec["dlogb"] == base + 16       # loss sum -> acc[0], dlogb -> slot of dataset 0
    assert rec["flux"] == ds[0].flux_s.data_ptr() and rec["dflux"] == ds[0].dflux_s.data_ptr()  # shifted in / out
    assert recorder[3][1][3] == base + 16                               # Adam on log(background norm) reads the slot
    shift_b = recorder[5][1]
    assert shift_b[0] == ds[0].dflux_s.data_ptr() and shift_b[6] == eng.parts.data_ptr() and shift_b[7] == 0
    assert shift_b[8] == base + 24                                      # dshift -> the two slots after dlogb
    assert recorder[6][1][5] == 2                                       # Adam on the (shift_x, shift_y) pair
    # joint step: both datasets in one launch per direction, each with its own accumulator slots and gradient part
    del recorder[:]
    eng.joint_step()
    seq = names(recorder)
    assert seq == ["jd_step_begin_flux", "jd_shift_forward", "jd_likelihood_forward", "jd_adam_scalar_step_dev",
                   "jd_adam_scalar_step_dev", "jd_likelihood_backward", "jd_shift_backward", "jd_adam_scalar_step_dev",
                   "jd_adam_joint_step_dev"]
    assert recorder[2][1][1] == 2 and recorder[-1][1][6] == 2
    key = [k for k in eng._tables if len(k[0]) == 2][0]
    recs = np.frombuffer(eng._tables[key].numpy().tobytes(), dtype=E.LIK_DTYPE)
    assert recs[1]["dlogb"] == base + 8 * (2 + 3) and recs[1]["dflux"] == eng.parts.data_ptr() + 4 * 32 * 32
    assert recs[1]["flux"] == eng.flux.data_ptr() and recs[0]["loss_const"] == 0.0   # counts of 1: no Stirling term


def test_calibration_routing(monkeypatch):
    cals = J.NPredCalibrations()
    cals["a"] = J.NPredCalibration(background_norm=1.2)
    assert MAPDeconvolver._calibrations_fusable(cals)            # shifts at 0: background norm only
    cals["b"] = J.NPredCalibration(shift_x=0.3, shift_y=0.0)
    monkeypatch.delenv("JD_FUSED_SHIFT", raising=False)
    assert MAPDeconvolver._calibrations_fusable(cals)            # non-zero shift: jd_shift_forward / backward
    monkeypatch.setenv("JD_FUSED_SHIFT", "0")
    assert not MAPDeconvolver._calibrations_fusable(cals)        # ... unless sent to the autograd path
    monkeypatch.delenv("JD_FUSED_SHIFT")
    cals["c"] = J.NPredCalibration(psf_scale=1.1)
    assert not MAPDeconvolver._calibrations_fusable(cals)        # psf rescaling is not covered
    assert MAPDeconvolver._shift_is_zero(cals["a"]) and not MAPDeconvolver._shift_is_zero(cals["b"])


def test_deconvolver_run_builds_the_engine_with_calibrations(recorder, monkeypatch):
    """`MAPDeconvolver.run` end to end on CPU with the kernels recorded: datasets -> TotalLoss -> DatasetBuffers (with
    the calibration's parameter storage) -> engine steps and per-epoch traces -> result object."""
    monkeypatch.setenv("JD_FUSED_SHIFT", "1")
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    rng = np.random.default_rng(0)
    A = rng.normal(0, 0.05, size=(3, 64, 64))
    gmm = J.GaussianMixtureModel.from_numpy(np.zeros((3, 64)), A @ A.transpose(0, 2, 1) + 0.01 * np.eye(64),
                                            np.full(3, 1 / 3), meta=J.GaussianMixtureModelMeta(stride=4))
    prior = J.GMMPatchPrior(gmm=gmm, generator=torch.Generator().manual_seed(0))
    comp = J.SpatialFluxComponent.from_numpy(flux=np.ones((24, 24)), upsampling_factor=2, prior=prior)
    ds = {n: dict(counts=rng.poisson(2.0, size=(24, 24)).astype(np.float32), psf=np.full((4, 4), 1 / 16.0),
                  exposure=np.ones((24, 24)), background=np.full((24, 24), 0.5)) for n in ("a", "b")}
    cals = J.NPredCalibrations()
    cals["a"] = J.NPredCalibration(shift_x=0.4, shift_y=-0.2, background_norm=1.1)
    cals["b"] = J.NPredCalibration(background_norm=0.9, frozen=True)
    deco = MAPDeconvolver(n_epochs=2, display_progress=False, use_cuda_graph=False, compute_error=True)
    deco.device = torch.device("cpu")  # dry run: no kernel is executed
    res = deco.run(datasets=ds, components=comp, calibrations=cals)
    eng = deco.engine
    a, b = eng.datasets
    assert a.f == 2 and a.shift_xy is not None and a.train_shift and a.train_bkg_norm
    assert a.shift_xy.data_ptr() == cals["a"].shift_xy.data_ptr()      # the engine updates the parameter in place
    assert b.shift_xy is None and not b.train_bkg_norm and b.bkg_log_norm is not None
    seq = names(recorder)
    assert seq.count("jd_shift_forward") == 2 * (1 + 1) + 1            # warm-up + 2 steps + 2 traces for dataset a
    assert seq.count("jd_adam_joint_step_dev") == 2 * 2 + 1           # 2 epochs x 2 datasets + warm-up
    assert len(res.trace_loss) == 2 and set(res.trace_loss.colnames) >= {"total", "dataset-a", "dataset-b"}
    assert res.flux_upsampled_total.shape == (48, 48)
    err = res.components["flux"].flux_upsampled_error_numpy   # compute_error: inf everywhere, as in the reference
    assert err.shape == (48, 48) and np.all(np.isinf(err))


def test_engine_support_matrix():
    """Which configurations take the fused engine and which the autograd path (core.py `_engine_supported`)."""
    rng = np.random.default_rng(1)
    A = rng.normal(0, 0.05, size=(2, 64, 64))
    gmm = J.GaussianMixtureModel.from_numpy(np.zeros((2, 64)), A @ A.transpose(0, 2, 1) + 0.01 * np.eye(64),
                                            np.full(2, 0.5), meta=J.GaussianMixtureModelMeta(stride=4))

    def comps(prior, **kw):
        c = J.FluxComponents()
        c["flux"] = J.SpatialFluxComponent.from_numpy(flux=np.ones((16, 16)), prior=prior, **kw)
        return c

    deco = MAPDeconvolver(display_progress=False)
    assert deco._engine_supported(comps(J.UniformPrior()), None)
    assert deco._engine_supported(comps(J.GMMPatchPrior(gmm=gmm)), None)
    assert deco._engine_supported(comps(J.GMMPatchPrior(gmm=gmm, norm=J.IdentityImageNorm())), None)
    assert not deco._engine_supported(comps(J.GMMPatchPrior(gmm=gmm, norm=J.ASinhImageNorm(alpha=0.5))), None)
    assert not deco._engine_supported(comps(J.UniformPrior(), frozen=True), None)
    assert not MAPDeconvolver(optimizer_type="sgd")._engine_supported(comps(J.UniformPrior()), None)
    assert not MAPDeconvolver(fused=False)._engine_supported(comps(J.UniformPrior()), None)
    two = comps(J.UniformPrior())
    two["more"] = J.SpatialFluxComponent.from_numpy(flux=np.ones((16, 16)), prior=J.UniformPrior())
    assert not deco._engine_supported(two, None)
    # a trainable norm's parameters reach the optimiser through the component (models/core.py: components.parameters())
    with_norm = comps(J.GMMPatchPrior(gmm=gmm, norm=J.ASinhImageNorm(alpha=0.5)))
    assert len(list(with_norm.parameters())) == 3


@pytest.mark.parametrize("backend,entry", [(4, "jd_gmm_prior_forward_tcm2"), (5, "jd_gmm_prior_forward_tc16x2"),
                                           (3, "jd_gmm_prior_forward_tcm")])
def test_two_tile_prior_kernels_and_bucketed_backward_are_wired(recorder, monkeypatch, backend, entry):
    """Backends 4 / 5 (jd_gmm_tcm2.cu) take the operand image of their recipe and a workspace of their own size; from
    ops.BWD_BUCKETED_MIN_PATCHES patches on the max-mode backward goes through the bucketed kernel (workspace given)."""
    sizes = []
    lib = types.SimpleNamespace(
        jd_gmm_tcm_workspace_bytes=lambda P, K: sizes.append(("tcm", P, K)) or 512,
        jd_gmm_tcm2_workspace_bytes=lambda P, K: sizes.append(("tcm2", P, K)) or 1024,
        jd_gmm_backward_workspace_elems=lambda P, K: sizes.append(("bwd", P, K)) or 4 * K + 2 + P,
        jd_likelihood_supported=lambda kh, kw, f: 1)
    monkeypatch.setattr(E._lib, "load", lambda: lib)
    monkeypatch.setattr(ops._lib, "load", lambda: lib)
    packed = fake_packed()
    images = {"tcm": (torch.zeros(8, dtype=torch.uint8), torch.zeros(4)), "tc16": (torch.zeros(4, dtype=torch.uint8), torch.zeros(4))}
    monkeypatch.setattr(ops, "_btm", lambda p: images["tcm"])
    monkeypatch.setattr(ops, "_bt16", lambda p: images["tc16"])
    prior = dict(packed=packed, stride=4, marginalize=False, backend=backend)
    monkeypatch.setattr(ops, "BWD_BUCKETED_MIN_PATCHES", 10)
    eng = E.MapEngine(torch.zeros(32, 32), [dataset()], prior=prior, use_graph=False)
    assert eng.P == 49 and eng.bwd_ws is not None and eng.bwd_ws.numel() == 4 * 4 + 2 + 49 and not eng.bwd_ws.any()
    assert sizes[0] == (("tcm2" if backend in (4, 5) else "tcm"), 49, 4)
    assert eng.sk_ws.numel() == (1024 if backend in (4, 5) else 512) and not eng.sk_ws.any()
    eng.step(0)
    assert names(recorder) == ["jd_step_begin_flux", "jd_likelihood_forward", "jd_likelihood_backward", entry,
                               "jd_gmm_prior_backward", "jd_adam_joint_step_dev"]
    fwd, bwd = recorder[3][1], recorder[4][1]
    want = images["tc16"] if backend == 5 else images["tcm"]
    assert fwd[7] == want[0].data_ptr() and fwd[8] == want[1].data_ptr()   # operand image, inverse component scales
    assert fwd[15] == eng.sk_ws.data_ptr() and bwd[-2] == eng.bwd_ws.data_ptr()
    # fewer patches than the threshold: the triangular warp-per-patch kernel, no workspace
    monkeypatch.setattr(ops, "BWD_BUCKETED_MIN_PATCHES", 1000)
    eng = E.MapEngine(torch.zeros(32, 32), [dataset()], prior=prior, use_graph=False)
    assert eng.bwd_ws is None
    del recorder[:]
    eng.step(0)
    assert names(recorder)[4] == "jd_gmm_prior_backward_max_tri"
    # logsumexp mode never buckets
    assert not ops.use_bwd_bucketed(10 ** 6, marginalize=True)


def test_overlap_forks_the_likelihood_chain_and_joins_before_the_update(recorder, monkeypatch):
    """JD_OVERLAP / overlap=True: likelihood kernels are enqueued on the side stream between a fork and a join, the
    prior chain on the main stream, and everything that reads both (Adam, fold into dflux_l) after the join."""
    events = recorder  # same list: interleave stream events with the recorded C-ABI calls

    class FakeStream:
        def __init__(self, name="side", **kw):
            self.name = name

        def wait_stream(self, other):
            events.append((f"{self.name}.wait({other.name})", ()))

    main = FakeStream("main")

    @contextlib.contextmanager
    def fake_stream_ctx(stream):
        events.append((f"enter({stream.name})", ()))
        yield
        events.append((f"exit({stream.name})", ()))

    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", fake_stream_ctx)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: main)
    prior = dict(packed=fake_packed(), stride=4, marginalize=False, backend=1)
    eng = E.MapEngine(torch.zeros(32, 32), [dataset(), dataset()], prior=prior, use_graph=False, overlap=True)
    eng.step(0)  # one dataset: a handful of launches, kept in order on the main stream
    assert names(events) == ["jd_step_begin_flux", "jd_likelihood_forward", "jd_likelihood_backward",
                             "jd_gmm_prior_forward_tc", "jd_gmm_prior_backward_max_tri", "jd_adam_joint_step_dev"]
    del events[:]
    eng.joint_step()
    seq = names(events)
    assert seq[:3] == ["jd_step_begin_flux", "side.wait(main)", "enter(side)"]
    assert seq.index("exit(side)") < seq.index("jd_gmm_prior_forward_tc") < seq.index("main.wait(side)")
    assert seq[seq.index("main.wait(side)") + 1:] == ["jd_adam_joint_step_dev"]  # parts + fold + Adam in one launch
    assert seq.count("jd_likelihood_backward") == 1
    del events[:]
    # trace after a joint step: the datasets' losses are the ones the step accumulated (copied on the device), only the
    # prior is evaluated again (with the trace's own cycle-spin draw)
    eng.trace_enqueue(torch.zeros(eng.n_trace, dtype=torch.float64))
    assert names(events) == ["jd_step_begin", "jd_gmm_prior_forward_tc"]
    del events[:]
    eng.step(0)  # ... after a reference step every dataset is evaluated: likelihood chain beside the prior
    del events[:]
    eng.trace_enqueue(torch.zeros(eng.n_trace, dtype=torch.float64))
    seq = names(events)
    assert seq[0] == "jd_step_begin" and seq[-2:] == ["jd_gmm_prior_forward_tc", "main.wait(side)"]
    assert "jd_likelihood_forward" in seq
    # JD_OVERLAP=0 (the fixture's setting): no side stream at all
    assert E.MapEngine(torch.zeros(32, 32), [dataset()], prior=prior, use_graph=False)._side is None


def test_one_dataset_step_puts_the_prior_forward_first_on_part_of_the_sm_pairs(recorder, monkeypatch):
    """JD_SPLIT_CLUSTERS=n (or the value tuned at warm-up): the two-tile prior forward is enqueued first through
    jd_gmm_prior_forward_tcx2_on(recipe, n, ...), then the likelihood chain on the side stream (forked BEFORE the prior
    forward, so it does not wait for it), the prior backward on the main stream, join, update.  0 / other backends /
    several datasets per step: unchanged."""
    events = recorder

    class FakeStream:
        def __init__(self, name="side", **kw):
            self.name = name

        def wait_stream(self, other):
            events.append((f"{self.name}.wait({other.name})", ()))

    main = FakeStream("main")

    @contextlib.contextmanager
    def fake_stream_ctx(stream):
        events.append((f"enter({stream.name})", ()))
        yield
        events.append((f"exit({stream.name})", ()))

    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "stream", fake_stream_ctx)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: main)
    lib = types.SimpleNamespace(jd_gmm_tcm_workspace_bytes=lambda P, K: 512, jd_gmm_tcm2_workspace_bytes=lambda P, K: 1024,
                                jd_gmm_backward_workspace_elems=lambda P, K: 4 * K + 2 + P,
                                jd_likelihood_supported=lambda kh, kw, f: 1)
    monkeypatch.setattr(E._lib, "load", lambda: lib)
    monkeypatch.setattr(ops._lib, "load", lambda: lib)
    images = (torch.zeros(8, dtype=torch.uint8), torch.zeros(4))
    monkeypatch.setattr(ops, "_btm", lambda p: images)
    monkeypatch.setattr(ops, "_bt16", lambda p: images)
    monkeypatch.setenv("JD_SPLIT_CLUSTERS", "20")
    for backend in (4, 5):
        prior = dict(packed=fake_packed(), stride=4, marginalize=False, backend=backend)
        eng = E.MapEngine(torch.zeros(32, 32), [dataset(), dataset()], prior=prior, use_graph=False, overlap=True)
        assert eng.split_clusters == 20 and not eng.split_auto
        del events[:]
        eng.step(1)
        seq = names(events)
        assert seq == ["jd_step_begin_flux", "side.wait(main)", "jd_gmm_prior_forward_tcx2_on", "enter(side)",
                       "jd_likelihood_forward", "jd_likelihood_backward", "exit(side)", "jd_gmm_prior_backward_max_tri",
                       "main.wait(side)", "jd_adam_joint_step_dev"]
        call = [a for n, a in events if n == "jd_gmm_prior_forward_tcx2_on"][0]
        assert call[0] == backend - 4 and call[1] == 20
        del events[:]
        eng.joint_step()  # two datasets per step: the multi-dataset overlap, prior forward on every SM pair
        assert "jd_gmm_prior_forward_tcx2_on" not in names(events)
        assert ops.TCM_ENTRY[backend] in names(events)
    # the one-tile kernels have no such entry: in order
    prior = dict(packed=fake_packed(), stride=4, marginalize=False, backend=3)
    eng = E.MapEngine(torch.zeros(32, 32), [dataset()], prior=prior, use_graph=False, overlap=True)
    del events[:]
    eng.step(0)
    assert names(events)[:4] == ["jd_step_begin_flux", "jd_likelihood_forward", "jd_likelihood_backward",
                                 "jd_gmm_prior_forward_tcm"]
    # default: "auto" = in order until warmup() has timed the candidates (CUDA only; nothing to tune on the host)
    monkeypatch.delenv("JD_SPLIT_CLUSTERS")
    prior = dict(packed=fake_packed(), stride=4, marginalize=False, backend=4)
    eng = E.MapEngine(torch.zeros(32, 32), [dataset()], prior=prior, use_graph=False, overlap=True)
    assert eng.split_auto and eng.split_clusters == 0
    eng._tune_split(joint=False)
    assert eng.split_clusters == 0


def test_peer_buffers_are_pooled_across_engines(monkeypatch):
    """collective='peer': allocation + rendezvous of the symmetric-memory buffers happens once per (device, group,
    pixel count); `release_peer` (end of MAPDeconvolver's joint run, same program point on every rank) syncs theta back
    into the component's storage and returns the set, the next engine takes it over with zeroed flags / gradient and
    its own theta; an engine that starts while another still holds the set allocates a second one."""
    import torch.distributed._symmetric_memory as symm

    made = []

    class Handle:
        buffer_ptrs_dev = 0

        def __init__(self):
            self.barriers = 0

        def barrier(self, channel=0):
            self.barriers += 1

    def empty(n, dtype=None, device=None):
        made.append(("empty", n))
        return torch.full((n,), 7, dtype=dtype)

    def rendezvous(tensor, group):
        made.append(("rendezvous", tensor.numel()))
        return Handle()

    monkeypatch.setattr(symm, "empty", empty)
    monkeypatch.setattr(symm, "rendezvous", rendezvous)
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(E, "_PEER_POOL", {})
    group = types.SimpleNamespace(group_name="g0")

    def engine(value):
        eng = object.__new__(E.MapEngine)
        eng.fH, eng.fW, eng.n, eng.world, eng.pg, eng.dev = 4, 8, 32, 2, group, torch.device("cpu")
        eng.theta = torch.full((4, 8), float(value))
        eng._theta_param, eng._graphs, eng._graph_nodes = None, {"joint": object()}, {"joint": 3}
        eng._enable_peer()
        return eng

    a = engine(1.0)
    assert [m[0] for m in made] == ["empty"] * 3 + ["rendezvous"] * 3
    assert a.theta.data_ptr() == a.sym_theta.data_ptr() and float(a.theta[0, 0]) == 1.0
    assert not a.sym_grad.any() and not a.sym_sig.any() and a.h_sig.barriers == 1
    param = a._theta_param
    a.theta += 2.0  # training moves the symmetric copy only
    a.sym_sig += 5  # barrier epochs of the run
    assert float(param[0, 0]) == 1.0
    first = a.sym_theta
    a.release_peer()
    assert a.theta is param and float(param[0, 0]) == 3.0 and a._graphs == {} and a.h_grad is None
    with pytest.raises(E._lib.JolidecoB200Error, match="released"):
        a._peer_update()
    a.release_peer()  # idempotent
    b = engine(10.0)
    assert len(made) == 6 and b.sym_theta is first  # taken over: no allocation, no rendezvous
    assert float(b.theta[0, 0]) == 10.0 and not b.sym_sig.any() and not b.sym_grad.any() and b.h_sig.barriers == 2
    c = engine(20.0)  # b still holds the set
    assert len(made) == 12 and c.sym_theta is not first
    b.release_peer()
    c.release_peer()
    assert len(E._PEER_POOL[(None, "g0", 32)]) == 2
    # another pixel count never takes a set of the wrong size
    d = object.__new__(E.MapEngine)
    d.fH, d.fW, d.n, d.world, d.pg, d.dev = 4, 4, 16, 2, group, torch.device("cpu")
    d.theta, d._theta_param, d._graphs, d._graph_nodes = torch.zeros(4, 4), None, {}, {}
    d._enable_peer()
    assert len(made) == 18 and d.sym_grad.numel() == 16


def test_split_tuner_decision_rules(recorder, monkeypatch):
    """MapEngine._tune_split: candidates 0 / 25 / 35 / 50 / 65 / 80 % of the SM pairs, each timed over three gradient
    passes after one untimed pass; the fastest split is kept only if it beats the in-order pass by more than 5 %; the
    decision is cached per engine shape; ranks of a joint run use the MAX of the timings over ranks and every rank takes
    part in that all-reduce, eligible or not."""
    lib = types.SimpleNamespace(jd_gmm_tcm_workspace_bytes=lambda P, K: 512, jd_gmm_tcm2_workspace_bytes=lambda P, K: 1024,
                                jd_gmm_backward_workspace_elems=lambda P, K: 4 * K + 2 + P,
                                jd_likelihood_supported=lambda kh, kw, f: 1)
    monkeypatch.setattr(E._lib, "load", lambda: lib)
    monkeypatch.setattr(ops._lib, "load", lambda: lib)
    images = (torch.zeros(8, dtype=torch.uint8), torch.zeros(4))
    monkeypatch.setattr(ops, "_btm", lambda p: images)
    monkeypatch.setattr(E, "_SPLIT_CACHE", {})
    passes = []

    class FakeDev:
        type, index = "cuda", 0

    class FakeEvent:
        table = {}
        current = [0]

        def __init__(self, enable_timing=True):
            pass

        def record(self):
            pass

        def synchronize(self):
            pass

        def elapsed_time(self, other):
            return FakeEvent.table[FakeEvent.current[0]]

    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda dev: types.SimpleNamespace(multi_processor_count=148))

    def build(n_datasets=1):
        prior = dict(packed=fake_packed(), stride=4, marginalize=False, backend=4)
        eng = E.MapEngine(torch.zeros(32, 32), [dataset() for _ in range(n_datasets)], prior=prior, use_graph=False,
                          overlap=False)
        eng.overlap, eng.dev = True, FakeDev()
        real = eng._gradients

        def gradients(entries, scale):
            FakeEvent.current[0] = eng.split_clusters
            passes.append(eng.split_clusters)

        eng._gradients = gradients
        return eng

    cands = [0, 18, 26, 37, 48, 59]
    # clearly faster split
    FakeEvent.table = {0: 3.0, 18: 4.0, 26: 3.3, 37: 2.4, 48: 2.7, 59: 2.9}
    eng = build()
    eng._tune_split(joint=False)
    assert passes == [c for c in cands for _ in range(4)]  # one untimed + three timed passes per candidate
    assert eng.split_clusters == 37 and eng.split_timings_ms[37] == pytest.approx(0.8)
    # same shape again: cached, nothing is timed
    del passes[:]
    eng = build()
    eng._tune_split(joint=False)
    assert passes == [] and eng.split_clusters == 37
    # less than 5 % faster: stay in order
    monkeypatch.setattr(E, "_SPLIT_CACHE", {})
    FakeEvent.table = {0: 3.0, 18: 4.0, 26: 3.3, 37: 2.9, 48: 2.95, 59: 3.1}
    eng = build()
    eng._tune_split(joint=False)
    assert eng.split_clusters == 0
    # several datasets per joint step: not a one-dataset step, nothing to tune
    del passes[:]
    eng = build(2)
    eng._tune_split(joint=True)
    assert passes == [] and eng.split_clusters == 0
    # joint run on several ranks: MAX over ranks decides (here another rank is slow at 37 pairs), and an ineligible rank
    # (two datasets) still takes part in the all-reduce
    monkeypatch.setattr(E, "_SPLIT_CACHE", {})
    FakeEvent.table = {0: 3.0, 18: 4.0, 26: 3.3, 37: 2.4, 48: 2.7, 59: 2.9}
    other = torch.tensor([3.0, 4.0, 3.3, 3.5, 2.6, 2.9], dtype=torch.float64)
    reduced = []

    def all_reduce(t, op=None, group=None):
        reduced.append(t.clone())
        t.copy_(torch.maximum(t, other))

    monkeypatch.setattr(torch.distributed, "all_reduce", all_reduce)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda data, **kw: real_tensor(data, **{k: v for k, v in kw.items() if k != "device"}))
    eng = build()
    eng.world, eng.pg = 2, object()
    eng._tune_split(joint=True)
    assert eng.split_clusters == 48 and reduced[0].tolist() == [3.0, 4.0, 3.3, 2.4, 2.7, 2.9]
    eng = build(2)
    eng.world, eng.pg = 2, object()
    eng._tune_split(joint=True)
    assert len(reduced) == 2 and not reduced[1].any() and eng.split_clusters == 0
