"""GPU tests of the tcgen05 tensor-core engine (K1b): FP16 operands, FP32 accumulation and cell state.

This is the REDUCED-PRECISION path of north_star ("the reduced-precision tensor-core path reports its RMSE
delta"); the FP32 engines carry the 1e-5 parity bar (test_gpu_forward.py).  Tolerances here are stated
against the float64 oracle on the same weights and inputs:
    max |y_tc - y_oracle| <= 2e-3 * max|y_oracle| + 2e-4         (3-factor cells, ranks <= 256)
and the RMSE delta vs the FP32 engine must stay below 5e-4 of the output scale.  FP16 has an 11-bit
significand, so one rounding of h / t per step is ~5e-4 relative; tanh.approx adds ~5e-4 absolute.
"""
import numpy as np
import pytest
import torch

import svdlstm
from helpers import oracle_twin

import os

pytestmark = pytest.mark.gpu
GOLDEN_W = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dropbear_weights.npz")


def _models(H, L, seed=0):
    layers, dense = svdlstm.synthetic_layers(16, H, L, seed=seed)
    full = svdlstm.full_model_from_weights(layers, dense)
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    return full, sm


def _check(y_tc, y_ref, what, rel=2e-3, abs_=2e-4):
    y_tc = np.asarray(y_tc, np.float64)
    y_ref = np.asarray(y_ref, np.float64)
    assert y_tc.shape == y_ref.shape, (what, y_tc.shape, y_ref.shape)
    assert np.isfinite(y_tc).all(), what
    err = np.abs(y_tc - y_ref).max()
    tol = rel * np.abs(y_ref).max() + abs_
    assert err <= tol, "%s: max abs err %.3e > tol %.3e" % (what, err, tol)
    return err


@pytest.mark.parametrize("H,L,rank", [(128, 1, 16), (128, 2, 8), (256, 2, 32), (256, 2, 24), (128, 3, 40)])
def test_tc_resident_matches_oracle(oracle, H, L, rank):
    """Weight stream resident in shared memory; ragged batch (not a multiple of the 32-sequence tile)."""
    _, sm = _models(H, L)
    m = svdlstm.truncate_singular_model(sm, rank)
    x = np.random.default_rng(1).standard_normal((45, 20, 16)).astype(np.float32)
    y = m.predict(x, engine="tc")
    assert m.last_engine() == svdlstm.ENGINE_TC
    _check(y, oracle_twin(oracle, m).predict(x), "tc resident H=%d L=%d r=%d" % (H, L, rank))


@pytest.mark.parametrize("rank", [64, 128, 200, 256])
def test_tc_streamed_matches_oracle(oracle, rank):
    """Ranks whose factors exceed one SM's shared memory are re-streamed from L2 every step (ring of 16 KB slots)."""
    _, sm = _models(256, 2)
    m = svdlstm.truncate_singular_model(sm, rank)
    x = np.random.default_rng(2).standard_normal((33, 12, 16)).astype(np.float32)
    y = m.predict(x, engine="tc")
    _check(y, oracle_twin(oracle, m).predict(x), "tc streamed r=%d" % rank)


def test_tc_reduced_form_and_rmse_delta(oracle):
    """2-factor (ReducedLSTMCell) weights on the tensor-core engine, and the RMSE delta vs the FP32 engine that
    bench.py reports.  C = inv(V1) V2 has entries in the hundreds -- FP16 is the wrong container for it (5 % errors in round 1)
    -- so the engine runs the block as its exact 3-factor equivalent in . (B P s) . Q^T with [I | C] = P s Q^T (K2 on
    device): the 2-factor form now meets the SAME bar as the 3-factor one (2e-3 relative + 2e-4)."""
    _, sm = _models(128, 2)
    x = torch.randn(64, 40, 16, generator=torch.Generator().manual_seed(3)).cuda()
    m3 = svdlstm.truncate_singular_model(sm, 32)
    y32 = m3(x, engine="general")
    ytc = m3(x, engine="tc")
    scale = float(y32.abs().max())
    rmse_delta = float(((ytc - y32) ** 2).mean().sqrt())
    assert rmse_delta < 5e-4 * max(scale, 1.0), rmse_delta
    m2 = svdlstm.make_LSTM_reduced_model(sm, rank=32)
    y2 = m2.predict(x.cpu().numpy(), engine="tc")
    _check(y2, oracle_twin(oracle, m2).predict(x.cpu().numpy()), "tc 2-factor r=32")
    # the shipped DROPBEAR model's 2-factor form (cond(V1) up to 136, max|C| ~ 650) at every rank, units padded 15 -> 128
    layers, dense = svdlstm.load_model_weights_npz(GOLDEN_W)
    dsm = svdlstm.make_LSTM_singular_model(svdlstm.full_model_from_weights(layers, dense), merged_kernel=True, return_sequences=True)
    xd = np.random.default_rng(4).standard_normal((40, 60, 16)).astype(np.float32)
    # (trained weights: |W| up to ~3 and saturating gates through 3 layers -- FP16 operand rounding shows as ~2.5e-3 of the output
    #  scale on this model in every form, full / 3-factor / 2-factor alike; bar 5e-3, RMSE delta reported by bench.py)
    for r in (15, 8, 3):
        md = svdlstm.make_LSTM_reduced_model(dsm, rank=r)
        _check(md.predict(xd, engine="tc"), oracle_twin(oracle, md).predict(xd), "tc DROPBEAR 2-factor r=%d" % r, rel=5e-3)
        m3 = svdlstm.truncate_singular_model(dsm, r)
        e2 = np.abs(md.predict(xd, engine="tc") - oracle_twin(oracle, md).predict(xd)).max()
        e3 = np.abs(m3.predict(xd, engine="tc") - oracle_twin(oracle, m3).predict(xd)).max()
        assert e2 < 3.0 * e3 + 1e-3, "2-factor form is not at 3-factor accuracy: %.3e vs %.3e" % (e2, e3)


def test_tc_batch_properties_full_tile_count():
    """Size-independent properties at a batch spanning many CTAs: per-sequence results do not depend on which
    tile / column a sequence lands in (bit-exact), and padding columns never leak."""
    _, sm = _models(128, 2)
    m = svdlstm.truncate_singular_model(sm, 16)
    x = torch.randn(300, 33, 16, generator=torch.Generator().manual_seed(4)).cuda()
    y = m(x, engine="tc")
    perm = torch.randperm(300, generator=torch.Generator().manual_seed(5)).cuda()
    assert torch.equal(m(x[perm], engine="tc"), y[perm])
    assert torch.equal(m(x[:7], engine="tc"), y[:7])
    assert torch.equal(m(x, engine="tc"), y)   # deterministic across launches


def test_tc_without_dense_top_returns_hidden_sequence():
    """A bare SingularLSTM layer (no Dense top to fuse): the engine returns the FP32 hidden sequence."""
    _, sm = _models(128, 1, seed=7)
    layer = svdlstm.truncate_singular_model(sm, 16).layers[0]
    x = torch.randn(40, 10, 16, generator=torch.Generator().manual_seed(8)).cuda()
    h32 = layer.call(x, engine="general")
    htc = layer.call(x, engine="tc")
    assert tuple(htc.shape) == (40, 10, 128)
    assert float((htc - h32).abs().max()) < 3e-3


@pytest.mark.parametrize("H,L", [(128, 2), (256, 2)])
def test_tc_split_forms_match_oracle(oracle, H, L):
    """merged_kernel=False -- the form the reference driver itself builds (svd_acceleration_v3.py:117,143) -- on the tensor-core
    engine: the four per-gate blocks of a side are merged when the weights are packed, either concatenated (block-diagonal right
    factor: truncated 3-factor models) or multiplied out to the dense matrix (full rank, 2-factor form).  Same bar as merged."""
    full, _ = _models(H, L)
    split = svdlstm.make_LSTM_singular_model(full, merged_kernel=False, return_sequences=True)
    x = np.random.default_rng(11).standard_normal((70, 16, 16)).astype(np.float32)
    y = split.predict(x, engine="tc")                                   # full rank: R = 4 min(D, H) >= kin -> dense route
    assert split.last_engine() == svdlstm.ENGINE_TC
    _check(y, oracle_twin(oracle, split).predict(x), "tc split 3F full rank H=%d" % H)
    for r in (8, 24):
        m3 = svdlstm.truncate_singular_model(split, r)                  # recurrent side: 4r < H -> concat route
        _check(m3.predict(x, engine="tc"), oracle_twin(oracle, m3).predict(x), "tc split 3F r=%d H=%d" % (r, H))
        m2 = svdlstm.make_LSTM_reduced_model(split, rank=r, merged_kernel=False)   # 2-factor split -> dense route
        _check(m2.predict(x, engine="tc"), oracle_twin(oracle, m2).predict(x), "tc split 2F r=%d H=%d" % (r, H))
    # the regime switch takes split models to the tensor cores as well
    xb = np.random.default_rng(12).standard_normal((256, 8, 16)).astype(np.float32)
    m3 = svdlstm.truncate_singular_model(split, 8)
    m3.predict(xb)
    assert m3.last_engine() == svdlstm.ENGINE_TC


def test_tc_split_ragged_ranks_from_cutoff(oracle):
    """make_LSTM_reduced_model(smodel, cutoff=c, merged_kernel=False) (svd_acceleration_v3.py:143) keeps a different rank in every
    gate block; the merged block the tensor-core engine builds from them must place every ragged piece where it belongs."""
    full, _ = _models(128, 2, seed=3)
    split = svdlstm.make_LSTM_singular_model(full, merged_kernel=False, return_sequences=True)
    gate_max = [np.abs(w).reshape(4, -1).max(1).min() for l in split.layers[:-1] for w in l.get_weights()[:2]]
    cutoff = 0.6 * float(min(gate_max))                # every gate block keeps something; how much differs from block to block
    red = svdlstm.make_LSTM_reduced_model(split, cutoff=cutoff, merged_kernel=False)
    ranks = {int(w.shape[1]) for l in red.layers[:-1] for w in l.get_weights()[:-1:2]}
    assert len(ranks) > 1, "the cutoff did not produce ragged ranks: %s" % ranks
    x = np.random.default_rng(14).standard_normal((130, 12, 16)).astype(np.float32)
    y = red.predict(x, engine="tc")
    assert red.last_engine() == svdlstm.ENGINE_TC
    _check(y, oracle_twin(oracle, red).predict(x), "tc split 2F ragged ranks %s" % sorted(ranks))
    assert np.abs(y - red.predict(x, engine="fp32")).max() < 2e-3 * np.abs(y).max() + 2e-4


def test_tc_split_dropbear_model(dropbear_weights):
    """The shipped 3 x 15 model in the reference's own configuration (split 3-factor and its 2-factor reduction) on the
    tensor-core engine, against the FP32 engine: RMSE <= 3e-3 of the output RMS (the bar of the merged DROPBEAR tests)."""
    layers, dense = dropbear_weights
    full = svdlstm.full_model_from_weights(layers, dense)
    split = svdlstm.make_LSTM_singular_model(full, merged_kernel=False, return_sequences=True)
    red = svdlstm.make_LSTM_reduced_model(split, rank=8, merged_kernel=False)
    x = torch.randn(64, 300, 16, generator=torch.Generator().manual_seed(13)).cuda()
    for m, what in ((split, "split 3F"), (red, "split 2F r=8")):
        y32 = m(x, engine="general")
        ytc = m(x, engine="tc")
        rms = float((y32 ** 2).mean().sqrt())
        rmse = float(((ytc - y32) ** 2).mean().sqrt())
        assert rmse <= 3e-3 * rms, (what, rmse, rms)


def test_tc_time_major_and_go_backwards():
    """SingularLSTM kwargs time_major / go_backwards (svd_classes_v3.py:377-437) in the dense regime: the host feeds the
    tensor-core kernel the batch-major / time-reversed sequence; same results as the FP32 engine's own handling of the flags."""
    _, sm = _models(128, 2)
    layer_model = svdlstm.truncate_singular_model(sm, 24)
    h = layer_model._fused_handle()
    x = torch.randn(160, 12, 16, generator=torch.Generator().manual_seed(21)).cuda()
    for kw in (dict(go_backwards=True), dict(time_major=True), dict(go_backwards=True, time_major=True),
               dict(go_backwards=True, return_sequences=False)):
        xin = x.transpose(0, 1).contiguous() if kw.get("time_major") else x
        y32, _, _ = h.forward(xin, engine="fp32", **kw)
        ytc, _, _ = h.forward(xin, **kw)                       # engine auto: 160 sequences x 128 units -> tensor cores
        assert h.last_engine() == svdlstm.ENGINE_TC, kw
        assert ytc.shape == y32.shape, kw
        assert float((ytc - y32).abs().max()) < 2e-3 * float(y32.abs().max()) + 2e-4, kw


def test_predict_with_the_upload_inside_the_forward(monkeypatch):
    """model.predict(pinned host x) in the dense regime uploads x in time slices while the tensor-core kernel already runs
    (svdlstm_forward_streamed_input; the layer-0 input warp follows a progress word).  Bit-identical to the forward on the
    fully uploaded array; launches that cannot do it (32-sequence tiles, models the engine does not take) fall back to
    upload-then-forward."""
    _, sm = _models(256, 2)
    x = torch.randn(2400, 96, 16, generator=torch.Generator().manual_seed(50))   # 64-sequence tiles (32-wide ones would not fit the SMs)
    xp = svdlstm.pinned_empty((2400, 96, 16))
    xp.copy_(x)
    assert xp.is_pinned()
    for rank, streamed in ((128, True), (16, True)):
        m = svdlstm.truncate_singular_model(sm, rank)
        y_dev = m(x.cuda(), engine="tc").cpu().numpy()
        l0 = svdlstm.launches()
        y1 = m.predict(xp)
        assert m.last_engine() == svdlstm.ENGINE_TC
        assert np.array_equal(y1, y_dev), rank
        assert (svdlstm.launches() - l0 == 1) == streamed, (rank, svdlstm.launches() - l0)   # streamed: the one pipelined launch, no packing pass
        y2 = m.predict(xp)                                 # second request: buffers and progress word are reused
        assert np.array_equal(y2, y_dev)
        monkeypatch.setenv("SVDLSTM_STREAMED_INPUT", "0")
        assert np.array_equal(m.predict(xp), y_dev)
        monkeypatch.delenv("SVDLSTM_STREAMED_INPUT")
    # a batch that runs as 32-sequence tiles (no raw-x build): the slices are uploaded, then the plain forward runs
    m = svdlstm.truncate_singular_model(sm, 128)
    xs = svdlstm.pinned_empty((300, 96, 16))
    xs.copy_(x[:300])
    l0 = svdlstm.launches()
    assert np.array_equal(m.predict(xs), m(xs.cuda(), engine="tc").cpu().numpy())
    # a model the tensor-core engine does not take at all (512 full-rank units): plain upload + FP32 engine, same answer as on device
    layers512, dense512 = svdlstm.synthetic_layers(16, 512, 1, seed=0)
    wide = svdlstm.full_model_from_weights(layers512, dense512, return_sequences=True)
    xw = svdlstm.pinned_empty((256, 64, 16))
    xw.copy_(torch.randn(256, 64, 16, generator=torch.Generator().manual_seed(52)))
    assert np.array_equal(wide.predict(xw), wide(xw.cuda()).cpu().numpy()) and wide.last_engine() == svdlstm.ENGINE_GENERAL
    # odd lengths / slices that do not divide T
    m = svdlstm.truncate_singular_model(sm, 128)
    xo = svdlstm.pinned_empty((2500, 71, 16))
    xo.copy_(torch.from_numpy(np.random.default_rng(51).standard_normal((2500, 71, 16)).astype(np.float32)))
    assert np.array_equal(m.predict(xo), m(xo.cuda(), engine="tc").cpu().numpy())


def test_tc_rejects_what_it_cannot_run():
    """No silent fallback: unsupported models / calls raise with the reason."""
    full, sm = _models(128, 1)
    x = torch.randn(4, 5, 16).cuda()
    y_full = full(x, engine="tc")                                 # unfactored cell: runs as I . W (units <= 256)
    assert float((y_full - full(x, engine="general")).abs().max()) < 2e-3
    layers512, dense512 = svdlstm.synthetic_layers(16, 512, 1, seed=0)
    with pytest.raises((RuntimeError, ValueError), match="ranks above 256"):
        svdlstm.full_model_from_weights(layers512, dense512, return_sequences=True)(x, engine="tc")
    layers, dense = svdlstm.synthetic_layers(16, 1100, 1, seed=0)   # more than 1024 units
    with pytest.raises((RuntimeError, ValueError), match="units|ranks"):
        svdlstm.full_model_from_weights(layers, dense, return_sequences=True)(x, engine="tc")


def test_tc_launch_shapes_agree(monkeypatch):
    """The engine picks its launch shape from the batch: all layers in one co-resident launch ("pipe", 32- or 64-sequence
    tiles) when layers x tiles fits the SMs, else one launch per layer ("seq").  The per-sequence arithmetic is the same
    in every shape, so the outputs must agree to float32 rounding of the Dense-top sum."""
    _, sm = _models(256, 2)
    m = svdlstm.truncate_singular_model(sm, 48)
    x = torch.randn(200, 24, 16, generator=torch.Generator().manual_seed(9)).cuda()
    outs = {}
    for mode, ns in (("seq", "32"), ("seq", "64"), ("pipe", "32"), ("pipe", "64")):
        monkeypatch.setenv("SVDLSTM_TC_MODE", mode)
        monkeypatch.setenv("SVDLSTM_TC_NS", ns)
        outs[(mode, ns)] = m(x, engine="tc").clone()
    monkeypatch.delenv("SVDLSTM_TC_MODE")
    monkeypatch.delenv("SVDLSTM_TC_NS")
    outs["auto"] = m(x, engine="tc").clone()
    ref = outs[("seq", "32")]
    for k, v in outs.items():
        assert float((v - ref).abs().max()) < 2e-6, k


def test_tc_three_layers_pipelined_and_large_batch_fallback(oracle):
    """3 layers x tiles <= SMs runs pipelined with one hand-off image per layer; a batch too large for that falls back
    to per-layer launches.  Both against the float64 oracle on a sample."""
    _, sm = _models(128, 3)
    m = svdlstm.truncate_singular_model(sm, 24)
    x = np.random.default_rng(10).standard_normal((70, 16, 16)).astype(np.float32)
    _check(m.predict(x, engine="tc"), oracle_twin(oracle, m).predict(x), "tc 3 layers pipelined")
    xb = torch.randn(148 * 64 + 40, 6, 16, generator=torch.Generator().manual_seed(11)).cuda()   # > 148 tiles of 64 per layer
    yb = m(xb, engine="tc")
    idx = torch.tensor([0, 31, 32, 63, 64, 5000, 148 * 64 + 39])
    _check(yb[idx.cuda()].cpu().numpy(), oracle_twin(oracle, m).predict(xb[idx.cuda()].cpu().numpy()), "tc large batch (per-layer launches)")


def test_tc_full_c3_size_properties(oracle):
    """BASELINE configs[2] at full size (B=4096, T=1024, H=256, L=2, rank 128 -> all layers in one pipelined launch,
    64-sequence tiles, streamed weights).  Size-independent properties: (a) a sequence's output does not depend on the
    batch it travels in -- rows taken from the big run equal the same rows run as a small batch (different tile, column,
    launch shape) to FP32 rounding of the Dense sum; (b) against the float64 oracle over all 1024 steps the error stays
    at the reduced-precision level and does not grow with time; (c) the run is deterministic."""
    _, sm = _models(256, 2)
    m = svdlstm.truncate_singular_model(sm, 128)
    x = torch.randn(4096, 1024, 16, generator=torch.Generator().manual_seed(12)).cuda()
    y = m(x, engine="tc")
    assert tuple(y.shape) == (4096, 1024, 1) and bool(torch.isfinite(y).all())
    idx = torch.tensor([0, 1, 63, 64, 65, 2047, 2048, 4032, 4095]).cuda()
    y_small = m(x[idx].contiguous(), engine="tc")
    assert float((y_small - y[idx]).abs().max()) < 2e-6
    # (b) against the float64 ORACLE (not another CUDA engine) on the same weights, all 1024 steps of the sampled sequences
    y_or = torch.from_numpy(oracle_twin(oracle, m).predict(x[idx].cpu().numpy())).float()
    err = (y[idx].cpu() - y_or).abs()
    scale = float(y_or.abs().max())
    assert float(err.max()) < 2e-3 * scale + 2e-4
    assert float(err[:, -128:].max()) < 2.0 * float(err[:, :512].max()) + 1e-4     # no drift over the sequence
    # ... and the FP32 general engine at this full-size shape holds the 1e-5 bar against the same oracle
    from helpers import assert_parity
    assert_parity(m(x[idx].contiguous(), engine="general").cpu().numpy(), y_or.double().numpy(), "C3 rank 128 general engine, T=1024")
    assert torch.equal(m(x, engine="tc"), y)


def test_tc_units_1024_matches_oracle(oracle):
    """BASELINE configs[4] shape (units = 1024, 3 layers, rank 128): 8 unit blocks, 16 epilogue warps (the cell state of
    1024 cells x 32 sequences lives in their registers), weights streamed through 16 KB ring slots; ragged batch."""
    _, sm = _models(1024, 3)
    m = svdlstm.truncate_singular_model(sm, 128)
    x = np.random.default_rng(9).standard_normal((37, 7, 16)).astype(np.float32)
    y = m.predict(x, engine="tc")
    assert m.last_engine() == svdlstm.ENGINE_TC
    _check(y, oracle_twin(oracle, m).predict(x), "tc units=1024 L=3 r=128")


def test_tc_large_batch_runs_as_pipelined_chunks():
    """A batch too large for one co-resident launch (layers x tiles > SM count) is cut into chunks that are each pipelined
    with 64-sequence tiles; per-sequence results must equal those of a small batch bit for bit (the chunking, the tile
    width and the position of a sequence never change its arithmetic)."""
    _, sm = _models(256, 2)
    m = svdlstm.truncate_singular_model(sm, 48)
    B = 74 * 64 + 200          # one full chunk of 148 / 2 tiles + a ragged one
    x = torch.randn(B, 5, 16, generator=torch.Generator().manual_seed(11)).cuda()
    y = m(x, engine="tc")
    assert bool(torch.isfinite(y).all())
    idx = torch.tensor([0, 63, 64, 4735, 4736, B - 1]).cuda()
    assert torch.equal(m(x[idx], engine="tc"), y[idx])
    assert torch.equal(m(x[4000:4100], engine="tc"), y[4000:4100])


def test_tc_last_output_only():
    """return_sequences=False (svd_classes_v3.py:428-431) on the tensor-core engine = the last step of the sequence it
    always computes; must equal the sliced full output bit for bit and match the FP32 engine's last output."""
    layers, dense = svdlstm.synthetic_layers(16, 128, 2, seed=3)
    full = svdlstm.full_model_from_weights(layers, dense)          # return_sequences=False model
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=False)
    m = svdlstm.truncate_singular_model(sm, 24)
    x = torch.randn(50, 12, 16, generator=torch.Generator().manual_seed(10)).cuda()
    y_tc = m(x, engine="tc")
    y32 = m(x, engine="general")
    assert tuple(y_tc.shape) == tuple(y32.shape) == (50, 1)
    assert float((y_tc - y32).abs().max()) < 2e-3 * float(y32.abs().max()) + 2e-4


@pytest.mark.parametrize("H,L,rank,B", [(256, 2, 128, 100), (128, 2, 16, 45), (256, 1, 64, 33)])
def test_tc_state_carried_over_time_chunks(H, L, rank, B):
    """initial_state / return_state on the tensor-core engine (svd_classes_v3.py:393,433-434): a sequence run in two
    time chunks with (h, c) carried over must equal the unchunked run BIT FOR BIT (h lives as FP16 between steps
    either way, c as FP32), and the returned state must match the FP32 engine's within the engine tolerance."""
    _, sm = _models(H, L)
    m = svdlstm.truncate_singular_model(sm, rank)
    x = torch.randn(B, 14, 16, generator=torch.Generator().manual_seed(12)).cuda()
    m(x[:2], engine="tc")                                   # builds the fused handle
    h = m._fused_handle()
    y_all, hs_all, cs_all = h.forward(x, want_state=True, engine="tc")
    y_a, hs_a, cs_a = h.forward(x[:, :5].contiguous(), want_state=True, engine="tc")
    y_b, hs_b, cs_b = h.forward(x[:, 5:].contiguous(), initial_state=(hs_a, cs_a), want_state=True, engine="tc")
    assert torch.equal(torch.cat([y_a, y_b], 1), y_all)
    for l in range(L):
        assert torch.equal(cs_b[l], cs_all[l]) and torch.equal(hs_b[l], hs_all[l])
    y32, hs32, cs32 = h.forward(x, want_state=True, engine="general")
    for l in range(L):
        assert float((hs_all[l] - hs32[l]).abs().max()) < 3e-3
        assert float((cs_all[l] - cs32[l]).abs().max()) < 3e-3 * max(1.0, float(cs32[l].abs().max()))


def test_tc_stream_api_matches_one_long_run():
    """Sequential.stream: three time chunks with the state handed back in == one run (tensor-core engine: bit for bit)."""
    _, sm = _models(256, 2)
    m = svdlstm.truncate_singular_model(sm, 96)
    x = torch.randn(70, 21, 16, generator=torch.Generator().manual_seed(13)).cuda()
    y_all = m(x, engine="tc")
    state, parts = None, []
    for lo, hi in ((0, 4), (4, 13), (13, 21)):
        y, state = m.stream(x[:, lo:hi].contiguous(), state, engine="tc")
        parts.append(y)
    assert torch.equal(torch.cat(parts, 1), y_all)


@pytest.mark.parametrize("T", [1, 2, 3])
def test_tc_very_short_sequences(oracle, T):
    """T = 1..3: the Dense-top / t_w rows of step T-1 come from the extra flush pass; nothing may depend on T >= a few steps."""
    _, sm = _models(256, 2)
    m = svdlstm.truncate_singular_model(sm, 128)
    x = np.random.default_rng(20 + T).standard_normal((70, T, 16)).astype(np.float32)
    _check(m.predict(x, engine="tc"), oracle_twin(oracle, m).predict(x), "tc T=%d" % T)


def test_tc_mixed_ranks_and_wide_dense(oracle):
    """Different W / U ranks per layer (the greedy sweep's ragged ranks) and a Dense top with several outputs: the t_w hand-off
    rows, the t_u rows and the Dense rows share the S1u tiles in every combination."""
    layers, dense = svdlstm.synthetic_layers(16, 256, 3, seed=5, n_out=5)
    full = svdlstm.full_model_from_weights(layers, dense)
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    for ranks in ((40, 72), (128, 24), (100, 128)):
        m = svdlstm.truncate_singular_model(sm, ranks)      # (rank of the input factors, rank of the recurrent factors)
        x = np.random.default_rng(31).standard_normal((50, 6, 16)).astype(np.float32)
        _check(m.predict(x, engine="tc"), oracle_twin(oracle, m).predict(x), "tc mixed ranks %s" % (ranks,))


def test_tc_full_model_matches_oracle(oracle):
    """The trained full (unfactored) LSTM on the tensor-core engine (as the factorisation I . W): the baseline the truncated
    models are compared with at equal engine; 2 layers so the hand-off path carries h itself."""
    layers, dense = svdlstm.synthetic_layers(16, 256, 2, seed=2)
    full = svdlstm.full_model_from_weights(layers, dense, return_sequences=True)
    x = np.random.default_rng(41).standard_normal((70, 9, 16)).astype(np.float32)
    y = full.predict(x, engine="tc")
    assert full.last_engine() == svdlstm.ENGINE_TC
    _check(y, oracle_twin(oracle, full).predict(x), "tc full model H=256 L=2")


def test_tc_dropbear_model_padded_units(oracle):
    """J2: the shipped 3 x 15 DROPBEAR model on the tensor-core engine.  Units are padded 15 -> 128 with zero weights (padded
    cells stay at h = c = 0 exactly), layers run pipelined.  Full model, 3-factor ranks and return_state round trip."""
    layers, dense = svdlstm.load_model_weights_npz(GOLDEN_W)
    full = svdlstm.full_model_from_weights(layers, dense)
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    x = np.random.default_rng(21).standard_normal((70, 200, 16)).astype(np.float32)
    y_or = oracle_twin(oracle, full).predict(x)
    def check_db(y_tc, y_ref, what):
        # Trained weights, saturating gates, 200 steps through 3 layers: a numpy emulation of nothing but the FP16 rounding of the
        # operands (weights, h, t) reproduces these errors (max 1.0e-2 / 3.3e-2 / 9.5e-3 / 1.1e-3 at ranks 15 / 10 / 4 / 1 on outputs
        # of |y| <= 2.8, RMSE 1e-3 .. 1.7e-3) -- the truncated models amplify single roundings more than the synthetic ones do.  Bars: RMSE <=
        # 3e-3 of the output RMS (the "RMSE delta" north_star asks for), max <= 2e-2 of the output scale.
        y_tc, y_ref = np.asarray(y_tc, np.float64), np.asarray(y_ref, np.float64)
        assert np.isfinite(y_tc).all(), what
        rms = np.sqrt(np.mean(y_ref ** 2))
        rmse = np.sqrt(np.mean((y_tc - y_ref) ** 2))
        assert rmse <= 3e-3 * rms, "%s: RMSE %.3e > 3e-3 x rms %.3e" % (what, rmse, rms)
        assert np.abs(y_tc - y_ref).max() <= 2e-2 * np.abs(y_ref).max(), "%s: max err %.3e" % (what, np.abs(y_tc - y_ref).max())

    check_db(full.predict(x, engine="tc"), y_or, "tc DROPBEAR full")
    assert full.last_engine() == svdlstm.ENGINE_TC
    for r in (15, 10, 4, 1):
        m = svdlstm.truncate_singular_model(sm, r)
        check_db(m.predict(x, engine="tc"), oracle_twin(oracle, m).predict(x), "tc DROPBEAR 3F r=%d" % r)
    # chunked == unchunked with the state carried (state arrays are (B, 15): the true units)
    m = svdlstm.truncate_singular_model(sm, 8)
    xd = torch.from_numpy(x).cuda()
    y_all = m(xd, engine="tc")
    ya, st = m.stream(xd[:, :77], None, engine="tc")
    assert tuple(st[0][0].shape) == (70, 15)
    yb, _ = m.stream(xd[:, 77:], st, engine="tc")
    assert torch.equal(torch.cat([ya, yb], 1), y_all)


def test_auto_engine_regime_switch(monkeypatch):
    """north_star: "switches to tensor-core tiles only when batch x 4H makes the thin contraction genuinely dense".
    engine=None / "auto": batch >= 128 and units >= 64 on a supported model -> tcgen05 engine; small batches, small models,
    unsupported forms and engine="fp32" -> FP32 engines.  This is what rmodel.predict(X) (svd_acceleration_v3.py:150-151) runs."""
    _, sm = _models(256, 2)
    m = svdlstm.truncate_singular_model(sm, 128)
    xb = torch.randn(4096, 8, 16, generator=torch.Generator().manual_seed(30)).cuda()
    y = m.predict(xb.cpu().numpy())                     # no engine argument, host array in / out
    assert m.last_engine() == svdlstm.ENGINE_TC
    y32 = m.predict(xb.cpu().numpy(), engine="fp32")
    assert m.last_engine() == svdlstm.ENGINE_GENERAL
    assert float(np.abs(y - y32).max()) < 2e-3 * float(np.abs(y32).max()) + 2e-4
    m(xb[:64])
    assert m.last_engine() == svdlstm.ENGINE_GENERAL      # 64 sequences: not dense enough
    r2 = svdlstm.make_LSTM_reduced_model(sm, rank=64)       # the reference's own timed object: a 2-factor model
    r2(xb)
    assert r2.last_engine() == svdlstm.ENGINE_TC
    split = svdlstm.make_LSTM_singular_model(_models(256, 1)[0], merged_kernel=False, return_sequences=True)
    split(xb[:256])
    assert split.last_engine() == svdlstm.ENGINE_TC       # split cells (the reference driver's form) take the fast path too
    split512 = svdlstm.make_LSTM_singular_model(_models(512, 1)[0], merged_kernel=False, return_sequences=True)
    split512(xb[:256])
    assert split512.last_engine() == svdlstm.ENGINE_GENERAL  # ... unless neither merge route fits (gate ranks sum > 256, kin > 256)
    layers, dense = svdlstm.load_model_weights_npz(GOLDEN_W)
    small = svdlstm.full_model_from_weights(layers, dense)
    small(xb[:512])
    assert small.last_engine() in (svdlstm.ENGINE_WAVEFRONT, svdlstm.ENGINE_GENERAL)   # 15 units: latency regime
    # return_sequences=False (the reference builders' default) also takes the fast path: last step of the sequence
    mlast = svdlstm.truncate_singular_model(svdlstm.make_LSTM_singular_model(_models(256, 2)[0], merged_kernel=True), 32)
    ylast = mlast(xb)
    assert tuple(ylast.shape) == (4096, 1) and mlast.last_engine() == svdlstm.ENGINE_TC


def test_tc_sees_dense_and_variable_updates(oracle):
    """the engine bakes the Dense kernel and the factors into FP16 images.  In-place updates through
    Dense.set_weights / TimeDistributed.set_weights / Variable.assign must invalidate them."""
    _, sm = _models(128, 2)
    m = svdlstm.truncate_singular_model(sm, 16)
    x = np.random.default_rng(40).standard_normal((40, 10, 16)).astype(np.float32)
    _check(m.predict(x, engine="tc"), oracle_twin(oracle, m).predict(x), "before update")
    dk, db = m.layers[-1].get_weights()
    m.layers[-1].set_weights([dk * -1.7, db + 0.3])
    _check(m.predict(x, engine="tc"), oracle_twin(oracle, m).predict(x), "after Dense.set_weights")
    cell = m.layers[0].cell
    cell.recurrent_kernel.assign(cell.recurrent_kernel.numpy() * 0.5)      # sigma_u through Variable.assign
    _check(m.predict(x, engine="tc"), oracle_twin(oracle, m).predict(x), "after Variable.assign")
    m.layers[-1].layer.bias.assign(np.array([1.25], np.float32))
    _check(m.predict(x, engine="tc"), oracle_twin(oracle, m).predict(x), "after Dense bias assign")


@pytest.mark.parametrize("D", [16, 10, 24])
def test_tc_raw_x_loader_paths(oracle, D):
    """The 64-sequence pipelined launch with a long weight stream reads the caller's float32 x itself (RAWX kernel build):
    vectorised register path (D % 4 == 0, D <= 16), generic path (D = 10: unaligned rows; D = 24: more than four float4 per row),
    ragged last tile.  Same rows through the small-batch launch (bulk-copy loader + pack_x) must agree to FP32 rounding of the
    Dense sum, and both with the float64 oracle."""
    layers, dense = svdlstm.synthetic_layers(D, 256, 2, seed=3)
    full = svdlstm.full_model_from_weights(layers, dense)
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    m = svdlstm.truncate_singular_model(sm, 128)
    x = torch.randn(2400 + 37, 5, D, generator=torch.Generator().manual_seed(31)).cuda()     # 39 tiles of 64 per layer: pipelined, ragged
    y = m(x, engine="tc")
    idx = torch.tensor([0, 63, 64, 1000, 2399, 2400, 2436]).cuda()
    y_small = m(x[idx].contiguous(), engine="tc")
    assert float((y_small - y[idx]).abs().max()) < 2e-6
    _check(y[idx].cpu().numpy(), oracle_twin(oracle, m).predict(x[idx].cpu().numpy()), "tc raw-x D=%d" % D)
    assert torch.equal(m(x, engine="tc"), y)
