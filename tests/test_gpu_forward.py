"""GPU parity tests (run with -m gpu on a B200): CUDA forward path vs the float64 oracle on the same
weights and inputs, through the reference-facing Python API -> C-ABI.

Tolerance (north_star: "within 1e-5 relative per output"): |y - ref| <= 1e-5*|ref| + 2e-6.  The
absolute term covers outputs crossing zero (h in (-1,1); 2e-6 is ~30 float32 ulps of 1)."""
import numpy as np
import pytest
import torch

import svdlstm

pytestmark = pytest.mark.gpu

from helpers import assert_parity, cond_slack, oracle_twin


@pytest.fixture(scope="module")
def full(dropbear_weights):
    layers, dense = dropbear_weights
    return svdlstm.full_model_from_weights(layers, dense, return_sequences=True)


@pytest.fixture(scope="module")
def x_small():
    return np.random.default_rng(0).standard_normal((3, 64, 16)).astype(np.float32)


@pytest.mark.parametrize("engine", ["general", "wavefront"])
def test_full_model_parity_and_kat(oracle, full, kat, x_small, engine):
    y = full.predict(kat["sin_x"].astype(np.float32), engine=engine)[0, :, 0]
    assert_parity(y, kat["sin_y_full"], "SURVEY App. D KAT (%s)" % engine)
    assert_parity(full.predict(x_small, engine=engine), oracle_twin(oracle, full).predict(x_small), "full " + engine)


@pytest.mark.parametrize("engine", ["general", "wavefront"])
@pytest.mark.parametrize("merged", [True, False])
def test_singular_and_reduced_parity_all_ranks(oracle, full, x_small, merged, engine):
    sm = svdlstm.make_LSTM_singular_model(full, hoyer=0.01, merged_kernel=merged, return_sequences=True)
    y_full = oracle_twin(oracle, full).predict(x_small)
    y_sm = sm.predict(x_small, engine=engine)
    assert_parity(y_sm, oracle_twin(oracle, sm).predict(x_small), "3-factor same-weights")
    # device SVD (Jacobi) + device forward == the full model (algebraic identity at full rank)
    assert np.max(np.abs(y_sm - y_full)) < 2e-5
    for r in (15, 12, 8, 4, 1):
        tm = svdlstm.truncate_singular_model(sm, r)
        assert_parity(tm.predict(x_small, engine=engine), oracle_twin(oracle, tm).predict(x_small), "3F top-%d" % r)
        rm = svdlstm.make_LSTM_reduced_model(sm, merged_kernel=merged, rank=r)
        assert_parity(rm.predict(x_small, engine=engine), oracle_twin(oracle, rm).predict(x_small), "2F top-%d" % r,
                      ref32=oracle_twin(oracle, rm, np.float32).predict(x_small), extra_atol=cond_slack(rm))
        # 2-factor and 3-factor forms of the same top-r truncation agree (they differ only in cost)
        assert np.max(np.abs(rm.predict(x_small, engine=engine) - tm.predict(x_small, engine=engine))) < 5e-5
    rm = svdlstm.make_LSTM_reduced_model(sm, cutoff=.05, merged_kernel=merged)
    assert_parity(rm.predict(x_small, engine=engine), oracle_twin(oracle, rm).predict(x_small), "2F cutoff .05",
                  ref32=oracle_twin(oracle, rm, np.float32).predict(x_small), extra_atol=cond_slack(rm))
    if merged:   # cutoff .05 prunes nothing on merged matrices => reduced == full
        assert np.max(np.abs(rm.predict(x_small, engine=engine) - y_full)) < 5e-5


def test_device_builders_match_oracle_builders(oracle, full, dropbear_weights, x_small):
    """End to end: device Jacobi SVD + K2b vs np.linalg.svd / inv (the reference's own calls)."""
    layers, dense = dropbear_weights
    ofull = oracle.model_from_weights(layers, dense, dtype=np.float64)
    for merged in (True, False):
        osm = oracle.make_LSTM_singular_model(ofull, merged_kernel=merged, return_sequences=True, svd_dtype=np.float64)
        sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=merged, return_sequences=True)
        for r in (12, 5):
            y_dev = svdlstm.make_LSTM_reduced_model(sm, merged_kernel=merged, rank=r).predict(x_small)
            y_or = oracle.make_LSTM_reduced_model(osm, merged_kernel=merged, rank=r).predict(x_small)
            assert np.max(np.abs(y_dev - y_or)) < 5e-5, (merged, r)
        # threshold semantics: same kept ranks as the oracle at a cutoff that really prunes
        cut = 1.0
        rm = svdlstm.make_LSTM_reduced_model(sm, cutoff=cut, merged_kernel=merged)
        orm = oracle.make_LSTM_reduced_model(osm, cutoff=cut, merged_kernel=merged)
        assert [w.shape for w in rm.get_weights()] == [np.shape(w) for w in orm.get_weights()]
        assert np.max(np.abs(rm.predict(x_small) - orm.predict(x_small))) < 1e-4
        assert svdlstm.count_weights(rm) == oracle.count_weights(orm)


def test_auto_engine_picks_wavefront_and_long_sequence(oracle, full):
    x = np.random.default_rng(1).standard_normal((1, 6000, 16)).astype(np.float32)
    y = full.predict(x)
    assert full.last_engine() == svdlstm.ENGINE_WAVEFRONT
    # 18 000 dependent cell updates: float32 rounding noise itself is ~5e-6 here (ref32 term)
    assert_parity(y, oracle_twin(oracle, full).predict(x), "T=6000 wavefront",
                  ref32=oracle_twin(oracle, full, np.float32).predict(x))
    yg = full.predict(x, engine="general")
    assert np.max(np.abs(y - yg)) < 1e-5


def test_layer_kwargs(oracle, dropbear_weights):
    (W, U, b) = dropbear_weights[0][0]
    rng = np.random.default_rng(2)
    x = rng.standard_normal((4, 20, 16)).astype(np.float32)
    ocell = oracle.FullCell(15, W, U, b)
    h0 = rng.standard_normal((4, 15)).astype(np.float32) * 0.3
    c0 = rng.standard_normal((4, 15)).astype(np.float32) * 0.3
    mask = np.ones((4, 20), bool)
    mask[1, 3:9] = False
    mask[2, :4] = False
    for engine in ("general", "wavefront"):
        def mk(**kw):
            return svdlstm.LSTM(15, weights=[W, U, b], engine=engine, **kw)
        out = mk(return_sequences=True, return_state=True)(x, initial_state=[h0, c0])
        ref = oracle.rnn_layer(ocell, x, initial_state=[h0, c0], return_sequences=True, return_state=True)
        for a, r in zip(out, ref):
            assert_parity(a.cpu().numpy(), r, "initial_state/return_state " + engine)
        assert_parity(mk()(x).cpu().numpy(), oracle.rnn_layer(ocell, x), "last output " + engine)
        assert_parity(mk(return_sequences=True, go_backwards=True)(x).cpu().numpy(),
                      oracle.rnn_layer(ocell, x, go_backwards=True, return_sequences=True), "go_backwards " + engine)
        xt = np.ascontiguousarray(np.swapaxes(x, 0, 1))
        assert_parity(mk(return_sequences=True, time_major=True)(xt).cpu().numpy(),
                      oracle.rnn_layer(ocell, xt, time_major=True, return_sequences=True), "time_major " + engine)
        # stateful: two chunks == one pass (svd_classes_v3.py:421-426)
        st = mk(return_sequences=True, stateful=True)
        y12 = torch.cat([st(x[:, :7]), st(x[:, 7:])], 1).cpu().numpy()
        assert_parity(y12, oracle.rnn_layer(ocell, x, return_sequences=True), "stateful " + engine)
        st.reset_states()
        assert_parity(st(x[:, :7]).cpu().numpy(), oracle.rnn_layer(ocell, x[:, :7], return_sequences=True), "reset_states")
    for zero in (False, True):
        lay = svdlstm.LSTM(15, weights=[W, U, b], return_sequences=True, return_state=True, zero_output_for_mask=zero)
        out = lay(x, mask=mask)
        ref = oracle.rnn_layer(ocell, x, mask=mask, return_sequences=True, return_state=True, zero_output_for_mask=zero)
        for a, r in zip(out, ref):
            assert_parity(a.cpu().numpy(), r, "mask zero=%s" % zero)
    with pytest.raises(ValueError):
        svdlstm.LSTM(15, weights=[W, U, b], engine="wavefront")(x, mask=mask)


@pytest.mark.parametrize("H", [15, 80])
def test_general_engine_tile_width_invariance(monkeypatch, H):
    """The FP32 general engine processes BT = 1..16 sequences per CTA (chosen from the batch); every sequence's arithmetic is
    the same at every width, so the outputs -- and the final states, with a mask, with both cell forms -- must be BIT-identical
    across widths.  H = 15 takes the narrow-layer path (thread per gate column), H = 80 the wide one (thread per unit)."""
    rng = np.random.default_rng(40 + H)
    W = (0.4 * rng.standard_normal((16, 4 * H))).astype(np.float32)
    U = (0.4 * rng.standard_normal((H, 4 * H))).astype(np.float32)
    b = (0.1 * rng.standard_normal(4 * H)).astype(np.float32)
    dense = ((0.3 * rng.standard_normal((H, 1))).astype(np.float32), np.zeros(1, np.float32))
    full_m = svdlstm.full_model_from_weights([(W, U, b)], dense)
    sm = svdlstm.make_LSTM_singular_model(full_m, merged_kernel=False, return_sequences=True)
    models = [full_m, svdlstm.truncate_singular_model(sm, 5), svdlstm.make_LSTM_reduced_model(sm, rank=6, merged_kernel=False)]
    x = torch.randn(2100, 5, 16, generator=torch.Generator().manual_seed(41)).cuda()
    mask = (torch.rand(2100, 5, generator=torch.Generator().manual_seed(42)) > 0.3).cuda()
    lay = svdlstm.LSTM(H, weights=[W, U, b], return_sequences=True, return_state=True, engine="general")
    ref = None
    for bt in ("1", "4", "16"):
        monkeypatch.setenv("SVDLSTM_GEN_BT", bt)
        outs = [m(x, engine="general") for m in models] + list(lay(x, mask=mask))
        if ref is None:
            ref = outs
        else:
            for i, (a, r) in enumerate(zip(outs, ref)):
                assert torch.equal(a, r), "BT=%s output %d differs from BT=1" % (bt, i)
    monkeypatch.delenv("SVDLSTM_GEN_BT")
    for m, r in zip(models, ref):          # the library's own choice (2100 sequences -> 16 per CTA)
        assert torch.equal(m(x, engine="general"), r)


def test_cell_call_contract(oracle, full, dropbear_weights):
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True)
    rm = svdlstm.make_LSTM_reduced_model(sm, rank=7)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 16)).astype(np.float32)
    h = rng.standard_normal((5, 15)).astype(np.float32) * 0.5
    c = rng.standard_normal((5, 15)).astype(np.float32) * 0.5
    for model in (full, sm, rm):
        cell = model.layers[0].cell
        out, (h1, c1) = cell.call(x, [h, c])
        oc = oracle_twin(oracle, model).cells[0]
        hr, cr = oc.step(x.astype(np.float64), h.astype(np.float64), c.astype(np.float64))
        assert out is h1
        slack = 1e-5 if model is rm else None     # 2-factor: see assert_parity
        assert_parity(h1.cpu().numpy(), hr, "cell h", ref32=None if slack is None else hr + slack / 3)
        assert_parity(c1.cpu().numpy(), cr, "cell c", ref32=None if slack is None else cr + slack / 3)
    assert [v.name for v in sm.layers[0].cell.weights] == ["kernel", "recurrent_kernel", "w_left", "w_right", "u_left", "u_right", "bias"]
    assert [p.name for p in sm.layers[0].get_prunable_weights()] == ["kernel", "recurrent_kernel"]
    assert sm.layers[0].cell.kernel.numpy().shape == (1, 16)


def test_set_weights_errors_and_update(oracle, full, x_small):
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    lay = sm.layers[1]
    w = lay.get_weights()
    with pytest.raises(ValueError, match="expecting 7 weights"):
        lay.set_weights(w[:6])
    bad = list(w)
    bad[2] = bad[2][:, :5]
    with pytest.raises(ValueError, match="not compatible"):
        lay.set_weights(bad)
    w2 = [a.copy() for a in w]
    w2[0][0, -3:] = 0.0            # zero three singular values (what Hoyer fine-tuning drives towards)
    lay.set_weights(w2)
    assert_parity(sm.predict(x_small), oracle_twin(oracle, sm).predict(x_small), "after set_weights")
    # S > cutoff now prunes those three
    rm = svdlstm.make_LSTM_reduced_model(sm, cutoff=.05)
    assert rm.layers[1].cell.ranks[0] == 12
    assert np.max(np.abs(rm.predict(x_small) - sm.predict(x_small))) < 5e-5
    with pytest.raises(ValueError):
        svdlstm.make_LSTM_reduced_model(sm, merged_kernel=False)
    with pytest.raises(ValueError):
        svdlstm.SingularLSTMCell(15, w=[w[2], w[0], w[3]], u=[w[4], w[1], w[5][:, :30]], b=w[6]).build((None, 15))


def test_medium_model_general_engine(oracle):
    """C3-shaped but small enough for the float64 oracle: L=2, H=64, D=16, ranks 8/32/64."""
    layers, dense = svdlstm.synthetic_layers(16, 64, 2, seed=0)
    full = svdlstm.full_model_from_weights(layers, dense)
    x = np.random.default_rng(4).standard_normal((9, 24, 16)).astype(np.float32)
    assert full.predict(x).shape == (9, 24, 1)
    assert full.last_engine() == svdlstm.ENGINE_GENERAL
    assert_parity(full.predict(x), oracle_twin(oracle, full).predict(x), "H=64 full")
    for merged in (True, False):
        sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=merged, return_sequences=True)
        for r in (8, 32, 64):
            tm = svdlstm.truncate_singular_model(sm, r)
            assert_parity(tm.predict(x), oracle_twin(oracle, tm).predict(x), "H=64 3F r=%d merged=%s" % (r, merged))
            rm = svdlstm.make_LSTM_reduced_model(sm, merged_kernel=merged, rank=r)
            assert_parity(rm.predict(x), oracle_twin(oracle, rm).predict(x), "H=64 2F r=%d merged=%s" % (r, merged),
                          ref32=oracle_twin(oracle, rm, np.float32).predict(x))


def test_batch_permutation_and_padding_properties(full):
    """Size-independent properties at a batch that is not a multiple of the CTA tile."""
    x = torch.randn(37, 50, 16, generator=torch.Generator().manual_seed(5)).cuda()
    perm = torch.randperm(37, generator=torch.Generator().manual_seed(6)).cuda()
    for engine in ("general", "wavefront"):
        y = full(x, engine=engine)
        assert torch.equal(full(x[perm], engine=engine), y[perm])
        assert torch.equal(full(x[:5], engine=engine), y[:5])
        assert torch.isfinite(y).all()


@pytest.mark.parametrize("rank", [8, 32, 128, 256])
def test_fp32_parity_at_c3_shape(oracle, rank):
    """BASELINE configs[2] model (L=2, H=256, D=16) on the FP32 general engine vs the float64 oracle on the same weights,
    <= 1e-5 relative (+2e-6 abs): 3-factor and 2-factor forms at the swept ranks, on a subsample of sequences/steps the
    oracle finishes in seconds.  (The full-size run, T=1024, is in test_gpu_tc.py::test_tc_full_c3_size_properties.)"""
    layers, dense = svdlstm.synthetic_layers(16, 256, 2, seed=0)
    full = svdlstm.full_model_from_weights(layers, dense)
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    x = np.random.default_rng(50 + rank).standard_normal((6, 96, 16)).astype(np.float32)
    tm = svdlstm.truncate_singular_model(sm, rank)
    assert_parity(tm.predict(x, engine="general"), oracle_twin(oracle, tm).predict(x), "C3 3F r=%d" % rank)
    rm = svdlstm.make_LSTM_reduced_model(sm, rank=rank)
    assert_parity(rm.predict(x, engine="general"), oracle_twin(oracle, rm).predict(x), "C3 2F r=%d" % rank,
                  ref32=oracle_twin(oracle, rm, np.float32).predict(x), extra_atol=cond_slack(rm))
    if rank == 256:     # full rank: both factored forms reproduce the unfactored model
        assert_parity(full.predict(x, engine="general"), oracle_twin(oracle, full).predict(x), "C3 full")
        assert np.max(np.abs(tm.predict(x, engine="general") - full.predict(x, engine="general"))) < 2e-5


def test_fp32_parity_at_c5_shape(oracle):
    """BASELINE configs[4] model (L=3, H=1024, rank 128) on the FP32 general engine vs the float64 oracle, <= 1e-5."""
    layers, dense = svdlstm.synthetic_layers(16, 1024, 3, seed=0)
    full = svdlstm.full_model_from_weights(layers, dense)
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    tm = svdlstm.truncate_singular_model(sm, 128)
    x = np.random.default_rng(60).standard_normal((3, 24, 16)).astype(np.float32)
    assert_parity(tm.predict(x, engine="general"), oracle_twin(oracle, tm).predict(x), "C5 3F r=128")


def test_derived_models_own_their_weights(oracle, full, x_small):
    """builders hand live tensors / views of the source model to the new cells; every cell must own its buffers
    (Keras variables are independent): set_weights on a derived model leaves the source -- and its siblings -- untouched."""
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    t8, t4 = svdlstm.truncate_singular_model(sm, 8), svdlstm.truncate_singular_model(sm, 4)
    rm = svdlstm.make_LSTM_reduced_model(sm, rank=8)
    snap = {id(mo): ([w.copy() for w in mo.get_weights()], mo.predict(x_small)) for mo in (full, sm, t4, rm)}
    for layer in t8.layers:
        layer.set_weights([w * 0.5 + 0.01 for w in layer.get_weights()])
    sm.layers[0].set_weights([w + 0.125 for w in sm.layers[0].get_weights()])
    for mo in (full, t4, rm):
        w0, y0 = snap[id(mo)]
        for a, b in zip(w0, mo.get_weights()):
            assert np.array_equal(a, b)
        assert np.array_equal(mo.predict(x_small), y0)
    assert not np.array_equal(sm.predict(x_small), snap[id(sm)][1])


def test_realtime_stream_matches_oracle_and_survives_idle(oracle, full):
    """the persistent real-time kernel fed through host-mapped rings.  One sample per call, state on the
    device: the per-sample outputs equal the whole-sequence run (FP32 1e-5 bar vs the float64 oracle); an idle gap longer than
    idle_ms parks the state and the next sample relaunches the kernel without losing it; reset_states / states round-trip;
    a weight update is picked up."""
    import time
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True)
    for label, model in (("full", full), ("3F r8", svdlstm.truncate_singular_model(sm, 8)), ("2F r8", svdlstm.make_LSTM_reduced_model(sm, rank=8)),
                         ("split 3F", svdlstm.make_LSTM_singular_model(full, merged_kernel=False, return_sequences=True))):
        x = np.random.default_rng(70).standard_normal((1, 300, 16)).astype(np.float32)
        y_ref = oracle_twin(oracle, model).predict(x)[0]
        with model.open_stream(idle_ms=100) as st:
            ys = [st.step(x[0, t]) for t in range(100)]
            assert st.kernel_launches() == 1
            time.sleep(0.4)                                    # > idle_ms: the kernel parks its state and leaves
            ys += [st.step(x[0, t]) for t in range(100, 150)]
            assert st.kernel_launches() == 2
            Y, lat = st.run(x[0, 150:], period_us=50.0)         # native paced loop
            assert lat.shape == (150,) and float(np.median(lat)) < 200.0
            hs, cs = st.states()
            got = np.concatenate([np.stack(ys), Y])
            extra = dict(ref32=oracle_twin(oracle, model, np.float32).predict(x)[0], extra_atol=cond_slack(model)) if "2F" in label else {}
            assert_parity(got, y_ref, "real-time stream " + label, **extra)
            # state round trip: restart from the state after 300 samples, replay the last 20 from the state after 280
            st.reset_states(None)
            for t in range(280):
                st.step(x[0, t])
            tail = np.stack([st.step(x[0, t]) for t in range(280, 300)])
            assert np.array_equal(tail, got[280:300])
            hs2, cs2 = st.states()
            assert all(np.array_equal(a, b) for a, b in zip(hs + cs, hs2 + cs2))
            st.reset_states((hs, cs))
            assert all(np.array_equal(a, b) for a, b in zip(hs + cs, sum(st.states(), [])))
    # weights re-bound under a live stream
    m = svdlstm.truncate_singular_model(sm, 8)
    x = np.random.default_rng(71).standard_normal((1, 40, 16)).astype(np.float32)
    with m.open_stream() as st:
        a = np.stack([st.step(x[0, t]) for t in range(20)])
        dk, db = m.layers[-1].get_weights()
        m.layers[-1].set_weights([dk * 2.0, db])
        st.reset_states(None)
        b = np.stack([st.step(x[0, t]) for t in range(20)])
        assert np.allclose(b - db, 2.0 * (a - db), rtol=1e-5, atol=1e-6)
    big, _ = None, None
    layers, dense = svdlstm.synthetic_layers(16, 64, 1, seed=0)
    with pytest.raises((ValueError, RuntimeError), match="wavefront"):
        svdlstm.full_model_from_weights(layers, dense).open_stream()
