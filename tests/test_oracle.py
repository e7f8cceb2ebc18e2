"""CPU tests that pin the oracle (oracle/svdlstm_oracle.py) -- the checker the GPU parity tests trust.

Pins: algebraic identities, an independent torch.nn.LSTM, the cross-session KAT of SURVEY
App. D, numpy-SVD goldens of the shipped matrices, fixture-derived RMSE/SNR goldens, slide-9 weight
counts, the toy 3x3 KAT of old_versions/svd_classes.py:237-249."""
import math

import numpy as np
import pytest
import torch


def test_kat_full_model_matches_survey(oracle, dropbear_weights, kat):
    layers, dense = dropbear_weights
    m = oracle.model_from_weights(layers, dense, dtype=np.float64)
    y = m.predict(kat["sin_x"])[0, :, 0]
    assert np.max(np.abs(y - kat["sin_y_full"])) < 1e-13
    # literal values recorded in SURVEY.md App. D
    assert abs(y[0] - 0.255471757839) < 5e-12 and abs(y[9] - 0.258164923839) < 5e-12


def test_torch_lstm_cross_check(oracle, dropbear_weights):
    layers, dense = dropbear_weights
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 40, 16))
    y_or = oracle.model_from_weights(layers, dense, dtype=np.float64).predict(x)
    a = torch.tensor(x, dtype=torch.float64)
    for W, U, b in layers:
        H = U.shape[0]
        lstm = torch.nn.LSTM(W.shape[0], H, batch_first=True).double()
        with torch.no_grad():
            # torch gate order i,f,g,o == Keras i,f,c,o
            lstm.weight_ih_l0.copy_(torch.tensor(W.T, dtype=torch.float64))
            lstm.weight_hh_l0.copy_(torch.tensor(U.T, dtype=torch.float64))
            lstm.bias_ih_l0.copy_(torch.tensor(b, dtype=torch.float64))
            lstm.bias_hh_l0.zero_()
            a, _ = lstm(a)
    y_t = (a @ torch.tensor(dense[0], dtype=torch.float64) + torch.tensor(dense[1], dtype=torch.float64)).numpy()
    assert np.max(np.abs(y_t - y_or)) < 1e-13


@pytest.mark.parametrize("merged", [True, False])
def test_identities_full_singular_reduced(oracle, dropbear_weights, merged):
    layers, dense = dropbear_weights
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 60, 16))
    full = oracle.model_from_weights(layers, dense, dtype=np.float64)
    y = full.predict(x)
    sm = oracle.make_LSTM_singular_model(full, merged_kernel=merged, return_sequences=True, svd_dtype=np.float64)
    ys = sm.predict(x)
    assert np.max(np.abs(ys - y)) < 5e-13
    # cutoff .05 prunes nothing for the MERGED matrices (min sigma 0.73); per-gate blocks do have
    # sigma < .05, so the split identity is checked at cutoff 0
    rm = oracle.make_LSTM_reduced_model(sm, cutoff=.05 if merged else 0.0, merged_kernel=merged)
    yr = rm.predict(x)
    assert np.max(np.abs(yr - y)) < 1e-10
    if not merged:
        pruned = oracle.make_LSTM_reduced_model(sm, cutoff=.05, merged_kernel=False)
        assert oracle.count_weights(pruned) < oracle.count_weights(rm)
    # weight ordering contracts (svd_classes_v3.py:113, :278, :308-315)
    ws = sm.cells[0].get_weights()
    assert [w.shape for w in ws][:2] == ([(1, 16), (1, 15)] if merged else [(1, 60), (1, 60)])
    assert len(rm.cells[0].get_weights()) == (5 if merged else 17)


def test_truncation_kat(oracle, dropbear_weights, kat):
    layers, dense = dropbear_weights
    full = oracle.model_from_weights(layers, dense, dtype=np.float64)
    sm = oracle.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True, svd_dtype=np.float64)
    for r, y9 in ((12, 0.458862722403), (8, 0.479061045407), (4, 0.573431658238)):
        y3 = oracle.truncate_singular_model(sm, r).predict(kat["sin_x"])[0, :, 0]
        y2 = oracle.make_LSTM_reduced_model(sm, rank=r).predict(kat["sin_x"])[0, :, 0]
        assert abs(y3[9] - y9) < 5e-12
        assert np.max(np.abs(y3 - y2)) < 1e-11
        assert np.max(np.abs(y3 - kat["sin_y_top%d" % r])) < 1e-12


def test_svd_goldens(dropbear_weights, kat):
    layers, _ = dropbear_weights
    mins = []
    for i, (W, U, b) in enumerate(layers):
        for nm, M in (("W", W), ("U", U)):
            s = np.linalg.svd(M.astype(np.float64), compute_uv=False)
            assert np.allclose(s, kat["svd_%s%d_s" % (nm, i)], rtol=2e-6)
            mins.append(s.min())
    assert abs(min(mins) - 0.730) < 1e-3            # cutoff .05 prunes nothing
    assert abs(kat["svd_W0_s"][0] - 4.337) < 1e-3


def test_metric_goldens(oracle, series):
    y, p = series["y_test"].astype(np.float64), series["pred"].astype(np.float64)
    assert y.shape == (29700,)
    assert abs(oracle.rmse(y, p) - 0.20285040751787883) < 1e-9
    assert abs(oracle.signaltonoise(y, p) - 12.433968928917704) < 1e-7
    assert abs(float(series["rmse"]) - 0.20285040751787883) < 1e-15
    assert int(series["n_total"]) == 106450 and int(series["n_train"]) == 76750
    # reference divisor quirk: sum over all y / len(y_test)
    assert abs(oracle.reference_rmse(y, p, y.size) - oracle.rmse(y, p)) < 1e-15


def test_weight_counts(oracle, dropbear_weights):
    layers, dense = dropbear_weights
    full = oracle.model_from_weights(layers, dense)
    assert oracle.count_weights(full) == 5656
    assert sum(oracle.full_weight_count(W.shape[0], U.shape[0]) for W, U, _ in layers) == 5640
    sm = oracle.make_LSTM_singular_model(full, merged_kernel=False, svd_dtype=np.float64)
    for r in (15, 8, 3):
        rm = oracle.make_LSTM_reduced_model(sm, rank=r, merged_kernel=False)
        expect = sum(oracle.reduced_split_weight_count(W.shape[0], 15, min(r, 15), min(r, 15)) for W, _, _ in layers) + 16
        assert oracle.count_weights(rm) == expect
    smm = oracle.make_LSTM_singular_model(full, merged_kernel=True, svd_dtype=np.float64)
    rm = oracle.make_LSTM_reduced_model(smm, rank=8, merged_kernel=True)
    expect = sum(oracle.reduced_merged_weight_count(W.shape[0], 15, 8, 8) for W, _, _ in layers) + 16
    assert oracle.count_weights(rm) == expect


def test_toy_rank_reduction(oracle, kat):
    A2 = oracle.reduce_matrix_rank(kat["toy_A"], 2)
    assert np.linalg.matrix_rank(A2) == 2
    assert np.allclose(A2, kat["toy_A_rank2"], atol=1e-12)
    M1, M2 = oracle.reduce_two_step(kat["toy_A"], 2)
    x = np.array([[0.3], [-1.0], [2.0]])
    y = np.concatenate([M1 @ x, M2 @ (M1 @ x)])
    assert np.allclose(y, A2 @ x, atol=1e-10)


def test_regularizers(oracle):
    rng = np.random.default_rng(3)
    s = rng.standard_normal((1, 15))
    assert math.isclose(oracle.hoyer_regularizer(s, 0.01), 0.01 * np.abs(s).sum() / (s ** 2).sum())
    q, _ = np.linalg.qr(rng.standard_normal((12, 12)))
    assert oracle.orthogonal_regularizer_rows(q, 0.5) < 1e-14
    assert oracle.orthogonality_fro_sq(q) < 1e-25
    X = rng.standard_normal((6, 20))
    raw = oracle.penalty_raw_sums(X)
    assert math.isclose(raw[0], np.abs(X).sum()) and math.isclose(raw[1], (X ** 2).sum())
    assert math.isclose(oracle.orthogonal_regularizer_rows(X, 2.0), 2.0 * 0.5 * raw[2] / 15.0)
    assert math.isclose(raw[3], oracle.orthogonality_fro_sq(X))


def test_rnn_loop_semantics(oracle, dropbear_weights):
    (W, U, b) = dropbear_weights[0][0]
    cell = oracle.FullCell(15, W, U, b)
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3, 12, 16))
    seq, h, c = oracle.rnn_layer(cell, x, return_sequences=True, return_state=True)
    assert np.allclose(seq[:, -1], h)
    # go_backwards == forward on the time-reversed input, outputs in processing order
    sb = oracle.rnn_layer(cell, x, go_backwards=True, return_sequences=True)
    assert np.allclose(sb, oracle.rnn_layer(cell, x[:, ::-1], return_sequences=True))
    # chunked streaming with carried state == one pass (stateful semantics, svd_classes_v3.py:421-426)
    s1, h1, c1 = oracle.rnn_layer(cell, x[:, :5], return_sequences=True, return_state=True)
    s2 = oracle.rnn_layer(cell, x[:, 5:], initial_state=[h1, c1], return_sequences=True)
    assert np.allclose(np.concatenate([s1, s2], 1), seq)
    # mask: masked steps carry state and repeat the previous output
    mask = np.ones((3, 12), bool)
    mask[1, 4:7] = False
    sm = oracle.rnn_layer(cell, x, mask=mask, return_sequences=True)
    assert np.allclose(sm[1, 4], sm[1, 3]) and np.allclose(sm[1, 6], sm[1, 3])
    assert np.allclose(sm[0], seq[0])
    x_skip = np.delete(x[1:2], [4, 5, 6], axis=1)
    assert np.allclose(oracle.rnn_layer(cell, x_skip, return_sequences=True)[0, 4:], sm[1, 7:])


def test_oracle_float32_close_to_float64(oracle, dropbear_weights):
    layers, dense = dropbear_weights
    x = np.random.default_rng(5).standard_normal((1, 300, 16))
    y64 = oracle.model_from_weights(layers, dense, dtype=np.float64).predict(x)
    y32 = oracle.model_from_weights(layers, dense, dtype=np.float32).predict(x)
    assert np.max(np.abs(y64 - y32)) < 2e-5


def test_greedy_sigma_sweep_oracle(oracle, dropbear_weights):
    """old_versions/svd_acceleration.py:61-88: the first evaluation happens before anything is removed (RMSE 0 against
    the model's own output), weights eliminated follow 2n-2r-1 (:86), and the error grows as sigmas are dropped."""
    layers, dense = dropbear_weights
    sub = [tuple(a.copy() for a in l) for l in (layers[1], layers[2])]
    om = oracle.model_from_weights(sub, dense)
    x = np.random.default_rng(3).standard_normal((2, 25, 15)).astype(np.float32)
    y = om.predict(x)
    rmse, w = oracle.greedy_sigma_sweep(om, x, y, 8)
    assert rmse[0] < 1e-12 and rmse[-1] > rmse[1] > 0
    assert list(w[:3]) == [0.0, 2 * 15 - 2 * 14 - 1, (2 * 15 - 2 * 14 - 1) * 2] or w[2] in (2.0, 4.0)
    assert np.all(np.diff(w) > 0)


def test_torch_ref_matches_numpy_oracle(oracle, dropbear_weights):
    """The torch-autograd checker of the training step (oracle/svdlstm_torch_ref.py) is itself pinned: its forward pass equals
    the numpy oracle's on the shipped model, merged and split."""
    import svdlstm_torch_ref as R
    layers, dense = dropbear_weights
    ofull = oracle.model_from_weights(layers, dense, dtype=np.float64)
    x = np.random.default_rng(3).standard_normal((3, 12, 16))
    for merged in (True, False):
        osm = oracle.make_LSTM_singular_model(ofull, merged_kernel=merged, return_sequences=True, svd_dtype=np.float64)
        tm = R.TorchSingularModel([c.get_weights() for c in osm.cells], osm.dense, [c.units for c in osm.cells], merged, True)
        assert np.max(np.abs(tm.forward(x).detach().numpy() - osm.predict(x))) < 1e-12
