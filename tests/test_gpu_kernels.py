"""GPU parity tests for K2 (Jacobi SVD), K2b (B, C construction), K3 (penalties), K4 (sweep SSE), the
rank sweep and the old explicit-rank API -- each against numpy / the oracle."""
import math

import numpy as np
import pytest
import torch

import svdlstm
from helpers import assert_parity, oracle_twin

pytestmark = pytest.mark.gpu


def _check_svd(A, U, S, Vt, rtol=1e-5):
    A64 = A.astype(np.float64)
    s_ref = np.linalg.svd(A64, compute_uv=False)
    k = min(A.shape)
    assert S.shape == (k,) and U.shape == (A.shape[0], k) and Vt.shape == (k, A.shape[1])
    assert np.all(np.diff(S) <= 0), "singular values must be sorted descending"
    assert np.max(np.abs(S - s_ref) / s_ref) < rtol, np.max(np.abs(S - s_ref) / s_ref)
    rec = (U.astype(np.float64) * S.astype(np.float64)) @ Vt.astype(np.float64)
    assert np.max(np.abs(rec - A64)) < 5e-6 * max(1.0, s_ref[0])
    assert np.max(np.abs(U.T.astype(np.float64) @ U - np.eye(k))) < 5e-6
    assert np.max(np.abs(Vt.astype(np.float64) @ Vt.T - np.eye(k))) < 5e-6


def test_svd_shipped_matrices_vs_numpy(dropbear_weights, kat):
    layers, _ = dropbear_weights
    for i, (W, U_, b) in enumerate(layers):
        for nm, M in (("W", W), ("U", U_)):
            U, S, Vt = (t.cpu().numpy() for t in svdlstm.svd_batched(M))
            _check_svd(M, U, S, Vt)
            # vs the numpy (LAPACK gesdd, float32) goldens = what the reference computes at :562
            assert np.max(np.abs(S - kat["svd_%s%d_s" % (nm, i)]) / kat["svd_%s%d_s" % (nm, i)]) < 1e-5
            l_ref, r_ref = kat["svd_%s%d_l" % (nm, i)], kat["svd_%s%d_r" % (nm, i)]
            sign = np.sign(np.sum(Vt * r_ref, axis=1))         # singular vectors up to sign
            assert np.max(np.abs(Vt * sign[:, None] - r_ref)) < 2e-4
            assert np.max(np.abs(U * sign[None, :] - l_ref)) < 2e-4
            # per-gate blocks, batched in one call (:482-502)
            H = U_.shape[0]
            blocks = np.stack([M[:, g * H:(g + 1) * H] for g in range(4)])
            Sg = svdlstm.svd_batched(blocks, compute_uv=False).cpu().numpy()
            for g in range(4):
                ref = kat["svd_%s%d_g%d_s" % (nm, i, g)]
                assert np.max(np.abs(Sg[g] - ref) / ref) < 1e-5


@pytest.mark.parametrize("shape", [(16, 1024), (60, 16), (64, 64), (1, 7), (5, 1), (256, 1024), (300, 40)])
def test_svd_random_shapes(shape):
    rng = np.random.default_rng(sum(shape))
    A = rng.standard_normal(shape).astype(np.float32)
    (U, S, Vt), sw = svdlstm.svd_batched(A, return_sweeps=True)
    _check_svd(A, U.cpu().numpy(), S.cpu().numpy(), Vt.cpu().numpy())
    assert 1 <= int(sw[0]) <= 40


def test_svd_batched_and_graded_spectrum():
    rng = np.random.default_rng(7)
    # graded spectrum: smallest sigma 1e-4 of the largest must still be accurate to 1e-5 relative
    q1, _ = np.linalg.qr(rng.standard_normal((24, 24)))
    q2, _ = np.linalg.qr(rng.standard_normal((96, 96)))
    s = np.logspace(0, -4, 24)
    A = ((q1 * s) @ q2[:24]).astype(np.float32)
    batch = np.stack([A, 2 * A, A[::-1].copy()])
    U, S, Vt = (t.cpu().numpy() for t in svdlstm.svd_batched(batch))
    for bi in range(3):
        _check_svd(batch[bi], U[bi], S[bi], Vt[bi])
    assert np.allclose(S[1], 2 * S[0], rtol=1e-6)


def test_reduce_factors_vs_oracle(oracle, dropbear_weights, kat):
    W = dropbear_weights[0][0][0]
    l, s, r = np.linalg.svd(W, full_matrices=False)
    for rank in (16, 8, 3, 1):
        B, Cm, pr = svdlstm.reduce_factors(l[:, :rank], s[:rank], r[:rank], return_pivot_ratio=True)
        Bo, Co = oracle._reduce_one(l.astype(np.float64), s.astype(np.float64), r.astype(np.float64), rank=rank)
        assert np.max(np.abs(B.cpu().numpy() - Bo)) < 2e-6
        assert np.max(np.abs(Cm.cpu().numpy() - Co)) < 1e-5 * max(1.0, np.abs(Co).max())
        assert 0 < float(pr[0]) <= 1
    # reference goldens (np.linalg.inv in float32, svd_classes_v3.py:625-626)
    B, Cm = svdlstm.reduce_factors(l[:, :8], s[:8], r[:8])
    assert np.max(np.abs(B.cpu().numpy() - kat["red_rank8_wB"])) < 1e-5
    assert np.max(np.abs(Cm.cpu().numpy() - kat["red_rank8_wC"])) < 2e-4
    # singular V1 is flagged, not silently inverted
    V = np.zeros((2, 5), np.float32)
    V[0, 2] = 1
    V[1, 3] = 1
    _, _, pr = svdlstm.reduce_factors(np.eye(4, 2, dtype=np.float32), np.ones(2, np.float32), V, return_pivot_ratio=True)
    assert float(pr[0]) == 0.0


def test_penalties_vs_oracle(oracle):
    rng = np.random.default_rng(8)
    items = [rng.standard_normal((1, 15)), rng.standard_normal((16, 60)), rng.standard_normal((70, 33)),
             np.linalg.qr(rng.standard_normal((40, 40)))[0], rng.standard_normal((130, 257)), rng.standard_normal((1, 5000))]
    items = [a.astype(np.float32) for a in items]
    wide = rng.standard_normal((40, 1000)).astype(np.float32)     # split-K Gram tiles: 4 feature ranges per tile
    tall = rng.standard_normal((600, 50)).astype(np.float32)      # columns mode with 600 features: 3 ranges
    spec = [(a, a.shape[0] > 1, False) for a in items] + [(items[2], True, True), (wide, True, False), (tall, True, True)]
    raw = svdlstm.evaluate_penalties(spec)
    raw2 = svdlstm.evaluate_penalties(spec)
    assert np.array_equal(raw, raw2), "fixed-order reduction must be bit-reproducible"
    for (a, gram, cols), got in zip(spec, raw):
        ref = oracle.penalty_raw_sums(a, mode="columns" if cols else "rows")
        assert math.isclose(got[0], ref[0], rel_tol=1e-6) and math.isclose(got[1], ref[1], rel_tol=1e-6)
        if gram:
            # GEMM-sized rows-mode items run on the tensor cores (3xTF32, FP32 accumulation that truncates): 2e-5; the rest in float64
            rt = 2e-5 if (not cols and min(a.shape) >= 64 and a.shape[0] ** 2 * a.shape[1] >= 2 ** 21) else 1e-6
            assert math.isclose(got[2], ref[2], rel_tol=rt, abs_tol=1e-4)
            assert math.isclose(got[3], ref[3], rel_tol=rt, abs_tol=1e-6)
    s = items[0]
    assert math.isclose(svdlstm.HoyerRegularizer(0.01)(s), oracle.hoyer_regularizer(s.astype(np.float64), 0.01), rel_tol=1e-5)
    assert math.isclose(svdlstm.HoyerRegularizer(1.0).l1_over_l2(s), oracle.hoyer_l1_over_l2(s), rel_tol=1e-6)
    X = items[1]
    assert math.isclose(svdlstm.OrthogonalRegularizer(0.3)(X), oracle.orthogonal_regularizer_rows(X, 0.3), rel_tol=1e-6)
    assert svdlstm.OrthogonalRegularizer(1.0)(items[3]) < 1e-6
    assert svdlstm.OrthogonalRegularizer(1.0).fro_sq(items[3]) < 1e-9
    with pytest.raises(ValueError):
        svdlstm.OrthogonalRegularizer(0.3)(np.ones(4, np.float32))


def test_model_losses_fused(oracle, dropbear_weights):
    layers, dense = dropbear_weights
    full = svdlstm.full_model_from_weights(layers, dense)
    sm = svdlstm.make_LSTM_singular_model(full, hoyer=0.01, orthogonal=0.1, merged_kernel=True)
    before = svdlstm.launches()
    losses = sm.losses
    assert svdlstm.launches() - before == 1          # ONE fused launch for all 18 regularised weights
    assert len(losses) == 3 * 6
    ref = []
    for layer in sm.layers[:-1]:
        w = layer.get_weights()
        ref += [oracle.hoyer_regularizer(w[0].astype(np.float64), 0.01), oracle.hoyer_regularizer(w[1].astype(np.float64), 0.01)]
        ref += [oracle.orthogonal_regularizer_rows(w[i], 0.1) for i in (2, 3, 4, 5)]
    assert np.allclose(losses, ref, rtol=1e-5, atol=1e-7)
    assert sm.layers[0].cell.train_uv and sm.layers[0].cell.w_left.trainable


def test_sweep_sse_and_metric_goldens(series):
    y, p = series["y_test"], series["pred"]
    assert abs(svdlstm.rmse(y, p) - 0.20285040751787883) < 1e-9
    assert abs(svdlstm.reference_rmse(y, p, y.size) - 0.20285040751787883) < 1e-9
    rng = np.random.default_rng(9)
    pred = rng.standard_normal((7, 100003)).astype(np.float32)
    tgt = rng.standard_normal(100003).astype(np.float32)
    sse = svdlstm.sweep_sse(pred, tgt).cpu().numpy()
    ref = ((pred.astype(np.float64) - tgt.astype(np.float64)) ** 2).sum(1)
    assert np.allclose(sse, ref, rtol=1e-13)
    assert np.array_equal(sse, svdlstm.sweep_sse(pred, tgt).cpu().numpy())


def test_rank_sweep_matches_oracle_and_is_shard_invariant(oracle, dropbear_weights):
    layers, dense = dropbear_weights
    full = svdlstm.full_model_from_weights(layers, dense)
    X = np.random.default_rng(10).standard_normal((13, 40, 16)).astype(np.float32)
    ranks = [16, 11, 6, 2]
    ofull = oracle.model_from_weights(layers, dense)
    osm = oracle.make_LSTM_singular_model(ofull, merged_kernel=True, return_sequences=True, svd_dtype=np.float64)
    y_full = ofull.predict(X)
    # 3-factor sweep: device SVD + truncation + forward + K4 vs the oracle pipeline (np.linalg.svd, float64)
    res3 = svdlstm.rank_sweep(full, X, ranks, form="singular")
    for i, r in enumerate(ranks):
        y3 = oracle.truncate_singular_model(osm, r).predict(X)
        assert abs(res3["rmse"][i] - oracle.rmse(y_full, y3)) < 2e-5
        assert np.max(np.abs(res3["preds"][i].cpu().numpy() - y3[..., 0])) < 5e-5, r
    # 2-factor sweep: RMSE vs the oracle pipeline; predictions vs the oracle on the SAME (B, C) factors
    # with the float32-conditioning-aware tolerance (C = inv(V1) V2 is ill-conditioned at some ranks)
    _, models = svdlstm.build_rank_models(full, ranks)
    res = svdlstm.rank_sweep(full, X, ranks, form="reduced", models=models)
    for i, r in enumerate(ranks):
        yr = oracle.make_LSTM_reduced_model(osm, rank=r).predict(X)
        assert abs(res["rmse"][i] - oracle.rmse(y_full, yr)) < 1e-4
        assert_parity(res["preds"][i].cpu().numpy()[..., None], oracle_twin(oracle, models[i]).predict(X), "sweep 2F r=%d" % r,
                      ref32=oracle_twin(oracle, models[i], np.float32).predict(X))
    assert res["rmse"][0] < 1e-5 and res["rmse"][-1] > res["rmse"][1]
    # "virtual ranks": any partition of the sequences gives bit-identical per-(rank, sequence) outputs
    for world in (2, 4):
        parts = []
        for vr in range(world):
            lo, hi = svdlstm.shard_bounds(13, world, vr)
            parts.append(svdlstm.rank_sweep(full, X[lo:hi], ranks, models=models)["preds"])
        assert torch.equal(torch.cat(parts, 1), svdlstm.rank_sweep(full, X, ranks, models=models)["preds"])


def test_old_explicit_rank_api(oracle, dropbear_weights, kat):
    A2 = svdlstm.reduce_matrix_rank(kat["toy_A"], 2)
    assert np.linalg.matrix_rank(A2.astype(np.float64), tol=1e-4) == 2
    assert np.max(np.abs(A2 - kat["toy_A_rank2"])) < 5e-6
    M1, M2 = svdlstm.reduce_two_step(kat["toy_A"], 2)
    O1, O2 = oracle.reduce_two_step(kat["toy_A"], 2)
    assert np.max(np.abs(M1 - O1)) < 1e-5 and np.max(np.abs(M2 - O2)) < 1e-5
    layers, dense = dropbear_weights
    sub = [layers[1], layers[2]]                       # square layers (the old API "assumes a square model")
    full = svdlstm.full_model_from_weights(sub, dense)
    sv = svdlstm.get_model_singular_values(full)
    ref = oracle.get_model_singular_values(oracle.model_from_weights(sub, dense))
    assert sv.shape == (2, 2, 4, 15) and np.max(np.abs(sv - ref) / ref) < 1e-5
    x = np.random.default_rng(11).standard_normal((2, 30, 15)).astype(np.float32)
    svdlstm.set_model_matrix_rank(full, (0, 1, 2), 9)
    om = oracle.model_from_weights([tuple(a.copy() for a in l) for l in sub], dense)
    oracle.set_model_matrix_rank(om, (0, 1, 2), 9)
    assert np.max(np.abs(full.predict(x) - om.predict(x))) < 2e-5
    assert np.linalg.matrix_rank(full.layers[0].get_weights()[1][:, 30:45].astype(np.float64), tol=1e-4) == 9


def test_greedy_sigma_sweep_matches_oracle(oracle, dropbear_weights):
    """LSTM_wrapper.iterate_reduce_model (old_versions/svd_classes.py:139-182; loop of
    old_versions/svd_acceleration.py:61-88) on device == the oracle's restatement, iteration by iteration."""
    layers, dense = dropbear_weights
    sub = [layers[1], layers[2]]                       # square layers, as the old sweep assumes
    x = np.random.default_rng(12).standard_normal((3, 40, 15)).astype(np.float32)
    om_full = oracle.model_from_weights([tuple(a.copy() for a in l) for l in sub], dense)
    y = om_full.predict(x)                             # target = full-model output (RMSE ratio plot semantics)
    n_it = 12
    ref_rmse, ref_w = oracle.greedy_sigma_sweep(om_full, x, y, n_it)
    wrap = svdlstm.LSTM_wrapper(svdlstm.full_model_from_weights(sub, dense))
    rmse, w, done = wrap.iterate_reduce_model(x, y.astype(np.float32), reductions=n_it)
    assert done == n_it and rmse.shape == (n_it,)
    assert rmse[0] < 1e-6                              # nothing removed yet: the model reproduces its own output
    assert np.array_equal(w, ref_w)
    assert np.max(np.abs(rmse - ref_rmse)) < 2e-5 + 1e-4 * np.max(ref_rmse)
    assert np.all(np.diff(w) > 0)
    order = svdlstm.sorted_sigma_indices(wrap.model_singular_values)
    assert order.shape == (2 * 2 * 4 * 15, 4)
    sv = wrap.model_singular_values
    assert np.all(np.diff(sv[tuple(order.T)]) >= 0)    # ascending global order
    assert int(wrap.model_ranks.sum()) == 2 * 2 * 4 * 15 - n_it


def test_weight_counts_device_models(dropbear_weights):
    layers, dense = dropbear_weights
    full = svdlstm.full_model_from_weights(layers, dense)
    assert svdlstm.count_weights(full) == 5656 == full.count_params()
    sm = svdlstm.make_LSTM_singular_model(full, merged_kernel=False)
    rm = svdlstm.make_LSTM_reduced_model(sm, merged_kernel=False, rank=8)
    expect = sum(svdlstm.reduced_split_weight_count(W.shape[0], 15, 8, 8) for W, _, _ in layers) + 16
    assert svdlstm.count_weights(rm) == expect
    assert rm._fused_handle().count_weights() == expect - 16
    assert 0 < svdlstm.weight_reduction_percent(full, rm) < 100


def test_svd_c5_size_matrix():
    """BASELINE configs[4] factor size: the merged recurrent matrix of an H=1024 layer (1024 x 4096), large Jacobi path.
    Singular values against LAPACK (float64) to 1e-5 relative including the smallest; reconstruction and orthogonality."""
    import time
    rng = np.random.default_rng(1024)
    A = (rng.standard_normal((1024, 4096)) / 64.0).astype(np.float32)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    (U, S, Vt), sw = svdlstm.svd_batched(A, return_sweeps=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    s_ref = np.linalg.svd(A.astype(np.float64), compute_uv=False)
    S = S.cpu().numpy().astype(np.float64).reshape(-1)
    assert np.max(np.abs(S - s_ref) / s_ref) < 1e-5
    Ud, Vd = U.double(), Vt.double()
    rec = (Ud * torch.from_numpy(S).cuda()) @ Vd
    assert float((rec - torch.from_numpy(A.astype(np.float64)).cuda()).abs().max()) < 2e-5
    eye = torch.eye(1024, dtype=torch.float64, device="cuda")
    assert float((Ud.reshape(1024, 1024).T @ Ud.reshape(1024, 1024) - eye).abs().max()) < 1e-5
    assert float((Vd.reshape(1024, 4096) @ Vd.reshape(1024, 4096).T - eye).abs().max()) < 1e-5
    print("SVD 1024x4096: %d sweeps, %.2f s" % (int(sw[0]), dt))


def test_penalties_tensor_core_gram(oracle, monkeypatch):
    """K3 kind 2: 128 x 128 Gram tiles on tcgen05 (kind::tf32, 3xTF32 split, FP32 accumulation in TMEM) for GEMM-sized rows-mode
    items -- the C5 factor shapes (128 x 4096 right factors: 8 feature splits; 1024 x 128 left factors: 36 tiles), ragged
    shapes, and orthonormal rows (penalty ~ 0).  Measured against the float64 oracle: 2e-6 .. 7e-6 relative, always LOW (the
    tensor core's FP32 accumulator truncates), ~5e-8 absolute per entry of the normalised Gram matrix.  Bars: 2e-5 relative,
    1e-7 x pairs absolute on the off-diagonal sum, 1e-8 absolute on ||X X^T - I||_F^2 of orthonormal rows."""
    rng = np.random.default_rng(20)
    q = np.linalg.qr(rng.standard_normal((256, 256)))[0].astype(np.float32)
    items = [(rng.standard_normal((128, 4096)) / 64.0).astype(np.float32), (rng.standard_normal((1024, 128)) / 11.0).astype(np.float32),
             rng.standard_normal((300, 700)).astype(np.float32), q, q[:100, :].copy(), rng.standard_normal((130, 257)).astype(np.float32)]
    spec = [(a, True, False) for a in items]
    raw = svdlstm.evaluate_penalties(spec)
    assert np.array_equal(raw, svdlstm.evaluate_penalties(spec)), "tensor-core Gram tiles must be bit-reproducible"
    for a, got in zip(items, raw):
        ref = oracle.penalty_raw_sums(a, mode="rows")
        pairs = a.shape[0] * (a.shape[0] - 1)
        assert math.isclose(got[0], ref[0], rel_tol=1e-6) and math.isclose(got[1], ref[1], rel_tol=1e-6)
        assert math.isclose(got[2], ref[2], rel_tol=2e-5, abs_tol=1e-7 * pairs), (a.shape, got[2], ref[2])
        assert math.isclose(got[3], ref[3], rel_tol=2e-5, abs_tol=1e-8), (a.shape, got[3], ref[3])
    monkeypatch.setenv("SVDLSTM_K3_NO_TC", "1")          # the float64 CUDA-core tiles (kept for columns mode and small items)
    raw64 = svdlstm.evaluate_penalties(spec)
    monkeypatch.delenv("SVDLSTM_K3_NO_TC")
    for a, got in zip(items, raw64):
        ref = oracle.penalty_raw_sums(a, mode="rows")
        assert math.isclose(got[2], ref[2], rel_tol=1e-6, abs_tol=1e-4) and math.isclose(got[3], ref[3], rel_tol=1e-6, abs_tol=1e-9)
