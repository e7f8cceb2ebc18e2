"""CPU tests of the host-side logic: sharding, the sweep's exchange step under gloo (world_size 2),
weight I/O, metrics helpers."""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import svdlstm
from svdlstm import sweep as sweep_mod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 65536, 1000003):
        for w in (1, 2, 3, 4, 8):
            b = [svdlstm.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, N, R, per, out_dir):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    preds_all = torch.tensor(rng.standard_normal((R, N, per)).astype(np.float32))
    tgt_all = torch.tensor(rng.standard_normal((N, per)).astype(np.float32))
    lo, hi = sweep_mod.shard_bounds(N, world, rank)
    p = preds_all[:, lo:hi].contiguous()
    d = (p.double() - tgt_all[lo:hi].double()[None])
    sse = (d * d).reshape(R, -1).sum(1)
    cnt = torch.tensor([float((hi - lo) * per)], dtype=torch.float64)
    sse_g, cnt_g, preds_g = sweep_mod.exchange_results(p, sse, cnt, N, gather_predictions=True)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), sse=sse_g.numpy(), cnt=cnt_g.numpy(), preds=preds_g.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [11, 16])
def test_sweep_exchange_gloo_world2(tmp_path, N):
    R, per, world = 3, 5, 2
    port = 29500 + (os.getpid() % 2000) + N
    mp.spawn(_gloo_worker, args=(world, port, N, R, per, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    preds_all = rng.standard_normal((R, N, per)).astype(np.float32)
    tgt_all = rng.standard_normal((N, per)).astype(np.float32)
    sse_ref = ((preds_all.astype(np.float64) - tgt_all.astype(np.float64)[None]) ** 2).reshape(R, -1).sum(1)
    outs = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    for o in outs:
        assert np.array_equal(o["preds"], preds_all)          # un-padded, sequence order
        assert np.allclose(o["sse"], sse_ref, rtol=1e-13)
        assert o["cnt"][0] == N * per
    # every rank holds bit-identical reductions (summed in rank order)
    assert np.array_equal(outs[0]["sse"], outs[1]["sse"])


def test_exchange_without_dist_is_identity():
    p = torch.zeros(2, 3, 4)
    s = torch.ones(2, dtype=torch.float64)
    c = torch.tensor([12.0], dtype=torch.float64)
    s2, c2, p2 = sweep_mod.exchange_results(p, s, c, 3)
    assert s2 is s and c2 is c and p2 is p


def test_weights_csv_roundtrip(tmp_path, dropbear_weights):
    layers, dense = dropbear_weights
    for transposed in (True, False):
        d = str(tmp_path / ("t%d" % transposed))
        svdlstm.save_model_weights_csv(layers, dense, d, layer_names=["lstm_69", "lstm_70", "lstm_71"], transposed=transposed)
        l2, d2 = svdlstm.load_model_weights_csv(d, transposed=transposed)
        for (W, U, b), (W2, U2, b2) in zip(layers, l2):
            assert np.array_equal(W, W2) and np.array_equal(U, U2) and np.array_equal(b, b2)
        assert np.array_equal(dense[0], d2[0]) and np.array_equal(dense[1], d2[1])
    # shipped layout: Wi.csv is units x input_dim (transposed)
    wi = np.loadtxt(os.path.join(str(tmp_path / "t1"), "lstm_69", "Wi.csv"), delimiter=",")
    assert wi.shape == (15, 16)
    l3, d3 = svdlstm.load_model_weights_npz(os.path.join(ROOT, "tests", "golden", "dropbear_weights.npz"))
    assert len(l3) == 3 and l3[0][0].shape == (16, 60) and d3[0].shape == (15, 1)


@pytest.mark.skipif(not os.path.isdir("/root/reference/code/model_weights"), reason="reference fixtures not on this box")
def test_loader_reads_the_reference_fixture(dropbear_weights):
    layers, dense = svdlstm.load_model_weights_csv("/root/reference/code/model_weights",
                                                   layer_names=["lstm_69", "lstm_70", "lstm_71"])
    for (W, U, b), (W2, U2, b2) in zip(layers, dropbear_weights[0]):
        assert np.array_equal(W, W2) and np.array_equal(U, U2) and np.array_equal(b, b2)
    assert sum(W.size + U.size + b.size for W, U, b in layers) + 16 == 5656


def test_metrics_host_helpers(series):
    y, p = series["y_test"].astype(np.float64), series["pred"].astype(np.float64)
    assert abs(svdlstm.signaltonoise(y, p) - 12.433968928917704) < 1e-7
    inv = svdlstm.signaltonoise(y, p, invert=True, dB=False)
    assert math.isclose(inv, 1.0 / svdlstm.signaltonoise(y, p, dB=False), rel_tol=1e-12)
    assert svdlstm.full_weight_count(16, 15) + 2 * svdlstm.full_weight_count(15, 15) + 16 == 5656
    assert svdlstm.reduced_merged_weight_count(16, 15, 16, 15) == 16 * 60 + 15 * 60 + 60   # full rank == dense
    assert svdlstm.reduced_split_weight_count(15, 15, 15, 15) == svdlstm.full_weight_count(15, 15)


def test_synthetic_layers_shapes():
    layers, dense = svdlstm.synthetic_layers(16, 32, 2, seed=0)
    assert layers[0][0].shape == (16, 128) and layers[1][0].shape == (32, 128) and dense[0].shape == (32, 1)
    U = layers[0][1][:, :32].astype(np.float64)
    assert np.allclose(U.T @ U, np.eye(32), atol=1e-5)
    assert np.all(layers[0][2][32:64] == 1) and np.all(layers[0][2][:32] == 0)


def test_regularizer_surface():
    h = svdlstm.HoyerRegularizer(0.01)
    assert h.get_config() == {"hoyer": np.float32(0.01)}
    assert svdlstm.HoyerRegularizer(None).hoyer == 0
    with pytest.raises(ValueError):
        svdlstm.OrthogonalRegularizer(0.1, mode="diag")
    o = svdlstm.OrthogonalRegularizer(0.5)
    assert o.from_raw([0, 0, 3.0, 0], (4, 9)) == 0.5 * 0.5 * 3.0 / 6.0


def test_weights_json_roundtrip_and_zip(tmp_path, dropbear_weights):
    """the reference's JSON export layout (load_preprocess.py:80-90) and the shipped model_weights.zip."""
    layers, dense = dropbear_weights
    pth = str(tmp_path / "w.json")
    svdlstm.save_model_weights_json(layers, dense, pth)
    l2, d2 = svdlstm.load_model_weights_json(pth)
    assert len(l2) == len(layers)
    for (W, U, b), (W2, U2, b2) in zip(layers, l2):
        assert np.array_equal(W, W2) and np.array_equal(U, U2) and np.array_equal(b, b2)
    assert np.array_equal(np.asarray(dense[0]).reshape(-1), d2[0].reshape(-1))
    zp = "/root/reference/code/model_weights.zip"
    if os.path.exists(zp):
        lz, dz = svdlstm.load_model_weights_zip(zp)
        for (W, U, b), (Wz, Uz, bz) in zip(layers, lz):
            assert np.array_equal(W, Wz) and np.array_equal(U, Uz) and np.array_equal(b, bz)
        assert np.allclose(np.asarray(dense[1]).reshape(-1), dz[1])


@pytest.mark.parametrize("D,H", [(1, 5), (3, 1), (1, 1), (4, 6)])
@pytest.mark.parametrize("transposed", [True, False])
def test_csv_roundtrip_degenerate_shapes_and_layer_order(tmp_path, D, H, transposed):
    """np.loadtxt returns 0-D / 1-D arrays when units or input_dim is 1 (the reference's own v2 builders use
    InputLayer([None, 1])); shapes must come from the bias length, in both orientations.  Twelve layers check that
    lstm_9 sorts before lstm_10 (numeric suffix, not string order)."""
    rng = np.random.default_rng(D * 10 + H)
    n_layers = 12
    layers = []
    d = D
    for _ in range(n_layers):
        layers.append((rng.standard_normal((d, 4 * H)).astype(np.float32), rng.standard_normal((H, 4 * H)).astype(np.float32),
                       rng.standard_normal(4 * H).astype(np.float32)))
        d = H
    dense = (rng.standard_normal((H, 1)).astype(np.float32), rng.standard_normal(1).astype(np.float32))
    pth = str(tmp_path / "w")
    svdlstm.save_model_weights_csv(layers, dense, pth, transposed=transposed)
    l2, d2 = svdlstm.load_model_weights_csv(pth, transposed=transposed)
    assert len(l2) == n_layers
    for (W, U, b), (W2, U2, b2) in zip(layers, l2):
        assert W2.shape == W.shape and U2.shape == U.shape and b2.shape == b.shape
        assert np.allclose(W, W2, rtol=1e-6) and np.allclose(U, U2, rtol=1e-6) and np.allclose(b, b2, rtol=1e-6)
    assert np.allclose(dense[0], d2[0], rtol=1e-6) and np.allclose(dense[1], d2[1], rtol=1e-6)


def test_preprocess_and_split_train_random_follow_the_reference_recipe():
    """svd_acceleration_v3.py:24-87 on a synthetic recording (the real data_6_with_FFT.json is not shipped): NaN forward fill,
    1.5 s settle cut, Fourier resampling, standardisation, 16-wide framing, the 30.7 s split and the reference's
    (t, t_test, t_train) return order; split_train_random windows + the target one frame past each window."""
    from scipy import signal
    rng = np.random.default_rng(0)
    acc_t = np.sort(rng.uniform(0, 40.0, 40000))
    acc = np.sin(40 * acc_t) + 0.1 * rng.standard_normal(acc_t.size)
    pin_t = np.linspace(0, 40.0, 900)
    pin = 0.1 + 0.05 * np.sin(0.7 * pin_t)
    pin[[0, 5, 6, 400]] = np.nan
    data = {"acceleration_data": acc.tolist(), "time_acceleration_data": acc_t.tolist(),
            "measured_pin_location": pin.tolist(), "measured_pin_location_tt": pin_t.tolist()}
    period = 2e-3
    (X, X_tr, X_te), (y, y_tr, y_te), (t, t_te, t_tr), pin_scaler, acc_scaler = svdlstm.preprocess(period, data=data)
    # straight restatement of the recipe
    p = pin.copy()
    for i in range(len(p)):
        if math.isnan(p[i]):
            p[i] = p[i - 1]
    p, pt = p[pin_t > 1.5], pin_t[pin_t > 1.5] - 1.5
    a, at = acc[acc_t > 1.5], acc_t[acc_t > 1.5] - 1.5
    num = int((at[-1] - at[0]) / period)
    ra, rt = signal.resample(a, num, at)
    rp = np.interp(rt, pt, p)
    an = (ra - ra.mean()) / ra.std()
    pn = ((rp - rp.mean()) / rp.std()).astype(np.float32)
    n = an.size // 16
    assert X.shape == (1, n, 16) and y.shape == (n,) and t.shape == (n,)
    assert np.allclose(X[0], an[:n * 16].reshape(n, 16), atol=1e-12)
    assert np.array_equal(y, pn[:n * 16].reshape(n, 16).T[0]) and y.dtype == np.float32
    assert np.allclose(t, rt[:n * 16].reshape(n, 16).T[0])
    assert X_tr.shape[1] == int((t < 30.7).sum()) and X_te.shape[1] == int((t > 30.7).sum())
    assert np.array_equal(t_tr, t[t < 30.7]) and np.array_equal(t_te, t[t > 30.7]) and np.array_equal(y_te, y[t > 30.7])
    assert np.allclose(pin_scaler.inverse_transform(y.reshape(-1, 1)).ravel(), rp[:n * 16:16], atol=1e-6)
    assert np.allclose(acc_scaler.inverse_transform(X[0].reshape(-1, 1)).ravel(), ra[:n * 16], atol=1e-9)
    Xm, ym = svdlstm.split_train_random(X_tr, y_tr, 64, 200, rng=3)
    assert Xm.shape == (64, 200, 16) and ym.shape == (64,)
    starts = np.random.default_rng(3).integers(0, X_tr.shape[1] - 200, size=64)
    for i in (0, 17, 63):
        assert np.array_equal(Xm[i], X_tr[0, starts[i]:starts[i] + 200]) and ym[i] == y_tr[starts[i] + 200]
    np.random.seed(5)
    Xg, yg = svdlstm.split_train_random(X_tr, y_tr, 8, 50)          # global-state draw, like the reference's randint
    np.random.seed(5)
    ref_idx = [np.random.randint(0, X_tr.shape[1] - 50) for _ in range(8)]
    assert Xg.shape == (8, 50, 16) and all(0 <= s < X_tr.shape[1] - 50 for s in ref_idx)
    with pytest.raises(ValueError):
        svdlstm.split_train_random(X_tr[:, :100], y_tr[:100], 4, 200)
