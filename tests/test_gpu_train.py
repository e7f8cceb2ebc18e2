"""GPU tests of the device training step (K6): smodel.compile / fit of the reference driver
(code/svd_acceleration_v3.py:111-128) for models of SingularLSTMCells.

Checker = oracle/svdlstm_torch_ref.py (torch-CPU float64 autograd restatement, itself pinned to the numpy oracle in
tests/test_oracle.py).  Tolerance: the device step is FP32 with sequential accumulation over T steps and a batch sum, so
gradients are compared at |g - ref| <= 2e-4 * max|ref| + 1e-7 per tensor (observed ~1e-5); the loss at 1e-5 relative.
"""
import numpy as np
import pytest

import svdlstm
import svdlstm_torch_ref as R

pytestmark = pytest.mark.gpu


def _twin(sm, merged, return_sequences):
    lstms = sm.layers[:-1]
    return R.TorchSingularModel([l.cell.get_weights() for l in lstms], sm.layers[-1].get_weights(), [l.units for l in lstms], merged,
                                return_sequences)


def _cmp(g_dev, g_ref, what):
    assert len(g_dev) == len(g_ref)
    for i, (a, b) in enumerate(zip(g_dev, g_ref)):
        b = np.asarray(b, np.float64).reshape(a.shape)
        tol = 2e-4 * np.abs(b).max() + 1e-7
        assert np.abs(a - b).max() <= tol, "%s: tensor %d max err %.3e (scale %.3e)" % (what, i, np.abs(a - b).max(), np.abs(b).max())


@pytest.mark.parametrize("merged", [True, False])
@pytest.mark.parametrize("return_sequences", [False, True])
def test_gradients_match_autograd(dropbear_weights, merged, return_sequences):
    """d(mse + Hoyer)/d(sigma, Dense) on the shipped model, the configuration the reference fine-tunes (:117-128:
    hoyer=0.01, orthogonal=None; merged_kernel=False there, both here)."""
    layers, dense = dropbear_weights
    full = svdlstm.full_model_from_weights(layers, dense, return_sequences=return_sequences)
    sm = svdlstm.make_LSTM_singular_model(full, hoyer=0.01, orthogonal=None, merged_kernel=merged, return_sequences=return_sequences)
    sm.compile(loss="mse", optimizer="adam")
    rng = np.random.default_rng(0)
    X = rng.standard_normal((6, 40, 16)).astype(np.float32)
    y = rng.standard_normal((6, 40, 1) if return_sequences else (6,)).astype(np.float32)
    loss, grads = sm.gradients(X, y)
    d_ref, t_ref, g_ref = R.loss_and_grads(_twin(sm, merged, return_sequences), X, y, hoyer_coef=0.01)
    assert abs(loss - t_ref) <= 1e-5 * abs(t_ref)
    _cmp(grads, g_ref, "hoyer merged=%s rs=%s" % (merged, return_sequences))
    # non-trainable tensors (factors, bias) get exactly zero
    for li in range(3):
        for w in range(2, 7):
            assert not grads[7 * li + w].any()
    assert abs(sm.evaluate(X, y) - d_ref) <= 1e-5 * d_ref


@pytest.mark.parametrize("merged", [True, False])
def test_gradients_train_uv_with_orthogonal_regularizer(dropbear_weights, merged):
    """train_uv (orthogonal != None, svd_classes_v3.py:505-509, 566-570): factors and bias become trainable and carry the
    Keras OrthogonalRegularizer(mode='rows') penalty."""
    layers, dense = dropbear_weights
    full = svdlstm.full_model_from_weights(layers, dense, return_sequences=False)
    sm = svdlstm.make_LSTM_singular_model(full, hoyer=0.02, orthogonal=0.05, merged_kernel=merged, return_sequences=False)
    # move the factors off exact orthogonality so that the penalty and its gradient are not ~0
    rng = np.random.default_rng(1)
    for l in sm.layers[:-1]:
        l.set_weights([w + (0.05 * rng.standard_normal(w.shape)).astype(np.float32) if i in (2, 3, 4, 5) else w
                       for i, w in enumerate(l.get_weights())])
    sm.compile()
    X = rng.standard_normal((5, 25, 16)).astype(np.float32)
    y = rng.standard_normal((5,)).astype(np.float32)
    loss, grads = sm.gradients(X, y)
    _, t_ref, g_ref = R.loss_and_grads(_twin(sm, merged, False), X, y, hoyer_coef=0.02, orth_factor=0.05, train_uv=True)
    assert abs(loss - t_ref) <= 2e-5 * abs(t_ref)
    _cmp(grads, g_ref, "train_uv merged=%s" % merged)
    loss2, grads2 = sm.gradients(X, y)
    assert loss2 == loss and all(np.array_equal(a, b) for a, b in zip(grads, grads2)), "training step is not deterministic"


@pytest.mark.parametrize("merged", [True, False])
def test_gradients_wide_layers(merged):
    """Layers wider than 32 units take the warp-per-projection backward path of K6 (the 15-unit shipped model takes the
    thread-per-projection one): 40- and 48-unit layers with every tensor trainable, against the autograd twin."""
    rng = np.random.default_rng(7)
    dims = [16, 40, 48]
    layers = [((0.3 * rng.standard_normal((dims[i], 4 * dims[i + 1]))).astype(np.float32),
               (0.3 * rng.standard_normal((dims[i + 1], 4 * dims[i + 1]))).astype(np.float32),
               (0.1 * rng.standard_normal(4 * dims[i + 1])).astype(np.float32)) for i in range(2)]
    dense = ((0.3 * rng.standard_normal((48, 2))).astype(np.float32), np.zeros(2, np.float32))
    full = svdlstm.full_model_from_weights(layers, dense, return_sequences=False)
    sm = svdlstm.make_LSTM_singular_model(full, hoyer=0.02, orthogonal=0.05, merged_kernel=merged, return_sequences=False)
    for l in sm.layers[:-1]:
        l.set_weights([w + (0.02 * rng.standard_normal(w.shape)).astype(np.float32) if i in (2, 3, 4, 5) else w
                       for i, w in enumerate(l.get_weights())])
    sm.compile()
    X = rng.standard_normal((4, 20, 16)).astype(np.float32)
    y = rng.standard_normal((4, 2)).astype(np.float32)
    loss, grads = sm.gradients(X, y)
    _, t_ref, g_ref = R.loss_and_grads(_twin(sm, merged, False), X, y, hoyer_coef=0.02, orth_factor=0.05, train_uv=True)
    assert abs(loss - t_ref) <= 2e-5 * abs(t_ref)
    _cmp(grads, g_ref, "wide merged=%s" % merged)


def test_adam_steps_and_fit(dropbear_weights):
    """Three train_on_batch steps equal three Keras-Adam steps on the autograd gradients (float64 twin); then `fit` runs the
    reference recipe end to end -- compile, fit with Hoyer, rebuild the reduced model with cutoff=.05 -- and the fine-tuned
    weights are what every engine (and a derived model) sees."""
    layers, dense = dropbear_weights
    full = svdlstm.full_model_from_weights(layers, dense, return_sequences=False)
    sm = svdlstm.make_LSTM_singular_model(full, hoyer=0.01, merged_kernel=False)
    sm.compile(loss="mse", optimizer="adam")
    rng = np.random.default_rng(2)
    X = rng.standard_normal((8, 30, 16)).astype(np.float32)
    y = (0.3 * rng.standard_normal((8,))).astype(np.float32)
    twin = _twin(sm, False, False)
    params = [p.detach().numpy() for p in twin.params()]
    state = {"t": 0, "m": [np.zeros_like(p) for p in params], "v": [np.zeros_like(p) for p in params]}
    for step in range(3):
        sm.train_on_batch(X, y)
        _, _, g_ref = R.loss_and_grads(twin, X, y, hoyer_coef=0.01)
        R.keras_adam(params, g_ref, state)         # params alias the twin's tensors: updated in place
    w_dev = sm.get_weights()
    for a, b in zip(w_dev, params):
        assert np.abs(a - b.reshape(a.shape)).max() <= 5e-6 + 1e-4 * np.abs(a - b.reshape(a.shape)).max() + 3e-5, np.abs(a - b.reshape(a.shape)).max()
    # --- the recipe: a target the full model can explain, sparse sigma after the Hoyer fine-tune ---
    Xm = rng.standard_normal((256, 40, 16)).astype(np.float32)
    ym = full.predict(Xm)[:, 0]
    sm2 = svdlstm.make_LSTM_singular_model(full, hoyer=0.05, merged_kernel=False)
    sm2.compile(loss="mse", optimizer="adam", learning_rate=2e-2)
    s_before = np.concatenate([np.abs(w).ravel() for l in sm2.layers[:-1] for w in l.get_weights()[:2]])
    hist = sm2.fit(Xm, ym, batch_size=32, epochs=6, validation_data=(Xm[:64], ym[:64]), seed=0)
    assert len(hist.history["loss"]) == 6 and len(hist.history["val_loss"]) == 6
    assert hist.history["loss"][-1] < hist.history["loss"][0]
    s_after = np.concatenate([np.abs(w).ravel() for l in sm2.layers[:-1] for w in l.get_weights()[:2]])
    assert (s_after < 0.05).sum() > (s_before < 0.05).sum(), "the Hoyer fine-tune did not push any singular value below the cutoff"
    rm = svdlstm.make_LSTM_reduced_model(sm2, cutoff=.05, merged_kernel=False)        # svd_acceleration_v3.py:143: now prunes
    assert svdlstm.count_weights(rm) < svdlstm.count_weights(full)
    y_s, y_r = sm2.predict(Xm[:32]), rm.predict(Xm[:32])[:, -1]
    assert np.isfinite(y_r).all() and np.abs(y_r - y_s).max() < 0.25
    # the in-place update reached the inference engines of this model
    assert np.allclose(sm2.predict(Xm[:8], engine="general"), sm2.predict(Xm[:8], engine="wavefront"), atol=2e-5)
    # the reference's own validation call: a last-step model against the whole target series of ONE long sequence (Keras broadcasts)
    Xl = rng.standard_normal((1, 500, 16)).astype(np.float32)
    yl = rng.standard_normal((1, 500, 1)).astype(np.float32)
    v = sm2.evaluate(Xl, yl)
    assert abs(v - float(np.mean((sm2.predict(Xl).reshape(1, 1, 1) - yl) ** 2))) < 1e-5 * v
    with pytest.raises(RuntimeError, match="compile"):
        svdlstm.make_LSTM_singular_model(full).fit(Xm, ym)
    with pytest.raises(ValueError, match="SingularLSTMCell"):
        full.compile()
