"""Generates tests/golden/*.npz from the reference's shipped fixtures.  TEST INFRASTRUCTURE.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python oracle/make_golden.py

Inputs  : /root/reference/code/model_weights/** (trained 3x15 DROPBEAR LSTM, CSVs stored
          transposed -- SURVEY fact 8), preprocessed_DROPBEAR_{t,y}.csv, model_prediction.csv.
Outputs : tests/golden/dropbear_weights.npz   W/U/b per layer (Keras layout) + Dense top
          tests/golden/dropbear_series.npz    y[t>30.7], model_prediction, their RMSE/SNR goldens
          tests/golden/kat.npz                known-answer vectors:
              * SURVEY App. D cross-session KAT (sin input, full + truncated y_t)
              * np.linalg.svd of every shipped merged and per-gate matrix (the reference's own
                SVD dependency executed here, svd_classes_v3.py:491,562)
              * reduced-model factors B, C from np.linalg.inv (svd_classes_v3.py:625-626)
              * toy 3x3 rank-2 KAT (old_versions/svd_classes.py:237-249)
No reference *source* is copied; only numeric fixtures derived from its data files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import svdlstm_oracle as O  # noqa: E402

REF = "/root/reference/code"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

SURVEY_KAT_FULL = np.array([0.255471757839, 0.440270522221, 0.607929977257, 0.712479293749,
                            0.736907363072, 0.688360866808, 0.589332776109, 0.469045401480,
                            0.353799635193, 0.258164923839])
SURVEY_KAT_Y9 = {12: 0.458862722403, 8: 0.479061045407, 4: 0.573431658238}


def main():
    os.makedirs(OUT, exist_ok=True)
    layers, dense = O.load_model_weights_csv(os.path.join(REF, "model_weights"),
                                             layer_names=["lstm_69", "lstm_70", "lstm_71"])
    w = {}
    for i, (W, U, b) in enumerate(layers):
        w["W%d" % i], w["U%d" % i], w["b%d" % i] = W, U, b
    w["dense_kernel"], w["dense_bias"] = dense
    np.savez_compressed(os.path.join(OUT, "dropbear_weights.npz"), **w)
    print("weights:", [(W.shape, U.shape, b.shape) for W, U, b in layers], dense[0].shape,
          "params", sum(W.size + U.size + b.size for W, U, b in layers) + dense[0].size + dense[1].size)

    # ---- series fixtures + metric goldens -------------------------------------------------
    t = np.loadtxt(os.path.join(REF, "preprocessed_DROPBEAR_t.csv"))
    y = np.loadtxt(os.path.join(REF, "preprocessed_DROPBEAR_y.csv"))
    pred = np.loadtxt(os.path.join(REF, "model_prediction.csv"))
    y_test = y[t > 30.7]
    assert y_test.shape == pred.shape == (29700,)
    g_rmse = O.rmse(y_test, pred)
    g_snr = O.signaltonoise(y_test, pred)
    print("RMSE", repr(g_rmse), "SNR", repr(g_snr))
    assert abs(g_rmse - 0.20285040751787883) < 1e-15
    assert abs(g_snr - 12.433968928917704) < 1e-12
    np.savez_compressed(os.path.join(OUT, "dropbear_series.npz"),
                        y_test=y_test.astype(np.float32), pred=pred.astype(np.float32),
                        n_total=np.int64(t.size), n_train=np.int64(np.sum(t < 30.7)),
                        dt=np.float64(t[1] - t[0]), rmse=np.float64(g_rmse), snr_db=np.float64(g_snr))
    # float32 storage must not move the goldens beyond 1e-9
    assert abs(O.rmse(y_test.astype(np.float32), pred.astype(np.float32)) - g_rmse) < 1e-9

    # ---- known-answer vectors -------------------------------------------------------------
    kat = {}
    tt = np.arange(10)[:, None]
    jj = np.arange(16)[None, :]
    x = np.sin(0.1 * tt + 0.3 * jj)[None]            # (1,10,16)
    full = O.model_from_weights(layers, dense, dtype=np.float64)
    y_full = full.predict(x)[0, :, 0]
    assert np.max(np.abs(y_full - SURVEY_KAT_FULL)) < 5e-12, y_full
    kat["sin_x"] = x
    kat["sin_y_full"] = y_full
    sm = O.make_LSTM_singular_model(full, merged_kernel=True, return_sequences=True,
                                    svd_dtype=np.float64)
    for r in (15, 12, 8, 4, 1):
        tr = O.truncate_singular_model(sm, r)
        y3 = tr.predict(x)[0, :, 0]
        rm = O.make_LSTM_reduced_model(sm, rank=r, merged_kernel=True)
        y2 = rm.predict(x)[0, :, 0]
        assert np.max(np.abs(y3 - y2)) < 1e-11
        if r in SURVEY_KAT_Y9:
            assert abs(y3[9] - SURVEY_KAT_Y9[r]) < 5e-12, (r, y3[9])
        kat["sin_y_top%d" % r] = y3

    # SVD goldens (float32 input as the reference does; numpy/LAPACK output)
    for i, (W, U, b) in enumerate(layers):
        H = U.shape[0]
        for nm, M in (("W", W), ("U", U)):
            l, s, r = np.linalg.svd(M, full_matrices=False)
            kat["svd_%s%d_s" % (nm, i)] = s
            kat["svd_%s%d_l" % (nm, i)] = l
            kat["svd_%s%d_r" % (nm, i)] = r
            for g in range(4):
                sg = np.linalg.svd(M[:, g * H:(g + 1) * H], compute_uv=False)
                kat["svd_%s%d_g%d_s" % (nm, i, g)] = sg
    # reduced factors of layer 0 at cutoff .05 (keeps everything) and rank 8
    sm32 = O.make_LSTM_singular_model(O.model_from_weights(layers, dense, dtype=np.float32),
                                      merged_kernel=True, dtype=np.float32)
    for tag, kw in (("cut05", dict(cutoff=.05)), ("rank8", dict(rank=8))):
        rm = O.make_LSTM_reduced_model(sm32, merged_kernel=True, dtype=np.float32, **kw)
        wl, wr, ul, ur, _ = rm.cells[0].get_weights()
        kat["red_%s_wB" % tag], kat["red_%s_wC" % tag] = wl, wr
        kat["red_%s_uB" % tag], kat["red_%s_uC" % tag] = ul, ur

    A = np.array([[1, 3, 5], [2, 2, 2], [5, 8, 9]], np.float64)
    A2 = O.reduce_matrix_rank(A, 2)
    assert np.linalg.matrix_rank(A2) == 2
    kat["toy_A"] = A
    kat["toy_A_rank2"] = A2
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **kat)
    print("min sigma over merged matrices:",
          min(kat["svd_%s%d_s" % (nm, i)].min() for nm in "WU" for i in range(3)))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
