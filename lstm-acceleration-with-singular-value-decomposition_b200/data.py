"""Host-side data preparation in front of the hot path: the DROPBEAR preprocessing and the random
training windows of the reference driver.

    preprocess(sampling_period)                              code/svd_acceleration_v3.py:24-80
    split_train_random(X_train, y_train, batch_size, train_len)   code/svd_acceleration_v3.py:82-87

Same names, argument order and return structure as the reference.  This is data plumbing on the host (numpy / scipy FFT
resampling), exactly where the reference does it; nothing of the LSTM path runs here.  The raw recording
``data_6_with_FFT.json`` is not shipped with the reference (.MISSING_LARGE_BLOBS), so ``preprocess`` also accepts the
already-parsed dict through ``data=`` -- which is how the tests drive it with a synthetic recording.
"""
from __future__ import annotations

import json

import numpy as np

FRAME = 16            # samples of acceleration per model input row (svd_acceleration_v3.py:61)
T_SETTLE = 1.5        # seconds dropped at the start of the recording (:45-48)
T_SPLIT = 30.7        # train / test boundary in seconds (:69-75)


class StandardScaler:
    """The three methods of sklearn.preprocessing.StandardScaler the reference uses (:55-60, :172-174): population
    standard deviation (ddof = 0), column-wise, zero variance scaled by 1."""

    def fit(self, a):
        a = np.asarray(a, np.float64)
        self.mean_ = a.mean(axis=0)
        self.var_ = a.var(axis=0)
        self.scale_ = np.where(self.var_ > 0, np.sqrt(self.var_), 1.0)
        return self

    def transform(self, a):
        return (np.asarray(a, np.float64) - self.mean_) / self.scale_

    def fit_transform(self, a):
        return self.fit(a).transform(a)

    def inverse_transform(self, a):
        return np.asarray(a, np.float64) * self.scale_ + self.mean_


def _forward_fill_nan(v):
    """Each NaN takes the value of the sample before it (:40-43; a leading NaN wraps to the last sample, as v[-1] does)."""
    v = np.array(v, dtype=np.float64)
    bad = np.isnan(v)
    if not bad.any():
        return v
    idx = np.where(bad, 0, np.arange(v.size))
    np.maximum.accumulate(idx, out=idx)
    filled = v[idx]
    if bad[0]:                                  # python's pin[-1]: the (already clean or not) last element
        lead = np.argmin(bad)                   # first valid index
        filled[:lead] = v[-1]
    return filled


def preprocess(sampling_period, path="data_6_with_FFT.json", data=None):
    """-> (X, X_train, X_test), (y, y_train, y_test), (t, t_test, t_train), pin_scaler, acc_scaler  -- the reference's tuple
    order, including its (t, t_test, t_train) quirk (:80).  X is (1, n_frames, 16); y, t are per frame."""
    from scipy import signal
    if data is None:
        with open(path) as f:
            data = json.load(f)
    acc = np.asarray(data["acceleration_data"], np.float64)
    acc_t = np.asarray(data["time_acceleration_data"], np.float64)
    pin = _forward_fill_nan(data["measured_pin_location"])
    pin_t = np.asarray(data["measured_pin_location_tt"], np.float64)

    keep_p, keep_a = pin_t > T_SETTLE, acc_t > T_SETTLE
    pin, pin_t = pin[keep_p], pin_t[keep_p] - T_SETTLE
    acc, acc_t = acc[keep_a], acc_t[keep_a] - T_SETTLE
    num = int((acc_t[-1] - acc_t[0]) / sampling_period)
    acc_rs, t_rs = signal.resample(acc, num, acc_t)                 # Fourier resampling onto `num` uniform samples
    pin_rs = np.interp(t_rs, pin_t, pin)

    acc_scaler, pin_scaler = StandardScaler(), StandardScaler()
    acc_n = acc_scaler.fit_transform(acc_rs.reshape(-1, 1)).ravel()
    pin_n = pin_scaler.fit_transform(pin_rs.reshape(-1, 1)).ravel().astype(np.float32)

    n_frames = acc_n.size // FRAME
    X = acc_n[:n_frames * FRAME].reshape(n_frames, FRAME)[None]     # (1, n_frames, 16)
    t = t_rs[:n_frames * FRAME:FRAME]                               # first sample time of every frame
    y = pin_n[:n_frames * FRAME:FRAME]
    tr, te = t < T_SPLIT, t > T_SPLIT
    return (X, X[:, tr], X[:, te]), (y, y[tr], y[te]), (t, t[te], t[tr]), pin_scaler, acc_scaler


def split_train_random(X_train, y_train, batch_size, train_len, rng=None):
    """``batch_size`` random windows of ``train_len`` frames from the one long training run and, for each, the target
    ONE frame past its end (:82-87).  ``rng`` (np.random.Generator or seed) makes the draw reproducible; the default draws
    from numpy's global state like the reference's ``randint``."""
    X_train = np.asarray(X_train)
    y_train = np.asarray(y_train)
    run = X_train.shape[1]
    if run - train_len <= 0 or y_train.shape[0] < run:
        raise ValueError("need more than train_len=%d frames (have %d) and one target per frame" % (train_len, run))
    if rng is None:
        starts = np.random.randint(0, run - train_len, size=batch_size)
    else:
        starts = np.random.default_rng(rng).integers(0, run - train_len, size=batch_size)
    win = starts[:, None] + np.arange(train_len)[None, :]
    return np.ascontiguousarray(X_train[0][win]), np.ascontiguousarray(y_train[starts + train_len])
