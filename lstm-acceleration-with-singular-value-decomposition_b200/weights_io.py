"""Weight / fixture I/O for the hot path: the shipped ``code/model_weights`` CSV layout.

The reference ships its trained full model as per-gate CSVs ``{W,U,b}{i,f,c,o}.csv`` per LSTM layer plus
``dense_top/{weights,bias}.csv``.  The shipped files are TRANSPOSED relative to Keras (``Wi.csv`` is
units x input_dim -- the column-vector convention of code/old_versions/svd_classes.py:125-126), while
code/load_preprocess.py:93-126 writes them un-transposed; ``transposed=`` selects the orientation.
"""
from __future__ import annotations

import os

import numpy as np

GATES = ("i", "f", "c", "o")


def load_model_weights_csv(path, layer_names=None, transposed=True):
    """-> ([(W (D,4H), U (H,4H), b (4H,)), ...], (dense_kernel (H,1), dense_bias (1,))) float32, Keras layout."""
    if layer_names is None:
        layer_names = sorted(d for d in os.listdir(path) if d.startswith("lstm"))
    layers = []
    for name in layer_names:
        d = os.path.join(path, name)
        Ws = [np.atleast_2d(np.loadtxt(os.path.join(d, "W%s.csv" % g), delimiter=",")) for g in GATES]
        Us = [np.atleast_2d(np.loadtxt(os.path.join(d, "U%s.csv" % g), delimiter=",")) for g in GATES]
        bs = [np.loadtxt(os.path.join(d, "b%s.csv" % g), delimiter=",").ravel() for g in GATES]
        if transposed:
            Ws = [w.T for w in Ws]
            Us = [u.T for u in Us]
        layers.append((np.concatenate(Ws, 1).astype(np.float32), np.concatenate(Us, 1).astype(np.float32),
                       np.concatenate(bs).astype(np.float32)))
    dk = np.loadtxt(os.path.join(path, "dense_top", "weights.csv"), delimiter=",").reshape(-1, 1).astype(np.float32)
    db = np.loadtxt(os.path.join(path, "dense_top", "bias.csv"), delimiter=",").reshape(1).astype(np.float32)
    return layers, (dk, db)


def save_model_weights_csv(layers, dense, path, layer_names=None, transposed=True):
    """Inverse of load_model_weights_csv (same file names as the shipped fixture)."""
    os.makedirs(path, exist_ok=True)
    layer_names = layer_names or ["lstm_%d" % i for i in range(len(layers))]
    for name, (W, U, b) in zip(layer_names, layers):
        d = os.path.join(path, name)
        os.makedirs(d, exist_ok=True)
        H = U.shape[0]
        for gi, g in enumerate(GATES):
            w = W[:, gi * H:(gi + 1) * H]
            u = U[:, gi * H:(gi + 1) * H]
            np.savetxt(os.path.join(d, "W%s.csv" % g), w.T if transposed else w, delimiter=",")
            np.savetxt(os.path.join(d, "U%s.csv" % g), u.T if transposed else u, delimiter=",")
            np.savetxt(os.path.join(d, "b%s.csv" % g), b[gi * H:(gi + 1) * H], delimiter=",")
    d = os.path.join(path, "dense_top")
    os.makedirs(d, exist_ok=True)
    np.savetxt(os.path.join(d, "weights.csv"), np.asarray(dense[0]).ravel(), delimiter=",")
    np.savetxt(os.path.join(d, "bias.csv"), np.asarray(dense[1]).ravel(), delimiter=",")


def load_model_weights_npz(path):
    """The same weights packed as one .npz (tests/golden/dropbear_weights.npz): W{i},U{i},b{i}, dense_*."""
    z = np.load(path)
    layers = []
    i = 0
    while "W%d" % i in z:
        layers.append((z["W%d" % i], z["U%d" % i], z["b%d" % i]))
        i += 1
    return layers, (z["dense_kernel"], z["dense_bias"])


def synthetic_layers(D, H, L, seed=0, n_out=1):
    """Synthetic full weights of SURVEY §8(d): Glorot-uniform W, per-gate orthogonal U, zero bias with
    forget-gate ones, Glorot Dense top.  numpy only (host-side data generation)."""
    import math
    rng = np.random.default_rng(seed)
    layers = []
    d_in = D
    for _ in range(L):
        lim = math.sqrt(6.0 / (d_in + 4 * H))
        W = rng.uniform(-lim, lim, size=(d_in, 4 * H))
        Us = []
        for _g in range(4):
            q, r = np.linalg.qr(rng.standard_normal((H, H)))
            Us.append(q * np.sign(np.diag(r)))
        U = np.concatenate(Us, 1)
        b = np.zeros(4 * H)
        b[H:2 * H] = 1.0
        layers.append((W.astype(np.float32), U.astype(np.float32), b.astype(np.float32)))
        d_in = H
    lim = math.sqrt(6.0 / (H + n_out))
    dense = (rng.uniform(-lim, lim, size=(H, n_out)).astype(np.float32), np.zeros(n_out, np.float32))
    return layers, dense
