"""Weight / fixture I/O for the hot path: the shipped ``code/model_weights`` CSV layout.

The reference ships its trained full model as per-gate CSVs ``{W,U,b}{i,f,c,o}.csv`` per LSTM layer plus
``dense_top/{weights,bias}.csv``.  The shipped files are TRANSPOSED relative to Keras (``Wi.csv`` is
units x input_dim -- the column-vector convention of code/old_versions/svd_classes.py:125-126), while
code/load_preprocess.py:93-126 writes them un-transposed; ``transposed=`` selects the orientation.
"""
from __future__ import annotations

import io
import json
import os
import zipfile

import numpy as np

GATES = ("i", "f", "c", "o")


def _layer_key(name):
    """lstm_9 < lstm_10: order layer directories by their numeric suffix (plain string sort gets this wrong)."""
    digits = "".join(ch for ch in name if ch.isdigit())
    return (int(digits) if digits else -1, name)


def _assemble_layer(read, transposed):
    """One layer from its 12 per-gate arrays.  ``read(fname) -> ndarray`` (any ndim: np.loadtxt returns 0-D / 1-D arrays
    when units or input_dim is 1, in either orientation).  The shapes are fixed explicitly: units from the length of
    b*.csv, the input width from the element count of W*.csv -- never from np.atleast_2d / .T guesses."""
    bs = [np.asarray(read("b%s.csv" % g), np.float64).reshape(-1) for g in GATES]
    H = int(bs[0].size)
    if H < 1 or any(b.size != H for b in bs):
        raise ValueError("per-gate bias files disagree on the number of units: %s" % [int(b.size) for b in bs])

    def mat(fname, rows=None):
        a = np.asarray(read(fname), np.float64)
        if a.size % H:
            raise ValueError("%s holds %d values, not a multiple of units=%d" % (fname, a.size, H))
        d = a.size // H
        if rows is not None and d != rows:
            raise ValueError("%s: expected %d x %d values, found %d" % (fname, rows, H, a.size))
        # file orientation: transposed = (units, d) row-major [the shipped fixture], else (d, units) [load_preprocess.py]
        return a.reshape(H, d).T if transposed else a.reshape(d, H)

    Ws = [mat("W%s.csv" % g) for g in GATES]
    if any(w.shape != Ws[0].shape for w in Ws):
        raise ValueError("per-gate W files disagree on the input width")
    Us = [mat("U%s.csv" % g, rows=H) for g in GATES]
    return (np.concatenate(Ws, 1).astype(np.float32), np.concatenate(Us, 1).astype(np.float32),
            np.concatenate(bs).astype(np.float32))


def load_model_weights_csv(path, layer_names=None, transposed=True):
    """-> ([(W (D,4H), U (H,4H), b (4H,)), ...], (dense_kernel (H,1), dense_bias (1,))) float32, Keras layout."""
    if layer_names is None:
        layer_names = sorted((d for d in os.listdir(path) if d.startswith("lstm")), key=_layer_key)
    layers = []
    for name in layer_names:
        d = os.path.join(path, name)
        layers.append(_assemble_layer(lambda f, d=d: np.loadtxt(os.path.join(d, f), delimiter=","), transposed))
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


def save_model_weights_json(layers, dense, path):
    """code/load_preprocess.py:80-90 layout: {"layer<i>": [get_weights() arrays as nested lists]}, LSTM layers then the Dense top."""
    data = {}
    for i, (W, U, b) in enumerate(layers):
        data["layer%d" % i] = [np.asarray(W).tolist(), np.asarray(U).tolist(), np.asarray(b).tolist()]
    data["layer%d" % len(layers)] = [np.asarray(dense[0]).tolist(), np.asarray(dense[1]).tolist()]
    with open(path, "w", encoding="utf-8") as f:
        json.dump(data, f, ensure_ascii=False, indent=4)


def load_model_weights_json(path):
    """Inverse of save_model_weights_json (and reader of the reference's own JSON export)."""
    with open(path, encoding="utf-8") as f:
        data = json.load(f)
    keys = sorted(data, key=lambda k: int(k[len("layer"):]))
    layers = []
    for k in keys[:-1]:
        W, U, b = (np.asarray(a, np.float32) for a in data[k])
        layers.append((W, U, b))
    dk, db = (np.asarray(a, np.float32) for a in data[keys[-1]])
    return layers, (dk.reshape(-1, dk.shape[-1] if dk.ndim > 1 else 1), db.reshape(-1))


def load_model_weights_zip(path, transposed=True):
    """The shipped code/model_weights.zip: the same per-gate CSV tree as code/model_weights/ (CRLF line ends), read
    without unpacking.  Returns the load_model_weights_csv structure."""
    with zipfile.ZipFile(path) as z:
        names = [n for n in z.namelist() if n.lower().endswith(".csv")]

        def find(layer, fname):
            for n in names:
                parts = n.replace("\\", "/").split("/")
                if len(parts) >= 2 and parts[-2] == layer and parts[-1] == fname:
                    return n
            raise FileNotFoundError("%s/%s not in %s" % (layer, fname, path))

        def read(layer, fname):
            return np.loadtxt(io.StringIO(z.read(find(layer, fname)).decode("utf-8")), delimiter=",")

        layer_names = sorted({n.replace("\\", "/").split("/")[-2] for n in names
                              if n.replace("\\", "/").split("/")[-2].startswith("lstm")}, key=_layer_key)
        layers = [_assemble_layer(lambda f, name=name: read(name, f), transposed) for name in layer_names]
        dk = read("dense_top", "weights.csv").reshape(-1, 1).astype(np.float32)
        db = read("dense_top", "bias.csv").reshape(1).astype(np.float32)
    return layers, (dk, db)
