"""TEST INFRASTRUCTURE ONLY (see oracle/svdlstm_oracle.py): a torch-CPU float64 restatement of the SingularLSTMCell model
with autograd, used by tests/ as the checker of the device training step (K6).  The product never imports it.

Follows code/svd_classes_v3.py:116-236 (cell step, merged and split), :455-462 (HoyerRegularizer) and the Keras 2.10
OrthogonalRegularizer(mode='rows') semantics of SURVEY App. B; loss = mean squared error + sum of regulariser terms
(svd_acceleration_v3.py:111-128).  Parity of this file itself: its forward pass is checked against the numpy oracle in
tests/test_oracle.py::test_torch_ref_matches_numpy_oracle.
"""
import numpy as np
import torch


def _t(a):
    return torch.tensor(np.asarray(a, np.float64), dtype=torch.float64, requires_grad=True)


class TorchSingularModel:
    """weights: list per layer of get_weights() order [sigma_w, sigma_u, w_left, w_right, u_left, u_right, bias]; dense = (kernel, bias)."""

    def __init__(self, layer_weights, dense, units, merged, return_sequences):
        self.layers = [[_t(w) for w in lw] for lw in layer_weights]
        self.dense = [_t(dense[0]), _t(dense[1])]
        self.units = list(units)
        self.merged = bool(merged)
        self.return_sequences = bool(return_sequences)

    def params(self):
        out = []
        for lw in self.layers:
            out += lw
        return out + self.dense

    def _step(self, w, H, x, h, c):
        s_w, s_u, w_l, w_r, u_l, u_r, b = w
        if self.merged:
            z = ((x @ w_l) * s_w) @ w_r + b + ((h @ u_l) * s_u) @ u_r
            zi, zf, zc, zo = torch.split(z, H, dim=1)
        else:
            kw, ku = w_r.shape[0], u_r.shape[0]
            zs = []
            for g in range(4):
                xg = ((x @ w_l[:, g * kw:(g + 1) * kw]) * s_w[:, g * kw:(g + 1) * kw]) @ w_r[:, g * H:(g + 1) * H] + b[g * H:(g + 1) * H]
                rg = ((h @ u_l[:, g * ku:(g + 1) * ku]) * s_u[:, g * ku:(g + 1) * ku]) @ u_r[:, g * H:(g + 1) * H]
                zs.append(xg + rg)
            zi, zf, zc, zo = zs
        i, f, o = torch.sigmoid(zi), torch.sigmoid(zf), torch.sigmoid(zo)
        c = f * c + i * torch.tanh(zc)
        h = o * torch.tanh(c)
        return h, c

    def forward(self, X):
        a = torch.tensor(np.asarray(X, np.float64), dtype=torch.float64)
        B, T = a.shape[0], a.shape[1]
        for w, H in zip(self.layers, self.units):
            h = torch.zeros(B, H, dtype=torch.float64)
            c = torch.zeros(B, H, dtype=torch.float64)
            outs = []
            for t in range(T):
                h, c = self._step(w, H, a[:, t], h, c)
                outs.append(h)
            a = torch.stack(outs, 1)
        y = a @ self.dense[0] + self.dense[1]
        return y if self.return_sequences else y[:, -1]


def hoyer(x, coef):
    return coef * torch.sum(torch.abs(x)) / torch.sum(x * x)


def orthogonal_rows(x, factor):
    xn = x / torch.clamp(torch.sqrt(torch.sum(x * x, dim=1, keepdim=True)), min=1e-6)
    p = xn @ xn.T
    n = x.shape[0]
    pairs = n * (n - 1.0) / 2.0
    off = torch.sum(torch.abs(p * (1.0 - torch.eye(n, dtype=x.dtype))))
    return factor * 0.5 * off / pairs


def loss_and_grads(model, X, y, hoyer_coef=0.0, orth_factor=0.0, train_uv=False):
    """-> (data loss, total loss, grads in model.params() order; non-trainable tensors get zeros)."""
    pred = model.forward(X)
    yt = torch.tensor(np.asarray(y, np.float64), dtype=torch.float64).reshape(pred.shape)
    data = torch.mean((pred - yt) ** 2)
    total = data
    for w in model.layers:
        if hoyer_coef:
            total = total + hoyer(w[0], hoyer_coef) + hoyer(w[1], hoyer_coef)
        if orth_factor and train_uv:
            for k in (2, 3, 4, 5):
                total = total + orthogonal_rows(w[k], orth_factor)
    params = model.params()
    grads = torch.autograd.grad(total, params, allow_unused=True)
    out = []
    n_layers = len(model.layers)
    for idx, (p, g) in enumerate(zip(params, grads)):
        is_dense = idx >= 7 * n_layers
        trainable = is_dense or (idx % 7 < 2) or train_uv
        out.append(g.detach().numpy() if (g is not None and trainable) else np.zeros(tuple(p.shape)))
    return float(data), float(total), out


def keras_adam(params, grads, state, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """One Keras-Adam update on numpy arrays (in place); state = {"t": int, "m": [...], "v": [...]}."""
    state["t"] += 1
    t = state["t"]
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    for i, (p, g) in enumerate(zip(params, grads)):
        state["m"][i] = b1 * state["m"][i] + (1 - b1) * g
        state["v"][i] = b2 * state["v"][i] + (1 - b2) * g * g
        p -= lr_t * state["m"][i] / (np.sqrt(state["v"][i]) + eps)
