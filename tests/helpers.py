"""Shared helpers of the GPU parity tests."""
import numpy as np

import svdlstm

RTOL, ATOL = 1e-5, 2e-6


def cond_slack(model):
    """Extra absolute tolerance for 2-factor models: 4e-7 * max|C| over the right factors.  z = [a | a C]
    sums terms of magnitude |a||C| that cancel to O(1); float32 rounding of that sum is ~eps*sqrt(K)*max|C|
    whatever the summation order (numpy/TF or ours)."""
    mx = 0.0
    for layer in model.layers[:-1]:
        c = layer.cell
        if isinstance(c, svdlstm.ReducedLSTMCell):
            for v in c.weights:
                if "right" in v.name and v.tensor.numel():
                    mx = max(mx, float(v.tensor.abs().max()))
    return 4e-7 * mx


def assert_parity(y, ref, what="", ref32=None, extra_atol=0.0):
    """ref = float64 oracle.  For the 2-factor form C = inv(V1) V2 has entries up to ~650 on the shipped
    weights (cond(V1) up to 136), so ANY float32 evaluation -- the reference's TF float32 included -- sits
    ~1e-5 from the float64 value.  There the bar is: as close to float64 as the reference-precision
    arithmetic itself (float32 oracle, ref32) is, times 5, on top of the base tolerance."""
    y = np.asarray(y, np.float64)
    ref = np.asarray(ref, np.float64)
    assert y.shape == ref.shape, (what, y.shape, ref.shape)
    err = np.abs(y - ref)
    tol = RTOL * np.abs(ref) + ATOL + extra_atol
    if ref32 is not None:
        tol = tol + 5.0 * np.max(np.abs(np.asarray(ref32, np.float64) - ref))
    worst = np.max(err - tol)
    assert worst <= 0, "%s: max abs err %.3e (rel %.3e) exceeds tolerance" % (
        what, err.max(), (err / (np.abs(ref) + 1e-6)).max())


def oracle_twin(oracle, model, dtype=np.float64):
    """Oracle model holding exactly the device model's weights (get_weights() orderings are the contract)."""
    cells = []
    lstms = model.layers[:-1]
    for layer in lstms:
        c = layer.cell
        w = c.get_weights()
        if isinstance(c, svdlstm.SingularLSTMCell):
            cells.append(oracle.SingularCell(c.units, [w[2], w[0], w[3]], [w[4], w[1], w[5]], w[6],
                                             merged_kernel=c.merged_kernel, dtype=dtype))
        elif isinstance(c, svdlstm.ReducedLSTMCell):
            if c.merged_kernel:
                cells.append(oracle.ReducedCell(c.units, [w[0], w[1]], [w[2], w[3]], w[4], True, dtype=dtype))
            else:
                ww = [[w[4 * g], w[4 * g + 1]] for g in range(4)]
                uu = [[w[4 * g + 2], w[4 * g + 3]] for g in range(4)]
                cells.append(oracle.ReducedCell(c.units, ww, uu, w[16], False, dtype=dtype))
        else:
            cells.append(oracle.FullCell(c.units, w[0], w[1], w[2], dtype=dtype))
    dk, db = model.layers[-1].get_weights()
    return oracle.Model(cells, (dk, db), return_sequences=lstms[-1].return_sequences)


