"""CPU (-m "not gpu") checks of the C-ABI boundary: the library loads, exports every symbol that
include/svdlstm.h declares, validates arguments, and REFUSES to compute without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import svdlstm
from svdlstm import _cabi as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "svdlstm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svdlstm_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = svdlstm.lib()
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), "libsvdlstm.so does not export %s" % n
    assert sorted(C.EXPORTS) == names
    assert b"sm_100a" in lib.svdlstm_version()


def test_create_validates_arguments():
    lib = svdlstm.lib()
    h = ctypes.c_void_p()
    assert lib.svdlstm_create(ctypes.byref(h), 0, 16, C.int_array([15])) < 0
    assert b"n_layers" in lib.svdlstm_last_error()
    assert lib.svdlstm_create(ctypes.byref(h), 1, 0, C.int_array([15])) < 0
    assert lib.svdlstm_create(ctypes.byref(h), 2, 16, C.int_array([15, 0])) < 0
    assert lib.svdlstm_create(ctypes.byref(h), 3, 16, C.int_array([15, 15, 15])) == 0
    assert lib.svdlstm_count_weights(h) == 0
    # rank checks of the weight setters are host-side (pointers are only stored)
    fake = (ctypes.c_void_p * 7)(*[8] * 7)
    assert lib.svdlstm_set_singular_weights(h, 0, 1, fake, 17, 15) < 0      # k_w > min(D,4H)
    assert lib.svdlstm_set_singular_weights(h, 0, 0, fake, 16, 15) < 0      # split: k_w > min(D,H)
    assert lib.svdlstm_set_singular_weights(h, 0, 1, fake, 16, 15) == 0
    assert lib.svdlstm_count_weights(h) == 16 * (1 + 16 + 60) + 15 * (1 + 15 + 60) + 60
    assert lib.svdlstm_set_singular_weights(h, 5, 1, fake, 16, 15) < 0      # layer out of range
    assert lib.svdlstm_set_reduced_weights(h, 1, 1, (ctypes.c_void_p * 5)(*[8] * 5), C.int_array([61, 8])) < 0
    assert lib.svdlstm_set_reduced_weights(h, 1, 1, (ctypes.c_void_p * 5)(*[8] * 5), C.int_array([8, 8])) == 0
    assert lib.svdlstm_set_dense_top(h, 8, 8, 100) < 0
    lib.svdlstm_destroy(h)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = svdlstm.lib()
    h = ctypes.c_void_p()
    assert lib.svdlstm_create(ctypes.byref(h), 1, 4, C.int_array([3])) == 0
    fake = 8
    assert lib.svdlstm_set_full_weights(h, 0, fake, fake, fake) == 0
    rc = lib.svdlstm_forward(h, fake, 1, 1, fake, None, None, None, None, None, 0, 0, None)
    assert rc != 0 and b"no CUDA device" in lib.svdlstm_last_error()
    lib.svdlstm_destroy(h)
    # the Python surface refuses as well, before touching any data
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        svdlstm.full_model_from_weights([(np.zeros((4, 12), np.float32), np.zeros((3, 12), np.float32),
                                          np.zeros(12, np.float32))], (np.zeros((3, 1), np.float32), np.zeros(1, np.float32)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        svdlstm.HoyerRegularizer(0.01)(np.ones((1, 4), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        svdlstm.svd_batched(np.eye(3, dtype=np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lstm-acceleration-with-singular-value-decomposition_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "svdlstm_oracle" not in txt and "import oracle" not in txt and "svdlstm_torch_ref" not in txt, f


def test_new_entry_points_validate_on_the_host():
    """Round-2 surface (real-time stream, trainer, batched K2b, K5): argument / model checks are host-side and run without a
    GPU; anything that would compute refuses."""
    lib = svdlstm.lib()
    fake7 = (ctypes.c_void_p * 7)(*[8] * 7)
    # a model the wavefront engine cannot hold -> the stream says why
    big = ctypes.c_void_p()
    assert lib.svdlstm_create(ctypes.byref(big), 1, 16, C.int_array([64])) == 0
    assert lib.svdlstm_set_singular_weights(big, 0, 1, fake7, 16, 64) == 0
    st = ctypes.c_void_p()
    assert lib.svdlstm_stream_open(big, 0, ctypes.byref(st)) < 0 and b"wavefront" in lib.svdlstm_last_error()
    # weights never set
    h = ctypes.c_void_p()
    assert lib.svdlstm_create(ctypes.byref(h), 2, 16, C.int_array([15, 15])) == 0
    assert lib.svdlstm_stream_open(h, 0, ctypes.byref(st)) < 0 and b"never set" in lib.svdlstm_last_error()
    tr = ctypes.c_void_p()
    assert lib.svdlstm_trainer_create(h, None, ctypes.byref(tr)) < 0 and b"never set" in lib.svdlstm_last_error()
    # trainer: only 3-factor cells; flat layout = get_weights() order per layer, then the Dense top
    assert lib.svdlstm_set_singular_weights(h, 0, 1, fake7, 16, 15) == 0
    assert lib.svdlstm_set_full_weights(h, 1, 8, 8, 8) == 0
    assert lib.svdlstm_trainer_create(h, None, ctypes.byref(tr)) < 0 and b"3-factor" in lib.svdlstm_last_error()
    assert lib.svdlstm_set_singular_weights(h, 1, 0, fake7, 15, 15) == 0          # split: 4 per-gate blocks of rank 15
    assert lib.svdlstm_set_dense_top(h, 8, 8, 1) == 0
    assert lib.svdlstm_trainer_create(h, C.int_array([1, 0]), ctypes.byref(tr)) == 0
    offs = (ctypes.c_int64 * (7 * 2 + 3))()
    assert lib.svdlstm_trainer_layout(tr, offs) == 0
    sizes0 = [16, 15, 16 * 16, 16 * 60, 15 * 15, 15 * 60, 60]                      # merged layer: D=16, H=15
    sizes1 = [60, 60, 15 * 60, 15 * 60, 15 * 60, 15 * 60, 60]                       # split layer: sigma (1,4k), left (D,4k), right (k,4H)
    exp, o = [], 0
    for sz in sizes0 + sizes1:
        exp.append(o)
        o += sz
    exp += [o, o + 15, o + 16]
    assert list(offs) == exp and lib.svdlstm_trainer_num_params(tr) == o + 16
    if not torch.cuda.is_available():
        assert lib.svdlstm_stream_open(h, 0, ctypes.byref(st)) != 0 and b"no CUDA device" in lib.svdlstm_last_error()
    lib.svdlstm_trainer_destroy(tr)
    # batched K2b / K5 argument checks
    item = C.ReduceItem(8, 8, 8, 8, None, None, 4, 4, 3, 5, 4)                       # r > n
    assert lib.svdlstm_reduce_factors_batched(ctypes.byref(item), 1, None) < 0
    assert lib.svdlstm_reduce_factors_batched(None, 0, None) < 0
    assert lib.svdlstm_scaled_matmul(8, 2, None, 8, 4, None, 3, 5, 4, 8, 4, None) < 0 and b"bad shape" in lib.svdlstm_last_error()   # lda < k
    # forward with the upload inside: argument checks first, then "this model cannot" (-3) before anything touches a device
    fsi = lib.svdlstm_forward_streamed_input
    assert fsi(h, None, 8, 4, 8, 8, 4, 8, 16) < 0 and b"null" in lib.svdlstm_last_error()
    assert fsi(h, 8, 8, 4, 8, 8, -1, 8, 16) < 0 and b"n_slices" in lib.svdlstm_last_error()
    assert fsi(h, 8, 8, 4, 8, 8, 4, 8, 8) < 0 and b"stream of its own" in lib.svdlstm_last_error()
    wide = ctypes.c_void_p()                                                           # 512 full-rank units: rank 512 > 256
    assert lib.svdlstm_create(ctypes.byref(wide), 1, 16, C.int_array([512])) == 0
    assert lib.svdlstm_set_full_weights(wide, 0, 8, 8, 8) == 0
    assert fsi(wide, 8, 8, 256, 64, 8, 4, 8, 16) == -3 and b"ranks above 256" in lib.svdlstm_last_error()
    lib.svdlstm_destroy(wide)
    lib.svdlstm_destroy(h)
    lib.svdlstm_destroy(big)
