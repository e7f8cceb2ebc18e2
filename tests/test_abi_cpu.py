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
                assert "svdlstm_oracle" not in txt and "import oracle" not in txt, f
