"""svdlstm -- B200-native (sm_100a) implementation of the SVD-factored LSTM hot path of
dncoble/LSTM-acceleration-with-singular-value-decomposition.

Drop-in for the reference's layer / builder API (code/svd_classes_v3.py) and driver metrics
(code/svd_acceleration_v3.py), Python host -> C-ABI (include/svdlstm.h) -> hand-written CUDA.
There is no CPU fallback: the numpy oracle under oracle/ is test infrastructure only.
"""
from ._cabi import (ENGINE_AUTO, ENGINE_FP32, ENGINE_GENERAL, ENGINE_TC, ENGINE_WAVEFRONT, EXPORTS, LIB_PATH, TC_MIN_BATCH,
                    TC_MIN_UNITS, lib, pinned_empty, require_cuda)
from . import _cabi
from .layers import (Dense, Handle, HoyerRegularizer, InputLayer, LSTM, LSTMCell, OrthogonalRegularizer,
                     PrunableTimeDistributed, ReducedLSTMCell, SingularLSTM, SingularLSTMCell, TimeDistributed,
                     Variable, evaluate_penalties, get_default_engine, set_default_engine)
from .models import (RealtimeStream, Sequential, full_model_from_weights, make_LSTM_reduced_model, make_LSTM_singular_model,
                     make_split_LSTM_singular_model, reduce_factors, reduce_factors_batched, svd_batched, truncate_singular_model)
from .metrics import (count_weights, full_weight_count, reduced_merged_weight_count, reduced_split_weight_count,
                      reference_rmse, rmse, signaltonoise, sweep_sse, weight_reduction_percent)
from .rank_reduce import (LSTM_wrapper, get_model_singular_values, reduce_matrix_rank, reduce_two_step,
                          set_model_matrix_rank, sorted_sigma_indices)
from .data import StandardScaler, preprocess, split_train_random
from .training import History, Trainer
from .sweep import build_rank_models, rank_sweep, shard_bounds
from .weights_io import (load_model_weights_csv, load_model_weights_json, load_model_weights_npz, load_model_weights_zip,
                         save_model_weights_csv, save_model_weights_json, synthetic_layers)

__version__ = "0.1.0"


def launches() -> int:
    """Kernels launched through the C-ABI since import (bench.py's gpu_launches)."""
    return _cabi.launch_counter
