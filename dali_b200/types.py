"""nvidia.dali.types subset (dali/python/nvidia/dali/types.py): enums with the reference's numeric values
(include/dali/core/dali_data_type.h:44-57, common.h:144-162)."""
import enum


class DALIDataType(enum.IntEnum):
    NO_TYPE = -1
    UINT8 = 0
    UINT16 = 1
    UINT32 = 2
    UINT64 = 3
    INT8 = 4
    INT16 = 5
    INT32 = 6
    INT64 = 7
    FLOAT16 = 8
    FLOAT = 9
    FLOAT64 = 10
    BOOL = 11


class DALIInterpType(enum.IntEnum):
    INTERP_NN = 0
    INTERP_LINEAR = 1
    INTERP_CUBIC = 2
    INTERP_LANCZOS3 = 3
    INTERP_TRIANGULAR = 4
    INTERP_GAUSSIAN = 5


class DALIImageType(enum.IntEnum):
    RGB = 0
    BGR = 1
    GRAY = 2
    YCbCr = 3
    ANY_DATA = 4


NO_TYPE, UINT8, UINT16, UINT32, UINT64, INT8, INT16, INT32, INT64, FLOAT16, FLOAT, FLOAT64, BOOL = [DALIDataType(v) for v in range(-1, 12)]
INTERP_NN, INTERP_LINEAR, INTERP_CUBIC, INTERP_LANCZOS3, INTERP_TRIANGULAR, INTERP_GAUSSIAN = [DALIInterpType(v) for v in range(6)]
RGB, BGR, GRAY, YCbCr, ANY_DATA = [DALIImageType(v) for v in range(5)]

_NP = {UINT8: "uint8", UINT16: "uint16", UINT32: "uint32", UINT64: "uint64", INT8: "int8", INT16: "int16", INT32: "int32",
       INT64: "int64", FLOAT16: "float16", FLOAT: "float32", FLOAT64: "float64", BOOL: "bool"}


def to_numpy_type(t):
    import numpy as np
    return np.dtype(_NP[DALIDataType(int(t))])


def from_numpy_type(dt):
    import numpy as np
    dt = np.dtype(dt)
    for k, v in _NP.items():
        if np.dtype(v) == dt:
            return k
    raise TypeError(f"Unsupported numpy type {dt}")


class SampleInfo:
    """types.py:697-711 in the reference"""

    def __init__(self, idx_in_epoch, idx_in_batch, iteration, epoch_idx=0):
        self.idx_in_epoch, self.idx_in_batch, self.iteration, self.epoch_idx = idx_in_epoch, idx_in_batch, iteration, epoch_idx


class BatchInfo:
    def __init__(self, iteration, epoch_idx=0):
        self.iteration, self.epoch_idx = iteration, epoch_idx
