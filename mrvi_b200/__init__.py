"""mrvi_b200 -- B200-native counterfactual local-sample-distance path of MrVI.

Drop-in for ``mrvi.MrVI.get_local_sample_distances`` (reference ``src/mrvi/_model.py:627-714``):
host side in Python mirroring the reference API, device side hand-written sm_100a CUDA behind a
C-ABI (``include/mrvi_b200.h``).  No CPU fallback.
"""
from ._constants import MRVI_REGISTRY_KEYS
from ._data import SimpleAnnData, synthetic_counts, synthetic_iid
from ._model import MrVI
from ._params import (MrviDims, expected_param_shapes, flatten_params, infer_dims, init_params,
                      load_params, save_params, validate_params)
from ._types import MrVIReduction
from ._xr import DataArray, Dataset

__version__ = "0.1.0"

__all__ = [
    "MrVI", "MrVIReduction", "MrviDims", "MRVI_REGISTRY_KEYS", "SimpleAnnData", "synthetic_iid",
    "synthetic_counts", "flatten_params", "infer_dims", "init_params", "validate_params",
    "expected_param_shapes", "save_params", "load_params", "DataArray", "Dataset",
]
