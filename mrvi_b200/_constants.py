"""Registry keys of the local-sample-distance path.

Mirrors ``src/mrvi/_constants.py:4-8`` of the reference (the one key the path uses)
plus the scvi-tools registry names the reference reads in ``_model.py:377-384``.
"""
from typing import NamedTuple


class _MRVI_REGISTRY_KEYS_NT(NamedTuple):
    SAMPLE_KEY: str = "sample"


MRVI_REGISTRY_KEYS = _MRVI_REGISTRY_KEYS_NT()

# scvi-tools REGISTRY_KEYS used by the path (reference `_model.py:377-378`, `_module.py:304`).
X_KEY = "X"
INDICES_KEY = "ind_x"

# Hard-coded in the reference: ResnetBlock.n_hidden default (`_components.py:31`).
RESNET_HIDDEN = 128
# flax.linen.LayerNorm default epsilon (third-party; see oracle header).
LN_EPS = 1e-6
