"""Path E host side: PySyft-shaped SPDZ / fixed-precision API over the primia_b200 C ABI."""
from .ops import *  # noqa: F401,F403
from .spdz import (  # noqa: F401
    EmptyCryptoPrimitiveStoreError,
    Party,
    PrimitiveStorage,
    spdz_compute,
    spdz_mask,
    spdz_mul,
)
from .tensors import AdditiveSharingTensor, FixedPrecisionTensor  # noqa: F401
from . import functional  # noqa: F401
from . import fss  # noqa: F401
from .resnet import EncryptedResNet18  # noqa: F401
