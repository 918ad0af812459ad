from ._irreps import Irrep, Irreps  # noqa: F401
from ._sh import SphericalHarmonics  # noqa: F401
from ._tp import (  # noqa: F401
    ElementwiseTensorProduct,
    FullyConnectedTensorProduct,
    Instruction,
    Linear,
    TensorProduct,
)
from ._wigner import wigner_3j  # noqa: F401
