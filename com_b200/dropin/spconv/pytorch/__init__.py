from com_b200.sparse import (SparseConv3d, SparseConvTensor, SparseConvolution, SparseInverseConv3d,  # noqa: F401
                             SparseModule, SparseSequential, SubMConv3d, ToDense)
from . import conv  # noqa: F401
