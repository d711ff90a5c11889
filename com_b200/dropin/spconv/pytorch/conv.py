from com_b200.sparse import SparseConv3d, SparseConvolution, SparseInverseConv3d, SubMConv3d  # noqa: F401
