from com_b200.voxel import TVTensor as Tensor, from_numpy  # noqa: F401
