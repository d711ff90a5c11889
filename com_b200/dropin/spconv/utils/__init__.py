from com_b200.voxel import Point2VoxelGPU3d as Point2VoxelCPU3d  # noqa: F401  (runs on the GPU)
from com_b200.voxel import Point2VoxelGPU3d as Point2VoxelGPU3d  # noqa: F401
