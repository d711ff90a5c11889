"""Host-side mirror of the voxel generator interface the reference uses.

Reference: VoxelGeneratorWrapper (pcdet/datasets/processor/data_processor.py:15-60) which wraps
spconv.utils.Point2VoxelCPU3d / VoxelGeneratorV2.  Same constructor keywords, same return triple
(voxels (M,T,C) f32, coordinates (M,3) i32 in z,y,x order, num_points (M,) i32); the work runs on
the GPU through comb_voxelize.  There is no CPU fallback.
"""
import numpy as np
import torch

from . import ops


class TVTensor:
    """Minimal stand-in for cumm.tensorview.Tensor: what data_processor.py:54-59 touches."""

    def __init__(self, array):
        self._a = array

    def numpy(self):
        return np.array(self._a, copy=True)

    def numpy_view(self):
        return self._a

    @property
    def shape(self):
        return self._a.shape


def from_numpy(arr):
    return TVTensor(np.ascontiguousarray(arr))


def _as_cuda_points(pc):
    if isinstance(pc, TVTensor):
        pc = pc.numpy_view()
    if isinstance(pc, np.ndarray):
        pc = torch.from_numpy(np.ascontiguousarray(pc, dtype=np.float32))
    if not isinstance(pc, torch.Tensor):
        raise TypeError("points must be a numpy array, a tensorview tensor or a torch tensor")
    dev = ops.host_op_device()      # worker-process policy: spawn -> lazy context on the rank's GPU, bad fork -> raises
    return pc.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()


class Point2VoxelGPU3d:
    """Drop-in for spconv.utils.Point2VoxelCPU3d (exported under that name by the spconv shim)."""

    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_points_per_voxel, max_num_voxels):
        self.vsize = [float(v) for v in vsize_xyz]
        self.coors_range = [float(v) for v in coors_range_xyz]
        self.num_point_features = int(num_point_features)
        self.max_points = int(max_num_points_per_voxel)
        self.max_voxels = int(max_num_voxels)
        rng = np.asarray(self.coors_range, dtype=np.float32)
        vs = np.asarray(self.vsize, dtype=np.float32)
        self.grid_size = np.round((rng[3:] - rng[:3]) / vs).astype(np.int64)

    def point_to_voxel_torch(self, pts):
        """pts (N,C) CUDA fp32 -> (voxels (M,T,C), coords (M,3) zyx, num (M,)) CUDA tensors."""
        n = int(pts.shape[0])
        r = ops.voxelize(pts, [0, n], self.vsize, self.coors_range, self.max_points, self.max_voxels)
        m = int(r["counts"][1].item())
        return r["voxels"][:m], r["coords"][:m, 1:].contiguous(), r["num_points"][:m]

    def point_to_voxel(self, pc, clear_voxels=True):
        pts = _as_cuda_points(pc)
        assert pts.shape[1] == self.num_point_features, "num_point_features mismatch"
        with torch.cuda.device(pts.device):
            v, c, n = self.point_to_voxel_torch(pts)
            return TVTensor(v.cpu().numpy()), TVTensor(c.cpu().numpy()), TVTensor(n.cpu().numpy())


class VoxelGeneratorWrapper:
    """Same interface as data_processor.py:15-60."""

    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_points_per_voxel, max_num_voxels):
        self.spconv_ver = 2
        self._voxel_generator = Point2VoxelGPU3d(
            vsize_xyz=vsize_xyz, coors_range_xyz=coors_range_xyz, num_point_features=num_point_features,
            max_num_points_per_voxel=max_num_points_per_voxel, max_num_voxels=max_num_voxels)

    def generate(self, points):
        tv_voxels, tv_coordinates, tv_num_points = self._voxel_generator.point_to_voxel(from_numpy(points))
        return tv_voxels.numpy(), tv_coordinates.numpy(), tv_num_points.numpy()
