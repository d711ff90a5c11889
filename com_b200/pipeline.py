"""Frame pipeline of the hot path — the public call a user makes for inference:

    points (per frame, host or device)  ->  voxelize + MeanVFE  ->  VoxelResBackBone8x  ->  HeightCompression

Reference call chain it replaces (one process per GPU, frames are independent units):
  DataProcessor.transform_points_to_voxels   pcdet/datasets/processor/data_processor.py:125-153
  collate_batch / load_data_to_gpu           pcdet/datasets/dataset.py:252-259, pcdet/models/__init__.py:23-37
  MeanVFE.forward                            pcdet/models/backbones_3d/vfe/mean_vfe.py:14-31
  VoxelResBackBone8x.forward                 pcdet/models/backbones_3d/spconv_backbone.py:241-293
  HeightCompression.forward                  pcdet/models/backbones_2d/map_to_bev/height_compression.py:10-26

Everything between the H2D copy of the points and the dense BEV tensor runs on the device without a
host synchronisation except one read of the per-level row counts (models.VoxelResBackBone8x.forward_fused).
"""
import numpy as np
import torch

from . import models, ops, synth


class FramePipeline:
    def __init__(self, input_channels=5, point_cloud_range=None, voxel_size=None, max_points_per_voxel=5,
                 max_voxels=150000, device="cuda", seed=0):
        self.range = list(point_cloud_range or synth.POINT_CLOUD_RANGE)
        self.vsize = list(voxel_size or synth.VOXEL_SIZE)
        self.T, self.max_voxels, self.C = int(max_points_per_voxel), int(max_voxels), int(input_channels)
        rng, vs = np.asarray(self.range, np.float32), np.asarray(self.vsize, np.float32)
        self.grid_size = np.round((rng[3:] - rng[:3]) / vs).astype(np.int64).tolist()      # x, y, z
        torch.manual_seed(seed)
        self.vfe = models.MeanVFE(None, input_channels)
        self.backbone = models.VoxelResBackBone8x(None, input_channels, self.grid_size).to(device).eval()
        self.to_bev = models.HeightCompression(None)
        self.device = torch.device(device)

    @torch.no_grad()
    def forward_device(self, points, frame_offsets):
        """points (N_total, C) fp32 CUDA, frames concatenated; frame_offsets python ints.
        -> batch_dict with voxel_coords, encoded_spconv_tensor, multi_scale_3d_features, spatial_features.
        Everything is enqueued without waiting for the device (row counts stay on the device, tensors are
        capacity-sized); ONE host read of all counts at the end sizes the returned views."""
        batch = len(frame_offsets) - 1
        r = ops.voxelize(points, frame_offsets, self.vsize, self.range, self.T, self.max_voxels, want_voxels=False,
                         mean_dtype=torch.bfloat16, mean_ld=16)
        n_dev = r["counts"][batch:batch + 1]
        worst = False
        while True:
            levels, counts, caps = self.backbone.fused_async(r["mean"], r["coords"], batch, n_dev=n_dev, worst=worst)
            x, c, shape = levels[-1]
            dense = ops.dense(x, c, batch, shape, n_dev=counts[4:5])
            host = torch.cat([r["counts"], counts]).tolist()          # the only host synchronisation
            m, cnt = host[batch], host[batch + 1:]
            outs = self.backbone.fused_finish(levels, cnt, caps, int(r["coords"].shape[0]), batch)
            if outs is not None or worst:
                break
            worst = True                                              # a learned capacity overflowed: redo
        x1, x2, x3, x4, out = outs
        n, ch, d, h, w = dense.shape
        return {"batch_size": batch, "voxel_features": r["mean"][:m], "voxel_coords": r["coords"][:m],
                "voxel_num_points": r["num_points"][:m], "voxel_counts": r["counts"],
                "encoded_spconv_tensor": out, "encoded_spconv_tensor_stride": 8,
                "multi_scale_3d_features": {"x_conv1": x1, "x_conv2": x2, "x_conv3": x3, "x_conv4": x4},
                "multi_scale_3d_strides": {"x_conv1": 1, "x_conv2": 2, "x_conv3": 4, "x_conv4": 8},
                "spatial_features": dense.view(n, ch * d, h, w), "spatial_features_stride": 8}

    @torch.no_grad()
    def forward_host(self, frames, pinned=None):
        """frames: list of (N_b, C) float32 numpy arrays (or one pinned CPU tensor + offsets in `pinned`).
        Copies host -> device on the current stream, then runs forward_device."""
        if pinned is not None:
            host, offs = pinned
        else:
            offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(int).tolist()
            host = torch.from_numpy(np.concatenate(frames, axis=0))
        dev = host.to(self.device, non_blocking=True)
        return self.forward_device(dev, offs)
