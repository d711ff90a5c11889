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
import os

import numpy as np
import torch

from . import _lib, models, ops, synth


class FramePipeline:
    def __init__(self, input_channels=5, point_cloud_range=None, voxel_size=None, max_points_per_voxel=5,
                 max_voxels=150000, device="cuda", seed=0, use_graph=False):
        self._init = dict(input_channels=input_channels, point_cloud_range=point_cloud_range, voxel_size=voxel_size,
                          max_points_per_voxel=max_points_per_voxel, max_voxels=max_voxels, device=device, seed=seed,
                          use_graph=use_graph)
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
        self.use_graph = bool(use_graph)     # replay the step as one CUDA graph (see _forward_graph)
        self._graph = None
        self._zero_stream = None             # side stream of the early BEV zero-fill (see _enqueue)
        self.graph_launches = 0              # kernels launched through graph replays (bench.py's gpu_launches)

    def clone_lane(self):
        """A second pipeline over the SAME modules (weights, folded plan) with its own graph and static buffers."""
        other = FramePipeline(**self._init)
        other.vfe, other.backbone, other.to_bev = self.vfe, self.backbone, self.to_bev
        return other

    # ------------------------------------------------------------------ enqueue (no host synchronisation)
    def _enqueue(self, points, offsets, batch, worst=False):
        """Enqueue voxelize -> backbone -> dense on the current stream.  `offsets`: python ints or a device int32
        tensor (batch+1).  Returns capacity-sized tensors and ONE device tensor holding every row count."""
        # the zero-fill of the dense BEV tensor (36 MB per frame) depends on nothing: start it on a side stream now and
        # only scatter the ~6k active cells per frame at the end (comb_dense_scatter)
        main = torch.cuda.current_stream(self.device)
        oshape = self.backbone.out_spatial_shape()
        dense = torch.empty((batch, self.backbone.num_point_features, *oshape), dtype=torch.float32, device=self.device)
        overlap = os.environ.get("COMB_OVERLAP", "1") != "0"
        if os.environ.get("COMB_DENSE_SCATTER", "1") == "0":      # A/B switch: one kernel writes the whole tensor
            r = ops.voxelize(points, offsets, self.vsize, self.range, self.T, self.max_voxels, want_voxels=False,
                             mean_dtype=torch.bfloat16, mean_ld=16)
            n_dev = r["counts"][batch:batch + 1]
            levels, counts, caps = self.backbone.fused_async(r["mean"], r["coords"], batch, n_dev=n_dev, worst=worst)
            x, c, shape = levels[-1]
            dense = ops.dense(x, c, batch, shape, n_dev=counts[4:5])
            return dict(r=r, levels=levels, caps=caps, dense=dense, all_counts=torch.cat([r["counts"], counts]))
        if overlap:
            if self._zero_stream is None:
                self._zero_stream = torch.cuda.Stream(self.device)
            fork, joined = torch.cuda.Event(), torch.cuda.Event()
            fork.record(main)
            self._zero_stream.wait_event(fork)
            with torch.cuda.stream(self._zero_stream):
                dense.zero_()
                joined.record(self._zero_stream)
        else:
            dense.zero_()
        r = ops.voxelize(points, offsets, self.vsize, self.range, self.T, self.max_voxels, want_voxels=False,
                         mean_dtype=torch.bfloat16, mean_ld=16)
        n_dev = r["counts"][batch:batch + 1]
        levels, counts, caps = self.backbone.fused_async(r["mean"], r["coords"], batch, n_dev=n_dev, worst=worst)
        x, c, shape = levels[-1]
        assert list(shape) == list(oshape)
        if overlap:
            main.wait_event(joined)
        ops.dense_scatter(x, c, batch, shape, dense, n_dev=counts[4:5])
        return dict(r=r, levels=levels, caps=caps, dense=dense, all_counts=torch.cat([r["counts"], counts]))

    def _finish(self, q, batch):
        """One host read of all row counts, then exact-size views (None when a learned capacity overflowed)."""
        host = q["all_counts"].tolist()                              # the only host synchronisation
        m, cnt = host[batch], host[batch + 1:]
        r = q["r"]
        outs = self.backbone.fused_finish(q["levels"], cnt, q["caps"], int(r["coords"].shape[0]), batch)
        if outs is None:
            return None
        x1, x2, x3, x4, out = outs
        n, ch, d, h, w = q["dense"].shape
        return {"batch_size": batch, "voxel_features": r["mean"][:m], "voxel_coords": r["coords"][:m],
                "voxel_num_points": r["num_points"][:m], "voxel_counts": r["counts"],
                "encoded_spconv_tensor": out, "encoded_spconv_tensor_stride": 8,
                "multi_scale_3d_features": {"x_conv1": x1, "x_conv2": x2, "x_conv3": x3, "x_conv4": x4},
                "multi_scale_3d_strides": {"x_conv1": 1, "x_conv2": 2, "x_conv3": 4, "x_conv4": 8},
                "spatial_features": q["dense"].view(n, ch * d, h, w), "spatial_features_stride": 8}

    @torch.no_grad()
    def forward_device(self, points, frame_offsets):
        """points (N_total, C) fp32 CUDA, frames concatenated; frame_offsets python ints.
        -> batch_dict with voxel_coords, encoded_spconv_tensor, multi_scale_3d_features, spatial_features.
        Everything is enqueued without waiting for the device (row counts stay on the device, tensors are
        capacity-sized); ONE host read of all counts at the end sizes the returned views."""
        batch = len(frame_offsets) - 1
        if self.use_graph:
            return self._forward_graph(points, frame_offsets, batch)
        out = self._finish(self._enqueue(points, frame_offsets, batch), batch)
        if out is None:                                              # a learned capacity overflowed: redo
            out = self._finish(self._enqueue(points, frame_offsets, batch, worst=True), batch)
        return out

    # ------------------------------------------------------------------ CUDA-graph replay
    def _forward_graph(self, points, frame_offsets, batch):
        """The whole step as ONE cudaGraphLaunch: the kernels read frame offsets and row counts from device
        memory, so the captured launch sequence is valid for any frames that fit the captured capacities
        (point capacity, voxel capacity, learned level capacities).  Host work per step: two small copies into
        the static input buffers, the launch, one read of the counts."""
        n = int(points.shape[0])
        g = self._graph
        if g is None or g["batch"] != batch or n > g["n_cap"]:
            g = self._capture(points.to(self.device, non_blocking=True), frame_offsets, batch)
        g["points"][:n].copy_(points, non_blocking=True)             # D2D, or H2D straight from pinned memory
        self._stage_offsets(g, frame_offsets)
        g["graph"].replay()
        self.graph_launches += g["launches"]
        out = self._finish(g["q"], batch)
        if out is None:                                              # capacity overflow: eager redo + recapture later
            self._graph = None
            out = self._finish(self._enqueue(points, frame_offsets, batch, worst=True), batch)
        return out

    @torch.no_grad()
    def enqueue_device(self, points, frame_offsets):
        """Asynchronous form of forward_device in CUDA-graph mode: copy the (device or pinned host) points into the
        graph's input buffer and replay, WITHOUT reading the row counts back — returns a handle for finish().  Steps
        enqueued back to back run on the device with no host round trip in between."""
        if not self.use_graph:
            raise RuntimeError("enqueue_device needs FramePipeline(use_graph=True)")
        batch, n = len(frame_offsets) - 1, int(points.shape[0])
        g = self._graph
        if g is None or g["batch"] != batch or n > g["n_cap"]:
            g = self._capture(points.to(self.device, non_blocking=True), frame_offsets, batch)
        g["points"][:n].copy_(points, non_blocking=True)
        self._stage_offsets(g, frame_offsets)
        g["graph"].replay()
        self.graph_launches += g["launches"]
        return (g, batch)

    @staticmethod
    def _stage_offsets(g, frame_offsets):
        """Frame offsets -> the graph's device buffer through a small RING of pinned staging buffers: a slot is only
        rewritten after the asynchronous H2D copy that last read it has completed (its event), so steps enqueued back
        to back without finish() never see each other's offsets."""
        ring = g["offs_ring"]
        k = g["offs_k"] % len(ring)
        g["offs_k"] += 1
        host, ev = ring[k]
        ev.synchronize()
        host[:] = torch.tensor(frame_offsets, dtype=torch.int32)
        g["offs"].copy_(host, non_blocking=True)
        ev.record(torch.cuda.current_stream(g["offs"].device))

    def finish(self, handle):
        """Read the row counts of the LAST replay of the handle's graph and build the batch_dict (None when a learned
        capacity overflowed: call forward_device for that batch)."""
        g, batch = handle
        return self._finish(g["q"], batch)

    def _capture(self, points, frame_offsets, batch):
        n = int(points.shape[0])
        n_cap = max((int(n * 1.25) + 65535) // 65536 * 65536, 65536)
        # eager warm-up: builds the weight plan, sets kernel attributes, learns the level capacities
        for _ in range(2):
            if self._finish(self._enqueue(points, frame_offsets, batch), batch) is None:
                self._finish(self._enqueue(points, frame_offsets, batch, worst=True), batch)
        g = {"batch": batch, "n_cap": n_cap,
             "points": torch.zeros((n_cap, int(points.shape[1])), dtype=torch.float32, device=points.device),
             "offs": torch.zeros((batch + 1,), dtype=torch.int32, device=points.device),
             "offs_ring": [(torch.zeros((batch + 1,), dtype=torch.int32).pin_memory(), torch.cuda.Event())
                           for _ in range(4)],
             "offs_k": 0}
        g["points"][:n].copy_(points)
        g["offs"].copy_(torch.tensor(frame_offsets, dtype=torch.int32))
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        lib = _lib.load()
        n0 = lib.comb_launch_count()
        with torch.cuda.graph(graph):
            g["q"] = self._enqueue(g["points"], g["offs"], batch)
        g["launches"] = int(lib.comb_launch_count() - n0)    # kernels of this library inside one replay
        g["graph"] = graph
        self._graph = g
        return g

    @torch.no_grad()
    def forward_host(self, frames, pinned=None):
        """frames: list of (N_b, C) float32 numpy arrays (or one pinned CPU tensor + offsets in `pinned`).
        Copies host -> device on the current stream, then runs forward_device."""
        if pinned is not None:
            host, offs = pinned
        else:
            offs = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(int).tolist()
            host = torch.from_numpy(np.concatenate(frames, axis=0))
        if self.use_graph:
            return self._forward_graph(host, offs, len(offs) - 1)    # H2D goes straight into the graph's input buffer
        dev = host.to(self.device, non_blocking=True)
        return self.forward_device(dev, offs)


class FrameStream:
    """Streaming front end of a FramePipeline in CUDA-graph mode — the serving loop:

        ticket = stream.submit(pinned_points, frame_offsets)      # returns at once
        ...                                                       # submit the next batch before collecting
        res = stream.result(ticket)                               # pinned host views of the encoded tensor

    Two LANES, each with its own captured step graph, static buffers and launch stream; batches alternate between
    them.  Batches are independent, so the head of batch k+1 (H2D straight into its lane's graph input on the copy-in
    stream, voxelizer, level-1 index) runs while the convolutions of batch k are still in flight, and the D2H of a
    lane's result (copy-out stream, straight from the graph's output buffers) overlaps the other lane's kernels.
    The streams are ordered with events only; the host blocks in result() alone.  (r1 measurement, device-resident:
    1.147 ms/step with one lane, 1.082 with two.)  Replaces the synchronous load_data_to_gpu -> model -> .cpu() loop of
    pcdet/models/__init__.py:23-37 and tools/eval_utils/eval_utils.py:58-71 for this path.
    Result views stay valid until the second submit() after their own."""

    LANES = 2

    def __init__(self, pipe, host_points, frame_offsets):
        if not pipe.use_graph:
            raise RuntimeError("FrameStream needs a FramePipeline(use_graph=True)")
        self.pipe, dev = pipe, pipe.device
        self.batch = len(frame_offsets) - 1
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        # the payload copy of step k is enqueued by the host AFTER step k+1 has been submitted: on s_out it would queue
        # behind the wait for step k+1's replay (measured: e2e 3520 -> 2560 frames/s), so it gets a stream of its own
        self.s_pay = torch.cuda.Stream(dev)
        ev = lambda: torch.cuda.Event(enable_timing=True)
        self.lanes = []
        for i in range(self.LANES):
            p = pipe if i == 0 else pipe.clone_lane()
            launch = torch.cuda.Stream(dev)
            with torch.cuda.stream(launch):
                g = p._graph
                if g is None or g["batch"] != self.batch or int(host_points.shape[0]) > g["n_cap"]:
                    g = p._capture(host_points.to(dev), frame_offsets, self.batch)
            q = g["q"]
            x, c, _ = q["levels"][-1]
            self.lanes.append(dict(
                pipe=p, g=g, launch=launch,
                offs_host=torch.zeros((self.batch + 1,), dtype=torch.int32).pin_memory(),
                h_feat=torch.empty(tuple(x.shape), dtype=x.dtype).pin_memory(),
                h_idx=torch.empty(tuple(c.shape), dtype=c.dtype).pin_memory(),
                h_cnt=torch.empty(tuple(q["all_counts"].shape), dtype=q["all_counts"].dtype).pin_memory(),
                ev_in=ev(), ev_out=ev(), ev_cnt=ev(), ev_done=ev(), src=None, pending=False))
        torch.cuda.synchronize(dev)
        g0 = self.lanes[0]["g"]
        x, c, _ = g0["q"]["levels"][-1]
        cnt = g0["q"]["all_counts"]
        self.k = 0
        self.h2d_bytes = 0
        self.d2h_bytes = int(x.numel() * x.element_size() + c.numel() * c.element_size() + cnt.numel() * cnt.element_size())

    @torch.no_grad()
    def submit(self, host_points, frame_offsets):
        lane = self.lanes[self.k % self.LANES]
        g, launch = lane["g"], lane["launch"]
        n = int(host_points.shape[0])
        if len(frame_offsets) - 1 != self.batch or n > g["n_cap"]:
            raise RuntimeError("FrameStream: batch shape differs from the captured one (build a new stream)")
        if lane.get("pending"):                           # submit twice on a lane without result(): fetch it now
            lane["stash"] = self._collect(lane)
        lane["ev_in"].synchronize()                       # the previous H2D out of offs_host has long finished
        lane["offs_host"][:] = torch.tensor(frame_offsets, dtype=torch.int32)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(lane["ev_out"])          # the lane's previous replay has consumed its input buffers
            g["points"][:n].copy_(host_points, non_blocking=True)
            g["offs"].copy_(lane["offs_host"], non_blocking=True)
            lane["ev_in"].record(self.s_in)
        with torch.cuda.stream(launch):
            launch.wait_event(lane["ev_in"])
            launch.wait_event(lane["ev_done"])            # the lane's previous result has left its output buffers
            g["graph"].replay()
            lane["ev_out"].record(launch)
        lane["pipe"].graph_launches += g["launches"]
        q = g["q"]
        with torch.cuda.stream(self.s_out):
            # first the row counts (a few bytes); the payload follows in _collect() with its EXACT size once the host
            # knows the count (the capacity-sized copy moved 10.1 MB for 6.5 MB of live rows per step)
            self.s_out.wait_event(lane["ev_out"])
            lane["h_cnt"].copy_(q["all_counts"], non_blocking=True)
            lane["ev_cnt"].record(self.s_out)
        lane["src"] = (host_points, list(frame_offsets))
        lane["pending"] = True
        self.h2d_bytes = n * int(host_points.shape[1]) * 4 + (self.batch + 1) * 4
        self.k += 1
        return self.k - 1

    def _collect(self, lane):
        """Second half of a step's device->host traffic: wait for the row counts, then copy exactly the live rows of the
        encoded tensor.  Returns (counts, overflowed)."""
        batch = self.batch
        lane["ev_cnt"].synchronize()
        cnt = lane["h_cnt"].tolist()
        lv = cnt[batch + 1:]
        q = lane["g"]["q"]
        caps = q["caps"]
        n1 = int(q["r"]["coords"].shape[0])
        hard = self.pipe.backbone._caps(n1, batch, worst=True)
        over = any(c >= caps[li] and caps[li] < hard[li] for c, li in zip(lv[1:], (2, 3, 4, 5)))
        n = 0 if over else lv[4]
        x, c, _ = q["levels"][-1]
        with torch.cuda.stream(self.s_pay):
            self.s_pay.wait_event(lane["ev_cnt"])
            if n > 0:
                lane["h_feat"][:n].copy_(x[:n], non_blocking=True)
                lane["h_idx"][:n].copy_(c[:n], non_blocking=True)
            lane["ev_done"].record(self.s_pay)
        lane["pending"] = False
        self.d2h_bytes = int(n * (x.shape[1] * x.element_size() + c.shape[1] * c.element_size())
                             + lane["h_cnt"].numel() * lane["h_cnt"].element_size())
        return cnt, over

    @property
    def graph_launches(self):
        return sum(l["pipe"].graph_launches for l in self.lanes)

    def done_event(self, ticket):
        return self.lanes[ticket % self.LANES]["ev_done"]

    @torch.no_grad()
    def result(self, ticket):
        lane, batch = self.lanes[ticket % self.LANES], self.batch
        cnt, over = self._collect(lane) if lane.get("pending") else lane.pop("stash")
        lane["ev_done"].synchronize()
        lv = cnt[batch + 1:]
        if over:
            # a learned level capacity overflowed: redo this batch synchronously with worst-case capacities
            # (on the lane's own launch stream, after everything in flight there); forward_host recaptures the lane's
            # graph with the capacities learned from this batch, and the lane adopts the new graph and output buffers
            host, offs = lane["src"]
            with torch.cuda.stream(lane["launch"]):
                lane["pipe"]._graph = None
                bd = lane["pipe"].forward_host(None, pinned=(host, offs))
                enc = bd["encoded_spconv_tensor"]
                res = {"features": enc.features.cpu(), "indices": enc.indices.cpu(),
                       "voxel_counts": bd["voxel_counts"].cpu(), "rows": int(enc.features.shape[0])}
                g = lane["pipe"]._graph
                if g is not None:
                    lane["g"] = g
                    x, c, _ = g["q"]["levels"][-1]
                    if tuple(x.shape) != tuple(lane["h_feat"].shape):
                        lane["h_feat"] = torch.empty(tuple(x.shape), dtype=x.dtype).pin_memory()
                        lane["h_idx"] = torch.empty(tuple(c.shape), dtype=c.dtype).pin_memory()
                lane["launch"].synchronize()
            return res
        n = lv[4]
        return {"features": lane["h_feat"][:n], "indices": lane["h_idx"][:n],
                "voxel_counts": torch.tensor(cnt[: batch + 1], dtype=torch.int32), "rows": n}
