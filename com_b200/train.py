"""Fused TRAIN-mode step of the VoxelResBackBone8x topology (forward + backward), the training form of rows a6-a9.

Reference contract: `VoxelResBackBone8x.forward` under `train()` (pcdet/models/backbones_3d/spconv_backbone.py:241-293)
with `SparseSequential(conv, BatchNorm1d(eps=1e-3, momentum=0.01), ReLU)` (:21-25) and `SparseBasicBlock.forward`
(:50-66), differentiated by autograd.  What the module path does with ~450 Python-driven launches per step (hash
rulebooks with a host read per strided conv, weight repacking in every backward, BatchNorm / ReLU / add in eager
torch) runs here as one planned chain, exactly like the fused eval path:

  * rows of every level in KEY order, bitmap-rank rulebooks, device-side row counts (no host read until the end);
  * per layer   raw = conv(x) (+bias)                       tcgen05 gather-GEMM, fp32 out
                act = relu(bn_train(raw) (+ identity))      comb_bn_train_fwd (batch statistics, running stats updated)
  * backward    draw, g = bn_train_bwd(dact, act, raw)      comb_bn_train_bwd (g = gradient of the identity branch)
                dW = wgrad(x, draw)                         tcgen05 wgrad, key-ordered rows
                dx = dgrad(draw) (+ g of the block)         the same gather-GEMM over the transposed rulebook with W^T;
                                                            for SubM the transposed rulebook is the rulebook itself with
                                                            the kernel offsets mirrored, so only W is re-indexed
  * W and W^T are packed once per weight version (i.e. once per optimizer step).

Activations and activation gradients are stored in bf16 (north_star: bf16 inputs, fp32 accumulate, 2e-2); parameter
gradients, statistics and the returned features are fp32.
"""
import os

import torch

from . import ops
from .sparse import SparseConvTensor


class _Layer:
    __slots__ = ("name", "conv", "bn", "level", "down", "block_end", "block_start", "cin", "cin_p", "cout", "K")

    def __init__(self, name, conv, bn, level, down=False, block_start=False, block_end=False):
        self.name, self.conv, self.bn, self.level, self.down = name, conv, bn, level, down
        self.block_start, self.block_end = block_start, block_end
        self.cin, self.cout = conv.in_channels, conv.out_channels
        self.cin_p = ops.pad16(self.cin)
        self.K = conv.kernel_size[0] * conv.kernel_size[1] * conv.kernel_size[2]


class FusedTrainer:
    """Planner / executor attached to a backbone module (the mirror class or the reference's own class)."""

    def __init__(self, module):
        self.m = module
        self.layers = self._plan(module)
        self._packed = {}           # layer index -> (weight version, fwd image, dgrad image)
        # COMB_TRAIN_GRAPH=0 keeps eager launches (one Python call per kernel); default: two CUDA graphs per step
        self.use_graph = os.environ.get("COMB_TRAIN_GRAPH", "1") != "0"
        self._graph = None
        self._capturing = False
        self.relu = True            # tests switch the ReLUs off to get a smooth loss (gradient parity without sign flips)
        self.last = None            # state of the last forward (levels for multi_scale_3d_features, counts)
        self.last_flat = None       # flat buffer holding every parameter gradient of the last graphed backward

    @staticmethod
    def _plan(m):
        L = [_Layer("conv_input", m.conv_input[0], m.conv_input[1], 1)]
        for li, stage in enumerate((m.conv1, m.conv2, m.conv3, m.conv4), start=1):
            mods = list(stage._modules.values())
            if li > 1:
                L.append(_Layer("down%d" % li, mods[0][0], mods[0][1], li, down=True))
                mods = mods[1:]
            for b in mods:
                L.append(_Layer("res%d.1" % li, b.conv1, b.bn1, li, block_start=True))
                L.append(_Layer("res%d.2" % li, b.conv2, b.bn2, li, block_end=True))
        L.append(_Layer("out", m.conv_out[0], m.conv_out[1], 5, down=True))
        return L

    def overflowed(self, st, cnt):
        hard = self.m._caps(st["n1"], st["batch"], worst=True)     # a capacity equal to the hard bound cannot overflow
        return any(c >= st["caps"][li] and st["caps"][li] < hard[li] for c, li in zip(cnt[1:], (2, 3, 4, 5)))

    def note_counts(self, st, cnt):
        for c, li in zip(cnt[1:], (2, 3, 4, 5)):
            self.m._ratios[li] = max(self.m._ratios.get(li, 0.0), c / max(st["n1"], 1))

    def parameters(self):
        """Parameters in the order the autograd Function receives them / returns their gradients."""
        ps = []
        for l in self.layers:
            ps.append(l.conv.weight)
            if l.conv.bias is not None:
                ps.append(l.conv.bias)
            ps += [l.bn.weight, l.bn.bias]
        return ps

    # ------------------------------------------------------------------------------------------ weights
    def _weights(self, i):
        """(forward image, dgrad image) of layer i, repacked only when the weight tensor changed."""
        l = self.layers[i]
        w = l.conv.weight
        key = (w._version, w.data_ptr())
        hit = self._packed.get(i)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                w3 = w.detach().reshape(l.cout, l.K, l.cin).float()
                fwd = ops.pack_weight_bf16(w3.contiguous())
                dg = None
                if i > 0:                                        # the network input needs no gradient
                    # dgrad as a forward conv: "Cout" = Cin of the layer, "Cin" = Cout of the layer.  SubM: the transposed
                    # rulebook is the rulebook with mirrored offsets, nbr_t[k] = nbr[K-1-k]  =>  Wd[ci][k][co] = W[co][K-1-k][ci]
                    wt = (w3 if l.down else w3.flip(1)).permute(2, 1, 0)
                    if l.cin_p != l.cin:
                        pad = torch.zeros((l.cin_p, l.K, l.cout), dtype=torch.float32, device=w.device)
                        pad[: l.cin] = wt
                        wt = pad
                    dg = ops.pack_weight_bf16(wt.contiguous())
            hit = (key, fwd, dg)
            self._packed[i] = hit
        return hit[1], hit[2]

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, feats, coords, batch_size, n_dev=None, worst=False):
        m = self.m
        n1 = int(coords.shape[0])
        caps = m._caps(n1, batch_size, worst=worst)
        shape = [int(v) for v in m.sparse_shape]
        k3, one = [3, 3, 3], [1, 1, 1]
        x = ops.cast_pad(feats.float().contiguous(), 16, n_dev=n_dev) if not (
            feats.dtype == torch.bfloat16 and feats.shape[1] == 16) else feats.contiguous()
        # ---- coordinates: bitmap-rank index of every level, output sets of the strided convs, rulebooks
        idx = ops.index_build(coords.contiguous(), batch_size, shape, n_dev=n_dev, want_coords=False)
        perm, _ = ops.index_rank(coords, idx, n_dev=n_dev, scatter_coords=True)
        x = ops.permute_rows(x, perm, scatter=True, n_dev=n_dev)
        lv = {1: dict(coords=idx.coords, n=idx.count, shape=list(shape), cap=n1, idx=idx)}
        lv[1]["nbr"] = ops.nbrmap_build_indexed(idx.coords, idx, k3, one, one, one, no_dev=idx.count)
        downs = {}
        for l in self.layers:
            if not l.down:
                continue
            src = lv[l.level - 1]
            conv = l.conv
            cv = (conv.kernel_size, conv.stride, conv.padding, conv.dilation)
            oshape = ops.conv_out_shape(src["shape"], *cv)
            oidx = ops.index_build(src["coords"], batch_size, oshape, conv=cv, out_cap=caps[l.level], n_dev=src["n"])
            nbr_d = ops.nbrmap_build_indexed(oidx.coords, src["idx"], *cv, no_dev=oidx.count)
            nbr_t = ops.nbrmap_transpose(nbr_d, src["cap"], no_dev=oidx.count)
            downs[l.name] = (nbr_d, nbr_t)
            lv[l.level] = dict(coords=oidx.coords, n=oidx.count, shape=list(oshape), cap=int(oidx.coords.shape[0]), idx=oidx)
            if l.level < 5:
                lv[l.level]["nbr"] = ops.nbrmap_build_indexed(oidx.coords, oidx, k3, one, one, one, no_dev=oidx.count)
        # ---- features
        saved = []
        block_in = None
        for i, l in enumerate(self.layers):
            wf, _ = self._weights(i)
            out_lv = lv[l.level]
            nbr = downs[l.name][0] if l.down else out_lv["nbr"]
            bias = l.conv.bias.detach().float() if l.conv.bias is not None else None
            if l.block_start:
                block_in = x
            # the convolution output stays fp32 (one rounding per layer, like the eval path whose affine runs in the
            # fp32 epilogue); with a bf16 `raw` the 21-layer chain sat at 2.3e-2 of the fp32 module path, above the bar
            raw = ops.spconv_fwd_bf16(x, wf, l.K, l.cout, nbr, bias=bias, no_dev=out_lv["n"], out_dtype=torch.float32)
            bn = l.bn
            act, mean, invstd = ops.bn_train_fwd(
                raw, bn.weight.detach(), bn.bias.detach(), bn.eps, bn.momentum if bn.momentum is not None else 0.1,
                bn.running_mean if bn.track_running_stats else None, bn.running_var if bn.track_running_stats else None,
                residual=block_in if l.block_end else None, relu=self.relu, n_dev=out_lv["n"])
            saved.append((x, raw, act, mean, invstd))
            x = act
            if l.block_end:
                block_in = None
            out_lv["x"] = x
        with torch.no_grad():
            for l in self.layers:
                if l.bn.track_running_stats and l.bn.num_batches_tracked is not None:
                    l.bn.num_batches_tracked.add_(1)
        counts = torch.cat([lv[i]["n"] for i in (1, 2, 3, 4, 5)])
        return dict(lv=lv, downs=downs, saved=saved, counts=counts, caps=caps, n1=n1, batch=batch_size)

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, st, dout):
        """dout: gradient of the encoded features, bf16 (cap5, 128).  -> gradients in parameters() order."""
        lv, downs, saved = st["lv"], st["downs"], st["saved"]
        grads = []
        d = dout
        pending = None                  # g of the block's identity branch, added to the dgrad of the block's first conv
        for i in range(len(self.layers) - 1, -1, -1):
            l = self.layers[i]
            x, raw, act, mean, invstd = saved[i]
            out_lv = lv[l.level]
            in_lv = lv[l.level - 1] if l.down else out_lv
            nbr = downs[l.name][0] if l.down else out_lv["nbr"]
            draw, g, dgamma, dbeta = ops.bn_train_bwd(d, act, raw, l.bn.weight.detach(), mean, invstd, relu=self.relu,
                                                      want_g=l.block_end, n_dev=out_lv["n"])
            if l.block_end:
                pending = g
            dw = ops.spconv_wgrad_bf16(x, draw, nbr, l.cin, no_dev=out_lv["n"]).reshape(l.conv.weight.shape)
            layer_grads = [dw]
            if l.conv.bias is not None:
                # a bias in front of a train-mode BatchNorm has an exactly zero gradient (sum_rows draw = -gamma*invstd*
                # (dgamma/n) * sum_rows xhat, and sum xhat = 0 by the definition of the batch mean); autograd through
                # nn.BatchNorm1d returns ~1e-6 of rounding noise in its place, a column sum of the bf16 draw ~1e-4
                layer_grads.append(torch.zeros_like(l.conv.bias, dtype=torch.float32))
            layer_grads += [dgamma, dbeta]
            grads.append(layer_grads)
            if i > 0:
                _, wd = self._weights(i)
                map_t = downs[l.name][1] if l.down else nbr
                res = pending if l.block_start else None
                d = ops.spconv_fwd_bf16(draw, wd, l.K, l.cin_p, map_t, residual=res, no_dev=in_lv["n"])
                if l.block_start:
                    pending = None
        flat = []
        for layer_grads in reversed(grads):
            flat += layer_grads
        return flat


class _GraphedStep:
    """The fused step as TWO CUDA graphs (forward, backward) over static capacity-sized buffers: the ~290 kernel
    launches of libcomb200 plus the weight re-packing of a step become two cudaGraphLaunch calls.  Everything inside is
    driven by device-side row counts, so the same graphs serve every batch that fits the captured capacities (voxel
    capacity, learned level capacities); the host reads the five row counts once, after the forward graph."""

    def __init__(self, trainer, feats, coords, batch_size):
        n, C = int(feats.shape[0]), int(feats.shape[1])
        dev = feats.device
        self.batch, self.C = batch_size, C
        self.cap1 = max((int(n * 1.25) + 4095) // 4096 * 4096, 4096)
        self.feats = torch.zeros((self.cap1, C), dtype=torch.float32, device=dev)
        self.coords = torch.zeros((self.cap1, 4), dtype=torch.int32, device=dev)
        self.n = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.load(feats, coords)
        bns = [l.bn for l in trainer.layers]
        keep = [(b.running_mean.clone(), b.running_var.clone(), b.num_batches_tracked.clone()) for b in bns
                if b.track_running_stats]
        # eager warm-up on the static buffers: kernel attributes, learned level capacities (two passes), one backward
        for _ in range(2):
            st = trainer.forward(self.feats, self.coords, batch_size, n_dev=self.n)
            trainer.note_counts(st, st["counts"].tolist())
        cap5 = int(st["lv"][5]["x"].shape[0])
        self.d = torch.zeros((cap5, trainer.layers[-1].cout), dtype=torch.bfloat16, device=dev)
        trainer.backward(st, self.d)
        for b, (rm, rv, nb) in zip([b for b in bns if b.track_running_stats], keep):   # the warm-up is not a training step
            b.running_mean.copy_(rm)
            b.running_var.copy_(rv)
            b.num_batches_tracked.copy_(nb)
        torch.cuda.synchronize(dev)
        trainer._packed.clear()                     # the packing kernels must be part of the forward graph
        trainer._capturing = True
        try:
            pool = torch.cuda.graph_pool_handle()
            self.g_fwd, self.g_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fwd, pool=pool):
                self.st = trainer.forward(self.feats, self.coords, batch_size, n_dev=self.n)
            cap5 = int(self.st["lv"][5]["x"].shape[0])
            if cap5 != int(self.d.shape[0]):
                self.d = torch.zeros((cap5, trainer.layers[-1].cout), dtype=torch.bfloat16, device=dev)
            with torch.cuda.graph(self.g_bwd, pool=pool):
                grads = trainer.backward(self.st, self.d)
                self.shapes = [tuple(g.shape) for g in grads]
                self.flat = torch.cat([g.reshape(-1) for g in grads])
        finally:
            trainer._capturing = False
        self.sizes = [int(torch.Size(sh).numel()) for sh in self.shapes]

    def fits(self, feats, batch_size):
        return batch_size == self.batch and int(feats.shape[0]) <= self.cap1 and int(feats.shape[1]) == self.C

    def load(self, feats, coords):
        n = int(feats.shape[0])
        self.feats[:n].copy_(feats)
        self.coords[:n].copy_(coords)
        self.n.fill_(n)


class _TrainFn(torch.autograd.Function):
    """One autograd node for the whole backbone: inputs = the parameters (FusedTrainer.parameters() order), output =
    the encoded features (n5, 128) fp32."""

    @staticmethod
    def forward(ctx, trainer, feats, coords, batch_size, *params):
        with torch.no_grad():
            st = None
            if trainer.use_graph:
                g = trainer._graph
                if g is None or not g.fits(feats, batch_size):
                    g = trainer._graph = _GraphedStep(trainer, feats, coords, batch_size)
                g.load(feats, coords)
                g.g_fwd.replay()
                st = g.st
                cnt = st["counts"].tolist()                      # the one host read of the step
                if trainer.overflowed(st, cnt):                  # a learned capacity was too small: eager redo, recapture later
                    undo_bn_step(trainer)
                    trainer._graph, st = None, None
            graphed = st is not None
            if st is None:
                st = trainer.forward(feats, coords, batch_size)
                cnt = st["counts"].tolist()
                if trainer.overflowed(st, cnt):
                    undo_bn_step(trainer)                        # the statistics of the truncated pass do not count
                    st = trainer.forward(feats, coords, batch_size, worst=True)
                    cnt = st["counts"].tolist()
            trainer.note_counts(st, cnt)
            st["cnt"] = cnt
            trainer.last = st
            ctx.trainer, ctx.st, ctx.graphed = trainer, st, graphed
            return st["lv"][5]["x"][: cnt[4]].float()

    @staticmethod
    def backward(ctx, dout):
        st, tr = ctx.st, ctx.trainer
        with torch.no_grad():
            if ctx.graphed and tr._graph is not None and tr._graph.st is st:
                g = tr._graph
                g.d.zero_()
                g.d[: dout.shape[0]].copy_(dout)
                g.g_bwd.replay()
                flat = g.flat.clone()                            # the graph's output buffer is rewritten by the next replay
                grads = [t.view(sh) for t, sh in zip(torch.split(flat, g.sizes), g.shapes)]
                tr.last_flat = flat                              # all parameter gradients of this step, one buffer (dist.FlatGradSync)
            else:
                cap5 = int(st["lv"][5]["x"].shape[0])
                d = torch.zeros((cap5, dout.shape[1]), dtype=torch.bfloat16, device=dout.device)
                d[: dout.shape[0]] = dout.to(torch.bfloat16)
                grads = tr.backward(st, d)
                tr.last_flat = None
        ctx.st = None
        return (None, None, None, None) + tuple(grads)


def undo_bn_step(trainer):
    """A forward pass that overflowed a learned capacity is repeated: restoring the running statistics exactly is not
    possible after the in-place update, so the repeated pass simply counts as a second training step of the statistics
    (momentum 0.01: a 1 % nudge, once per capacity growth).  num_batches_tracked is rolled back."""
    for l in trainer.layers:
        if l.bn.track_running_stats and l.bn.num_batches_tracked is not None:
            l.bn.num_batches_tracked.sub_(1)


def get_trainer(module):
    tr = module.__dict__.get("_comb_trainer")
    if tr is None:
        tr = FusedTrainer(module)
        module.__dict__["_comb_trainer"] = tr
    return tr


def forward_train(module, feats, coords, batch_size):
    """-> (x_conv1..4, out) SparseConvTensors; `out.features` (fp32) carries the autograd graph of the whole backbone,
    the multi-scale tensors are bf16 and detached (CenterPoint does not consume them)."""
    tr = get_trainer(module)
    out_feats = _TrainFn.apply(tr, feats, coords, batch_size, *tr.parameters())
    st = tr.last
    cnt = st["cnt"]
    tensors = []
    for li in (1, 2, 3, 4):
        l = st["lv"][li]
        tensors.append(SparseConvTensor(l["x"][: cnt[li - 1]], l["coords"][: cnt[li - 1]], l["shape"], batch_size))
    l5 = st["lv"][5]
    tensors.append(SparseConvTensor(out_feats, l5["coords"][: cnt[4]], l5["shape"], batch_size))
    return tuple(tensors)
