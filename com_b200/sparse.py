"""Host-side mirror of the `spconv.pytorch` surface the reference uses, on top of libcomb200.

Names, constructor arguments and behaviour follow what the reference calls
(pcdet/utils/spconv_utils.py:3-34, pcdet/models/backbones_3d/spconv_backbone.py:8-66,183-293):
SparseConvTensor, SparseModule, SparseSequential, SubMConv3d, SparseConv3d, SparseInverseConv3d and
conv.SparseConvolution.  Weights use the spconv-2.x layout (Cout, kz, ky, kx, Cin) so checkpoints
load through the reference's own adapter (pcdet/models/detectors/detector3d_template.py:337-348).

Row-order contract: SubM keeps the input row order; a strided SparseConv3d emits its output rows in
ascending order of the linear key ((b*D+z)*H+y)*W+x (canonical order, SURVEY hard part 1).
"""
import math
import os

import torch
from torch import nn

from . import ops


class _Config:
    # "bf16" (default): tensor-core path (bf16 operands, fp32 accumulate; north_star tolerance 2e-2): no-grad forward
    #         passes, and — with autograd — forward, dgrad (the same gather-GEMM over the transposed rulebook with W^T)
    #         and wgrad on the tcgen05 kernels.  Channel counts outside {16,32,64,128} fall back to the fp32 kernels.
    # "f32":  fp32 check mode (CUDA cores, fixed summation order; tolerance 1e-4) — COMB200_COMPUTE=f32.
    compute = os.environ.get("COMB200_COMPUTE", "bf16")
    # wgrad of the "bf16" training form: "bf16" = tcgen05 kernel (conv_wgrad.cu), "f32" = fp32 check kernel
    wgrad = os.environ.get("COMB200_WGRAD", "bf16")
    # f4 — the 2D backbone (BaseBEVBackbone) on bf16 NHWC: "bf16" makes the reference's HeightCompression hand it the BEV
    # image as channels-last bf16 (comb_dense_scatter_nhwc_bf16) and runs its forward under bf16 autocast; "f32"
    # (default) is the reference's fp32 NCHW tensor.  COMB200_BEV=bf16.
    bev = os.environ.get("COMB200_BEV", "f32")


config = _Config()


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return [int(x) for x in v]
    return [int(v)] * 3


class Rulebook:
    """Gather-form rulebook of one (indice_key) convolution."""

    def __init__(self, nbr, in_indices, out_indices, in_shape, out_shape, ksize, stride, pad, dil, subm):
        self.nbr = nbr                      # (K, No) int32, input row or -1
        self.in_indices = in_indices
        self.out_indices = out_indices
        self.in_shape = list(in_shape)
        self.out_shape = list(out_shape)
        self.ksize, self.stride, self.pad, self.dil, self.subm = ksize, stride, pad, dil, subm
        self._nbr_t = None
        self._pairs = None

    @property
    def nbr_t(self):                        # (K, Ni) int32, output row or -1 (scatter form)
        if self._nbr_t is None:
            self._nbr_t = ops.nbrmap_transpose(self.nbr, int(self.in_indices.shape[0]))
        return self._nbr_t

    def pairs(self):
        """spconv-style (indice_pairs (2,K,No), indice_pair_num (K,))."""
        if self._pairs is None:
            self._pairs = ops.nbrmap_to_pairs(self.nbr)
        return self._pairs


class SparseConvTensor:
    """Mirror of spconv.pytorch.SparseConvTensor (attributes used by the reference:
    features, indices, spatial_shape, batch_size, indice_dict, replace_feature, dense)."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None,
                 benchmark=False, **kwargs):
        if indices.dtype != torch.int32:
            indices = indices.int()
        self.features = features
        self.indices = indices.contiguous()
        self.spatial_shape = [int(x) for x in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self.grid = grid
        self.voxel_num = voxel_num
        self.benchmark = benchmark
        self._tables = kwargs.get("_tables", {})   # id(indices storage) -> (table, slots)

    def replace_feature(self, feature):
        t = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid, self.voxel_num,
                             self.indice_dict, self.benchmark, _tables=self._tables)
        return t

    def shadow_copy(self):
        return self.replace_feature(self.features)

    @property
    def spatial_size(self):
        return int(torch.tensor(self.spatial_shape).prod())

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key, None)

    def _hash_table(self):
        key = (self.indices.data_ptr(), int(self.indices.shape[0]))
        if key not in self._tables:
            self._tables[key] = ops.hash_build(self.indices, self.batch_size, self.spatial_shape)
        return self._tables[key]

    def dense(self, channels_first=True):
        # autograd-aware: gradients of the 2D backbone / CenterHead / COMLoss flow back to the rows through
        # comb_dense_gather (the reference trains through encoded_spconv_tensor.dense(), height_compression.py:21)
        if torch.is_grad_enabled() and self.features.requires_grad:
            out = ops.DenseFunction.apply(self.features, self.indices, self.batch_size, self.spatial_shape)
        else:
            out = ops.dense(self.features.contiguous(), self.indices, self.batch_size, self.spatial_shape)
        if not channels_first:
            out = out.permute(0, 2, 3, 4, 1).contiguous()
        return out

    def dense_bev_bf16(self):
        """f4: HeightCompression's (N, C*D, H, W) image as channels-last bf16 (height_compression.py:21-24 in one kernel)."""
        if torch.is_grad_enabled() and self.features.requires_grad:
            return ops.DenseNHWCFunction.apply(self.features, self.indices, self.batch_size, self.spatial_shape)
        return ops.dense_nhwc_bf16(self.features.contiguous(), self.indices, self.batch_size, self.spatial_shape)

    @property
    def sparity(self):
        return self.indices.shape[0] / (self.spatial_size * self.batch_size)


class SparseModule(nn.Module):
    """Marker base class: modules that take and return a SparseConvTensor."""
    pass


def is_spconv_module(m):
    return isinstance(m, SparseModule)


class SparseSequential(SparseModule):
    """nn.Sequential that applies plain nn.Modules to `.features` (spconv_backbone.py:21-25,187)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        from collections import OrderedDict
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        it = iter(self._modules.values())
        for _ in range(idx):
            next(it)
        return next(it)

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        for k, module in self._modules.items():
            if is_spconv_module(module):
                assert isinstance(input, SparseConvTensor)
                input = module(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input = input.replace_feature(module(input.features))
            else:
                input = module(input)
        return input


class _SpConvFunction(torch.autograd.Function):
    """out[o] = sum_k in[nbr[k][o]] @ W[:,k,:]^T (+bias); backward = dgrad over the transposed
    rulebook + wgrad (fp32 kernels)."""

    @staticmethod
    def forward(ctx, feats, weight, bias, rb, gather_map, scatter_map_fn):
        Cout, Cin = weight.shape[0], weight.shape[-1]
        w3 = weight.reshape(Cout, -1, Cin).contiguous()
        feats = feats.contiguous()
        out = ops.spconv_fwd_f32(feats, w3, gather_map, bias=bias)
        ctx.save_for_backward(feats, weight)
        ctx.gather_map, ctx.scatter_map_fn, ctx.has_bias = gather_map, scatter_map_fn, bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        feats, weight = ctx.saved_tensors
        Cout, Cin = weight.shape[0], weight.shape[-1]
        w3 = weight.reshape(Cout, -1, Cin).contiguous()
        dout = dout.contiguous().float()
        din = dw = db = None
        if ctx.needs_input_grad[0]:
            din = ops.spconv_dgrad_f32(dout, w3, ctx.scatter_map_fn())
        if ctx.needs_input_grad[1]:
            dw = ops.spconv_wgrad_f32(feats, dout, ctx.gather_map).reshape(weight.shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dout.sum(0)
        return din, dw, db, None, None, None


class _SpConvFunctionBF16(torch.autograd.Function):
    """Mixed-precision training form (config.compute == "bf16"): forward, dgrad and wgrad all run on the tcgen05
    kernels with bf16 operands and fp32 accumulation.  dgrad is the forward gather-GEMM over the transposed rulebook
    with the transposed weights, din[i] = sum_k dout[nbr_t[k][i]] @ W[:,k,:]; wgrad reduces over the output rows with
    the gathered rows as MN-major operands (conv_wgrad.cu).  The bf16 image of the input is what is saved for the
    backward pass.  Outputs and gradients are fp32 tensors.  config.wgrad = "f32" keeps wgrad on the fp32 check
    kernel (fp32 features and gradients)."""

    @staticmethod
    def forward(ctx, feats, weight, bias, rb, gather_map, scatter_map_fn):
        Cout, Cin = weight.shape[0], weight.shape[-1]
        w3 = weight.detach().reshape(Cout, -1, Cin).contiguous().float()
        feats = feats.contiguous()
        xb = ops.cast_pad(feats, ops.pad16(Cin))
        out = ops.spconv_fwd_bf16(xb, ops.pack_weight_bf16(w3), int(w3.shape[1]), Cout, gather_map,
                                  bias=bias.detach().float() if bias is not None else None, out_dtype=torch.float32)
        ctx.wgrad_f32 = config.wgrad == "f32"
        ctx.save_for_backward(feats if ctx.wgrad_f32 else xb, weight)
        ctx.gather_map, ctx.scatter_map_fn, ctx.has_bias = gather_map, scatter_map_fn, bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        feats, weight = ctx.saved_tensors
        Cout, Cin = weight.shape[0], weight.shape[-1]
        w3 = weight.detach().reshape(Cout, -1, Cin).contiguous().float()
        K = int(w3.shape[1])
        dout = dout.contiguous().float()
        doutb = ops.cast_pad(dout, Cout)
        din = dw = db = None
        if ctx.needs_input_grad[0]:
            cin_p = ops.pad16(Cin)                       # the kernel's N must be 16 / 32 / 64 / 128: zero rows beyond Cin
            if cin_p == Cin:
                wt = w3.permute(2, 1, 0).contiguous()    # W^T as (Cin, K, Cout): one copy kernel
            else:
                wt = torch.zeros((cin_p, K, Cout), dtype=torch.float32, device=w3.device)
                wt[:Cin] = w3.permute(2, 1, 0)
            din = ops.spconv_fwd_bf16(doutb, ops.pack_weight_bf16(wt), K, cin_p, ctx.scatter_map_fn(),
                                      out_dtype=torch.float32)[:, :Cin].contiguous()
        if ctx.needs_input_grad[1]:
            if ctx.wgrad_f32:
                dw = ops.spconv_wgrad_f32(feats, dout, ctx.gather_map).reshape(weight.shape)
            else:
                dw = ops.spconv_wgrad_bf16(feats, doutb, ctx.gather_map, Cin).reshape(weight.shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dout.sum(0)
        return din, dw, db, None, None, None


class SparseConvolution(SparseModule):
    """Mirror of spconv.pytorch.conv.SparseConvolution for ndim=3."""

    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 algo=None, fp32_accum=None, record_voxel_count=False, act_type=None, act_alpha=0, act_beta=0,
                 large_kernel_fast_algo=False, name=None, **kwargs):
        super().__init__()
        assert ndim == 3, "only 3D sparse convolution is on the COM hot path"
        assert groups == 1, "groups != 1 is not supported"
        assert not transposed, "SparseConvTranspose3d is not on the COM hot path"
        self.ndim = ndim
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size = _triple(kernel_size)
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.dilation = _triple(dilation)
        self.output_padding = _triple(output_padding)
        self.conv1x1 = all(k == 1 for k in self.kernel_size) and all(s == 1 for s in self.stride)
        self.subm, self.inverse, self.transposed = subm, inverse, transposed
        self.groups = groups
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(self.out_channels, *self.kernel_size, self.in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(self.out_channels))
        else:
            self.register_parameter("bias", None)
        self._packed = None   # (weight version, packed bf16 image)
        self.reset_parameters()

    def extra_repr(self):
        s = "{in_channels}, {out_channels}, kernel_size={kernel_size}, stride={stride}, padding={padding}"
        if self.bias is None:
            s += ", bias=False"
        return s.format(**self.__dict__) + ", subm=%s, indice_key=%s" % (self.subm, self.indice_key)

    def reset_parameters(self):
        # same scheme as spconv / torch conv: kaiming_uniform(a=sqrt(5)) with fan_in = K*Cin
        K = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        fan_in = K * self.in_channels
        gain = math.sqrt(2.0 / (1 + 5.0))
        bound = gain * math.sqrt(3.0 / fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                b = 1 / math.sqrt(fan_in)
                self.bias.uniform_(-b, b)

    # ---- rulebook -------------------------------------------------------------------------------
    def _rulebook(self, input):
        rb = input.find_indice_pair(self.indice_key)
        if self.inverse:
            assert rb is not None and self.indice_key is not None, "inverse conv needs the forward conv's indice_key"
            return rb
        if rb is not None and self.subm:
            return rb
        if rb is not None and not self.subm:
            raise RuntimeError("indice_key %r is already used by another non-submanifold conv" % self.indice_key)
        k, d = self.kernel_size, self.dilation
        table, slots = input._hash_table()
        if self.subm:
            pad = [(kk // 2) * dd for kk, dd in zip(k, d)]
            stride = [1, 1, 1]
            out_shape = input.spatial_shape
            out_indices = input.indices
        else:
            pad, stride = self.padding, self.stride
            out_shape = ops.conv_out_shape(input.spatial_shape, k, stride, pad, d)
            n_in = int(input.indices.shape[0])
            vol = input.batch_size * out_shape[0] * out_shape[1] * out_shape[2]
            contrib = 1
            for kk, ss in zip(k, stride):
                contrib *= (kk + ss - 1) // ss
            cap = min(vol, n_in * contrib)
            out_indices, cnt = ops.conv_out_coords(input.indices, input.batch_size, out_shape, k, stride, pad, d, cap)
            out_indices = out_indices[: int(cnt.item())]
        nbr = ops.nbrmap_build(out_indices, table, slots, input.batch_size, input.spatial_shape, k, stride, pad, d)
        rb = Rulebook(nbr, input.indices, out_indices, input.spatial_shape, out_shape, k, stride, pad, d, self.subm)
        if self.indice_key is not None:
            input.indice_dict[self.indice_key] = rb
        return rb

    def _packed_weight(self):
        v = self.weight._version
        if self._packed is None or self._packed[0] != v or self._packed[1].device != self.weight.device:
            w3 = self.weight.detach().reshape(self.out_channels, -1, self.in_channels).contiguous().float()
            self._packed = (v, ops.pack_weight_bf16(w3))
        return self._packed[1]

    def forward(self, input):
        assert isinstance(input, SparseConvTensor)
        feats = input.features
        if not feats.is_cuda:
            raise RuntimeError("libcomb200 sparse convolution needs CUDA tensors (no CPU fallback)")
        rb = self._rulebook(input)
        if self.inverse:
            gather_map, scatter_fn = rb.nbr_t, (lambda: rb.nbr)
            out_indices, out_shape = rb.in_indices, rb.in_shape
        else:
            gather_map, scatter_fn = rb.nbr, (lambda: rb.nbr_t)
            out_indices, out_shape = rb.out_indices, rb.out_shape
        K = gather_map.shape[0]
        needs_grad = torch.is_grad_enabled() and (feats.requires_grad or self.weight.requires_grad)
        use_tc = (config.compute == "bf16" or feats.dtype == torch.bfloat16) and not needs_grad \
            and self.out_channels in (16, 32, 64, 128) and self.in_channels <= 128
        if use_tc:
            cin_p = ops.pad16(self.in_channels)
            if feats.dtype == torch.bfloat16 and feats.shape[1] == cin_p:
                xb = feats.contiguous()
            else:
                xb = ops.cast_pad(feats.float().contiguous(), cin_p)
            bias = self.bias.detach().float() if self.bias is not None else None
            out = ops.spconv_fwd_bf16(xb, self._packed_weight(), K, self.out_channels, gather_map, bias=bias,
                                      out_dtype=feats.dtype if feats.dtype == torch.bfloat16 else torch.float32)
        elif needs_grad and config.compute == "bf16" and self.out_channels in (16, 32, 64, 128) \
                and ops.pad16(self.in_channels) in (16, 32, 64, 128):
            out = _SpConvFunctionBF16.apply(feats.float(), self.weight, self.bias, rb, gather_map, scatter_fn)
        else:
            out = _SpConvFunction.apply(feats.float(), self.weight, self.bias, rb, gather_map, scatter_fn)
        res = SparseConvTensor(out, out_indices, out_shape, input.batch_size, input.grid, input.voxel_num,
                               input.indice_dict, input.benchmark,
                               _tables=input._tables if (self.subm or self.inverse) else {})
        return res


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, fp32_accum=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key, **kwargs)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, fp32_accum=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key, **kwargs)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True, algo=None, fp32_accum=None,
                 **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key,
                         **kwargs)


class ToDense(SparseModule):
    def forward(self, x):
        return x.dense()
