"""One layer shape of the bench batch, launched a few times (for ncu): python scripts/conv_one.py <level> <cin> <cout>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from com_b200 import ops
import conv_trace_ts as ct

lv, cin, cout = (int(a) for a in sys.argv[1:4])
cd, idx = ct.level_coords(lv)
n = int(cd.shape[0])
nbr = ops.nbrmap_build_indexed(cd, idx, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1])
x = torch.randn((n, cin), device="cuda").to(torch.bfloat16)
w = ops.pack_weight_bf16(torch.randn((cout, 27, cin), device="cuda") / 20)
for _ in range(6):
    ops.spconv_fwd_bf16(x, w, 27, cout, nbr)
torch.cuda.synchronize()
