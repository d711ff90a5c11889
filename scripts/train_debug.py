"""Per-layer comparison of the fused train step against the module path (fp32 check mode), in layer order."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from com_b200 import sparse
import test_gpu_train_fused as T

sparse.config.compute = "f32"
feats, coords = T.make_inputs()
if "--small" in sys.argv:
    keep = coords[:, 0] == 0
    feats, coords = feats[keep][:4000].contiguous(), coords[keep][:4000].contiguous()
ref_bb, fus_bb = T.build(), T.build()
fus_bb.load_state_dict(ref_bb.state_dict())
os.environ["COMB_FUSED_TRAIN"] = "0"
bd_r, sf_r, wgt = T.run(ref_bb, feats, coords)
os.environ.pop("COMB_FUSED_TRAIN")
bd_f, sf_f, _ = T.run(fus_bb, feats, coords, wgt)
torch.cuda.synchronize()
print("dense err", T.rel(sf_f.detach(), sf_r.detach()))
for (k, p), (_, q) in zip(ref_bb.named_parameters(), fus_bb.named_parameters()):
    print("%-28s ref|max| %.3e  fused|max| %.3e  err %.4f" % (k, float(p.grad.abs().max()), float(q.grad.abs().max()), T.rel(q.grad, p.grad)))
