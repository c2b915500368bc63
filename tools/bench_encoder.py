"""Time the tensor-core encoder (MODE=split|fast) alone at the bench batch; HP3D_CONV_DEBUG / HP3D_CONV_PATCH select experiments."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from types import SimpleNamespace as NS
import hierarchicalprobabilistic3dhuman_b200 as hp
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
B = int(os.environ.get("B", 256))
cfg = NS(MODEL=NS(NUM_IN_CHANNELS=18, NUM_RESNET_LAYERS=18, EMBED_DIM=256, DELTA_I=True, DELTA_I_WEIGHT=1.0, NUM_SMPL_BETAS=10))
net = hp.PoseMFShapeGaussianNet(syn.SMPL_PARENTS.tolist(), cfg, encoder_mode=os.environ.get("MODE", "split"))
net.load_state_dict(syn.synthetic_state_dict(0)); net = net.cuda().eval()
x = torch.from_numpy(syn.synthetic_proxy_rep(16, seed=1)).repeat(B // 16, 1, 1, 1).cuda()
for _ in range(3): net.encode(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): net.encode(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"debug": os.environ.get("HP3D_CONV_DEBUG", "0"), "patch": os.environ.get("HP3D_CONV_PATCH", "1"), "encoder_ms": ms,
                  "TFLOPs": 6.279e-3 * B / ms}))
