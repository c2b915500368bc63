"""Signed per-layer error statistics of the tensor-core encoder modes against the oracle (run on a B200).
A systematic negative mean of (ours - ref) / ref on positive activations = accumulator truncation (round toward zero);
zero-mean noise = operand rounding."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from oracle import net_oracle
from hierarchicalprobabilistic3dhuman_b200 import synthetic as syn
from test_gpu_net import make_model

NAMES = ["stem", "pool"] + [f"layer{l}.{b}" for l in range(1, 5) for b in range(2)]
B = int(os.environ.get("DIAG_B", "2"))
sd = syn.synthetic_state_dict(0)
x = torch.from_numpy(syn.synthetic_proxy_rep(B, seed=1))
taps_ref = {}
with torch.no_grad():
    feats_ref = net_oracle.encoder_forward(sd, x, taps=taps_ref)
for mode in sys.argv[1:] or ["split", "parity"]:
    m = make_model(mode)
    feats, taps = m.encode_taps(x.cuda())
    out = {"mode": mode, "env": {k: v for k, v in os.environ.items() if k.startswith("HP3D_")}}
    for name, t in list(zip(NAMES, taps)) + [("feats", feats)]:
        ref = (taps_ref[name] if name != "feats" else feats_ref).double()
        o = (t.permute(0, 3, 1, 2) if name != "feats" else t).double().cpu()
        e = o - ref
        big = ref > 0.1 * ref.max()
        r = (e[big] / ref[big])
        out[name] = {"maxrel": float(e.abs().max() / ref.abs().max()), "mean_rel_big": float(r.mean()), "std_rel_big": float(r.std()),
                     "n_big": int(big.sum())}
    print(json.dumps(out))
