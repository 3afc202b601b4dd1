"""Developer aid: bf16 tensor-core path vs the reference's fp32 golden outputs -- max-abs, rms and the PSNR delta against a
realistic (~32 dB) synthetic ground truth, for every whole-model golden case.  Run on a GPU box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import synth
from studiosr_b200 import models as M

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
cases = json.load(open(os.path.join(GOLD, "meta.json")))["cases"]
refs = {}
p = os.path.join(GOLD, "bf16_autocast_ref.json")
if os.path.exists(p):
    refs = json.load(open(p))["cases"]
W = {"swinir": synth.swinir_weights, "hat": synth.hat_weights, "edsr": synth.edsr_weights, "rcan": synth.rcan_weights}
C = {"swinir": M.SwinIR, "hat": M.HAT, "edsr": M.EDSR, "rcan": M.RCAN}
for name, c in sorted(cases.items()):
    fam = name.split("_")[0]
    if fam not in W or "shape" not in c:
        continue
    kw = dict(c["cfg"])
    if fam in ("swinir", "hat"):
        kw["drop_path_rate"] = 0.0
    m = C[fam](**kw)
    m.load_state_dict(W[fam](c["cfg"], c["wseed"]), strict=True)
    m = m.cuda().train(bool(c.get("training", False)))
    m.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    x = synth.image_batch(tuple(c["shape"]), c["xseed"]).cuda()
    with torch.no_grad():
        y = m(x).float().cpu()
    ref = torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["y"])
    e = (y - ref).double()
    g = torch.Generator().manual_seed(99)
    gt = (ref + 0.025 * torch.randn(ref.shape, generator=g)).clip(0, 1)
    q = lambda t: (t * 255.0).round().clip(0, 255).double()
    ps = lambda a: 10 * torch.log10(255.0 ** 2 / ((q(a) - q(gt)) ** 2).mean()).item()
    r = refs.get(name, {})
    print(f"{name:34s} max_abs {e.abs().max():.3e} rms {e.pow(2).mean().sqrt():.3e} | ref autocast max_abs {r.get('max_abs', float('nan')):.3e} "
          f"rms {r.get('rms', float('nan')):.3e} | PSNR(ref,gt) {ps(ref):.2f} dB, delta {abs(ps(y) - ps(ref)):.4f} dB", flush=True)
