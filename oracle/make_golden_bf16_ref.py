"""TEST INFRASTRUCTURE ONLY -- how far does the UNMODIFIED REFERENCE itself move when it runs under stock bf16 autocast?

    python -m oracle.make_golden_bf16_ref        (build container only: imports /root/reference through oracle/ref_shim.py)

For every whole-model golden case the reference forward is repeated under `torch.autocast("cpu", dtype=torch.bfloat16)` (the
Trainer's / a user's stock mixed-precision switch, trainer.py:69,80) and the error against the committed fp32 golden output is
stored in tests/golden/bf16_autocast_ref.json: {case: {max_abs, rms}}.  tests/test_gpu_models.py bounds the bf16 tensor-core
path of this repo by a multiple of these numbers -- a bound that can fail, unlike a PSNR against a random ground truth."""
import json
import os

import numpy as np
import torch

from . import synth
from .ref_shim import import_reference

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main() -> None:
    torch.set_num_threads(8)
    ref = import_reference()
    models = ref.models
    with open(os.path.join(GOLD, "meta.json")) as f:
        cases = json.load(f)["cases"]
    builders = {
        "swinir": lambda c: (models.SwinIR(drop_path_rate=0.0, **c["cfg"]), synth.swinir_weights(c["cfg"], c["wseed"])),
        "hat": lambda c: (models.HAT(drop_path_rate=0.0, **c["cfg"]), synth.hat_weights(c["cfg"], c["wseed"])),
        "edsr": lambda c: (models.EDSR(**c["cfg"]), synth.edsr_weights(c["cfg"], c["wseed"])),
        "rcan": lambda c: (models.RCAN(**c["cfg"]), synth.rcan_weights(c["cfg"], c["wseed"])),
    }
    out = {}
    for name, c in sorted(cases.items()):
        fam = name.split("_")[0]
        if fam not in builders or "shape" not in c:
            continue
        m, P = builders[fam](c)
        m.load_state_dict(P, strict=True)
        m.train(bool(c.get("training", False)))
        x = synth.image_batch(tuple(c["shape"]), c["xseed"])
        gold = torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["y"])
        with torch.no_grad():
            y32 = m(x)
            assert (y32 - gold).abs().max().item() < 1e-5, name  # this really is the golden's model / input
            with torch.autocast("cpu", dtype=torch.bfloat16):
                y16 = m(x).float()
        e = (y16 - gold).double()
        out[name] = dict(max_abs=float(e.abs().max()), rms=float(e.pow(2).mean().sqrt()))
        print(name, out[name], flush=True)
    with open(os.path.join(GOLD, "bf16_autocast_ref.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, how="reference forward under torch.autocast('cpu', dtype=torch.bfloat16) vs fp32 golden",
                       cases=out), f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
