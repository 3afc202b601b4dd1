"""CPU: the C-ABI library loads and exports every symbol include/ssr_b200.h declares; the
product path refuses to run without a GPU (no fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ssr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from studiosr_b200 import _lib

    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ssr_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes prototypes out of sync with the header"
    assert lib.ssr_version() == 100


def test_state_dict_contract_matches_reference_keys():
    from oracle import synth
    from studiosr_b200.models import EDSR, SwinIR

    m = SwinIR()
    sd = synth.swinir_weights(synth.swinir_config(), 0)  # key list pinned to the reference by make_golden (strict=True)
    assert list(m.state_dict().keys()) == list(sd.keys()) and len(sd) == 532
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape) and v.dtype == sd[k].dtype, k
    e = EDSR()
    sd = synth.edsr_weights(synth.EDSR_DEFAULT, 0)
    assert list(e.state_dict().keys()) == list(sd.keys()) and len(sd) == 142
    assert sum(p.numel() for p in m.parameters()) == 11900199
    assert sum(p.numel() for p in e.parameters() if p.requires_grad) == 43089923


def test_no_cpu_fallback():
    from studiosr_b200.models import SwinIR

    m = SwinIR(embed_dim=60, depths=[2], num_heads=[6])
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(1, 3, 8, 8))


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "studiosr_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in txt.replace("oracle:", "").replace("(oracle", ""), f"{f} references the oracle"


def test_model_config_roundtrip():
    from studiosr_b200.models import SwinIR

    m = SwinIR(scale=2, embed_dim=60, depths=[2, 2], num_heads=[6, 6])
    cfg = m.get_model_config()
    assert cfg["scale"] == 2 and cfg["embed_dim"] == 60 and cfg["window_size"] == 8 and cfg["upsampler"] == "pixelshuffle"
    assert m.get_training_config()["batch_size"] == 32
