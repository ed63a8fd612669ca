import json
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLD = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(GOLD / name, weights_only=False)


@pytest.fixture(scope="session")
def tiny_dot():
    return load_golden("tiny_dot.pt")


@pytest.fixture(scope="session")
def tiny_mlp():
    return load_golden("tiny_mlp.pt")


@pytest.fixture(scope="session")
def full_de():
    return load_golden("full_de_b32.pt")


@pytest.fixture(scope="session")
def beam_kat():
    return json.loads((GOLD / "beam_kat.json").read_text())


def build_mm(cfg, seed, device=None, attn_model=None):
    """Drop-in multimodal model with the reference's init under `seed` (bit-identical, see oracle/make_golden.py)."""
    import vag_nmt_b200 as vag
    torch.manual_seed(seed)
    m = vag.NMT_AttentionImagine_Seq2Seq_Beam_V11(
        cfg["src_size"], cfg["tgt_size"], cfg["im_feats_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
        cfg["hidden_size"], cfg["shared_embedding_size"], 0.99, attn_model=attn_model or cfg.get("attn_model", "dot"),
        tied_emb=True, init_split=0.5).eval()
    return m.to(device) if device is not None else m


def build_tm(cfg, seed, device=None):
    import vag_nmt_b200 as vag
    torch.manual_seed(seed)
    m = vag.NMT_Seq2Seq_Beam_V2(cfg["src_size"], cfg["tgt_size"], cfg["src_embedding_size"], cfg["tgt_embedding_size"],
                                cfg["hidden_size"], tied_emb=True).eval()
    return m.to(device) if device is not None else m


def cpu_params(model, dtype=torch.float32):
    return {k: v.detach().cpu().to(dtype) for k, v in model.state_dict().items()}


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
