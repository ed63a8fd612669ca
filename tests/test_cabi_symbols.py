"""CPU: libvagnmt.so loads and exports every symbol include/vag_nmt.h declares (no compute calls here)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "vag_nmt.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vag_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for need in ("vag_encoder_fwd_f32", "vag_attn_keys_f32", "vag_decoder_step_f32", "vag_beam_decode_f32", "vag_beam_select_f32",
                 "vag_vse_pool_fwd_f32", "vag_rank_loss_f32", "vag_recall_ranks_f32", "vag_linear_f32", "vag_attention_f32"):
        assert need in syms


def test_library_exports_every_declared_symbol():
    from vag_nmt_b200 import _cabi
    from vag_nmt_b200.build import build_library
    lib_path = build_library()
    lib = ctypes.CDLL(str(lib_path))
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vag_nmt.h but not exported"
        assert s in _cabi.SIGNATURES, f"{s} has no ctypes signature in vag_nmt_b200/_cabi.py"
    assert set(_cabi.SIGNATURES) <= set(syms), "ctypes binds a symbol the header does not declare"
    l2 = _cabi.load_library()
    assert l2.vag_abi_version() == 1
    assert isinstance(l2.vag_last_error(), bytes)


def test_struct_layouts_match_header():
    """sizeof of the ctypes mirrors == the C structs (pointer-sized fields, natural alignment)."""
    from vag_nmt_b200 import _cabi
    assert ctypes.sizeof(_cabi.EncoderWeights) == 16 + 8 + 8 + 4 * 16     # E, H, precision (+pad), vocab, emb, 4 x 2 pointers
    assert ctypes.sizeof(_cabi.VseWeights) == 24 + 7 * 8          # 6 ints
    assert ctypes.sizeof(_cabi.DecoderWeights) == 16 + 8 + 23 * 8 + 16  # 4 ints, int64 V, 23 pointers, prepared + size
