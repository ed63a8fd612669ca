"""CPU: host-side logic of the drop-in package — constructors, state_dict layout, synthetic inputs, sharding,
and the "fail loudly without the CUDA extension / device" contract."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from conftest import GOLD, build_mm, build_tm

ROOT = Path(__file__).resolve().parent.parent

EXPECTED_KEYS_MM = [
    "encoder.embedding.weight", "encoder.gru.weight_ih_l0", "encoder.gru.weight_hh_l0", "encoder.gru.bias_ih_l0",
    "encoder.gru.bias_hh_l0", "encoder.gru.weight_ih_l0_reverse", "encoder.gru.weight_hh_l0_reverse",
    "encoder.gru.bias_ih_l0_reverse", "encoder.gru.bias_hh_l0_reverse", "decoder.embedding.weight",
    "decoder.gru_1.weight_ih_l0", "decoder.gru_1.weight_hh_l0", "decoder.gru_1.bias_ih_l0", "decoder.gru_1.bias_hh_l0",
    "decoder.attn.v", "decoder.attn.attn_h.weight", "decoder.attn.attn_e.weight", "decoder.context2hid.weight",
    "decoder.gru_2.weight_ih_l0", "decoder.gru_2.weight_hh_l0", "decoder.gru_2.bias_ih_l0", "decoder.gru_2.bias_hh_l0",
    "decoder.W1.weight", "decoder.W1.bias", "decoder.W2.weight", "decoder.W2.bias", "decoder.W3.weight", "decoder.W3.bias",
    "decoder.out.weight", "decoder.out.bias", "vse_imagine.imagine_attn.ctx2ctx.weight",
    "vse_imagine.imagine_attn.emb2ctx.weight", "vse_imagine.im_embedding.weight", "vse_imagine.im_embedding.bias",
    "vse_imagine.text_embedding.weight", "vse_imagine.text_embedding.bias", "decoderini.weight", "decoderini.bias"]


def test_state_dict_layout_and_param_counts():
    from vag_nmt_b200 import synthetic
    mm = build_mm(synthetic.DE, 1234)
    assert list(mm.state_dict().keys()) == EXPECTED_KEYS_MM
    assert sum(p.numel() for p in mm.parameters()) == 15_944_623                  # probed on the reference
    assert mm.decoder.out.weight is mm.decoder.embedding.weight                   # tied (NMT_Decoder.py:105-106)
    assert float(mm.decoder.embedding.weight[0].abs().sum()) > 1.0                # quirk 1: pad row is NOT zero
    assert float(mm.decoder.W1.bias.abs().sum()) == 0.0 and float(mm.decoder.out.bias.abs().sum()) == 0.0
    assert mm.vse_imagine.dropout_im_emb == 0.0 and mm.shared_embedding_size == 512
    tm = build_tm(synthetic.DE, 1235)
    assert sum(p.numel() for p in tm.parameters()) == 12_797_871
    assert [k for k in tm.state_dict() if k.startswith("vse")] == []


def test_init_matches_reference_checksums(full_de):
    mm = build_mm(full_de["cfg"], full_de["seed"])
    for k, v in mm.state_dict().items():
        s, a = full_de["param_checksums"]["mm"][k]
        assert abs(float(v.double().sum()) - s) <= 1e-9 * max(1.0, abs(a)), k


def test_synthetic_inputs_are_deterministic_and_well_formed():
    from vag_nmt_b200 import synthetic
    a = synthetic.make_batch(32, 8507, 9391, 2048, seed=7)
    b = synthetic.make_batch(32, 8507, 9391, 2048, seed=7)
    assert torch.equal(a.src, b.src) and torch.equal(a.tgt, b.tgt) and torch.equal(a.im, b.im)
    assert a.src_lengths == sorted(a.src_lengths, reverse=True) and a.src.shape[1] == a.src_lengths[0]
    for r, L in enumerate(a.src_lengths):
        assert int(a.src[r, L - 1]) == 3 and (a.src[r, :L - 1] >= 4).all() and (a.src[r, L:] == 0).all()
    assert (a.tgt != 0).sum(1).unique().numel() == 1         # bucketed: one target length per training batch
    sents, im = synthetic.make_corpus(50, 8507, 2048, seed=7)
    src, lens, im_s, order = synthetic.pad_and_sort(sents, im)
    assert sorted(order) == list(range(50)) and lens == sorted(lens, reverse=True)
    assert all(src[r, :lens[r]].tolist() == sents[order[r]] for r in range(50)) and torch.equal(im_s[3], im[order[3]])


def test_shard_ranges_cover_exactly():
    from vag_nmt_b200.translate import shard_range
    for n in (0, 1, 7, 1000, 1014):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_decode_corpus_restores_order_with_oracle_decoder():
    """Host batching + un-sort (preprocessing.py:234-306, :475-486) with the CPU oracle standing in for the model."""
    from oracle import vag_oracle as O
    from vag_nmt_b200 import synthetic
    from vag_nmt_b200.translate import decode_corpus
    from conftest import cpu_params
    cfg = synthetic.TINY
    p = cpu_params(build_mm(cfg, 3))
    sents, im = synthetic.make_corpus(11, cfg["src_size"], cfg["im_feats_size"], seed=2, max_len=9, min_len=1, mean=5, std=3)
    fn = lambda src, lens, im_b, K, L: O.multimodal_beamsearch_decode(p, src, lens, im_b, K, L)
    whole = decode_corpus(fn, sents, im, 3, 10)
    by4 = decode_corpus(fn, sents, im, 3, 10, batch_size=4)
    single = [decode_corpus(fn, [s], im[i:i + 1], 3, 10)[0] for i, s in enumerate(sents)]
    assert whole == by4 == single            # sentences are independent: batching never changes a translation


def test_no_cuda_means_loud_failure_not_fallback():
    import vag_nmt_b200 as vag
    from vag_nmt_b200 import _cabi, ops
    if torch.cuda.is_available():
        pytest.skip("CPU-only contract")
    with pytest.raises(_cabi.VagError):
        ops.linear(torch.zeros(2, 4), torch.zeros(3, 4))
    m = build_mm(__import__("vag_nmt_b200").synthetic.TINY, 1)
    with pytest.raises(RuntimeError):
        m.beamsearch_decode(torch.tensor([[5, 3]]), [2], torch.rand(1, 24), beam_size=2, max_length=3)
    with pytest.raises(RuntimeError):
        vag.t2i(torch.rand(3, 4), torch.rand(3, 4))


def test_missing_library_is_reported(tmp_path):
    from vag_nmt_b200 import _cabi
    with pytest.raises(_cabi.VagError, match="not found"):
        _cabi.load_library(tmp_path / "libvagnmt.so")


def test_product_never_imports_the_oracle():
    import re
    for f in (ROOT / "vag_nmt_b200").rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f       # no import of oracle/
        assert "vag_oracle" not in src and "oracle." not in src.replace("CPU oracle.", ""), f
    for f in (ROOT / "vag_nmt_b200" / "csrc").glob("*.cu*"):
        assert "oracle" not in f.read_text(), f


def test_reference_whole_module_checkpoints_load_without_the_reference(tiny_dot, tmp_path):
    """torch.save(model) files written by the REAL reference (oracle/make_golden.py; the format of nmt_multimodal_beam_DE.py:491-519)
    load into the drop-in classes with the reference package absent: same state_dict, same hyper-parameters."""
    import pickle
    import sys
    from vag_nmt_b200.checkpoint_compat import load_reference_module, load_reference_stub
    assert not any(m == "machine_translation_vision" or m.startswith("machine_translation_vision.") for m in sys.modules)
    import vag_nmt_b200 as vag
    mm = load_reference_module(GOLD / "ref_module_tiny_mm.pt")
    tm = load_reference_module(GOLD / "ref_module_tiny_tm.pt")
    assert isinstance(mm, vag.NMT_AttentionImagine_Seq2Seq_Beam_V11) and isinstance(tm, vag.NMT_Seq2Seq_Beam_V2)
    for model, want in ((mm, tiny_dot["params_mm"]), (tm, tiny_dot["params_tm"])):
        sd = model.state_dict()
        assert list(sd.keys()) == list(want.keys())
        for k in want:
            assert torch.equal(sd[k], want[k]), k
    cfg = tiny_dot["cfg"]
    assert (mm.hidden_size, mm.shared_embedding_size, mm.tgt_size, mm.loss_w, mm.init_split, mm.tied_emb) == \
        (cfg["hidden_size"], cfg["shared_embedding_size"], cfg["tgt_size"], 0.99, 0.5, True)
    assert mm.decoder.out.weight is mm.decoder.embedding.weight            # the tie survives the conversion
    stub = load_reference_stub(GOLD / "ref_module_tiny_mm.pt")
    assert stub._ref_class.endswith("NMT_AttentionImagine_Seq2Seq_Beam_V11.NMT_AttentionImagine_Seq2Seq_Beam_V11")
    # the resolver is an allow-list: a pickle naming anything else is refused
    evil = tmp_path / "evil.pt"
    class Boom:
        def __reduce__(self):
            import os
            return (os.system, ("true",))
    torch.save({"x": Boom()}, evil)
    with pytest.raises(pickle.UnpicklingError):
        load_reference_stub(evil)


def test_large_decode_shapes_are_captured_only_when_they_repeat_back_to_back():
    """models._capture_large_now: the graph capture of a > 2048-row decode shape waits for the same shape twice in a row (one-off
    calls and callers cycling through many large shapes keep the enqueued loop)."""
    from conftest import build_mm
    from vag_nmt_b200 import synthetic
    m = build_mm(dict(synthetic.TINY), 1)
    a, b = ("", 1000, 30, 12, 80, "fp32", True, 8, 0), ("", 1000, 31, 12, 80, "fp32", True, 8, 0)
    assert m._capture_large_now(a) is False          # first sight: enqueue
    assert m._capture_large_now(a) is True           # came straight back: capture
    assert m._capture_large_now(b) is False
    assert m._capture_large_now(a) is False          # alternating shapes never capture
    assert m._capture_large_now(b) is False
    assert m._capture_large_now(b) is True


def test_top2_epilogue_algorithm_is_canonical_on_ties(tmp_path):
    """tools/top2_crosscheck.cpp restates the per-chunk top-2 of the vocabulary kernel's epilogue (csrc/linear_tc.cu) and checks it
    against a brute-force canonical top-2 on random chunks full of exact ties; the insertion chains it replaced fail the same check
    (the defect this harness found).  The kernel itself is pinned on the GPU by tests/test_gpu_tc.py."""
    import shutil
    import subprocess
    from pathlib import Path
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    src = Path(__file__).resolve().parent.parent / "tools" / "top2_crosscheck.cpp"
    exe = tmp_path / "top2_crosscheck"
    subprocess.run(["g++", "-O2", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe), "1000000"], capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "new_vs_ref" and int(out[1]) == 0
    assert out[2] == "old_vs_ref" and int(out[3]) > 0
