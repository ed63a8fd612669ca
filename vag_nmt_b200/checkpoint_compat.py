"""Loading the reference's WHOLE-MODULE checkpoints into the drop-in classes.

The reference saves ``torch.save(imagine_model, path)`` (nmt_multimodal_beam_DE.py:491-519) and restores with
``best_model = torch.load(path)`` (:534): the pickle names the reference's own classes
(``machine_translation_vision.models.….NMT_AttentionImagine_Seq2Seq_Beam_V11`` and its layer classes), which do not exist
next to this package.  ``load_reference_module`` un-pickles such a file WITHOUT the reference installed: every
``machine_translation_vision.*`` class is resolved to an inert ``nn.Module`` stub that merely holds the pickled ``__dict__``
(parameters, sub-modules, hyper-parameter attributes); the constructor arguments are read back from those attributes and tensor
shapes, a drop-in model is built and the stub tree's ``state_dict`` — same keys, SURVEY.md section 8b — is loaded into it.

Un-pickling executes constructors named by the file, so the resolver is an allow-list: torch / collections / numpy core /
builtins needed by tensors and ``nn.Module`` state, the stubs, nothing else (a checkpoint naming any other class is refused).
"""
from __future__ import annotations

import pickle
import types
from typing import Union

import torch
import torch.nn as nn

_REF_PREFIX = "machine_translation_vision"
_ALLOWED_PREFIXES = ("torch", "collections", "numpy", "_codecs")
_ALLOWED_BUILTINS = {"set", "frozenset", "list", "dict", "tuple", "int", "float", "bool", "str", "bytes", "bytearray", "complex", "slice",
                     "range", "getattr", "object"}


class ReferenceStub(nn.Module):
    """Holds the pickled state of one reference module (any class under machine_translation_vision.*)."""

    _ref_class = "?"

    def __getattr__(self, name):
        # The reference stores BOUND METHODS of itself as attributes (ImagineAttn: ``self.score = self.score_dot``); pickle
        # restores them with getattr(instance, "score_dot") while the instance is still empty.  Anything that is not a parameter,
        # buffer or sub-module resolves to an inert placeholder — the stub only ever serves state_dict() and attribute reads.
        d = self.__dict__
        for store in ("_parameters", "_buffers", "_modules"):
            if store in d and name in d[store]:
                return d[store][name]
        if name.startswith("__") or name in ("_parameters", "_buffers", "_modules"):
            raise AttributeError(name)
        return _placeholder

    def forward(self, *a, **k):  # pragma: no cover - never called
        raise RuntimeError("a ReferenceStub only carries a reference checkpoint's state; convert it with load_reference_module")


def _placeholder(*a, **k):  # pragma: no cover - never called
    raise RuntimeError("method of a reference class: not available on a ReferenceStub")


def _stub_for(module: str, name: str):
    return type(name, (ReferenceStub,), {"_ref_class": f"{module}.{name}", "__module__": __name__})


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == _REF_PREFIX or module.startswith(_REF_PREFIX + "."):
            return _stub_for(module, name)
        if module in ("builtins", "__builtin__") and name in _ALLOWED_BUILTINS:     # protocol-2 pickles spell it __builtin__
            return super().find_class("builtins", name)
        if any(module == p or module.startswith(p + ".") for p in _ALLOWED_PREFIXES):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"reference checkpoint names {module}.{name}, which is outside the allow-list")


def _pickle_module():
    m = types.ModuleType("vag_reference_pickle")
    m.__dict__.update({k: getattr(pickle, k) for k in dir(pickle) if not k.startswith("__")})
    m.Unpickler = _Unpickler
    m.load = lambda f, **kw: _Unpickler(f, **kw).load()
    return m


def load_reference_stub(path) -> ReferenceStub:
    """The un-pickled module tree as stubs (CPU tensors)."""
    obj = torch.load(path, map_location="cpu", pickle_module=_pickle_module(), weights_only=False)
    if not isinstance(obj, ReferenceStub):
        raise TypeError(f"{path} does not hold a pickled reference module (got {type(obj).__name__})")
    return obj


def _attr(stub, name, default):
    return stub.__dict__.get(name, default)


def convert_reference_module(stub: ReferenceStub) -> nn.Module:
    """Stub tree → drop-in model (CPU, eval mode) carrying the checkpoint's weights and hyper-parameters."""
    from .models import NMT_AttentionImagine_Seq2Seq_Beam_V11, NMT_Seq2Seq_Beam_V2
    sd = stub.state_dict()
    kind = stub._ref_class.rsplit(".", 1)[-1]
    enc, dec = stub._modules["encoder"], stub._modules["decoder"]
    src_size, src_emb = sd["encoder.embedding.weight"].shape
    tgt_size, tgt_emb = sd["decoder.embedding.weight"].shape
    hidden = sd["encoder.gru.weight_hh_l0"].shape[1]
    tied = sd["decoder.out.weight"].data_ptr() == sd["decoder.embedding.weight"].data_ptr() or bool(_attr(stub, "tied_emb", False))
    common = dict(n_layers=int(_attr(stub, "n_layers", 1)), dropout_ctx=float(_attr(enc, "dropout_ctx", 0.0)),
                  dropout_emb=float(_attr(enc, "dropout_emb", 0.0)), dropout_out=float(_attr(dec, "dropout_out", 0.0)), tied_emb=tied)
    if kind == "NMT_AttentionImagine_Seq2Seq_Beam_V11":
        vse = stub._modules["vse_imagine"]
        model = NMT_AttentionImagine_Seq2Seq_Beam_V11(
            src_size, tgt_size, sd["vse_imagine.im_embedding.weight"].shape[1], src_emb, tgt_emb, hidden,
            sd["vse_imagine.im_embedding.weight"].shape[0], float(_attr(stub, "loss_w", 0.99)), beam_size=int(_attr(stub, "beam_size", 1)),
            attn_model=str(_attr(vse, "attn_type", _attr(stub, "attn_model", "dot"))), activation_vse=bool(_attr(vse, "activation_vse", True)),
            init_split=float(_attr(stub, "init_split", 0.5)), dropout_rnn_enc=float(_attr(enc, "dropout_rnn", 0.0)), **common)
    elif kind == "NMT_Seq2Seq_Beam_V2":
        model = NMT_Seq2Seq_Beam_V2(src_size, tgt_size, src_emb, tgt_emb, hidden, beam_size=int(_attr(stub, "beam_size", 1)),
                                    dropout_rnn=float(_attr(enc, "dropout_rnn", 0.0)), **common)
    else:
        raise TypeError(f"no drop-in for reference class {stub._ref_class} (SURVEY.md section 2 lists the other variants as out of scope)")
    missing, unexpected = model.load_state_dict(sd, strict=False)
    if missing or unexpected:
        raise RuntimeError(f"reference checkpoint does not match the drop-in's state_dict: missing {missing}, unexpected {unexpected}")
    return model.eval()


def load_reference_module(path, device: Union[str, torch.device, None] = None) -> nn.Module:
    """``torch.load(path)`` of the reference (nmt_multimodal_beam_DE.py:534) → the equivalent drop-in model."""
    model = convert_reference_module(load_reference_stub(path))
    return model.to(device) if device is not None else model
