"""The optimiser shell around the step driver: learning-rate decay on a validation plateau,
early stopping on BLEU, best-checkpoint bookkeeping — nmt_multimodal_beam_DE.py:332-335, 469, 491-527.

``ReduceLROnPlateau`` follows ``torch.optim.lr_scheduler.ReduceLROnPlateau`` in its default 'min' / relative-threshold mode
(the reference constructs it with factor 0.2, patience 10 and steps it with the mean dev translation loss) but drives
``ClipAdam.param_groups`` — the fused optimiser is not a ``torch.optim.Optimizer``.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch


class ReduceLROnPlateau:
    def __init__(self, optimizer, factor: float = 0.1, patience: int = 10, threshold: float = 1e-4, cooldown: int = 0,
                 min_lr: float = 0.0, eps: float = 1e-8):
        if factor >= 1.0:
            raise ValueError("Factor should be < 1.0.")
        self.optimizer = optimizer
        self.factor, self.patience, self.threshold, self.cooldown, self.min_lr, self.eps = factor, patience, threshold, cooldown, min_lr, eps
        self.best = math.inf
        self.num_bad_epochs = 0
        self.cooldown_counter = 0
        self.last_epoch = 0

    def _better(self, value: float) -> bool:
        return value < self.best * (1.0 - self.threshold)

    def step(self, metric) -> None:
        value = float(metric)
        self.last_epoch += 1
        if self._better(value):
            self.best = value
            self.num_bad_epochs = 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            for group in self.optimizer.param_groups:
                old = float(group["lr"])
                new = max(old * self.factor, self.min_lr)
                if old - new > self.eps:
                    group["lr"] = new
            self.cooldown_counter = self.cooldown
            self.num_bad_epochs = 0

    def state_dict(self) -> Dict:
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, state: Dict) -> None:
        self.__dict__.update(state)


class EarlyStopping:
    """The BLEU-driven counter of nmt_multimodal_beam_DE.py:389,496-502,521-527: reset to ``patience`` when the dev BLEU
    improves, decremented otherwise; training stops when it reaches 0."""

    def __init__(self, patience: int):
        self.patience = int(patience)
        self.counter = int(patience)
        self.best = 0.0

    def update(self, bleu: float) -> bool:
        """→ True when this evaluation set a new best (the caller saves the 'best_BLEU' checkpoint)."""
        if bleu > self.best:
            self.best = float(bleu)
            self.counter = self.patience
            return True
        self.counter -= 1
        return False

    @property
    def should_stop(self) -> bool:
        return self.counter == 0


def save_checkpoint(path: str, model: torch.nn.Module, optimizer=None, scheduler: Optional[ReduceLROnPlateau] = None, extra: Optional[Dict] = None) -> None:
    """state_dict-based checkpoint (the reference pickles the whole module, nmt_multimodal_beam_DE.py:491-519; the keys of
    ``model`` are the reference's, so the weights interchange through ``load_state_dict`` in both directions)."""
    blob = {"model": {k: v.detach().cpu() for k, v in model.state_dict().items()}, "extra": extra or {}}
    if optimizer is not None:
        names = {id(p): n for n, p in model.named_parameters()}
        blob["optimizer"] = {"step_count": optimizer.step_count,
                             "lr": [float(g["lr"]) for g in optimizer.param_groups],
                             "state": {names[id(p)]: (m.detach().cpu(), v.detach().cpu()) for p, (m, v) in optimizer.state.items() if id(p) in names}}
    if scheduler is not None:
        blob["scheduler"] = scheduler.state_dict()
    tmp = path + ".tmp"
    torch.save(blob, tmp)
    os.replace(tmp, path)


def load_checkpoint(path: str, model: torch.nn.Module, optimizer=None, scheduler: Optional[ReduceLROnPlateau] = None) -> Dict:
    blob = torch.load(path, map_location="cpu", weights_only=True)   # tensors, dicts, floats, ints only
    model.load_state_dict(blob["model"])
    if optimizer is not None and "optimizer" in blob:
        o = blob["optimizer"]
        optimizer.step_count = int(o["step_count"])
        for g, lr in zip(optimizer.param_groups, o["lr"]):
            g["lr"] = float(lr)
        params = dict(model.named_parameters())
        optimizer.state = {params[n]: (m.to(params[n].device), v.to(params[n].device)) for n, (m, v) in o["state"].items() if n in params}
        optimizer._table_key = None
    if scheduler is not None and "scheduler" in blob:
        scheduler.load_state_dict(blob["scheduler"])
    return blob.get("extra", {})
