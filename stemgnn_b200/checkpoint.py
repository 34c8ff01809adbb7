"""state_dict checkpoints with resume.

The reference pickles the WHOLE module every epoch (handler.py:16-24) and never saves optimiser /
scheduler / epoch, so training cannot be resumed and loading needs `torch.load(weights_only=False)`.
These helpers add a tensor-only format next to it (the reference-style files keep being written by
`models.handler.save_model`, so `handler.test` and reference tooling are unaffected):

    save_checkpoint(path, model, optimizer, scheduler, epoch)      # plain dict of tensors / numbers
    model, ckpt = load_checkpoint(path, device)                     # torch.load(weights_only=True)
    restore_training(ckpt, optimizer, scheduler) -> next epoch
    convert_module_pickle(src, dst)                                 # reference-style pickle -> this format
"""
import torch

FORMAT = "stemgnn_b200.state/1"


def _ctor_args(model):
    return {"units": int(model.unit), "stack_cnt": int(model.stack_cnt), "time_step": int(model.time_step),
            "multi_layer": int(model.multi_layer), "horizon": int(model.horizon),
            "dropout_rate": float(getattr(model, "dropout_rate", 0.5)), "leaky_rate": float(model.alpha)}


def save_checkpoint(path, model, optimizer=None, scheduler=None, epoch=None, extra=None):
    ckpt = {"format": FORMAT, "ctor": _ctor_args(model),
            "model": {k: v.detach().cpu() for k, v in model.state_dict().items()},
            "optimizer": optimizer.state_dict() if optimizer is not None else None,
            "scheduler": scheduler.state_dict() if scheduler is not None else None,
            "epoch": None if epoch is None else int(epoch), "extra": extra or {}}
    torch.save(ckpt, path)
    return path


def load_checkpoint(path, device="cpu"):
    """Rebuilds `models.base_model.Model` from a checkpoint written by save_checkpoint."""
    from models.base_model import Model
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    if ckpt.get("format") != FORMAT:
        raise RuntimeError(f"{path}: not a {FORMAT} checkpoint")
    c = ckpt["ctor"]
    model = Model(c["units"], c["stack_cnt"], c["time_step"], c["multi_layer"], horizon=c["horizon"],
                  dropout_rate=c["dropout_rate"], leaky_rate=c["leaky_rate"])
    model.load_state_dict(ckpt["model"])
    return model.to(device), ckpt


def restore_training(ckpt, optimizer=None, scheduler=None):
    """Loads optimiser / scheduler state; returns the epoch to continue with."""
    if optimizer is not None and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    if scheduler is not None and ckpt.get("scheduler") is not None:
        scheduler.load_state_dict(ckpt["scheduler"])
    return 0 if ckpt.get("epoch") is None else ckpt["epoch"] + 1


def convert_module_pickle(src, dst):
    """Reference-style whole-module pickle (`<epoch>_stemgnn.pt`, handler.py:24) -> tensor-only checkpoint.
    Needs `models.base_model` importable as the pickled class path (this repo's drop-in or the reference)."""
    with open(src, "rb") as f:
        module = torch.load(f, map_location="cpu", weights_only=False)
    return save_checkpoint(dst, module)
