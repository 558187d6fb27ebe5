"""Checkpoint round trip in the reference's format.

The reference writes ``<name><epoch>.pth.tar`` = ``torch.save({'epoch', 'autoencoder_state_dict',
'optimizer_state_dict', 'scheduler_state_dict'})`` (train_funcs.py:451-455, 563-567) and resumes / fine-tunes from it in
main.py:277-290.  The drop-in models keep the reference's ``state_dict`` keys and parameter registration order, so these
files load into them unchanged (optimizer state is positional: it relies on that order); the helpers below only package
the two code paths.
"""
import torch

KEYS = ("epoch", "autoencoder_state_dict", "optimizer_state_dict", "scheduler_state_dict")


def save_checkpoint(path, model, optimizer, scheduler, epoch):
    """train_funcs.py:451-455.  `scheduler` may be None (cfg.TRAIN.scheduler[0] false): an empty dict is stored."""
    torch.save({"epoch": int(epoch),
                "autoencoder_state_dict": model.state_dict(),
                "optimizer_state_dict": optimizer.state_dict(),
                "scheduler_state_dict": scheduler.state_dict() if scheduler is not None else {}}, path)


def load_checkpoint(path, model, optimizer=None, scheduler=None, finetune=False, map_location=None):
    """main.py:277-290.  Returns the epoch to start from: 1 when fine-tuning (weights only, cfg.TRAIN.resume[2]), else
    stored epoch + 1 with optimizer (and scheduler, when both sides have one) state restored."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    missing = [k for k in KEYS[:2] if k not in ckpt]
    if missing:
        raise KeyError(f"{path} is not a reference checkpoint (missing {missing})")
    model.load_state_dict(ckpt["autoencoder_state_dict"])
    if finetune:
        return 1
    if optimizer is not None:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    if scheduler is not None and ckpt.get("scheduler_state_dict"):
        scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    return int(ckpt["epoch"]) + 1
