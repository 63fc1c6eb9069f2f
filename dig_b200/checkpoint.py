"""Checkpoint wire format of the reference (utils/utils.py:546-669): `{'model','optimizer','epoch','scaler','args'}`
written with torch.save on rank 0 as `<output_dir>/checkpoint-<epoch>.pth`; state-dict keys are the reference's,
so checkpoints interchange with the fine-tune loader (run_class_finetuning.py:362-442)."""
import glob
import os

import torch

from . import utils


def save_model(args, epoch, model, model_without_ddp, optimizer, loss_scaler, model_ema=None):
    if not utils.is_main_process():
        return
    path = os.path.join(args.output_dir, "checkpoint-%s.pth" % str(epoch))
    to_save = {"model": model_without_ddp.state_dict(), "optimizer": optimizer.state_dict(), "epoch": epoch,
               "scaler": loss_scaler.state_dict() if loss_scaler is not None else None, "args": args}
    os.makedirs(args.output_dir, exist_ok=True)
    torch.save(to_save, path)


def auto_load_model(args, model, model_without_ddp, optimizer, loss_scaler, model_ema=None):
    """Resume from the newest `checkpoint-<int>.pth` in output_dir (U:581-669)."""
    if getattr(args, "auto_resume", True) and not getattr(args, "resume", ""):
        latest = -1
        for ckpt in glob.glob(os.path.join(args.output_dir, "checkpoint-*.pth")):
            t = ckpt.split("-")[-1].split(".")[0]
            if t.isdigit():
                latest = max(int(t), latest)
        if latest >= 0:
            args.resume = os.path.join(args.output_dir, "checkpoint-%d.pth" % latest)
        print("Auto resume checkpoint: %s" % getattr(args, "resume", ""))
    if getattr(args, "resume", ""):
        ckpt = torch.load(args.resume, map_location="cpu", weights_only=False)
        model_without_ddp.load_state_dict(ckpt["model"])
        print("Resume checkpoint %s" % args.resume)
        if "optimizer" in ckpt and "epoch" in ckpt:
            optimizer.load_state_dict(ckpt["optimizer"])
            args.start_epoch = ckpt["epoch"] + 1 if isinstance(ckpt["epoch"], int) else args.start_epoch
            if loss_scaler is not None and ckpt.get("scaler") is not None:
                loss_scaler.load_state_dict(ckpt["scaler"])
            print("With optim & sched!")
