"""Evaluation loops in the shape of the reference's drivers (evaluate.py:111-143 validate_sintel, :146-182
validate_kitti, :25-55 create_sintel_submission with warm start), runnable on any iterable of samples — the datasets
themselves are out of scope (no data offline), so `bench.py` / the tests feed synthetic pairs.

Differences from the reference, all behaviour-preserving: pairs of equal size are evaluated in batches instead of one at a
time (the model's batch items are independent; frames of a different size — KITTI's vary — start a new batch), and the
warm-start forward interpolation runs on the GPU.  Metrics aggregate exactly as the reference's loops do: Sintel-style
(`valid` absent) pools all pixels (evaluate.py:131-137), KITTI-style averages per-image means (evaluate.py:172-179).
"""
import numpy as np
import torch

from utils.utils import InputPadder, forward_interpolate


@torch.no_grad()
def validate(model, samples, iters=32, mode="sintel", batch_size=8, device="cuda"):
    """samples: iterable of (image1 [3,H,W], image2 [3,H,W], flow_gt [2,H,W], valid [H,W] or None).
    Returns the metrics validate_sintel / validate_kitti print: EPE, 1px/3px/5px, and KITTI F1 when `valid` is given."""
    model.eval()
    epe_all, f1_all, batch = [], [], []
    epe_img = []                                            # KITTI: per-image mean EPE (evaluate.py:172)

    def flush():
        if not batch:
            return
        im1 = torch.stack([b[0] for b in batch]).to(device).float()
        im2 = torch.stack([b[1] for b in batch]).to(device).float()
        padder = InputPadder(im1.shape, mode=mode)
        p1, p2 = padder.pad(im1, im2)
        _, flow_pr = model(p1, p2, iters=iters, test_mode=True)
        flow = padder.unpad(flow_pr).cpu()
        for k, (_, _, gt, valid) in enumerate(batch):
            epe = torch.sum((flow[k] - gt) ** 2, dim=0).sqrt()
            if valid is None:
                epe_all.append(epe.view(-1).numpy())
            else:                                           # evaluate.py:163-171
                mag = torch.sum(gt ** 2, dim=0).sqrt().view(-1)
                val = valid.view(-1) >= 0.5
                e = epe.view(-1)
                out = ((e > 3.0) & ((e / mag) > 0.05)).float()
                epe_img.append(e[val].mean().item())
                epe_all.append(e[val].numpy())
                f1_all.append(out[val].numpy())
        batch.clear()

    for s in samples:
        s = s if len(s) == 4 else (s[0], s[1], s[2], None)
        if batch and batch[0][0].shape != s[0].shape:       # frame sizes differ (KITTI): close the batch
            flush()
        batch.append(s)
        if len(batch) == batch_size:
            flush()
    flush()
    e = np.concatenate(epe_all)
    res = {"epe": float(np.mean(e)), "1px": float(np.mean(e < 1)), "3px": float(np.mean(e < 3)), "5px": float(np.mean(e < 5))}
    if f1_all:
        res["epe"] = float(np.mean(epe_img))                # evaluate.py:178: mean of the per-image means
        res["f1"] = float(100 * np.mean(np.concatenate(f1_all)))   # evaluate.py:175,179: pooled over all valid pixels
    return res


@torch.no_grad()
def run_sequence(model, frames, iters=32, warm_start=False, mode="sintel", device="cuda"):
    """create_sintel_submission's inner loop (evaluate.py:31-44): consecutive frame pairs of one sequence, optionally
    warm-starting each pair from the forward-interpolated low-resolution flow of the previous one.  Returns the list of
    unpadded [2,H,W] flows (CPU)."""
    model.eval()
    flows, flow_prev = [], None
    for f1, f2 in zip(frames[:-1], frames[1:]):
        im1, im2 = f1[None].to(device).float(), f2[None].to(device).float()
        padder = InputPadder(im1.shape, mode=mode)
        p1, p2 = padder.pad(im1, im2)
        flow_low, flow_pr = model(p1, p2, iters=iters, flow_init=flow_prev, test_mode=True)
        flows.append(padder.unpad(flow_pr[0]).cpu())
        if warm_start:
            flow_prev = forward_interpolate(flow_low[0])[None]
    return flows


def load_checkpoint(model, state):
    """Checkpoints are saved from an nn.DataParallel wrapper (train.py:231): strip the `module.` prefix when present, then
    load strictly (evaluate.py:252-257 loads into the wrapper instead)."""
    if isinstance(state, str):
        state = torch.load(state, map_location="cpu")
    state = {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}
    return model.load_state_dict(state)
