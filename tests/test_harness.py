"""The "next" rows of SURVEY.md §8f-2/f-4: evaluation loops, warm start, on-disk flow formats, checkpoint prefixes."""
import os

import numpy as np
import pytest
import torch

from conftest import build_model, frames


# ----------------------------------------------------------------------------- host-only


def test_flo_roundtrip_and_header(tmp_path):
    from utils.frame_utils import readFlow, writeFlow
    g = np.random.default_rng(0)
    flow = g.normal(size=(7, 11, 2)).astype(np.float32) * 20
    fn = str(tmp_path / "a.flo")
    writeFlow(fn, flow)
    raw = open(fn, "rb").read()
    assert raw[:4] == np.float32(202021.25).tobytes() and raw[4:12] == np.int32([11, 7]).tobytes()   # frame_utils.py:7,86-88
    assert len(raw) == 12 + 7 * 11 * 2 * 4
    assert np.array_equal(readFlow(fn), flow)
    writeFlow(fn, flow[..., 0], flow[..., 1])                                # (u, v) calling form, frame_utils.py:70
    assert np.array_equal(readFlow(fn), flow)
    open(fn, "wb").write(b"\x00" * 32)
    assert readFlow(fn) is None


def test_kitti_png_roundtrip(tmp_path):
    pytest.importorskip("cv2")
    from utils.frame_utils import readFlowKITTI, writeFlowKITTI
    g = np.random.default_rng(1)
    flow = (g.integers(-2000, 2000, size=(5, 9, 2)) / 64.0).astype(np.float32)   # exactly representable at 1/64 px
    fn = str(tmp_path / "k.png")
    writeFlowKITTI(fn, flow)
    f2, valid = readFlowKITTI(fn)
    assert np.array_equal(f2, flow) and (valid == 1).all()


def test_validate_metrics_with_a_stub_model():
    """validate() reproduces evaluate.py:131-140 (EPE / 1px / 3px / 5px) and :163-179 (KITTI F1) on a model stub."""
    from rnc.harness import validate

    class Stub(torch.nn.Module):
        def forward(self, im1, im2, iters=12, test_mode=True, flow_init=None):
            flow = (im1[:, :2] - im2[:, :2])                 # "prediction" = channel difference, at padded resolution
            return flow[:, :, ::8, ::8], flow
    g = torch.Generator().manual_seed(0)
    samples = []
    for k in range(5):
        a, b = torch.rand(3, 20, 30, generator=g) * 4, torch.rand(3, 20, 30, generator=g) * 4
        gt = (a[:2] - b[:2]) + torch.randn(2, 20, 30, generator=g)
        samples.append((a, b, gt, (torch.rand(20, 30, generator=g) > 0.3 + 0.1 * k).float()))   # valid counts differ per image
    res = validate(Stub(), [s[:3] for s in samples], iters=1, batch_size=2, device="cpu")
    e = torch.cat([torch.sum((s[0][:2] - s[1][:2] - s[2]) ** 2, 0).sqrt().view(-1) for s in samples])
    assert abs(res["epe"] - e.mean().item()) < 1e-6 and abs(res["3px"] - (e < 3).float().mean().item()) < 1e-6
    resk = validate(Stub(), samples, iters=1, mode="kitti", batch_size=3, device="cpu")
    ek, f1 = [], []
    for a, b, gt, v in samples:
        ee = torch.sum((a[:2] - b[:2] - gt) ** 2, 0).sqrt().view(-1)
        mag = torch.sum(gt ** 2, 0).sqrt().view(-1)
        m = v.view(-1) >= 0.5
        ek.append(ee[m].mean().item())                      # evaluate.py:172: per-image mean ...
        f1.append(((ee > 3) & (ee / mag > 0.05)).float()[m])
    assert abs(resk["epe"] - sum(ek) / len(ek)) < 1e-6      # ... averaged over images (evaluate.py:178)
    assert abs(resk["f1"] - 100 * torch.cat(f1).mean().item()) < 1e-4
    # frames of different sizes (KITTI) are batched by shape instead of failing in torch.stack
    odd = (torch.rand(3, 16, 40, generator=g), torch.rand(3, 16, 40, generator=g), torch.zeros(2, 16, 40), torch.ones(16, 40))
    resm = validate(Stub(), samples[:2] + [odd] + samples[2:], iters=1, mode="kitti", batch_size=4, device="cpu")
    assert np.isfinite(resm["epe"])


def test_load_checkpoint_strips_dataparallel_prefix():
    from rnc.harness import load_checkpoint
    m = build_model("raft_nc_dbl")
    m2 = build_model("raft_nc_dbl", seed=2)
    load_checkpoint(m2, {"module." + k: v.clone() for k, v in m.state_dict().items()})
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    load_checkpoint(m2, dict(m.state_dict()))                                 # un-prefixed works too


def test_oracle_forward_interpolate_basics():
    from oracle import raft_oracle as orc
    z = orc.forward_interpolate(torch.zeros(2, 6, 9))                          # zero flow: border samples (x=0 / y=0) are dropped
    assert z.abs().max() == 0
    f = torch.zeros(2, 8, 12)
    f[0] += 1.5                                                                # uniform shift: every kept sample carries (1.5, 0)
    out = orc.forward_interpolate(f)
    assert torch.allclose(out[0], torch.full((8, 12), 1.5)) and out[1].abs().max() == 0


# ----------------------------------------------------------------------------- GPU


@pytest.mark.gpu
def test_forward_interpolate_matches_scipy_reference():
    from oracle import raft_oracle as orc
    from utils.utils import forward_interpolate
    g = torch.Generator().manual_seed(4)
    for shape, scale in (((2, 55, 128), 6.0), ((2, 17, 21), 3.0), ((2, 9, 9), 40.0)):
        flow = torch.randn(*shape, generator=g) * scale
        ref = orc.forward_interpolate(flow)
        out = forward_interpolate(flow.cuda())
        assert out.is_cuda and out.shape == flow.shape
        assert torch.equal(out.cpu(), ref)                                     # nearest-sample selection is exact
    f2 = torch.randn(2, 2, 30, 40, generator=g) * 5
    b = forward_interpolate(f2.cuda())                                         # batched form
    assert torch.equal(b[1].cpu(), orc.forward_interpolate(f2[1])) and torch.equal(b[0].cpu(), orc.forward_interpolate(f2[0]))


@pytest.mark.gpu
def test_sequence_with_warm_start_matches_oracle():
    """Two consecutive pairs, the second warm-started from the first (evaluate.py:36-40), vs the CPU oracle chain."""
    from oracle import raft_oracle as orc
    from rnc.harness import run_sequence, validate
    m = build_model("raft_nc_dbl").cuda()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    a, b = frames(1, 128, 256, seed=21)
    c, _ = frames(1, 128, 256, seed=22)
    seq = [a[0], b[0], c[0]]
    flows = run_sequence(m, seq, iters=4, warm_start=True)
    lo0, up0, _ = orc.raft_forward(sd, a, b, iters=4, upsample_every_iter=False)
    init = orc.forward_interpolate(lo0[0])[None]
    lo1, up1, _ = orc.raft_forward(sd, b, c, iters=4, flow_init=init, upsample_every_iter=False)
    epe = lambda x, y: (x - y).pow(2).sum(0).sqrt().mean().item()
    assert epe(flows[0], up0[0]) < 1e-3 and epe(flows[1], up1[0]) < 1e-3
    # validate(): batched evaluation with padding (436 -> 440) against a synthetic ground truth
    im1, im2 = frames(3, 436, 1024, seed=5)
    gt = torch.zeros(3, 2, 436, 1024)
    res = validate(m, [(im1[i], im2[i], gt[i]) for i in range(3)], iters=2, batch_size=2)
    assert np.isfinite(res["epe"]) and 0 <= res["1px"] <= 1
