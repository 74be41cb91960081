"""Pins the CPU oracle (oracle/raft_oracle.py) against outputs of the unmodified reference (tests/golden/, produced by
oracle/make_golden.py in the authoring container).  The reference has no tests or vectors of its own."""
import torch

from conftest import frames
from oracle import raft_oracle as orc

TOL = 1e-4   # fp32 re-ordering noise between the reference's grid_sample/conv kernels and the restatement is ~1e-5


def close(a, b, tol=TOL):
    return (a - b).abs().max().item() <= tol


def test_frames_are_reproducible(meta):
    from oracle.make_golden import tensor_sha
    im1, im2 = frames(1, 128, 256)
    assert [tensor_sha(im1), tensor_sha(im2)] == meta["frames_sha_cfg1"]


def test_end_to_end_ncup(gold, sd_ncup):
    im1, im2 = frames(1, 128, 256)
    lo, up, ups = orc.raft_forward(sd_ncup, im1, im2, iters=4, model="raft_nc_dbl")
    assert close(lo, gold["raft_nc_dbl_flow_low"])
    assert close(up, gold["raft_nc_dbl_flow_up"], 2e-4)
    assert close(ups[0], gold["raft_nc_dbl_pred0"], 2e-4)


def test_end_to_end_convex(gold, sd_raft):
    im1, im2 = frames(1, 128, 256)
    lo, up, ups = orc.raft_forward(sd_raft, im1, im2, iters=4, model="raft")
    assert close(lo, gold["raft_flow_low"])
    assert close(up, gold["raft_flow_up"], 2e-4)
    assert close(ups[0], gold["raft_pred0"], 2e-4)


def test_upsample_once_equals_every_iteration(gold, sd_ncup):
    # the per-iteration NCUP calls of test_mode never feed back (SURVEY.md finding 9)
    im1, im2 = frames(1, 128, 256)
    lo, up, _ = orc.raft_forward(sd_ncup, im1, im2, iters=4, model="raft_nc_dbl", upsample_every_iter=False)
    assert close(up, gold["raft_nc_dbl_flow_up"], 2e-4)


def test_kitti_config_no_bn(gold):
    from conftest import build_model
    sd = {k: v.detach() for k, v in build_model("raft_nc_dbl", "kitti").state_dict().items()}
    im1, im2 = frames(1, 128, 256)
    _, up, _ = orc.raft_forward(sd, im1, im2, iters=4, model="raft_nc_dbl", use_bn=False)
    assert close(up, gold["raft_nc_dbl_kitti_flow_up"], 2e-4)


def test_warm_start(gold, sd_ncup):
    im1, im2 = frames(1, 128, 256)
    lo, up, _ = orc.raft_forward(sd_ncup, im1, im2, iters=2, model="raft_nc_dbl", flow_init=gold["warm_flow_init"])
    assert close(lo, gold["warm_flow_low"])
    assert close(up, gold["warm_flow_up"], 2e-4)


def test_corr_lookup_volume_path(gold):
    pyr = orc.corr_pyramid(gold["fmap1"], gold["fmap2"])
    for it in (0, 3):
        assert close(orc.corr_lookup(pyr, gold[f"coords_it{it}"]), gold[f"corr_it{it}"])


def test_corr_lookup_direct_equals_reference(gold):
    # Appendix A.1: lookups against the pooled FEATURE pyramid equal lookups into the pooled 4-D volume
    for it in (0, 3):
        out = orc.corr_lookup_direct(gold["fmap1"], gold["fmap2"], gold[f"coords_it{it}"])
        assert close(out, gold[f"corr_it{it}"])


def test_corr_lookup_odd_sizes(gold):
    out = orc.corr_lookup_direct(gold["odd_f1"], gold["odd_f2"], gold["odd_coords"])
    assert close(out, gold["odd_corr"])
    pyr = orc.corr_pyramid(gold["odd_f1"], gold["odd_f2"])
    assert close(orc.corr_lookup(pyr, gold["odd_coords"]), gold["odd_corr"])


def test_update_block_teacher_forced(gold, sd_ncup):
    for it in (0, 3):
        flow = gold[f"coords_it{it}"] - orc.coords_grid(1, 16, 32)
        net, mask, delta = orc.update_block(sd_ncup, gold[f"net_in_it{it}"], gold["inp"], gold[f"corr_it{it}"], flow, False)
        assert mask is None
        assert close(net, gold[f"net_out_it{it}"], 2e-5)
        assert close(delta, gold[f"delta_it{it}"], 2e-5)


def test_ncup_teacher_forced(gold, sd_ncup):
    for it in (0, 3):
        out = orc.ncup_upsample(sd_ncup, gold[f"ncup_in_flow_it{it}"], gold[f"net_out_it{it}"], use_bn=True)
        assert close(out, gold[f"ncup_out_it{it}"], 2e-5)


def test_convex_upsampler(gold):
    assert close(orc.convex_upsample(gold["convex_flow"], gold["convex_mask"]), gold["convex_out"], 1e-5)


def test_input_pad(meta):
    assert orc.input_pad((1, 3, 436, 1024), "sintel") == meta["sintel_pad"] == [0, 0, 2, 2]
    assert orc.input_pad((1, 3, 375, 1242), "kitti") == meta["kitti_pad"]
    x = torch.arange(2 * 3 * 5 * 7, dtype=torch.float32).view(2, 3, 5, 7)
    pad = orc.input_pad(x.shape, "sintel")
    (y,) = orc.pad_images(pad, x)
    assert y.shape[-2] % 8 == 0 and y.shape[-1] % 8 == 0
    assert torch.equal(orc.unpad(pad, y), x)


def test_sequence_loss_matches_definition():
    g = torch.Generator().manual_seed(0)
    preds = [torch.randn(2, 2, 8, 8, generator=g) for _ in range(3)]
    gt = torch.randn(2, 2, 8, 8, generator=g) * 5
    valid = (torch.rand(2, 8, 8, generator=g) > 0.3).float()
    loss = orc.sequence_loss(preds, gt, valid, gamma=0.85)
    ref = sum(0.85 ** (2 - i) * (valid[:, None] * (p - gt).abs()).mean() for i, p in enumerate(preds))
    assert abs(loss.item() - ref.item()) < 1e-6
