"""Operator seams of SURVEY.md §8b on the GPU, each through the C ABI, against outputs of the reference's own modules
(tests/golden/r2.npz, oracle/make_golden_r2.py); plus the device / nn.DataParallel contract of the boundary."""
import copy
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import ROOT, build_model, frames
from oracle import raft_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def g2():
    z = np.load(os.path.join(ROOT, "tests", "golden", "r2.npz"))
    return {k: torch.from_numpy(np.ascontiguousarray(z[k])) for k in z.files}


def epe(a, b):
    return (a - b).pow(2).sum(1).sqrt().mean().item()


def test_nconv2d_layer_matches_reference(g2):
    from nconv_modules import NConv2d
    layer = NConv2d(2, 2, (5, 5))
    with torch.no_grad():
        layer.weight_p.copy_(g2["nconv_weight_p"])
    layer = layer.to(DEV)
    with torch.no_grad():
        y, c = layer((g2["nconv_data"].to(DEV), g2["nconv_conf"].to(DEV)))
    assert (y.cpu() - g2["nconv_y"]).abs().max() < 1e-5 and (c.cpu() - g2["nconv_cout"]).abs().max() < 1e-6
    assert y[2].abs().max() == 0 and torch.isfinite(y).all()                     # all-zero confidences: 0, never NaN


def test_nconv_unet_matches_reference(g2):
    from nconv_modules import NConvUNet
    net = NConvUNet(in_ch=1, channels_multiplier=2, num_downsampling=1, encoder_filter_sz=5, decoder_filter_sz=3,
                    out_filter_sz=1, use_bias=False, data_pooling="conf_based", shared_encoder=True, use_double_conv=False)
    net.load_state_dict({k[len("unet_sd_"):]: v for k, v in g2.items() if k.startswith("unet_sd_")})
    net = net.to(DEV)
    with torch.no_grad():
        x, c = net((g2["unet_data"].to(DEV), g2["unet_conf"].to(DEV)))
    assert (x.cpu() - g2["unet_xout"]).abs().max() < 1e-4 and (c.cpu() - g2["unet_cout"]).abs().max() < 1e-6


def test_update_block_submodules_match_reference(g2, gold):
    m = build_model("raft_nc_dbl").to(DEV)
    ub = m.update_block
    flow = (gold["coords_it3"] - orc.coords_grid(1, 16, 32)).to(DEV)
    with torch.no_grad():
        mot = ub.encoder(flow, gold["corr_it3"].to(DEV))
        h = ub.gru(gold["net_in_it3"].to(DEV), torch.cat([gold["inp"].to(DEV), mot], 1))
        df = ub.flow_head(h)
        w = m.upsampler.weights_est_net(g2["seam_simple_in"].to(DEV))
    assert (mot.cpu() - g2["seam_motion"]).abs().max() < 2e-5
    assert (h.cpu() - g2["seam_gru"]).abs().max() < 2e-5
    assert (df.cpu() - g2["seam_flow_head"]).abs().max() < 2e-5
    assert (w.cpu() - g2["seam_simple_out"]).abs().max() < 2e-6


def test_bilinear_sampler_matches_reference(g2):
    from utils.utils import bilinear_sampler
    out, mask = bilinear_sampler(g2["bs_img"].to(DEV), g2["bs_coords"].to(DEV), mask=True)
    assert (out.cpu() - g2["bs_out"]).abs().max() < 1e-6 and torch.equal(mask.cpu(), g2["bs_mask"])
    assert torch.equal(bilinear_sampler(g2["bs_img"].to(DEV), g2["bs_coords"].to(DEV)), out)


# ----------------------------------------------------------------------------- boundary contract: devices, replicas, copies


def test_model_is_deepcopyable_and_picklable_after_a_forward():
    m = build_model("raft_nc_dbl").to(DEV)
    im1, im2 = frames(1, 128, 256)
    with torch.no_grad():
        _, up = m(im1.to(DEV), im2.to(DEV), iters=2, test_mode=True)
        m2 = copy.deepcopy(m)
        m3 = pickle.loads(pickle.dumps(m))
        assert torch.equal(m2(im1.to(DEV), im2.to(DEV), iters=2, test_mode=True)[1], up)
        assert torch.equal(m3(im1.to(DEV), im2.to(DEV), iters=2, test_mode=True)[1], up)


def test_data_mutation_needs_invalidate_or_checksum(monkeypatch):
    from rnc.engine import invalidate_packed
    m = build_model("raft_nc_dbl").to(DEV)
    im1, im2 = frames(1, 128, 256)
    a, b = im1.to(DEV), im2.to(DEV)
    with torch.no_grad():
        up0 = m(a, b, iters=2, test_mode=True)[1]
        m.update_block.flow_head.conv2.weight.data.mul_(0.5)                     # bypasses the version counter
        invalidate_packed()
        up1 = m(a, b, iters=2, test_mode=True)[1]
        assert not torch.equal(up0, up1)
        monkeypatch.setenv("RNC_PARAM_CHECK", "checksum")                         # content-keyed: no invalidate needed
        up1b = m(a, b, iters=2, test_mode=True)[1]
        m.update_block.flow_head.conv2.weight.data.mul_(2.0)
        up2 = m(a, b, iters=2, test_mode=True)[1]
    assert torch.equal(up1, up1b) and epe(up2, up0) < 1e-4


def test_mismatched_devices_are_rejected():
    m = build_model("raft_nc_dbl").to(DEV)
    im = torch.zeros(1, 3, 128, 256)
    with torch.no_grad(), pytest.raises(Exception):
        m(im.to(DEV), im, iters=1, test_mode=True)                               # CPU tensor


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_non_current_device_and_dataparallel_match_single_gpu():
    """evaluate.py:246-252 / train.py:175 wrap the model in nn.DataParallel: replicas run concurrently from one thread per
    device; a model on cuda:1 must also work while cuda:0 is the current device."""
    m = build_model("raft_nc_dbl").to("cuda:0")
    im1, im2 = frames(4, 128, 256, seed=9)
    with torch.no_grad():
        _, ref = m(im1.to("cuda:0"), im2.to("cuda:0"), iters=3, test_mode=True)
        m1 = copy.deepcopy(m).to("cuda:1")
        assert torch.cuda.current_device() == 0
        lo1, up1 = m1(im1.to("cuda:1"), im2.to("cuda:1"), iters=3, test_mode=True)
        assert up1.device.index == 1 and epe(up1.cpu(), ref.cpu()) < 1e-4
        with pytest.raises(ValueError):
            m1(im1.to("cuda:0"), im2.to("cuda:0"), iters=1, test_mode=True)       # parameters on cuda:1, images on cuda:0
        dp = torch.nn.DataParallel(m, device_ids=[0, 1])
        for _ in range(2):                                                         # second pass: replicas get fresh copies
            lo, up = dp(im1.to("cuda:0"), im2.to("cuda:0"), iters=3, test_mode=True)
            assert up.shape == ref.shape and epe(up.cpu(), ref.cpu()) < 1e-4
        # weights updated between DataParallel forwards (as an optimizer step would): replicas must not serve stale packs
        m.update_block.flow_head.conv2.weight.mul_(0.5)
        _, ref2 = m(im1.to("cuda:0"), im2.to("cuda:0"), iters=3, test_mode=True)
        _, up2 = dp(im1.to("cuda:0"), im2.to("cuda:0"), iters=3, test_mode=True)
        assert epe(up2.cpu(), ref2.cpu()) < 1e-4 and epe(ref2.cpu(), ref.cpu()) > 1e-3
