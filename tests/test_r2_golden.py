"""Round-2 fixtures from the unmodified reference (oracle/make_golden_r2.py): the oracle restatements of the operator
seams, byte-exact flow codecs, and the differentiable oracle's gradients against the reference's autograd."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, build_model
from oracle import raft_oracle as orc
from oracle.make_golden_r2 import GRAD_ITERS, grad_fixture, tied_leaves, train_inputs


@pytest.fixture(scope="module")
def g2():
    z = np.load(os.path.join(ROOT, "tests", "golden", "r2.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def meta2():
    with open(os.path.join(ROOT, "tests", "golden", "r2_meta.json")) as f:
        return json.load(f)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_oracle_nconv2d_matches_reference_layer(g2):
    w = orc.softplus10(T(g2["nconv_weight_p"]))
    y, c = orc.nconv2d(T(g2["nconv_data"]), T(g2["nconv_conf"]), w)
    assert (y - T(g2["nconv_y"])).abs().max() < 1e-5 and (c - T(g2["nconv_cout"])).abs().max() < 1e-6
    assert T(g2["nconv_y"])[2].abs().max() == 0          # all-zero confidence -> 0 / (0 + 1e-20) = 0, not NaN


def test_oracle_unet_live_path_matches_reference(g2):
    sd = {"upsampler.interpolation_net." + k[len("unet_sd_"):]: T(v) for k, v in g2.items() if k.startswith("unet_sd_")}
    x, c = orc.nconv_unet_live(sd, T(g2["unet_data"]), T(g2["unet_conf"]))
    assert (x - T(g2["unet_xout"])).abs().max() < 1e-4 and (c - T(g2["unet_cout"])).abs().max() < 1e-6


def test_oracle_submodules_match_reference(g2, gold, sd_ncup):
    flow = gold["coords_it3"] - orc.coords_grid(1, 16, 32)
    mot = orc.motion_encoder(sd_ncup, flow, gold["corr_it3"])
    assert (mot - T(g2["seam_motion"])).abs().max() < 1e-5
    h = orc.sep_conv_gru(sd_ncup, gold["net_in_it3"], torch.cat([gold["inp"], mot], 1))
    assert (h - T(g2["seam_gru"])).abs().max() < 1e-5
    assert (orc.flow_head(sd_ncup, h) - T(g2["seam_flow_head"])).abs().max() < 1e-5
    w = orc.weights_net(sd_ncup, T(g2["seam_simple_in"]), use_bn=True)
    assert (w - T(g2["seam_simple_out"])).abs().max() < 1e-6


def test_oracle_bilinear_sampler_matches_reference(g2):
    img, co = T(g2["bs_img"]), T(g2["bs_coords"])
    n, c, h, w = img.shape
    out = torch.stack([orc._bilinear_zero(img[:, k:k + 1], co[..., 0].reshape(n, -1), co[..., 1].reshape(n, -1)).view(n, 5, 7)
                       for k in range(c)], 1)
    assert (out - T(g2["bs_out"])).abs().max() < 1e-6


# ----------------------------------------------------------------------------- codecs: bytes written by the reference


def test_flo_writer_is_byte_identical_to_reference(g2, tmp_path):
    from utils.frame_utils import readFlow, writeFlow
    fn = str(tmp_path / "a.flo")
    writeFlow(fn, g2["flo_flow"])
    assert open(fn, "rb").read() == g2["flo_bytes"].tobytes()
    writeFlow(fn, g2["flo_flow"][..., 0], g2["flo_flow"][..., 1])
    assert open(fn, "rb").read() == g2["flo_bytes_uv"].tobytes()
    assert np.array_equal(readFlow(fn), g2["flo_flow"])


def test_kitti_png_writer_matches_reference(g2, tmp_path):
    cv2 = pytest.importorskip("cv2")
    from utils.frame_utils import readFlowKITTI, writeFlowKITTI
    fn = str(tmp_path / "k.png")
    writeFlowKITTI(fn, g2["kitti_flow"])
    ours = cv2.imread(fn, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)
    open(fn, "wb").write(g2["kitti_png_bytes"].tobytes())
    theirs = cv2.imread(fn, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)
    assert ours.dtype == np.uint16 and np.array_equal(ours, theirs)              # same pixels (incl. the truncation to uint16)
    flow, valid = readFlowKITTI(fn)                                               # our reader on the reference's file
    assert np.array_equal(flow, g2["kitti_read_flow"]) and np.array_equal(valid, g2["kitti_read_valid"])
    writeFlowKITTI(fn, g2["kitti_flow"])
    assert open(fn, "rb").read() == g2["kitti_png_bytes"].tobytes()               # same encoder, same bytes


@pytest.mark.parametrize("tag", ["pfm_le_color", "pfm_be_grey"])
def test_pfm_reader_matches_reference(g2, tmp_path, tag):
    from utils.frame_utils import read_gen, readPFM
    fn = str(tmp_path / (tag + ".pfm"))
    open(fn, "wb").write(g2[tag + "_bytes"].tobytes())
    out = readPFM(fn)
    assert out.shape == g2[tag + "_read"].shape and np.array_equal(out.astype(np.float32), g2[tag + "_read"])
    gen = read_gen(fn)
    assert gen.shape == (g2[tag + "_read"].shape[:2] + (2,) if out.ndim == 3 else out.shape)
    open(fn, "wb").write(b"P6\n1 1\n")
    with pytest.raises(Exception):
        readPFM(fn)


# ----------------------------------------------------------------------------- training oracle vs the reference's autograd


def test_oracle_gradients_match_reference_autograd(meta2):
    """Loss and per-parameter gradients of the differentiable oracle (train mode + frozen BN, 128x160, 3 iterations) against
    the reference's own backward pass, pinned as norms + seeded projections per parameter."""
    name = "raft_nc_dbl"
    m = build_model(name)
    im1, im2, gt, valid = train_inputs()
    sd, leaves = tied_leaves(m)
    _, _, ups = orc.raft_forward_graph(sd, im1, im2, iters=GRAD_ITERS, model=name)
    loss = orc.sequence_loss(ups, gt, valid, gamma=0.85)
    assert abs(float(loss) - meta2[f"train_loss_{name}"]) < 1e-4
    loss.backward()
    fix = grad_fixture({k: p.grad for k, p in leaves.items()})
    ref = meta2[f"train_grads_{name}"]
    gmax = meta2[f"train_grad_norm_max_{name}"]
    assert set(fix) == set(ref)
    for k in ref:
        n = np.sqrt(leaves[k].numel())
        tol = 1e-3 * ref[k][0] + 1e-5 * gmax
        assert abs(fix[k][0] - ref[k][0]) < tol, k
        assert all(abs(a - b) < tol * n for a, b in zip(fix[k][1:], ref[k][1:])), k
