import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "raft-ncup_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def gold():
    z = np.load(os.path.join(ROOT, "tests", "golden", "cfg1.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def meta():
    with open(os.path.join(ROOT, "tests", "golden", "meta.json")) as f:
        return json.load(f)


def ref_args(dataset="sintel"):
    """The flag values every reference script ships (eval_raft_nc_sintel.sh:12-34)."""
    import argparse
    return argparse.Namespace(
        small=False, mixed_precision=False, load_pretrained=None, freeze_raft=False, dataset=dataset, align_corners=True,
        final_upsampling="NConvUpsampler", final_upsampling_scale=4, final_upsampling_use_data_for_guidance=True,
        final_upsampling_channels_to_batch=True, final_upsampling_use_residuals=False, final_upsampling_est_on_high_res=False,
        interp_net="NConvUNet", interp_net_channels_multiplier=2, interp_net_num_downsampling=1,
        interp_net_data_pooling="conf_based", interp_net_encoder_filter_sz=5, interp_net_decoder_filter_sz=3,
        interp_net_out_filter_sz=1, interp_net_shared_encoder=True, interp_net_use_double_conv=False, interp_net_use_bias=False,
        weights_est_net="Simple", weights_est_net_num_ch=[64, 32], weights_est_net_filter_sz=[3, 3, 1],
        weights_est_net_dilation=[1, 1, 1])


def build_model(name="raft_nc_dbl", dataset="sintel", seed=1234):
    """Seeded model on CPU in eval mode — bit-identical weights to the reference built with the same seed."""
    import importlib
    torch.manual_seed(seed)
    mod = importlib.import_module(name)
    return mod.RAFT(ref_args(dataset)).eval()


def frames(b, h, w, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, 3, h, w, generator=g) * 255, torch.rand(b, 3, h, w, generator=g) * 255


@pytest.fixture(scope="session")
def sd_ncup():
    return {k: v.detach() for k, v in build_model("raft_nc_dbl").state_dict().items()}


@pytest.fixture(scope="session")
def sd_raft():
    return {k: v.detach() for k, v in build_model("raft").state_dict().items()}
