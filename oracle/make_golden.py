"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference; it does not exist on the GPU box):

    python oracle/make_golden.py

The reference is a pure-Python/PyTorch repo with no tests or vectors of its own, so the
fixtures are outputs of the reference's own modules on seeded inputs:
  * weights:  torch.manual_seed(1234) then RAFT(args) on CPU (train.py:345 uses the same seed)
  * frames:   torch.Generator().manual_seed(7); rand(B,3,H,W)*255
Everything here is TEST INFRASTRUCTURE.  Nothing in the product imports it.
"""
import argparse
import hashlib
import json
import os
import sys
import warnings

import numpy as np
import torch

REF = os.environ.get("RNC_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def ref_args(dataset="sintel"):
    """The flag values every shipped script uses (eval_raft_nc_sintel.sh:12-34)."""
    return argparse.Namespace(
        small=False, mixed_precision=False, load_pretrained=None, freeze_raft=False, dataset=dataset,
        align_corners=True,
        final_upsampling="NConvUpsampler", final_upsampling_scale=4, final_upsampling_use_data_for_guidance=True,
        final_upsampling_channels_to_batch=True, final_upsampling_use_residuals=False,
        final_upsampling_est_on_high_res=False,
        interp_net="NConvUNet", interp_net_channels_multiplier=2, interp_net_num_downsampling=1,
        interp_net_data_pooling="conf_based", interp_net_encoder_filter_sz=5, interp_net_decoder_filter_sz=3,
        interp_net_out_filter_sz=1, interp_net_shared_encoder=True, interp_net_use_double_conv=False,
        interp_net_use_bias=False,
        weights_est_net="Simple", weights_est_net_num_ch=[64, 32], weights_est_net_filter_sz=[3, 3, 1],
        weights_est_net_dilation=[1, 1, 1])


def tensor_sha(t):
    return hashlib.sha256(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


def frames(b, h, w, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, 3, h, w, generator=g) * 255, torch.rand(b, 3, h, w, generator=g) * 255


def main():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, os.path.join(REF, "core"))
    sys.path.insert(0, ROOT)
    import raft as ref_raft                     # reference core/raft.py
    import raft_nc_dbl as ref_nc                # reference core/raft_nc_dbl.py
    from corr import CorrBlock                  # reference core/corr.py
    from utils.utils import InputPadder         # reference core/utils/utils.py
    from oracle import raft_oracle as orc

    os.makedirs(OUT, exist_ok=True)
    meta = {"reference_commit": "51ac387", "torch": torch.__version__, "weights_seed": 1234, "frames_seed": 7}
    torch.set_num_threads(os.cpu_count())

    models = {}
    for name, mod, dataset in (("raft_nc_dbl", ref_nc, "sintel"), ("raft_nc_dbl_kitti", ref_nc, "kitti"),
                               ("raft", ref_raft, "sintel")):
        torch.manual_seed(1234)
        m = mod.RAFT(ref_args(dataset)).eval()
        models[name] = m
        sd = m.state_dict()
        meta[f"state_sha_{name}"] = {k: tensor_sha(v) for k, v in sd.items()}
        meta[f"state_shape_{name}"] = {k: list(v.shape) for k, v in sd.items()}

    # ---------------- cfg 1: 128x256, B=1, 4 iters (BASELINE.json configs[0]) -----------------
    im1, im2 = frames(1, 128, 256)
    meta["frames_sha_cfg1"] = [tensor_sha(im1), tensor_sha(im2)]
    gold = {}
    for name in ("raft_nc_dbl", "raft"):
        m = models[name]
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        # reference outputs, test mode and train-mode list
        with torch.no_grad():
            lo, up = m(im1, im2, iters=4, test_mode=True)
            preds = m(im1, im2, iters=4, test_mode=False)
        gold[f"{name}_flow_low"] = lo.numpy()
        gold[f"{name}_flow_up"] = up.numpy()
        gold[f"{name}_pred0"] = preds[0].numpy()
        # oracle replay with teacher-forcing trace; pin oracle == reference
        tr = {}
        olo, oup, oups = orc.raft_forward(sd, im1, im2, iters=4, model=name, trace=tr)
        d_lo = (olo - lo).abs().max().item()
        d_up = (oup - up).abs().max().item()
        d_p0 = (oups[0] - preds[0]).abs().max().item()
        print(f"[{name}] oracle vs reference: flow_low {d_lo:.3e}  flow_up {d_up:.3e}  pred0 {d_p0:.3e}")
        assert max(d_lo, d_up, d_p0) < 2e-4, "oracle does not reproduce the reference"
        meta[f"oracle_vs_reference_{name}"] = {"flow_low": d_lo, "flow_up": d_up, "pred0": d_p0}
        if name == "raft_nc_dbl":
            gold["fmap1"] = tr["fmap1"].numpy()
            gold["fmap2"] = tr["fmap2"].numpy()
            gold["net0"] = tr["net0"].numpy()
            gold["inp"] = tr["inp"].numpy()
            # reference CorrBlock on the reference's own fmaps, teacher-forced coords of iterations 0 and 3
            with torch.no_grad():
                cb = CorrBlock(tr["fmap1"], tr["fmap2"], radius=4)
                for it in (0, 3):
                    c = tr["iters"][it]["coords"]
                    ref_corr = cb(c)
                    d = (ref_corr - tr["iters"][it]["corr"]).abs().max().item()
                    print(f"  corr lookup it{it}: oracle vs reference CorrBlock {d:.3e}")
                    assert d < 1e-4
                    gold[f"coords_it{it}"] = c.numpy()
                    gold[f"corr_it{it}"] = ref_corr.numpy()
                    # reference update block, teacher forced
                    flow = c - orc.coords_grid(1, 16, 32)
                    net_in = tr["net0"] if it == 0 else tr["iters"][it - 1]["net"]
                    rnet, _, rdelta = m.update_block(net_in, tr["inp"], ref_corr, flow)
                    gold[f"net_in_it{it}"] = net_in.numpy()
                    gold[f"net_out_it{it}"] = rnet.numpy()
                    gold[f"delta_it{it}"] = rdelta.numpy()
                    # reference NCUP on (flow after update, net)
                    fl = c + rdelta - orc.coords_grid(1, 16, 32)
                    gold[f"ncup_in_flow_it{it}"] = fl.numpy()
                    gold[f"ncup_out_it{it}"] = m.upsample_flow(fl, rnet).numpy()
        else:
            # convex upsampler (raft.py:73-84) on a small seeded case
            g = torch.Generator().manual_seed(5)
            cm = torch.randn(2, 576, 6, 10, generator=g) * 3
            cf = torch.randn(2, 2, 6, 10, generator=g) * 4
            with torch.no_grad():
                gold["convex_mask"] = cm.numpy()
                gold["convex_flow"] = cf.numpy()
                gold["convex_out"] = m.upsample_flow(cf, cm).numpy()

    # kitti config (no BatchNorm in the weights net, upsampler.py:42): final outputs only
    m = models["raft_nc_dbl_kitti"]
    with torch.no_grad():
        lo, up = m(im1, im2, iters=4, test_mode=True)
    gold["raft_nc_dbl_kitti_flow_up"] = up.numpy()

    # warm start (flow_init) path, raft_nc_dbl.py:144-145
    m = models["raft_nc_dbl"]
    g = torch.Generator().manual_seed(11)
    finit = torch.randn(1, 2, 16, 32, generator=g) * 2
    with torch.no_grad():
        lo, up = m(im1, im2, iters=2, flow_init=finit, test_mode=True)
    gold["warm_flow_init"] = finit.numpy()
    gold["warm_flow_low"] = lo.numpy()
    gold["warm_flow_up"] = up.numpy()

    # InputPadder behaviour at the Sintel shape (utils.py:7-25)
    pd = InputPadder((1, 3, 436, 1024), "sintel")
    meta["sintel_pad"] = list(pd._pad)
    pk = InputPadder((1, 3, 375, 1242), "kitti")
    meta["kitti_pad"] = list(pk._pad)

    # odd-size correlation lookup (pooling drops odd rows/cols; SURVEY finding 7): reference CorrBlock
    g = torch.Generator().manual_seed(3)
    f1 = torch.randn(1, 256, 17, 21, generator=g) * 1.5
    f2 = torch.randn(1, 256, 17, 21, generator=g) * 1.5
    co = orc.coords_grid(1, 17, 21) + torch.randn(1, 2, 17, 21, generator=g) * 6
    with torch.no_grad():
        gold["odd_corr"] = CorrBlock(f1, f2, radius=4)(co).numpy()
    gold["odd_coords"] = co.numpy()
    gold["odd_f1"] = f1.numpy()
    gold["odd_f2"] = f2.numpy()
    meta["odd_seed"] = 3

    np.savez_compressed(os.path.join(OUT, "cfg1.npz"), **gold)
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    sz = os.path.getsize(os.path.join(OUT, "cfg1.npz")) / 1e6
    print(f"wrote {OUT}/cfg1.npz ({sz:.2f} MB) and meta.json")


if __name__ == "__main__":
    main()
