"""Round-2 golden fixtures from the UNMODIFIED reference (authoring container only; needs /root/reference):

    python oracle/make_golden_r2.py     ->  tests/golden/r2.npz + r2_meta.json

  * operator seams: the reference's own NConv2d / NConvUNet / Simple / FlowHead / SepConvGRU / BasicMotionEncoder /
    bilinear_sampler outputs on seeded inputs (incl. all-zero confidences)
  * codec bytes: files written by the reference's frame_utils.writeFlow / writeFlowKITTI, PFM files read by its readPFM
  * training: loss and per-parameter gradients of the reference model (train mode + freeze_bn, as train.py:185-186,215)
    under sequence_loss (train.py:46-71 restated: train.py itself cannot be imported here, SURVEY.md finding 3); the
    gradients are pinned as seeded random projections + norms per parameter, and the differentiable oracle
    (oracle/raft_oracle.py:raft_forward_graph) is asserted against them here.
TEST INFRASTRUCTURE ONLY.
"""
import importlib.util
import json
import os
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from make_golden import OUT, REF, frames, ref_args, tensor_sha   # noqa: E402

GRAD_SHAPE = (128, 160)       # smallest valid size (every pyramid level >= 2 px)
GRAD_ITERS = 3
N_PROJ = 4


def proj_vectors(name, numel):
    g = torch.Generator().manual_seed(abs(hash_name(name)) % (2 ** 31))
    return torch.randn(N_PROJ, numel, generator=g)


def hash_name(name):
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 2147483647
    return h


def grad_fixture(grads):
    """{param name: tensor} -> {name: [norm, proj_0..proj_3]} (float64)."""
    out = {}
    for k, g in grads.items():
        v = g.detach().double().reshape(-1)
        out[k] = [float(v.norm())] + [float(x) for x in (proj_vectors(k, v.numel()).double() @ v)]
    return out


def tied_leaves(model):
    """state_dict copy whose aliased entries (shared modules: norm3 == downsample.1, the NConv encoder aliases) share ONE
    leaf tensor, plus {first parameter name: leaf} in named_parameters() order."""
    by_ptr, sd = {}, {}
    for k, v in model.state_dict().items():
        key = (v.data_ptr(), tuple(v.shape))
        if key not in by_ptr:
            by_ptr[key] = v.detach().clone()
        sd[k] = by_ptr[key]
    leaves = {k: sd[k].requires_grad_(True) for k, _ in model.named_parameters()}
    return sd, leaves


def train_inputs(b=2):
    h, w = GRAD_SHAPE
    im1, im2 = frames(b, h, w, seed=31)
    g = torch.Generator().manual_seed(32)
    gt = torch.randn(b, 2, h, w, generator=g) * 5           # train.py-style synthetic ground truth
    valid = (torch.rand(b, h, w, generator=g) > 0.1).float()
    return im1, im2, gt, valid


def main():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, os.path.join(REF, "core"))
    sys.path.insert(0, ROOT)
    import raft as ref_raft
    import raft_nc_dbl as ref_nc
    from nconv_modules import NConv2d, NConvUNet
    from utils.utils import bilinear_sampler
    from oracle import raft_oracle as orc
    spec = importlib.util.spec_from_file_location("ref_frame_utils", os.path.join(REF, "core", "utils", "frame_utils.py"))
    rfu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rfu)

    torch.set_num_threads(os.cpu_count())
    gold, meta = {}, {"reference_commit": "51ac387", "torch": torch.__version__}

    # ------------------------------------------------------------------ NConv2d / NConvUNet seams
    torch.manual_seed(77)
    layer = NConv2d(2, 2, (5, 5))
    g = torch.Generator().manual_seed(78)
    data = torch.randn(3, 2, 19, 23, generator=g) * 4
    conf = torch.rand(3, 2, 19, 23, generator=g)
    conf[conf < 0.3] = 0.0
    conf[2] = 0.0                                             # a sample whose confidences are all zero (eps = 1e-20 path)
    with torch.no_grad():
        y, c = layer((data, conf))
    gold.update(nconv_weight_p=layer.weight_p.detach().numpy(), nconv_data=data.numpy(), nconv_conf=conf.numpy(),
                nconv_y=y.numpy(), nconv_cout=c.numpy())
    torch.manual_seed(79)
    unet = NConvUNet(in_ch=1, channels_multiplier=2, num_downsampling=1, encoder_filter_sz=5, decoder_filter_sz=3,
                     out_filter_sz=1, use_bias=False, data_pooling="conf_based", shared_encoder=True, use_double_conv=False)
    d1 = torch.randn(2, 1, 24, 40, generator=g) * 3
    c1 = (torch.rand(2, 1, 24, 40, generator=g) > 0.8).float() * torch.rand(2, 1, 24, 40, generator=g)
    with torch.no_grad():
        xo, co = unet((d1, c1))
    for k, v in unet.state_dict().items():
        gold["unet_sd_" + k] = v.numpy()
    gold.update(unet_data=d1.numpy(), unet_conf=c1.numpy(), unet_xout=xo.numpy(), unet_cout=co.numpy())

    # ------------------------------------------------------------------ update-block sub-module seams + Simple (seed-1234 model)
    torch.manual_seed(1234)
    m = ref_nc.RAFT(ref_args("sintel")).eval()
    z = np.load(os.path.join(OUT, "cfg1.npz"))
    net_in, inp, corr, coords = (torch.from_numpy(z[k]) for k in ("net_in_it3", "inp", "corr_it3", "coords_it3"))
    flow = coords - orc.coords_grid(1, 16, 32)
    with torch.no_grad():
        mot = m.update_block.encoder(flow, corr)
        h = m.update_block.gru(net_in, torch.cat([inp, mot], 1))
        df = m.update_block.flow_head(h)
        x130 = torch.randn(1, 130, 32, 64, generator=g)
        sw = m.upsampler.weights_est_net(x130)
    gold.update(seam_motion=mot.numpy(), seam_gru=h.numpy(), seam_flow_head=df.numpy(), seam_simple_in=x130.numpy(),
                seam_simple_out=sw.numpy())

    # ------------------------------------------------------------------ bilinear_sampler
    img = torch.randn(2, 3, 9, 13, generator=g)
    co = torch.rand(2, 5, 7, 2, generator=g) * torch.tensor([16.0, 12.0]) - 2.0       # some samples outside
    with torch.no_grad():
        bs, bm = bilinear_sampler(img, co, mask=True)
    gold.update(bs_img=img.numpy(), bs_coords=co.numpy(), bs_out=bs.numpy(), bs_mask=bm.numpy())

    # ------------------------------------------------------------------ codec bytes written / read by the reference
    rng = np.random.default_rng(5)
    fl = (rng.normal(size=(7, 11, 2)) * 20).astype(np.float32)
    flk = (rng.integers(-3000, 3000, size=(6, 10, 2)) / 64.0 + rng.random((6, 10, 2)) * 0.01).astype(np.float32)
    with tempfile.TemporaryDirectory() as td:
        rfu.writeFlow(os.path.join(td, "a.flo"), fl)
        gold["flo_bytes"] = np.frombuffer(open(os.path.join(td, "a.flo"), "rb").read(), np.uint8)
        rfu.writeFlow(os.path.join(td, "b.flo"), fl[..., 0], fl[..., 1])
        gold["flo_bytes_uv"] = np.frombuffer(open(os.path.join(td, "b.flo"), "rb").read(), np.uint8)
        rfu.writeFlowKITTI(os.path.join(td, "k.png"), flk)
        gold["kitti_png_bytes"] = np.frombuffer(open(os.path.join(td, "k.png"), "rb").read(), np.uint8)
        kf, kv = rfu.readFlowKITTI(os.path.join(td, "k.png"))
        gold.update(kitti_read_flow=kf, kitti_read_valid=kv)
        for tag, header, arr, dt in (("pfm_le_color", b"PF\n5 4\n-1.0\n", rng.normal(size=(4, 5, 3)), "<f4"),
                                     ("pfm_be_grey", b"Pf\n6 3\n1.0\n", rng.normal(size=(3, 6)), ">f4")):
            raw = header + arr.astype(dt).tobytes()
            open(os.path.join(td, tag + ".pfm"), "wb").write(raw)
            gold[tag + "_bytes"] = np.frombuffer(raw, np.uint8)
            gold[tag + "_read"] = np.ascontiguousarray(rfu.readPFM(os.path.join(td, tag + ".pfm"))).astype(np.float32)
    gold.update(flo_flow=fl, kitti_flow=flk)

    # ------------------------------------------------------------------ training gradients (reference autograd, CPU)
    im1, im2, gt, valid = train_inputs()
    for name, mod in (("raft_nc_dbl", ref_nc), ("raft", ref_raft)):
        torch.manual_seed(1234)
        model = mod.RAFT(ref_args("sintel"))
        model.train()
        model.freeze_bn()                                      # train.py:185-186 (every stage but chairs)
        preds = model(im1, im2, iters=GRAD_ITERS)
        loss = orc.sequence_loss(preds, gt, valid, gamma=0.85)
        loss.backward()
        grads = {k: p.grad for k, p in model.named_parameters()}
        assert all(v is not None for v in grads.values()), "a parameter received no gradient"
        meta[f"train_loss_{name}"] = float(loss)
        meta[f"train_grads_{name}"] = grad_fixture(grads)
        # differentiable oracle vs reference autograd
        sd, leaves = tied_leaves(model)
        _, _, ups = orc.raft_forward_graph(sd, im1, im2, iters=GRAD_ITERS, model=name)
        oloss = orc.sequence_loss(ups, gt, valid, gamma=0.85)
        oloss.backward()
        worst = 0.0
        gmax = max(float(v.norm()) for v in grads.values())
        for k, p in leaves.items():
            # error relative to the parameter's gradient, with a floor for gradients that are mathematically zero (biases in
            # front of an InstanceNorm) or pure cancellation noise (the scale-invariant 1x1 nconv_out): 1e-5 of the largest
            rel = float((p.grad - grads[k]).norm() / (grads[k].norm() + 1e-5 * gmax))
            if rel > 1e-4:
                print(f"   {k}: rel {rel:.2e} |g| {float(grads[k].norm()):.3e}")
            worst = max(worst, rel)
        meta[f"train_grad_norm_max_{name}"] = gmax
        print(f"[{name}] loss ref {float(loss):.6f} oracle {float(oloss):.6f}; worst per-parameter rel. gradient error {worst:.2e}")
        assert abs(float(loss) - float(oloss)) < 1e-4 and worst < 1e-3
        meta[f"train_oracle_vs_reference_{name}"] = {"loss_abs": abs(float(loss) - float(oloss)), "grad_rel_worst": worst}
    meta["train_inputs_sha"] = [tensor_sha(t) for t in (im1, im2, gt, valid)]
    meta["train_cfg"] = {"shape": list(GRAD_SHAPE), "iters": GRAD_ITERS, "batch": 2, "gamma": 0.85, "n_proj": N_PROJ}

    np.savez_compressed(os.path.join(OUT, "r2.npz"), **gold)
    with open(os.path.join(OUT, "r2_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(f"wrote r2.npz ({os.path.getsize(os.path.join(OUT, 'r2.npz')) / 1e6:.2f} MB) and r2_meta.json")


if __name__ == "__main__":
    main()
