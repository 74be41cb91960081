"""Time corr_lookup_umma_kernel alone for several compile-time variants (developer tool, not part of the product path).

    python tools/lookup_probe.py build NAME:-DFOO=1,-DBAR NAME2:...     # here (nvcc cross-compiles): variant .so files
    python tools/lookup_probe.py run [B H W]                              # on the GPU box: times every built variant

Variants are librnc builds whose corr_lookup_umma.cu was compiled with extra -D flags; they land under
raft-ncup_b200/build/variants/ (git-ignored, shipped by gpurun).  Inputs: random unit-variance features, a coherent
flow field (constant + small noise) like the benchmark's.
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "raft-ncup_b200")
VAR = os.path.join(PKG, "build", "variants")
sys.path.insert(0, PKG)


def build(specs):
    from rnc import build as B
    B.build()
    os.makedirs(VAR, exist_ok=True)
    nvcc = B.nvcc_path()
    inc = ["-I", os.path.join(ROOT, "include"), "-I", B.CSRC]
    for spec in specs:
        name, _, defs = spec.partition(":")
        defs = [d for d in defs.split(",") if d]
        obj = os.path.join(VAR, f"corr_lookup_umma_{name}.o")
        subprocess.check_call([nvcc, *B.ARCH, *B.NVCC_FLAGS, *inc, *defs, "-c", os.path.join(B.CSRC, "corr_lookup_umma.cu"), "-o", obj])
        objs = [os.path.join(B.OBJ, s[:-3] + ".o") for s in B.sources() if s != "corr_lookup_umma.cu"] + [obj]
        lib = os.path.join(VAR, f"librnc_{name}.so")
        subprocess.check_call([nvcc, *B.ARCH, "-shared", "-o", lib, *objs, "-lcuda"])
        print("built", lib)


def real_inputs(B):
    """Feature maps and the iteration-32 coords of the benchmark forward (bench.py's stimulus): the realistic window spread."""
    import torch
    sys.path.insert(0, ROOT)
    from rnc.synth import build_model, frames
    from utils.utils import InputPadder
    m = build_model("raft_nc_dbl").cuda()
    im1, im2 = frames(B, 436, 1024)
    p1, p2 = InputPadder(im1.shape, "sintel").pad(im1, im2)
    with torch.no_grad():
        m(p1.cuda(), p2.cuda(), iters=32, test_mode=True)
    eng = m.engine()
    ws = next(w for k, w in eng._ws.items() if k[0] == "umma" and w.B == B)
    return ws.f1_cl.clone(), ws.f2_pyr.clone(), ws.coords1.clone()


def run(B=8, H=55, W=128):
    import torch
    torch.manual_seed(0)
    dev = "cuda:0"
    D, L = 256, 4
    if os.environ.get("PROBE_REAL", "1") == "1" and (H, W) == (55, 128):
        f1, f2, coords = real_inputs(B)
    else:
        f1 = torch.randn(B, H, W, D, device=dev)
        lv = [torch.randn(B, H >> l, W >> l, D, device=dev) for l in range(L)]
        f2 = torch.cat([t.reshape(-1) for t in lv])
        ys, xs = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
        coords = torch.stack([xs + 3.3, ys - 2.1], 0).float()[None].repeat(B, 1, 1, 1) + 0.3 * torch.randn(B, 2, H, W, device=dev)
        coords = coords.contiguous()
    f1h, f2h = f1.half().contiguous(), f2.half().contiguous()
    hi = torch.zeros(B * H * W, 352, dtype=torch.float16, device=dev)
    lo = torch.zeros_like(hi)
    libs = sorted(f for f in os.listdir(VAR) if f.endswith(".so")) if os.path.isdir(VAR) else []
    libs = [os.path.join(PKG, "rnc", "librnc.so")] + [os.path.join(VAR, f) for f in libs]
    ref = None
    for path in libs:
        lib = C.CDLL(path)
        lib.rnc_corr_lookup_umma_workspace_bytes.restype = C.c_size_t
        nb = lib.rnc_corr_lookup_umma_workspace_bytes(B, H, W)
        flags = torch.zeros(nb // 4 + 4, dtype=torch.int32, device=dev)
        fn = lib.rnc_corr_lookup_umma_fwd
        fn.restype = C.c_int
        vp = C.c_void_p
        args = [vp(f1h.data_ptr()), vp(f2h.data_ptr()), vp(f1.data_ptr()), vp(f2.data_ptr()), vp(coords.data_ptr()),
                B, D, H, W, L, 4, vp(hi.data_ptr()), vp(lo.data_ptr()), 352, 88, vp(flags.data_ptr()), C.c_size_t(nb),
                vp(torch.cuda.current_stream().cuda_stream)]
        hi.zero_(); lo.zero_()
        for _ in range(5):
            st = fn(*args)
            assert st == 0, st
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(int(os.environ.get("PROBE_REPS", "50"))):
            fn(*args)
        e1.record()
        torch.cuda.synchronize()
        val = (hi.float() + lo.float())
        if ref is None:
            ref = val.clone()
        err = (val - ref).abs().max().item()
        print(f"{os.path.basename(path):40s} {e0.elapsed_time(e1) / int(os.environ.get('PROBE_REPS', '50')) * 1000:8.1f} us/launch  "
              f"flags={int(flags.sum())}  maxdiff_vs_first={err:.3g}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(*[int(a) for a in sys.argv[2:5]])
