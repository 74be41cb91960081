"""Drop-in for `core/utils/utils.py`: InputPadder, coords_grid, upflow8, bilinear_sampler (host-side helpers the
evaluation loops call around the model, evaluate.py:125-129)."""
import torch
import torch.nn.functional as F


class InputPadder:
    """Replicate-pads frames to a multiple of 8 (utils.py:7-25): 'sintel' splits the padding evenly, otherwise it is
    put at the bottom."""

    def __init__(self, dims, mode="sintel"):
        self.ht, self.wd = dims[-2:]
        ph = (8 - self.ht % 8) % 8
        pw = (8 - self.wd % 8) % 8
        if mode == "sintel":
            self._pad = [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]
        else:
            self._pad = [pw // 2, pw - pw // 2, 0, ph]

    def pad(self, *inputs):
        return [F.pad(x, self._pad, mode="replicate") for x in inputs]

    def unpad(self, x):
        l, r, t, b = self._pad
        return x[..., t:x.shape[-2] - b, l:x.shape[-1] - r]


def coords_grid(batch, ht, wd):
    ys, xs = torch.meshgrid(torch.arange(ht), torch.arange(wd), indexing="ij")
    return torch.stack([xs, ys], 0).float()[None].repeat(batch, 1, 1, 1)


def upflow8(flow, mode="bilinear"):
    return 8 * F.interpolate(flow, size=(8 * flow.shape[2], 8 * flow.shape[3]), mode=mode, align_corners=True)
