"""Drop-in for the flow codecs of `core/utils/frame_utils.py` that the evaluation writers use
(evaluate.py:46-47 `.flo` submissions, :101-103 KITTI 16-bit PNG).  Host-side I/O, no GPU involved."""
import numpy as np

FLO_MAGIC = np.float32(202021.25)          # frame_utils.py:7 TAG_CHAR


def writeFlow(filename, uv, v=None):
    """Middlebury .flo (frame_utils.py:70-99): float32 magic, int32 width, int32 height, then row-major interleaved (u, v)."""
    uv = np.asarray(uv)
    if v is None:
        if uv.ndim != 3 or uv.shape[2] != 2:
            raise ValueError("expected an [H, W, 2] flow array")
        u, v = uv[:, :, 0], uv[:, :, 1]
    else:
        u = uv
    if u.shape != v.shape:
        raise ValueError("u and v must have the same shape")
    h, w = u.shape
    data = np.empty((h, w, 2), np.float32)
    data[..., 0], data[..., 1] = u, v
    with open(filename, "wb") as f:
        f.write(FLO_MAGIC.tobytes())
        f.write(np.int32(w).tobytes())
        f.write(np.int32(h).tobytes())
        f.write(data.tobytes())


def readFlow(filename):
    """frame_utils.py:11-30 — returns [H, W, 2] float32, or None on a bad magic number (as the reference does)."""
    with open(filename, "rb") as f:
        magic = np.frombuffer(f.read(4), np.float32, count=1)
        if magic.size != 1 or magic[0] != FLO_MAGIC:
            print("Magic number incorrect. Invalid .flo file")
            return None
        w = int(np.frombuffer(f.read(4), np.int32, count=1)[0])
        h = int(np.frombuffer(f.read(4), np.int32, count=1)[0])
        data = np.frombuffer(f.read(8 * w * h), np.float32, count=2 * w * h)
    return data.reshape(h, w, 2).copy()


def writeFlowKITTI(filename, uv):
    """frame_utils.py:116-120 — uint16 PNG, channels (valid=1, v, u) in file order, value = 64*flow + 2^15."""
    import cv2
    enc = 64.0 * np.asarray(uv, np.float64) + 2 ** 15
    valid = np.ones(enc.shape[:2] + (1,))
    cv2.imwrite(filename, np.concatenate([enc, valid], -1).astype(np.uint16)[..., ::-1])


def readFlowKITTI(filename):
    """frame_utils.py:102-107 — returns (flow [H,W,2] float32, valid [H,W] float32)."""
    import cv2
    raw = cv2.imread(filename, cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)[:, :, ::-1].astype(np.float32)
    return (raw[:, :, :2] - 2 ** 15) / 64.0, raw[:, :, 2]


def readPFM(filename):
    """frame_utils.py:33-68 — PFM ('PF' colour / 'Pf' grey), scale < 0 = little endian; rows are stored bottom-up."""
    import re
    with open(filename, "rb") as f:
        header = f.readline().rstrip()
        if header == b"PF":
            color = True
        elif header == b"Pf":
            color = False
        else:
            raise Exception("Not a PFM file.")
        dim = re.match(rb"^(\d+)\s(\d+)\s$", f.readline())
        if not dim:
            raise Exception("Malformed PFM header.")
        width, height = int(dim.group(1)), int(dim.group(2))
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        data = np.frombuffer(f.read(), dtype=endian + "f4")
    shape = (height, width, 3) if color else (height, width)
    return np.flipud(np.reshape(data, shape))


def readDispKITTI(filename):
    """frame_utils.py:109-113 — 16-bit disparity PNG -> (flow [H,W,2] with u = -disp, valid)."""
    import cv2
    disp = cv2.imread(filename, cv2.IMREAD_ANYDEPTH) / 256.0
    valid = disp > 0.0
    return np.stack([-disp, np.zeros_like(disp)], -1), valid


def read_gen(file_name, pil=False):
    """frame_utils.py:123-142 — dispatch on the file extension, as the datasets do."""
    from os.path import splitext
    ext = splitext(file_name)[-1]
    if ext in (".png", ".jpeg", ".ppm", ".jpg", ".webp"):
        from PIL import Image
        return Image.open(file_name)
    if ext in (".bin", ".raw"):
        return np.load(file_name)
    if ext == ".flo":
        return readFlow(file_name).astype(np.float32)
    if ext == ".pfm":
        flow = readPFM(file_name).astype(np.float32)
        return flow if flow.ndim == 2 else flow[:, :, :-1]
    if ext == ".npz":
        return np.load(file_name)["optical_flow"].astype(np.float32).transpose(1, 2, 0)
    return []
