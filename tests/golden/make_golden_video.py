"""Generates tests/golden/video_chunks.npz by RUNNING THE REFERENCE's utils/video_utils.py:7-33 (VideoDataset)
on synthetic MJPG videos whose frame k is a flat colour encoding k (frame_code), and main.py:155-159 (LR maker).

    python tests/golden/make_golden_video.py        (build container only: needs /root/reference)

Stored per video length: the number of chunks, and for every chunk the index of the first frame of each of
its windows (decoded back from the colour) -- the structure the recurrence reset follows (main.py:196).
"""
import os
import sys
import tempfile

import cv2
import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def frame_code(k, h, w):
    """frame k as a flat colour that survives the lossy codec: BGR = (16, 32*((k//8)%8)+16, 32*(k%8)+16)"""
    f = np.empty((h, w, 3), np.uint8)
    f[..., 0], f[..., 1], f[..., 2] = 16, 32 * ((k // 8) % 8) + 16, 32 * (k % 8) + 16
    return f


def frame_decode(rgb):
    """inverse of frame_code on a decoded RGB frame"""
    r, g = float(rgb[..., 0].mean()), float(rgb[..., 1].mean())
    return int(round((r - 16) / 32)) + 8 * int(round((g - 16) / 32))


def write_video(path, n, h=32, w=48):
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (w, h))
    assert vw.isOpened()
    for k in range(n):
        vw.write(frame_code(k, h, w))
    vw.release()


def main():
    sys.path.insert(0, REF)
    from utils.video_utils import VideoDataset  # noqa: E402
    from utils.tools import transpose1312, transpose1323  # noqa: E402
    from torch.nn.functional import interpolate  # noqa: E402
    out = {}
    for n in (40, 60, 64, 50):
        with tempfile.TemporaryDirectory() as d:
            write_video(os.path.join(d, "v.avi"), n)
            ds = VideoDataset(d)
            ds.read_video(ds.video_paths[0])
            chunks = ds.data
            first = [[frame_decode(win[0]) for win in ch] for ch in chunks]
            out[f"{n}/num_chunks"] = np.array(len(chunks))
            out[f"{n}/lengths"] = np.array([len(c) for c in chunks])
            out[f"{n}/first_frames"] = np.array([f for c in first for f in c])
            assert all(len(win) == 3 for ch in chunks for win in ch)
    # main.py:155-159: LR = interpolate(transpose1323(d.float()), (H/4, W/4)), default nearest
    g = torch.Generator().manual_seed(5)
    for name, (H, W) in {"div": (32, 48), "ragged": (30, 50)}.items():
        d = torch.randint(0, 256, (3, H, W, 3), generator=g, dtype=torch.uint8)
        lr = transpose1312(interpolate(transpose1323(d.type(torch.float32)), (int(d.shape[1] / 4), int(d.shape[2] / 4))))
        out[f"lr/{name}/hr"] = d.numpy()
        out[f"lr/{name}/lr"] = lr.contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "video_chunks.npz"), **out)
    print({k: (v.tolist() if v.size < 30 else v.shape) for k, v in out.items() if not k.startswith("lr/")})


if __name__ == "__main__":
    main()
