"""Video I/O either side of the hot path: the reference's window / chunk reader,
the LR maker of its training loop, a pinned-memory staging ring and a u8 frame writer.

    ref: utils/video_utils.py:7-33 (VideoDataset: cv2 decode, 3-frame windows, 20(+1) chunks),
         main.py:155-159 (LR = nearest-neighbour /4 of the decoded frames), main.py:190-203 (the
         recurrence over a chunk: estimated_image = None at the chunk start, then the previous output)

The on-disk input is whatever cv2.VideoCapture reads; the reference defines no output format
(testing / inference are TODOs in its ReadMe), so FrameWriter offers the two obvious ones: a
cv2.VideoWriter container or a raw .npy stack of (H,W,3) u8 RGB frames.

Host-side only: nothing here touches the GPU except PinnedFrameRing's async H2D copies.
"""
from __future__ import annotations

import os
from glob import glob

import numpy as np
import torch

SPLIT_VIDEO_NUM = 20   # utils/video_utils.py:11


def window_starts(n_frames: int, T: int = 3) -> range:
    """Windows of T consecutive frames, one per start index (utils/video_utils.py:25, T = 3 there)."""
    return range(max(n_frames - (T - 1), 0))


def chunk_windows(n_frames: int, T: int = 3, splitvideonum: int = SPLIT_VIDEO_NUM) -> list[range]:
    """The reference's chunking of a video's windows (utils/video_utils.py:26-27):
    `for i in range(0, length, int(length/splitvideonum)): data[i:i+int(length/splitvideonum)]` with
    length = the FRAME count, data = the list of windows.  That yields splitvideonum (+1 when the
    division leaves a remainder) slices; trailing slices can be short or empty -- kept as they are,
    because the recurrence restarts at every chunk (main.py:196) and the chunk boundaries therefore
    define the result."""
    n_win = len(window_starts(n_frames, T))
    step = int(n_frames / splitvideonum)
    if step < 1:
        raise ValueError(f"video of {n_frames} frames is shorter than splitvideonum={splitvideonum} "
                         "(the reference divides by zero here, utils/video_utils.py:26)")
    return [range(min(i, n_win), min(i + step, n_win)) for i in range(0, n_frames, step)]


def read_video(path: str) -> np.ndarray:
    """All frames of a video as (N,H,W,3) u8 RGB (utils/video_utils.py:17-24)."""
    import cv2
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise IOError(f"cannot open video {path!r}")
    frames = []
    while True:
        ret, img = cap.read()
        if not ret:
            break
        frames.append(cv2.cvtColor(img, cv2.COLOR_BGR2RGB))
    cap.release()
    if not frames:
        raise IOError(f"no frames decoded from {path!r}")
    return np.stack(frames)


class VideoDataset:
    """Same surface as the reference's torch Dataset (utils/video_utils.py:7-33): item k is one chunk =
    a list of windows, each a list of T (H,W,3) u8 RGB frames; a video is decoded when its first chunk
    is requested.  `__len__` is videos * (splitvideonum + 1) as in the reference."""

    def __init__(self, path, window: int = 3):
        self.video_paths = sorted(glob(os.path.join(path, '*')))
        self.data = []
        self.window = window
        self.splitvideonum = SPLIT_VIDEO_NUM
        self.truthsplitvideonum = self.splitvideonum + 1

    def __len__(self):
        return len(self.video_paths) * self.truthsplitvideonum

    def read_video(self, file):
        imgs = read_video(file)
        T = self.window
        windows = [list(imgs[i:i + T]) for i in window_starts(len(imgs), T)]
        for r in chunk_windows(len(imgs), T, self.splitvideonum):
            self.data.append(windows[r.start:r.stop])

    def __getitem__(self, idx):
        if idx % self.truthsplitvideonum == 0:
            self.read_video(self.video_paths[idx // self.truthsplitvideonum])
        data = self.data[0]
        self.data = self.data[1:]
        return data


def make_lr(frames_u8: torch.Tensor, scale: int = 4) -> torch.Tensor:
    """(N,H,W,3) u8/float -> (N,H//s,W//s,3) fp32: the LR input of the training loop, main.py:155-159
    (`interpolate(transpose1323(d.float()), (H/4, W/4))`, default mode nearest = every s-th pixel when s
    divides the size, ATen's floor(dst*in/out) otherwise)."""
    f = frames_u8.to(torch.float32)
    N, H, W, _ = f.shape
    h, w = int(H / scale), int(W / scale)
    if h * scale == H and w * scale == W:
        return f[:, ::scale, ::scale].contiguous()
    x = torch.nn.functional.interpolate(f.permute(0, 3, 1, 2), (h, w))
    return x.permute(0, 2, 3, 1).contiguous()


class PinnedFrameRing:
    """Host staging for the frames of a rank's shard: `depth` pinned buffers, filled by the caller
    (decoder thread or synthetic generator) and copied to the device on a dedicated copy stream, so
    the H2D of window k+1 overlaps the kernels of window k.  `put` returns a device tensor plus an
    event the compute stream must wait on."""

    def __init__(self, shape, dtype=torch.float32, depth: int = 3, device="cuda:0"):
        self.device = torch.device(device)
        self.host = [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(depth)]
        self.dev = [torch.empty(shape, dtype=dtype, device=self.device) for _ in range(depth)]
        self.free_ev = [None] * depth          # compute-side event: slot's device tensor no longer read
        self.ready_ev = [None] * depth         # copy-side event: slot's previous H2D has left the host buffer
        self.stream = torch.cuda.Stream(device=self.device)
        self.k = 0

    def put(self, src: torch.Tensor):
        i = self.k % len(self.host)
        self.k += 1
        if self.ready_ev[i] is not None:
            self.ready_ev[i].synchronize()     # the slot's previous H2D may still be reading the pinned buffer (also
                                               # when the caller never called release() for it)
        if self.free_ev[i] is not None:
            self.free_ev[i].synchronize()      # the device tensor is about to be overwritten
            self.stream.wait_event(self.free_ev[i])
        self.host[i].copy_(src)
        with torch.cuda.stream(self.stream):
            self.dev[i].copy_(self.host[i], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        self.ready_ev[i] = ready
        return self.dev[i], ready, i

    def release(self, slot: int):
        """Call on the compute stream after the last kernel that reads the slot has been enqueued."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free_ev[slot] = ev


class FrameWriter:
    """(H,W,3) u8 RGB frames -> a video container (cv2.VideoWriter, by extension) or a .npy stack."""

    def __init__(self, path: str, fps: float = 30.0, fourcc: str = "mp4v"):
        self.path = path
        self.fps = fps
        self.fourcc = fourcc
        self._raw = path.endswith(".npy")
        self._frames = []
        self._vw = None

    def write(self, frame_u8):
        a = frame_u8.cpu().numpy() if isinstance(frame_u8, torch.Tensor) else np.asarray(frame_u8)
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
            raise ValueError("FrameWriter: (H,W,3) u8 frames expected")
        if self._raw:
            self._frames.append(a.copy())
            return
        import cv2
        if self._vw is None:
            self._vw = cv2.VideoWriter(self.path, cv2.VideoWriter_fourcc(*self.fourcc), self.fps, (a.shape[1], a.shape[0]))
            if not self._vw.isOpened():
                raise IOError(f"cannot open {self.path!r} for writing")
        self._vw.write(cv2.cvtColor(a, cv2.COLOR_RGB2BGR))

    def close(self):
        if self._raw:
            np.save(self.path, np.stack(self._frames) if self._frames else np.zeros((0, 0, 0, 3), np.uint8))
        elif self._vw is not None:
            self._vw.release()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
