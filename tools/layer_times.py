"""Per-launch device times of one conv-stack forward at the C2 shape (--c4: the C4 shape), and plain memory-bandwidth
reference points (write-only / read-only / copy) on the same box.
    gpurun -- 'python tools/layer_times.py > gpurun_out/layer_times.log'"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import _lib  # noqa: E402
from video_super_resolution_b200.my_packages.SRProjection.SRProjectionModule import SRProjectionModule  # noqa: E402


def bw():
    n = 1 << 30
    a = torch.empty(n, dtype=torch.float32, device="cuda")
    b = torch.empty_like(a)

    def t(fn, nbytes, name):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name:28s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s", flush=True)

    t(lambda: a.zero_(), 4 * n, "write-only (zero_ 4 GiB)")
    t(lambda: a.sum(), 4 * n, "read-only (sum 4 GiB)")
    t(lambda: b.copy_(a), 8 * n, "copy (4 GiB -> 4 GiB)")
    del a, b


def main():
    if "--no-bw" not in sys.argv:
        bw()
    M, h, w, scale = 20, 270, 480, 4
    if "--c4" in sys.argv:        # config C4: 2x, 1080p -> 4K, 5-frame window
        M, h, w, scale = 14, 1080, 1920, 2
    torch.manual_seed(0)
    sr = SRProjectionModule(num_maps=M, upscale_factor=scale)
    x = (torch.rand((M, 3, h, w)) * 255).cuda()
    for _ in range(2):
        sr(x)
    sr.profile(True)
    sr(x)
    L = _lib.lib()
    ent = next(iter(sr._plans.values()))
    cap = 512
    ms = (ctypes.c_float * cap)()
    kc = (ctypes.c_int32 * cap)()
    n = L.vsr_srfbn_profile_launches(ent["plan"], ms, kc, cap)
    tot = {}
    for i in range(n):
        name = L.vsr_srfbn_kernel_class_name(kc[i]).decode()
        tot[name] = tot.get(name, 0.0) + ms[i]
        if "--summary" not in sys.argv:
            print(f"{i:4d} {name:28s} {ms[i] * 1e3:9.1f} us")
    for k, v in tot.items():
        print(f"class {k:28s} {v:9.3f} ms")
    print("total", sum(ms[i] for i in range(n)), "ms")


if __name__ == "__main__":
    main()
