import os, sys, torch
sys.path.insert(0, "/root/repo" if os.path.isdir("/root/repo/tools") else os.getcwd())
from video_super_resolution_b200 import ops, synthetic
from tools.time_ops import timeit, PEAK
B,h,w=8,1080,1920
P=B*h*w
f=synthetic.smooth_flow(B,h,w,8.0,seed=1).cuda()
src=(torch.rand((B,h,w,3),device="cuda")*255)
lab=synthetic.labels(B,h,w).cuda()
for _ in range(3):
    ms=timeit(lambda: ops.warp(src,f,2,ref=src)); print("warp C3 fast+norm", ms*1e3/B, 48*P/ms/1e6/PEAK)
    ms=timeit(lambda: ops.warp(src,f,2)); print("warp C3 fast", ms*1e3/B, 32*P/ms/1e6/PEAK)
    ms=timeit(lambda: ops.warp_labels(lab,f)); print("labels", ms*1e3/B, 10*P/ms/1e6/PEAK)
