"""SRProjectionModule -- the fusion / upsampling convolutions (SRFBN + per-pixel fc over the stacked
maps) with the reference's surface:

    SRProjectionModule(in_channels=3, out_channels=3, num_features=32, upscale_factor=4,
                       num_steps=3, num_groups=6, act_type='prelu', norm_type=None)
    .forward(x (M,3,h,w) fp32 0..255) -> (1,3,s*h,s*w) fp32
    ref: my_packages/SRProjection/SRProjectionModule.py:96-150

Parameters carry the reference's state-dict names and shapes, so a
checkpoint written by the reference's main.py:233-237 loads with load_state_dict unchanged; the only
generalisation is `num_maps` (the reference hard-wires fc in-features to 8 = 3*3-1 stacked maps,
:127).  The arithmetic does not run in torch: forward packs the weights to BF16 GEMM operands
(once per weight version) and calls the tcgen05 implicit-GEMM plan of libvsr_b200.so through the
C ABI (vsr_srfbn_*).  The FeedbackBlock follows the INTENDED dense-concat dataflow (Appendix C);
the reference's own forward reads uninitialised memory (:55-59,:70-74).

Inference only (the reference's hot loop runs under no_grad, main.py:199-203): there is no
autograd through the CUDA path.
"""
import ctypes

import torch
import torch.nn as nn

from ... import _lib

RGB_MEAN = (0.4488, 0.4371, 0.4040)   # SRProjectionModule.py:105
# (kernel, stride, padding) of the up/down projection units.  x4 is the reference's hard-wired geometry
# (SRProjectionModule.py:10-12,101-103); x2 is SRFBN's, needed by BASELINE config C4.
GEOMETRY = {4: (8, 4, 2), 2: (6, 2, 2)}


def _block(conv, act=True):
    """ConvBlock / DeconvBlock as nn.Sequential so that parameter names are `<name>.0.weight`,
    `<name>.0.bias`, `<name>.1.weight` (blocks.py:7-43,64-74)."""
    return nn.Sequential(conv, nn.PReLU(num_parameters=1, init=0.2)) if act else nn.Sequential(conv)


class _MeanShift(nn.Conv2d):
    """blocks.py:46-55: frozen 1x1 identity conv with bias sign*255*mean."""

    def __init__(self, rgb_mean, sign=-1):
        super().__init__(3, 3, kernel_size=1)
        self.weight.data = torch.eye(3).view(3, 3, 1, 1)
        self.bias.data = sign * 255.0 * torch.tensor(rgb_mean)
        for p in self.parameters():
            p.requires_grad = False


class _FeedbackParams(nn.Module):
    """Parameter container with FeedbackBlock's names (SRProjectionModule.py:7-42)."""

    def __init__(self, nf, num_groups, ksp=(8, 4, 2)):
        super().__init__()
        self.compress_in = _block(nn.Conv2d(2 * nf, nf, 1))
        self.upBlocks = nn.ModuleList([_block(nn.ConvTranspose2d(nf, nf, *ksp)) for _ in range(num_groups)])
        self.downBlocks = nn.ModuleList([_block(nn.Conv2d(nf, nf, *ksp)) for _ in range(num_groups)])
        self.uptranBlocks = nn.ModuleList([_block(nn.Conv2d(nf * (i + 2), nf, 1)) for i in range(num_groups - 1)])
        self.downtranBlocks = nn.ModuleList([_block(nn.Conv2d(nf * (i + 2), nf, 1)) for i in range(num_groups - 1)])
        self.compress_out = _block(nn.Conv2d(num_groups * nf, nf, 1))


class SRProjectionModule(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, num_features=32, upscale_factor=4, num_steps=3, num_groups=6,
                 act_type='prelu', norm_type=None, num_maps=8, workspace_cap_bytes=None):
        super(SRProjectionModule, self).__init__()
        if (in_channels, out_channels, num_features, num_groups) != (3, 3, 32, 6) \
                or upscale_factor not in GEOMETRY or act_type != 'prelu' or norm_type is not None:
            raise NotImplementedError("the B200 path implements 3->3 channels, 32 features, 6 groups, PReLU, no "
                                      "norm, x4 (k8 s4 p2, the reference's geometry) or x2 (k6 s2 p2)")
        ksp = GEOMETRY[upscale_factor]
        self.num_steps = num_steps
        self.num_features = num_features
        self.upscale_factor = upscale_factor
        self.num_maps = num_maps
        # optional cap on the activation workspace: the plan then processes the maps in chunks (bit-identical results)
        self.workspace_cap_bytes = workspace_cap_bytes
        nf = num_features
        self.sub_mean = _MeanShift(RGB_MEAN, -1)
        self.conv_in = _block(nn.Conv2d(in_channels, 4 * nf, 3, padding=1))
        self.feat_in = _block(nn.Conv2d(4 * nf, nf, 1))
        self.block = _FeedbackParams(nf, num_groups, ksp)
        self.out = _block(nn.ConvTranspose2d(nf, nf, *ksp))
        self.conv_out = _block(nn.Conv2d(nf, out_channels, 3, padding=1), act=False)
        self.add_mean = _MeanShift(RGB_MEAN, 1)
        self.fc = nn.Sequential(nn.Linear(num_maps, 32), nn.ReLU(), nn.Linear(32, 1), nn.ReLU())
        self._plans = {}          # (h, w, device) -> dict(plan, weights, workspace, version)
        self._keep = []

    # -- C ABI plumbing ---------------------------------------------------------------------------
    def _weights_struct(self):
        """vsr_srfbn_weights over contiguous fp32 host copies of the parameters."""
        keep = []

        def fp(t):
            a = t.detach().to("cpu", torch.float32).contiguous()
            keep.append(a)
            return ctypes.cast(a.data_ptr(), ctypes.POINTER(ctypes.c_float))

        def sl(seq):
            return float(seq[1].weight.detach().reshape(-1)[0])

        W = _lib.SrfbnWeights()
        W.sub_mean_bias, W.add_mean_bias = fp(self.sub_mean.bias), fp(self.add_mean.bias)
        W.conv_in_w, W.conv_in_b, W.conv_in_slope = fp(self.conv_in[0].weight), fp(self.conv_in[0].bias), sl(self.conv_in)
        W.feat_in_w, W.feat_in_b, W.feat_in_slope = fp(self.feat_in[0].weight), fp(self.feat_in[0].bias), sl(self.feat_in)
        b = self.block
        W.compress_in_w, W.compress_in_b, W.compress_in_slope = \
            fp(b.compress_in[0].weight), fp(b.compress_in[0].bias), sl(b.compress_in)
        for i in range(6):
            W.up_w[i], W.up_b[i], W.up_slope[i] = fp(b.upBlocks[i][0].weight), fp(b.upBlocks[i][0].bias), sl(b.upBlocks[i])
            W.down_w[i], W.down_b[i], W.down_slope[i] = \
                fp(b.downBlocks[i][0].weight), fp(b.downBlocks[i][0].bias), sl(b.downBlocks[i])
        for i in range(5):
            W.uptran_w[i], W.uptran_b[i], W.uptran_slope[i] = \
                fp(b.uptranBlocks[i][0].weight), fp(b.uptranBlocks[i][0].bias), sl(b.uptranBlocks[i])
            W.downtran_w[i], W.downtran_b[i], W.downtran_slope[i] = \
                fp(b.downtranBlocks[i][0].weight), fp(b.downtranBlocks[i][0].bias), sl(b.downtranBlocks[i])
        W.compress_out_w, W.compress_out_b, W.compress_out_slope = \
            fp(b.compress_out[0].weight), fp(b.compress_out[0].bias), sl(b.compress_out)
        W.out_w, W.out_b, W.out_slope = fp(self.out[0].weight), fp(self.out[0].bias), sl(self.out)
        W.conv_out_w, W.conv_out_b = fp(self.conv_out[0].weight), fp(self.conv_out[0].bias)
        W.fc0_w, W.fc0_b = fp(self.fc[0].weight), fp(self.fc[0].bias)
        W.fc2_w, W.fc2_b = fp(self.fc[2].weight), fp(self.fc[2].bias)
        return W, keep

    def _version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _plan_for(self, M, h, w, device):
        key = (M, h, w, device.index)
        ent = self._plans.get(key)
        ver = self._version()
        L = _lib.lib()
        if ent is None:
            cfg = _lib.SrfbnConfig(M, h, w, self.num_steps, 6, self.num_features, self.upscale_factor)
            plan = ctypes.c_void_p()
            _lib.check(L.vsr_srfbn_plan_create(ctypes.byref(cfg), ctypes.byref(plan)), "srfbn_plan_create")
            if self.workspace_cap_bytes:
                _lib.check(L.vsr_srfbn_plan_set_workspace_cap(plan, int(self.workspace_cap_bytes)), "srfbn_plan_set_workspace_cap")
            ent = {"plan": plan, "version": None, "chunk_maps": int(L.vsr_srfbn_chunk_maps(plan)),
                   "weights": torch.empty(int(L.vsr_srfbn_weight_bytes(plan)), dtype=torch.uint8, device=device),
                   "workspace": torch.empty(int(L.vsr_srfbn_workspace_bytes(plan)), dtype=torch.uint8, device=device)}
            self._plans[key] = ent
        if ent["version"] != ver:
            W, keep = self._weights_struct()
            host = torch.empty(ent["weights"].numel(), dtype=torch.uint8).pin_memory()
            _lib.check(L.vsr_srfbn_pack_weights(ent["plan"], ctypes.byref(W), host.data_ptr()), "srfbn_pack_weights")
            ent["weights"].copy_(host)
            torch.cuda.current_stream(device).synchronize()
            del keep
            _lib.check(L.vsr_srfbn_bind(ent["plan"], ent["weights"].data_ptr(), ent["workspace"].data_ptr(),
                                        ent["workspace"].numel()), "srfbn_bind")
            ent["version"] = ver
            ent["have_premix"] = False          # bind drops the second layer list and the per-map images
            ent["refresh_first"] = None
        return ent

    def forward(self, x, out_u8=None, want_f32=True, changed_from=None):
        """out_u8: optional (s*h, s*w, 3) u8 CUDA tensor that additionally receives the frame as clamp(y,0,255) rounded
        half to even (the loader's pixel format, utils/video_utils.py:23), written by the fc-fuse kernel itself;
        want_f32=False skips the fp32 frame (then the u8 frame is the only output and is what is returned).
        changed_from=k: the caller states that, since the previous forward of this module on a stack of this shape, only
        the maps x[k:] changed (the fuse pass, network/video_super_resolution.py:62, feeds the frames unchanged): the
        maps are independent until the per-pixel fc, so only x[k:] go through the conv stack again and the per-map images
        of x[:k] are reused -- bit-identical to a full forward."""
        if not x.is_cuda:
            raise RuntimeError("SRProjectionModule: CUDA tensors only (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[0] != self.num_maps:
            raise ValueError(f"SRProjectionModule: expected ({self.num_maps},3,h,w), got {tuple(x.shape)}")
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.to(torch.float32).contiguous()
        M, _, h, w = x.shape
        ent = self._plan_for(M, h, w, x.device)
        s = self.upscale_factor
        if out_u8 is not None and (not out_u8.is_cuda or out_u8.dtype != torch.uint8 or not out_u8.is_contiguous()
                                   or tuple(out_u8.shape) != (s * h, s * w, 3)):
            raise ValueError(f"SRProjectionModule: out_u8 must be a contiguous CUDA u8 tensor of shape {(s * h, s * w, 3)}")
        if out_u8 is None and not want_f32:
            raise ValueError("SRProjectionModule: nothing to compute")
        y = torch.empty((1, 3, s * h, s * w), dtype=torch.float32, device=x.device) if want_f32 else None
        L = _lib.lib()
        with torch.cuda.device(x.device):
            yp = y.data_ptr() if y is not None else None
            up = out_u8.data_ptr() if out_u8 is not None else None
            st = torch.cuda.current_stream().cuda_stream
            if changed_from is None or changed_from <= 0:
                _lib.check(L.vsr_srfbn_forward_u8(ent["plan"], x.data_ptr(), yp, up, st), "srfbn_forward")
                ent["have_premix"] = True
            else:
                if not ent.get("have_premix"):
                    raise RuntimeError("SRProjectionModule: changed_from needs a preceding full forward on this shape")
                if ent.get("refresh_first") != int(changed_from):
                    _lib.check(L.vsr_srfbn_prepare_refresh(ent["plan"], int(changed_from)), "srfbn_prepare_refresh")
                    ent["refresh_first"] = int(changed_from)
                _lib.check(L.vsr_srfbn_forward_refresh_u8(ent["plan"], x.data_ptr(), yp, up, int(changed_from), st),
                           "srfbn_forward_refresh")
        return y if want_f32 else out_u8

    def premix(self, x):
        """Test hook: per-map outputs before the fc fuse, (M,3,4h,4w) (SRProjectionModule.py:143)."""
        self.forward(x)
        M, _, h, w = x.shape
        ent = self._plans[(M, h, w, x.device.index)]
        s = self.upscale_factor
        out = torch.empty((M, 3, s * h, s * w), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().vsr_srfbn_debug_premix(ent["plan"], out.data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream), "srfbn_debug_premix")
        return out

    def group_error(self):
        """Test hook: True if a co-scheduled group launch gave up waiting (see vsr_srfbn_debug_group_error)."""
        L = _lib.lib()
        return any(L.vsr_srfbn_debug_group_error(ent["plan"], torch.cuda.current_stream().cuda_stream) != 0
                   for ent in self._plans.values())

    def profile(self, enable=True):
        """Bracket every kernel launch of subsequent forwards with CUDA events (bench.py)."""
        for ent in self._plans.values():
            _lib.check(_lib.lib().vsr_srfbn_profile_enable(ent["plan"], int(bool(enable))), "srfbn_profile_enable")

    def profile_read(self):
        """{kernel class: dict(ms, launches, flops, bytes)} of the last forward of each plan, summed."""
        L = _lib.lib()
        n = 11            # VSR_SRFBN_KERNEL_CLASSES
        out = {}
        for ent in self._plans.values():
            ms = (ctypes.c_double * n)()
            la = (ctypes.c_int32 * n)()
            fl = (ctypes.c_double * n)()
            by = (ctypes.c_double * n)()
            _lib.check(L.vsr_srfbn_profile_read(ent["plan"], ms, la, fl, by), "srfbn_profile_read")
            for k in range(n):
                name = L.vsr_srfbn_kernel_class_name(k).decode()
                d = out.setdefault(name, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
                d["ms"] += ms[k]
                d["launches"] += la[k]
                d["flops"] += fl[k]
                d["bytes"] += by[k]
        return out

    def __del__(self):
        try:
            L = _lib.lib()
            for ent in self._plans.values():
                L.vsr_srfbn_plan_destroy(ent["plan"])
        except Exception:
            pass
