"""Layout helpers with the reference's names (utils/tools.py:76-77,102-123).
They return stride-permuted views, exactly like the reference's chained transposes."""
import torch


def maskprocess(mask: torch.Tensor) -> torch.Tensor:
    """(h,w) -> (3,h,w): tile a single-channel map over RGB (utils/tools.py:76-77)."""
    return mask.unsqueeze(0).expand(3, *mask.shape).contiguous()


def transpose1323(x):  # (N,H,W,C) -> (N,C,H,W)
    return x.permute(0, 3, 1, 2)


def transpose1223(x):  # (N,C,H,W) -> (N,H,W,C)
    return x.permute(0, 2, 3, 1)


def transpose1312(x):  # (N,C,H,W) -> (N,H,W,C)
    return x.permute(0, 2, 3, 1)


def transpose1201(x):  # (H,W,C) -> (C,H,W)
    return x.permute(2, 0, 1)


def transpose030112(x):  # (M,C,H,W) -> (C,H,W,M)
    return x.permute(1, 2, 3, 0)


def transpose031323(x):  # (C,H,W,1) -> (1,C,H,W)
    return x.permute(3, 0, 1, 2)


class StaticCenterCrop(object):
    """utils/tools.py:8-14: centre crop of an (h,w,c) image to crop_size."""

    def __init__(self, image_size, crop_size):
        self.th, self.tw = crop_size
        self.h, self.w = image_size

    def __call__(self, img):
        y0 = (self.h - self.th) // 2
        x0 = (self.w - self.tw) // 2
        return img[y0:y0 + self.th, x0:x0 + self.tw, :]
