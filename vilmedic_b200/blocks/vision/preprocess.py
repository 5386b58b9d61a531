"""On-GPU tail of the reference's image transform.

The reference's train transform is `Resize(resize) -> RandomCrop(crop) -> RandomHorizontalFlip() -> ToTensor() ->
Normalize(mean, std)` per image on DataLoader workers (vilmedic/datasets/base/ImageDataset.py:97-104), handing fp32
[B,3,crop,crop] tensors (602 KB / image) to the model.  `GpuImageTransform` runs the whole chain on the device — `GpuResize` is
Pillow's fixed-point separable resampling (what torchvision's Resize does to a PIL image), bit for bit, for batches of equally sized
images (`resize=` given); without it the Resize stays on the host (PIL) and the rest moves to the device: the batch crosses PCIe as uint8 HWC (4x fewer bytes than fp32, and before the crop only
resize^2 / crop^2 = 1.3x more pixels), the random crop origin / flip decisions are drawn on the host with EXACTLY the calls
torchvision makes, in the same order (RandomCrop.get_params: two `torch.randint`; RandomHorizontalFlip: one
`torch.rand(1) < p`), so a seeded run selects the same pixels as the reference pipeline, and one kernel produces the
normalised fp32 NCHW batch, bit-identical to the CPU transform.
"""
import math

import torch

from ... import ops

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def pil_resample_tables(in_size, out_size):
    """Coefficient tables of Pillow's bilinear resampling for one axis (src/libImaging/Resample.c: precompute_coeffs +
    normalize_coeffs_8bpc), in the same double arithmetic: -> (bounds int32 [out, 2], coefs int32 [out, ksize])."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale                      # bilinear_filter support = 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = torch.zeros((out_size, 2), dtype=torch.int32)
    coefs = torch.zeros((out_size, ksize), dtype=torch.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = []
        ww = 0.0
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            w = 1.0 - abs(v) if abs(v) < 1.0 else 0.0
            k.append(w)
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
            coefs[xx, x] = int(0.5 + k[x] * (1 << 22)) if k[x] >= 0 else int(-0.5 + k[x] * (1 << 22))
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return bounds, coefs


def resize_output_size(h, w, size):
    """torchvision.transforms.Resize(int): the smaller edge becomes `size`, aspect ratio kept (functional._compute_resized_output_size)."""
    if isinstance(size, (tuple, list)) and len(size) == 2:
        return int(size[0]), int(size[1])
    size = int(size[0]) if isinstance(size, (tuple, list)) else int(size)
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long_ / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


class GpuResize:
    """`transforms.Resize(resize)` of the reference's transform (ImageDataset.py:99) on uint8 HWC batches on the device, bit-identical to
    Pillow's BILINEAR resize (tests/test_ops_gpu.py).  Tables are cached per (input size, output size)."""

    def __init__(self, size):
        self.size = size
        self._tables = {}

    def __call__(self, images_u8):
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise TypeError("GpuResize expects uint8 [B,H,W,3] images")
        x = images_u8.contiguous()
        if not x.is_cuda:
            x = x.cuda(non_blocking=True)
        B, H, W, _ = x.shape
        oh, ow = resize_output_size(H, W, self.size)
        if (oh, ow) == (H, W):
            return x
        key = (H, W, oh, ow, x.device)
        if key not in self._tables:
            bh, ch = pil_resample_tables(W, ow)
            bv, cv = pil_resample_tables(H, oh)
            self._tables[key] = tuple(t.to(x.device) for t in (bh, ch, bv, cv))
        bh, ch, bv, cv = self._tables[key]
        if ow != W:                                  # Pillow: horizontal pass first, vertical pass on its 8-bit result
            x = ops.image_resample_u8(x, bh, ch, H, ow, 0)
        if oh != H:
            x = ops.image_resample_u8(x, bv, cv, oh, ow, 1)
        return x


class GpuImageTransform:
    def __init__(self, crop=224, mean=IMAGENET_MEAN, std=IMAGENET_STD, train=True, flip_p=0.5, resize=None):
        self.crop, self.mean, self.std, self.train, self.flip_p = int(crop), tuple(mean), tuple(std), bool(train), float(flip_p)
        self.resize = GpuResize(resize) if resize is not None else None       # Resize(resize) on the device too (same-size batches)

    def draw(self, B, H, W):
        """Per-image (top, left, flip) with torchvision's RNG call sequence; evaluation: no crop offset randomness."""
        th = tw = self.crop
        if H < th or W < tw:
            raise ValueError("Required crop size %s is larger than input image size %s" % ((th, tw), (H, W)))
        top, left, flip = [], [], []
        for _ in range(B):
            if not self.train:
                top.append(0), left.append(0), flip.append(0)
                continue
            if W == tw and H == th:
                i = j = 0
            else:
                i = torch.randint(0, H - th + 1, size=(1,)).item()
                j = torch.randint(0, W - tw + 1, size=(1,)).item()
            f = bool(torch.rand(1) < self.flip_p)
            top.append(i), left.append(j), flip.append(int(f))
        return (torch.tensor(top, dtype=torch.int32), torch.tensor(left, dtype=torch.int32), torch.tensor(flip, dtype=torch.uint8))

    def __call__(self, images_u8, params=None, device=None):
        """images_u8: uint8 [B,H,W,3] (host, ideally pinned, or device).  Returns fp32 [B,3,crop,crop] on the device."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise TypeError("GpuImageTransform expects uint8 [B,H,W,3] images")
        if self.resize is not None:
            images_u8 = self.resize(images_u8)
        B, H, W, _ = images_u8.shape
        if not self.train and (H != self.crop or W != self.crop):
            raise ValueError("evaluation images must already be resized to (%d, %d)" % (self.crop, self.crop))
        top, left, flip = params if params is not None else self.draw(B, H, W)
        dev = torch.device(device) if device is not None else (images_u8.device if images_u8.is_cuda else torch.device("cuda"))
        x = images_u8.contiguous().to(dev, non_blocking=True)
        return ops.image_crop_flip_normalize(x, top.to(dev), left.to(dev), flip.to(dev), self.crop, self.mean, self.std)
