"""On-GPU tail of the reference's image transform.

The reference's train transform is `Resize(resize) -> RandomCrop(crop) -> RandomHorizontalFlip() -> ToTensor() ->
Normalize(mean, std)` per image on DataLoader workers (vilmedic/datasets/base/ImageDataset.py:97-104), handing fp32
[B,3,crop,crop] tensors (602 KB / image) to the model.  `GpuImageTransform` keeps Resize on the host (PIL) and moves the
rest to the device: the batch crosses PCIe as uint8 HWC (4x fewer bytes than fp32, and before the crop only
resize^2 / crop^2 = 1.3x more pixels), the random crop origin / flip decisions are drawn on the host with EXACTLY the calls
torchvision makes, in the same order (RandomCrop.get_params: two `torch.randint`; RandomHorizontalFlip: one
`torch.rand(1) < p`), so a seeded run selects the same pixels as the reference pipeline, and one kernel produces the
normalised fp32 NCHW batch, bit-identical to the CPU transform.
"""
import torch

from ... import ops

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class GpuImageTransform:
    def __init__(self, crop=224, mean=IMAGENET_MEAN, std=IMAGENET_STD, train=True, flip_p=0.5):
        self.crop, self.mean, self.std, self.train, self.flip_p = int(crop), tuple(mean), tuple(std), bool(train), float(flip_p)

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
        B, H, W, _ = images_u8.shape
        if not self.train and (H != self.crop or W != self.crop):
            raise ValueError("evaluation images must already be resized to (%d, %d)" % (self.crop, self.crop))
        top, left, flip = params if params is not None else self.draw(B, H, W)
        dev = torch.device(device) if device is not None else (images_u8.device if images_u8.is_cuda else torch.device("cuda"))
        x = images_u8.contiguous().to(dev, non_blocking=True)
        return ops.image_crop_flip_normalize(x, top.to(dev), left.to(dev), flip.to(dev), self.crop, self.mean, self.std)
