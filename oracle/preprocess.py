"""Oracle for the image transform tail: RandomCrop -> RandomHorizontalFlip -> ToTensor -> Normalize exactly as
vilmedic/datasets/base/ImageDataset.py:97-104 composes torchvision's transforms, applied per image to uint8 HWC arrays
(after the reference's Resize).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import torch


def crop_flip_normalize(images_u8, top, left, flip, crop, mean, std):
    """Plain-torch restatement: F.crop -> F.hflip -> to_tensor (float32 / 255) -> normalize (sub mean, div std)."""
    out = []
    m = torch.tensor(mean, dtype=torch.float32).view(3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(3, 1, 1)
    for b in range(images_u8.shape[0]):
        img = images_u8[b].permute(2, 0, 1)                                  # HWC -> CHW (ToTensor)
        img = img[:, int(top[b]):int(top[b]) + crop, int(left[b]):int(left[b]) + crop]
        if int(flip[b]):
            img = img.flip(-1)
        x = img.to(torch.float32).div(255)
        out.append(x.sub(m).div(s))
    return torch.stack(out)


def torchvision_train_transform(images_u8, crop, mean, std):
    """The reference's own Compose (minus Resize), on PIL images, consuming the global torch RNG like a DataLoader worker."""
    from PIL import Image
    from torchvision import transforms
    t = transforms.Compose([transforms.RandomCrop(crop), transforms.RandomHorizontalFlip(), transforms.ToTensor(),
                            transforms.Normalize(mean, std)])
    return torch.stack([t(Image.fromarray(images_u8[b].numpy())) for b in range(images_u8.shape[0])])
