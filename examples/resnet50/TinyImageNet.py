"""Tiny-ImageNet-200 dataset reader (the reference ships one next to its ResNet-50 provider: examples/resnet50/TinyImageNet.py).

Directory layout of the public archive:
    <root>/wnids.txt                                  200 class ids, one per line
    <root>/train/<wnid>/images/<wnid>_<k>.JPEG        500 images per class, 64 x 64
    <root>/val/images/val_<k>.JPEG + <root>/val/val_annotations.txt   (file, wnid, box) per line
There is no network in the build image, so the provider falls back to synthetic tensors of the same shape unless
``TINY_IMAGENET_ROOT`` points at an extracted copy.
"""
import os
from typing import Callable, List, Optional, Tuple

import torch
from torch.utils.data import Dataset


class TinyImageNet(Dataset):
    def __init__(self, root: str, split: str = 'train', transform: Optional[Callable] = None, image_size: int = 224):
        assert split in ('train', 'val')
        self.root, self.split, self.transform, self.image_size = root, split, transform, image_size
        with open(os.path.join(root, 'wnids.txt')) as fh:
            self.classes = [line.strip() for line in fh if line.strip()]
        self.class_to_idx = {w: i for i, w in enumerate(self.classes)}
        self.samples: List[Tuple[str, int]] = []
        if split == 'train':
            for wnid in self.classes:
                folder = os.path.join(root, 'train', wnid, 'images')
                for name in sorted(os.listdir(folder)):
                    if name.lower().endswith(('.jpeg', '.jpg', '.png')):
                        self.samples.append((os.path.join(folder, name), self.class_to_idx[wnid]))
        else:
            with open(os.path.join(root, 'val', 'val_annotations.txt')) as fh:
                for line in fh:
                    parts = line.split('\t')
                    if len(parts) >= 2 and parts[1] in self.class_to_idx:
                        self.samples.append((os.path.join(root, 'val', 'images', parts[0]), self.class_to_idx[parts[1]]))

    def __len__(self) -> int:
        return len(self.samples)

    def __getitem__(self, i: int):
        from PIL import Image
        path, label = self.samples[i]
        with Image.open(path) as im:
            im = im.convert('RGB')
            if self.image_size and im.size != (self.image_size, self.image_size):
                im = im.resize((self.image_size, self.image_size), Image.BILINEAR)
            x = torch.frombuffer(bytearray(im.tobytes()), dtype=torch.uint8).view(im.size[1], im.size[0], 3)
        x = x.permute(2, 0, 1).float().div_(255.0)
        mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
        x = (x - mean) / std
        if self.transform is not None:
            x = self.transform(x)
        return x, label
