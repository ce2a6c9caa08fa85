"""GPU input pipeline: the reference's per-sample image transform with the pixel work on the device.

Mirrors data/base_dataset.py:17-55 (`get_transform`) and data/single_dataset.py:8-35 (`SingleDataset`).  The reference
decodes with PIL and then, per sample and on `nThreads` CPU workers, runs Scale -> RandomCrop -> RandomHorizontalFlip ->
random 90-degree rotation -> ToTensor -> Normalize(0.5, 0.5); at B200 step rates (200 steps/s at batch 8) that starves the
GPU.  Here an image is decoded (and, for the `resize*` / `scale_width*` modes, resized by the very same PIL call) ONCE, kept as
uint8 HWC in device memory, and every later sample is one kernel launch (`sgk_image_transform_u8`: crop + flip + rotation +
ToTensor + Normalize + channel selection, written straight into the batch buffer).  The random decisions are drawn from
Python's `random` in the order the reference's transform list consumes them (crop top, crop left, flip, rotation), so a
seeded run sees the same augmentations; the arithmetic is bit-identical to ToTensor + Normalize in fp32.
"""
import os
import random

import torch

from . import _lib as L

IMG_EXTENSIONS = ('.jpg', '.jpeg', '.png', '.ppm', '.bmp', '.tif', '.tiff')


def draw_params(opt, H0, W0):
    """(y0, x0, flip, rot) for one sample, consuming Python `random` like get_transform's list does
    (torchvision RandomCrop.get_params: no draw when the image already has the crop size)."""
    S = opt.fineSize
    mode = opt.resize_or_crop
    y0 = x0 = 0
    if mode in ('resize_and_crop', 'crop', 'scale_width_and_crop'):
        if H0 < S or W0 < S:
            raise ValueError("image %dx%d is smaller than fineSize %d" % (H0, W0, S))
        if not (H0 == S and W0 == S):
            y0 = random.randint(0, H0 - S)
            x0 = random.randint(0, W0 - S)
    flip = 0
    if opt.isTrain and not opt.no_flip:
        flip = 1 if random.random() < 0.5 else 0
    rot = 0
    if opt.isTrain and not getattr(opt, 'no_rotate', True):
        rot = random.randint(0, 3)
    return y0, x0, flip, rot


def decode(path, opt):
    """PIL decode + the deterministic, size-changing head of get_transform (Scale / scale_width), as uint8 HWC on the host."""
    import numpy as np
    from PIL import Image
    img = Image.open(path).convert('RGB')
    mode = opt.resize_or_crop
    if mode == 'resize_and_crop':
        img = img.resize((opt.loadSize, opt.loadSize), Image.BILINEAR)
    elif mode in ('scale_width', 'scale_width_and_crop'):
        tw = opt.fineSize if mode == 'scale_width' else opt.loadSize
        ow, oh = img.size
        if ow != tw:
            img = img.resize((tw, int(tw * oh / ow)), Image.BILINEAR)
    return torch.from_numpy(np.asarray(img, dtype=np.uint8).copy())


class GpuTransform(object):
    """transform(img_u8, out) -> out: one decoded uint8 [H, W, C] image (CUDA, or pinned host memory mapped by the driver) to
    fp32 [len(channels), S, S] planes."""

    def __init__(self, opt, channels=(0, 1, 2)):
        self.opt = opt
        self.channels = tuple(int(c) for c in channels)
        if not 1 <= len(self.channels) <= 4:
            raise ValueError("GpuTransform: 1..4 channels")

    def __call__(self, img_u8, out=None, params=None):
        if img_u8.dtype != torch.uint8 or img_u8.dim() != 3 or not img_u8.is_contiguous():
            raise RuntimeError("GpuTransform: expected a contiguous uint8 [H, W, C] image")
        if not img_u8.is_cuda:
            raise RuntimeError("GpuTransform: the image must be in device memory (no CPU fallback); use .cuda() once at load time")
        H0, W0, C0 = img_u8.shape
        S = self.opt.fineSize
        if self.opt.resize_or_crop == 'scale_width':
            raise NotImplementedError("GpuTransform: 'scale_width' yields non-square samples; only the cropping modes and 'none' "
                                      "have a kernel")
        y0, x0, flip, rot = params if params is not None else draw_params(self.opt, H0, W0)
        if out is None:
            out = torch.empty((len(self.channels), S, S), dtype=torch.float32, device=img_u8.device)
        elif out.shape != (len(self.channels), S, S) or out.dtype != torch.float32 or not out.is_contiguous() or not out.is_cuda:
            raise RuntimeError("GpuTransform: `out` must be a contiguous fp32 CUDA tensor [channels, fineSize, fineSize]")
        import ctypes
        chan = (ctypes.c_int * len(self.channels))(*self.channels)
        L.check(L.load().sgk_image_transform_u8(img_u8.data_ptr(), H0, W0, C0, out.data_ptr(), S, y0, x0, flip, rot, chan,
                                                len(self.channels), torch.cuda.current_stream().cuda_stream), "image_transform_u8")
        return out


class GpuSingleDataset(object):
    """SingleDataset (data/single_dataset.py) whose samples are produced on the device.  Images are decoded lazily, once,
    and cached as uint8 in HBM (a 512x512 RGB image is 0.75 MB: 100 000 images fit a B200 twice over)."""

    def initialize(self, opt, device=None):
        self.opt = opt
        self.root = opt.dataroot
        self.dir_A = os.path.join(opt.dataroot, opt.phase)
        paths = []
        for root, _, fnames in sorted(os.walk(self.dir_A)):
            for fname in fnames:
                if fname.lower().endswith(IMG_EXTENSIONS):
                    paths.append(os.path.join(root, fname))
        self.A_paths = sorted(paths)
        self.transform = GpuTransform(opt)
        self.device = torch.device(device if device is not None else "cuda")
        self._cache = {}
        return self

    def _image(self, index):
        img = self._cache.get(index)
        if img is None:
            img = decode(self.A_paths[index], self.opt).to(self.device)
            self._cache[index] = img
        return img

    def __getitem__(self, index):
        return {'A': self.transform(self._image(index)), 'A_paths': self.A_paths[index]}

    def batch(self, indices, out=None):
        """{'A': [B, 3, S, S] device tensor, 'A_paths': [...]} -- every sample written in place, no stacking copy."""
        S = self.opt.fineSize
        if out is None:
            out = torch.empty((len(indices), 3, S, S), dtype=torch.float32, device=self.device)
        for b, i in enumerate(indices):
            self.transform(self._image(i), out=out[b])
        return {'A': out, 'A_paths': [self.A_paths[i] for i in indices]}

    def __len__(self):
        return len(self.A_paths)

    def name(self):
        return 'SingleImageDataset'
