"""GPU side of the clip loader: the reference's test-time transform
`Resize(min_size, max_size)` -> `ToTensor()` (mega_core/data/transforms/build.py:75-83, transforms.py:31-67,295-297)
followed by the collate padding (`to_image_list(size_divisible)`, structures/image_list.py:36-66), for frames that are
still the decoder's output: uint8 HWC.

    loader = GpuFrameTransform(min_size=600, max_size=1000, size_divisible=32)
    images = loader(frames_u8_hwc)        # torch.uint8 [n, H, W, 3], host (pinned) or device
    -> ImageList(uint8 [n, 3, Hp, Wp] on the device, image_sizes = [(oh, ow)] * n)

The ImageList goes to DiffusionDet.forward as it is: the model's first kernel evaluates ToTensor and the normalizer on
the uint8 planes (dvid_preprocess_u8).  The resize reproduces Pillow's bytes (dvid_resize_bilinear_u8), so the
detections are those of the reference pipeline on the same decoded frames, while the host never touches a pixel and
3 bytes per source pixel cross PCIe instead of 12 per resized pixel.
"""
import torch

from . import ops
from .structures import ImageList


def get_size(image_size, min_size, max_size):
    """Resize.get_size (transforms.py:38-59) for a single min_size: (w, h) -> (oh, ow)."""
    w, h = image_size
    size = min_size
    if max_size is not None:
        lo, hi = float(min(w, h)), float(max(w, h))
        if hi / lo * size > max_size:
            size = int(round(max_size * lo / hi))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        return (int(size * h / w), size)
    return (size, int(size * w / h))


class GpuFrameTransform:
    def __init__(self, min_size=600, max_size=1000, size_divisible=32, device="cuda"):
        self.min_size = int(min_size[0] if isinstance(min_size, (list, tuple)) else min_size)
        self.max_size = max_size
        self.size_divisible = int(size_divisible)
        self.device = torch.device(device)
        self.h2d_bytes = 0

    def __call__(self, frames):
        if frames.dtype != torch.uint8 or frames.dim() not in (3, 4) or frames.shape[-1] != 3:
            raise ValueError("GpuFrameTransform expects decoded frames: uint8 [n, H, W, 3] or [H, W, 3]")
        if frames.dim() == 3:
            frames = frames[None]
        if not frames.is_cuda:
            self.h2d_bytes += frames.numel()
            frames = frames.to(self.device, non_blocking=True)
        n, h, w, _ = frames.shape
        oh, ow = get_size((w, h), self.min_size, self.max_size)
        planes = ops.resize_frames_u8(frames.contiguous(), oh, ow, pad_to=self.size_divisible)
        return ImageList(planes, [(oh, ow)] * n)

    def from_jpeg(self, files):
        """Compressed frames (an iterable of JPEG byte strings of one video, all the same size) -> the same ImageList:
        nvJPEG decode on the GPU (ops.decode_jpeg), then the resize above.  Only the compressed bytes cross PCIe."""
        frames = [ops.decode_jpeg(f, self.device) for f in files]
        self.h2d_bytes += sum(len(f) for f in files)
        return self(torch.stack(frames))
