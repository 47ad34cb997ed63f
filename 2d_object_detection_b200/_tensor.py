"""Tensor plumbing: torch owns device memory and streams (DLPack in, torch tensors out)."""
from __future__ import annotations

import torch

from . import _lib


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("2d_object_detection_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_device(x, dtype=torch.float32) -> torch.Tensor:
    """fp32 C-contiguous tensor on the current CUDA device.

    torch CUDA tensors pass through zero-copy; anything exporting __dlpack__ (TF via
    tf.experimental.dlpack, CuPy, JAX) is imported zero-copy; host arrays are uploaded."""
    dev = require_cuda()
    if not isinstance(x, torch.Tensor):
        if hasattr(x, "__dlpack__") and not type(x).__module__.startswith("numpy"):
            x = torch.from_dlpack(x)
        else:
            x = torch.as_tensor(x)
    if x.dtype != dtype:
        x = x.to(dtype)
    if x.device != dev:
        x = x.to(dev, non_blocking=True)
    return x.contiguous()


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def workspace(nbytes: int) -> torch.Tensor:
    """Scratch from torch's caching allocator (stream-ordered, so safe to reuse across calls)."""
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=require_cuda())


lib = _lib.load
check = _lib.check
