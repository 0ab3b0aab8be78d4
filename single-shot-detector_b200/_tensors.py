"""Argument adaptation for the Python mirror: torch CUDA tensors are passed zero-copy (data_ptr);
NumPy arrays / lists are staged to the current CUDA device and results come back as NumPy."""
import numpy as np
import torch

from . import _lib


class Call:
    """Collects inputs for one library call, remembers whether the caller speaks NumPy or torch."""

    def __init__(self, device=None):
        self.numpy_mode = False
        self.device = device
        self._keep = []

    def _pick_device(self, x):
        if self.device is None:
            if isinstance(x, torch.Tensor) and x.is_cuda:
                self.device = x.device
            else:
                if not torch.cuda.is_available():
                    raise _lib.SsdkError('no CUDA device: this package runs only on the GPU (no CPU fallback)')
                self.device = torch.device('cuda', torch.cuda.current_device())
        return self.device

    def tensor(self, x, dtype, shape=None):
        """-> contiguous CUDA tensor of `dtype` (no copy when x already is one)."""
        if isinstance(x, torch.Tensor):
            dev = self._pick_device(x)
            if not x.is_cuda:
                self.numpy_mode = self.numpy_mode  # CPU torch tensors: stage, but keep torch outputs
            t = x.to(device=dev, dtype=dtype).contiguous()
        else:
            self.numpy_mode = True
            dev = self._pick_device(None)
            t = torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(dev)
        if shape is not None:
            t = t.reshape(shape)
        self._keep.append(t)
        return t

    def empty(self, shape, dtype):
        t = torch.empty(shape, dtype=dtype, device=self._pick_device(None))
        self._keep.append(t)
        return t

    def ctx(self):
        dev = self._pick_device(None)
        h = _lib.context(dev.index if dev.index is not None else torch.cuda.current_device())
        _lib.check(_lib.load().ssdk_ctx_set_stream(h, torch.cuda.current_stream(dev).cuda_stream))
        return h

    def result(self, *tensors):
        if self.numpy_mode:
            out = tuple(t.cpu().numpy() for t in tensors)
        else:
            out = tensors
        return out[0] if len(out) == 1 else out


def ptr(t):
    return None if t is None else t.data_ptr()
