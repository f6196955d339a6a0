"""TEST INFRASTRUCTURE (RDN_SIMT_EMU=1 only): the GPU tests use torch CUDA tensors as plain device memory (``data_ptr()``,
streams, synchronize).  Under the CPU emulation of the kernels "device memory" is host memory, so here every ``device="cuda"``
allocation becomes a CPU tensor, ``.cuda()`` is the identity and streams are dummies — the tests themselves stay unchanged."""
from __future__ import annotations

import contextlib
import os


class _Stream:
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def synchronize(self):
        pass

    def wait_stream(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _is_cuda(dev) -> bool:
    return dev is not None and str(dev).startswith("cuda")


def install() -> None:
    import torch

    def host_device(fn):
        def wrapped(*a, **k):
            if _is_cuda(k.get("device")):
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped

    for name in ("zeros", "ones", "full", "empty", "arange", "tensor", "zeros_like", "empty_like", "full_like", "ones_like", "randint", "rand"):
        setattr(torch, name, host_device(getattr(torch, name)))
    def cuda(self, *a, **k):
        out = torch.empty_like(self)  # torch's CPU allocator aligns to 64 bytes (from_numpy memory need not be 32-byte aligned)
        out.copy_(self)
        return out

    torch.Tensor.cuda = cuda
    orig_to = torch.Tensor.to

    def to(self, *a, **k):
        moved = any(isinstance(x, (str, torch.device)) and _is_cuda(x) for x in a) or _is_cuda(k.get("device"))
        a = tuple("cpu" if isinstance(x, (str, torch.device)) and _is_cuda(x) else x for x in a)
        if _is_cuda(k.get("device")):
            k["device"] = "cpu"
        out = orig_to(self, *a, **k)
        return cuda(out) if moved else out

    torch.Tensor.to = to
    torch.cuda.is_available = lambda: True
    torch.cuda.device_count = lambda: int(os.environ.get("RDN_SIMT_DEVICES", "1"))
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.cuda.Stream = _Stream
    torch.cuda.stream = lambda s=None: contextlib.nullcontext()
    torch.cuda.device = lambda d=None: contextlib.nullcontext()
    torch.cuda.set_device = lambda d: None
