"""Device-buffer plumbing: torch CUDA tensors as the owners of device memory handed to the C ABI."""
import numpy as np

_torch = None


def torch():
  global _torch
  if _torch is None:
    import torch as t
    _torch = t
  return _torch


def require_cuda():
  t = torch()
  if not t.cuda.is_available():
    raise RuntimeError("pdhg_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
  return t


def is_tensor(x):
  return _torch is not None and isinstance(x, _torch.Tensor) or type(x).__module__.startswith("torch")


def current_device():
  """The process's current CUDA device (one process per GPU: `torch.cuda.set_device(local_rank)` selects it)."""
  return int(require_cuda().cuda.current_device())


def to_dev(x, device=None):
  """numpy / torch (any device) -> contiguous fp64 CUDA tensor on `device` (default: the current CUDA device)."""
  t = require_cuda()
  if device is None:
    device = current_device()
  if is_tensor(x):
    return x.to(device="cuda:%d" % device, dtype=t.float64).contiguous()
  return t.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to("cuda:%d" % device)


def like_input(dev_tensor, template):
  """Return the result in the caller's array family (numpy in -> numpy out)."""
  if is_tensor(template):
    return dev_tensor
  return dev_tensor.cpu().numpy()


def stream_ptr(device=None):
  t = torch()
  if device is None:
    device = current_device()
  return int(t.cuda.current_stream(device).cuda_stream)
