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
  # (non_blocking: a pinned source — e.g. an array from to_host() or torch's pin_memory() — is copied asynchronously)
  return t.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to("cuda:%d" % device, non_blocking=True)


# Results that go back to NumPy callers are copied device -> PINNED host memory (a pageable destination halves the PCIe rate and
# blocks the stream).  Page-locking hundreds of MB per call would cost more than the copy, so the pinned buffers are pooled; a
# buffer is reused only when the ndarray that was handed out for it (and with it every view of it) has been garbage-collected.
_pinned_pool = {}       # (shape, dtype) -> list of [pinned tensor, weakref to the ndarray handed out (or None)]
_PINNED_POOL_BYTES = 4 << 30


def _pinned_entry(dev_tensor):
  t = torch()
  key = (tuple(dev_tensor.shape), dev_tensor.dtype)
  bucket = _pinned_pool.setdefault(key, [])
  for ent in bucket:
    if ent[1] is None or ent[1]() is None:
      return ent
  total = sum(e[0].numel() * e[0].element_size() for bs in _pinned_pool.values() for e in bs)
  if total + dev_tensor.numel() * dev_tensor.element_size() > _PINNED_POOL_BYTES:
    for bs in _pinned_pool.values():       # drop the idle buffers of every shape before growing further
      bs[:] = [e for e in bs if e[1] is not None and e[1]() is not None]
  ent = [t.empty(dev_tensor.shape, dtype=dev_tensor.dtype, pin_memory=True), None]
  bucket.append(ent)
  return ent


def to_host(dev_tensors):
  """Device tensors -> NumPy arrays backed by pooled pinned memory: asynchronous copies, ONE synchronisation."""
  import weakref
  t = torch()
  ents = []
  for d in dev_tensors:
    ent = _pinned_entry(d)
    ent[1] = lambda: True                  # reserved while this call is in flight
    ent[0].copy_(d, non_blocking=True)
    ents.append(ent)
  if dev_tensors:
    t.cuda.current_stream(dev_tensors[0].device).synchronize()
  outs = []
  for ent in ents:
    arr = ent[0].numpy()
    ent[1] = weakref.ref(arr)
    outs.append(arr)
  return outs


def like_input(dev_tensor, template):
  """Return the result in the caller's array family (numpy in -> numpy out)."""
  if is_tensor(template):
    return dev_tensor
  return to_host([dev_tensor])[0]


def stream_ptr(device=None):
  t = torch()
  if device is None:
    device = current_device()
  return int(t.cuda.current_stream(device).cuda_stream)
