"""Probe: does torch's symmetric memory (peer-mapped buffers + device-side barrier over NVLink) work on this box?  torchrun, 1 rank / GPU."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
  n = 1 << 16
  t = symm_mem.empty(n, dtype=torch.float64, device=torch.device("cuda", local))
  hdl = symm_mem.rendezvous(t, dist.group.WORLD)
  t.fill_(float(rank))
  hdl.barrier()
  right = (rank + 1) % world
  peer = hdl.get_buffer(right, (n,), torch.float64)
  mine = torch.full((16,), 100.0 + rank, dtype=torch.float64, device="cuda")
  peer[:16].copy_(mine)                       # store into the neighbour's memory
  hdl.barrier()
  torch.cuda.synchronize()
  left = (rank - 1) % world
  ok = bool((t[:16] == 100.0 + left).all()) and bool((t[16:32] == float(rank)).all())
  # timing: 1000 x (peer copy of 3 x 16 KB + barrier)
  rows = torch.randn(3 * 2048, dtype=torch.float64, device="cuda")
  for _ in range(20):
    peer[:rows.numel()].copy_(rows); hdl.barrier()
  torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
  for _ in range(1000):
    peer[:rows.numel()].copy_(rows); hdl.barrier()
  torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 1000
  # one-shot all-reduce of 84 doubles through peer reads
  s = symm_mem.empty(128, dtype=torch.float64, device=torch.device("cuda", local)); hs = symm_mem.rendezvous(s, dist.group.WORLD)
  s.fill_(rank + 1.0); hs.barrier()
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for _ in range(1000):
    tot = hs.get_buffer(0, (128,), torch.float64).clone()
    for r in range(1, world):
      tot += hs.get_buffer(r, (128,), torch.float64)
    hs.barrier()
  torch.cuda.synchronize(); dt2 = (time.perf_counter() - t0) / 1000
  print("rank %d ok=%s  peer copy 48 KB + barrier: %.1f us   one-shot all-reduce (128 doubles) + barrier: %.1f us  total[0]=%g"
        % (rank, ok, dt * 1e6, dt2 * 1e6, float(tot[0])), flush=True)
except Exception as e:
  import traceback; traceback.print_exc()
  print("rank %d SYMM_MEM_FAILED %r" % (rank, e), flush=True)
dist.barrier()
dist.destroy_process_group()
