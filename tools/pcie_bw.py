"""Page-locked host <-> device copy bandwidth of the box: the ceiling of bench.py's `e2e`.

    python tools/pcie_bw.py                                              # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_bw.py
                                                                         # N concurrent streams, one process per GPU

Every rank copies 1 GiB device -> pinned host (and back) `REPS` times between barriers; the aggregate
is total bytes / max time over the ranks -- what N GPUs streaming FASTQ to the host can reach at best."""
import json
import os
import time

import torch

REPS = 8


def main():
  rank, local, world = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('LOCAL_RANK', '0'), ('WORLD_SIZE', '1')))
  torch.cuda.set_device(local)
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  nb = 1 << 30
  d = torch.empty(nb, dtype=torch.uint8, device='cuda')
  h = torch.empty(nb, dtype=torch.uint8).pin_memory()
  out = {}
  for name, dst, src in (('d2h', h, d), ('h2d', d, h)):
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    t0 = time.perf_counter()
    for _ in range(REPS):
      dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
      t = torch.tensor([dt], dtype=torch.float64, device='cuda')
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      dt = float(t[0])
    out[name + '_gbs_aggregate'] = world * REPS * nb / dt / 1e9
    out[name + '_gbs_per_gpu'] = REPS * nb / dt / 1e9
  if rank == 0:
    print(json.dumps(dict(out, n_gpus=world, cores=os.cpu_count())))
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
