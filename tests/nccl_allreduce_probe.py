"""Development probe (torchrun): NCCL all-reduce time by dtype / op / size on this box."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    ops = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}
    for mb in (32, 290):
        n = mb * 1024 * 1024 // 8
        for dt in (torch.float64, torch.int64, torch.float32, torch.int32):
            t = torch.ones(n * (8 // torch.empty(0, dtype=dt).element_size()), dtype=dt, device="cuda")
            for name, op in ops.items():
                for _ in range(2):
                    dist.all_reduce(t, op=op)
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                for _ in range(5):
                    dist.all_reduce(t, op=op)
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) / 5 * 1e3
                if rank == 0:
                    print(f"{mb:4d} MB {str(dt):14s} {name}: {ms:7.3f} ms  "
                          f"{mb / 1024 / (ms * 1e-3):6.1f} GB/s (alg)", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
