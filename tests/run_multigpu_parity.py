"""torchrun entry of tests/test_gpu_multi.py: the multi-GPU parity check of tests/mgpu_parity.py on WORLD_SIZE GPUs."""
import json
import os
import sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mgpu_parity   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for k, s in ((501, 31), (101, 11)):
        res = mgpu_parity.run(dist, rank, world, local, k=k, s=s)
        ok = ok and res["ok"]
        if rank == 0:
            print("MULTIGPU_PARITY", "OK" if res["ok"] else "FAIL", json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
