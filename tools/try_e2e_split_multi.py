"""e2e ql! n=16384 through host pointers on every rank at once (torchrun): which host schedule copes best when the GPUs share the
host's memory system.  Prints the max over ranks per configuration."""
import os, sys, time, json
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
rank = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n = 16384
q = load_package(); h = q.Handle(rank)
hp = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
src = torch.randn((n, n), dtype=torch.float64)
Ah = hp.numpy().T
flag = torch.zeros(1, device="cuda")
for split, cw in [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or ((2, 0), (1, 0), (0, 0)):
    h.set_option("host_split", split); h.set_option("host_split_chunk", cw)
    ts = []
    for i in range(4):
        hp.copy_(src); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); q.ql_(Ah, handle=h); ts.append(time.perf_counter() - t0)
    t = torch.tensor([min(ts[1:])], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"gpus": world, "host_split": split, "chunk": cw, "rate_rank0": h.upload_rate(), "ms_max": t.item() * 1e3,
                          "tflops_total": world * 4 / 3 * n ** 3 / t.item() / 1e12}), flush=True)
dist.destroy_process_group()
