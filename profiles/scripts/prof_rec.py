import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # the repository root
import torch, bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
rows = 1024
wl = bench.make_workload(rows, seed=1000)
m = bench.build_model(wl, dev)
m._ensure_rec(65536)
m.compress_round(apply=False)
spread = (torch.arange(rows, device=dev, dtype=torch.int32) % wl["G"]).contiguous()
m.compress_round(blocks=spread, apply=False)
torch.cuda.synchronize()
print("done")
