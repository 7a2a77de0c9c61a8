"""Short run for ncu: 3 fit steps (every kernel a launch of its own), REC rounds, 3 prior-training steps."""
import sys, os, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # the repository root
os.environ.setdefault("RECOMBINER_GRAPH", "0")
import torch, bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
rows = 1024
wl = bench.make_workload(rows, seed=1000)
m = bench.build_model(wl, dev)
x, y = wl["x"][:1].to(dev).expand(rows, -1, -1), wl["y"].to(dev)
opt = torch.optim.Adam(m.parameters(), lr=2e-4)
cfg = m._adam_config(opt)
m._ensure_rec(65536)
for i in range(4):
    m.fit_step(x, y, i + 1, cfg, 5)
torch.cuda.synchronize()
m.compress_round()
G = wl["G"]
spread = (torch.arange(rows, device=dev, dtype=torch.int32) % G).contiguous()
m.compress_round(blocks=spread, apply=False)
torch.cuda.synchronize()
if "--prior" in sys.argv:
    print("prior ms", bench.prior_training_ms(wl, dev, steps=3, warmup=1))
print("done")
