import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # the repository root
os.environ["RECOMBINER_GRAPH"] = "0"
import torch, bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
rows = 1024
wl = bench.make_workload(rows, seed=1000)
m = bench.build_model(wl, dev)
x, y = wl["x"][:1].to(dev).expand(rows, -1, -1), wl["y"].to(dev)
opt = torch.optim.Adam(m.parameters(), lr=2e-4)
cfg = m._adam_config(opt)
for i in range(3):
    m.fit_step(x, y, i + 1, cfg, 5)
torch.cuda.synchronize()
print("done")
