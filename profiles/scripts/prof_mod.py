import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # the repository root
os.environ["RECOMBINER_GRAPH"] = "0"
import numpy as np, torch, bench
from recombiner_b200.config import configs
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
for name in sys.argv[1:]:
    if name == "prior":
        wl = bench.make_workload(1024, seed=1000)
        print("prior ms", bench.prior_training_ms(wl, dev, steps=2, warmup=1))
        continue
    cfg = configs[name]
    R = int(np.prod(cfg["patch_nums"])) if cfg["patch"] else 1
    rows = bench.MODALITY_DATA[name] * R
    wl = bench.make_workload(rows, seed=77, dataset=name)
    m = bench.build_model(wl, dev)
    x, y = wl["x"][:1].to(dev).expand(rows, -1, -1), wl["y"].to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=2e-4)
    c = m._adam_config(opt)
    for i in range(3):
        m.fit_step(x, y, i + 1, c, 5)
    torch.cuda.synchronize()
    print("done", name, m.engine.x_generated)
