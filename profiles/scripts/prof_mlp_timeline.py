"""In-kernel timeline (RCB_MLP_PROFILE build) of the MLP kernel in the fit step's own configuration."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RECOMBINER_GRAPH", "0")
import numpy as np, torch
from recombiner_b200 import _lib
_lib.LIB_PATH = os.path.abspath("scratch/libprof.so")
import bench
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
eng = m.engine
ws = eng.workspace(rows, 5)
coef = 2.0 / (5 * eng.pix * eng.out)
for _ in range(2):
    eng.mlp(ws, rows, 5, x, mode=1, y=y, coef=coef)
torch.cuda.synchronize()
lib = _lib.load()
lib.rcb_mlp_prof_read.argtypes = [C.c_void_p]; lib.rcb_mlp_prof_read.restype = C.c_int
buf = np.zeros(4 * 1024, dtype=np.int64)
assert lib.rcb_mlp_prof_read(buf.ctypes.data) == 0
names = {100: "wait", 200: "epi", 300: "pub", 301: "iss"}
for slot in range(2):
    ev = buf[slot * 1024:(slot + 1) * 1024].reshape(512, 2)
    ev = ev[ev[:, 1] != 0]
    t0 = ev[0, 1]
    print(f"--- slot {slot}: {len(ev)} events")
    print(" ".join(f"{int(e[0])}:{int(e[1]-t0)}" for e in ev))
    # per-phase sums over the loop: time from each event to the next
    tot = {}
    for a, b in zip(ev[:-1], ev[1:]):
        tot[int(a[0])] = tot.get(int(a[0]), 0) + int(b[1] - a[1])
    print({names.get(k, k): v for k, v in tot.items()})
