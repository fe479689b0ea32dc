"""How much of the step is launch gaps?  20 steps (a) as bench.py runs them, (b) with dt given (no host read),
(c) one step captured in a CUDA graph and replayed."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyrmt_b200.driver import make_case, fsi_step
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
state, prm = make_case(N, scheme="weno5")
for _ in range(6):
    state, dt, _ = fsi_step(state, prm)
torch.cuda.synchronize()
def timed(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
box = {"s": state}
def a():
    box["s"] = fsi_step(box["s"], prm)[0]
def b():
    box["s"] = fsi_step(box["s"], prm, dt)[0]
out = {"N": N, "dt": dt, "ms_host_dt": timed(a), "ms_given_dt": timed(b)}
try:
    static = tuple(t.clone() for t in box["s"])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fsi_step(static, prm, dt)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        res = fsi_step(static, prm, dt)[0]
    out["ms_graph"] = timed(g.replay)
    ref = fsi_step(static, prm, dt)[0]
    g.replay(); torch.cuda.synchronize()
    out["graph_equals_eager"] = all(bool(torch.equal(x, y)) for x, y in zip(res, ref))
except Exception as e:
    out["graph_error"] = repr(e)[:600]
print(json.dumps(out))
