"""Experiment: one 65536-column forward vs two 32768-column forwards on two streams with 74-CTA grids (DLADMM_GRID_CAP=74)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
m, d, K = 250, 500, 15
halves = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = 65536 // halves
models, datas, streams = [], [], []
for h in range(halves):
    data = dl.gen_syn_data(B, m=m, d=d, seed=1, col_offset=h * B)
    z = lambda r: torch.zeros(r, B, device="cuda")
    torch.manual_seed(1)
    models.append(dl.DLADMMNetScalar(m, 1, d, B, data.A, torch.rand(d, B, device="cuda") / d, z(m), z(m), K))
    datas.append(data); streams.append(torch.cuda.Stream())
def step():
    outs = []
    for h in range(halves):
        with torch.cuda.stream(streams[h]), torch.no_grad():
            outs.append(models[h](datas[h].X))
    return outs
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in streams: s.wait_event(e0)
n = 10
for _ in range(n): step()
for s in streams: torch.cuda.current_stream().wait_stream(s)
e1.record(); torch.cuda.synchronize()
print("halves=%d grid_cap=%s: %.3f ms per 65536-column forward" % (halves, os.environ.get("DLADMM_GRID_CAP"), e0.elapsed_time(e1) / n))
