"""Host-side cost of one forward call (small batch: the GPU is never the limit), split into Python and library time."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
m, d, K, B = 250, 500, 15, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
z = lambda r: torch.zeros(r, B, device="cuda")
model = dl.DLADMMNetScalar(m, 1, d, B, data.A, torch.rand(d, B, device="cuda") / d, z(m), z(m), K)
def run(n):
    with torch.no_grad():
        for _ in range(n):
            model(data.X)
    torch.cuda.synchronize()
run(20)
t0 = time.perf_counter(); run(200); t1 = time.perf_counter()
print("B=%d: %.3f ms per forward call (wall incl. final sync)" % (B, (t1 - t0) / 200 * 1e3))
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.no_grad():
    for _ in range(50):
        model(data.X)
t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print("B=%d: host enqueue %.3f ms per call; GPU drained %.3f ms after the last enqueue" % (B, (t1 - t0) / 50 * 1e3, (t2 - t1) * 1e3))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(50); e1.record(); torch.cuda.synchronize()
print("B=%d: device time %.3f ms per call" % (B, e0.elapsed_time(e1) / 50))
pr = cProfile.Profile(); pr.enable(); run(100); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
