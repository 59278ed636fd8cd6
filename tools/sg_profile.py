"""Where the safeguarded evaluation's time goes (scalar K=20, 4096 columns): wall clock, device time, host profile."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
B, K, m, d = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 20, 250, 500
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
z = lambda r: torch.zeros(r, B, device="cuda")
torch.manual_seed(1126)
mdl = dl.DLADMMNetScalar(m, 1, d, B, data.A, torch.rand(d, B, device="cuda") / d, z(m), z(m), K)
kw = dict(delta=0.0, mu_k_method="EMA", mu_k_param=0.5)
for _ in range(3):
    mdl.forward_safeguarded(data.X, True, True, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = mdl.forward_safeguarded(data.X, True, True, **kw)
e1.record(); th = time.perf_counter() - t0
torch.cuda.synchronize(); tw = time.perf_counter() - t0
print("per call: host enqueue+final read %.3f ms, wall %.3f ms, device span %.3f ms" % (th * 100, tw * 100, e0.elapsed_time(e1) / 10))
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    mdl.forward_safeguarded(data.X, True, True, **kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
