"""torchrun --nproc-per-node N tools/c5_train_sync_ab.py [columns] : training step at the large-scale shape (m 1000, d 2000, K 40:
320 MB of weight gradients) with the gradient sync as ONE collective after the backward and as 32 MB buckets overlapped with the
backward (model.sync_gradients(bucket_mb=...)).  Environment switches worth sweeping: NCCL_MAX_CTAS, DLADMM_GRID_CAP."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import dladmm_b200 as dl
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m, d, K = 1000, 2000, 40
data = dl.gen_syn_data(B, m=m, d=d, seed=1, col_offset=rank * B)
z = lambda r: torch.zeros(r, B, device="cuda")
torch.manual_seed(1126)
model = dl.DLADMMNetScalar(m, 1, d, B, data.A, torch.rand(d, B, device="cuda") / d, z(m), z(m), K)
# where the time goes: events around the sync inside the backward (a = every backward kernel enqueued, b = gradients summed)
from dladmm_b200 import function as F
_orig = F._sync_gradients
marks = []
MODE = ["real"]
def _wrapped(spec, flat, plan=None):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if MODE[0] == "real":
        _orig(spec, flat, plan)
    elif MODE[0] == "tiny":                      # a 16-byte collective in place of the 320 MB one
        dist.all_reduce(flat[:4])
    elif MODE[0] == "hostsync":                  # no collective: the host waits for the device once per step
        torch.cuda.synchronize()
    b.record()
    marks.append((a, b))
F._sync_gradients = _wrapped
starts = []
def step():
    model.zero_grad(set_to_none=True)
    s = torch.cuda.Event(enable_timing=True); s.record(); starts.append(s)
    loss, _ = model.l1l1_loss(data.X, 0.001)
    loss.backward()
def timed(n=3):
    step(); torch.cuda.synchronize(); dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n):
        step()
    t1.record(); torch.cuda.synchronize()
    t = torch.tensor([t0.elapsed_time(t1) / n], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
# the chip is power-capped at this shape and drifts for the first second of load: warm up, then interleave the schedules
model.sync_gradients(False)
for _ in range(20):
    step()
res = {}
MODES = (("no sync", False, None), ("single", True, 0), ("bucket 32 MB", True, 32), ("bucket 64 MB", True, 64))
for rep in range(3):
    for label, sync, mb in MODES:
        model.sync_gradients(sync, bucket_mb=mb)
        del marks[:], starts[:]
        t = timed(8)
        comp = sum(s_.elapsed_time(a) for s_, (a, b) in zip(starts[1:], marks[1:])) / (len(marks) - 1)
        tail = sum(a.elapsed_time(b) for a, b in marks[1:]) / (len(marks) - 1)
        res.setdefault(label, []).append((t, comp, tail))
if rank == 0:
    print("N=%d B=%d NCCL_MAX_CTAS=%s DLADMM_GRID_CAP=%s (ms per step: total / fwd+bwd kernels / sync tail, three interleaved rounds of 8 steps): " % (world, B, os.environ.get("NCCL_MAX_CTAS"), os.environ.get("DLADMM_GRID_CAP"))
          + " | ".join("%s %s" % (k, ", ".join("%.2f/%.2f/%.2f" % r for r in v)) for k, v in res.items()))
dist.destroy_process_group()
