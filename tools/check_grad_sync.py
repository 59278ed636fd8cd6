"""torchrun --nproc-per-node 2 tools/check_grad_sync.py : the in-backward gradient allreduce (model.sync_gradients) gives the
same .grad on every rank as backward + allreduce_gradients, and as one process over the concatenated columns."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import dladmm_b200 as dl
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m, d, K, Bl = 60, 100, 4, 256
def make(B, off):
    data = dl.gen_syn_data(B, m=m, d=d, seed=3, col_offset=off)
    z = lambda r: torch.zeros(r, B, device="cuda")
    torch.manual_seed(5)
    return dl.DLADMMNetFull(m, 1, d, B, data.A, torch.full((d, B), 0.001, device="cuda"), z(m), z(m), K), data
def grads(model, data, total_B, sync, post, bucket_mb=None):
    model.zero_grad(set_to_none=True)
    model.sync_gradients(sync, bucket_mb=bucket_mb)
    loss, _ = model.l1l1_loss(data.X, 0.01, global_batch=total_B)     # mean over the GLOBAL batch (the default when sync is on)
    loss.backward()
    if post:
        dl.allreduce_gradients(list(model.parameters()))
    return [p.grad.clone() for p in model.parameters()]
model, data = make(Bl, rank * Bl)
g_post = grads(model, data, world * Bl, False, True)
big, bigdata = make(world * Bl, 0)
g_one = grads(big, bigdata, world * Bl, False, False)
rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
# one collective after the backward (the default at this size), then the bucketed schedule forced by a 30 KB bucket: the four
# 24 KB weight gradients go as two buckets released by the events of layers 2 and 0, the rest after the last kernel
from dladmm_b200.function import plan_gradient_buckets
for label, bucket_mb in (("single", None), ("bucketed", 0.03)):
    for rep in range(3):                                              # the events and the side stream are reused across steps
        g_sync = grads(model, data, world * Bl, True, False, bucket_mb)
    spec, params = model._spec_and_params()
    plan = plan_gradient_buckets(spec, params, [True] * len(params), spec.grad_bucket_bytes)
    assert (plan is not None and len(plan.buckets) == 2) if bucket_mb else plan is None
    e1 = max(rel(a, b) for a, b in zip(g_sync, g_post))
    e2 = max(rel(a, b) for a, b in zip(g_sync, g_one))
    print("rank %d [%s]: sync vs post-allreduce %.2e, sync vs single process over all columns %.2e" % (rank, label, e1, e2))
    assert e1 < 1e-6 and e2 < 2e-3
dist.destroy_process_group()
