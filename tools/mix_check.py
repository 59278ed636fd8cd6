"""Quick check of the split-precision mode tf32_bf16x2 against fp64 and against tf32x3: one product, a short forward on both
schedules, a training step, and C1 timings.   python tools/mix_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dladmm_b200 as dl
import dladmm_oracle as orc

torch.backends.cuda.matmul.allow_tf32 = False
def one_product(m, d, B):
    torch.manual_seed(3)
    A = torch.randn(m, d, device="cuda"); A = A / A.norm(dim=0, keepdim=True)
    Z = torch.randn(d, B, device="cuda")
    z = torch.zeros(m, B, device="cuda")
    want = A.double() @ Z.double()
    for mode in ("fp32", "tf32x3", "tf32_bf16x2", "tf32"):
        model = dl.DLADMMNetScalar(m, 1, d, B, A, Z, z, z, 1, precision=mode)
        got = model._t0(z)
        print("product %dx%d B=%d %-12s rel err %.2e  mean signed rel %.2e" % (m, d, B, mode, ((got.double() - want).norm() / want.norm()).item(),
              (((got.double() - want) * want.sign()).sum() / want.abs().sum()).item()))

def fwd(m, d, K, B, persistent):
    if persistent: os.environ.pop("DLADMM_NO_PERSISTENT", None)
    else: os.environ["DLADMM_NO_PERSISTENT"] = "1"
    data = dl.gen_syn_data(B, m=m, d=d, seed=11)
    Z0 = torch.rand(d, B, device="cuda") / d
    E0 = torch.zeros(m, B, device="cuda"); L0 = torch.zeros(m, B, device="cuda")
    d64 = lambda t: t.detach().double().cpu()
    ref = None
    for mode in ("fp32", "tf32x3", "tf32_bf16x2"):
        torch.manual_seed(7)
        model = dl.DLADMMNetScalar(m, 1, d, B, data.A, Z0, E0, L0, K, precision=mode)
        if ref is None:
            sd = {k: d64(v) for k, v in model.state_dict().items()}
            ref = orc.forward("scalar", sd, d64(data.A), d64(data.X), d64(Z0), d64(E0), d64(L0), K)
        with torch.no_grad():
            out = model(data.X)
        errs = []
        for i in range(3):
            errs.append(["%.1e" % ((out[i][k].double().cpu() - ref[i][k]).norm() / ref[i][k].norm().clamp_min(1e-30)).item() for k in (0, K // 2, K - 1)])
        print("forward m=%d d=%d K=%d B=%d persistent=%d %-12s Z %s E %s L %s" % (m, d, K, B, persistent, mode, errs[0], errs[1], errs[2]))

def grads(m, d, K, B):
    data = dl.gen_syn_data(B, m=m, d=d, seed=5)
    Z0 = torch.zeros(d, B, device="cuda"); E0 = torch.zeros(m, B, device="cuda"); L0 = torch.zeros(m, B, device="cuda")
    res = {}
    for mode in ("fp32", "tf32x3", "tf32_bf16x2"):
        torch.manual_seed(7)
        model = dl.DLADMMNetScalar(m, 1, d, B, data.A, Z0, E0, L0, K, precision=mode)
        loss = model.l1l1_loss(data.X, 0.01)
        loss = loss[0] if isinstance(loss, tuple) else loss
        loss.backward()
        res[mode] = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]).double()
        print("train %-12s loss %.8e" % (mode, loss.item()))
    for mode in ("tf32x3", "tf32_bf16x2"):
        print("grad  %-12s rel err vs fp32 path %.2e" % (mode, ((res[mode] - res["fp32"]).norm() / res["fp32"].norm()).item()))

def timing(mode, B=65536, m=250, d=500, K=15, n=10):
    os.environ.pop("DLADMM_NO_PERSISTENT", None)
    data = dl.gen_syn_data(B, m=m, d=d, seed=1)
    Z0 = torch.zeros(d, B, device="cuda"); E0 = torch.zeros(m, B, device="cuda"); L0 = torch.zeros(m, B, device="cuda")
    model = dl.DLADMMNetScalar(m, 1, d, B, data.A, Z0, E0, L0, K, precision=mode)
    with torch.no_grad():
        for _ in range(3): out = model(data.X)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): out = model(data.X)
        e1.record(); torch.cuda.synchronize()
    f = e0.elapsed_time(e1) / n
    del out
    for _ in range(2):
        model.zero_grad(); l = model.l1l1_loss(data.X, 0.01); l = l[0] if isinstance(l, tuple) else l; l.backward()
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        model.zero_grad(); l = model.l1l1_loss(data.X, 0.01); l = l[0] if isinstance(l, tuple) else l; l.backward()
    e1.record(); torch.cuda.synchronize()
    print("timing %-12s C1 forward %.3f ms   train step %.3f ms" % (mode, f, e0.elapsed_time(e1) / 5))

if __name__ == "__main__":
    one_product(250, 500, 4096); one_product(500, 250, 4096); one_product(1000, 2000, 2048)
    fwd(250, 500, 15, 2048, False); fwd(250, 500, 15, 2048, True); fwd(120, 300, 4, 200, False)
    grads(250, 500, 6, 2048)
    for mode in ("tf32x3", "tf32_bf16x2", "tf32x3", "tf32_bf16x2"):
        timing(mode)
