"""Per-layer relative L2 error of the iterates against the fp64 oracle, per precision mode, at the C1 shape and depth
(m 250, d 500, K 15) and at the C5 depth (m 1000, d 2000, K 40), plus the error of ONE product per mode.  Run on a B200:

    python tools/precision_table.py > profiles/r02_precision_table.md
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dladmm_b200 as dl
import dladmm_oracle as orc


def table(m, d, K, B, modes):
    torch.manual_seed(1126)
    data = dl.gen_syn_data(B, m=m, d=d, seed=1126)
    Z0 = torch.rand(d, B, device="cuda") / d
    E0 = torch.zeros(m, B, device="cuda"); L0 = torch.zeros(m, B, device="cuda")
    d64 = lambda t: t.detach().double().cpu()
    ref = None
    rows = {}
    for mode in modes:
        torch.manual_seed(7)
        model = dl.DLADMMNetScalar(m, 1, d, B, data.A, Z0, E0, L0, K, precision=mode)
        if ref is None:
            sd = {k: d64(v) for k, v in model.state_dict().items()}
            ref = orc.forward("scalar", sd, d64(data.A), d64(data.X), d64(Z0), d64(E0), d64(L0), K)
            # the reference's own fp32 arithmetic (ATen on the GPU, TF32 off) as the floor
            torch.backends.cuda.matmul.allow_tf32 = False
            sd32 = {k: v.detach() for k, v in model.state_dict().items()}
            with torch.no_grad():
                e32 = orc.forward("scalar", sd32, data.A, data.X, Z0, E0, L0, K)
            rows["aten fp32 (eager port)"] = [[((e32[i][k].double().cpu() - ref[i][k]).norm() / ref[i][k].norm().clamp_min(1e-30)).item()
                                              for k in range(K)] for i in range(4)]
        with torch.no_grad():
            out = model(data.X)
        if len(out) == 3:
            out = list(out) + [None]
        rows[mode] = [[((out[i][k].double().cpu() - ref[i][k]).norm() / ref[i][k].norm().clamp_min(1e-30)).item() for k in range(K)]
                      for i in range(3)] + [[((out[3][k + 1].double().cpu() - ref[3][k + 1]).norm() / ref[3][k + 1].norm().clamp_min(1e-30)).item()
                                             for k in range(K)]]
    print("\n## m=%d d=%d K=%d B=%d, scalar variant, default init: relative L2 error vs the fp64 oracle\n" % (m, d, K, B))
    ks = sorted(set([0, 1, 2, 4, K // 2, K - 2, K - 1]))
    print("| mode | iterate | " + " | ".join("k=%d" % k for k in ks) + " | max |")
    print("|---|---|" + "---|" * (len(ks) + 1))
    for mode, r in rows.items():
        for i, nm in enumerate(("Z", "E", "L", "T")):
            print("| %s | %s | " % (mode, nm) + " | ".join("%.1e" % r[i][k] for k in ks) + " | %.1e |" % max(r[i]))


def one_product(m, d, B):
    """error of ONE A.Z product (T_0 = A Z0 + E0 - X with E0 = X = 0) per mode on dense data"""
    torch.manual_seed(3)
    A = torch.randn(m, d, device="cuda"); A = A / A.norm(dim=0, keepdim=True)
    Z = torch.randn(d, B, device="cuda")
    z = torch.zeros(m, B, device="cuda")
    want = A.double() @ Z.double()
    print("\n## one product (%d x %d)(%d x %d), dense Gaussian operands: relative L2 error vs fp64\n" % (m, d, d, B))
    print("| mode | rel. L2 error | mean signed error (toward zero < 0) |\n|---|---|---|")
    for mode in ("fp32", "tf32_bf16x2", "tf32x3", "tf32", "bf16"):
        model = dl.DLADMMNetScalar(m, 1, d, B, A, Z, z, z, 1, precision=mode)
        got = model._t0(z)
        print("| %s | %.2e | %.2e |" % (mode, ((got.double() - want).norm() / want.norm()).item(), (((got.double() - want) * want.sign()).sum() / want.abs().sum()).item()))
    torch.backends.cuda.matmul.allow_tf32 = False
    got = A @ Z
    print("| aten fp32 (cuBLAS SGEMM) | %.2e | %.2e |" % (((got.double() - want).norm() / want.norm()).item(), (((got.double() - want) * want.sign()).sum() / want.abs().sum()).item()))


if __name__ == "__main__":
    print("# r02: precision modes against the fp64 oracle (tools/precision_table.py on one B200)")
    one_product(250, 500, 4096)
    one_product(500, 250, 4096)
    one_product(1000, 2000, 4096)
    table(250, 500, 15, 4096, ("fp32", "tf32_bf16x2", "tf32x3", "tf32", "bf16"))
    table(1000, 2000, 40, 1024, ("tf32_bf16x2", "tf32x3", "tf32", "bf16"))
    one_product(24, 40, 4096)
    one_product(64, 128, 4096)
    one_product(128, 256, 4096)
