"""Timeline of the all-layer persistent forward kernel (DLADMM_PF_TRACE=1): per unit of the first CTAs, when the MMA warp got its
accumulator, when the first operands landed, when the last MMA was issued, when the epilogue started and ended.

    DLADMM_PF_TRACE=1 python tools/pf_trace.py [B] [precision] > gpurun_out/pf_trace.txt
"""
import ctypes as C, os, sys
os.environ["DLADMM_PF_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, dladmm_b200 as dl
from dladmm_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
variant = sys.argv[3] if len(sys.argv) > 3 else "scalar"
m, d, K = (int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])) if len(sys.argv) > 6 else (250, 500, 15)
data = dl.gen_syn_data(B, m=m, d=d, seed=1)
Z0 = torch.rand(d, B, device="cuda") / d
z = lambda r: torch.zeros(r, B, device="cuda")
torch.manual_seed(1126)
model = dl.VARIANT_CLASSES[variant](m, 1, d, B, data.A, Z0, z(m), z(m), K, precision=prec)
with torch.no_grad():
    for _ in range(3):
        model(data.X)
torch.cuda.synchronize()
n = 4 * 512 * 8
buf = (C.c_int64 * n)()
got = _lib.load().dladmm_debug_trace(buf, n)
assert got == n, got
tr = np.frombuffer(buf, dtype=np.int64).reshape(4, 512, 8)
GHZ = 1.9
names = {0: "T0", 1: "Z", 2: "E"}
for cta in range(2):
    t = tr[cta]
    nun = int((t[:, 3] > 0).sum())
    t0 = t[0, 1]
    print("CTA %d: %d units, kernel span %.1f us" % (cta, nun, (t[nun - 1, 5] - t0) / GHZ / 1e3))
    rows = []
    for i in range(nun):
        ty = int(t[i, 0]); st, bt = int(t[i, 7] >> 32), int(t[i, 7] & 0xffffffff)
        acc_wait = (t[i, 1] - (t[i - 1, 3] if i else t0)) / GHZ / 1e3          # MMA warp idle between units (waiting for a free accumulator)
        op_wait = (t[i, 2] - t[i, 1]) / GHZ / 1e3                               # ... then for the first operands
        mma = (t[i, 3] - t[i, 2]) / GHZ / 1e3
        epi_wait = (t[i, 4] - (t[i - 1, 5] if i else t0)) / GHZ / 1e3           # epilogue idle before this unit
        epi = (t[i, 5] - t[i, 4]) / GHZ / 1e3
        rows.append((ty, acc_wait, op_wait, mma, epi_wait, epi))
        if i < 40:
            print("  unit %3d %-2s stage %2d bt %4d | mma: wait_acc %6.2f wait_op %6.2f issue %6.2f | epi: idle %6.2f run %6.2f" %
                  (i, names[ty], st, bt, acc_wait, op_wait, mma, epi_wait, epi))
    rows = np.array(rows)
    for ty in (1, 2):
        r = rows[rows[:, 0] == ty]
        print("  %s units: n=%d  mean wait_acc %.2f wait_op %.2f mma_issue %.2f | epi idle %.2f run %.2f  (us)" %
              (names[ty], len(r), r[:, 1].mean(), r[:, 2].mean(), r[:, 3].mean(), r[:, 4].mean(), r[:, 5].mean()))
    print("  totals (us): mma busy %.1f, mma wait_acc %.1f, wait_op %.1f | epi busy %.1f idle %.1f" %
          (rows[:, 3].sum(), rows[:, 1].sum(), rows[:, 2].sum(), rows[:, 5].sum(), rows[:, 4].sum()))
