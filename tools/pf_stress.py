"""Stress of the all-layer persistent forward kernel: random shapes / variants / modes, each compared bit for bit with the
per-layer schedule; repeated back-to-back calls.  python tools/pf_stress.py [rounds]"""
import os, sys, random, time
os.environ["DLADMM_PERSISTENT"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dladmm_b200 as dl
from dladmm_b200.function import run_forward
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = random.Random(1234)
t0 = time.time()
for it in range(rounds):
    variant = rng.choice(["scalar", "full", "tied", "lasso", "ltheta"])
    m = rng.choice([24, 64, 250, 256, 300]); d = rng.choice([40, 128, 500, 512, 700])
    B = rng.choice([4, 20, 128, 132, 1000, 4096, 18944, 19072, 40000]); K = rng.choice([1, 2, 3, 7])
    prec = rng.choice(["tf32x3", "tf32"])
    torch.manual_seed(it)
    data = dl.gen_syn_data(B, m=m, d=d, seed=it)
    z = lambda r: torch.zeros(r, B, device="cuda")
    model = dl.VARIANT_CLASSES[variant](m, 1, d, B, data.A, torch.rand(d, B, device="cuda") / d, z(m), z(m), K, precision=prec)
    spec, params = model._spec_and_params()
    params = [p.detach() for p in params]
    kw = rng.choice([dict(want_masks=False), dict(want_masks=False, last_only=True), dict(want_masks=True, extras={})])
    obj = rng.choice([None, 0.01])
    if obj is not None and "extras" not in kw:
        kw = dict(kw, extras={})
    outs = []
    for rep in range(3):
        outs.append(run_forward(spec, model.A, data.X, model.Z0, model.E0, model.L0, params, objective_alpha=obj, **kw))
    os.environ["DLADMM_NO_PERSISTENT"] = "1"
    ref = run_forward(spec, model.A, data.X, model.Z0, model.E0, model.L0, params, objective_alpha=obj, **kw)
    os.environ.pop("DLADMM_NO_PERSISTENT")
    def written(o):      # last_only: 2 ping-pong slabs, of which K = 1 only ever writes one (the other is uninitialised memory)
        if not kw.get("last_only"):
            return o
        Z, E, L, T = o[:4]
        return (Z[(K - 1) % 2], E[(K - 1) % 2], L[(K - 1) % 2], T[K % 2]) + tuple(o[4:])
    for o in outs:
        for x, y in zip(written(o), written(ref)):
            assert (x is None) == (y is None)
            if x is not None:
                assert torch.equal(x, y), (it, variant, m, d, B, K, prec, kw.keys())
    torch.cuda.synchronize()
print("pf_stress ok: %d random cases x 3 calls, %.1f s" % (rounds, time.time() - t0))
