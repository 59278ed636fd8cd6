#!/usr/bin/env python
"""bench.py -- D-LADMM hot-path benchmark (driver contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision P]

Workload (BASELINE.json metric "D-LADMM fwd instances/sec (K=15, batch 64K)"; SURVEY 8(a) "C1"):
    scalar variant (main_syn_l1l1_scalar.py), m=250, d=500, K=15 layers, B=65536 problem instances per GPU,
    synthetic data with gen_syn_data.py semantics (p=0.1, sigma=1), reference default weight init,
    forward returning EVERY layer's iterates (the reference API, a10).
A "step" is one K-layer forward over one batch.  value = instances/s with inputs resident in HBM;
e2e = the same through DLADMMNet.forward with the observations in pinned HOST memory (H2D copy inside the
timed region) plus a D2H read of the per-layer objective.  Weak scaling: every rank owns its own 65536
columns, no collective on the data path.

--impl reference: the reference's algorithm on the host CPU (the oracle port -- the reference is Python
and cannot travel to the GPU box), all host threads, each step a bounded column sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

M, D, K_LAYERS, B_PER_GPU = 250, 500, 15, 65536
VARIANT = "scalar"
METRIC = "dladmm_fwd_instances_per_sec"
UNIT = "instances/s"
CPU_SAMPLE_COLS, CPU_SAMPLE_REPS = 32768, 10      # bounded sample of the workload for the host-CPU legs (~10 s on 16 cores)
F_FWD = 2.0 * M * D * (2 * K_LAYERS + 1)          # algorithmic flops per instance
F_GEMM_PER_COL = 2.0 * M * D                      # one (d x m)(m x 1) or (m x d)(d x 1) product


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed `ncu --set full` capture of the final build
# (profiles/r01_ncu_full_summary_v7.md; a training-mode launch: it also writes the 1-byte prox masks, 16.4 MB for `gemm_elt`
# and 32.8 MB for `gemm_z`, which the inference launch timed here does not), keyed by (precision, kernel kind)
TRAFFIC_NCU = {("tf32x3", "gemm_elt"): 332.131072e6 + 228.349184e6, ("tf32x3", "gemm_z"): 197.948416e6 + 119.981056e6}
try:      # round 2: the persistent kernel's capture (profiles/r02_ncu_persistent.json, written by tools/ncu_summary.py)
    with open(os.path.join(ROOT, "profiles", "r02_ncu_persistent.json")) as _fh:
        _j = json.load(_fh)
    TRAFFIC_NCU[("tf32x3", "fwd_persistent")] = _j["dram__bytes_read.sum"] + _j["dram__bytes_write.sum"]
except Exception:
    pass
try:      # ... and of the default split-precision mode (profiles/r02_ncu_persistent_mixed.json)
    with open(os.path.join(ROOT, "profiles", "r02_ncu_persistent_mixed.json")) as _fh:
        _j = json.load(_fh)
    TRAFFIC_NCU[("tf32_bf16x2", "fwd_persistent")] = _j["dram__bytes_read.sum"] + _j["dram__bytes_write.sum"]
except Exception:
    pass


def _peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            j = json.load(fh)
        p.update(hbm_gbs=j["hbm_gbs"], bf16_tflops=j["bf16_tflops"], bf16_tflops_sustained=j.get("bf16_tflops_sustained"),
                 source="measured")
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "measured_tf32_peak.json")) as fh:
            p["tf32_tflops"] = json.load(fh)["tf32_tflops"]
            p["tf32_source"] = "measured (profiles/measured_tf32_peak.json)"
    except Exception:
        p["tf32_tflops"] = p["bf16_tflops"] / 2.0
        p["tf32_source"] = "bf16 peak / 2 (TF32 not measured)"
    return p


class ClockSampler(object):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                mx = float(f[2])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on host cores (oracle port of the reference's ATen op sequence)
# ---------------------------------------------------------------------------------------------------
def cpu_forward_rate(sample_cols, reps, seed=1126):
    """-> (seconds per timed forward, threads, kind): the reference's DLADMMNet (or the oracle port, see _RefModel) on the host cores"""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, D, generator=g)
    A = A / A.pow(2).sum(dim=0, keepdim=True).sqrt()
    Zs = (torch.rand(D, sample_cols, generator=g) < 0.1).float() * torch.randn(D, sample_cols, generator=g)
    Es = (torch.rand(M, sample_cols, generator=g) < 0.1).float() * torch.randn(M, sample_cols, generator=g)
    X = A.mm(Zs) + Es
    Z0 = torch.rand(D, sample_cols, generator=g) / D
    E0 = torch.zeros(M, sample_cols); L0 = torch.zeros(M, sample_cols)
    rm = _RefModel(VARIANT, M, D, K_LAYERS, sample_cols, A, Z0, E0, L0, "cpu", seed)
    times = []
    with torch.no_grad():
        rm.forward(X)                                                      # warm-up
        for _ in range(reps):
            t0 = time.perf_counter()
            rm.forward(X)
            times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads(), rm.kind


def run_reference(args):
    rank, world, _ = _dist_env()
    if rank != 0:
        return 0
    sample = CPU_SAMPLE_COLS
    times, threads, kind = cpu_forward_rate(sample, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    value = sample * len(timed) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C1 scalar D-LADMM forward, m=250 d=500 K=15, all iterates returned",
                   "columns_per_step": sample, "variant": VARIANT},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "%d-column forward per step (full workload is %d columns per GPU)" % (sample, B_PER_GPU)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def _reference_kind():
    """"reference" when the literal reference classes can be loaded here (/root/reference in the build container, the files
    oracle/fetch_ref.py staged under oracle/_ref/ on the GPU box), else "port" (oracle/dladmm_oracle.py, the same ATen ops)."""
    try:
        import load_reference as lr
        return "reference" if lr.reference_available() else "port"
    except Exception:
        return "port"


class _RefModel(object):
    """The reference's DLADMMNet for `variant` on `device` ("cpu": .cuda() patched to identity; "cuda": unmodified, stock
    PyTorch eager on the B200), or the oracle port of the same op sequence when the class cannot be loaded."""

    def __init__(self, variant, m, d, K, B, A, Z0, E0, L0, device, seed=1126):
        import torch
        import dladmm_oracle as orc
        self.variant, self.K, self.kind, self.device = variant, K, _reference_kind(), device
        self.A = A.to(device)
        torch.manual_seed(seed)
        if self.kind == "reference":
            import load_reference as lr
            cls = lr.load_class(variant)
            args = dict(m=m, n=10000, d=d, batch_size=B, A=A.cpu(), Z0=Z0.cpu(), E0=E0.cpu(), L0=L0.cpu(), layers=K)
            if device == "cpu":
                with lr.cuda_is_identity():
                    self.model = cls(**args)
            else:
                self.model = cls(**args).cuda()
            self.lr = lr
        else:
            self.sd = {k: v.to(device).requires_grad_(True) for k, v in orc.default_state_dict(variant, A.cpu(), K, B).items()}
            self.state = [t.to(device) for t in (Z0, E0, L0)]
            self.orc = orc

    def forward(self, x):
        if self.kind == "reference":
            if self.device == "cpu":
                with self.lr.cuda_is_identity():
                    return self.model(x)
            return self.model(x)
        return self.orc.forward(self.variant, self.sd, self.A, x, self.state[0], self.state[1], self.state[2], self.K)

    def params(self):
        return list(self.model.parameters()) if self.kind == "reference" else list(self.sd.values())

    def train_step(self, x, alpha=0.001):
        """forward + the script's training loss + backward (main_syn_l1l1_scalar.py:283-301 / main_syn_lasso_scalar.py:268-285)"""
        import torch
        for p in self.params():
            p.grad = None
        out = self.forward(x)
        Z = out[0]
        total = 0.0
        for k in range(self.K):
            res = x - torch.mm(self.A, Z[k])
            r = 0.5 * torch.sum(res ** 2.0, dim=0).mean() if self.variant == "lasso" else torch.sum(torch.abs(res), dim=0).mean()
            total = total + (alpha * torch.sum(torch.abs(Z[k]), dim=0).mean() + r) * (0.6 ** 3 if k < self.K - 1 else 1.0)
        total.backward()
        return total


def eager_b200_leg(dev, A, X, budget_s=25.0):
    """The reference's own classes under stock PyTorch eager ON THE B200 ("the real bar to beat"): forward under
    no_grad and forward + script loss + backward, wall clock per call with a synchronize on both sides, median of 3."""
    import torch
    out = {"kind": _reference_kind(), "what": "unmodified reference DLADMMNet, PyTorch %s eager on the same GPU, fp32 (TF32 off), "
                                              "wall clock per call incl. host launch overhead" % torch.__version__, "cases": []}
    torch.backends.cuda.matmul.allow_tf32 = False
    t_leg = time.time()
    cases = [("scalar", 15, 25, True), ("scalar", 15, 4096, True), ("scalar", 15, 65536, True), ("full", 20, 20, True),
             ("full", 20, 4096, True), ("lasso", 15, 20, True), ("lasso", 15, 4096, True)]
    for variant, K, B, with_bwd in cases:
        if time.time() - t_leg > budget_s:
            break
        g = torch.Generator().manual_seed(5)
        Z0 = torch.rand(D, B, generator=g) / D
        E0 = torch.zeros(M, B); L0 = torch.zeros(M, B)
        rm = _RefModel(variant, M, D, K, B, A.cpu(), Z0, E0, L0, "cuda")
        x = X[:, :B].contiguous()

        def timed(fn, n=3):
            ts = []
            for i in range(n + 1):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize(dev)
                ts.append(time.perf_counter() - t0)
            ts = sorted(ts[1:])
            return ts[len(ts) // 2]

        def fwd():
            with torch.no_grad():
                return rm.forward(x)
        row = {"variant": variant, "K": K, "columns": B, "fwd_ms": 1e3 * timed(fwd)}
        row["fwd_instances_per_s"] = B / (row["fwd_ms"] * 1e-3)
        if with_bwd and B <= 65536:
            row["train_ms"] = 1e3 * timed(lambda: rm.train_step(x))
            row["train_samples_per_s"] = B / (row["train_ms"] * 1e-3)
        out["cases"].append(row)
        del rm
        torch.cuda.empty_cache()
    return out


def small_batch_leg(dl, dev, A, X, precision):
    """The reference scripts' own regime (bs 20-25) and up: wall clock of one K=15 forward call (host enqueue + device),
    median of 20 after warm-up."""
    import torch
    rows = []
    for B in (25, 256, 1024, 4096):
        z = lambda r: torch.zeros(r, B, device=dev)
        mdl = dl.VARIANT_CLASSES[VARIANT](M, 10000, D, B, A, torch.rand(D, B, device=dev) / D, z(M), z(M), K_LAYERS,
                                          precision=precision, device=dev)
        x = X[:, :B].contiguous()
        ts = []
        with torch.no_grad():
            for i in range(25):
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                mdl(x)
                torch.cuda.synchronize(dev)
                ts.append(time.perf_counter() - t0)
        ts = sorted(ts[5:])
        # ... and the same call in a loop that does not read its results every iteration (calls pipeline: max(host enqueue, device))
        with torch.no_grad():
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for i in range(100):
                mdl(x)
            torch.cuda.synchronize(dev)
            loop = (time.perf_counter() - t0) / 100
        rows.append({"columns": B, "fwd_ms": 1e3 * ts[len(ts) // 2], "fwd_instances_per_s": B / ts[len(ts) // 2],
                     "fwd_ms_back_to_back": 1e3 * loop})
    return {"what": "one DLADMMNet.forward(x) call, scalar K=15 m=250 d=500: fwd_ms = wall clock of ONE call bracketed by device "
                    "synchronisation (host enqueue + device time), median of 20; fwd_ms_back_to_back = per call in a loop of 100 "
                    "calls synchronised once at the end", "cases": rows}


def safeguard_leg(dl, dev, A, X, precision, B=4096, K=20):
    """Safeguarded evaluation (test_syn_l1l1_scalar.py:179-317): classical + learned + safeguard step per layer, EMA updater."""
    import torch
    z = lambda r: torch.zeros(r, B, device=dev)
    torch.manual_seed(3)
    mdl = dl.VARIANT_CLASSES["scalar"](M, 10000, D, B, A, z(D), z(M), z(M), K, precision=precision, device=dev)
    x = X[:, :B].contiguous()
    kw = dict(delta=0.0, mu_k_method="EMA", mu_k_param=0.5, alpha=0.01)
    mdl.forward_safeguarded(x, True, True, **kw)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        out = mdl.forward_safeguarded(x, True, True, **kw)
        torch.cuda.synchronize(dev)
        ts.append(time.perf_counter() - t0)
    t = sorted(ts)[1]
    return {"what": "DLADMMNet.forward_safeguarded(x, use_learned, use_safeguard), scalar K=%d, %d columns, EMA mu updater" % (K, B),
            "ms": 1e3 * t, "value": B / t, "unit": UNIT, "fallback_columns_per_layer": out[4]}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dladmm_b200 as dl
    from dladmm_b200 import _lib

    rank, world, local = _dist_env()
    # keep NCCL's banner off stdout (rank 0 prints exactly one JSON line): it is printed at NCCL_DEBUG=VERSION and above
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG", None)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    precision = dl.resolve_precision(args.precision or dl.default_precision(), M, D)     # ("auto" resolves by shape)
    B = args.columns
    # this rank's column shard of the global data set (weak scaling: B columns per rank)
    data = dl.gen_syn_data(B, m=M, d=D, p=0.1, sigma=1.0, seed=1126, device=dev, col_offset=rank * B)
    A_host = data.A
    if world > 1:   # A is generated from the seed alone, identical on every rank; assert it
        chk = data.A.sum().reshape(1).clone()
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        assert all(torch.equal(lst[0], t) for t in lst)
    g = torch.Generator(device=dev).manual_seed(1126 + rank)
    Z0 = torch.rand(D, B, device=dev, generator=g) / D                # main_syn_l1l1_scalar.py:226
    E0 = torch.zeros(M, B, device=dev); L0 = torch.zeros(M, B, device=dev)
    torch.manual_seed(1126)
    model = dl.VARIANT_CLASSES[VARIANT](M, 10000, D, B, A_host, Z0, E0, L0, K_LAYERS, precision=precision, device=dev)
    X = data.X

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        with torch.no_grad():
            return model(X)

    # ---- kernel-only: inputs resident in HBM -------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    barrier()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step_resident()          # outputs are dropped each step exactly as in the warm-up (allocator steady state)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = _lib.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    if rank == 0 and ms_total < 400.0:
        # the timed region (steps x ~3 ms) is shorter than a few nvidia-smi samples: keep the SAME step running, untimed, for
        # ~0.4 s more so that the clock / throttle record is taken under this load (the timing above is not affected)
        t_probe = time.time()
        while time.time() - t_probe < 0.4:
            for _ in range(10):
                step_resident()
            torch.cuda.synchronize(dev)
        t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "the %d timed steps plus ~0.4 s of the same step, untimed" % args.steps
    barrier()

    # ---- per-kernel pass (CUDA events around every launch, same steps) --------------------------------
    _lib.profile_start()
    for _ in range(args.steps):
        step_resident()
    prof = _lib.profile_stop()

    # ---- end to end: pinned host observations -> H2D -> forward -> objective -> D2H -------------------
    X_host = X.cpu().pin_memory()
    obj_host = [torch.empty(K_LAYERS, dtype=torch.float32).pin_memory() for _ in range(2)]
    obj_total = torch.zeros(K_LAYERS, dtype=torch.float64)

    def run_e2e(nsteps):
        # dl.HostFeed uploads batch i+1 on a copy stream while batch i computes; every step still copies its own
        # 65.5 MB of observations from pinned host memory.  Every step's per-layer objective is copied to pinned host memory
        # and READ on the host (accumulated, as the reference's test loop does, main_syn_l1l1_scalar.py:333-334): the read of
        # step i-1 happens while step i runs, the last one after the loop -- all inside the timed region.
        feed = dl.HostFeed((X_host for _ in range(nsteps)), dev)
        done = [None, None]
        obj_total.zero_()
        for i, x_dev in enumerate(feed):
            obj, _outs = model.forward_objective(x_dev, 0.001)      # all K iterates returned + fused objective
            obj_host[i & 1].copy_(obj, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            done[i & 1] = ev
            if i > 0:
                done[(i - 1) & 1].synchronize()
                obj_total.add_(obj_host[(i - 1) & 1].double())
        if nsteps > 0:
            done[(nsteps - 1) & 1].synchronize()
            obj_total.add_(obj_host[(nsteps - 1) & 1].double())
        return feed.bytes_copied

    run_e2e(max(2, args.warmup // 2))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h2d_bytes = run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    # ---- training throughput (forward + fused loss + backward + parameter gradients) ---------------------------------
    #   scalar K=15 (C1/C3 shape), full K=20 (BASELINE configs[1]), lasso K=15 (configs[2])
    def train_leg(variant, Kt, Bt, nst):
        tm = dl.VARIANT_CLASSES[variant](M, 10000, D, Bt, A_host, Z0[:, :Bt].contiguous(), E0[:, :Bt].contiguous(),
                                         L0[:, :Bt].contiguous(), Kt, precision=precision, device=dev)
        Xt = X[:, :Bt].contiguous()
        if world > 1:
            tm.sync_gradients(True)
        wts = [0.6 ** 3] * (Kt - 1) + [1.0]
        loss_fn = tm.lasso_loss if variant == "lasso" else tm.l1l1_loss     # main_syn_lasso_scalar.py:276-281 / main_syn_l1l1_scalar.py:289-299

        def step_train():
            tm.zero_grad(set_to_none=True)
            loss, _ = loss_fn(Xt, 0.001, wts)
            loss.backward()                              # world > 1: the backward allreduces the gradient buffer in place
        for _ in range(2):
            step_train()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(nst):
            step_train()
        t1.record()
        barrier()
        ms = t0.elapsed_time(t1) / nst
        _lib.profile_start()
        step_train()
        tprof = _lib.profile_stop()
        if world > 1:     # every rank must hold the same (globally summed) gradients after the in-backward allreduce
            chk = torch.stack([p.grad.double().sum() for p in tm.parameters()])
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), "parameter gradients differ between ranks after sync_gradients"
        return {"variant": variant, "K": Kt, "columns_per_gpu": Bt, "ms_per_step": ms,
                "kernel_ms": {k: v[0] for k, v in tprof.items() if v[1] > 0},
                "kernel_launches": {k: v[1] for k, v in tprof.items() if v[1] > 0}}

    trains = []
    if args.train_columns > 0:
        nst = max(2, args.steps // 2)
        trains.append(train_leg(VARIANT, K_LAYERS, args.train_columns, nst))
        if not args.quick:
            trains.append(train_leg("full", 20, args.train_columns, 2))
            trains.append(train_leg("lasso", 15, args.train_columns, 2))
        torch.cuda.empty_cache()
    train = trains[0] if trains else None

    # ---- other precision modes and API shapes of the C1 forward (3 steps each) ----------------------------------------
    variants_ms = {}
    if not args.quick:
        other = "tf32x3" if precision != "tf32x3" else "tf32_bf16x2"
        for label, prec, kw in ((other, other, {}), ("tf32", "tf32", {}), ("bf16", "bf16", {}), (precision + "_last_only", precision, {"last_only": True})):
            mm = dl.VARIANT_CLASSES[VARIANT](M, 10000, D, B, A_host, Z0, E0, L0, K_LAYERS, precision=prec, device=dev)
            mm.load_state_dict(model.state_dict())

            def stepv():
                with torch.no_grad():
                    return mm(X, **kw)
            for _ in range(2):
                stepv()
            barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(3):
                stepv()
            t1.record()
            barrier()
            variants_ms[label] = t0.elapsed_time(t1) / 3
            del mm
        torch.cuda.empty_cache()

    # ---- end to end with the FINAL iterate brought back to the host (what an inference caller wants) -------------------
    e2e_z = None
    if not args.quick:
        def run_e2e_z(nsteps):
            # upload of batch i+1 (HostFeed) and download of result i-1 (HostDrain) both overlap the forward of batch i
            feed = dl.HostFeed((X_host for _ in range(nsteps)), dev)
            acc = 0.0
            for x_dev in feed:
                with torch.no_grad():
                    Zl = model(x_dev, last_only=True)[0]
                hz = drain.push(Zl[0])
                if hz is not None:
                    acc += float(hz[0, 0])            # the host reads the result
                    drain.recycle(hz)
            for hz in drain.flush():
                acc += float(hz[0, 0])
                drain.recycle(hz)
            return feed.bytes_copied
        # one drain for the warm-up and the timed loop: its three pinned 131 MB result buffers are allocated (page-locked) during
        # the warm-up and recycled afterwards -- a caller streaming batches keeps its drain the same way
        drain = dl.HostDrain(dev)
        run_e2e_z(4)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hb = run_e2e_z(args.steps)
        e1.record()
        barrier()
        e2e_z = {"ms_per_step": e0.elapsed_time(e1) / args.steps, "h2d_bytes_per_step": int(hb // args.steps),
                 "d2h_bytes_per_step": int(4 * D * B)}

    # ---- optional: the large-scale shape (BASELINE configs[4] "C5": A 1000x2000, K=40), tensor-bound regime ---------
    c5 = None
    if args.c5_columns > 0:
        torch.cuda.empty_cache()
        m5, d5, K5, B5 = 1000, 2000, 40, args.c5_columns
        data5 = dl.gen_syn_data(B5, m=m5, d=d5, p=0.1, sigma=1.0, seed=1126, device=dev, col_offset=rank * B5)
        z5 = lambda r: torch.zeros(r, B5, device=dev)
        torch.manual_seed(1126)
        model5 = dl.VARIANT_CLASSES[VARIANT](m5, 10000, d5, B5, data5.A, torch.rand(d5, B5, device=dev) / d5, z5(m5), z5(m5), K5,
                                             precision=precision, device=dev)
        c5 = {"columns_per_gpu": B5}
        for label, last_only in (("all_iterates", False), ("last_only", True)):
            def step5():
                with torch.no_grad():
                    return model5(data5.X, last_only=last_only)
            for _ in range(2):
                step5()
            barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n5 = 3
            t0.record()
            for _ in range(n5):
                step5()
            t1.record()
            barrier()
            c5[label] = t0.elapsed_time(t1) / n5
        # BASELINE configs[4] at its stated size: 1 048 576 instances in total, sharded over the ranks (STRONG scaling), each
        # rank walking its share in chunks of up to 131 072 columns; final iterate + fused per-layer objective (last_only)
        if args.c5_total > 0:
            share = args.c5_total // world
            chunk = min(131072, share)
            nchunk = (share + chunk - 1) // chunk
            datac = dl.gen_syn_data(chunk, m=m5, d=d5, p=0.1, sigma=1.0, seed=1126, device=dev, col_offset=rank * share)
            zc = lambda r: torch.zeros(r, chunk, device=dev)
            torch.manual_seed(1126)
            modelc = dl.VARIANT_CLASSES[VARIANT](m5, 10000, d5, chunk, data5.A, torch.rand(d5, chunk, device=dev) / d5, zc(m5), zc(m5), K5,
                                                 precision=precision, device=dev)
            modelc.load_state_dict(model5.state_dict())
            objs = torch.zeros(K5, device=dev, dtype=torch.float64)

            def pass_1m():
                for _ in range(nchunk):       # (synthetic: every chunk of a rank reuses the same generated columns)
                    obj, _o = modelc.forward_objective(datac.X, 0.001, last_only=True)
                    objs.add_(obj.double())
            pass_1m() if nchunk <= 2 else modelc.forward_objective(datac.X, 0.001, last_only=True)
            barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            pass_1m()
            t1.record()
            barrier()
            c5["total_1m_ms"] = t0.elapsed_time(t1)
            c5["total_1m_chunks"] = nchunk
            c5["total_1m_chunk_columns"] = chunk
            del modelc, datac
        # training at the C5 shape (320 MB of weight gradients per step: the case where the allreduce matters)
        if args.c5_train_columns > 0:
            Bt5 = args.c5_train_columns
            torch.cuda.empty_cache()
            zt = lambda r: torch.zeros(r, Bt5, device=dev)
            tm5 = dl.VARIANT_CLASSES[VARIANT](m5, 10000, d5, Bt5, data5.A, torch.rand(d5, Bt5, device=dev) / d5, zt(m5), zt(m5), K5,
                                              precision=precision, device=dev)
            if world > 1:
                tm5.sync_gradients(True)
            Xt5 = data5.X[:, :Bt5].contiguous()

            def step5t():
                tm5.zero_grad(set_to_none=True)
                loss, _ = tm5.l1l1_loss(Xt5, 0.001)
                loss.backward()

            def time5t():
                step5t()
                step5t()
                barrier()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(4):
                    step5t()
                t1.record()
                barrier()
                return t0.elapsed_time(t1) / 4
            if world > 1:
                # the same step with the opt-in schedule: weight gradients reduced in 32 MB buckets from a side stream while the
                # layers below run; then the default, ONE collective after the backward; same sums asserted
                tm5.sync_gradients(True, bucket_mb=32)
                c5["train_bucketed_ms"] = time5t()
                g_bucketed = [p_.grad.clone() for p_ in tm5.parameters()]
                tm5.sync_gradients(True)
            c5["train_ms"] = time5t()
            if world > 1:
                worst = max(((a - b.grad).norm() / a.norm().clamp_min(1e-20)).item() for a, b in zip(g_bucketed, tm5.parameters()))
                assert worst < 1e-5, "bucketed and single-collective gradient sync disagree: %g" % worst
                del g_bucketed
            c5["train_columns"] = Bt5
            del tm5, Xt5
        del model5, data5
        torch.cuda.empty_cache()

    # ---- max over ranks -------------------------------------------------------------------------------
    extra_keys = [("train%d" % i, t["ms_per_step"]) for i, t in enumerate(trains)] + sorted(variants_ms.items())
    if e2e_z:
        extra_keys.append(("e2e_z", e2e_z["ms_per_step"]))
    if c5:
        extra_keys += [(k, c5[k]) for k in ("total_1m_ms", "train_ms", "train_bucketed_ms") if k in c5]
    stats = torch.tensor([ms_total, ms_e2e, train["ms_per_step"] if train else 0.0,
                          c5["all_iterates"] if c5 else 0.0, c5["last_only"] if c5 else 0.0] + [v for _, v in extra_keys],
                         device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    vals = [float(v) for v in stats.tolist()]
    ms_total, ms_e2e, ms_train_step, ms_c5_all, ms_c5_last = vals[:5]
    extra = {k: v for (k, _), v in zip(extra_keys, vals[5:])}

    if rank == 0:
        peaks = _peaks()
        ms_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total * 1e-3)
        e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
        # dominant kernel = the kind with the most device time in the per-kernel pass
        kinds = {k: v for k, v in prof.items() if v[1] > 0}
        dom = max(kinds, key=lambda k: kinds[k][0])
        dom_ms, dom_n = kinds[dom]
        gemm_kinds = ("gemm_t0", "gemm_z", "gemm_elt", "fwd_persistent")
        # the all-layer persistent kernel runs every product of the forward in one launch
        flops_per_launch = F_FWD * B if dom == "fwd_persistent" else F_GEMM_PER_COL * B
        mma_passes = {"fp32": 1, "tf32": 1, "tf32x3": 3, "bf16": 1, "tf32_bf16x2": 2}[precision]
        # tensor-pipe peak of the headline arithmetic, from the two measured dense peaks (TF32 and bf16 cuBLAS burst figures):
        # time per product = one kind::tf32 pass per TF32 pass + one kind::f16 pass (twice the rate) per bf16 correction pass
        tensor_peak_prec = {"fp32": peaks["tf32_tflops"], "tf32": peaks["tf32_tflops"], "tf32x3": peaks["tf32_tflops"] / 3.0,
                            "bf16": peaks["bf16_tflops"],
                            "tf32_bf16x2": 1.0 / (1.0 / peaks["tf32_tflops"] + 2.0 / peaks["bf16_tflops"])}[precision]
        peak_note = {"tf32_bf16x2": "1 / (1/TF32 dense + 2/bf16 dense): one kind::tf32 pass + two kind::f16 correction passes",
                     "tf32x3": "TF32 dense / 3 passes", "bf16": "bf16 dense"}.get(precision, "TF32 dense")
        # algorithmic HBM bytes per launch of each fused product kernel: the activation
        # operand + the epilogue inputs read + the outputs written (weights are L2-resident, 3xTF32 splits never
        # leave shared memory, prox masks are not written in inference)
        bytes_per_launch = {
            "gemm_t0": 4.0 * B * (D + 3 * M + 2 * M),        # Z0 | E0, X, L0 | T0, V
            "gemm_z": 4.0 * B * (M + D + D),                 # V | Z_{k-1} | Z_k
            "gemm_elt": 4.0 * B * (D + 3 * M + 4 * M),       # Z_k | X, E_{k-1}, L_{k-1} | E_k, T_{k+1}, L_k, V
        }
        # one launch of the persistent kernel = T_0 + K x (W V kernel + A Z kernel) worth of algorithmic traffic
        bytes_per_launch["fwd_persistent"] = (bytes_per_launch["gemm_t0"] +
                                              K_LAYERS * (bytes_per_launch["gemm_z"] + bytes_per_launch["gemm_elt"]))
        avg_ms = dom_ms / dom_n
        tensor_peak = tensor_peak_prec
        tensor_ach = flops_per_launch / (avg_ms * 1e-3) / 1e12 if dom in gemm_kinds else None
        hbm_ach = bytes_per_launch.get(dom, 0.0) / (avg_ms * 1e-3) / 1e9 if dom in gemm_kinds else None
        tensor_frac = tensor_ach / tensor_peak if tensor_ach else None
        hbm_frac = hbm_ach / peaks["hbm_gbs"] if hbm_ach else None
        total_prof = sum(v[0] for v in kinds.values())
        # the binding roofline of the dominant kernel is the one it sits closer to
        hbm_bound = (hbm_frac or 0) >= (tensor_frac or 0)
        roofline = {
            "bound": "hbm" if hbm_bound else "tensor", "kernel": dom,
            "achieved": hbm_ach if hbm_bound else tensor_ach,
            "peak": peaks["hbm_gbs"] if hbm_bound else tensor_peak,
            "unit": "GB/s" if hbm_bound else "TFLOP/s",
            "frac": hbm_frac if hbm_bound else tensor_frac,
            "traffic": TRAFFIC_NCU.get((precision, dom)),
            "peak_source": "HBM copy %s; tensor peak of precision %s = %s (TF32 dense %s, bf16 dense %s)" %
                           (peaks["source"], precision, peak_note, peaks["tf32_source"], peaks["source"]),
            "tensor": {"achieved_tflops": tensor_ach, "peak_tflops": tensor_peak, "frac": tensor_frac,
                       "frac_of_tf32x3_peak": (tensor_ach / (peaks["tf32_tflops"] / 3.0)) if tensor_ach else None,
                       "tf32x3_peak_tflops": peaks["tf32_tflops"] / 3.0},
            "hbm": {"achieved_gbs": hbm_ach, "peak_gbs": peaks["hbm_gbs"], "frac": hbm_frac,
                    "algorithmic_bytes_per_launch": bytes_per_launch.get(dom)},
            "share_of_step": dom_ms / total_prof if total_prof else None,
            "avg_launch_ms": avg_ms,
            "algorithmic_flops_per_launch": flops_per_launch,
            "per_kind_ms_per_step": {k: v[0] / args.steps for k, v in kinds.items()},
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32x3": "tf32x3", "tf32": "tf32", "bf16": "bf16", "tf32_bf16x2": "tf32_bf16x2"}[precision], "data": "synthetic",
            "config": {"workload": "C1 scalar D-LADMM forward (main_syn_l1l1_scalar.py), m=250 d=500 K=15, "
                                   "%d instances per GPU, all K iterates of Z,E,L,T returned (reference API)" % B,
                       "variant": VARIANT, "precision": precision, "columns_per_gpu": B,
                       "l2": "inputs+outputs per step (%.1f GB) far exceed the 126 MB L2; no explicit flush" %
                             (4.0 * (K_LAYERS * (D + 3 * M) + 4 * M + D) * B / 1e9),
                       "parallelism": "columns sharded, %d rank(s), no data-path collective" % world},
            "algorithmic_tflops": value * F_FWD / 1e12,
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes // args.steps),
                    "d2h_bytes_per_step": int(K_LAYERS * 4), "ms_per_step": ms_e2e / args.steps,
                    "api": "for x in dl.HostFeed(pinned host batches): DLADMMNet.forward_objective(x, alpha) -> all K iterates "
                           "+ per-layer objective copied to the host and read there every step (the read of step i-1 overlaps step i); "
                           "the upload of batch i+1 overlaps the forward of batch i (double buffer), each batch is copied "
                           "H2D once inside the timed region"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        tp = tensor_peak_prec

        def train_block(t, ms):
            Kt, Bt = t["K"], t["columns_per_gpu"]
            f_train = 2.0 * M * D * (5 * Kt + 1)                          # F_fwd + F_bwd per sample
            per_launch = {}
            for kind in ("gemm_z", "gemm_elt", "bwd_gemm_dz", "bwd_gemm_dw", "bwd_gemm_dv"):
                if kind in t["kernel_ms"]:
                    avg = t["kernel_ms"][kind] / t["kernel_launches"][kind]
                    per_launch[kind] = {"avg_launch_ms": avg, "tflops": F_GEMM_PER_COL * Bt / (avg * 1e-3) / 1e12,
                                        "frac_of_tensor_peak": F_GEMM_PER_COL * Bt / (avg * 1e-3) / 1e12 / tp}
            dom_t = max(per_launch, key=lambda k: t["kernel_ms"][k]) if per_launch else None
            ach = Bt * f_train / (ms * 1e-3) / 1e12
            return {"metric": "dladmm_train_samples_per_sec", "variant": t["variant"], "K": Kt,
                    "value": world * Bt / (ms * 1e-3), "unit": "samples/s", "columns_per_gpu": Bt, "ms_per_step": ms,
                    "library_kernel_ms_per_step": t["kernel_ms"],
                    "roofline": {"bound": "tensor", "achieved": ach, "peak": tp, "unit": "TFLOP/s", "frac": ach / tp,
                                 "frac_of_tf32x3_peak": ach / (peaks["tf32_tflops"] / 3.0),
                                 "algorithmic_flops_per_sample": f_train, "dominant_kernel": dom_t, "per_kernel": per_launch,
                                 "what": "whole training step per GPU against the issue-rate tensor peak of precision %s (%s); frac_of_tf32x3_peak = the same "
                                         "against TF32 dense / 3, the peak round 1 quoted; per_kernel = one launch of each product kernel" % (precision, peak_note)},
                    "what": "DLADMMNet.%s(x).backward(): forward + fused objective + backward + parameter gradients%s"
                            % ("lasso_loss" if t["variant"] == "lasso" else "l1l1_loss", " + in-backward NCCL allreduce (rank-equality asserted)" if world > 1 else "")}
        if trains:
            line["train"] = train_block(trains[0], ms_train_step)
            for i, t in enumerate(trains[1:], 1):
                line["train_" + t["variant"]] = train_block(t, extra["train%d" % i])
        if variants_ms:
            line["c1_forward_other_modes"] = {k: {"ms_per_step": extra[k], "value": world * B / (extra[k] * 1e-3)} for k in variants_ms}
            line["c1_forward_other_modes"]["what"] = ("the same C1 forward in the other fp32-class mode (tf32x3: three kind::tf32 passes / tf32_bf16x2: one + two bf16 "
                                                      "correction passes), with single-pass TF32 / bf16 operands (stated-tolerance modes, all "
                                                      "iterates returned) and with last_only=True in the headline precision")
        if e2e_z:
            line["e2e_final_z"] = dict(e2e_z, ms_per_step=extra["e2e_z"], value=world * B / (extra["e2e_z"] * 1e-3), unit=UNIT,
                                       api="HostFeed -> DLADMMNet.forward(x, last_only=True) -> HostDrain: the final Z of every step copied to pinned host memory "
                                           "(copy of step i-1 overlaps the forward of step i) and read on the host")
        if c5:
            f5 = 2.0 * 1000 * 2000 * (2 * 40 + 1)                  # algorithmic flops per instance (324 MFLOP)
            B5 = c5["columns_per_gpu"]
            line["c5"] = {
                "workload": "large-scale shape of BASELINE configs[4]: scalar D-LADMM forward, m=1000 d=2000 K=40, %d instances per GPU" % B5,
                "unit": UNIT, "tensor_peak_tflops": tp, "tf32x3_peak_tflops": peaks["tf32_tflops"] / 3.0,
                "peak_note": "frac_of_tensor_peak is against the issue-rate peak of precision %s (%s); round 1 quoted the 3xTF32 peak "
                             "(tf32x3_peak_tflops): divide algorithmic_tflops_per_gpu by it for that figure" % (precision, peak_note),
                "all_iterates": {"ms_per_step": ms_c5_all, "value": world * B5 / (ms_c5_all * 1e-3),
                                 "algorithmic_tflops_per_gpu": B5 * f5 / (ms_c5_all * 1e-3) / 1e12,
                                 "frac_of_tensor_peak": B5 * f5 / (ms_c5_all * 1e-3) / 1e12 / tp},
                "last_only": {"ms_per_step": ms_c5_last, "value": world * B5 / (ms_c5_last * 1e-3),
                              "algorithmic_tflops_per_gpu": B5 * f5 / (ms_c5_last * 1e-3) / 1e12,
                              "frac_of_tensor_peak": B5 * f5 / (ms_c5_last * 1e-3) / 1e12 / tp},
                "roofline": {"bound": "tensor", "achieved": B5 * f5 / (ms_c5_all * 1e-3) / 1e12, "peak": tp, "unit": "TFLOP/s",
                             "frac": B5 * f5 / (ms_c5_all * 1e-3) / 1e12 / tp,
                             "hbm_gbs_all_iterates": 4.0 * (40 * (2000 + 3 * 1000) + 2 * 1000) * B5 / (ms_c5_all * 1e-3) / 1e9},
            }
            if "total_1m_ms" in extra:
                tot = args.c5_total // world * world
                line["c5"]["total_1m"] = {"instances_total": tot, "scaling": "strong", "ms": extra["total_1m_ms"],
                                          "value": tot / (extra["total_1m_ms"] * 1e-3), "unit": UNIT,
                                          "chunks_per_rank": c5["total_1m_chunks"], "chunk_columns": c5["total_1m_chunk_columns"],
                                          "algorithmic_tflops_per_gpu": tot / world * f5 / (extra["total_1m_ms"] * 1e-3) / 1e12,
                                          "frac_of_tensor_peak": tot / world * f5 / (extra["total_1m_ms"] * 1e-3) / 1e12 / tp,
                                          "what": "BASELINE configs[4] at its stated size: %d instances split over %d GPU(s), final iterate + fused "
                                                  "per-layer objective (last_only), no data-path collective" % (tot, world)}
            if "train_ms" in extra:
                f5t = 2.0 * 1000 * 2000 * (5 * 40 + 1)
                Bt5 = c5["train_columns"]
                line["c5"]["train"] = {"columns_per_gpu": Bt5, "ms_per_step": extra["train_ms"], "value": world * Bt5 / (extra["train_ms"] * 1e-3),
                                       "unit": "samples/s", "algorithmic_tflops_per_gpu": Bt5 * f5t / (extra["train_ms"] * 1e-3) / 1e12,
                                       "frac_of_tensor_peak": Bt5 * f5t / (extra["train_ms"] * 1e-3) / 1e12 / tp,
                                       "gradient_bytes_allreduced": int(4 * 40 * (1000 * 2000 + 6)) if world > 1 else 0}
                if "train_bucketed_ms" in extra:
                    line["c5"]["train"]["allreduce"] = {
                        "schedule": "one collective on the step's flat gradient buffer, stream-ordered after the last backward kernel",
                        "ms_per_step_bucketed_overlap": extra["train_bucketed_ms"],
                        "bucketed_overlap": "opt-in sync_gradients(bucket_mb=32): weight gradients in 32 MB buckets (4 layers each), reverse layer "
                                            "order, issued from a side stream when the library's per-layer event fires (timed first, on a "
                                            "cooler chip; interleaved A/B: profiles/r02_c5_sync_sweep.md)"}
        if world == 1 and not args.quick:
            line["eager_b200"] = eager_b200_leg(dev, A_host, X)
            line["small_batch"] = small_batch_leg(dl, dev, A_host, X, precision)
            line["safeguard"] = safeguard_leg(dl, dev, A_host, X, precision)
        if world == 1 and not args.no_cpu_baseline:
            times, threads, kind = cpu_forward_rate(CPU_SAMPLE_COLS, 1 + CPU_SAMPLE_REPS)
            cpu_val = CPU_SAMPLE_COLS * CPU_SAMPLE_REPS / sum(times[1:])
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": "%d timed forwards of %d columns (same m,d,K, all iterates; %.1f s of host time) after 1 warm-up"
                                              % (CPU_SAMPLE_REPS, CPU_SAMPLE_COLS, sum(times[1:]))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None, choices=[None, "fp32", "tf32x3", "tf32", "bf16", "tf32_bf16x2"])
    ap.add_argument("--columns", type=int, default=B_PER_GPU, help="problem instances per GPU")
    ap.add_argument("--train-columns", type=int, default=65536, help="columns per GPU for the training leg (0 = skip)")
    ap.add_argument("--c5-columns", type=int, default=32768,
                    help="columns per GPU for the m=1000 d=2000 K=40 leg (BASELINE configs[4]); 0 = skip")
    ap.add_argument("--c5-total", type=int, default=1048576,
                    help="total instances of the strong-scaling C5 leg, split over the ranks (0 = skip)")
    ap.add_argument("--c5-train-columns", type=int, default=8192, help="columns per GPU of the C5 training leg (0 = skip)")
    ap.add_argument("--quick", action="store_true", help="headline legs only (forward, e2e, scalar training, C5 forward)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.quick:
        args.c5_total = args.c5_train_columns = 0
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
