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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
    import torch
    import dladmm_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, D, generator=g)
    A = A / A.pow(2).sum(dim=0, keepdim=True).sqrt()
    Zs = (torch.rand(D, sample_cols, generator=g) < 0.1).float() * torch.randn(D, sample_cols, generator=g)
    Es = (torch.rand(M, sample_cols, generator=g) < 0.1).float() * torch.randn(M, sample_cols, generator=g)
    X = A.mm(Zs) + Es
    Z0 = torch.rand(D, sample_cols, generator=g) / D
    E0 = torch.zeros(M, sample_cols); L0 = torch.zeros(M, sample_cols)
    sd = orc.default_state_dict(VARIANT, A, K_LAYERS, sample_cols, generator=g)
    times = []
    with torch.no_grad():
        orc.forward(VARIANT, sd, A, X[:, :512].contiguous(), Z0[:, :512].contiguous(), E0[:, :512].contiguous(),
                    L0[:, :512].contiguous(), K_LAYERS)                    # warm-up
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.forward(VARIANT, sd, A, X, Z0, E0, L0, K_LAYERS)
            times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def run_reference(args):
    rank, world, _ = _dist_env()
    if rank != 0:
        return 0
    sample = CPU_SAMPLE_COLS
    times, threads = cpu_forward_rate(sample, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    value = sample * len(timed) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C1 scalar D-LADMM forward, m=250 d=500 K=15, all iterates returned",
                   "columns_per_step": sample, "variant": VARIANT},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d-column forward per step (full workload is %d columns per GPU)" % (sample, B_PER_GPU)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dladmm_b200 as dl
    from dladmm_b200 import _lib

    rank, world, local = _dist_env()
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's banner off stdout: rank 0 prints exactly one JSON line
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    precision = args.precision or dl.default_precision()
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
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

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

    # ---- optional training throughput (forward + backward + parameter gradients) ----------------------
    train = None
    if args.train_columns > 0:
        Bt = args.train_columns
        tm = dl.VARIANT_CLASSES[VARIANT](M, 10000, D, Bt, A_host, Z0[:, :Bt].contiguous(), E0[:, :Bt].contiguous(),
                                         L0[:, :Bt].contiguous(), K_LAYERS, precision=precision, device=dev)
        Xt = X[:, :Bt].contiguous()
        if world > 1:
            tm.sync_gradients(True)

        wts = [0.6 ** 3] * (K_LAYERS - 1) + [1.0]

        def step_train():
            tm.zero_grad(set_to_none=True)
            loss, _ = tm.l1l1_loss(Xt, 0.001, wts)       # main_syn_l1l1_scalar.py:289-299, fused
            loss.backward()                              # world > 1: the backward allreduces the gradient buffer in place
        for _ in range(2):
            step_train()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nst = max(2, args.steps // 2)
        t0.record()
        for _ in range(nst):
            step_train()
        t1.record()
        barrier()
        ms_train = t0.elapsed_time(t1)
        _lib.profile_start()
        step_train()
        tprof = _lib.profile_stop()
        train = {"columns_per_gpu": Bt, "ms_per_step": ms_train / nst,
                 "kernel_ms_per_step": {k: v[0] for k, v in tprof.items() if v[1] > 0}}

    # ---- optional: the large-scale shape (BASELINE configs[4] "C5": A 1000x2000, K=40), tensor-bound regime ---------
    c5 = None
    if args.c5_columns > 0:
        if args.train_columns > 0:
            del tm, Xt
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
        del model5, data5
        torch.cuda.empty_cache()

    # ---- max over ranks -------------------------------------------------------------------------------
    stats = torch.tensor([ms_total, ms_e2e, train["ms_per_step"] if train else 0.0,
                          c5["all_iterates"] if c5 else 0.0, c5["last_only"] if c5 else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_train_step, ms_c5_all, ms_c5_last = [float(v) for v in stats.tolist()]

    if rank == 0:
        peaks = _peaks()
        ms_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total * 1e-3)
        e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
        # dominant kernel = the kind with the most device time in the per-kernel pass
        kinds = {k: v for k, v in prof.items() if v[1] > 0}
        dom = max(kinds, key=lambda k: kinds[k][0])
        dom_ms, dom_n = kinds[dom]
        gemm_kinds = ("gemm_t0", "gemm_z", "gemm_elt")
        flops_per_launch = F_GEMM_PER_COL * B
        mma_passes = {"fp32": 1, "tf32": 1, "tf32x3": 3}[precision]
        # algorithmic HBM bytes per launch of each fused product kernel: the activation
        # operand + the epilogue inputs read + the outputs written (weights are L2-resident, 3xTF32 splits never
        # leave shared memory, prox masks are not written in inference)
        bytes_per_launch = {
            "gemm_t0": 4.0 * B * (D + 3 * M + 2 * M),        # Z0 | E0, X, L0 | T0, V
            "gemm_z": 4.0 * B * (M + D + D),                 # V | Z_{k-1} | Z_k
            "gemm_elt": 4.0 * B * (D + 3 * M + 4 * M),       # Z_k | X, E_{k-1}, L_{k-1} | E_k, T_{k+1}, L_k, V
        }
        avg_ms = dom_ms / dom_n
        tensor_peak = peaks["tf32_tflops"] / mma_passes
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
            "peak_source": "HBM copy %s; TF32 dense %s / %d MMA pass(es) for precision %s" %
                           (peaks["source"], peaks["tf32_source"], mma_passes, precision),
            "tensor": {"achieved_tflops": tensor_ach, "peak_tflops": tensor_peak, "frac": tensor_frac},
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
            "dtype": {"fp32": "f32", "tf32x3": "tf32x3", "tf32": "tf32"}[precision], "data": "synthetic",
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
        if train:
            line["train"] = {"metric": "dladmm_train_samples_per_sec", "value": world * train["columns_per_gpu"] / (ms_train_step * 1e-3),
                             "unit": "samples/s", "columns_per_gpu": train["columns_per_gpu"], "ms_per_step": ms_train_step,
                             "library_kernel_ms_per_step": train["kernel_ms_per_step"],
                             "what": "DLADMMNet.l1l1_loss(x).backward(): forward + fused L1-L1 objective + backward + parameter gradients%s, scalar K=15" % (" + NCCL allreduce" if world > 1 else "")}
        if c5:
            f5 = 2.0 * 1000 * 2000 * (2 * 40 + 1)                  # algorithmic flops per instance (324 MFLOP)
            B5 = c5["columns_per_gpu"]
            tp = peaks["tf32_tflops"] / mma_passes
            line["c5"] = {
                "workload": "large-scale shape of BASELINE configs[4]: scalar D-LADMM forward, m=1000 d=2000 K=40, %d instances per GPU" % B5,
                "unit": UNIT, "tensor_peak_tflops": tp,
                "all_iterates": {"ms_per_step": ms_c5_all, "value": world * B5 / (ms_c5_all * 1e-3),
                                 "algorithmic_tflops_per_gpu": B5 * f5 / (ms_c5_all * 1e-3) / 1e12,
                                 "frac_of_tensor_peak": B5 * f5 / (ms_c5_all * 1e-3) / 1e12 / tp},
                "last_only": {"ms_per_step": ms_c5_last, "value": world * B5 / (ms_c5_last * 1e-3),
                              "algorithmic_tflops_per_gpu": B5 * f5 / (ms_c5_last * 1e-3) / 1e12,
                              "frac_of_tensor_peak": B5 * f5 / (ms_c5_last * 1e-3) / 1e12 / tp},
            }
        if world == 1 and not args.no_cpu_baseline:
            times, threads = cpu_forward_rate(CPU_SAMPLE_COLS, 1 + CPU_SAMPLE_REPS)
            cpu_val = CPU_SAMPLE_COLS * CPU_SAMPLE_REPS / sum(times[1:])
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
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
    ap.add_argument("--precision", default=None, choices=[None, "fp32", "tf32x3", "tf32"])
    ap.add_argument("--columns", type=int, default=B_PER_GPU, help="problem instances per GPU")
    ap.add_argument("--train-columns", type=int, default=65536, help="columns per GPU for the training leg (0 = skip)")
    ap.add_argument("--c5-columns", type=int, default=32768,
                    help="columns per GPU for the m=1000 d=2000 K=40 leg (BASELINE configs[4]); 0 = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
