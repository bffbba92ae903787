#!/usr/bin/env python
"""Benchmark of the sweeping MPS-classifier hot path (BASELINE.json metric: bond-updates/s and s/sweep at
Ns = 60k samples, bond dimension D = 64, 196 sites, 10 labels, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full sweep over the batch: forward() (environment build) + S-1 = 195 bond updates, directions
alternating like Network.train (NC:310-335).  Prints ONE JSON line (rank 0).

  value     device-resident throughput: inputs (phi, labels) already in HBM, whole sweeps enqueued without host syncs
  e2e       same metric through the reference-facing API (Network.forward / Network.sweep) with HOST NumPy buffers:
            the host->device copy of X and the device->host read of f are inside the timed region
  roofline  the dominant kernel, timed live with CUDA events around its launches inside the timed region
  cpu_baseline  the NumPy oracle (a port of the reference's algorithm; the Python reference cannot travel to the GPU
            box) timed on this box's host cores on a bounded sample of the same workload

N > 1 (torchrun): the 60k samples are sharded across ranks (strong scaling), dB + metrics are all-reduced (NCCL)
once per bond update, the batch-independent part (bond update, SVD) is replicated.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(S=196, L=10, D=64, Ns=60000, lr=1e-4, wd=1e-3, act="linear", loss="MSE", L2=True, seed=2)
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "roofline_traffic.json")


def synthetic_data(Ns, S, L, seed):
    """10-class stripe templates + uniform noise, feature-mapped on the host."""
    import tensornetworkforml_b200.data_generator as gen
    np.random.seed(seed)
    side = int(round(S ** 0.5))
    data, labels = gen.create_multiclass_dataset(Ns, side, L, 0.7)
    return gen.psi(data.reshape(Ns, -1)), labels.astype(np.int64)


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle on a bounded sample (interior bond updates at the full Ns and D)
# ---------------------------------------------------------------------------------------------------
def cpu_bond_updates(n_warm, n_timed, Ns, D, L, lr, wd, act, loss):
    from oracle import mps_oracle as O
    rng = np.random.default_rng(0)
    n = n_warm + n_timed
    S = n + 4
    sites = [rng.standard_normal((D, 2, D)) / np.sqrt(2 * D) for _ in range(S)]
    sites[0] = rng.standard_normal((1, 2, D)) / np.sqrt(2)
    sites[-1] = rng.standard_normal((D, 2, 1)) / np.sqrt(2 * D)
    p0 = 1
    sites[p0] = rng.standard_normal((D, 2, L, D)) / np.sqrt(2 * D)
    net = O.OracleMPS(sites, L, act_fn=act, loss_fn=loss, rule="fixed", max_bond=D, l_pos=p0)
    net.phi = O.feature_map(rng.random((Ns, S)))
    net.env = [None] * (S + 1)
    net.env[0] = np.ones((Ns, 1))
    net.env[1] = rng.standard_normal((Ns, D))
    for p in range(p0 + 2, S):
        net.env[p] = rng.standard_normal((Ns, D)) / np.sqrt(D)
    net.env[S] = np.ones((Ns, 1))
    net._norm = [np.eye(D) for _ in range(S + 1)]
    y = rng.integers(0, L, Ns)
    y1h = np.zeros((Ns, L))
    y1h[np.arange(Ns), y] = 1
    f = rng.standard_normal((Ns, L)) * 0.1
    times = []
    for i in range(n):
        t0 = time.perf_counter()
        f = net.sweep_step(f, y1h, lr, wd, True, False)
        times.append(time.perf_counter() - t0)
    return times[n_warm:]


def workload_config(S, L, D, Ns, dtype):
    """The `config` block both arms print (identical text: the driver compares them)."""
    default_cfg = (S, L, D, Ns) == (CFG["S"], CFG["L"], CFG["D"], CFG["Ns"])
    side = int(round(S ** 0.5))
    return dict(workload="%s: %dx%d synthetic %d-label stripes, S=%d sites, L=%d, D=%d (fixed-D truncation), Ns=%d "
                         "samples total, %s, linear/MSE, L2 norm-environment term on"
                         % ("config3" if default_cfg else "variant of config3", side, side, L, S, L, D, Ns,
                            "FP64" if dtype == "f64" else "FP32/TF32"),
                S=S, L=L, D=D, Ns=Ns)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def blas_threads(n):
    """All host threads for the CPU arm, also under torchrun (which exports OMP_NUM_THREADS=1 to its children: with
    it the BLAS-backed port ran on ONE core and the N >= 2 ratios of round 1 were inflated ~2x)."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def run_reference_arm(args, rank):
    if rank != 0:
        return
    c = dict(CFG, Ns=args.ns, D=args.D, S=args.S, L=args.L)
    nthr = host_threads()
    with blas_threads(nthr):
        times = cpu_bond_updates(args.warmup, args.steps, c["Ns"], c["D"], c["L"], c["lr"], c["wd"], c["act"], c["loss"])
    per = float(np.mean(times))
    val = 1.0 / per
    line = dict(impl="reference", metric="bond_updates_per_s", value=val, unit="bond-updates/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=per * 1e3, higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=workload_config(c["S"], c["L"], c["D"], c["Ns"], "f64"),
                cpu_baseline=dict(value=val, unit="bond-updates/s", cores=nthr, kind="port",
                                  sample="%d interior bond updates (D=%d both sides, L=%d) at the full Ns=%d, "
                                         "NumPy/OpenBLAS oracle port of the reference sweep_step on %d BLAS threads; "
                                         "each step = 1 bond update" % (args.steps, c["D"], c["L"], c["Ns"], nthr)),
                e2e=dict(value=val, unit="bond-updates/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                s_per_sweep_extrapolated=per * (c["S"] - 1))
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        if gpu_index is None:
            return
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms",
                                          os.environ.get("TNML_BENCH_SMI_MS", "100")], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ns", type=int, default=CFG["Ns"], help="total samples (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the API arm (bond-dimension sweeps: device numbers only)")
    ap.add_argument("--no-kernel-pass", action="store_true",
                    help="skip the second, event-bracketed pass (no per-kernel table and no roofline object: A/B runs only)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="f64 = the parity path (default, the BASELINE metric); f32 = FP32 storage / TF32 tcgen05 variant")
    ap.add_argument("--D", type=int, default=CFG["D"], help="bond dimension (other BASELINE.json configs)")
    ap.add_argument("--S", type=int, default=CFG["S"], help="number of sites (a square number)")
    ap.add_argument("--L", type=int, default=CFG["L"], help="number of labels")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import tensornetworkforml_b200 as tn
    from tensornetworkforml_b200 import _lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    c = dict(CFG, Ns=args.ns, D=args.D, S=args.S, L=args.L)
    S, L, D, Ns = c["S"], c["L"], c["D"], c["Ns"]
    X_all, y_all = synthetic_data(Ns, S, L, c["seed"])
    from tensornetworkforml_b200.parallel import shard_bounds
    lo, hi = shard_bounds(Ns, rank, world)
    X, y = np.ascontiguousarray(X_all[lo:hi]), np.ascontiguousarray(y_all[lo:hi])
    del X_all
    np.random.seed(c["seed"])                      # identical initial weights on every rank
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        net = tn.Network(N=S, M=D, L=L, normalize=True, calibration_X=X[:min(len(X), 2048)], act_fn=c["act"],
                         loss_fn=c["loss"], truncation="fixed", max_bond=D, device="cuda:%d" % local_rank,
                         dtype="float64" if args.dtype == "f64" else "float32")
    eng = net._engine()
    lib = _lib.lib()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def device_step():
        eng.forward()
        left = eng.l_pos == S - 1
        eng.begin_sweep(y_dev, left, c["L2"])
        for _ in range(S - 1):
            eng.sweep_step(c["lr"], c["wd"], c["L2"], left)
        # the sweep's record (metrics, update statistics, singular values incl. the batched deferred-tail solve) is
        # read once per sweep, as Network.sweep does: same work as the API arm
        return eng.history()

    api_t = dict(forward=0.0, sweep=0.0, n=0)

    def api_step():
        t0 = time.perf_counter()
        f = net.forward(X)
        t1 = time.perf_counter()
        left = net.l_pos == S - 1
        out = net.sweep(X, y, f, c["lr"], c["wd"], L2_flag=c["L2"], left_dir=left)
        api_t["forward"] += t1 - t0
        api_t["sweep"] += time.perf_counter() - t1
        api_t["n"] += 1
        return out

    # ---- device-resident arm ------------------------------------------------------------------
    eng.load_input(X)
    y_dev = torch.from_numpy(y.astype(np.int32)).to(eng.device)
    for _ in range(args.warmup):
        device_step()
    barrier()
    # one sampler (rank 0's GPU) is enough: eight 10 Hz nvidia-smi pollers measurably slowed the 8-rank run
    clocks = ClockSampler(local_rank if rank == 0 else None)
    k0 = lib.tnml_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    marks = []
    for _ in range(args.steps):
        device_step()
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
    e1.record()
    barrier()
    sweep_ms = [(e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]) for i in range(len(marks))]
    launches = lib.tnml_kernel_launches() - k0
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    # per-kernel live timing in a SECOND pass of the same K steps (the event pairs around every call would perturb the
    # headline number of the short-step configurations); `share` below is relative to that pass
    eng.timers = {}
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record()
    for _ in range(0 if args.no_kernel_pass else args.steps):
        device_step()
    t1e.record()
    barrier()
    ms_timed_pass = t0e.elapsed_time(t1e)
    timers, eng.timers = eng.timers, None
    t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    n_updates = args.steps * (S - 1)
    value = n_updates / (ms * 1e-3)
    hist = eng.history()
    finite = bool(np.isfinite(hist["mae"]).all())
    sv_raw = eng.hist["svals"][:eng.hist["n"]].cpu().numpy()
    nsv = eng.hist["nsv"]
    jac = np.array([[sv_raw[i, nsv[i]], sv_raw[i, nsv[i] + 1]] for i in range(len(nsv)) if nsv[i] == 2 * D])
    jacobi_sweeps = None
    if len(jac):
        fast = jac[:, 0] >= 100                       # 100 + sweeps marks a warm-started (deflation) split
        p1 = np.where(fast, jac[:, 0] - 100, jac[:, 0])
        jacobi_sweeps = dict(pass1_mean=float(np.nanmean(p1)), pass1_max=float(np.nanmax(p1)),
                             pass2_mean=float(np.nanmean(jac[:, 1])), pass2_max=float(np.nanmax(jac[:, 1])),
                             warm_started_fraction=float(fast.mean()),
                             refusals={("code%d" % int(c)): int(n) for c, n in zip(*np.unique(
                                 [sv_raw[i, nsv[i] + 2] for i in range(len(nsv)) if nsv[i] == 2 * D and
                                  not sv_raw[i, nsv[i]] >= 100 and np.isfinite(sv_raw[i, nsv[i] + 2])], return_counts=True))},
                             note="last sweep, splits with short side 2D; warm-started = svd_fast.cuh path taken")
    smin = np.array([hist["svals"][i][-1] / hist["svals"][i][0] for i in range(len(nsv)) if nsv[i] == 2 * D])
    spectrum = dict(sigma_min_over_max_median=float(np.median(smin)), sigma_min_over_max_min=float(smin.min())) if len(smin) else None

    # ---- per-kernel live timing -> roofline --------------------------------------------------------
    kern = {}
    for name, evs in timers.items():
        tot = sum(a.elapsed_time(b) for a, b, _ in evs)
        fl = sum(f for _, _, f in evs)
        kern[name] = dict(calls=len(evs), ms_total=tot, share=tot / ms_timed_pass,
                          tflops=(fl / (tot * 1e-3) / 1e12) if tot else 0.0,
                          avg_ms=tot / max(1, len(evs)))
    peak = 37.06
    peak_src = "fallback constant (profiles/fp64_peak_r01.json missing)"
    if os.path.exists(FP64_PEAK_FILE):
        pk = json.load(open(FP64_PEAK_FILE))
        peak = float(pk["fp64_tflops_sustained"])
        peak_src = ("measured FP64 DMMA peak, tools/fp64_peak.cu on this pool's B200 (profiles/fp64_peak_r01.json); "
                    "MEASURED_PEAKS.json has no FP64 entry")
    # k_grad: the K = Ns tensor-core reduction, the Ns-proportional kernel on the critical path (it has the whole GPU;
    # k_project does the same FLOPs but is deliberately capped to ~110 SMs because it runs beside the SVD split)
    top = "grad"
    traffic = None
    if os.path.exists(TRAFFIC_FILE):
        traffic = json.load(open(TRAFFIC_FILE)).get(top)
    kname = "k_grad<FULL>"
    if args.dtype == "f32":
        # no measured TF32 peak on this pool: half of MEASURED_PEAKS.json's bf16 burst figure (TF32 runs at half the bf16
        # rate on tcgen05; nominal 1.1 PFLOP/s dense)
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = (json.load(open(mp))["bf16_tflops"] / 2) if os.path.exists(mp) else 1100.0
        peak_src = "half of the measured bf16 burst peak in MEASURED_PEAKS.json (TF32 = bf16 / 2 on tcgen05); not measured directly"
        kname, traffic = "k_grad_tc (tcgen05 kind::tf32)", None
    if args.no_kernel_pass:
        kern[top] = dict(calls=0, ms_total=0.0, share=0.0, tflops=0.0, avg_ms=0.0)
    roofline = dict(kernel=kname, bound="tensor", achieved=kern[top]["tflops"], peak=peak, unit="TFLOP/s",
                    frac=kern[top]["tflops"] / peak, traffic=traffic, peak_source=peak_src,
                    flops_per_launch="8*Ns*L*Dl*Dr per launch (2 flops x Ns x (2 Dl) x (2 L Dr)), summed over the "
                                     "launches of the timed region / summed CUDA-event time")

    # whole bond update against the FP64 tensor-pipe roofline: algorithmic FLOPs of SURVEY.md section 8d (uniform bond
    # dimension D in the interior; the chain ends are smaller, so this slightly over-counts) / measured step time
    if args.dtype == "f64":
        F_update = Ns * D * D * (16 * L + 4) + 4 * Ns * L * D
        t_update = ms * 1e-3 / n_updates
        roofline["north_star"] = ("whole bond update at %d GPU(s): %.3f of the FP64 tensor-pipe roofline (target >= 0.5 "
                                  "at 8 GPUs)" % (world, F_update / (peak * 1e12) / world / t_update))
        roofline["whole_update"] = dict(gflop=F_update / 1e9, ideal_ms=F_update / (peak * 1e12) * 1e3 / world,
                                        measured_ms=t_update * 1e3, frac=F_update / (peak * 1e12) / world / t_update,
                                        note="all kernels of a bond update incl. the latency-bound SVD split and launch "
                                             "gaps, against the same FP64 DMMA peak (x n_gpus)")

    # ---- end-to-end arm through the reference-facing API (host buffers) -------------------------------
    e2e = None
    if not args.no_e2e:
        net.l_pos, net._host_fresh = eng.l_pos, False       # the device arm drove the engine directly
        net.register_input(X)                               # the caller's batch, page-locked in place (explicit opt-in)
        for _ in range(max(1, min(args.warmup, 2))):
            api_step()
        barrier()
        api_t.update(forward=0.0, sweep=0.0, n=0)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            f_host = api_step()
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([max(e0.elapsed_time(e1), wall)], dtype=torch.float64, device=eng.device)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = dict(value=n_updates / (e2e_ms * 1e-3), unit="bond-updates/s", h2d_bytes_per_step=int(X.nbytes + y.size * 4),
                   d2h_bytes_per_step=int(f_host.elem.nbytes + (S - 1) * (4 + 6 + 4 * D) * 8),
                   ms_per_step=e2e_ms / args.steps, api="Network.forward(X_host) + Network.sweep(X_host, y_host, f)",
                   forward_ms=api_t["forward"] / max(1, api_t["n"]) * 1e3,
                   sweep_ms=api_t["sweep"] / max(1, api_t["n"]) * 1e3)

    line = dict(metric="bond_updates_per_s", value=value, unit="bond-updates/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms / args.steps, s_per_sweep=ms / args.steps * 1e-3,
                higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f64" if args.dtype == "f64" else "tf32 (fp32 storage, fp64 bond algebra + SVD)", data="synthetic",
                config=dict(workload_config(S, L, D, Ns, args.dtype), samples_per_gpu=hi - lo,
                            parallelism="sample-shard x%d" % world,
                            l2_flush="inputs exceed L2 (env cache %.1f GB per GPU)" % (eng.env.numel() * eng.esz / 1e9),
                            bond_updates_per_step=S - 1),
                clocks=clk, e2e=e2e, gpu_launches=int(launches), roofline=roofline, kernels=kern,
                sweep_ms=sweep_ms, timed_kernel_pass_ms_per_step=ms_timed_pass / args.steps,
                finite=finite, bonds_mid=eng.bond_dims()[S // 2], jacobi_sweeps=jacobi_sweeps, spectrum=spectrum)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        with blas_threads(host_threads()):
            times = cpu_bond_updates(1, 3, Ns, D, L, c["lr"], c["wd"], c["act"], c["loss"])
        per = float(np.mean(times))
        line["cpu_baseline"] = dict(value=1.0 / per, unit="bond-updates/s", cores=host_threads(), kind="port",
                                    sample="3 interior bond updates (D=%d both sides, L=%d) at the full Ns=%d with the "
                                           "NumPy/OpenBLAS oracle port of sweep_step (1 warm-up update)" % (D, L, Ns),
                                    s_per_sweep_extrapolated=per * (S - 1),
                                    verbatim_reference=dict(
                                        note="the reference itself (only tensor_svd's m replaced) cannot run this shape "
                                             "(78 GB intermediate in update_B) and does not exist on the GPU box; measured "
                                             "in the build container by tools/ref_verbatim_point.py, BASELINE.md section 2",
                                        measured_s_per_bond_update={"Ns=64": 0.54, "Ns=128": 0.75, "Ns=256": 1.58},
                                        config="S=16, D=64, L=10, L2_flag=False, interior bonds",
                                        extrapolated_s_per_bond_update_at_Ns_60000=324.0,
                                        l2_term_extra_s_per_bond_update={"S=16": 4.7, "S=196 (O(S))": 57.0}))
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
