#!/usr/bin/env python
"""Headline benchmark: MPO x MPS apply + SVD-round sweeps/s on BASELINE.json configs[1]
(single chain N=64, d=2, chi=256, MPO chi_W=16, rounded back to chi=256, FP64), one chain per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full sweep of the hot path over one synthetic chain: apply the MPO and round back to chi=256 with
optimal (SVD) truncation.  Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.

  value     device-resident inputs, CUDA-event timed, max over ranks (one independent chain per rank: "replicas only")
  e2e       the same sweep through the public API `syn.mul(W, X, mode="optimized", bond=256)` with HOST (pinned) cores:
            H2D of every input core and D2H of every output core inside the timed region
  roofline  the strided DMMA GEMM (dominant kernel): algorithmic FLOPs of all its launches in one sweep / their summed
            CUDA-event durations, against the cuBLAS DGEMM rate measured in the same process (MEASURED_PEAKS.json
            has no FP64 entry)
  cpu_baseline
            ONE full C2 sweep of the like-for-like CPU algorithm (oracle/svd_numpy.StructuredDensityMatrixSweep: the same
            density-matrix contraction sequence the GPU runs, numpy/BLAS/LAPACK on all host threads), measured, no scaling;
            the GPU result is compared with it in the same run (config.parity_vs_cpu_sweep)
  --impl reference
            the same CPU sweep, cut into K consecutive chunks of its 127 unit operations: every step is a bounded sample
            (one chunk), the K steps together are ONE complete sweep on real data, value = 1 / (sum of the step times)
  --workload c4 | c5
            the batched configs of BASELINE.json as the headline line (states/s, samples/s), batch-sharded over the ranks
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SITES, D_PHYS, CHI, CHI_W = 64, 2, 256, 16
METRIC = "mpo_mps_apply_svd_round_sweeps_per_s"
UNIT = "sweeps/s"
WORKLOAD = "C2 single chain: N=64 d=2 random MPS chi=256 x random MPO chi_W=16, apply + SVD-round to chi=256, FP64"


def capped_bonds(n, d_eff, chi):
    return [1] + [int(min(chi, d_eff ** min(k, n - k, 40))) for k in range(1, n)] + [1]


def make_chain(seed, n=N_SITES, d=D_PHYS, chi=CHI, chiw=CHI_W):
    """SURVEY 8(d) C2 inputs: cores ~ N(0, 1/(l d)) drawn on the host so CPU and GPU see identical bits."""
    rng = np.random.default_rng(seed)
    bx = capped_bonds(n, d, chi)
    bw = capped_bonds(n, d * d, chiw)
    X = [rng.normal(size=(bx[k], d, bx[k + 1])) / np.sqrt(bx[k] * d) for k in range(n)]
    W = [rng.normal(size=(bw[k], d, d, bw[k + 1])) / np.sqrt(bw[k] * d) for k in range(n)]
    return X, W


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the like-for-like density-matrix sweep of the oracle, MEASURED at full size (no flop-model scaling anywhere)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_full_sweep(X, W, chi):
    """One complete C2 sweep on the host cores: (seconds, cores, spectra, discarded)."""
    from oracle import svd_numpy as S
    t0 = time.perf_counter()
    sweep = S.StructuredDensityMatrixSweep(X, W, chi)
    for op in sweep.operations():
        sweep.run(op)
    return time.perf_counter() - t0, sweep.out, sweep.spectra, sweep.discarded


def balanced_chunks(costs, k):
    """Cut a list of costs into k consecutive chunks of roughly equal total cost (boundaries only; never scales a timing)."""
    k = max(1, min(k, len(costs)))
    total, bounds, acc, nxt = float(sum(costs)), [0], 0.0, 1
    for idx, c in enumerate(costs):
        acc += c
        while nxt < k and acc >= total * nxt / k and len(costs) - (idx + 1) >= k - nxt:
            bounds.append(idx + 1)
            nxt += 1
    bounds += [len(costs)] * (k + 1 - len(bounds))
    return [(bounds[j], bounds[j + 1]) for j in range(k)]


def ncu_traffic():
    """`roofline.traffic`: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, parsed from an
    `ncu --set full` capture by tools/ncu_traffic.py into profiles/gemm_traffic.json (committed beside the .txt summary).  It cannot
    be measured inside a timed run (a number taken under a profiler is never a bench value), so it is read from that file with its
    provenance -- or null when the file is absent."""
    path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return {"traffic": float(t["dram_bytes"]), "traffic_note": "%s; algorithmic bytes of that launch: %.4g" % (t["source"], t["algorithmic_bytes"])}
    except Exception:
        return {"traffic": None, "traffic_note": "no ncu capture parsed (profiles/gemm_traffic.json absent)"}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core it can (BLAS/LAPACK thread pools)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=host_threads())
    except Exception:
        pass


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: ONE complete sweep of the like-for-like algorithm on all host threads, cut into `steps` consecutive chunks."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    from oracle import svd_numpy as S
    X, W = make_chain(2)
    # warm-up: the BLAS/LAPACK thread pools on the cheap tail of the chain (environments of the last sites), then start over
    for _ in range(max(args.warmup, 1)):
        warm = S.StructuredDensityMatrixSweep(X, W, CHI)
        for op in warm.operations()[:8]:
            warm.run(op)
    sweep = S.StructuredDensityMatrixSweep(X, W, CHI)
    ops_list = sweep.operations()
    chunks = balanced_chunks([sweep.op_flops(op) for op in ops_list], args.steps)
    secs = []
    for lo, hi in chunks:
        t0 = time.perf_counter()
        for op in ops_list[lo:hi]:
            sweep.run(op)
        secs.append(time.perf_counter() - t0)
    secs += [0.0] * (args.steps - len(secs))             # more steps than unit operations: the surplus steps are empty
    total = float(sum(secs))
    value = 1.0 / total
    sample = ("ONE complete C2 sweep (all 64 sites: 63 environment updates + 64 truncation steps on real data) of "
              "oracle/svd_numpy.StructuredDensityMatrixSweep, cut into %d consecutive chunks = the %d timed steps (%.1f s in all, "
              "longest step %.1f s); value = 1 sweep / total measured time, nothing is extrapolated" % (len(chunks), args.steps, total, max(secs)))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "algorithm": "density-matrix SVD rounding through the (X, W) structure -- the same contraction sequence the GPU arm "
                                                     "executes (oracle/svd_numpy.py), numpy + BLAS/LAPACK eigh",
                   "step": "1/%d of one sweep (consecutive unit operations); steps x ms_per_step = one whole measured sweep" % args.steps,
                   "note": "the reference has no SVD rounding and cannot run d=2 chains with dim>2 (SURVEY fact 4): this is the oracle port",
                   "discarded_weight_total": float(sum(sweep.discarded))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": host_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


class Runtime:
    """torch.distributed / CUDA plumbing shared by the workloads: one process per GPU, barrier + synchronize around every timed
    region, CUDA events on the launching stream, MAX over ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def finish(self, line):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()
        if self.rank == 0:
            emit(line)


def cublas_peak(torch, dev, dtype, allow_tf32=False):
    """Tensor-pipe roofline denominator measured in this process with a library GEMM (8192^3, burst, best of 6): MEASURED_PEAKS.json
    carries HBM and bf16 only."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    try:
        a = torch.randn(8192, 8192, dtype=dtype, device=dev)
        b = torch.randn(8192, 8192, dtype=dtype, device=dev)
        best = 1e30
        for _ in range(6):
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(); torch.matmul(a, b); p1.record(); torch.cuda.synchronize()
            best = min(best, p0.elapsed_time(p1))
        return 2 * 8192.0 ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: TensorDense forward, MPO (16,16,16)x(16,16,16) bond 16, batch 65536 x 4096, batch-sharded
# ---------------------------------------------------------------------------------------------------------------------
C5_BATCH, C5_FLOP_PER_SAMPLE = 65536, 3.7748736e7          # 2 * (16 + 256 + 16) * 65536  


def c5_cpu_forward(x, cores, bias):
    """The reference's float32 arithmetic (layers/TensorDense.py:103-142) restated with numpy tensordot on the host cores."""
    T = x.reshape(x.shape[0], 16, 16, 16)
    T = np.tensordot(T, cores[0], axes=([1], [0]))                 # (s, i2, i3, o1, b1)
    T = np.tensordot(T, cores[1], axes=([1, 4], [0, 2]))           # (s, i3, o1, o2, b2)
    T = np.tensordot(T, cores[2], axes=([1, 4], [0, 2]))           # (s, o1, o2, o3)
    return np.maximum(T.reshape(x.shape[0], -1) + bias.reshape(-1), 0)


def run_c5(args):
    rt = Runtime()
    torch = rt.torch
    from syngular.layers import TensorDense
    from syngular_b200 import ops, parallel
    lo, hi = parallel.shard_bounds(C5_BATCH, rt.rank, rt.world)
    rng = np.random.default_rng(5)
    cores = [rng.normal(scale=0.05, size=s).astype(np.float32) for s in ((16, 16, 16), (16, 16, 16, 16), (16, 16, 16))]
    bias = np.zeros((16, 16, 16), np.float32)
    layer = TensorDense((16, 16, 16), (16, 16, 16), (16, 16), precision="tf32").build(cores, bias)
    g = torch.Generator(device=rt.dev).manual_seed(500 + rt.rank)
    x = torch.randn((hi - lo, 4096), dtype=torch.float32, device=rt.dev, generator=g)
    xh = x.cpu().pin_memory()
    yh = torch.empty_like(xh).pin_memory()
    out = torch.empty_like(x)

    def step_device():
        return ops.tt_dense3_tf32(x, layer._packed, layer._bias32, relu=True, out=out)

    def step_e2e():
        layer(xh, out=yh)            # the public API call on a pinned host batch: H2D, kernel and D2H of the pieces overlap on three streams

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    sampler = ClockSampler(rt.local)
    if rt.rank == 0:
        sampler.start()
    l0 = ops.lib.syn_launch_count()
    ms = rt.timed(step_device, args.steps)
    launches = (ops.lib.syn_launch_count() - l0) / float(args.steps)
    clocks = sampler.stop() if rt.rank == 0 else None
    value = C5_BATCH * args.steps / (ms * 1e-3)
    step_e2e()
    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e = rt.timed(step_e2e, e2e_steps)
    line = None
    if rt.rank == 0:
        # parity inside the bench: the first 512 samples against the float64 restatement (stated TF32 tolerance 4e-3 of the scale)
        from oracle import tensordense_numpy as TD
        ref = TD.forward(x[:512].cpu().numpy().astype(np.float64), [c.astype(np.float64) for c in cores], bias.astype(np.float64), "relu")
        err = float(np.max(np.abs(out[:512].cpu().numpy() - ref)) / np.max(np.abs(ref)))
        peak = cublas_peak(torch, rt.dev, torch.float32, allow_tf32=True)
        per_launch_ms = ms / args.steps
        achieved = C5_FLOP_PER_SAMPLE * (hi - lo) / (per_launch_ms * 1e-3) / 1e12
        cpu = None
        if rt.world == 1 and not args.no_cpu:
            use_all_host_threads()
            n_cpu = 4096
            xs = x[:n_cpu].cpu().numpy()
            c5_cpu_forward(xs[:256], cores, bias)
            t0 = time.perf_counter()
            c5_cpu_forward(xs, cores, bias)
            sec = time.perf_counter() - t0
            cpu = {"value": n_cpu / sec, "unit": "samples/s", "cores": host_threads(), "kind": "port",
                   "sample": "float32 numpy tensordot restatement of layers/TensorDense.py:103-142 on the first %d samples of the batch (%.2f s)" % (n_cpu, sec)}
        line = {
            "metric": "tensordense_forward_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": rt.world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": "C5 TensorDense forward: MPO (16,16,16)x(16,16,16) bond 16, batch 65536 x 4096 float32, bias + ReLU, batch-sharded",
                       "l2": "inputs larger than L2: 1.07 GB of x and 1.07 GB of y per step", "multi_gpu": "batch sharded in contiguous blocks, no collective (outputs stay sharded)",
                       "parity_rel_err_first_512": err, "tolerance": 4e-3},
            "clocks": clocks,
            "e2e": {"value": C5_BATCH * e2e_steps / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(xh.numel() * 4), "d2h_bytes_per_step": int(yh.numel() * 4),
                    "steps": e2e_steps, "api": "TensorDense(..., precision='tf32')(x_host, out=y_host): pinned host batch in, pinned host batch out, streamed in pieces of 8192 samples (upload / kernel / download overlapped)"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "tt_dense3_tf32_kernel (tcgen05.mma.kind::tf32)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": "cuBLAS TF32 GEMM 8192^3 measured in this process",
                         "flop_per_sample": C5_FLOP_PER_SAMPLE},
            "cpu_baseline": cpu,
        }
    rt.finish(line)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[3]: 8192 independent random MPS (N=32, d=2, chi=64): overlaps <A_b|B_b>, batch-sharded, ONE all-gather
# ---------------------------------------------------------------------------------------------------------------------
C4_BATCH, C4_SITES, C4_CHI = 8192, 32, 64


def run_c4(args):
    rt = Runtime()
    torch = rt.torch
    from syngular_b200 import ops, parallel
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    lo, hi = parallel.shard_bounds(C4_BATCH, rt.rank, rt.world)
    bonds = capped_bonds(C4_SITES, 2, C4_CHI)[1:-1]
    A = BMPS.random(hi - lo, (2,) * C4_SITES, bonds, seed=1000 + rt.rank, device=rt.dev)
    B = BMPS.random(hi - lo, (2,) * C4_SITES, bonds, seed=2000 + rt.rank, device=rt.dev)
    full = [1] + list(bonds) + [1]
    flops = sum(4.0 * 2 * a * b * max(a, b) for a, b in zip(full[:-1], full[1:])) * C4_BATCH
    bytes_ = 2 * 8.0 * sum(a * 2 * b for a, b in zip(full[:-1], full[1:])) * C4_BATCH

    def step_device():
        return parallel.gather_scalars(A.overlap(B), C4_BATCH)

    # end to end: a sub-batch of E2E_N pairs per rank from pinned host memory (the full sets are 2 x 12 GB), overlaps copied back
    e2e_n = min(1024, hi - lo)
    Ah = [c[:e2e_n].cpu().pin_memory() for c in A.sites]
    Bh = [c[:e2e_n].cpu().pin_memory() for c in B.sites]
    res_h = torch.empty(e2e_n, dtype=torch.float64).pin_memory()

    def step_e2e():
        a = BMPS([c.to(rt.dev, non_blocking=True) for c in Ah])
        b = BMPS([c.to(rt.dev, non_blocking=True) for c in Bh])
        res_h.copy_(a.overlap(b), non_blocking=True)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        ov = step_device()
    sampler = ClockSampler(rt.local)
    if rt.rank == 0:
        sampler.start()
    l0 = ops.lib.syn_launch_count()
    ms = rt.timed(step_device, args.steps)
    launches = (ops.lib.syn_launch_count() - l0) / float(args.steps)
    clocks = sampler.stop() if rt.rank == 0 else None
    value = C4_BATCH * args.steps / (ms * 1e-3)
    step_e2e()
    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e = rt.timed(step_e2e, e2e_steps)
    line = None
    if rt.rank == 0:
        # parity inside the bench: the first 4 pairs against the oracle's transfer-matrix contraction
        from oracle import ref_numpy as R
        errs = []
        for b in range(4):
            got = float(ov[b].item())
            want = float(R.overlap([c[b].cpu().numpy() for c in A.sites], [c[b].cpu().numpy() for c in B.sites]))
            errs.append(abs(got - want) / max(abs(want), 1e-300))
        peak = cublas_peak(torch, rt.dev, torch.float64)
        achieved = flops / rt.world / (ms / args.steps * 1e-3) / 1e12
        cpu = None
        if rt.world == 1 and not args.no_cpu:
            use_all_host_threads()
            n_cpu = 64
            host = [([c[b].cpu().numpy() for c in A.sites], [c[b].cpu().numpy() for c in B.sites]) for b in range(n_cpu)]
            t0 = time.perf_counter()
            for a, b in host:
                R.overlap(a, b)
            sec = time.perf_counter() - t0
            cpu = {"value": n_cpu / sec, "unit": "states/s", "cores": host_threads(), "kind": "port",
                   "sample": "oracle transfer-matrix `|` (ref_numpy.overlap, MPS:116-129) looped over the first %d pairs (%.2f s)" % (n_cpu, sec)}
        line = {
            "metric": "batched_overlap_states_per_s", "value": value, "unit": "states/s", "n_gpus": rt.world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4 batched inner products: 8192 independent random MPS pairs (N=32, d=2, chi=64), batch-sharded, one all-gather of 8192 float64",
                       "l2": "inputs larger than L2: 24 GB of cores per step over all ranks", "multi_gpu": "contiguous batch blocks per rank; the only collective is the all-gather of the scalars",
                       "parity_rel_err_first_4": max(errs), "hbm_gbs_algorithmic": bytes_ / rt.world / (ms / args.steps * 1e-3) / 1e9,
                       "hbm_peak_gbs": measured_peaks().get("hbm_gbs")},
            "clocks": clocks,
            "e2e": {"value": rt.world * e2e_n * e2e_steps / (ms_e2e * 1e-3), "unit": "states/s",
                    "h2d_bytes_per_step": int(sum(c.numel() for c in Ah + Bh) * 8), "d2h_bytes_per_step": int(e2e_n * 8), "steps": e2e_steps,
                    "api": "BatchedMatrixProductState(host cores).overlap(...) on a %d-pair sub-batch per rank from pinned host memory" % e2e_n},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "overlap_chain_kernel (DMMA.8x8x4)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": "cuBLAS DGEMM 8192^3 measured in this process",
                         "note": "FP64 compute governs this kernel in FP64 (SURVEY 8(d) K8): 4 d chi^3 flop against 16 d chi^2 bytes per plateau site"},
            "cpu_baseline": cpu,
        }
    rt.finish(line)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[2]: 50-qubit brick-wall circuit of Haar-random two-qubit unitaries, depth 20, chi_max = 512, complex128
# ---------------------------------------------------------------------------------------------------------------------
def run_c3(args):
    """One step = one whole circuit (490 gates) from |0...0>; replicas only (a chain is sequential): one circuit per rank."""
    rt = Runtime()
    torch = rt.torch
    from syngular.quantum import Circuit
    from syngular_b200 import ops
    rng3 = np.random.default_rng(3)

    def haar4():
        z = rng3.normal(size=(4, 4)) + 1j * rng3.normal(size=(4, 4))
        q, r = np.linalg.qr(z)
        return (q * (np.diag(r) / np.abs(np.diag(r)))).reshape(2, 2, 2, 2)
    nq, depth, chi = 50, 20, args.chi_max
    structure = [(haar4(), i) for layer in range(depth) for i in range(layer % 2, nq - 1, 2)]
    gate_bytes = int(sum(g.nbytes for g, _ in structure))
    state = {}

    def step():
        circ = Circuit(nq, structure=structure, chi_max=chi)          # the gates are host arrays: every step uploads them (H2D inside the step)
        circ.run()
        state["st"] = circ.get().state

    warm = max(args.warmup, 1)                                         # a circuit at chi_max = 512 takes seconds: one warm-up run settles every workspace
    for _ in range(warm):
        step()
    sampler = ClockSampler(rt.local)
    if rt.rank == 0:
        sampler.start()
    l0 = ops.lib.syn_launch_count()
    ms = rt.timed(step, args.steps)
    launches = (ops.lib.syn_launch_count() - l0) / float(args.steps)
    clocks = sampler.stop() if rt.rank == 0 else None
    value = rt.world * len(structure) * args.steps / (ms * 1e-3)
    line = None
    if rt.rank == 0:
        st = state["st"]
        norm2 = float(np.real(st.conj() | st))                        # D2H read of the result (the truncation loss of the run)
        line = {
            "metric": "circuit_gates_per_s", "value": value, "unit": "gates/s", "n_gpus": rt.world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": {"workload": "C3 quantum circuit: 50 qubits, brick-wall, depth 20, Haar-random two-qubit unitaries (seed 3), SVD truncation chi_max=%d, complex128" % chi,
                       "multi_gpu": "replicas only: one circuit per rank", "max_bond": int(max(c.shape[2] for c in st.sites[:-1])), "norm2": norm2,
                       "l2": "the saturated two-site updates work on 2048 x 2048 real embeddings (32 MB each) and their GEMM workspaces"},
            "clocks": clocks,
            "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": gate_bytes, "d2h_bytes_per_step": 16,
                    "api": "Circuit(50, structure=host gates, chi_max).run(): every step starts from host gate arrays and ends with a host read of the norm"},
            "gpu_launches": launches,
            "roofline": None, "cpu_baseline": None,
        }
    rt.finish(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import syngular as syn
    from syngular.tensor import MatrixProductOperator as MPO, MatrixProductState as MPS, _sweeps as sw
    from syngular_b200 import ops

    dev = torch.device("cuda", local)
    Xh, Wh = make_chain(2 + rank)
    Xp = [torch.from_numpy(x).pin_memory() for x in Xh]
    Wp = [torch.from_numpy(w).pin_memory() for w in Wh]
    Xd = [t.to(dev) for t in Xp]
    Wd = [t.to(dev) for t in Wp]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return sw.apply_round_dm(Xd, Wd, CHI)

    def step_e2e(out_host):
        X = MPS.from_sites(Xp)                          # H2D from pinned host memory
        Wm = MPO.from_sites(Wp)
        # D2H of the result cores into pinned host buffers: issued from inside the sweep as each core becomes final (host_out)
        return syn.mul(Wm, X, mode="optimized", bond=CHI, host_out=out_host)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # warm-up (also sizes every cached workspace)
    for _ in range(max(args.warmup, 3)):
        out, trunc = step_device()
    # parity guard inside the bench: the state must be left-canonical and normalisation-consistent
    L0 = out[20].reshape(-1, out[20].shape[-1])
    gram_err = float((ops.matmul(L0.t(), L0) - torch.eye(L0.shape[1], dtype=torch.float64, device=dev)).abs().max().item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.lib.syn_launch_count()
    ms = timed(step_device, args.steps)
    launches = (ops.lib.syn_launch_count() - launches0) / float(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = world * args.steps / (ms * 1e-3)

    # end to end through the public API with host buffers
    out_host = [torch.empty(CHI * D_PHYS * CHI, dtype=torch.float64).pin_memory() for _ in range(N_SITES)]
    step_e2e(out_host)
    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e = timed(lambda: step_e2e(out_host), e2e_steps)
    e2e_value = world * e2e_steps / (ms_e2e * 1e-3)
    h2d = int(sum(t.numel() for t in Xp + Wp) * 8)
    d2h = int(sum(int(np.prod(s.shape)) for s in out) * 8)

    # secondary: the reference-semantic (QR truncation) sweep on the same chain
    def step_qr():
        return sw.apply_round_qr(Xd, Wd, CHI)
    for _ in range(2):
        step_qr()
    ms_qr = timed(step_qr, max(2, min(args.steps, 5)))
    qr_value = world * max(2, min(args.steps, 5)) / (ms_qr * 1e-3)

    # BASELINE configs[3]: 8192 independent random MPS (N=32, d=2, chi=64), overlaps of state b of set A with state b of set B,
    # batch-sharded in contiguous blocks, ONE all-gather of the B scalars; reported as states/s
    from syngular_b200 import parallel
    from syngular_b200.batched import BatchedMatrixProductState as BMPS
    B_total, n4 = 8192, 32
    lo, hi = parallel.shard_bounds(B_total, rank, world)
    bonds4 = capped_bonds(n4, 2, 64)[1:-1]
    A4 = BMPS.random(hi - lo, (2,) * n4, bonds4, seed=1000 + rank, device=dev)
    B4 = BMPS.random(hi - lo, (2,) * n4, bonds4, seed=2000 + rank, device=dev)

    def step_c4():
        return parallel.gather_scalars(A4.overlap(B4), B_total)
    for _ in range(2):
        ov = step_c4()
    c4_reps = 3
    ms_c4 = timed(step_c4, c4_reps)
    c4_value = B_total * c4_reps / (ms_c4 * 1e-3)
    c4_flops = sum(4.0 * 2 * a * b * max(a, b) for a, b in zip([1] + list(bonds4), list(bonds4) + [1])) * B_total
    c4_bytes = 2 * 8.0 * sum(a * 2 * b for a, b in zip([1] + list(bonds4), list(bonds4) + [1])) * B_total
    ov_ok = bool(torch.isfinite(ov).all().item()) and ov.numel() == B_total
    # configs[3] (ii): one shared MPO (chi_W = 4) applied to the states and rounded back to chi = 64, both roundings, on the
    # first 256 states of this rank's shard (the rate is per state; the full shard just repeats it)
    from syngular.tensor import MatrixProductOperator as MPO4
    W4 = MPO4.random_cores((2,) * n4, (2,) * n4, capped_bonds(n4, 4, 4)[1:-1], seed=7).sites
    sub = BMPS([c[:256] for c in A4.sites])
    sub.apply_round(W4, 64); sub.apply_round_svd(W4, 64, chunk=256)
    ms_c4q = timed(lambda: sub.apply_round(W4, 64), 2)
    ms_c4s = timed(lambda: sub.apply_round_svd(W4, 64, chunk=256), 2)
    c4_apply_qr = world * 256 * 2 / (ms_c4q * 1e-3)
    c4_apply_svd = world * 256 * 2 / (ms_c4s * 1e-3)
    del A4, B4, sub
    # BASELINE configs[4]: TensorDense forward, MPO (16,16,16)x(16,16,16) bond 16, batch 65536 x 4096, batch-sharded
    from syngular.layers import TensorDense
    layer5 = TensorDense((16, 16, 16), (16, 16, 16), (16, 16), seed=5, precision="tf32").build()
    lo5, hi5 = parallel.shard_bounds(65536, rank, world)
    g5 = torch.Generator(device=dev).manual_seed(500 + rank)
    x5 = torch.randn((hi5 - lo5, 4096), dtype=torch.float32, device=dev, generator=g5)
    y5 = torch.empty_like(x5)
    step_c5 = lambda: ops.tt_dense3_tf32(x5, layer5._packed, layer5._bias32, relu=True, out=y5)
    step_c5(); step_c5()
    ms_c5 = timed(step_c5, 3) / 3
    c5_value = 65536 / (ms_c5 * 1e-3)
    layer5_f64 = TensorDense((16, 16, 16), (16, 16, 16), (16, 16), seed=5).build()
    x5d = x5[:8192].to(torch.float64)
    layer5_f64(x5d)
    ms_c5_f64 = timed(lambda: layer5_f64(x5d), 1)
    c5_f64_value = world * 8192 / (ms_c5_f64 * 1e-3)
    del x5, x5d, y5

    # BASELINE configs[0]: the README chain (readme.md:53-73) from the reference's decomposed cores (tests/golden), GPU vs the
    # numpy oracle on the host; tiny tensors (d = 16, bonds <= 24): launch-latency territory, reported for completeness
    c1 = None
    if rank == 0:
        try:
            gold = np.load(os.path.join(ROOT, "tests", "golden", "readme_chain.npz"))

            def cores(prefix):
                return [gold["%s/site%d" % (prefix, k)] for k in range(int(gold[prefix + "/n"]))]

            def readme_chain(MPSc, MPOc):
                Wc, Xc, Tc = MPOc.from_sites(cores("W0")), MPSc.from_sites(cores("X0")), MPOc.from_sites(cores("T0"))
                Wc = Wc >> 4
                Tc = Tc >> 2
                Z = ((Tc + Wc) @ Tc) @ Xc
                v = Z | Xc
                Z = Z >> 16
                Z.left_orthonormalization()
                return float(v)
            v_gpu = readme_chain(MPS, MPO)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                readme_chain(MPS, MPO)
            torch.cuda.synchronize()
            gpu_ms = (time.perf_counter() - t0) / 3 * 1e3
            from oracle import ref_numpy as RN
            t0 = time.perf_counter()
            for _ in range(3):
                v_cpu = readme_chain(RN.MPS, RN.MPO)
            cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
            c1 = {"gpu_ms": gpu_ms, "cpu_oracle_ms": cpu_ms, "Z_X_gpu": v_gpu, "Z_X_golden": float(gold["Z_X"]),
                  "rel_err": abs(v_gpu - float(gold["Z_X"])) / abs(float(gold["Z_X"]))}
        except Exception as exc:                     # never let the side measurement break the headline line
            c1 = {"error": repr(exc)}

    # BASELINE configs[2]: 50-qubit brick-wall circuit of Haar-random two-qubit unitaries, depth 20 (seed 3), complex128
    # through the planar compositions on the FP64 kernels.  The default bench line uses chi_max = 64 to stay short; configs[2]'s chi_max = 512
    # is timed by tools/circuit_bench.py.
    c3 = None
    if rank == 0:
        try:
            from syngular.quantum import Circuit
            rng3 = np.random.default_rng(3)

            def haar4():
                z = rng3.normal(size=(4, 4)) + 1j * rng3.normal(size=(4, 4))
                q, r = np.linalg.qr(z)
                return (q * (np.diag(r) / np.abs(np.diag(r)))).reshape(2, 2, 2, 2)
            nq, depth, chi3 = 50, 20, 64
            structure = [(haar4(), i) for layer in range(depth) for i in range(layer % 2, nq - 1, 2)]
            Circuit(8, structure=[(g, i % 7) for g, i in structure[:40]], chi_max=16).run()         # warm-up of every code path
            torch.cuda.synchronize()
            l0 = ops.lib.syn_launch_count()
            t0 = time.perf_counter()
            circ = Circuit(nq, structure=structure, chi_max=chi3)
            circ.run()
            torch.cuda.synchronize()
            sec3 = time.perf_counter() - t0
            st3 = circ.get().state
            c3 = {"gates_per_s": len(structure) / sec3, "seconds_per_circuit": sec3, "gates": len(structure), "qubits": nq, "depth": depth,
                  "chi_max": chi3, "max_bond": int(max(c.shape[2] for c in st3.sites[:-1])), "norm2": float(np.real(st3.conj() | st3)),
                  "kernel_launches": int(ops.lib.syn_launch_count() - l0),
                  "note": "configs[2] at chi_max=64 to keep the default run short (complex128 planar on the real FP64 kernels; saturated bonds "
                          "are cut by the fused spectral-projection kernel on the interleaved embedding, so chi_max=512 runs too: "
                          "tools/circuit_bench.py); norm2 < 1 is the truncation loss"}
        except Exception as exc:
            c3 = {"error": repr(exc)}

    line = None
    if rank == 0:
        # roofline of the dominant kernel: profile one sweep with per-launch CUDA events on the launching stream
        step_device()                                # settle the caching allocator after the side measurements above
        ops.GEMM_PROFILE = []
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); step_device(); t1.record()
        torch.cuda.synchronize()
        prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
        g_ms = sum(a.elapsed_time(b) for a, b, _, _, _ in prof)
        g_fl = sum(f for _, _, f, _, _ in prof)
        g_by = sum(by for _, _, _, by, _ in prof)
        sweep_ms = t0.elapsed_time(t1)
        # FP64 peak measured here: cuBLAS DGEMM 8192^3 (burst, best of 5)
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        best = 1e30
        for _ in range(6):
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(); torch.matmul(a, b); p1.record(); torch.cuda.synchronize()
            best = min(best, p0.elapsed_time(p1))
        del a, b
        peak = 2 * 8192.0 ** 3 / (best * 1e-3) / 1e12
        achieved = g_fl / (g_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "gemm_f64_kernel (DMMA.8x8x4)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, **ncu_traffic(),
                    "peak_source": "cuBLAS DGEMM 8192^3 measured in this process (MEASURED_PEAKS.json has no FP64 entry); DMMA pipe microbench: 37.1",
                    "launches_per_sweep": len(prof), "flops_per_sweep": g_fl, "algorithmic_bytes_per_sweep": g_by,
                    "kernel_ms_per_sweep": g_ms, "share_of_step": g_ms / (ms / args.steps),
                    "profiled_sweep_ms": sweep_ms,
                    "share_note": "kernel time of one extra sweep with CUDA events around every GEMM launch, over the timed ms_per_step"}
        cpu, parity = None, None
        if world == 1 and not args.no_cpu:
            # ONE full sweep of the like-for-like CPU algorithm on the host cores (measured, not scaled), and the GPU result
            # checked against it: per-bond discarded weights, the norm, and the overlap of the two rounded states
            use_all_host_threads()
            sec, cpu_out, cpu_spec, cpu_disc = cpu_full_sweep(Xh, Wh, CHI)
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": host_threads(), "kind": "port",
                   "sample": "ONE complete C2 sweep (all 64 sites) of oracle/svd_numpy.StructuredDensityMatrixSweep -- the same density-matrix "
                             "contraction sequence the GPU runs, numpy + BLAS/LAPACK on all host threads: %.1f s measured, nothing scaled" % sec}
            out_g, trunc_g = step_device()
            _, keep_g, disc_g = trunc_g.host()
            ref = [torch.from_numpy(np.ascontiguousarray(c)).to(dev) for c in cpu_out]
            n_gg = float(sw.overlap(out_g, out_g).item()); n_cc = float(sw.overlap(ref, ref).item()); n_gc = float(sw.overlap(out_g, ref).item())
            parity = {"norm2_rel_diff": abs(n_gg - n_cc) / n_cc, "state_rel_dist2": abs(n_gg + n_cc - 2.0 * n_gc) / n_cc,
                      "discarded_weight_max_abs_diff_over_norm2": float(max(abs(a - b) for a, b in zip(disc_g, cpu_disc)) / n_cc),
                      "kept_ranks_equal": [int(k) for k in keep_g] == [min(CHI, int(np.count_nonzero(sp > 3.2e-7 * sp[0]))) for sp in cpu_spec]}
            del ref, cpu_out
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rounding": "SVD truncation by the density-matrix algorithm (Gram environments; dominant eigenspace of every bond by the fused SP2 spectral-projection kernel, one-sided Jacobi below 256 or without a gap)",
                       "l2": "inputs larger than L2: 63 right environments of up to 134 MB (8.5 GB) are streamed every sweep",
                       "multi_gpu": "replicas only: one independent chain per rank, no collective on the data path",
                       "left_gram_err_site20": gram_err, "parity_vs_cpu_sweep": parity},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "syn.mul(W, X, mode='optimized', bond=256, host_out=pinned buffers) on pinned host cores: inputs uploaded, result cores downloaded while the sweep runs"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            # the batched configs of BASELINE.json ("batched states/s at 1-8 GPU"), whole-job aggregates of this run; `--workload c4 | c5`
            # print them as full lines of their own (e2e, roofline, cpu_baseline)
            "batched": [{"metric": "batched_overlap_states_per_s", "value": c4_value, "unit": "states/s", "n_gpus": world, "scaling": "strong",
                         "workload": "C4: 8192 MPS pairs (N=32, d=2, chi=64), batch-sharded, one all-gather"},
                        {"metric": "tensordense_forward_samples_per_s", "value": c5_value, "unit": "samples/s", "n_gpus": world, "scaling": "strong",
                         "workload": "C5: TensorDense 65536 x 4096 float32 on the fused tcgen05 TF32 kernel, batch-sharded"}],
            "extra": {"c1_readme_chain": c1, "c3_circuit": c3, "c4_batched_overlaps_states_per_s": c4_value, "c4_ms_per_batch": ms_c4 / c4_reps,
                      "c4_note": "BASELINE configs[3]: 8192 x (N=32, d=2, chi=64) overlaps, batch-sharded over %d GPU(s), one all-gather of 8192 "
                                 "float64; %.2f TFLOP/s FP64, %.0f GB/s of core traffic; gathered %s" % (
                                     world, c4_flops / (ms_c4 / c4_reps * 1e-3) / 1e12, c4_bytes / (ms_c4 / c4_reps * 1e-3) / 1e9,
                                     "ok" if ov_ok else "BAD"),
                      "c5_tensordense_samples_per_s": c5_value,
                      "c5_note": "configs[4] forward, float32 in/out on the fused tcgen05.mma.kind::tf32 kernel (%.0f TFLOP/s; `--workload c5` prints the "
                                 "full line); the FP64 DMMA-GEMM path of the same layer: %.0f samples/s" % (3.7748736e7 * 65536 / (ms_c5 * 1e-3) / 1e12, c5_f64_value),
                      "c4_apply_qr_round_states_per_s": c4_apply_qr, "c4_apply_svd_round_states_per_s": c4_apply_svd,
                      "c4_apply_note": "configs[3](ii): shared MPO chi_W=4 applied + rounded to chi=64, batched over 256 states per rank; SVD rounding by the batched projection kernel (one CTA per state and bond, csrc/purify_batched.cu), Jacobi for the members it rejects",
                      "qr_round_sweeps_per_s": qr_value, "qr_round_ms_per_sweep": 1e3 / (qr_value / world),
                      "qr_round_note": "reference-semantic `>>` (QR truncation, fused apply+round) on the same chain"},
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


_REAL_STDOUT = None


def protect_stdout():
    """Everything except the final JSON line goes to stderr: libraries (NCCL's version banner, cuBLAS warnings) write to fd 1."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(json.dumps(line), flush=True)
    else:
        os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 (default, the headline): single-chain apply + SVD-round; c3: 50-qubit circuit (gates/s, one circuit per step); "
                         "c4: batched overlaps (states/s); c5: TensorDense forward (samples/s)")
    ap.add_argument("--chi-max", type=int, default=512, help="bond cap of the c3 circuit (BASELINE configs[2]: 512)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c3":
        run_c3(args)
    elif args.workload == "c4":
        run_c4(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
