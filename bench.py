#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path (see the contract in DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload n16384|n4096|n32768|batched64|n65536|fp64_n4096]
    python bench.py --impl reference ...      # the reference's own CPU path on the host cores

A "step" is one pass of the hot path over one batch of synthetic input:
  n16384 / n4096  one FP32 inversion of the named order (BASELINE.json configs[2] / configs[1]);
                  with N > 1 GPUs every rank inverts its own matrix (split by matrix index, no
                  collective: "scaling": "weak")
  batched64       2^20 (per GPU: 2^20 / 8 ... see --batch) 64x64 inversions, split by index (configs[3])
  n65536          one inversion column-sharded over the ranks (configs[4], strong scaling)

`value` is whole-job GFLOP/s counting 2 N^3 flops per inversion with inputs resident in HBM;
`e2e` is the same metric through the host-pointer C-ABI call (pinned host buffers, H2D + D2H inside
the timed region).  The oracle / numpy legs (`cpu_baseline`, --impl reference) are the only places
this file touches oracle/ or CPU math, and only as the thing reported next to the GPU number.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "GFLOP/s (2N^3/t) and fraction of FP32 peak at N=16384; batched 64x64 inv/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="n16384", choices=["n16384", "n4096", "n32768", "batched64", "n65536", "fp64_n4096"])
    ap.add_argument("--kind", default="uniform", choices=["uniform", "diagdom"])
    ap.add_argument("--batch", type=int, default=1 << 20, help="batched64: matrices per GPU")
    ap.add_argument("--order", type=int, default=0, help="n65536: override the order (e.g. 16384 for a quick sharded run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="default workload: skip the batched64 / sharded_n65536 objects")
    ap.add_argument("--sharded-order", type=int, default=65536, help="order of the sharded_n65536 object (multi-GPU default line)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks sampler

class ClockSampler:
    """nvidia-smi poller running DURING the timed region (B200_PROFILING.md 'clocks' line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                 "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- reference CPU path

_REF_PY = Path("/root/reference/matrix_inv_numpy.py")
_ref_just_inv = None
if _REF_PY.exists():   # this container: time the reference's OWN function; the GPU box has no /root/reference
    try:
        import importlib.util

        _spec = importlib.util.spec_from_file_location("reference_matrix_inv_numpy", str(_REF_PY))
        _mod = importlib.util.module_from_spec(_spec)
        _spec.loader.exec_module(_mod)
        _ref_just_inv = _mod.just_inv
    except Exception:
        _ref_just_inv = None
REF_KIND = "reference" if _ref_just_inv is not None else "port"


def just_inv(K: int, dtype):
    """The reference's CPU path, /root/reference/matrix_inv_numpy.py:39-46: U(0,100) K x K matrix,
    np.linalg.inv(np.matrix(a)), timed with time.monotonic() around the inversion only.  When the reference file is
    present (this container) its own `just_inv` runs and the time it prints is parsed (kind "reference"); on the GPU box
    the file does not exist and the five lines are restated here (kind "port").  The float32 twin is always the port."""
    import numpy as np

    if _ref_just_inv is not None and dtype == np.float64:
        import contextlib
        import io

        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            _ref_just_inv(K)
        return float(buf.getvalue().split("TIME:")[1].split()[0])
    a = np.random.uniform(0, 100, (K, K)).astype(dtype)
    start = time.monotonic()
    res = np.linalg.inv(np.matrix(a))
    end = time.monotonic()
    assert res.shape == (K, K)
    return end - start


def pick_reference_order(n_target: int, budget_s: float):
    """Largest K in {n_target, n_target/2, ...} whose np.linalg.inv is projected to fit the budget."""
    import numpy as np

    t = just_inv(2048, np.float64)
    gflops = 2 * 2048 ** 3 / t / 1e9
    K = n_target
    while K > 2048 and 2 * K ** 3 / (gflops * 1e9) > budget_s:
        K //= 2
    return K, gflops


def cpu_baseline(n_target: int):
    """The reference's CPU path (numpy LAPACK getrf+getri, matrix_inv_numpy.py) and the oracle port
    (OpenMP Gauss-Jordan replay) on bounded samples, timed on the host cores."""
    import numpy as np

    from oracle import gj_oracle as o

    cores = os.cpu_count() or 1
    K, _ = pick_reference_order(n_target, 20.0)
    t64 = just_inv(K, np.float64)
    t32 = just_inv(K, np.float32)
    best = min(t64, t32)
    out = {"value": 2 * K ** 3 / best / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": REF_KIND,
           "sample": f"matrix_inv_numpy.py just_inv semantics: np.linalg.inv(np.matrix(U(0,100))) at N={K} "
                     f"(float64 {t64:.2f}s as written, float32 twin {t32:.2f}s; value = the faster); "
                     f"{'full order' if K == n_target else f'bounded sample of the N={n_target} workload'}"}
    Kp = 2048
    A = o.uniform(Kp)
    t0 = time.monotonic()
    X, piv, info = o.invert_inplace(A)
    tp = time.monotonic() - t0
    out["port"] = {"value": 2 * Kp ** 3 / tp / 1e9, "unit": "GFLOP/s", "cores": o.threads(), "kind": "port",
                   "sample": f"oracle gj_inplace_f32 (OpenMP Gauss-Jordan replay) at N={Kp}, {tp:.2f}s"}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (numpy.linalg.inv exactly
    as matrix_inv_numpy.py does it) on all host cores; each step one inversion of a bounded order."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core (numpy is not imported yet)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(os.cpu_count() or 1)
    import numpy as np

    n_target = {"n16384": 16384, "n4096": 4096, "n32768": 32768, "n65536": 65536, "batched64": 64, "fp64_n4096": 4096}[args.workload]
    cores = os.cpu_count() or 1
    if args.workload == "batched64":
        B = 65536
        a = np.random.default_rng(0).uniform(0, 100, (B, 64, 64)).astype(np.float32)
        for _ in range(args.warmup):
            np.linalg.inv(a)
        t0 = time.monotonic()
        for _ in range(args.steps):
            np.linalg.inv(a)
        dt = (time.monotonic() - t0) / args.steps
        value, unit = B / dt, "inv/s"
        sample = f"np.linalg.inv on a ({B},64,64) float32 stack per step (1/16 of the 2^20 batch)"
        K = 64
    else:
        # The metric's own order when the timed steps fit ~15 minutes (the headline N=1 run; well inside the driver's
        # 1800 s limit), else -- and for the N>1 runs of the scaling sweep, which only repeat it -- a bounded order.  Warm-up
        # steps of a CPU path only have to load LAPACK and spin up its threads: they run at N=2048.
        budget = 900.0 if args.gpus <= 1 else 120.0
        K, _ = pick_reference_order(n_target, budget / max(args.steps, 1))
        for _ in range(args.warmup):
            just_inv(min(K, 2048), np.float64)
        ts = [just_inv(K, np.float64) for _ in range(args.steps)]
        dt = sum(ts) / len(ts)
        value, unit = 2 * K ** 3 / dt / 1e9, "GFLOP/s"
        sample = (f"matrix_inv_numpy.py just_inv (np.linalg.inv(np.matrix(U(0,100) float64))) at N={K} per step"
                  + ("" if K == n_target else f"; bounded sample of the N={n_target} workload")
                  + f"; warm-up steps at N={min(K, 2048)}; just_inv {'imported from the reference' if REF_KIND == 'reference' else 'restated (reference file absent on this box)'}")
    line = {"metric": METRIC, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if args.workload != "batched64" else "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": (f"batched {1 << 20} x 64x64 FP32 inversions per GPU, split by matrix index" if args.workload == "batched64"
                                    else f"N={n_target} random-{args.kind} FP32 single inversion per GPU, partial pivoting"),
                       "n": n_target, "same_order_as_metric": bool(args.workload != "batched64" and K == n_target),
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": REF_KIND, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- column-sharded run (torchrun ranks)

_BCAST_GROUP = None   # (group,) once created


def make_bcast_group(dist, world):
    """NCCL process group for the per-block panel broadcasts, limited to MATINV_SHARD_BCAST_CTAS (default 4) CTAs and on a
    high-priority stream.  A receiver posts the broadcast of panel J+1 a whole block step ahead and its kernel spins until
    the owner has factored the panel: every CTA it holds is an SM slot the trailing GEMM does not get (measured with the
    C-ABI driver on 8 GPUs, N=65536: 1493 ms with NCCL's default configuration, 1375 ms with 2 CTAs, 1314 ms with 4, 1738 ms
    with 1 -- then the broadcast is too slow to hide behind the GEMM)."""
    if world <= 1:
        return None
    ctas = int(os.environ.get("MATINV_SHARD_BCAST_CTAS", "4"))
    if ctas <= 0:
        return None
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        opts.config.min_ctas = 1
        opts.config.max_ctas = ctas
        return dist.new_group(ranks=list(range(world)), pg_options=opts)
    except Exception:
        return None


def run_sharded_workload(m, torch, dist, n, kind, seed, K, W, rank, world, local, check_single_gpu):
    """K timed column-sharded inversions of one n x n matrix over all ranks (owner factors -> NCCL broadcast -> all apply;
    one all-to-all for the deferred column permutation), CUDA events, max over ranks.  With check_single_gpu every rank
    then inverts the SAME matrix alone on its own GPU (the 1-GPU base of the strong-scaling figure, timed in this job) and
    compares its column blocks of the sharded result and the pivot sequence bit for bit; the O(n^2) probe gives the
    residual.  Raises on a mismatch."""
    from gpu_matrix_inversion_b200.sharded import BLOCK, CudaShardBackend, ShardedInverter

    dev = torch.device("cuda", local)
    backend = CudaShardBackend(n, rank, world, dev)
    global _BCAST_GROUP
    if world > 1 and _BCAST_GROUP is None:
        _BCAST_GROUP = (make_bcast_group(dist, world),)
    bgroup = _BCAST_GROUP[0] if _BCAST_GROUP else None
    inv = ShardedInverter(backend, dist if world > 1 else None, bcast_group=bgroup)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def one_step():
        backend.generate(seed, kind)      # the factorisation is in place: regenerate (cheap, on device)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        e0.record()
        info, piv = inv.factorize()
        out = inv.exchange_columns(piv)
        e1.record()
        torch.cuda.synchronize()
        assert info == 0
        return e0.elapsed_time(e1), piv, out

    for _ in range(W):
        one_step()
    times = []
    for _ in range(K):
        t, piv, out = one_step()
        times.append(t)
    ms = max_over_ranks(sum(times) / len(times))
    res = {"n": n, "ms_per_step": ms, "value": 2.0 * n ** 3 / (ms * 1e-3) / 1e9, "unit": "GFLOP/s", "steps": K, "warmup": W,
           "n_gpus": world, "scaling": "strong",
           "comm": {"collective": "ncclBroadcast (torch.distributed, high-priority stream"
                                  + (", process group limited to %s CTAs" % os.environ.get("MATINV_SHARD_BCAST_CTAS", "4") if bgroup is not None else "")
                                  + ") of the factored panel per 128-column block + one all_to_all for the deferred column permutation",
                    "bytes_per_block_step": backend.msg_bytes, "block_steps": (n + BLOCK - 1) // BLOCK,
                    "all_to_all_bytes_per_rank": 4 * n * len(backend.blocks) * BLOCK}}
    if check_single_gpu:
        A = m.generate_dev(n, seed, kind)
        Xs = torch.empty_like(A)
        pivs = torch.empty(n, dtype=torch.int32, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rc, _ = m.invert_dev(A, Xs, piv=pivs)          # untimed: sizes the workspace for this order
        assert rc == m.OK, m.last_error()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        rc, _ = m.invert_dev(A, Xs, piv=pivs)
        e1.record()
        torch.cuda.synchronize()
        assert rc == m.OK, m.last_error()
        ms1 = max_over_ranks(e0.elapsed_time(e1))
        same = bool(torch.equal(pivs.cpu(), torch.from_numpy(piv)))
        for J, blk in out.items():
            same = same and bool(torch.equal(blk.contiguous().view(torch.int32), Xs[:, J * BLOCK:J * BLOCK + blk.shape[1]].contiguous().view(torch.int32)))
        flag = torch.tensor([1.0 if same else 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same_all = bool(flag.item() == 1.0)
        est = m.probe_residual_dev(A, Xs)
        res.update({"t_1gpu_ms": ms1, "speedup_vs_1gpu": ms1 / ms, "bitwise_equal_single_gpu": same_all, "residual": est,
                    "residual_note": "O(n^2) randomised estimate of ||AX-I||_F/(n||A||_F||X||_F) on the single-GPU inverse every "
                                     "rank computed in this job; the sharded result's column blocks and pivot sequence equal it "
                                     "bit for bit on every rank, so it is the sharded result's residual too",
                    "t_1gpu_note": "same matrix inverted alone on every GPU (max over ranks), timed in this job"})
        del A, Xs
        if not same_all:
            raise SystemExit(f"sharded N={n}: result differs from the single-GPU result (rank {rank})")
        if not (est <= 1e-5):
            raise SystemExit(f"sharded N={n}: residual estimate {est} above 1e-5")
    del out
    backend.close()
    torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------- our arm

def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import gpu_matrix_inversion_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or m.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # high-priority NCCL streams: with look-ahead the panel broadcast must not queue behind the trailing GEMM
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from oracle.gj_oracle import SEED_BATCHED, SEED_DIAGDOM, SEED_UNIFORM  # constants only

    sampler = ClockSampler(local)
    K, W = args.steps, max(args.warmup, 0)
    extra = {}

    if args.workload in ("n16384", "n4096", "n32768"):
        n = {"n16384": 16384, "n4096": 4096, "n32768": 32768}[args.workload]
        seed = (SEED_UNIFORM if args.kind == "uniform" else SEED_DIAGDOM) + n + rank * 7919
        A = m.generate_dev(n, seed, args.kind)
        X = torch.empty_like(A)
        for _ in range(W):
            rc, _ = m.invert_dev(A, X)
            assert rc == m.OK, m.last_error()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m.profile_enable(True)
        barrier()
        sampler.start()
        ev0.record()
        for _ in range(K):
            rc, _ = m.invert_dev(A, X)
        ev1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        barrier()
        prof = m.profile_read()
        m.profile_enable(False)
        assert rc == m.OK
        ms = max_over_ranks(ev0.elapsed_time(ev1) / K)
        flops = 2.0 * n ** 3
        value = world * flops / (ms * 1e-3) / 1e9
        res, _ = m.residual_dev(A, X) if n <= 16384 else (None, None)
        extra["residual"] = res

        # e2e: same metric through the host-pointer C-ABI entry, pinned host buffers, H2D + D2H inside
        Ah = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
        Xh = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
        Ah.copy_(A)
        torch.cuda.synchronize()
        for _ in range(1):
            rc = m.lib.matinv_invert_f32(Ah.data_ptr(), n, Xh.data_ptr(), None, 0)
            assert rc == m.OK, m.last_error()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            rc = m.lib.matinv_invert_f32(Ah.data_ptr(), n, Xh.data_ptr(), None, 0)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / K)
        assert rc == m.OK
        # the host entry uploads A in column windows while it already factors (pipelined schedule): same bits as the device entry
        Xchk = torch.empty_like(X)
        Xchk.copy_(Xh)
        e2e_same = bool(torch.equal(Xchk.view(torch.int32), X.view(torch.int32)))
        del Xchk
        assert e2e_same, "host-entry result differs from the device-entry result"
        e2e = {"value": world * flops / e2e_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": 4 * n * n,
               "bitwise_equal_device_result": e2e_same,
               "d2h_bytes_per_step": 4 * n * n, "ms_per_step": e2e_s * 1e3,
               "phases_s": m.last_phases(),   # last step: setup / H2D / factorisation / extraction + D2H / total (matinv_last_phases)
               "api": "matinv_invert_f32 (what matrix_inv_32 calls), pinned host buffers; upload pipelined in column windows "
                      "(MATINV_H2D_PIPELINE, default on for n >= 8192)"}

        # optional 3xTF32 tcgen05 trailing update (north_star config 3: "FP32 SIMT vs 3xTF32 tcgen05"), same input, same K;
        # the residual gate (O(N^2) probe + status read-back) is inside the timed region.  Its dominant kernel is HBM-bound.
        for _ in range(2):
            rc, _ = m.invert_dev(A, X, flags=m.FLAG_TF32X3)
            assert rc == m.OK, m.last_error()
        m.profile_enable(True)
        barrier()
        ev0.record()
        for _ in range(K):
            rc, _ = m.invert_dev(A, X, flags=m.FLAG_TF32X3)
        ev1.record()
        torch.cuda.synchronize()
        barrier()
        prof_tc = m.profile_read()
        m.profile_enable(False)
        st_tc = m.tf32x3_status()
        ms_tc = max_over_ranks(ev0.elapsed_time(ev1) / K)
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
        hbm = peaks.get("hbm_gbs", 6650.0)
        tc_ms = prof_tc["gemm_ms"] / max(prof_tc["gemm_launches"], 1)
        tc_bytes = prof_tc["gemm_flops"] / max(prof_tc["gemm_launches"], 1) / 32.0   # 8 B (read + write) per 2*128 flops
        tc_gbs = tc_bytes / (tc_ms * 1e-3) / 1e9 if tc_ms else 0.0
        tc_traffic = None
        ttf = ROOT / "profiles" / "r01_tf32x3_traffic.json"
        if n == 16384 and ttf.exists():   # DRAM bytes of one strip-kernel launch, from the committed ncu --set full capture
            tc_traffic = json.loads(ttf.read_text())["dram_bytes_total"]
        extra["tf32x3"] = {
            "value": world * flops / (ms_tc * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": ms_tc,
            "speedup_vs_fp32_simt": ms / ms_tc, "residual_estimate": st_tc["estimate"], "fell_back": st_tc["fell_back"],
            "residual": m.residual_dev(A, X)[0] if n <= 16384 else None, "gate": m.TF32X3_GATE,
            "gpu_launches": prof_tc["launches"],
            "roofline": {"bound": "hbm", "kernel": "tf32_split_kernel + trailing_tf32x3_strip_kernel (tcgen05.mma kind::tf32, TMEM)",
                         "achieved": tc_gbs, "peak": hbm, "unit": "GB/s", "frac": tc_gbs / hbm, "traffic": tc_traffic,
                         "traffic_note": "static: DRAM bytes per launch from the committed round-1 ncu --set full capture (profiles/r01_tf32x3_traffic.json), not re-measured in this run",
                         "algorithmic_bytes_per_launch": tc_bytes, "ms_per_launch": tc_ms,
                         "tensor_tflops_3x": 3.0 * prof_tc["gemm_flops"] / max(prof_tc["gemm_launches"], 1) / (tc_ms * 1e-3) / 1e12 if tc_ms else 0.0,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                         "kernel_share_of_step": prof_tc["gemm_ms"] / K / ms_tc},
            "note": "not bit-identical to the reference's FMA chain: accepted per inversion by the residual gate, else the "
                    "FP32 SIMT schedule is rerun (fell_back)"}

        peak = m.ffma_peak_tflops()
        gemm_ms = prof["gemm_ms"] / max(prof["gemm_launches"], 1)
        gemm_tflops = prof["gemm_flops"] / max(prof["gemm_launches"], 1) / (gemm_ms * 1e-3) / 1e12 if gemm_ms else 0.0
        traffic = None
        tf = ROOT / "profiles" / "r01_gemm_traffic.json"
        if n == 16384 and tf.exists():   # DRAM bytes of one trailing-update launch, from the committed ncu --set full capture
            traffic = json.loads(tf.read_text())["dram_bytes_total"]
        roofline = {"bound": "fp32_simt", "kernel": "trailing_gemm_kernel", "achieved": gemm_tflops, "peak": peak,
                    "unit": "TFLOP/s", "frac": gemm_tflops / peak if peak else None, "traffic": traffic,
                    "traffic_note": "static: DRAM bytes per launch from the committed round-1 ncu --set full capture (profiles/r01_gemm_traffic.json), not re-measured in this run; algorithmic bytes 2.098e9",
                    "peak_source": "measured live: matinv_ffma_peak_tflops (FFMA register-tile probe); "
                                   "MEASURED_PEAKS.json has no FP32 SIMT entry",
                    "kernel_share_of_step": prof["gemm_ms"] / K / ms,
                    "whole_inversion_frac_of_peak": (flops / (ms * 1e-3) / 1e12) / peak if peak else None,
                    "nominal_fp32_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12}
        launches = prof["launches"]
        if args.workload == "n16384" and not args.no_extras:
            # ---- BASELINE config 4 on the same line: 2^20 x 64x64 per GPU, split by index (no collective)
            del Ah, Xh
            bn, bb = 64, args.batch
            Ab = m.generate_batched_dev(bn, rank * bb, bb, SEED_BATCHED)
            Xb = torch.empty_like(Ab)
            ib = torch.empty(bb, dtype=torch.int32, device="cuda")
            for _ in range(2):
                m.invert_batched_dev(Ab, Xb, ib)
            barrier()
            bK = max(3, min(K, 10))
            ev0.record()
            for _ in range(bK):
                m.invert_batched_dev(Ab, Xb, ib)
            ev1.record()
            torch.cuda.synchronize()
            assert int((ib != 0).sum()) == 0
            bms = max_over_ranks(ev0.elapsed_time(ev1) / bK)
            gbs = bb * 2 * bn * bn * 4 / (bms * 1e-3) / 1e9
            extra["batched64"] = {
                "value": world * bb / (bms * 1e-3), "unit": "inv/s", "ms_per_step": bms, "steps": bK, "batch_per_gpu": bb, "n": bn,
                "scaling": "weak", "parallelism": f"index_split_x{world}",
                "roofline": {"bound": "hbm", "kernel": "batched_pk_kernel (one warp per matrix, register resident)", "achieved": gbs,
                             "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                             "algorithmic_bytes_per_matrix": 2 * bn * bn * 4,
                             "fp32_frac": (bb * 2.0 * bn ** 3 / (bms * 1e-3) / 1e12) / peak if peak else None, "fp32_peak_tflops": peak,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                             "note": "AI = 16 flop/B sits above the FP32 ridge: the FP32-bound ceiling is ~72 % of HBM peak (SURVEY Appendix C)"}}
            launches += bK
            del Ab, Xb, ib
            torch.cuda.empty_cache()
            if world > 1:
                # ---- BASELINE config 5 on the same line: one N=65536 matrix column-sharded over all ranks (strong scaling)
                del A, X
                torch.cuda.empty_cache()
                ns = args.sharded_order
                extra["sharded_n%d" % ns] = run_sharded_workload(m, torch, dist, ns, args.kind, (SEED_UNIFORM if args.kind == "uniform" else SEED_DIAGDOM) + ns,
                                                                  2, 1, rank, world, local, check_single_gpu=True)
        config = {"workload": f"N={n} random-{args.kind} FP32 single inversion per GPU, partial pivoting",
                  "n": n, "nb": 128, "l2": "inputs_larger_than_l2" if n >= 8192 else "l2_resident_input",
                  "parallelism": f"replicas_x{world}" if world > 1 else "single_gpu"}
        unit, scaling = "GFLOP/s", "weak"
    elif args.workload == "fp64_n4096":
        # FP64 entry point (matrix_inversion_FP64), blocked schedule of gj_f64.cu
        n = 4096
        gen = torch.Generator(device="cuda").manual_seed(0xB2006400 + rank)
        A = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen) * 100.0
        X = torch.empty_like(A)
        for _ in range(W):
            rc, _ = m.invert_f64_dev(A, X)
            assert rc == m.OK, m.last_error()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.start()
        ev0.record()
        for _ in range(K):
            rc, _ = m.invert_f64_dev(A, X)
        ev1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        barrier()
        assert rc == m.OK
        # kernel-level pass: one more inversion with an event pair around each trailing update
        m.profile_enable(True)
        rc, _ = m.invert_f64_dev(A, X)
        torch.cuda.synchronize()
        prof = m.profile_read()
        m.profile_enable(False)
        ms = max_over_ranks(ev0.elapsed_time(ev1) / K)
        flops = 2.0 * n ** 3
        value = world * flops / (ms * 1e-3) / 1e9
        R = A @ X - torch.eye(n, dtype=torch.float64, device="cuda")
        extra["residual"] = float(R.norm() / (n * A.norm() * X.norm()))
        Ah = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        Xh = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        Ah.copy_(A)
        torch.cuda.synchronize()
        rc = m.lib.matinv_invert_f64(Ah.data_ptr(), n, Xh.data_ptr(), None, 0)
        assert rc == m.OK, m.last_error()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            rc = m.lib.matinv_invert_f64(Ah.data_ptr(), n, Xh.data_ptr(), None, 0)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / K)
        e2e = {"value": world * flops / e2e_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": 8 * n * n,
               "d2h_bytes_per_step": 8 * n * n, "ms_per_step": e2e_s * 1e3,
               "api": "matinv_invert_f64 (what matrix_inversion_FP64 calls), pinned host buffers"}
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
        hbm = peaks.get("hbm_gbs", 6650.0)
        # the hooks bracket the trailing updates of the blocked schedule: per launch (n-64)^2 doubles read + written and
        # 2 (n-64)^2 64 flop; both ceilings are reported because at 64-wide panels the kernel sits near the ridge
        k_ms = prof["gemm_ms"] / max(prof["gemm_launches"], 1)
        mrows = float(n - 64)
        gbs = 16.0 * mrows * mrows / (k_ms * 1e-3) / 1e9 if k_ms else 0.0
        tfl = 2.0 * mrows * mrows * 64 / (k_ms * 1e-3) / 1e12 if k_ms else 0.0
        fp64_nominal = 148 * 64 * 2 * 1.965e9 / 1e12
        roofline = {"bound": "hbm", "kernel": "trailing_f64_kernel", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                    "frac": gbs / hbm, "traffic": None,
                    "traffic_note": "algorithmic bytes per launch 16 (n-64)^2 = 2.60e8 (the 134 MB matrix is about the size of "
                                    "L2, so part of it is served from L2 at this order)",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                    "fp64_tflops": tfl, "fp64_frac_of_nominal": tfl / fp64_nominal, "fp64_nominal_peak_tflops": fp64_nominal,
                    "kernel_share_of_step": prof["gemm_ms"] / ms}
        launches = prof["launches"] * K
        config = {"workload": f"N={n} U(0,100) FP64 single inversion per GPU, partial pivoting (matrix_inversion_FP64)",
                  "n": n, "l2": "input_about_l2_size", "parallelism": f"replicas_x{world}" if world > 1 else "single_gpu",
                  "path": "blocked, 64-column panels (two launches per column + recurrence + trailing update per panel)"}
        unit, scaling = "GFLOP/s", "weak"
    elif args.workload == "batched64":
        n, batch = 64, args.batch
        A = m.generate_batched_dev(n, rank * batch, batch, SEED_BATCHED)
        X = torch.empty_like(A)
        info = torch.empty(batch, dtype=torch.int32, device="cuda")
        for _ in range(W):
            m.invert_batched_dev(A, X, info)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m.profile_enable(True)
        barrier()
        sampler.start()
        ev0.record()
        for _ in range(K):
            m.invert_batched_dev(A, X, info)
        ev1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        barrier()
        launches = m.profile_read()["launches"]
        m.profile_enable(False)
        assert int((info != 0).sum()) == 0
        ms = max_over_ranks(ev0.elapsed_time(ev1) / K)
        value = world * batch / (ms * 1e-3)
        nb_e2e = min(batch, 1 << 17)
        Ah = torch.empty((nb_e2e, n, n), dtype=torch.float32, pin_memory=True)
        Xh = torch.empty_like(Ah).pin_memory()
        Ah.copy_(A[:nb_e2e])
        ih = torch.empty(nb_e2e, dtype=torch.int32)
        m.lib.matinv_invert_batched_f32(Ah.data_ptr(), n, nb_e2e, Xh.data_ptr(), ih.data_ptr(), 0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            m.lib.matinv_invert_batched_f32(Ah.data_ptr(), n, nb_e2e, Xh.data_ptr(), ih.data_ptr(), 0)
        e2e_s = max_over_ranks((time.perf_counter() - t0) / K)
        e2e = {"value": world * nb_e2e / e2e_s, "unit": "inv/s", "h2d_bytes_per_step": 4 * n * n * nb_e2e,
               "d2h_bytes_per_step": 4 * n * n * nb_e2e + 4 * nb_e2e, "batch": nb_e2e}
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
        hbm = peaks.get("hbm_gbs", 6650.0)
        gbs = batch * 2 * n * n * 4 / (ms * 1e-3) / 1e9
        peak32 = m.ffma_peak_tflops()
        roofline = {"bound": "hbm", "kernel": "batched kernel", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                    "frac": gbs / hbm, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s",
                    "fp32_frac": (batch * 2.0 * n ** 3 / (ms * 1e-3) / 1e12) / peak32 if peak32 else None,
                    "fp32_peak_tflops": peak32}
        config = {"workload": f"batched {batch} x 64x64 FP32 inversions per GPU, split by matrix index", "n": 64,
                  "batch_per_gpu": batch, "l2": "inputs_larger_than_l2", "parallelism": f"index_split_x{world}"}
        unit, scaling = "inv/s", "weak"
    else:
        # ---- one matrix column-sharded over the ranks (strong scaling): owner factors -> NCCL broadcast -> all apply
        n = args.order or 65536
        seed = (SEED_UNIFORM if args.kind == "uniform" else SEED_DIAGDOM) + n
        barrier()
        sampler.start()
        sres = run_sharded_workload(m, torch, dist, n, args.kind, seed, K, min(W, 1), rank, world, local, check_single_gpu=True)
        clocks = sampler.stop()
        barrier()
        ms, value = sres["ms_per_step"], sres["value"]
        extra["sharded"] = {k: sres[k] for k in ("t_1gpu_ms", "speedup_vs_1gpu", "bitwise_equal_single_gpu", "residual", "residual_note", "comm")}
        peak = m.ffma_peak_tflops()
        nblk = (n + 127) // 128
        launches = (nblk // max(world, 1)) * 16 + nblk * 4
        e2e = {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * n + 4,
               "note": "input generated on the device (16 GiB at N=65536 does not round-trip the host); "
                       "status word + pivots are read back every step; the host-buffer form of this path is the C-ABI entry "
                       "matinv_invert_sharded_f32 (tests/test_gpu_multi.py)"}
        roofline = {"bound": "fp32_simt", "kernel": "trailing_gemm_kernel", "achieved": value / 1e3 / world, "peak": peak,
                    "unit": "TFLOP/s", "frac": value / 1e3 / world / peak if peak else None, "traffic": None,
                    "peak_source": "measured live: matinv_ffma_peak_tflops; achieved = whole-job rate per GPU"}
        config = {"workload": f"N={n} random-{args.kind} FP32 single inversion, 1-D block-cyclic column sharding (128)",
                  "n": n, "nb": 128, "l2": "inputs_larger_than_l2", "parallelism": f"column_shards_x{world}",
                  "exchange": "NCCL broadcast of the factored panel per block step + one all-to-all for the column permutation"}
        unit, scaling = "GFLOP/s", "strong"

    line = {"metric": METRIC, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64" if args.workload == "fp64_n4096" else "f32",
            "data": "synthetic", "config": config, "roofline": roofline, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks}
    line.update(extra)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(4096 if args.workload in ("n4096", "fp64_n4096") else 16384)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
