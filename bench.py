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

def just_inv(K: int, dtype):
    """/root/reference/matrix_inv_numpy.py:39-46 restated (it cannot be imported on the GPU box):
    U(0,100) K x K matrix, np.linalg.inv(np.matrix(a)), timed with time.monotonic()."""
    import numpy as np

    a = np.random.default_rng(K).uniform(0, 100, (K, K)).astype(dtype)
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
    out = {"value": 2 * K ** 3 / best / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "reference",
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
        flops_per_step = None
    else:
        K, _ = pick_reference_order(n_target, 15.0)
        for _ in range(args.warmup):
            just_inv(K, np.float64)
        ts = [just_inv(K, np.float64) for _ in range(args.steps)]
        dt = sum(ts) / len(ts)
        value, unit = 2 * K ** 3 / dt / 1e9, "GFLOP/s"
        sample = (f"matrix_inv_numpy.py just_inv (np.linalg.inv(np.matrix(U(0,100) float64))) at N={K} per step"
                  + ("" if K == n_target else f"; bounded sample of the N={n_target} workload"))
    line = {"metric": METRIC, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if args.workload != "batched64" else "f32", "data": "synthetic",
            "impl": "reference", "config": {"workload": args.workload, "sample": sample},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
        e2e = {"value": world * flops / e2e_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": 4 * n * n,
               "d2h_bytes_per_step": 4 * n * n, "ms_per_step": e2e_s * 1e3,
               "api": "matinv_invert_f32 (what matrix_inv_32 calls), pinned host buffers"}

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
                         "traffic_note": "DRAM bytes per launch (ncu, profiles/r01_tf32x3_traffic.json)",
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
                    "traffic_note": "DRAM bytes per launch (ncu, profiles/r01_gemm_traffic.json); algorithmic bytes 2.098e9",
                    "peak_source": "measured live: matinv_ffma_peak_tflops (FFMA register-tile probe); "
                                   "MEASURED_PEAKS.json has no FP32 SIMT entry",
                    "kernel_share_of_step": prof["gemm_ms"] / K / ms,
                    "whole_inversion_frac_of_peak": (flops / (ms * 1e-3) / 1e12) / peak if peak else None,
                    "nominal_fp32_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12}
        launches = prof["launches"]
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
        from gpu_matrix_inversion_b200.sharded import CudaShardBackend, ShardedInverter

        n = args.order or 65536
        dev = torch.device("cuda", local)
        backend = CudaShardBackend(n, rank, world, dev)
        inv = ShardedInverter(backend, dist if world > 1 else None)
        seed = (SEED_UNIFORM if args.kind == "uniform" else SEED_DIAGDOM) + n

        def one_step():
            backend.generate(seed, args.kind)      # the factorisation is in place: regenerate (cheap, on device)
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
            return e0.elapsed_time(e1), out

        for _ in range(min(W, 1)):
            one_step()
        barrier()
        sampler.start()
        times = []
        for _ in range(K):
            t, out = one_step()
            times.append(t)
        clocks = sampler.stop()
        barrier()
        del out
        ms = max_over_ranks(sum(times) / len(times))
        flops = 2.0 * n ** 3
        value = flops / (ms * 1e-3) / 1e9
        peak = m.ffma_peak_tflops()
        nblk = (n + 127) // 128
        launches = (nblk // max(world, 1)) * 16 + nblk * 4
        e2e = {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * n + 4,
               "note": "input generated on the device (16 GiB at N=65536 does not round-trip the host); "
                       "status word + pivots are read back every step"}
        roofline = {"bound": "fp32_simt", "kernel": "trailing_gemm_kernel", "achieved": value / 1e3 / world, "peak": peak,
                    "unit": "TFLOP/s", "frac": value / 1e3 / world / peak if peak else None, "traffic": None,
                    "peak_source": "measured live: matinv_ffma_peak_tflops; achieved = whole-job rate per GPU"}
        config = {"workload": f"N={n} random-{args.kind} FP32 single inversion, 1-D block-cyclic column sharding (128)",
                  "n": n, "nb": 128, "l2": "inputs_larger_than_l2", "parallelism": f"column_shards_x{world}",
                  "exchange": "NCCL broadcast of the factored panel per block step + one all-to-all for the column permutation"}
        unit, scaling = "GFLOP/s", "strong"
        backend.close()

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
