#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one process per GPU.

Headline workload (configs[1]): batched Chebyshev transform (plan_transform(Chebyshev(), 4096) on a
4096 x 65536 Float64 matrix = FFTW REDFT10 + scaling in the reference); metric = coefficients/s.
A "step" is one pass of the hot path over one such batch (one kernel launch, 2 GiB in, 2 GiB out).

  value      device-resident throughput (inputs already in HBM), CUDA events on the launching stream
  e2e        same metric through the C-ABI host call (afb_plan_execute_host) with pinned HOST buffers:
             H2D + kernels + D2H inside the timed region
  roofline   HBM roofline of the dominant kernel: 16 B per coefficient (8 B read + 8 B written)
  cpu_baseline  the oracle port (scipy pocketfft DCT-II + the reference's scaling) on the host cores
  extras     the other BASELINE.json workloads (inverse / Fourier transforms, Clenshaw, sampling)

`--impl reference` times the reference-side CPU implementation (the oracle port: Julia and FFTW are absent
from this image) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LEN = 4096
BATCH = 65536
ALG_BYTES_PER_COEF = 16.0
METRIC = "chebyshev_transform_coefficients_per_s"


# ------------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
def cpu_transform_sample(sample_batch: int, min_seconds: float, workers: int):
    """Reference-side CPU path: DCT-II (pocketfft) + /n, c0/2 -- the oracle port, all host threads."""
    import numpy as np
    from oracle import approxfun_oracle as orc

    rng = np.random.default_rng(0)
    v = rng.standard_normal((sample_batch, N_LEN))
    orc.cheb_transform(v[:64], axis=1, workers=workers)  # warm
    t0 = time.perf_counter()
    reps = 0
    while True:
        orc.cheb_transform(v, axis=1, workers=workers)
        reps += 1
        el = time.perf_counter() - t0
        if el >= min_seconds or reps >= 200:
            break
    return reps * sample_batch * N_LEN / el, el, reps


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores (oracle port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import approxfun_oracle as orc

    cores = os.cpu_count() or 1
    sample_batch = 2048  # 2048 x 4096 doubles = 64 MiB per step
    rng = np.random.default_rng(0)
    v = rng.standard_normal((sample_batch, N_LEN))
    for _ in range(max(args.warmup, 1)):
        orc.cheb_transform(v, axis=1, workers=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.cheb_transform(v, axis=1, workers=cores)
    el = time.perf_counter() - t0
    val = args.steps * sample_batch * N_LEN / el
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "coefficients/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"batched Chebyshev transform n={N_LEN}, Float64 (BASELINE.json configs[1]); "
                               f"each step a bounded sample of {sample_batch} of the {BATCH} functions"},
        "cpu_baseline": {"value": val, "unit": "coefficients/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_batch} x {N_LEN} f64 per step, scipy.fft.dct(type=2, workers={cores}) + /n, c0/2 "
                                   "(oracle port of FastTransforms.plan_chebyshevtransform; Julia/FFTW absent)"},
        "e2e": {"value": val, "unit": "coefficients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def protect_stdout():
    """Libraries (NCCL's version banner, ...) print to fd 1; keep fd 1 for the ONE JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="smaller Clenshaw/sampling point sets (development)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    numa_note = None
    if world > 1:
        # one process per GPU: run (and first-touch the pinned e2e buffers) on the CPUs NVML reports as local to this
        # GPU, so that eight ranks do not all stage their PCIe traffic through one socket's memory
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            ncpu = os.cpu_count() or 1
            masks = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(masks) for b in range(64) if (m >> b) & 1}
            allowed = os.sched_getaffinity(0)
            pick = sorted(cpus & allowed)
            if pick:
                os.sched_setaffinity(0, pick)
                numa_note = f"rank pinned to {len(pick)} GPU-local CPUs"
        except Exception as exc:  # affinity is an optimisation only
            numa_note = f"no CPU affinity ({type(exc).__name__})"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import approxfun_b200 as af
    from approxfun_b200 import _lib as L

    lib = L.load()
    L.check(lib, lib.afb_init(local))
    n, B = N_LEN, args.batch
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_loop(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps  # ms per step, max over ranks

    # ---- headline: device-resident batched Chebyshev transform (weak scaling: every rank owns B functions)
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    v = torch.randn(B, n, dtype=torch.float64, device="cuda", generator=g)
    c = torch.empty_like(v)
    fwd = af.api.DevicePlan(L.CHEB1_FWD, n, 0, B, n)
    launches_per_step = fwd.launches
    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        fwd.execute(v.data_ptr(), c.data_ptr(), stream)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fwd.execute(v.data_ptr(), c.data_ptr(), stream)
    e1.record()
    barrier()
    # the timed region lasts ~K x 0.75 ms, shorter than nvidia-smi's sampling period: keep the SAME kernel running
    # (untimed) for about a second so the clock / throttle record describes this load, then stop the sampler
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 1.0:
        for _ in range(50):
            fwd.execute(v.data_ptr(), c.data_ptr(), stream)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["note"] = "sampled over the timed region plus a 1 s untimed continuation of the same kernel"
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    coefs_per_step = float(B) * n
    value = world * coefs_per_step / (ms * 1e-3)

    peak, peak_src = measured_peaks()
    achieved = ALG_BYTES_PER_COEF * coefs_per_step / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "afb::line_kernel<2048, LK_CHEB_FWD>",
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_COEF * coefs_per_step, "peak_source": peak_src,
                "frac_of_8TBs_spec": achieved / 8000.0}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("cheb_fwd_4096_dram_bytes_per_launch")
        except Exception:
            pass

    # keep four input/output columns of the timed kernel: the cpu_baseline leg (the only place bench.py touches oracle/)
    # checks them against the reference-side CPU result
    idx = [0, 1, B // 2, B - 1]
    chk_in, chk_out = v[idx].cpu().numpy(), c[idx].cpu().numpy()
    chk_e2e = None

    # ---- e2e: host buffers through the C-ABI host call (pipelined H2D | kernel | D2H)
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(2, min(args.steps, 4))
        hv = torch.empty(B, n, dtype=torch.float64).pin_memory()
        hc = torch.empty(B, n, dtype=torch.float64).pin_memory()
        hv.copy_(v)
        hplan = af.api.DevicePlan(L.CHEB1_FWD, n, 0, B, n)
        hin, hout = hv.numpy(), hc.numpy()
        hplan.execute_host(hin, hout)  # warm (allocates the pipeline buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hplan.execute_host(hin, hout)
        torch.cuda.synchronize()
        el = (time.perf_counter() - t0) / e2e_steps
        el = max_over_ranks(el * 1e3) * 1e-3
        chk_e2e = hout[idx].copy()
        # the bound of this number: the same two pinned buffers copied H2D and D2H concurrently, no kernel
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        dv2 = torch.empty_like(v)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            with torch.cuda.stream(s_up):
                dv2.copy_(hv, non_blocking=True)
            with torch.cuda.stream(s_dn):
                hc.copy_(c, non_blocking=True)
        torch.cuda.synchronize()
        copy_ms = max_over_ranks((time.perf_counter() - t0) / 2 * 1e3)
        del dv2
        e2e = {"value": world * coefs_per_step / el, "unit": "coefficients/s", "h2d_bytes_per_step": int(8 * B * n),
               "d2h_bytes_per_step": int(8 * B * n), "ms_per_step": el * 1e3, "steps": e2e_steps,
               "api": "afb_plan_execute_host (plan*v through the C ABI, pinned host buffers)",
               "pcie_copy_only_ms": copy_ms, "frac_of_pcie_copy_only": copy_ms / (el * 1e3),
               "host_affinity": numa_note}
        hplan.close()
        del hv, hc

    # ---- extras: the other BASELINE.json workloads, short runs
    extras = {}
    if not args.no_extras:
        inv = af.api.DevicePlan(L.CHEB1_INV, n, 0, B, n)
        ff = af.api.DevicePlan(L.FOURIER_FWD, n, 0, B, n)
        fi = af.api.DevicePlan(L.FOURIER_INV, n, 0, B, n)
        for name, pl in (("chebyshev_itransform", inv), ("fourier_transform", ff), ("fourier_itransform", fi)):
            t = time_loop(lambda: pl.execute(v.data_ptr(), c.data_ptr(), stream), 5, 3)
            extras[name] = {"coefficients_per_s": world * coefs_per_step / (t * 1e-3), "ms_per_step": t,
                            "hbm_frac": ALG_BYTES_PER_COEF * coefs_per_step / (t * 1e-3) / 1e9 / peak}
        del c
        # Laurent (complex): 32 B per coefficient; same batch and length (4 GiB in, 4 GiB out)
        if B * n * 32 <= (24 << 30):
            vz = torch.randn(B, n, dtype=torch.complex128, device="cuda", generator=g)
            cz = torch.empty_like(vz)
            for name, kind in (("laurent_transform", L.LAURENT_FWD), ("laurent_itransform", L.LAURENT_INV)):
                pz = af.api.DevicePlan(kind, n, 0, B, n)
                t = time_loop(lambda: pz.execute(vz.data_ptr(), cz.data_ptr(), stream), 3, 3)
                extras[name] = {"coefficients_per_s": world * coefs_per_step / (t * 1e-3), "ms_per_step": t,
                                "hbm_frac": 32.0 * coefs_per_step / (t * 1e-3) / 1e9 / peak}
                pz.close()
            del vz, cz
        # config 3: Clenshaw, N = 10^4 coefficients at P = 10^9 points per rank (BASELINE.json configs[2], full size)
        Nc, P = 10000, (10 ** 9 if not args.quick else 1 << 26)
        xs = af.points(af.Chebyshev(), Nc)
        cc = torch.tensor(af.transform(af.Chebyshev(), np.sin(3000 * xs ** 2) + np.cos(977 * xs)), device="cuda")
        x = torch.rand(P, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
        y = torch.empty_like(x)
        t = time_loop(lambda: af.api.clenshaw_device(cc.data_ptr(), Nc, x.data_ptr(), y.data_ptr(), P, stream), 2, 1)
        evals = world * P / (t * 1e-3)
        sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
        fp64_peak = 148 * 64 * 2 * sm_mhz * 1e6
        extras["clenshaw_N1e4"] = {"evals_per_s": evals, "ms_per_step": t, "points_per_rank": P,
                                   "fp64_flops": 3.0 * Nc * P / (t * 1e-3),
                                   "fp64_frac_of_nominal_peak": 3.0 * Nc * P / (t * 1e-3) / fp64_peak,
                                   "nominal_peak_flops": fp64_peak}
        cl_chk_x, cl_chk_y = x[:4096].cpu().numpy(), y[:4096].cpu().numpy()
        # config 5: sampling exp(-x^2) on -5..5 (N ~ 66), on-device Philox uniforms
        f = af.Fun(lambda t_: np.exp(-t_ * t_), af.Chebyshev(-5, 5))
        cdf = af.normalizedcumsum(f).coefficients
        dc = torch.tensor(cdf, device="cuda")
        t = time_loop(lambda: af.api.sample_device(dc.data_ptr(), cdf.size, -5.0, 5.0, 42, 0, y.data_ptr(), P, stream), 3, 1)
        extras["sample_gaussian"] = {"samples_per_s": world * P / (t * 1e-3), "ms_per_step": t, "N": int(cdf.size),
                                     "fp64_frac_of_nominal_peak": 47 * 3.0 * cdf.size * P / (t * 1e-3) / fp64_peak}
        # measured FP64 issue rates (replaces the nominal peak as the denominator): 8 chains per thread
        import ctypes as _C
        blocks, iters = 148 * 8, 20000
        pbuf = torch.empty(blocks * 256, dtype=torch.float64, device="cuda")
        probe = {}
        for mode, name, fl in ((0, "dfma_3_register_operands", 2.0), (1, "dfma_repeated_operand", 2.0),
                               (2, "clenshaw_step_dadd_dfma", 3.0)):
            tp = time_loop(lambda: L.check(lib, lib.afb_fp64_probe_device(mode, iters, blocks, _C.c_void_p(pbuf.data_ptr()),
                                                                        _C.c_void_p(stream))), 3, 2)
            probe[name] = fl * blocks * 256 * 8 * iters / (tp * 1e-3)
        extras["fp64_probe_flops"] = probe
        extras["clenshaw_N1e4"]["frac_of_measured_clenshaw_step_rate"] = extras["clenshaw_N1e4"]["fp64_flops"] / probe["clenshaw_step_dadd_dfma"]
        extras["clenshaw_N1e4"]["frac_of_measured_dfma_rate"] = extras["clenshaw_N1e4"]["fp64_flops"] / probe["dfma_repeated_operand"]
        extras["sample_gaussian"]["frac_of_measured_clenshaw_step_rate"] = (
            47 * 3.0 * cdf.size * P / (extras["sample_gaussian"]["ms_per_step"] * 1e-3) / probe["clenshaw_step_dadd_dfma"])
        if rank == 0 and world == 1 and not args.no_cpu:
            # reference-side CPU baselines for configs 3 and 5 (oracle port, reference loop order: whole-vector passes
            # through the ClenshawPlan work vectors, all host threads), on bounded point sets
            from oracle import oracle_c as orcc
            cores = os.cpu_count() or 1
            xc = np.random.default_rng(1).uniform(-1, 1, 200000)
            ccpu = cc.cpu().numpy()
            extras["clenshaw_N1e4"]["bit_exact_vs_oracle"] = bool(np.array_equal(cl_chk_y, orcc.clenshaw(ccpu, cl_chk_x)))
            t0 = time.perf_counter()
            orcc.clenshaw_planned(ccpu, xc, threads=cores)
            tc = time.perf_counter() - t0
            extras["clenshaw_N1e4"]["cpu_baseline"] = {
                "value": xc.size / tc, "unit": "evaluations/s", "cores": cores, "kind": "port",
                "sample": f"{xc.size} points, N=10^4, k-outer/i-inner plan loop order (UPSTREAM AFOP), C + OpenMP"}
            uc = np.random.default_rng(2).random(2000000)
            t0 = time.perf_counter()
            orcc.cheb_bisect_planned(cdf, uc, threads=cores)
            tc = time.perf_counter() - t0
            extras["sample_gaussian"]["cpu_baseline"] = {
                "value": uc.size / tc, "unit": "samples/s", "cores": cores, "kind": "port",
                "sample": f"{uc.size} samples, 47 whole-vector Clenshaw passes (sample.jl:58-75 loop order), C + OpenMP"}
        del x, y
        # config 1: one long transform, n = 2^20 (transform + itransform of sin(x^2) on 0..10); 16 MiB per direction
        # lives in L2, so rotate over 16 distinct buffers (> 126 MB L2) and report it as a latency-bound number
        n1 = 1 << 20
        xs1 = torch.tensor(np.sin(af.points(af.Chebyshev(0, 10), n1) ** 2), device="cuda")
        bufs = [xs1.clone() for _ in range(16)]
        outs = [torch.empty_like(xs1) for _ in range(16)]
        p1f = af.api.DevicePlan(L.CHEB1_FWD, n1, 0, 1, n1)
        p1i = af.api.DevicePlan(L.CHEB1_INV, n1, 0, 1, n1)
        def cfg1():
            for b_, o_ in zip(bufs, outs):
                p1f.execute(b_.data_ptr(), o_.data_ptr(), stream)
                p1i.execute(o_.data_ptr(), o_.data_ptr(), stream)
        t = time_loop(cfg1, 3, 3) / 16.0
        rt = float((outs[0] - xs1).abs().max())
        extras["cheb_2pow20_fwd_plus_inv"] = {"ms_per_pair": t, "coefficients_per_s": 2.0 * n1 / (t * 1e-3),
                                              "launches_per_pair": p1f.launches + p1i.launches,
                                              "roundtrip_max_abs_err": rt,
                                              "note": "latency/L2 bound: 16 rotating buffers, fwd+inv per buffer"}
        del bufs, outs
        # config 4: 2-D Chebyshev x Chebyshev 16384 x 16384 on one GPU (contiguous columns, then the row-pair four-step)
        n2d = 16384
        V = torch.randn(n2d, n2d, dtype=torch.float64, device="cuda", generator=g)
        p2 = af.api.DevicePlan(L.CHEB1_2D_FWD, n2d, n2d, 1, n2d)
        t = time_loop(lambda: p2.execute(V.data_ptr(), V.data_ptr(), stream), 3, 2)
        extras["cheb2d_16384"] = {"ms_per_step": t, "elements_per_s": float(n2d) * n2d / (t * 1e-3),
                                  "hbm_frac_of_32B_per_element": 32.0 * n2d * n2d / (t * 1e-3) / 1e9 / peak,
                                  "launches": p2.launches}
        del V
        # "next" rows, one number each: arbitrary-length transforms (Bluestein in one kernel) and the Padua transform
        ng, bg = 1000, 65536
        vg = torch.randn(bg, ng, dtype=torch.float64, device="cuda", generator=g)
        pg = af.api.DevicePlan(L.CHEB1_FWD, ng, 0, bg, ng)
        t = time_loop(lambda: pg.execute(vg.data_ptr(), vg.data_ptr(), stream), 5, 3)
        extras["cheb_n1000_batched"] = {"ms_per_step": t, "coefficients_per_s": float(ng) * bg / (t * 1e-3),
                                        "hbm_frac": 16.0 * ng * bg / (t * 1e-3) / 1e9 / peak, "launches": pg.launches}
        del vg
        dpad = 2047
        npad = (dpad + 1) * (dpad + 2) // 2
        vp = torch.randn(npad, dtype=torch.float64, device="cuda", generator=g)
        pp = af.api.DevicePlan(L.PADUA_FWD, npad, 0, 1, npad)
        t = time_loop(lambda: pp.execute(vp.data_ptr(), vp.data_ptr(), stream), 5, 3)
        extras["padua_degree2047"] = {"ms_per_step": t, "coefficients_per_s": float(npad) / (t * 1e-3),
                                      "launches": pp.launches}
        del vp
        if world > 1:
            # config 4 across GPUs (strong scaling: ONE 16384 x 16384 matrix, column slabs -> all-to-all -> row slabs)
            from approxfun_b200 import parallel as par
            nc, nr = n2d // world, n2d // world
            blk = torch.randn(nc, n2d, dtype=torch.float64, device="cuda", generator=g)   # column-major n x nc
            rows = torch.empty(nr, n2d, dtype=torch.float64, device="cuda")
            nv_bytes = 8.0 * n2d * nc * (world - 1) / world       # bytes this rank sends over NVLink per transform
            for mode, p2p in (("nccl_sendrecv", False), ("peer_memory_fused", True)):
                sp = par.Sharded2DPlan(L.CHEB1_2D_FWD, n2d, n2d, p2p=p2p)
                t = time_loop(lambda: sp.execute(blk.data_ptr(), rows.data_ptr(), stream), 5, 3)
                extras[f"cheb2d_16384_sharded_{mode}"] = {
                    "ms_per_step": t, "elements_per_s": float(n2d) * n2d / (t * 1e-3), "scaling": "strong",
                    "nvlink_bytes_per_rank": nv_bytes,
                    "nvlink_frac_of_770GBs_if_transfer_alone": nv_bytes / (t * 1e-3) / 1e9 / 770.0}
                barrier()
                sp.close()
            del blk, rows

    # the CPU legs run at N = 1 only: at N > 1 the other ranks busy-wait in the barrier on the same host cores
    # (measured: the 32-thread port drops from 8.7e8 to 2.0e8 coefficients/s beside seven waiting ranks)
    cpu_baseline = None if world == 1 else {"value": None, "unit": "coefficients/s", "cores": 0, "kind": "port",
                                            "sample": "not timed at N > 1 (rank 0 at N = 1 only); see the N = 1 line"}
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        val, el, reps = cpu_transform_sample(2048, 10.0, cores)
        from oracle import approxfun_oracle as orc
        want = orc.cheb_transform(chk_in, axis=1)
        cpu_baseline = {"value": val, "unit": "coefficients/s", "cores": cores, "kind": "port",
                        "sample": f"{reps} x (2048 x {n}) f64 in {el:.1f} s, scipy.fft.dct(type=2, workers={cores}) + /n, c0/2 "
                                  "(oracle port; the Julia/FFTW reference cannot run in this image)",
                        "gpu_max_abs_err_vs_this_baseline": float(np.max(np.abs(chk_out - want))),
                        "e2e_max_abs_err_vs_this_baseline": None if chk_e2e is None else float(np.max(np.abs(chk_e2e - want)))}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "coefficients/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"batched Chebyshev transform: {B} functions x n={n} Float64 per GPU "
                                   "(BASELINE.json configs[1]), plan_transform(Chebyshev(), n) forward",
                       "batch_per_gpu": B, "n": n, "sharding": "by function batch, no collective",
                       "l2": "inputs (2 GiB per step) are larger than the 126 MB L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks, "extras": extras,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
