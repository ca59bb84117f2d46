#!/usr/bin/env python
"""bench.py -- PairHMM forward throughput (GCUPS) on B200, BASELINE.json's metric on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2] [--impl ours|reference]

One "step" = one pass of the hot path over one batch: row parameters, float forward kernels (which also build the
fallback list), double re-run of the fallback list
(pmm_launch of include/pairhmm_cuda.h).  At N = 1 the workload is configs[1] of BASELINE.json (config 2, the
GATK-HaplotypeCaller-like active region: 1000 reads x 151 bp against 64 haplotypes of 300-600 bp, ~4.4e9 cells).
With N > 1 (launched by torch.distributed.run, one rank per GPU) every rank runs a batch of the same shape with its
own seed: pairs are independent, nothing is reduced, so there is no data-path collective ("weak" scaling); NCCL is
used only for the barrier and for the max over ranks of the device time.

Numbers in the JSON line
  value        whole-job GCUPS (cells = sum(read_len) * sum(hap_len), /root/reference/pairhmm/host/main.cpp:305-313)
               with inputs resident in HBM, device time from CUDA events, L2 flushed between steps
  e2e          the same metric through the public C ABI with HOST buffers, wall clock: every step packs its inputs into
               pinned memory, copies them to the GPU, builds the haplotype stream, runs the kernels, copies the results
               back and takes log10 on the host.  `value` runs the steps through the work queue (pmm_pool_*, three
               contexts on the GPU, so neighbouring steps overlap); `serial_value` runs them one at a time on one context
  roofline     the float forward kernel against the measured FP32 instruction-issue rate of this GPU
               (12 FP32 instructions per cell: 8 FMUL + 4 FADD)
  cpu_baseline the reference's own AVX implementation (oracle/_ref, built from /root/reference with pinned flags) on
               the host cores of this box -- a reported baseline, not the target
--impl reference times that CPU implementation as the step itself.
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

import numpy as np  # noqa: E402

METRIC = "pairhmm_forward_gcups"
UNIT = "GCUPS"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json configs index (1-based), default 2")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; reported in config)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the extra fast-mode measurement")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        import datetime
        rows = []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        use = inside or rows                      # the timed region can be shorter than nvidia-smi's first sample
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = set()
        for r in use:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": float(max(r[2] for r in use)),
                "power_w": float(np.median([r[3] for r in use])), "reasons": sorted(reasons),
                "samples": len(use), "samples_inside_timed_region": len(inside)}


def workload(cfg: int, seed: int, scale: float):
    from acc_genomics_b200 import synth
    return synth.config(cfg, seed=seed, scale=scale)


def hbm_view(traffic_bytes, kernel_ms):
    """The same kernel against the HBM roofline of MEASURED_PEAKS.json -- to show that this path is not memory-bound."""
    peak, src = 6547.5, "fallback: this pool's measured copy bandwidth quoted in the task (6547.5 GB/s)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"]); src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    if not traffic_bytes:
        return {"achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "peak_source": src}
    ach = traffic_bytes / (kernel_ms * 1e-3) * 1e-9
    return {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src}


def cells_of(batches) -> int:
    return int(sum(b.num_cells for b in batches))


def cpu_reference_pass(lib, batches, threads: int):
    """One pass of the reference's batch loop (FalconPairHMM::computePairhmmAVX) over the workload; seconds."""
    t0 = time.perf_counter()
    nfb = 0
    for b in batches:
        _, _, fb = lib.batch(b, threads=threads)
        nfb += int(fb.sum())
    return time.perf_counter() - t0, nfb


def run_reference(args, rank: int, world: int, out=sys.stdout):
    """--impl reference: the reference's own CPU implementation on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import oracle
    lib = oracle.reference()
    kind = "reference"
    if lib is None:
        lib, kind = oracle.port(), "port"
    threads = os.cpu_count() or 1
    batches = workload(args.config, args.config, args.scale)
    cells = cells_of(batches)
    for _ in range(args.warmup):
        cpu_reference_pass(lib, batches, threads)
    t = [cpu_reference_pass(lib, batches, threads)[0] for _ in range(args.steps)]
    total = sum(t)
    val = cells * args.steps / total * 1e-9
    from acc_genomics_b200 import synth
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": synth.CONFIG_NAMES[args.config], "scale": args.scale, "cells_per_step": cells,
                   "note": "reference AVX PairHMM (float pass, double re-run, log10) on host cores; no GPU involved"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "the full workload per step, all host threads, static split over reads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=out, flush=True)


_T0 = time.perf_counter()


def log(msg: str):
    print(f"[bench +{time.perf_counter() - _T0:7.2f}s] {msg}", file=sys.stderr, flush=True)


def main():
    args = parse()
    # Libraries (NCCL's version banner, for one) write to fd 1; the contract is ONE JSON line on stdout.  Keep the real
    # stdout aside and point fd 1 at stderr until the line is printed.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    try:
        _main(args, real_stdout)
    finally:
        real_stdout.flush()


def _main(args, real_stdout):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, real_stdout)
        return

    import torch
    import torch.distributed as dist
    from acc_genomics_b200 import synth
    from acc_genomics_b200.engine import PairHMMEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the PairHMM engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- workload: same shape on every rank, rank-specific seed (weak scaling, no exchange) ---------------------
    batches = workload(args.config, args.config + 1000 * rank, args.scale)
    cells = cells_of(batches)
    eng = PairHMMEngine(local)
    side = torch.cuda.Stream()                 # a real (non-default) stream: torch.cuda.Event records on the current stream
    torch.cuda.set_stream(side)
    eng.set_option("stream", side.cuda_stream)
    log("workload generated")
    eng.stage(batches)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2

    # ---- device-resident timing -------------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        eng.launch()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.t0 = time.time()
    for a, b in ev:
        flush.fill_(1)                      # evict L2 between steps (outside the timed bracket)
        a.record()
        eng.launch()
        b.record()
        b.synchronize()
    barrier()
    sampler.t1 = time.time()
    log("timed region done")
    clocks = sampler.stop()
    log("clock sampler stopped")
    step_ms = [a.elapsed_time(b) for a, b in ev]
    eng.sync()
    st = eng.stats()
    dev_s = sum(step_ms) * 1e-3

    # per-kernel time of the float forward kernel(s), from the engine's own events on the same stream
    f32_ms, fb_ms = [], []
    for _ in range(min(args.steps, 10)):
        flush.fill_(1)
        eng.launch(); eng.sync()
        f32_ms.append(eng.stats()["ms_f32"]); fb_ms.append(eng.stats()["ms_fallback"])
    f32_ms_avg = float(np.mean(f32_ms))
    raw = eng.fetch_raw()
    out, nfb = eng.fetch_log10()
    assert np.isfinite(raw).all() and not np.isnan(out).any()

    # ---- end to end through the C ABI with host buffers -----------------------------------------------------------
    # (a) one job at a time on one context: stage (pack, H2D) -> launch -> fetch (D2H, host log10), nothing overlapped
    res = np.empty(st["pairs"], dtype=np.float64)
    for _ in range(3):
        eng.restage(); eng.launch(); eng.fetch_log10(res)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.restage(); eng.launch(); eng.fetch_log10(res)
    torch.cuda.synchronize()
    e2e_serial_s = time.perf_counter() - t0
    st2 = eng.stats()
    # (b) the same steps through the work queue a host application uses (pmm_pool_*): every step still packs its
    # inputs from host memory, copies them to the GPU, runs, copies the results back and takes log10 on the host, but
    # three contexts keep the copies and the host work of neighbouring steps under the kernels
    from acc_genomics_b200.engine import PairHMMPool
    from collections import deque
    depth = 3
    pool = PairHMMPool(devices=[local], contexts_per_device=depth)
    outs = [np.empty(st["pairs"], dtype=np.float64) for _ in range(depth + 1)]

    def run_pool(n):
        live = deque()
        for k in range(n):
            if len(live) > depth:
                pool.wait(live.popleft())
            live.append(pool.submit(None, out=outs[k % (depth + 1)], job=eng._job))
        while live:
            pool.wait(live.popleft())
    run_pool(2 * depth)
    barrier()
    t0 = time.perf_counter()
    run_pool(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(outs[0].view(np.uint64), out.view(np.uint64)), "pool result differs from the single-context result"
    pool.close()
    log("e2e done")

    # ---- fast mode (opt-in: contracted float kernel + exact re-check of the guard band), same workload, reported aside
    fast = None
    if not args.no_fast_mode:
        eng.set_option("mode", "fast")
        eng.restage()
        for _ in range(3):
            eng.launch()
        barrier()
        fev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 50))]
        for a, b in fev:
            flush.fill_(1)
            a.record(); eng.launch(); b.record(); b.synchronize()
        fast_s = sum(a.elapsed_time(b) for a, b in fev) * 1e-3
        fms = []
        for _ in range(5):
            flush.fill_(1); eng.launch(); eng.sync(); fms.append(eng.stats()["ms_f32"])
        out_f, nfb_f = eng.fetch_log10()
        fin = np.isfinite(out)
        fast = {"steps": len(fev), "dev_s": fast_s, "f32_ms": float(np.mean(fms)), "recheck_pairs": int(eng.stats()["recheck_pairs"]),
                "fallback_pairs": int(nfb_f), "same_decision": bool(nfb_f == nfb and np.array_equal(np.isfinite(out_f), fin)),
                "max_rel_log10_vs_exact": float(np.max(np.abs(out_f[fin] - out[fin]) / np.abs(out[fin])))}
        eng.set_option("mode", "exact")
        log("fast mode done")

    # ---- FP32 issue peak, measured on this GPU ----------------------------------------------------------------------
    peak_lane_instr, _ = eng.measure_fp32_peak()

    # ---- max over ranks -----------------------------------------------------------------------------------------------
    times = torch.tensor([dev_s, e2e_s, e2e_serial_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        total_cells = torch.tensor([float(cells)], dtype=torch.float64, device="cuda")
        dist.all_reduce(total_cells, op=dist.ReduceOp.SUM)
        job_cells = float(total_cells.item())
    else:
        job_cells = float(cells)
    dev_s, e2e_s, e2e_serial_s = float(times[0].item()), float(times[1].item()), float(times[2].item())

    if rank == 0:
        value = job_cells * args.steps / dev_s * 1e-9
        e2e = job_cells * args.steps / e2e_s * 1e-9
        # roofline of the dominant kernel (float forward): 12 FP32 instructions per cell
        f32_cells_per_s = cells / (f32_ms_avg * 1e-3)
        achieved = f32_cells_per_s * 12 * 1e-12
        peak = peak_lane_instr * 1e-12
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        nominal_peak = sms * 128 * clocks["sm_mhz"] * 1e6 * 1e-12 if clocks.get("sm_mhz") else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (exact op order, no FMA contraction) + f64 re-run of underflowed pairs", "data": "synthetic",
            "config": {"workload": synth.CONFIG_NAMES[args.config], "scale": args.scale, "cells_per_step_per_gpu": cells,
                       "pairs_per_step_per_gpu": int(st["pairs"]), "fallback_pairs": int(nfb), "flush_pairs": int(st["flush_pairs"]),
                       "mode": "exact", "l2": "flushed between steps (256 MiB fill outside the timed bracket)",
                       "parallelism": f"{world} x independent region batches, no collective"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(st2["h2d_bytes"]), "d2h_bytes_per_step": int(st2["d2h_bytes"]),
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "path": "pmm_pool_submit_flat / pmm_pool_wait, 3 contexts on the GPU: per step pack + H2D + kernels + D2H + "
                            "host log10, host numpy buffers in, float64 log10 out, steps overlapped by the queue",
                    "serial_value": job_cells * args.steps / e2e_serial_s * 1e-9,
                    "serial_ms_per_step": e2e_serial_s / args.steps * 1e3,
                    "serial_path": "pmm_stage_flat + pmm_launch + pmm_fetch_log10 on one context, nothing overlapped"},
            "gpu_launches": int(st["kernel_launches"]) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "fp32_issue", "kernel": "pmm_forward_kernel<float,K,W> (float pass)", "achieved": achieved, "peak": peak,
                         "unit": "T FP32 lane-instr/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one float-pass launch on this workload, from the
                         # ncu --set full capture under profiles/ (r01h): the kernel is nowhere near a memory bound
                         "traffic": 5702144 if (args.config == 2 and args.scale == 1.0) else None,
                         "traffic_source": "profiles/r01h_f32_K19W8_ncu_metrics.txt (5.70 MB read, 0 written per launch; algorithmic input 1.2 MB + 4.9 MB row parameters)",
                         "kernel_ms": f32_ms_avg, "kernel_gcups": f32_cells_per_s * 1e-9,
                         "kernel_ms_note": "CUDA events around read_params_kernel + the float forward launch(es) on the engine's stream",
                         "fallback_pass_ms": float(np.mean(fb_ms)),
                         "algorithmic": "12 FP32 instr per cell (8 FMUL + 4 FADD) x cells per launch",
                         "peak_source": "measured live: independent FMUL/FADD streams (pmm_measure_fp32_peak); "
                                        "MEASURED_PEAKS.json has no FP32 figure",
                         "frac_flop_convention": achieved / (2 * peak),
                         "nominal_peak": nominal_peak, "frac_vs_nominal": (achieved / nominal_peak) if nominal_peak else None,
                         "nominal_peak_note": "SMs x 128 FP32 lanes x the SM clock sampled during the timed region",
                         "hbm_view": hbm_view(5702144 if (args.config == 2 and args.scale == 1.0) else None, f32_ms_avg)},
        }
        if fast:
            fcps = cells / (fast["f32_ms"] * 1e-3)
            line["fast_mode"] = {
                "note": "opt-in pmm_set_option(mode=fast): float cell update contracted to 4 FMUL + 4 FFMA (8 instr/cell), exact "
                        "re-check of results within 2^-7 of the 1e-28f threshold; decision identical, log10 within 1e-5 relative. "
                        "This rank only; not the headline value",
                "value": cells * fast["steps"] / fast["dev_s"] * 1e-9, "unit": UNIT, "ms_per_step": fast["dev_s"] / fast["steps"] * 1e3,
                "kernel_gcups": fcps * 1e-9, "roofline_frac_8_instr_per_cell": fcps * 8 / peak_lane_instr,
                "recheck_pairs": fast["recheck_pairs"], "same_decision_as_exact": fast["same_decision"],
                "max_rel_log10_vs_exact": fast["max_rel_log10_vs_exact"]}
        log("gpu side done")
        if not args.no_cpu_baseline and world >= 1:
            try:
                import oracle
                lib = oracle.reference()
                kind = "reference"
                if lib is None:
                    lib, kind = oracle.port(), "port"
                threads = os.cpu_count() or 1
                cpu_reference_pass(lib, batches, threads)
                best = min(cpu_reference_pass(lib, batches, threads)[0] for _ in range(3))
                one = cpu_reference_pass(lib, [batches[0].slice_reads(0, max(1, batches[0].num_read // 16))], 1)[0]
                line["cpu_baseline"] = {"value": cells / best * 1e-9, "unit": UNIT, "cores": threads, "kind": kind,
                                        "sample": "the full workload (all pairs incl. double re-run and log10), best of 3 passes",
                                        "one_thread_gcups": cells_of([batches[0].slice_reads(0, max(1, batches[0].num_read // 16))]) / one * 1e-9,
                                        "build": "g++ -O3 -mavx -ffp-contract=off (the arithmetic of Intel GKL's AVX build: no FMA contraction)"}
                # the build the reference's own Makefile makes (-march=native => FMA contraction, different low bits)
                try:
                    with open("/proc/cpuinfo") as f:
                        flags = f.read()
                except OSError:
                    flags = ""
                fma = oracle.reference(fma=True) if kind == "reference" and " fma" in flags and " avx2" in flags else None
                if fma is not None:
                    cpu_reference_pass(fma, batches, threads)
                    tf = min(cpu_reference_pass(fma, batches, threads)[0] for _ in range(3))
                    line["cpu_baseline"]["fma_contracted_build_gcups"] = cells / tf * 1e-9   # -mavx2 -mfma: what -march=native yields here
                try:
                    with open("/proc/cpuinfo") as f:
                        model = [ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")]
                    line["cpu_baseline"]["cpu_model"] = model[0] if model else None
                except OSError:
                    pass
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}
        log("cpu baseline done")
        print(json.dumps(line), file=real_stdout, flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
