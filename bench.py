#!/usr/bin/env python
"""bench.py — headline benchmark of the gorp batch extraction path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A step = one pass of the hot path (newline index -> combined DFA -> capture -> span offsets -> histogram) over one
batch of synthetic access-log text: config #2 of BASELINE.json (README Put/Get/OtherRequest definition, 100 M lines
per GPU, a seeded 1 M-line block tiled in HBM). N > 1: one process per GPU (torchrun), lines sharded as independent
contiguous ranges, no data-path collective ("weak" scaling: 100 M lines per GPU).

  value        whole-job lines/s with the text already resident in HBM (device-resident C-ABI entry point)
  e2e          same metric through gorp_extract_text with HOST buffers: pinned host text -> H2D -> kernels ->
               D2H of every result array, all inside the timed region
  roofline     dominant kernel: algorithmic input bytes per launch / its CUDA-event time, vs measured HBM peak
  cpu_baseline the oracle's C restatement of the reference loop on the host cores (JVM unavailable here)

--impl reference times that CPU restatement alone (the reference is pure Java; no JVM exists in this image).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "log lines/sec and input GB/s per B200 (match+capture)"
# --workload: BASELINE.json configs; the headline (default) is config #2, the one the metric is quoted on for 1 GPU
WORKLOADS = {
    "readme": ("config#2 README Put/Get/OtherRequest 3-extraction definition, synthetic access-log lines", 1_000_000),
    "simple": ("config#1 samples/simple.grp, synthetic matching/non-matching lines", 1_000_000),
    "weblog": ("config#3 ~20-extraction nginx/Apache access+error definition with parametric templates, mixed line lengths", 200_000),
    "syslog200": ("config#4 200-extraction definition, combined DFA outgrows shared memory (L2-resident table)", 200_000),
    "utf16mix": ("config#5 nginx/Apache definition, non-ASCII UTF-16 + divergence characters + 10 KB outlier lines", 200_000),
}
WORKLOAD, BLOCK_LINES = WORKLOADS["readme"]
NCU_DRAM_BYTES_PER_TEXT_BYTE = {"k0_chunkwalk_extract": (1.370340e9 + 0.319971e9) / (8_000_000 * 62.612088 * 2)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0, self.t1 = index, None, [], None, None

    def launch(self):
        """Start nvidia-smi early (it needs ~0.1-0.3 s to come up); only the samples taken between start() and stop()
        are reported."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for ln in self.proc.stdout:
                    self.lines.append((time.perf_counter(), ln))
            self.t = threading.Thread(target=pump, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def start(self):
        if self.proc is None:
            self.launch()
        self.t0 = time.perf_counter()

    def stop(self):
        self.t1 = time.perf_counter()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)  # let the sample that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (t, ln) in self.lines if self.t0 <= t <= self.t1 + 0.06]
        if not inside and self.lines:  # region shorter than the sampling period: the sample closest to it
            inside = [min(self.lines, key=lambda x: abs(x[0] - self.t1))[1]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_block(rank, workload="readme"):
    from gorp_b200 import corpus
    return corpus.CONFIGS[workload][1](BLOCK_LINES, seed=0x5EED0000 + {"simple": 1, "readme": 2, "weblog": 3, "syslog200": 4, "utf16mix": 5}[workload] + 16 * rank)


def definition_of(workload):
    from gorp_b200 import corpus
    return corpus.CONFIGS[workload][0]


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU algorithm (C restatement, all host threads) on a bounded sample per step."""
    if rank != 0:
        return
    from gorp_b200 import corpus
    from oracle import gorp_oracle
    cores = os.cpu_count() or 1
    block = make_block(0, args.workload)
    reps = max(1, args.ref_lines // BLOCK_LINES)
    text = np.tile(block, reps)
    starts, ends = gorp_oracle.split_lines(text)
    o = gorp_oracle.Gorp(definition_of(args.workload))
    o.extract_batch(block, gorp_oracle.split_lines(block), threads=cores)  # build + page-in
    for _ in range(args.warmup):
        o.extract_batch(text, (starts, ends), threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.extract_batch(text, (starts, ends), threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    n = len(starts)
    val = n / dt
    sample = "%d lines (%.2f GB UTF-16) of the same synthetic workload per step" % (n, text.nbytes / 1e9)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "lines/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16", "data": "synthetic", "input_gb_per_s": text.nbytes / dt / 1e9,
        "config": {"workload": WORKLOAD, "lines_per_step": n, "note": "reference is pure Java and no JVM exists in this image: "
                   "C restatement of Gorp.extract (oracle/gorp_oracle.c), one extract per line, static partition over host threads"},
        "cpu_baseline": {"value": val, "unit": "lines/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lines-per-gpu", type=int, default=100_000_000)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-lines", type=int, default=16_000_000)
    ap.add_argument("--cpu-lines", type=int, default=16_000_000)
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostics: do not run nvidia-smi beside the timed region")
    ap.add_argument("--workload", default="readme", choices=sorted(WORKLOADS), help="BASELINE.json config (default: #2, the headline)")
    args = ap.parse_args()
    global WORKLOAD, BLOCK_LINES
    WORKLOAD, BLOCK_LINES = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from gorp_b200 import build as gbuild
    gbuild.build()
    from gorp_b200 import _ffi, corpus
    from gorp_b200.api import Blob, _check
    lib = _ffi.lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload: seeded block tiled to lines_per_gpu in HBM
    block = make_block(rank, args.workload)
    reps = max(1, args.lines_per_gpu // BLOCK_LINES)
    n_lines = reps * BLOCK_LINES
    d_block = torch.from_numpy(block.view(np.int16)).to(dev)
    d_text = d_block.repeat(reps)
    n_units = d_text.numel()
    in_bytes = n_units * 2
    blob = Blob.from_definition(definition_of(args.workload))
    eng = C.c_void_p()
    devs = (C.c_int * 1)(local_rank)
    _check(lib.gorp_engine_create(blob._ptr, blob.length, devs, 1, C.byref(eng)))
    stream = torch.cuda.current_stream().cuda_stream
    dres = _ffi.DeviceResult()

    # N > 1: the one (optional) collective of the path — all-reduce of the per-extraction histogram (E + 2 int64) over
    # NCCL, every step, inside the timed region
    from gorp_b200 import sharding
    n_bins = blob.info()[2] + 2
    d_hist = torch.zeros(n_bins, dtype=torch.int64, device=dev)
    try:
        from cuda.bindings import runtime as cudart
    except Exception:  # noqa: BLE001
        from cuda import cudart
    # the batch is the seeded block tiled `reps` times: its histogram must be reps x the block's (checked after warm-up,
    # an end-to-end sanity check of the timed path at full size; the parity tests proper are tests/ -m gpu)
    res0 = _ffi.Result()
    _check(lib.gorp_extract_text(eng, block.ctypes.data, block.size, C.byref(res0)))
    block_hist = np.ctypeslib.as_array(res0.histogram, (n_bins,)).copy()
    lib.gorp_result_release(eng, C.byref(res0))

    def step(flags=0):
        _check(lib.gorp_extract_text_device(eng, 0, d_text.data_ptr(), n_units, stream, flags, C.byref(dres)))
        if world > 1:
            (err,) = cudart.cudaMemcpyAsync(d_hist.data_ptr(), dres.d_histogram, n_bins * 8, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, stream)
            assert int(err) == 0, err
            sharding.allreduce_histogram(d_hist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if not args.no_clock_sampler:
        sampler.launch()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    assert dres.n_lines == n_lines, (dres.n_lines, n_lines)
    (err,) = cudart.cudaMemcpyAsync(d_hist.data_ptr(), dres.d_histogram, n_bins * 8, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, stream)
    assert int(err) == 0, err
    assert (d_hist.cpu().numpy() == reps * block_hist).all(), (d_hist.cpu().numpy().tolist(), (reps * block_hist).tolist())
    names = (C.c_char_p * 16)()
    tot = (C.c_double * 16)()
    cnt, calls, launches = C.c_int(), C.c_int64(), C.c_int64()
    _check(lib.gorp_kernel_times(eng, 0, names, tot, 16, C.byref(cnt), C.byref(calls), C.byref(launches), 1))

    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    per_step = []
    gap_ms = float(os.environ.get("GORP_BENCH_GAP_MS", "0"))  # diagnostics only: idle time between steps
    for _ in range(args.steps):
        if os.environ.get("GORP_BENCH_PER_STEP"):  # diagnostics only: a pair of events per step
            if gap_ms:
                torch.cuda.synchronize()
                time.sleep(gap_ms / 1e3)
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        step(_ffi.FLAG_TIME_KERNELS)
        if os.environ.get("GORP_BENCH_PER_STEP"):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            per_step.append((ev0, ev))
    e1.record()
    barrier()
    if per_step:
        print("[bench] ms per step:", [round(a.elapsed_time(b), 2) for a, b in per_step], file=sys.stderr)
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:  # every rank holds the job-wide histogram: its sum is the job's line count
        assert int(d_hist.sum().item()) == n_lines * world, (int(d_hist.sum().item()), n_lines * world)
    _check(lib.gorp_kernel_times(eng, 0, names, tot, 16, C.byref(cnt), C.byref(calls), C.byref(launches), 1))
    kern = {names[i].decode(): tot[i] / max(calls.value, 1) for i in range(cnt.value)}
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_lines = n_lines * world
    value = total_lines / (ms_max / 1e3)

    # ---- end to end through the host-buffer C ABI (H2D + kernels + D2H inside the timed region)
    e2e = None
    try:
        if args.skip_e2e:
            raise RuntimeError("skipped (--skip-e2e)")
        avail_gb = 0.0
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                avail_gb = int(ln.split()[1]) / 1e6
        per_rank_gb = in_bytes / 1e9 * 1.6
        e2e_reps = reps if avail_gb > per_rank_gb * world * 1.5 + 16 else max(1, int(reps * (avail_gb - 16) / (per_rank_gb * world * 1.5)))
        e2e_lines = e2e_reps * BLOCK_LINES
        h_text = torch.empty(e2e_reps * block.size, dtype=torch.int16).pin_memory()
        h_np = h_text.numpy().view(np.uint16)
        for r in range(e2e_reps):
            h_np[r * block.size:(r + 1) * block.size] = block
        res = _ffi.Result()

        def e2e_step():
            _check(lib.gorp_extract_text(eng, h_np.ctypes.data, h_np.size, C.byref(res)))
            nl, ns = res.n_lines, res.n_lines * res.span_stride
            lib.gorp_result_release(eng, C.byref(res))
            return nl, ns
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            nl, ns = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        d2h = nl * 4 + (nl + 1) * 8 + ns * 4 + 5 * 8
        e2e = {"value": e2e_lines * world / float(tt.item()), "unit": "lines/s", "h2d_bytes_per_step": int(h_np.size * 2),
               "d2h_bytes_per_step": int(d2h), "lines_per_step_per_gpu": int(e2e_lines), "ms_per_step": float(tt.item()) * 1e3,
               "timing": "host wall clock around gorp_extract_text (it returns after the last D2H), max over ranks"}
        # the same call fed with ISO-8859-1 bytes (what a JDK 9+ String with the LATIN1 coder holds; the synthetic corpus is
        # ASCII): gorp_extract_text_latin1 widens on the device, the host-to-device copy moves half the bytes
        if int(block.max()) < 256:
            h8_np = h_text.numpy().view(np.uint8)[:e2e_reps * block.size]  # reuses the pinned buffer of the UTF-16 run
            b8 = block.astype(np.uint8)
            for r in range(e2e_reps):
                h8_np[r * block.size:(r + 1) * block.size] = b8

            def e2e8_step():
                _check(lib.gorp_extract_text_latin1(eng, h8_np.ctypes.data, h8_np.size, C.byref(res)))
                nl8 = res.n_lines
                lib.gorp_result_release(eng, C.byref(res))
                return nl8
            assert e2e8_step() == nl
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e8_step()
            torch.cuda.synchronize()
            dt8 = (time.perf_counter() - t0) / args.e2e_steps
            tt8 = torch.tensor([dt8], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt8, op=dist.ReduceOp.MAX)
            e2e["latin1_input"] = {"value": e2e_lines * world / float(tt8.item()), "unit": "lines/s", "h2d_bytes_per_step": int(h8_np.size),
                                   "d2h_bytes_per_step": int(d2h), "ms_per_step": float(tt8.item()) * 1e3,
                                   "call": "gorp_extract_text_latin1 (ISO-8859-1 bytes in, widened to UTF-16 on the device)"}
        del h_text
    except Exception as ex:  # noqa: BLE001
        e2e = {"value": None, "unit": "lines/s", "error": str(ex)[:200]}

    if rank == 0:
        peak, peak_src = hbm_peak()
        dom = max(kern, key=kern.get) if kern else None
        roof = None
        if dom:
            ach = in_bytes / (kern[dom] / 1e3) / 1e9
            # DRAM bytes per launch from the committed ncu --set full capture of the same kernel (profiles/
            # r1h_ncu_summary_readme.txt: 1.370 GB read + 0.320 GB written per 1.002 GB of text), scaled to this launch
            traffic = NCU_DRAM_BYTES_PER_TEXT_BYTE.get(dom)
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic * in_bytes if traffic else None,
                    "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per text byte (profiles/r1h_ncu_summary_readme.txt) x bytes of this launch" if traffic else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": in_bytes,
                    "kernel_ms": kern[dom], "all_kernels_ms": kern,
                    "whole_step_frac": (in_bytes / (ms_max / 1e3) / 1e9) / peak,
                    "whole_step_frac_of_8TBps": (in_bytes / (ms_max / 1e3) / 1e9) / 8000.0}
        # CPU baseline on rank 0: bounded sample of the same workload
        from oracle import gorp_oracle
        cores = os.cpu_count() or 1
        if args.skip_cpu:
            args.cpu_lines = BLOCK_LINES
        creps = max(1, args.cpu_lines // BLOCK_LINES)
        ctext = np.tile(block, creps)
        cst = gorp_oracle.split_lines(ctext)
        o = gorp_oracle.Gorp(definition_of(args.workload))
        o.extract_batch(block, gorp_oracle.split_lines(block), threads=cores)
        t0 = time.perf_counter()
        o.extract_batch(ctext, cst, threads=cores)
        cdt = time.perf_counter() - t0
        t0 = time.perf_counter()
        o.extract_batch(block, gorp_oracle.split_lines(block), threads=1)
        cdt1 = time.perf_counter() - t0
        cpu = {"value": len(cst[0]) / cdt, "unit": "lines/s", "cores": cores, "kind": "port",
               "sample": "%d lines (%.2f GB) of the same workload, all %d host threads; 1 thread: %.0f lines/s"
                         % (len(cst[0]), ctext.nbytes / 1e9, cores, BLOCK_LINES / cdt1),
               "note": "C restatement of the reference's Gorp.extract loop (JVM unavailable in this image)"}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "lines/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16",
            "data": "synthetic", "input_gb_per_s": in_bytes * world / (ms_max / 1e3) / 1e9,
            "config": {"workload": WORKLOAD, "lines_per_gpu": n_lines, "units_per_gpu": n_units,
                       "bytes_per_gpu": in_bytes, "block": "%d-line seeded block tiled %dx in HBM" % (BLOCK_LINES, reps),
                       "l2": "input (%.1f GB) is far larger than L2, no flush needed" % (in_bytes / 1e9),
                       "parallelism": "lines sharded per GPU as contiguous ranges, tables replicated; no data-path collective"
                                      + (", one NCCL all-reduce of the %d-bin histogram per step" % n_bins if world > 1 else "")},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(launches.value),
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
