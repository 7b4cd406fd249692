#!/usr/bin/env python
"""bench.py -- headline benchmark: spectrogram Msamples/s (cf32, 1024-pt FFT) on N B200s.

Workload (BASELINE.json configs[0] parameters at the roofline-capable size SURVEY.md 8d names):
cf32_le synthetic recording, 2^28 complex samples PER GPU (2 GiB in, 2 GiB of float32 dB out),
1024-point Hann window, 50 % overlap (hop 512), 20*log10(|X|+1e-10), fft-shifted rows.
A "step" is one pass of the whole hot path (decode -> frame -> window -> FFT -> |X| -> dB) over
that block.  Multi-GPU: the recording is time-sharded, every rank owns a contiguous block plus an
(nfft - hop) halo it reads itself; no collective on the data path ("scaling": "weak").

  value      device-resident throughput (inputs already in HBM), CUDA-event timed, max over ranks
  e2e        same metric through the public host API (Engine.spectrogram on pinned HOST buffers):
             H2D of the step's samples and D2H of its dB image inside the timed region
  roofline   algorithmic bytes / kernel time against the measured HBM copy bandwidth
  cpu_baseline  the oracle (C FP64 port of the reference's Java path) on this box's host cores

--impl reference times that CPU port alone (the reference itself is Java; no JVM in the image).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NFFT, HOP, WINDOW, DATATYPE = 1024, 512, "hann", "cf32_le"
LOG2_SAMPLES_PER_GPU = 28
METRIC = "spectrogram Msamples/s (cf32, 1024-pt FFT)"
UNIT = "Msamples/s"
WORKLOAD = "cf32_le 2^%d samples/GPU, nfft 1024, Hann, hop 512 (50%% overlap), f32 dB out" % LOG2_SAMPLES_PER_GPU


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # warm the two queries of the sampling loop (the first call of each is tens of ms)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.001)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def bind_near_gpu(torch, index):
    """Run this process on the CPUs of the GPU's NUMA node (sysfs local_cpulist) when the cpuset allows it, so that
    the pinned host buffers of the end-to-end leg are allocated on the memory next to the GPU's PCIe root (on a
    two-socket host the far node costs a third of the H2D/D2H rate).  Returns (previous affinity, info)."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bdf
        with open(base + "/local_cpulist") as f:
            txt = f.read().strip()
        try:
            with open(base + "/numa_node") as f:
                node = int(f.read())
        except Exception:
            node = None
        local = set()
        for part in txt.split(","):
            if part:
                lo, _, hi = part.partition("-")
                local.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        both = allowed & local
        info = {"gpu_numa_node": node, "allowed_cpus": len(allowed), "gpu_local_allowed_cpus": len(both), "bound": False}
        if both and both != allowed:
            os.sched_setaffinity(0, both)
            info["bound"] = True
        return allowed, info
    except Exception as e:                                    # no sysfs entry in this container: leave the affinity alone
        return None, {"error": str(e)[:100]}


def make_device_recording(torch, n, seed, device):
    """Three tones + white noise (synth.TONES), generated on the device in chunks."""
    from spectral_analyzer_b200 import synth
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(2 * n, dtype=torch.float32, device=device)
    chunk = 1 << 24
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        t = torch.arange(s, s + m, dtype=torch.float64, device=device)
        re = torch.randn(m, generator=g, device=device, dtype=torch.float32) * synth.NOISE_SIGMA
        im = torch.randn(m, generator=g, device=device, dtype=torch.float32) * synth.NOISE_SIGMA
        for f, a in synth.TONES:
            ph = torch.remainder(f * t, 1.0) * (2 * 3.141592653589793)
            re += (a * torch.cos(ph)).float()
            im += (a * torch.sin(ph)).float()
        out[2 * s:2 * (s + m):2] = re
        out[2 * s + 1:2 * (s + m):2] = im
    return out


def cpu_port_throughput(n_samples, nthreads=0, min_seconds=0.0):
    """Times the oracle's spectrogram (C FP64 restatement of SpectralService.computeMagnitudes +
    the updateDisplay frame loop) on host cores; returns (Msamples/s, threads, seconds)."""
    import numpy as np
    from oracle import c_oracle as co
    from spectral_analyzer_b200 import synth
    block = synth.recording(1 << 20, DATATYPE, seed=1)
    raw = np.tile(block, max(1, n_samples >> 20))[: n_samples * 8]
    frames = (n_samples - NFFT) // HOP + 1
    threads = nthreads if nthreads > 0 else (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    co.spectrogram(raw[: 8 * (1 << 16)], DATATYPE, 0, NFFT, HOP, WINDOW, 64, nthreads=threads)     # warm
    reps, dt = 0, 0.0
    while reps < 1 or (dt < min_seconds and reps < 64):
        t0 = time.perf_counter()
        co.spectrogram(raw, DATATYPE, 0, NFFT, HOP, WINDOW, frames, nthreads=threads)
        dt += time.perf_counter() - t0
        reps += 1
    return reps * frames * HOP / dt / 1e6, threads, dt


def run_reference(args, rank):
    """Reference arm: the CPU implementation of the path (oracle port; the Java reference cannot
    run here) on the box's host cores, all threads."""
    if rank != 0:
        return
    n = 1 << 25          # bounded sample per step: 2^25 samples = 65535 frames (the size at which the port peaks)
    vals = []
    threads = 0
    for i in range(args.warmup + args.steps):
        v, threads, dt = cpu_port_throughput(n)
        if i >= args.warmup:
            vals.append((v, dt))
        if sum(d for _, d in vals) > 150:
            break
    ms = 1e3 * sum(d for _, d in vals) / len(vals)
    value = sum(v for v, _ in vals) / len(vals)
    sample = "2^25 samples per step (same cf32/1024/Hann/hop-512 parameters), %d steps" % len(vals)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-samples", type=int, default=LOG2_SAMPLES_PER_GPU)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg")
    ap.add_argument("--sustained-steps", type=int, default=400,
                    help="extra back-to-back steps timed as one region to show the power-capped rate (0: skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import spectral_analyzer_b200 as sa

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = 1 << args.log2_samples
    halo = NFFT - HOP
    # rank r owns samples [r*n, (r+1)*n) of the recording and reads the halo that follows it
    n_local = n + (halo if rank < world - 1 else 0)
    frames = (n_local - NFFT) // HOP + 1
    d_iq = make_device_recording(torch, n_local, seed=1 + rank, device=device)
    d_out = torch.empty((frames, NFFT), dtype=torch.float32, device=device)
    eng = sa.Engine(local_rank)
    params = eng.make_params(DATATYPE, NFFT, HOP, WINDOW, n_frames=frames)
    stream = torch.cuda.current_stream().cuda_stream
    iq_bytes, out_bytes = d_iq.numel() * 4, d_out.numel() * 4

    def step():
        eng.spectrogram_device(d_iq.data_ptr(), iq_bytes, params, d_out.data_ptr(), out_bytes, stream)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    launches0 = eng.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    sampler.start()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    barrier()
    sampler.stop_flag = True
    sampler.join()
    launches = eng.kernel_launches - launches0
    per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    ms_local = ev[0].elapsed_time(ev[args.steps]) / args.steps
    t = torch.tensor([ms_local], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    samples_per_step_local = frames * HOP
    tot = torch.tensor([float(samples_per_step_local)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    value = tot.item() / (ms * 1e-3) / 1e6

    # roofline of the (single) kernel of the step: algorithmic bytes = every input byte once +
    # every output byte once (8 + 4*nfft/hop = 16 B per input sample)
    alg_bytes = n_local * 8 + frames * NFFT * 4
    kern_ms = sum(per_step) / len(per_step)            # average launch duration (one launch per step)
    kern_best, kern_median = min(per_step), sorted(per_step)[len(per_step) // 2]
    peak, peak_kind = hbm_peak()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("spectrogram_f32_1024_cf32_hann_bytes_per_launch")
    except Exception:
        pass

    # ---- end-to-end leg: public host API, pinned host buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        prev_aff, numa = bind_near_gpu(torch, local_rank)
        h_iq = torch.empty(d_iq.numel(), dtype=torch.float32, pin_memory=True)
        h_iq.copy_(d_iq)
        h_out = torch.empty((frames, NFFT), dtype=torch.float32, pin_memory=True)
        h_iq_np, h_out_np = h_iq.numpy(), h_out.numpy()
        e2e_steps = max(3, min(args.steps, 8))
        for _ in range(2):
            eng.spectrogram(h_iq_np, DATATYPE, NFFT, frames, hop=HOP, window=WINDOW, out=h_out_np)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.spectrogram(h_iq_np, DATATYPE, NFFT, frames, hop=HOP, window=WINDOW, out=h_out_np)
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        td = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        e2e = {"value": round(tot.item() / td.item() / 1e6, 3), "unit": UNIT,
               "h2d_bytes_per_step": int(n_local * 8), "d2h_bytes_per_step": int(frames * NFFT * 4),
               "ms_per_step": round(td.item() * 1e3, 3), "steps": e2e_steps,
               "api": "Engine.spectrogram (sa_spectrogram C-ABI), pinned host in/out"}
        same = bool(torch.equal(h_out.to(device), d_out))
        e2e["matches_device_path"] = same
        e2e["numa"] = numa
        if prev_aff is not None and numa.get("bound"):
            os.sched_setaffinity(0, prev_aff)

    # ---- sustained leg: the same step back to back for ~0.4 s; on a 1 kW part the SM clock drops under
    # sw_power_cap, so this is the rate a long recording sees (reported beside, not instead of, `value`)
    sustained = None
    if args.sustained_steps > 0:
        s2 = ClockSampler(local_rank)
        s2.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.sustained_steps):
            step()
        e1.record()
        barrier()
        s2.stop_flag = True
        s2.join()
        ts = torch.tensor([e0.elapsed_time(e1) / args.sustained_steps], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sustained = {"steps": args.sustained_steps, "ms_per_step": round(ts.item(), 4),
                     "value": round(tot.item() / (ts.item() * 1e-3) / 1e6, 3), "unit": UNIT,
                     "roofline_frac": round(alg_bytes / (ts.item() * 1e-3) / 1e9 / peak, 4), "clocks": s2.summary()}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v1, _, dt1 = cpu_port_throughput(1 << 21, nthreads=1)
        # bounded sample: 2^25 samples repeated until about 4 s of wall time on all host threads
        log2n = 25
        vall, threads, dtall = cpu_port_throughput(1 << log2n, min_seconds=4.0)
        # second opinion: numpy / pocketfft FP64 on one thread, same framing, window and dB form
        import numpy as np
        from spectral_analyzer_b200 import synth
        nn = 1 << 21
        xs = np.frombuffer(synth.recording(nn, DATATYPE, seed=1).tobytes(), np.complex64).astype(np.complex128)
        wnd = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(NFFT) / NFFT)
        t0 = time.perf_counter()
        fr = np.lib.stride_tricks.sliding_window_view(xs, NFFT)[::HOP] * wnd
        _ = np.fft.fftshift(20 * np.log10(np.abs(np.fft.fft(fr, axis=1)) + 1e-10), axes=1)
        v_np = fr.shape[0] * HOP / (time.perf_counter() - t0) / 1e6
        cpu = {"value": round(vall, 3), "unit": UNIT, "cores": threads, "kind": "port",
               "numpy_pocketfft_1thread_value": round(v_np, 3),
               "sample": "2^%d samples (repeated), same parameters, %.1f s on %d threads; 1 thread (the reference's FX-thread "
                         "concurrency): %.3f Msamples/s on 2^21 samples" % (log2n, dtall, threads, v1),
               "single_thread_value": round(v1, 3)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "datatype": DATATYPE, "nfft": NFFT, "hop": HOP, "window": WINDOW,
                       "samples_per_gpu": n, "frames_per_gpu": frames, "sharding": "time blocks + %d-sample halo" % halo,
                       "l2": "inputs (2 GiB) and outputs (2 GiB) per step are far larger than the 126 MB L2"},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "peak_kind": peak_kind,
                         "alg_bytes_per_launch": int(alg_bytes), "kernel_ms": round(kern_ms, 4),
                         "kernel_ms_best": round(kern_best, 4), "kernel_ms_median": round(kern_median, 4),
                         "kernel": "spectrogram_tma_kernel<float,1024,cf32,window>"},
            "clocks": sampler.summary(),
            "gpu_launches": int(launches),
        }
        # the FP32 pipe bound beside the HBM one: 5 nfft log2(nfft) / hop + 6 flop per input sample against
        # 148 SMs x 128 FMA lanes x 2 flop at the SM clock sampled during the run
        import math
        flops_per_sample = 5.0 * NFFT * math.log2(NFFT) / HOP + 6.0
        sm_mhz = (line["clocks"] or {}).get("sm_mhz") or 1965
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        fp32_ach = samples_per_step_local * flops_per_sample / (kern_ms * 1e-3) / 1e12
        line["fp32_pipe"] = {"flops_per_sample": round(flops_per_sample, 1), "achieved": round(fp32_ach, 2),
                             "peak": round(fp32_peak, 1), "unit": "TFLOP/s", "frac": round(fp32_ach / fp32_peak, 4)}
        if e2e:
            line["e2e"] = e2e
        if sustained:
            line["sustained"] = sustained
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
