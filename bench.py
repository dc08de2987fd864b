#!/usr/bin/env python
"""Headline benchmark: decoded information-bit throughput (Gbps) of the batched decoders.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload scl8|bp504|bp504es|sc256|scl32|ms2016]
    python bench.py --impl reference ...      # CPU arm: the oracle port on all host cores

One "step" = one pass of the decode hot path over one batch of synthetic AWGN frames.
Default workload = BASELINE.json configs[1]: Polar SCL L=8 N=1024 K=512 (frozen set:
Bhattacharyya @ 2 dB), AWGN 2 dB; the second headline config (LDPC BP n=504, 20 iterations,
early_stop off) is measured in the same run and reported under "secondary".

  value    device-resident throughput: LLRs already in HBM, bits left in HBM, CUDA events on
           the launching stream, max over ranks; per-step input (>= 512 MiB) exceeds L2.
  e2e      same metric through the C-ABI host-buffer call (pcl_*_decode_host): pinned host
           LLRs -> H2D -> decode -> D2H bits, copies inside the timed region.
  roofline algorithmic on-chip bytes per frame x frames / kernel time, against
           the measured HBM copy bandwidth of MEASURED_PEAKS.json (the contract's denominator)
           and, under "onchip", against the SMEM crossbar peak the path is actually bound by.
  cpu_baseline  the fp64 oracle port (oracle/pcl_oracle.c, OpenMP over frames) on the host cores.
Multi-GPU: frames shard by rank (weak scaling, no data-path collective); the timed region ends
with the one allreduce of the error counters.
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

# BASELINE.json metric; `value` is the first config named (SCL-8 N=1024), BP n=504 it=20 is "secondary"
METRIC = "decoded info-bit Gbps (SCL-8 N=1024; BP n=504 it=20)"

WORKLOADS = {
    # name: kind, params, frames per GPU per step, info bits per frame, algorithmic bytes/frame
    "scl8": dict(kind="polar", N=1024, K=512, L=8, snr=2.0, frames=131072,
                 desc="Polar SCL L=8 N=1024 K=512, AWGN 2 dB (BASELINE configs[1])"),
    "scl32": dict(kind="polar", N=1024, K=512, L=32, snr=2.0, frames=32768,
                  desc="Polar SCL L=32 N=1024 K=512, AWGN 2 dB (BASELINE configs[3])"),
    "sc256": dict(kind="polar", N=256, K=128, L=1, snr=3.0, frames=524288,
                  desc="Polar SC N=256 K=128, AWGN 3 dB (BASELINE configs[0])"),
    "bp504": dict(kind="ldpc", n=504, k=252, mode="bp", iters=20, snr=1.0, frames=262144,
                  desc="LDPC BP n=504 (3,6) Gallager H seed 42, 20 iterations, early_stop off (BASELINE configs[2])"),
    "bp504es": dict(kind="ldpc", n=504, k=252, mode="bp", iters=20, snr=1.0, frames=262144, early_stop=True,
                    desc="LDPC BP n=504, max_iter=20 WITH syndrome early stop at 1 dB (mean ~3 iterations)"),
    "ms2016": dict(kind="ldpc", n=2016, k=1008, mode="ms", iters=20, snr=1.0, frames=65536,
                   desc="LDPC Min-Sum(0.75) n=2016 (3,6), 20 iterations, early_stop off (BASELINE configs[3])"),
}


def algorithmic_bytes(w):
    """fp32 on-chip bytes + HBM I/O per frame."""
    if w["kind"] == "polar":
        N, K, L = w["N"], w["K"], w["L"]
        return L * N * int(np.log2(N)) * 12 + 4 * N + K
    n, E = w["n"], 3 * w["n"]
    return w["iters"] * (16 * E + 4 * n) + 5 * n


def info_bits(w):
    return w["K"] if w["kind"] == "polar" else w["k"]


def measured_traffic(name, frames):
    """dram__bytes_read.sum + dram__bytes_write.sum of the workload's kernel from the committed
    `ncu --set full` capture (profiles/latest.json: bytes per frame), scaled to one launch."""
    p = os.path.join(ROOT, "profiles", "latest.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p)).get(name)
    if not d:
        return None, None
    return d["dram_bytes_per_frame"] * frames, d["source"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """SM clock / throttle reasons / power sampled DURING the timed region.  NVML through
    nvidia_ml_py in a thread (a query costs microseconds and does not disturb the kernels; a
    polling `nvidia-smi -lms` process was seen to slow individual steps by 10-30 %); falls back
    to nvidia-smi when the module is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.nvml, self.stop_flag, self.thread = None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = (pynvml, h)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nvml
        while not self.stop_flag:
            try:
                sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                try:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append((sm, mask, pw))
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            sm = [r[0] for r in self.rows]
            reasons = [n for n in self.NAMES if any(r[1] & self.BITS[n] for r in self.rows)]
            pw = [r[2] for r in self.rows]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                    "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons, "source": "nvidia-smi"}


# ------------------------------------------------------------------ inputs ------
_GEN_CACHE = {}


def make_inputs(w, torch, device, seed, frame0=0):
    """Synthetic frames of the workload's shape from the library's own on-device generator
    (csrc/framegen.cuh: random message -> encode -> BPSK + AWGN -> LLR = 2y/sigma^2, reference
    channel/awgn.py:47,75).  Returns (llr, reference bits the decoder output is compared with, code)."""
    import polarcode_and_ldpc_b200 as P
    F = w["frames"]
    if w["kind"] == "polar":
        key = ("polar", w["N"], w["K"])
        if key not in _GEN_CACHE:
            frozen = P.bhattacharyya_frozen_set(w["N"], w["K"], 2.0)
            _GEN_CACHE[key] = (P.FrameGenerator.polar(w["N"], w["K"], frozen), dict(frozen=frozen))
        gen, code = _GEN_CACHE[key]
        llr, msg, _ = gen.generate(F, w["snr"], seed=seed, frame0=frame0, device=device, want_codeword=False)
        return llr, msg, code
    key = ("ldpc", w["n"])
    if key not in _GEN_CACHE:
        H = P.gallager_parity_check(w["n"], 3, 6, 42)
        G, _ = P.generator_from_parity(H)
        _GEN_CACHE[key] = (P.FrameGenerator.ldpc(G), dict(H=H, k_true=G.shape[0]))
    gen, code = _GEN_CACHE[key]
    llr, _, cw = gen.generate(F, w["snr"], seed=seed, frame0=frame0, device=device, want_message=False)
    return llr, cw, code


def make_decoder(w, code, dtype="float32"):
    import polarcode_and_ldpc_b200 as P
    if w["kind"] == "polar":
        if w["L"] == 1:
            return P.SCDecoder(w["N"], w["K"], frozen_bits=code["frozen"], dtype=dtype)
        return P.SCLDecoder(w["N"], w["K"], list_size=w["L"], frozen_bits=code["frozen"], dtype=dtype)
    if w["mode"] == "bp":
        return P.BPDecoder(code["H"], max_iter=w["iters"], early_stop=bool(w.get("early_stop", False)), dtype=dtype)
    return P.MSDecoder(code["H"], max_iter=w["iters"], normalization=0.75, early_stop=False, dtype=dtype)


def run_gpu_workload(name, args, torch, dist, rank, world, device, with_e2e=True):
    import polarcode_and_ldpc_b200 as P
    w = dict(WORKLOADS[name])
    if args.frames:
        w["frames"] = args.frames
    llr, ref, code = make_inputs(w, torch, device, seed=1234, frame0=rank * w["frames"])
    dec = make_decoder(w, code)
    F = w["frames"]
    kbits = info_bits(w)
    counters = P.ErrorCounters(1, device=device)
    ncmp = None

    def step():
        bits = dec.decode_batch(llr)
        counters.add(0, bits, ref, ncmp)
        return bits

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    counters.t.zero_()
    barrier()
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else 0)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for s in range(args.steps):
        kev[s][0].record()
        bits = dec.decode_batch(llr)
        kev[s][1].record()
        counters.add(0, bits, ref, ncmp)
    counters.allreduce()                       # the path's one collective (no-op at N=1)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms_total, ms_kernel], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_kernel = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    gbps = world * F * kbits / (ms_step * 1e-3) / 1e9
    c = counters.t[0].cpu().tolist()

    # ---- end to end through the C-ABI host-buffer call --------------------------
    e2e = None
    if with_e2e:
        llr_host = torch.empty(llr.shape, dtype=llr.dtype, pin_memory=True)
        llr_host.copy_(llr)
        width = kbits if w["kind"] == "polar" else w["n"]
        bits_host = torch.empty((F, width), dtype=torch.uint8, pin_memory=True)
        dec.decode_batch_host(llr_host, bits_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dec.decode_batch_host(llr_host, bits_host)     # synchronous: returns with bits on host
        torch.cuda.synchronize()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e_gbps = world * F * kbits * args.steps / float(el[0]) / 1e9
        ok = bool((bits_host.to(device)[:, :ref.shape[1]] == ref).all(dim=1).float().mean() > 0.5)
        e2e = {"value": e2e_gbps, "unit": "Gbps", "h2d_bytes_per_step": int(llr_host.numel() * 4),
               "d2h_bytes_per_step": int(bits_host.numel()), "api": "pcl_*_decode_host (C ABI, pinned host buffers)",
               "sane": ok}
    info = dec.launch_info()
    alg = algorithmic_bytes(w)
    hbm_peak, peak_src, sm_max = peaks()
    achieved = alg * F / (ms_kernel * 1e-3) / 1e9
    smem_peak = 148 * 128 * sm_max * 1e6 / 1e9
    traffic, traffic_src = measured_traffic(name, F)
    res = {
        "workload": name, "desc": w["desc"], "frames_per_gpu_per_step": F, "gbps": gbps, "ms_per_step": ms_step,
        "ms_kernel": ms_kernel, "frames_per_s": world * F / (ms_step * 1e-3), "e2e": e2e,
        "fer": c[1] / max(c[2], 1), "ber": c[0] / max(c[3], 1), "frames_counted": c[2],
        "launch": info, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_bytes_per_frame": alg,
                     "kernel": info.get("kernel", "ldpc_decode_kernel"),
                     "onchip": {"bound": "smem", "peak": smem_peak, "unit": "GB/s", "frac": achieved / smem_peak,
                                "note": "148 SMs x 128 B/clk x max SM clock; the path is SMEM/issue bound, HBM sees only "
                                        f"{(4 * (w.get('N') or w.get('n')) + (w.get('K') or w.get('n')))} B/frame"}},
    }
    return res


# -------------------------------------------------------------- CPU baselines ---
def cpu_port_rate(name, seconds, threads):
    """Oracle port (fp64 C restatement) on `threads` host threads over a bounded sample."""
    import polarcode_and_ldpc_b200 as P
    from oracle import oracle
    w = WORKLOADS[name]
    rng = np.random.default_rng(5)
    np.random.seed(5)

    def frames(F):
        if w["kind"] == "polar":
            frozen = P.bhattacharyya_frozen_set(w["N"], w["K"], 2.0)
            cw = P.PolarEncoder(w["N"], w["K"], frozen).encode_batch(rng.integers(0, 2, size=(F, w["K"])))
            return frozen, P.AWGNChannel(w["snr"]).transmit_batch(cw)
        H = P.gallager_parity_check(w["n"], 3, 6, 42)
        return H, P.AWGNChannel(w["snr"]).transmit_batch(np.zeros((F, w["n"]), dtype=int))

    def run(code, llr):
        t0 = time.perf_counter()
        if w["kind"] == "polar":
            if w["L"] == 1:
                oracle.polar_sc(w["N"], code, llr, nthreads=threads)
            else:
                oracle.polar_scl(w["N"], w["L"], code, llr, nthreads=threads)
        else:
            oracle.ldpc(code, llr, w["mode"], max_iter=w["iters"], normalization=0.75 if w["mode"] == "ms" else 1.0,
                        early_stop=False, nthreads=threads)
        return time.perf_counter() - t0

    code, llr = frames(threads * 16)
    dt = run(code, llr)
    F = int(max(threads * 16, min(200000, seconds / (dt / llr.shape[0]))))
    F -= F % threads
    code, llr = frames(F)
    dt = run(code, llr)
    return F * info_bits(w) / dt / 1e9, F, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    threads = os.cpu_count() or oracle.max_threads()
    name = args.workload
    w = WORKLOADS[name]
    vals = []
    for s in range(args.warmup + args.steps):
        gbps, F, dt = cpu_port_rate(name, seconds=max(2.0, 60.0 / max(1, args.steps + args.warmup)), threads=threads)
        if s >= args.warmup:
            vals.append((gbps, F, dt))
    gb = float(np.mean([v[0] for v in vals]))
    Fs, dts = vals[-1][1], float(np.mean([v[2] for v in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": gb, "unit": "Gbps",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dts * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "frames_per_step": Fs},
        "cpu_baseline": {"value": gb, "unit": "Gbps", "cores": threads, "kind": "port",
                         "sample": f"{Fs} frames per step, oracle/pcl_oracle.c (fp64 restatement of the reference's "
                                   "numpy decoder; the Python reference itself cannot travel to this box), OpenMP over frames"},
        "e2e": {"value": gb, "unit": "Gbps", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="scl8", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: workload's)")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    main_res = run_gpu_workload(args.workload, args, torch, dist, rank, world, device)
    secondary = None
    if not args.no_secondary and args.workload == "scl8":
        secondary = run_gpu_workload("bp504", args, torch, dist, rank, world, device)

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        from oracle import oracle
        threads = os.cpu_count() or oracle.max_threads()
        g, Fs, dt = cpu_port_rate(args.workload, seconds=12.0, threads=threads)
        cpu = {"value": g, "unit": "Gbps", "cores": threads, "kind": "port",
               "sample": f"{Fs} frames of the same workload in {dt:.1f} s, oracle/pcl_oracle.c fp64, OpenMP over frames"}
        if secondary is not None:
            g2, F2, dt2 = cpu_port_rate("bp504", seconds=8.0, threads=threads)
            secondary["cpu_baseline"] = {"value": g2, "unit": "Gbps", "cores": threads, "kind": "port",
                                         "sample": f"{F2} frames in {dt2:.1f} s"}
    if rank == 0:
        w = WORKLOADS[args.workload]
        line = {
            "metric": METRIC, "value": main_res["gbps"], "unit": "Gbps",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "frames_per_gpu_per_step": main_res["frames_per_gpu_per_step"],
                       "l2": "per-step input (frames x N x 4 B >= 512 MiB) exceeds the 126 MB L2",
                       "sharding": f"frames sharded over {world} rank(s), one allreduce of error counters",
                       "launch": main_res["launch"], "fer": main_res["fer"], "ber": main_res["ber"]},
            "e2e": main_res["e2e"], "gpu_launches": 2 * args.steps,
            "clocks": main_res["clocks"], "roofline": main_res["roofline"], "cpu_baseline": cpu,
            "kernel_ms": main_res["ms_kernel"], "frames_per_s": main_res["frames_per_s"],
        }
        if secondary is not None:
            line["secondary"] = {k: secondary[k] for k in
                                 ("workload", "desc", "gbps", "ms_per_step", "ms_kernel", "frames_per_s", "e2e", "fer",
                                  "roofline", "launch", "frames_per_gpu_per_step") if k in secondary}
            if "cpu_baseline" in secondary:
                line["secondary"]["cpu_baseline"] = secondary["cpu_baseline"]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
