#!/usr/bin/env python
"""Headline benchmark: decoded information-bit throughput (Gbps) of the batched decoders.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--quick]
    python bench.py --impl reference ...      # CPU arm: the reference's numpy decoders on all host cores

One "step" = one pass of the decode hot path over one batch of synthetic AWGN frames.
Headline workload = BASELINE.json configs[1]: Polar SCL L=8 N=1024 K=512 (frozen set:
Bhattacharyya @ 2 dB) over the config's SNR sweep; it is timed at -2, 0, 2 and 5 dB and `value`
is the LOWEST of the four (the prune shortcuts make throughput data dependent).  The second
headline config (LDPC BP n=504, 20 iterations, early_stop off) and the other BASELINE configs
(SC N=256, SCL-32 at four rates, Min-Sum n=2016, BP with early stop, the fp64 validation build)
are measured in the same run and listed under "secondary".

  value    device-resident throughput: LLRs already in HBM, bits left in HBM; every step is timed
           with CUDA events on the launching stream and `value` uses the MEDIAN step (min / mean /
           max beside it) plus the amortised tail (the one allreduce of the error counters);
           max over ranks; per-step input (>= 512 MiB) exceeds L2.
  e2e      same metric through the C-ABI host-buffer call (pcl_*_decode_host): pinned host
           LLRs -> H2D -> decode -> D2H bits, copies inside the timed region.  `e2e_dropin` is the
           reference call shape itself: decode_batch(np.float64[F, N]) from pageable memory.
  roofline algorithmic on-chip bytes per frame x frames / median kernel time against
           the MEASURED shared-memory bandwidth (profiles/onchip_peaks.json, scripts/peaks_microbench.cu);
           the HBM view (I/O bytes, ncu DRAM traffic) sits under "hbm_io", the MUFU view (BP) under "mufu".
  cpu_baseline  the reference's own numpy decoders (baseline/_ref/refsrc, multiprocessing.Pool over
           the host cores) on frames of the same workload, their bits compared with the GPU's; the
           fp64 C port of the oracle is the second entry.
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

# BASELINE.json metric; `value` is the first config named (SCL-8 N=1024), BP n=504 it=20 leads "secondary"
METRIC = "decoded info-bit Gbps (SCL-8 N=1024; BP n=504 it=20)"
SNR_SWEEP = (-2.0, 0.0, 2.0, 5.0)

WORKLOADS = {
    # name: kind, params, frames per GPU per step
    "scl8": dict(kind="polar", N=1024, K=512, L=8, snr=2.0, frames=131072,
                 desc="Polar SCL L=8 N=1024 K=512, AWGN SNR sweep -2/0/2/5 dB (BASELINE configs[1])"),
    "scl8_f64": dict(kind="polar", N=1024, K=512, L=8, snr=2.0, frames=32768, dtype="float64",
                     desc="Polar SCL L=8 N=1024 K=512, fp64 validation build (bit-exact with the reference)"),
    "scl32": dict(kind="polar", N=1024, K=512, L=32, snr=2.0, frames=32768,
                  desc="Polar SCL L=32 N=1024 K=512, AWGN 2 dB (BASELINE configs[3], rate 0.50)"),
    "scl32_k686": dict(kind="polar", N=1024, K=686, L=32, snr=2.0, frames=32768,
                       desc="Polar SCL L=32 N=1024 K=686 (rate 0.67)"),
    "scl32_k768": dict(kind="polar", N=1024, K=768, L=32, snr=2.0, frames=32768,
                       desc="Polar SCL L=32 N=1024 K=768 (rate 0.75)"),
    "scl32_k849": dict(kind="polar", N=1024, K=849, L=32, snr=2.0, frames=32768,
                       desc="Polar SCL L=32 N=1024 K=849 (rate 0.83)"),
    "sc256": dict(kind="polar", N=256, K=128, L=1, snr=3.0, frames=524288,
                  desc="Polar SC N=256 K=128, AWGN 3 dB (BASELINE configs[0])"),
    "sc1024": dict(kind="polar", N=1024, K=512, L=1, snr=2.0, frames=262144,
                   desc="Polar SC N=1024 K=512, AWGN 2 dB"),
    "bp504": dict(kind="ldpc", n=504, k=252, mode="bp", iters=20, snr=1.0, frames=262144,
                  desc="LDPC BP n=504 (3,6) Gallager H seed 42, 20 iterations, early_stop off (BASELINE configs[2])"),
    "bp504es": dict(kind="ldpc", n=504, k=252, mode="bp", iters=20, snr=1.0, frames=262144, early_stop=True,
                    desc="LDPC BP n=504, max_iter=20 WITH syndrome early stop at 1 dB (mean ~3 iterations)"),
    "ms2016": dict(kind="ldpc", n=2016, k=1008, mode="ms", iters=20, snr=1.0, frames=65536,
                   desc="LDPC Min-Sum(0.75) n=2016 (3,6), 20 iterations, early_stop off (BASELINE configs[3])"),
}
SECONDARY = ("bp504", "sc256", "sc1024", "scl32", "scl32_k686", "scl32_k768", "scl32_k849", "ms2016", "bp504es",
             "scl8_f64")


def algorithmic_bytes(w):
    """fp32 on-chip bytes + HBM I/O per frame."""
    if w["kind"] == "polar":
        N, K, L = w["N"], w["K"], w["L"]
        return L * N * int(np.log2(N)) * 12 + 4 * N + K
    n, E = w["n"], 3 * w["n"]
    # with the syndrome early stop the figure scales with the MEAN iteration count of the batch
    return w.get("mean_iters", w["iters"]) * (16 * E + 4 * n) + 5 * n


def io_bytes(w):
    """HBM I/O per frame: fp32 LLRs in, one byte per decoded bit out."""
    return 4 * w["N"] + w["K"] if w["kind"] == "polar" else 5 * w["n"]


def info_bits(w):
    return w["K"] if w["kind"] == "polar" else w["k"]


def measured_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per frame of the workload's kernel from the
    committed `ncu --set full` capture (profiles/latest.json)."""
    p = os.path.join(ROOT, "profiles", "latest.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p)).get(name)
    if not d:
        return None, None
    return d["dram_bytes_per_frame"], d["source"]


def peaks():
    """Roofline denominators: driver-measured HBM copy bandwidth + our measured on-chip peaks."""
    out = {"hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)", "sm_max_mhz": 1965.0}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        out.update(hbm_gbs=float(d["hbm_gbs"]), hbm_source="measured (MEASURED_PEAKS.json)",
                   sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)))
    out["smem_gbs"] = 148 * 128 * out["sm_max_mhz"] * 1e6 / 1e9
    out["smem_source"] = "computed: 148 SMs x 128 B/clk x max SM clock (no measurement committed)"
    out["mufu_gops"] = 148 * 16 * out["sm_max_mhz"] * 1e6 / 1e9
    out["mufu_source"] = "computed: 148 SMs x 16 MUFU/clk x max SM clock"
    p = os.path.join(ROOT, "profiles", "onchip_peaks.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if d.get("smem_lds128_gbs"):
            out["smem_gbs"] = float(d["smem_lds128_gbs"])
            out["smem_source"] = "measured (profiles/onchip_peaks.json: conflict-free LDS.128 stream on all SMs)"
        if d.get("mufu_ex2_gops"):
            out["mufu_gops"] = float(d["mufu_ex2_gops"])
            out["mufu_source"] = "measured (profiles/onchip_peaks.json: independent ex2.approx streams)"
    return out


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


def merge_clocks(a, b):
    """Clock records of two timed regions -> one (median of medians is not needed: keep the lower clock)."""
    if a is None:
        return b
    if b is None:
        return a
    out = dict(a)
    if b.get("sm_mhz") is not None and (a.get("sm_mhz") is None or b["sm_mhz"] < a["sm_mhz"]):
        out["sm_mhz"] = b["sm_mhz"]
    out["reasons"] = sorted(set(a.get("reasons", [])) | set(b.get("reasons", [])))
    out["samples"] = (a.get("samples") or 0) + (b.get("samples") or 0)
    if a.get("power_w_max") is not None and b.get("power_w_max") is not None:
        out["power_w_max"] = max(a["power_w_max"], b["power_w_max"])
    return out


# ------------------------------------------------------------------ inputs ------
_GEN_CACHE = {}
_CODE_CACHE = {}


def polar_code(N, K):
    import polarcode_and_ldpc_b200 as P
    key = ("polar", N, K)
    if key not in _CODE_CACHE:
        _CODE_CACHE[key] = dict(frozen=P.bhattacharyya_frozen_set(N, K, 2.0))
    return _CODE_CACHE[key]


def ldpc_code(n):
    import polarcode_and_ldpc_b200 as P
    key = ("ldpc", n)
    if key not in _CODE_CACHE:
        H = P.gallager_parity_check(n, 3, 6, 42)
        G, _ = P.generator_from_parity(H)
        _CODE_CACHE[key] = dict(H=H, G=G, k_true=G.shape[0])
    return _CODE_CACHE[key]


def make_inputs(w, torch, device, seed, frame0=0):
    """Synthetic frames of the workload's shape from the library's own on-device generator
    (csrc/framegen.cuh: random message -> encode -> BPSK + AWGN -> LLR = 2y/sigma^2, reference
    channel/awgn.py:47,75).  Returns (llr, reference bits the decoder output is compared with, code)."""
    import polarcode_and_ldpc_b200 as P
    F = w["frames"]
    dt = w.get("dtype", "float32")
    if w["kind"] == "polar":
        code = polar_code(w["N"], w["K"])
        key = ("polar", w["N"], w["K"], str(device))
        if key not in _GEN_CACHE:
            _GEN_CACHE[key] = P.FrameGenerator.polar(w["N"], w["K"], code["frozen"])
        llr, msg, _ = _GEN_CACHE[key].generate(F, w["snr"], seed=seed, frame0=frame0, device=device, want_codeword=False,
                                               dtype=dt)
        return llr, msg, code
    code = ldpc_code(w["n"])
    key = ("ldpc", w["n"], str(device))
    if key not in _GEN_CACHE:
        _GEN_CACHE[key] = P.FrameGenerator.ldpc(code["G"])
    llr, _, cw = _GEN_CACHE[key].generate(F, w["snr"], seed=seed, frame0=frame0, device=device, want_message=False,
                                          dtype=dt)
    return llr, cw, code


def make_decoder(w, code, dtype=None):
    import polarcode_and_ldpc_b200 as P
    dtype = dtype or w.get("dtype", "float32")
    if w["kind"] == "polar":
        if w["L"] == 1:
            return P.SCDecoder(w["N"], w["K"], frozen_bits=code["frozen"], dtype=dtype)
        return P.SCLDecoder(w["N"], w["K"], list_size=w["L"], frozen_bits=code["frozen"], dtype=dtype)
    if w["mode"] == "bp":
        return P.BPDecoder(code["H"], max_iter=w["iters"], early_stop=bool(w.get("early_stop", False)), dtype=dtype)
    return P.MSDecoder(code["H"], max_iter=w["iters"], normalization=0.75, early_stop=False, dtype=dtype)


_DEC_CACHE = {}


def cached_decoder(name, w, code):
    key = (name, w.get("dtype", "float32"))
    if key not in _DEC_CACHE:
        _DEC_CACHE[key] = make_decoder(w, code)
    return _DEC_CACHE[key]


def stats(xs):
    xs = [float(x) for x in xs]
    return {"median": float(np.median(xs)), "min": min(xs), "mean": float(np.mean(xs)), "max": max(xs), "n": len(xs)}


def run_gpu_workload(name, steps, warmup, env, snr=None, frames=None, with_e2e=False, with_dropin=False):
    """Device-resident timing of one workload: every step bracketed by CUDA events; the single
    allreduce of the error counters closes the timed region (warmed before it, like the decoder)."""
    import polarcode_and_ldpc_b200 as P
    torch, dist, rank, world, device = env
    w = dict(WORKLOADS[name])
    if frames:
        w["frames"] = frames
    if snr is not None:
        w["snr"] = snr
    llr, ref, code = make_inputs(w, torch, device, seed=1234, frame0=rank * w["frames"])
    dec = cached_decoder(name, w, code)
    F = w["frames"]
    kbits = info_bits(w)
    counters = P.ErrorCounters(1, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        bits = dec.decode_batch(llr)
        counters.add(0, bits, ref, None)
    counters.allreduce()                       # warms the exact collective of the timed region
    counters.t.zero_()
    start_line = torch.zeros(1, device=device)
    if world > 1:
        dist.all_reduce(start_line)
    barrier()
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else 0)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 2)]
    if world > 1:
        # GPU-side start line: the host threads leave barrier() milliseconds apart; this collective is
        # queued ahead of the first event, so every rank's timed region opens when the LAST rank's
        # stream arrives and the closing allreduce no longer measures the hosts' skew
        dist.all_reduce(start_line)
    ev[0].record()
    for s in range(steps):
        bits = dec.decode_batch(llr)
        ev[2 * s + 1].record()                 # end of the decode kernel of step s
        counters.add(0, bits, ref, None)
        ev[2 * s + 2].record()                 # end of step s
    counters.allreduce()                       # the path's one collective (no-op at N=1)
    ev[2 * steps + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [ev[2 * s].elapsed_time(ev[2 * s + 2]) for s in range(steps)]
    kern_ms = [ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(steps)]
    total_ms = ev[0].elapsed_time(ev[2 * steps + 1])
    tail_ms = ev[2 * steps].elapsed_time(ev[2 * steps + 1])
    mine = torch.tensor([float(np.median(step_ms)), float(np.median(kern_ms)), total_ms, tail_ms, min(step_ms),
                         max(step_ms), float(np.mean(step_ms))], dtype=torch.float64, device=device)
    if world > 1:
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = torch.stack(allr).cpu().numpy()
    else:
        per_rank = mine.cpu().numpy()[None, :]
    med_step, med_kern, total_ms, tail_ms = (float(per_rank[:, i].max()) for i in range(4))
    ms_eff = med_step + tail_ms / steps                      # median step + amortised collective
    gbps = world * F * kbits / (ms_eff * 1e-3) / 1e9
    c = counters.t[0].cpu().tolist()

    res = {
        "workload": name, "desc": w["desc"], "snr_db": w["snr"], "frames_per_gpu_per_step": F, "gbps": gbps,
        "ms_per_step": total_ms / steps, "ms_step": {"median": med_step, "min": float(per_rank[:, 4].min()),
                                                     "mean": float(per_rank[:, 6].max()), "max": float(per_rank[:, 5].max())},
        "ms_kernel": med_kern, "ms_tail_allreduce": tail_ms, "gbps_mean_of_steps": world * F * kbits / (total_ms / steps * 1e-3) / 1e9,
        "frames_per_s": world * F / (ms_eff * 1e-3),
        "fer": c[1] / max(c[2], 1), "ber": c[0] / max(c[3], 1), "frames_counted": c[2],
        "launch": dec.launch_info(), "clocks": clocks, "dtype": w.get("dtype", "float32"),
    }
    if world > 1:
        res["per_rank_ms_kernel"] = [float(x) for x in per_rank[:, 1]]
    if w["kind"] == "ldpc" and w.get("early_stop"):
        _, its = dec.decode_batch(llr[: min(F, 65536)], return_iterations=True)       # outside the timed region
        w["mean_iters"] = float(its.float().mean().item())
        res["mean_iterations"] = w["mean_iters"]
    res["roofline"] = roofline(name, w, med_kern, F, res["launch"])

    # ---- end to end through the C-ABI host-buffer call --------------------------
    if with_e2e:
        width = kbits if w["kind"] == "polar" else w["n"]
        words = (width + 31) // 32
        shifts = torch.arange(32, device=device, dtype=torch.int32)

        def e2e_leg(llr_host, label):
            bits_host = torch.empty((F, words), dtype=torch.int32, pin_memory=True)
            dec.decode_batch_host(llr_host, bits_host, packed=True)
            dec.decode_batch_host(llr_host, bits_host, packed=True)
            times = []
            barrier()
            t_all = time.perf_counter()
            for _ in range(steps):
                t0 = time.perf_counter()
                dec.decode_batch_host(llr_host, bits_host, packed=True)   # synchronous: returns with the bits on the host
                times.append(time.perf_counter() - t0)
            torch.cuda.synchronize()
            el = torch.tensor([float(np.median(times)), time.perf_counter() - t_all], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(el, op=dist.ReduceOp.MAX)
            # unpack a slice on the device and compare with the transmitted bits
            sl = bits_host[:4096].to(device)
            got = ((sl.unsqueeze(-1) >> shifts) & 1).reshape(sl.shape[0], -1)[:, :ref.shape[1]].to(torch.uint8)
            ok = bool((got == ref[:4096]).all(dim=1).float().mean() > 0.5)
            return {"value": world * F * kbits / float(el[0]) / 1e9, "unit": "Gbps",
                    "h2d_bytes_per_step": int(llr_host.numel() * llr_host.element_size()),
                    "d2h_bytes_per_step": int(bits_host.numel() * 4),
                    "value_mean_of_steps": world * F * kbits * steps / float(el[1]) / 1e9,
                    "api": f"pcl_*_decode_host_ex (C ABI): pinned {label} LLRs in, bit-packed rows out (packed on the device)",
                    "sane": ok}

        llr_host = torch.empty(llr.shape, dtype=llr.dtype, pin_memory=True)
        llr_host.copy_(llr)
        res["e2e"] = e2e_leg(llr_host, "float32" if llr.dtype == torch.float32 else "float64")
        if llr.dtype == torch.float32:
            # opt-in transport format: half the PCIe bytes, the decoder sees fp16-rounded LLRs (reported
            # separately; parity of this mode is judged against the oracle fed the same rounded values)
            llr16 = torch.empty(llr.shape, dtype=torch.float16, pin_memory=True)
            llr16.copy_(llr)
            res["e2e_fp16_transport"] = e2e_leg(llr16, "float16")
            del llr16
        del llr_host
    if with_dropin:
        # the reference call shape: decode_batch(np.float64[F, N]) from pageable memory, int64 bits back
        Fd = min(F, 32768)
        llr_np = llr[:Fd].double().cpu().numpy()
        out = dec.decode_batch(llr_np)
        times = []
        barrier()
        for _ in range(max(3, steps // 2)):
            t0 = time.perf_counter()
            out = dec.decode_batch(llr_np)
            times.append(time.perf_counter() - t0)
        el = torch.tensor([float(np.median(times))], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        ok = bool((out[:, :ref.shape[1]] == ref[:Fd].cpu().numpy()).all(axis=1).mean() > 0.5)
        res["e2e_dropin"] = {"value": world * Fd * kbits / float(el[0]) / 1e9, "unit": "Gbps", "frames": Fd,
                             "h2d_bytes_per_step": int(llr_np.size * 4), "host_bytes_read_per_step": int(llr_np.nbytes),
                             "d2h_bytes_per_step": int(out.shape[0] * out.shape[1]),
                             "api": "decode_batch(np.float64[F, N]) -> np.int64[F, K] (reference call shape, pageable memory)",
                             "sane": ok}
    return res


def roofline(name, w, ms_kernel, F, launch):
    pk = peaks()
    alg = algorithmic_bytes(w)
    achieved = alg * F / (ms_kernel * 1e-3) / 1e9
    traffic_pf, traffic_src = measured_traffic(name)
    io = io_bytes(w)
    r = {"bound": "smem", "achieved": achieved, "peak": pk["smem_gbs"], "unit": "GB/s", "frac": achieved / pk["smem_gbs"],
         "traffic": traffic_pf * F if traffic_pf is not None else None,
         "peak_source": pk["smem_source"], "algorithmic_bytes_per_frame": alg, "frames_per_launch": F,
         "kernel": launch.get("kernel", "ldpc_decode_kernel"), "kernel_ms_median": ms_kernel,
         "hbm_io": {"io_bytes_per_frame": io, "achieved_gbs": io * F / (ms_kernel * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                    "peak_source": pk["hbm_source"], "frac": io * F / (ms_kernel * 1e-3) / 1e9 / pk["hbm_gbs"],
                    "traffic_bytes_per_frame": traffic_pf, "traffic_source": traffic_src,
                    "traffic_over_io": (traffic_pf / io) if traffic_pf is not None else None}}
    if w["kind"] == "ldpc" and w["mode"] == "bp":
        # MUFU per edge per iteration: ex2 in, lg2 out, and a reciprocal per edge (3) or per PAIR of edges (2.5:
        # the banked kernel decodes two checks per lane and shares the reciprocal, ldpc_bp.cuh cn_bp_core2)
        per_edge = 2.5 if launch.get("kernel") == "ldpc_banked_kernel" else 3.0
        ops = per_edge * 3 * w["n"] * w.get("mean_iters", w["iters"])
        g = ops * F / (ms_kernel * 1e-3) / 1e9
        r["mufu"] = {"ops_per_frame": ops, "mufu_per_edge": per_edge, "achieved_gops": g, "peak": pk["mufu_gops"], "peak_source": pk["mufu_source"],
                     "frac": g / pk["mufu_gops"]}
    return r


# -------------------------------------------------------------- CPU baselines ---
def host_frames(name, F, seed=5, snr=None):
    """Host-generated frames of the workload (numpy), the way the reference's callers make them."""
    import polarcode_and_ldpc_b200 as P
    w = WORKLOADS[name]
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    snr = w["snr"] if snr is None else snr
    if w["kind"] == "polar":
        frozen = polar_code(w["N"], w["K"])["frozen"]
        cw = P.PolarEncoder(w["N"], w["K"], frozen).encode_batch(rng.integers(0, 2, size=(F, w["K"])))
        return frozen, P.AWGNChannel(snr).transmit_batch(cw)
    H = ldpc_code(w["n"])["H"]
    return H, P.AWGNChannel(snr).transmit_batch(np.zeros((F, w["n"]), dtype=int))


def cpu_port_rate(name, seconds, threads):
    """Oracle port (fp64 C restatement) on `threads` host threads over a bounded sample."""
    from oracle import oracle
    w = WORKLOADS[name]

    def run(code, llr):
        t0 = time.perf_counter()
        if w["kind"] == "polar":
            if w["L"] == 1:
                oracle.polar_sc(w["N"], code, llr, nthreads=threads)
            else:
                oracle.polar_scl(w["N"], w["L"], code, llr, nthreads=threads)
        else:
            oracle.ldpc(code, llr, w["mode"], max_iter=w["iters"], normalization=0.75 if w["mode"] == "ms" else 1.0,
                        early_stop=bool(w.get("early_stop", False)), nthreads=threads)
        return time.perf_counter() - t0

    code, llr = host_frames(name, threads * 16)
    dt = run(code, llr)
    F = int(max(threads * 16, min(200000, seconds / (dt / llr.shape[0]))))
    F -= F % threads
    code, llr = host_frames(name, F)
    dt = run(code, llr)
    return F * info_bits(w) / dt / 1e9, F, dt


def ref_spec(name):
    w = WORKLOADS[name]
    if w["kind"] == "polar":
        return dict(kind="sc" if w["L"] == 1 else "scl", N=w["N"], K=w["K"], L=w["L"],
                    frozen=np.asarray(polar_code(w["N"], w["K"])["frozen"]))
    return dict(kind=w["mode"], H=np.asarray(ldpc_code(w["n"])["H"]), iters=w["iters"],
                early_stop=bool(w.get("early_stop", False)), normalization=0.75)


def numpy_reference_rate(pool, name, frames_per_core, gpu_check=None, snr=None):
    """The reference's own decode(llr) over a Pool of all host cores on `frames_per_core` x cores
    frames of the workload.  gpu_check(llr) -> bits of the CUDA path on the same LLRs (compared)."""
    from oracle import ref_worker
    w = WORKLOADS[name]
    F = pool.procs * frames_per_core
    _, llr = host_frames(name, F, seed=17, snr=snr)
    bits, wall, cpu_s = pool.decode(ref_spec(name), llr)
    gb = F * info_bits(w) / wall / 1e9
    out = {"value": gb, "unit": "Gbps", "cores": pool.procs, "kind": "reference", "cpu_model": ref_worker.cpu_model(),
           "per_core_kbps": F * info_bits(w) / cpu_s / 1e3, "frames": F, "wall_s": wall,
           "sample": f"{F} frames ({frames_per_core} per core) of the same workload in {wall:.1f} s through the reference's own "
                     f"numpy decoder (baseline/_ref/refsrc, multiprocessing.Pool({pool.procs}))"}
    if gpu_check is not None:
        got = gpu_check(llr)
        width = bits.shape[1]
        out["gpu_bits_equal_reference_frames"] = int((got[:, :width] == bits).all(axis=1).sum())
    return out


def cpu_baselines(name, threads, pool, gpu_check, port_seconds, frames_per_core):
    g, Fs, dt = cpu_port_rate(name, seconds=port_seconds, threads=threads)
    port = {"value": g, "unit": "Gbps", "cores": threads, "kind": "port",
            "sample": f"{Fs} frames of the same workload in {dt:.1f} s, oracle/pcl_oracle.c fp64, OpenMP over frames"}
    if pool is None:
        return port
    ref = numpy_reference_rate(pool, name, frames_per_core, gpu_check)
    ref["port"] = port
    return ref


def run_reference_arm(args):
    """CPU arm: the reference's numpy decoders (kind "reference") on all host cores, a bounded
    sample per step; the C port only when the staged reference is missing."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_worker
    name = args.workload
    w = WORKLOADS[name]
    threads = ref_worker.host_cores()
    vals = []
    if ref_worker.available():
        pool = ref_worker.ReferencePool(threads)
        per_core = 2 if w["kind"] == "polar" and w["L"] > 1 else 4
        for s in range(args.warmup + args.steps):
            snr = SNR_SWEEP[s % len(SNR_SWEEP)] if name == "scl8" else None
            r = numpy_reference_rate(pool, name, per_core if s >= args.warmup else 1, snr=snr)
            if s >= args.warmup:
                vals.append((r["value"], r["frames"], r["wall_s"]))
        pool.close()
        kind = "reference"
        how = (f"{vals[-1][1]} frames per step through the reference's own numpy decoders "
               f"(baseline/_ref/refsrc, /root/reference/src/polar/decoder.py:225 / src/ldpc/decoder.py:124), "
               f"multiprocessing.Pool({threads}), {ref_worker.cpu_model()}")
    else:
        for s in range(args.warmup + args.steps):
            g, F, dt = cpu_port_rate(name, seconds=max(2.0, 60.0 / max(1, args.steps + args.warmup)), threads=threads)
            if s >= args.warmup:
                vals.append((g, F, dt))
        kind = "port"
        how = (f"{vals[-1][1]} frames per step, oracle/pcl_oracle.c (fp64 restatement; baseline/_ref/refsrc was not "
               "staged), OpenMP over frames")
    gb = float(np.median([v[0] for v in vals]))
    Fs, dts = vals[-1][1], float(np.median([v[2] for v in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": gb, "unit": "Gbps",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dts * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "snr_db": list(SNR_SWEEP) if name == "scl8" else [w["snr"]], "frames_per_step": Fs},
        "cpu_baseline": {"value": gb, "unit": "Gbps", "cores": threads, "kind": kind, "sample": how},
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
    ap.add_argument("--quick", action="store_true", help="headline workload only: no secondary list, no CPU arms")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.quick:
        args.no_secondary = args.no_cpu = True

    if args.impl == "reference":
        run_reference_arm(args)
        return

    # the CPU pool is spawned before CUDA is touched in this process
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pool = None
    threads = 1
    if rank == 0 and not args.no_cpu and world == 1:
        from oracle import ref_worker
        threads = ref_worker.host_cores()
        if ref_worker.available():
            pool = ref_worker.ReferencePool(threads)

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    env = (torch, dist, rank, world, device)
    F = args.frames or None

    launches = 0
    headline = args.workload == "scl8"
    sweep = []
    if headline:
        for snr in SNR_SWEEP:
            sweep.append(run_gpu_workload("scl8", args.steps, args.warmup, env, snr=snr, frames=F,
                                          with_e2e=(snr in (SNR_SWEEP[0], 2.0)), with_dropin=(snr == 2.0)))
            launches += 2 * args.steps
        main_res = min(sweep, key=lambda r: r["gbps"])
    else:
        main_res = run_gpu_workload(args.workload, args.steps, args.warmup, env, frames=F, with_e2e=True, with_dropin=True)
        launches += 2 * args.steps
    secondary = []
    if not args.no_secondary and headline:
        for nm in SECONDARY:
            first = nm == "bp504"
            secondary.append(run_gpu_workload(nm, args.steps if first else 3, 3, env, with_e2e=first, with_dropin=first))
            launches += 2 * (args.steps if first else 3)

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        def gpu_bits(nm):
            def fn(llr):
                w = WORKLOADS[nm]
                return cached_decoder(nm, w, polar_code(w["N"], w["K"]) if w["kind"] == "polar" else ldpc_code(w["n"])
                                      ).decode_batch(llr)
            return fn
        cpu = cpu_baselines(args.workload, threads, pool, gpu_bits(args.workload), 10.0, 8)
        for s in secondary:
            if s["workload"] == "bp504":
                s["cpu_baseline"] = cpu_baselines("bp504", threads, pool, gpu_bits("bp504"), 6.0, 8)
        if pool is not None:
            pool.close()
    if rank == 0:
        w = WORKLOADS[args.workload]
        clocks = None
        for r in (sweep or [main_res]):
            clocks = merge_clocks(clocks, r["clocks"])
        e2e_src = [r for r in (sweep or [main_res]) if "e2e" in r]
        e2e = min((r["e2e"] for r in e2e_src), key=lambda e: e["value"]) if e2e_src else None
        line = {
            "metric": METRIC, "value": main_res["gbps"], "unit": "Gbps",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "snr_db": list(SNR_SWEEP) if headline else [main_res["snr_db"]],
                       "quoted_at_snr_db": main_res["snr_db"],
                       "quoted": "the slowest sweep point (throughput is data dependent)" if headline else "single point",
                       "frames_per_gpu_per_step": main_res["frames_per_gpu_per_step"],
                       "timing": "CUDA events around every step; value from the median step + amortised allreduce tail",
                       "l2": "per-step input (frames x N x 4 B >= 512 MiB) exceeds the 126 MB L2",
                       "sharding": f"frames sharded over {world} rank(s), one allreduce of error counters",
                       "launch": main_res["launch"], "fer": main_res["fer"], "ber": main_res["ber"]},
            "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "roofline": main_res["roofline"], "cpu_baseline": cpu,
            "kernel_ms": main_res["ms_kernel"], "ms_step": main_res["ms_step"], "frames_per_s": main_res["frames_per_s"],
            "value_mean_of_steps": main_res["gbps_mean_of_steps"],
        }
        if "per_rank_ms_kernel" in main_res:
            line["per_rank_ms_kernel"] = main_res["per_rank_ms_kernel"]
        for key in ("e2e_dropin", "e2e_fp16_transport"):
            hit = [r[key] for r in (sweep or [main_res]) if key in r]
            if hit:
                line[key] = min(hit, key=lambda e: e["value"])
        keep = ("workload", "desc", "snr_db", "gbps", "ms_per_step", "ms_step", "ms_kernel", "frames_per_s", "e2e",
                "e2e_dropin", "e2e_fp16_transport", "fer", "roofline", "launch", "frames_per_gpu_per_step", "cpu_baseline", "dtype",
                "per_rank_ms_kernel")
        if sweep:
            line["snr_sweep"] = [{k: r[k] for k in ("snr_db", "gbps", "ms_step", "ms_kernel", "fer", "ber") if k in r} |
                                 ({"e2e_gbps": r["e2e"]["value"]} if "e2e" in r else {}) |
                                 {"smem_frac": r["roofline"]["frac"]} for r in sweep]
        if secondary:
            line["secondary"] = [{k: s[k] for k in keep if k in s} for s in secondary]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
