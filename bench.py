#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 batched Viterbi decoder (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

A "step" is one pass of the hot path (ingest -> add-compare-select -> traceback -> result gather) over one batch of synthetic
BPSK-over-AWGN frames (viterbidecodercpp_b200/synth.py, the statistics of the reference's run_snr_ber generator).
`value`  : decoded Mbit/s with the soft symbols already resident in HBM (device-pointer C ABI call, CUDA events).
`e2e`    : the same metric through the host-pointer C ABI call (pinned host buffers; H2D and D2H inside the timed region).
`roofline`: the dominant kernel (add-compare-select) against the measured HBM copy bandwidth; `roofline_alu` against the
            packed-integer issue roofline the north star names.
`cpu_baseline`: the reference's own AVX2 decoder (oracle/_ref, built from /root/reference in place) on all host threads.
--impl reference times that same CPU implementation as the reference arm.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# BASELINE.json configs.  cfg2 is the configuration the metric is quoted on.
WORKLOADS = {
    "cfg1": dict(code="Voyager", decode="SOFT16", frames=16384, bits=1024, ebno=4.0, desc="K=7 R=1/2 u16 soft, 16384 x 1024-bit frames"),
    "cfg2": dict(code="Voyager", decode="HARD8", frames=65536, bits=2048, ebno=4.0, desc="K=7 R=1/2 u8 hard, 65536 x 2048-bit frames"),
    "cfg3": dict(code="CDMA IS-95A", decode="SOFT16", frames=16384, bits=8192, ebno=4.0, desc="K=9 R=1/2 u16 soft, 16384 x 8192-bit frames"),
    "cfg4": dict(code="DAB Radio", decode="SOFT16", frames=65536, bits=768, ebno=4.0, punctured=True,
                 desc="DAB K=7 R=1/4 u16 soft punctured FIC frames, 65536 x 768-bit"),
    "cfg5": dict(code="Cassini", decode="SOFT16", frames=1024, bits=16384, ebno=4.0, desc="Cassini K=15 R=1/6 u16 soft, 1024 x 16384-bit frames"),
    "run_simple": dict(code="DAB Radio", decode="SOFT16", frames=8192, bits=8192, ebno=None, desc="run_simple: K=7 R=1/4 u16 soft, 8192-bit frames"),
}


def read_traffic(kernel_name, frames):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` captures (profiles/r02_traffic.json), scaled
    to this run's frame count; None when no capture exists for this kernel variant.  A call that used two kernels (K = 15 wave
    split) is looked up by the first, which decodes most of the frames."""
    kernel_name = kernel_name.split(" for ")[0]
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    e = t.get(kernel_name)
    if not e:
        return None
    return (e["dram_read_bytes"] + e["dram_write_bytes"]) * frames / e["frames"]


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def gpu_numa_cpus(torch, local_rank):
    """(node, cpus) of the NUMA node this rank's GPU hangs off, from sysfs; (None, None) when that cannot be determined"""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return None, None
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        return (node, cpus) if cpus else (None, None)
    except Exception:
        return None, None


def build_workload(name, seed, n_frames=None):
    import viterbidecodercpp_b200 as v
    from viterbidecodercpp_b200 import synth
    w = dict(WORKLOADS[name])
    if n_frames:
        w["frames"] = n_frames
    code = {c.name: c for c in v.COMMON_CODES}[w["code"]]
    dc = v.DECODE_TYPES[w["decode"]](code.R)
    F, L = w["frames"], w["bits"]
    dtype = np.int8 if dc.soft_bytes == 1 else np.int16
    keep = np.asarray(v.dab_fic_keep_schedule(), dtype=bool) if w.get("punctured") else None
    # generated in slabs to bound host memory; every frame is unique
    tx_all, sym_all = [], []
    slab = 4096
    for f0 in range(0, F, slab):
        n = min(slab, F - f0)
        tx, sym = synth.make_frames(code.K, code.R, code.G, n, L, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes,
                                    w["ebno"], seed + f0)
        if keep is not None:
            sym = synth.puncture(sym, keep)
        tx_all.append(tx); sym_all.append(sym.astype(dtype))
    w.update(code_obj=code, dc=dc, tx=np.concatenate(tx_all), sym=np.concatenate(sym_all), keep=keep)
    return w


def algorithmic_bytes(w):
    """per frame  symbols read + decision bits written (ACS kernel)  [+ traceback read + output + 16 for the pipeline]"""
    code, dc = w["code_obj"], w["dc"]
    L, S, N = w["bits"], w["bits"] + code.K - 1, 1 << (code.K - 1)
    sym = w["sym"].shape[1] * dc.soft_bytes
    dec = S * N // 8
    return {"acs_kernel": sym + dec + 16, "pipeline": sym + dec + 4 * L + L // 8 + 16, "acs_ops": S * N}


def reference_inputs(w):
    """what the reference decodes: the depunctured stream for punctured workloads (zeros inserted; same trellis work)"""
    sym = w["sym"]
    if w["keep"] is not None:
        dep = np.zeros((sym.shape[0], w["keep"].size), dtype=sym.dtype)
        dep[:, w["keep"]] = sym
        sym = dep
    return sym


def run_reference(args, rank):
    """reference arm: the reference's AVX2 decoder (unmodified headers, oracle/_ref/libvitref.so) on all host threads, on the SAME
    batch the GPU arm decodes (same generator, seed and frame count).  One step = one pass over the batch; the W warm-up passes and
    the K timed passes each run inside ONE harness call, so the pinned worker threads (one ViterbiDecoder_Core each) persist."""
    if rank != 0:
        return
    import oracle_binding as ob
    if not ob.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libvitref.so missing (built from /root/reference by oracle/Makefile)"}))
        return
    w = build_workload(args.workload, 1234, args.frames or None)
    code, dc, cfg = w["code_obj"], w["dc"], w["dc"].decoder_config
    sym = reference_inputs(w)
    threads = ob.ref_lib().vitref_host_threads()
    cfgl = [cfg.soft_decision_max_error, cfg.initial_start_error, cfg.initial_non_start_error, cfg.renormalisation_threshold]

    def run(passes):
        return ob.ref_decode(code.K, code.R, code.G, dc.soft_bytes, dc.soft_decision_high, dc.soft_decision_low, cfgl, ob.IMPL_AVX,
                             sym, sym.shape[0], w["bits"], n_threads=threads, passes=passes)["seconds"]
    if args.warmup:
        run(args.warmup)
    t = run(max(args.steps, 1))
    mbit = sym.shape[0] * w["bits"] / t / 1e6
    ab = algorithmic_bytes(w)
    line = {
        "impl": "reference", "metric": "decoded_mbit_per_s", "value": mbit, "unit": "Mbit/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8" if dc.soft_bytes == 1 else "u16", "data": "synthetic BPSK/AWGN (run_snr_ber statistics), fixed seed, every frame unique",
        "config": {"workload": f"{args.workload}: {w['desc']}", "frames_per_gpu": int(sym.shape[0]), "bits_per_frame": w["bits"],
                   "EbNo_dB": w["ebno"]},
        "acs_gops": sym.shape[0] * ab["acs_ops"] / t / 1e9,
        "cpu_baseline": {"value": mbit, "unit": "Mbit/s", "cores": threads, "kind": "reference",
                         "sample": f"the whole batch: {sym.shape[0]} frames x {w['bits']} bits per step, ViterbiDecoder_AVX_{'u8' if dc.soft_bytes == 1 else 'u16'}, "
                                   f"reset+update+chainback, {threads} pinned threads kept alive over the {args.steps} timed passes"},
        "e2e": {"value": mbit, "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def strong_cfg5(torch, dist, v, world, rank, local_rank, dev, steps=3, warmup=1):
    """BASELINE.json config 5 at 1/2/4/8 GPUs, STRONG scaling: 1024 Cassini K=15 R=1/6 frames of 16384 bits in total, 1024 / N per
    rank, soft symbols generated on the device (vitb_synth_frames_dev, Eb/N0 4 dB) so that no host generator sits in the way.
    Timed like the main leg: CUDA events around `steps` device-resident batch calls, barrier on both sides, max over ranks."""
    code = {c.name: c for c in v.COMMON_CODES}["Cassini"]
    dc = v.DECODE_TYPES["SOFT16"](code.R)
    from viterbidecodercpp_b200 import sharding
    total, L = 1024, 16384
    f0, f1 = sharding.frame_range(rank, world, total)          # the split tests/test_host_cpu.py checks under gloo
    F = f1 - f0
    row = (L + code.K - 1) * code.R
    bt = v.ViterbiBranchTable(code.K, code.R, code.G, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes)
    dec = v.ViterbiDecoder_CUDA(bt, dc.decoder_config, device=local_rank)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    d_sym = torch.empty((F, row), dtype=torch.int16, device=dev)
    d_tx = torch.empty((F, L // 8), dtype=torch.uint8, device=dev)
    d_out = torch.zeros((F, L // 8), dtype=torch.uint8, device=dev)
    d_acc = torch.zeros(F, dtype=torch.int64, device=dev)
    d_fin = torch.zeros(F, dtype=torch.int32, device=dev)
    dec.synth_frames_dev(F, L, d_tx.data_ptr(), d_sym.data_ptr(), EbNo_dB=4.0, seed=777 + f0, row_stride=row, stream=sptr)

    def step():
        dec.decode_batch_dev(d_sym.data_ptr(), F, L, d_out.data_ptr(), d_acc.data_ptr(), d_fin.data_ptr(), stream=sptr, row_stride=row)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    byte_errors = int((d_out != d_tx).sum().item())
    dec.set_profiling(True)
    step()
    torch.cuda.synchronize()
    st = dec.stage_ms()
    dec.set_profiling(False)
    ms = sharding.max_over_ranks(ms, dev)                       # slowest rank
    acs_ms = sharding.max_over_ranks(st["acs"], dev)
    byte_errors = int(sharding.sum_over_ranks(byte_errors, dev))
    name = dec.kernel_name
    dec.close()
    del d_sym, d_tx, d_out
    torch.cuda.empty_cache()
    ops = total * (L + code.K - 1) * (1 << (code.K - 1))
    return {"workload": "cfg5: Cassini K=15 R=1/6 u16 soft, 1024 x 16384-bit frames in total (strong scaling)", "scaling": "strong",
            "n_gpus": world, "frames_total": total, "frames_per_gpu": F, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "value": total * L / (ms * 1e-3) / 1e6, "unit": "Mbit/s", "acs_ms": acs_ms, "acs_gops": ops / (ms * 1e-3) / 1e9,
            "byte_errors": byte_errors, "kernel": name, "data": "device-generated BPSK/AWGN at 4 dB (Philox), inputs resident in HBM"}


def multi_in_process(torch, v, world, w, steps=3, warmup=1):
    """the north star's "per-GPU streams and a host scatter" inside ONE process: vitb_decode_batch_multi with one handle per device,
    each device decoding a full batch of the workload from one pinned host array (weak scaling, like the torchrun leg).  The call is
    synchronous (it returns when every device has finished), so it is timed with the host clock around it."""
    code, dc = w["code_obj"], w["dc"]
    F, L = w["sym"].shape[0], w["bits"]
    bt = v.ViterbiBranchTable(code.K, code.R, code.G, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes)
    decs = [v.ViterbiDecoder_CUDA(bt, dc.decoder_config, device=d) for d in range(world)]
    for d in decs:
        if w["keep"] is not None:
            d.set_puncture_schedule(w["keep"].astype(np.uint8), 0)
    h_sym = torch.from_numpy(np.concatenate([w["sym"]] * world)).pin_memory()
    n = F * world
    h_out = torch.zeros((n, (L + 7) // 8), dtype=torch.uint8).pin_memory()
    h_acc = torch.zeros(n, dtype=torch.int64).pin_memory()
    h_fin = torch.zeros(n, dtype=torch.int32).pin_memory()

    def step():
        v.decode_batch_multi_raw(decs, h_sym.data_ptr(), n, L, h_out.data_ptr(), h_acc.data_ptr(), h_fin.data_ptr(), row_stride=w["sym"].shape[1])
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    ms = (time.perf_counter() - t0) / steps * 1e3
    ber = float(np.unpackbits(h_out.numpy()[-F:] ^ w["tx"]).mean())
    for d in decs:
        d.close()
    return {"call": "vitb_decode_batch_multi, one handle per device, pinned host memory, one host thread", "n_devices": world,
            "frames_total": n, "steps": steps, "ms_per_step": ms, "value": n * L / (ms * 1e-3) / 1e6, "unit": "Mbit/s",
            "timing": "host clock around the synchronous call", "ber_last_device": ber}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="override the number of frames (debug)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=0, help="pin the lanes-per-frame-pair kernel variant (0 = automatic)")
    ap.add_argument("--window", type=int, default=0, help="K=15: sliding-window traceback with this many bits per window (0 = exact mode)")
    ap.add_argument("--no-strong", action="store_true", help="skip the config-5 strong-scaling leg (1024 Cassini frames over all GPUs)")
    ap.add_argument("--no-pipelining", action="store_true", help="time the device-resident leg with the stages of a batch in sequence")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import viterbidecodercpp_b200 as v
    from viterbidecodercpp_b200 import sharding

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- workload: every rank decodes a full batch of its own frames (weak scaling, no data-path collective) ----
    w = build_workload(args.workload, 1234 + 1000003 * rank, args.frames or None)
    code, dc = w["code_obj"], w["dc"]
    F, L = w["sym"].shape[0], w["bits"]
    bt = v.ViterbiBranchTable(code.K, code.R, code.G, dc.soft_decision_high, dc.soft_decision_low, dc.soft_bytes)
    dec = v.ViterbiDecoder_CUDA(bt, dc.decoder_config, device=local_rank)
    if w["keep"] is not None:
        dec.set_puncture_schedule(w["keep"].astype(np.uint8), 0)
    if args.lanes:
        dec.set_variant(args.lanes)
    if args.window:
        dec.set_traceback_window(args.window)

    out_stride = (L + 7) // 8
    # pinned host buffers of the end-to-end leg: allocated (first touch) while this thread sits on the NUMA node of its GPU, so that
    # with several ranks per box every GPU pulls its symbols from local memory instead of across the socket interconnect
    affinity0 = os.sched_getaffinity(0)
    numa_node, numa_cpus = gpu_numa_cpus(torch, local_rank)
    if numa_cpus:
        os.sched_setaffinity(0, numa_cpus)
    h_sym = torch.from_numpy(w["sym"]).pin_memory()
    h_out = torch.zeros((F, out_stride), dtype=torch.uint8).pin_memory()
    h_acc = torch.zeros(F, dtype=torch.int64).pin_memory()
    h_fin = torch.zeros(F, dtype=torch.int32).pin_memory()
    if numa_cpus:
        os.sched_setaffinity(0, affinity0)          # the CPU-baseline leg wants every host thread again
    d_sym = h_sym.to(dev)
    d_out = torch.zeros((F, out_stride), dtype=torch.uint8, device=dev)
    d_acc = torch.zeros(F, dtype=torch.int64, device=dev)
    d_fin = torch.zeros(F, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def step_dev():
        dec.decode_batch_dev(d_sym.data_ptr(), F, L, d_out.data_ptr(), d_acc.data_ptr(), d_fin.data_ptr(), stream=sptr,
                             row_stride=w["sym"].shape[1])

    def step_e2e():
        dec.decode_batch_async(h_sym.data_ptr(), F, L, h_out.data_ptr(), h_acc.data_ptr(), h_fin.data_ptr(), stream=sptr,
                               row_stride=w["sym"].shape[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, after=None):
        """exactly `steps` steps between two CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks;
        `after` (the flush of the pipelined calls) runs inside the timed region"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        if after:
            after()
        e1.record(stream)
        barrier()
        return sharding.max_over_ranks(e0.elapsed_time(e1), dev) / steps       # max over ranks

    def stage_breakdown(steps):
        """average device time of each pipeline stage over `steps` more steps (CUDA events recorded by the library on the same stream)"""
        stages = {"ingest": 0.0, "acs": 0.0, "traceback": 0.0, "gather": 0.0}
        dec.set_profiling(True)
        for _ in range(steps):
            step_dev()
            torch.cuda.synchronize()
            for name, ms in dec.stage_ms().items():
                stages[name] += ms
        dec.set_profiling(False)
        return {k2: v2 / steps for k2, v2 in stages.items()}

    # ---- warm-up, then the timed region (inputs 269 MB >> 126 MB L2, so no explicit L2 flush is needed for cfg2) ----
    # `value`: K back-to-back batch calls with the library's pipelining on (vitb_set_pipelining: the traceback of batch i runs on the
    # handle's second stream next to the add-compare-select of batch i+1; vitb_batch_flush inside the timed region waits for the
    # last one), i.e. the sustained rate of a caller that decodes one batch after another.  `ms_per_step_serial` is the same call
    # with its stages in sequence (latency of one isolated batch).
    pipelined = not args.no_pipelining
    for _ in range(args.warmup):
        step_dev()
    ms_serial = timed(step_dev, args.steps)
    dec.set_pipelining(pipelined)
    for _ in range(args.warmup):
        step_dev()
    dec.batch_flush(sptr)
    launches0 = dec.kernel_launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_step = timed(step_dev, args.steps, after=lambda: dec.batch_flush(sptr))
    launches = dec.kernel_launch_count - launches0
    dec.set_pipelining(False)
    kernel_name = dec.kernel_name
    stages = stage_breakdown(min(args.steps, 10))      # kernel-level times for the roofline, same inputs, live in this run

    # sanity: the timed path really decoded the frames (bit error rate against the transmitted bytes)
    torch.cuda.synchronize()
    ber = float(np.unpackbits(d_out.cpu().numpy() ^ w["tx"]).mean())

    for _ in range(2):
        step_e2e()
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = timed(step_e2e, e2e_steps)
    clocks = sampler.stop()
    torch.cuda.synchronize()
    assert (h_out.numpy() == d_out.cpu().numpy()).all(), "host-pointer path and device-pointer path disagree"

    # ceiling of the end-to-end leg: the same bytes moved by plain cudaMemcpyAsync (pinned host <-> device), no kernels, all ranks at
    # once, timed the same way.  e2e / this = how much of the host link the library's pipelined path reaches.
    def step_copy():
        d_sym.copy_(h_sym, non_blocking=True)
        h_out.copy_(d_out, non_blocking=True)
        h_acc.copy_(d_acc, non_blocking=True)
        h_fin.copy_(d_fin, non_blocking=True)
    step_copy()
    ms_copy = timed(step_copy, e2e_steps)

    # single-frame streaming calls (the reference's own call protocol; run_punctured_decoder calls update once per trellis step):
    # host time per vitb_update of one step, and of one whole frame in one call
    streaming = None
    if rank == 0:
        one = np.ascontiguousarray(w["sym"][0] if w["keep"] is None else reference_inputs(w)[0])
        dec.set_traceback_length(L)
        dec.reset()
        R_ = code.R
        n_small = min(200, one.size // R_)
        dec.update(one[:R_])
        dec.reset()
        t0 = time.perf_counter()
        for k in range(n_small):
            dec.update(one[k * R_:(k + 1) * R_])
        t_small = (time.perf_counter() - t0) / n_small
        dec.reset()
        dec.update(one)
        dec.reset()
        t0 = time.perf_counter()
        dec.update(one)
        err = dec.get_error()
        out1 = dec.chainback(L)
        t_frame = time.perf_counter() - t0
        assert (out1 == d_out[0].cpu().numpy()).all(), "streaming calls and batch call disagree on frame 0"
        streaming = {"update_one_step_us": t_small * 1e6, "whole_frame_update_get_error_chainback_ms": t_frame * 1e3,
                     "note": "host clock, synchronous calls through the C ABI (calls of up to 64 KB of symbols: staged in mapped pinned memory, ingest kernel + ACS kernel + one synchronize, the result read from mapped memory; larger calls: H2D copy, the two kernels, D2H copy)"}

    strong = None if args.no_strong else strong_cfg5(torch, dist, v, world, rank, local_rank, dev)
    multi = None
    if world > 1 and not args.no_strong:
        # rank 0 drives every device from one process; the other ranks wait ON THE HOST (a key in the rendezvous store): a NCCL
        # barrier would park a spinning kernel on their GPUs for the whole leg
        barrier()
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                multi = multi_in_process(torch, v, world, w)
            except Exception as e:
                multi = {"failed": str(e)}
            store.set("multi_in_process_done", "1")
        else:
            store.wait(["multi_in_process_done"])
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src, sm_max = read_peaks()
    ab = algorithmic_bytes(w)
    bits_total = world * F * L
    value = bits_total / (ms_step * 1e-3) / 1e6
    acs_ms = stages["acs"]
    achieved_gbs = F * ab["acs_kernel"] / (acs_ms * 1e-3) / 1e9
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    acs_gacs = F * ab["acs_ops"] / (acs_ms * 1e-3) / 1e9
    # Issue-slot roofline of the kernel that ran (the binding bound for every configuration: DESIGN.md section 3): one
    # warp-instruction per clock per SM sub-partition (32 lanes) at the kernel's own MINIMUM instruction count per add-compare-select
    # of one frame: survivor-history kernels 4 per butterfly = 1.0 (uint8 metrics, two frames per register) / 2.0 (uint16 metrics,
    # one frame per register); decision-row kernels 10 per butterfly of two frames = 2.5.
    if kernel_name.startswith("acs_hist"):
        ipa = 1.0 if dc.soft_bytes == 1 else 2.0
    else:
        ipa = 2.5
    issue_peak_gacs = n_sm * 4 * 32 * sm_max * 1e6 / ipa / 1e9
    # What the two integer pipes sustain on exactly this instruction mix, measured on the B200 with every operand in its own register
    # (profiles/microbench/r02_pipe_rates3.txt: VIADD.16x2 : VIADDMNMX.U16x2 = 1:1 issues at 1.46 clocks per warp-instruction, any
    # number of warps; profiles/microbench/r02_hist_mix3.txt: 6.3 clocks per butterfly = 4 instructions): the register file feeds
    # about 1.65 operands per clock and the butterfly reads 10.
    mix_clk = 1.46 if dc.soft_bytes == 1 else 1.49
    line = {
        "metric": "decoded_mbit_per_s", "value": value, "unit": "Mbit/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8" if dc.soft_bytes == 1 else "u16", "data": "synthetic BPSK/AWGN (run_snr_ber statistics), fixed seed, every frame unique",
        "config": {"workload": f"{args.workload}: {w['desc']}", "frames_per_gpu": F, "bits_per_frame": L, "EbNo_dB": w["ebno"],
                   "l2": "inputs larger than L2 (no flush)" if w["sym"].nbytes > 130e6 else "inputs smaller than L2; decision buffer larger than L2",
                   "kernel": kernel_name, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                   "pipelining": "traceback of batch i next to the ACS of batch i+1 (vitb_set_pipelining), flush inside the timed region" if pipelined else "off",
                   "host_numa_node": numa_node, "traceback_window_bits": args.window, "workspace_bytes": dec.workspace_bytes(F, L)},
        "ms_per_step_serial": ms_serial,
        "acs_gops": world * F * ab["acs_ops"] / (ms_step * 1e-3) / 1e9,
        "ber": ber,
        "stage_ms": stages,
        "roofline": {"bound": "issue", "achieved": acs_gacs, "peak": issue_peak_gacs, "unit": "GACS/s", "frac": acs_gacs / issue_peak_gacs,
                     "traffic": read_traffic(kernel_name, F), "kernel": "add-compare-select", "kernel_ms": acs_ms,
                     "definition": f"1 warp-instruction per clock per SM sub-partition at {sm_max:.0f} MHz, {ipa} instructions per add-compare-select of one frame (this kernel's minimum)",
                     "attainable_on_this_mix": {"clk_per_instruction": mix_clk, "frac": acs_gacs / issue_peak_gacs * mix_clk,
                                                "source": "profiles/microbench/r02_pipe_rates3.txt, r02_hist_mix3.txt (register-file bound: ~1.65 operand reads per clock)"}},
        "roofline_hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                         "traffic": read_traffic(kernel_name, F), "kernel": "add-compare-select", "peak_source": peak_src,
                         "algorithmic_bytes_per_frame": ab["acs_kernel"], "kernel_ms": acs_ms},
        "pipeline_hbm": {"achieved": F * ab["pipeline"] / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": F * ab["pipeline"] / (ms_step * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_frame": ab["pipeline"]},
        "e2e": {"value": bits_total / (ms_e2e * 1e-3) / 1e6, "unit": "Mbit/s", "h2d_bytes_per_step": int(w["sym"].nbytes),
                "d2h_bytes_per_step": int(h_out.numel() + h_acc.numel() * 8 + h_fin.numel() * 4), "ms_per_step": ms_e2e},
        "e2e_roofline": {"bound": "host link: the step's H2D + D2H bytes by plain cudaMemcpyAsync from/to pinned memory, all ranks at once, same run",
                         "copy_ms_per_step": ms_copy, "h2d_gbs_per_gpu": w["sym"].nbytes / (ms_copy * 1e-3) / 1e9,
                         "h2d_gbs_all_gpus": world * w["sym"].nbytes / (ms_copy * 1e-3) / 1e9, "frac": ms_copy / ms_e2e},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if streaming is not None:
        line["streaming"] = streaming
    if strong is not None:
        line["strong_cfg5"] = strong
    if multi is not None:
        line["multi_in_process"] = multi

    if world == 1 and not args.no_cpu_baseline:
        try:
            import oracle_binding as ob
            if ob.have_ref():
                cfg = dc.decoder_config
                cfgl = [cfg.soft_decision_max_error, cfg.initial_start_error, cfg.initial_non_start_error, cfg.renormalisation_threshold]
                sym = reference_inputs(w)
                threads = ob.ref_lib().vitref_host_threads()
                n_cpu = min(sym.shape[0], {"cfg3": 2048, "cfg5": 32}.get(args.workload, sym.shape[0]))

                def cpu_run(passes):
                    return ob.ref_decode(code.K, code.R, code.G, dc.soft_bytes, dc.soft_decision_high, dc.soft_decision_low, cfgl,
                                         ob.IMPL_AVX, sym[:n_cpu], n_cpu, L, n_threads=threads, passes=passes)
                r = cpu_run(1)                                # untimed first pass: page in, and the bytes for the mismatch count
                reps = int(min(50, max(2, np.ceil(2.0 / max(r["seconds"], 1e-3)))))      # >= 2 s on all threads, pinned, threads kept alive
                t_total = cpu_run(reps)["seconds"] * reps
                cpu_mbit = reps * n_cpu * L / t_total / 1e6
                mism = int((r["bytes"] != d_out.cpu().numpy()[:n_cpu]).any(axis=1).sum())
                line["cpu_baseline"] = {"value": cpu_mbit, "unit": "Mbit/s", "cores": threads, "kind": "reference",
                                        "sample": f"{n_cpu} frames x {L} bits, {reps} passes in one call, reference AVX2 decoder (reset+update+chainback), "
                                                  f"{threads} pinned threads kept alive; AVX2 tie-break differs from the scalar oracle: {mism}/{n_cpu} frames "
                                                  f"decode to different bytes than the GPU (scalar-exact) output",
                                        "acs_gops": reps * n_cpu * ab["acs_ops"] / t_total / 1e9}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "Mbit/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
        except Exception as e:   # the baseline must never take the bench line down with it
            line["cpu_baseline"] = {"value": None, "unit": "Mbit/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
