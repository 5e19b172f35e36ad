#!/usr/bin/env python
"""
Benchmark of the haplotype-transform hot path (BASELINE.json metric:
"Haplotypes.transform hap-calls/s at 1/2/4/8 B200; achieved HBM GB/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ukb|1000g|ancestry]
    python bench.py --impl reference ...       # the reference's CPU algorithm (oracle port)

A step = one pass of the hot path (pack kernel + transform kernel) over one synthetic cohort
shard that is already resident in HBM.  hap-calls = samples x haplotypes x 2.  With N > 1
(torchrun, one rank per GPU) every rank owns its own shard of samples (weak scaling: the
per-GPU shard is fixed, the cohort grows with N); there is no communication on the path.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "Haplotypes.transform hap-calls/s"
UNIT = "hap-calls/s"

WORKLOADS = {
    # BASELINE.json configs[2]: the config the metric is quoted on (sample-sharded 1/2/4/8 GPUs)
    "ukb": dict(n=500_000, p=50_000, H=100_000, layout="sample_major", n_pops=0, seed=2,
                name="synthetic UKB-scale: 500k samples x 50k variants x 100k haplotypes (2-20 alleles)"),
    # BASELINE.json configs[1]
    "1000g": dict(n=2_504, p=1_000_000, H=10_000, layout="variant_major", n_pops=0, seed=1,
                  name="synthetic 1000G-scale: 2,504 samples x 1M variants x 10k haplotypes (2-20 alleles)"),
    # BASELINE.json configs[3]
    "ancestry": dict(n=100_000, p=200_000, H=50_000, layout="sample_major", n_pops=5, seed=4,
                     name="synthetic ancestry-aware: 100k samples x 200k variants x 50k haplotypes, 5 populations"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ukb", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="override samples per GPU")
    ap.add_argument("--haps", type=int, default=0, help="override the number of haplotypes")
    ap.add_argument("--e2e-samples", type=int, default=16_384, help="samples of the host-buffer end-to-end leg")
    ap.add_argument("--cpu-samples", type=int, default=0,
                    help="samples of the CPU sample (default: 8192 for the cpu_baseline leg, in slices of 2048; 512 per process and step for --impl reference)")
    ap.add_argument("--cpu-procs", type=int, default=0, help="processes of --impl reference (0 = all cores, max 16)")
    ap.add_argument("--sorted-haps", action="store_true",
                    help="order the synthetic haplotypes by position, as .hap files are after `haptools index` "
                         "(default: generation order, SURVEY.md 8d)")
    ap.add_argument("--method", default="auto", choices=["auto", "direct", "local"],
                    help="transform form (engine.transform_u8): one kernel, two-phase local, or the engine's choice")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on a bounded sample of the same workload
# ----------------------------------------------------------------------------------------
def _cpu_slice(job):
    """Regenerate a sample slice on the host and run the oracle on it; returns seconds."""
    from haptools_b200 import synth
    from oracle import transform_oracle as orc

    (w, plan_arrays, s0, ns, hap_label) = job
    key_col, key_allele, hap_off, hap_key = plan_arrays
    cols = np.unique(key_col)
    gt = synth.synth_gt(ns, w["p"], w["seed"] + 1, mode=1, sample_offset=s0, variants=cols)
    anc = synth.synth_ancestry(ns, w["p"], w["seed"] + 2, w["n_pops"], 2000, sample_offset=s0, variants=cols) if w["n_pops"] else None
    kc = np.searchsorted(cols, key_col)
    t0 = time.perf_counter()
    out = orc.transform_from_csr(gt, kc, key_allele, hap_off, hap_key, anc, hap_label)
    dt = time.perf_counter() - t0
    return dt, int(out.sum())


def _oracle_plan_arrays(w, H):
    """Haplotype set in the oracle's (reference-like) form: one key per distinct
    (variant, allele), per-haplotype population label kept apart."""
    from haptools_b200 import synth
    from haptools_b200.plan import plan_from_refs

    ref_col, ref_allele, hap_off, hap_label = synth.synth_haplotype_refs(
        w["p"], H, w["seed"], n_pops=w["n_pops"], sort=w.get("sort", False))
    plain = plan_from_refs(ref_col, ref_allele, hap_off, w["p"])
    arrays = (plain.key_col().astype(np.int64), plain.key_allele, plain.hap_off.astype(np.int64), plain.hap_key.astype(np.int64))
    return arrays, (hap_label.astype(np.uint8) if hap_label is not None else None)


def cpu_baseline(w, H, n_samples, procs=1, reps=1):
    """hap-calls/s of the oracle port on `procs` processes, each on its own n_samples slice.
    Time = the slowest process's time inside the transform (inputs are prebuilt, like the
    reference's benchmark script does, tests/bench_transform.py:219-227)."""
    arrays, hap_label = _oracle_plan_arrays(w, H)
    jobs = [(w, arrays, 10_000 * i, n_samples, hap_label) for i in range(procs)]
    best = None
    for _ in range(reps):
        if procs == 1:
            res = [_cpu_slice(jobs[0])]
        else:
            import multiprocessing as mp

            with mp.get_context("fork").Pool(procs) as pool:
                res = pool.map(_cpu_slice, jobs)
        wall = max(r[0] for r in res)
        best = wall if best is None else min(best, wall)
    calls = procs * n_samples * H * 2
    return calls / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = dict(WORKLOADS[args.workload])
    w["sort"] = bool(args.sorted_haps)
    H = args.haps or w["H"]
    procs = args.cpu_procs or min(os.cpu_count() or 1, 16)
    ns = args.cpu_samples or 512
    for _ in range(args.warmup):
        cpu_baseline(w, H, max(64, ns // 8), procs=procs)
    times = []
    for _ in range(args.steps):
        v, wall = cpu_baseline(w, H, ns, procs=procs)
        times.append(wall)
    wall = float(np.mean(times))
    value = procs * ns * H * 2 / wall
    sample = f"{procs} processes x {ns} samples x all {w['p']} variants x all {H} haplotypes per step (oracle port of the numpy path, one slice per process)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": w["name"], "samples_per_step": procs * ns, "haplotypes": H, "variants": w["p"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an in-process NVML poller
    (no start-up delay, 20 ms period; the main thread sits in cudaStreamSynchronize without the
    GIL meanwhile), `nvidia-smi -lms` as the fallback.  Samples carry a host timestamp and
    stop(t0, t1) keeps those inside the timed window."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.how = index, [], None, None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {
                "hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap,
            }

            def poll():
                while not self._stop.is_set():
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append((time.perf_counter(), mhz, [k for k, b in bits.items() if mask & b]))
                    except Exception:
                        pass
                    self._stop.wait(0.02)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.how = "nvml"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            self.how = "nvidia-smi"
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                self.max_mhz = float(r[1])
                self.rows.append((time.perf_counter(), float(r[0]),
                                  [nm for nm, val in zip(self.NAMES, r[2:6]) if val.lower().startswith("active")]))
            except Exception:
                pass

    def stop(self, t0=None, t1=None):
        self._stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if not self.how:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)]
        reasons = sorted({x for r in rows for x in r[2]})
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None,
                "sm_max_mhz": getattr(self, "max_mhz", None), "reasons": reasons, "samples": len(rows),
                "source": self.how + ", sampled inside the timed region"}


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Pin this process to the CPUs next to its GPU (sysfs), so that the pinned host buffers of the
    end-to-end leg are first-touched on the GPU's own NUMA node.  Best effort."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        cpulist = Path(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read_text().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{bdf}: cpus {cpulist}"
    except Exception as e:  # sysfs not visible, no such attribute, ...
        return f"not bound ({type(e).__name__})"
    return "not bound"


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from haptools_b200 import _cuda, engine, synth
    from oracle import transform_oracle as orc  # checker + cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    engine.require_cuda()
    numa = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = dict(WORKLOADS[args.workload])
    n = args.samples or w["n"]  # samples per GPU (weak scaling)
    p, H = w["p"], args.haps or w["H"]
    lib = _cuda.lib()
    sample_offset = rank * n

    # ---- inputs, resident in HBM before the timed region ------------------------------
    w["sort"] = bool(args.sorted_haps)
    plan = synth.synth_plan(p, H, seed=w["seed"], n_pops=w["n_pops"], sort=w["sort"])
    dplan = engine.DevicePlan(plan, dev)
    def alloc(n):
        if w["layout"] == "sample_major":
            gt = torch.empty((n, p, 2), dtype=torch.uint8, device=dev)
        else:
            gt = torch.empty((p, n, 2), dtype=torch.uint8, device=dev).permute(1, 0, 2)
        anc = torch.empty((n, p, 2), dtype=torch.uint8, device=dev) if w["n_pops"] else None
        planes = engine.alloc_planes(n, plan.n_keys, dev)
        out = torch.empty((n, engine.out_pitch(H)), dtype=torch.uint8, device=dev)
        return gt, anc, planes, out

    n_requested = n
    while True:  # the full UKB-scale shard needs 162.5 GB of HBM: shrink it rather than die
        try:
            gt, anc, planes, out = alloc(n)
            break
        except torch.cuda.OutOfMemoryError:
            gt = anc = planes = out = None
            torch.cuda.empty_cache()
            if n <= 8192:
                raise
            n = (n // 2 + 1023) // 1024 * 1024
            sample_offset = rank * n
    pitch = engine.out_pitch(H)
    _cuda.check(lib.ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, sample_offset, w["seed"] + 1, 1, 0, 0), "synth")
    if anc is not None:
        _cuda.check(lib.ht_synth_ancestry(anc.data_ptr(), *anc.stride(), n, p, sample_offset, w["seed"] + 2, w["n_pops"], 2000, 0), "synth_anc")
    torch.cuda.synchronize()

    anc_ptr = anc.data_ptr() if anc is not None else 0
    anc_strides = anc.stride() if anc is not None else (0, 0, 0)

    two_phase = args.method == "local" or (args.method == "auto" and dplan.use_local())
    workspace = engine.local_workspace(dplan, n, dev) if two_phase else None
    method = "local" if two_phase else "direct"

    def step():
        engine.pack(dplan, gt.data_ptr(), gt.stride(), n, planes, anc_ptr, anc_strides)
        engine.transform_u8(dplan, planes, n, out.data_ptr(), pitch, method=method, workspace=workspace)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching (current) stream ----------
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    host_t0 = time.perf_counter()
    t_begin.record()
    for k in range(K):
        ev[k][0].record()
        engine.pack(dplan, gt.data_ptr(), gt.stride(), n, planes, anc_ptr, anc_strides)
        ev[k][1].record()
        engine.transform_u8(dplan, planes, n, out.data_ptr(), pitch, method=method, workspace=workspace)
        ev[k][2].record()
    t_end.record()
    barrier()
    clocks = sampler.stop(host_t0, time.perf_counter()) if rank == 0 else None
    total_ms = t_begin.elapsed_time(t_end)
    pack_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    tr_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / K
    calls_per_gpu = n * H * 2
    value = world * calls_per_gpu / (ms_per_step * 1e-3)

    # ---- correctness of what was just timed (outside the timed region) -----------------
    verified = True
    for s0 in (0, max(0, n - 32)):
        ns = min(32, n - s0)
        cols = plan.var_col.astype(np.int64)
        sub = synth.synth_gt(ns, p, w["seed"] + 1, mode=1, sample_offset=sample_offset + s0, variants=cols)
        arrays, hap_label = _oracle_plan_arrays(w, H)
        key_col, key_allele, hap_off, hap_key = arrays
        sub_anc = synth.synth_ancestry(ns, p, w["seed"] + 2, w["n_pops"], 2000, sample_offset=sample_offset + s0, variants=cols) if w["n_pops"] else None
        want = orc.transform_from_csr(sub, np.searchsorted(cols, key_col), key_allele, hap_off, hap_key, sub_anc, hap_label)
        got = out[s0 : s0 + ns, : 2 * H].cpu().numpy().reshape(ns, H, 2).view(np.bool_)
        verified = verified and bool(np.array_equal(got, want))
    ones_fraction = float(out[: min(n, 4096), : 2 * H].float().mean().item())
    if world > 1:  # every rank checked its own shard
        t = torch.tensor([1.0 if verified else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        verified = bool(t.item() == 1.0)

    # ---- end to end through the host-buffer API (H2D + kernels + D2H inside the timing) --
    # every rank runs it at the same time (each GPU has its own PCIe link; host memory is
    # shared), the time is the max over ranks
    e2e = None
    if not args.no_e2e:
        ne = min(args.e2e_samples, n)
        del out, planes, workspace
        torch.cuda.empty_cache()
        def host_array():
            # the workload's memory layout: variant-major = the VCF reader's (p, n, 2) array seen as (n, p, 2)
            if w["layout"] == "variant_major":
                return engine.pinned_empty((p, ne, 2)).transpose(1, 0, 2)
            return engine.pinned_empty((ne, p, 2))

        host_gt = host_array()
        host_gt[:] = synth.synth_gt(min(ne, 256), p, w["seed"] + 1, mode=1, sample_offset=sample_offset)[np.arange(ne) % min(ne, 256)]
        host_anc = None
        if w["n_pops"]:
            host_anc = host_array()
            host_anc[:] = synth.synth_ancestry(min(ne, 256), p, w["seed"] + 2, w["n_pops"], 2000, sample_offset=sample_offset)[np.arange(ne) % min(ne, 256)]
        host_out = engine.pinned_empty((ne, H, 2))
        for _ in range(2):
            engine.transform_host(host_gt, plan, ancestry=host_anc, out=host_out, dplan=dplan)
        times = []
        for _ in range(max(3, min(K, 5))):
            barrier()
            t0 = time.perf_counter()
            res = engine.transform_host(host_gt, plan, ancestry=host_anc, out=host_out, dplan=dplan)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        dt = float(np.mean(times))
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {
            "value": world * ne * H * 2 / dt, "unit": UNIT,
            "h2d_bytes_per_step": engine.upload_bytes(host_gt, plan, host_anc) * world,
            "d2h_bytes_per_step": int(res.nbytes) * world,
            "samples_per_gpu": ne, "ms": dt * 1e3, "cpu_binding_rank0": numa,
            "api": "haptools_b200.engine.transform_host (the call Haplotypes.transform makes) on pinned host arrays; "
                   "all ranks concurrently, max time over ranks",
        }

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (transform) ------------------------------------
    peaks_file = REPO / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ab = plan.algorithmic_bytes(n)
    tr_gbs = ab["transform"] / (tr_ms * 1e-3) / 1e9
    pack_gbs = ab["pack"] / (pack_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": ("transform_u8_kernel" if os.environ.get("HT_TRANSFORM_KERNEL", "")[:1] == "t" else "transform_u8_half_kernel") if not two_phase else "and_local_bits_kernel + unpack_bits_u8_kernel", "achieved": tr_gbs, "peak": peak, "unit": "GB/s",
        "frac": tr_gbs / peak, "traffic": int(ab["transform"] * 1.027), "peak_source": peak_src,
        "traffic_source": "ncu --set full (profiles/r1/ncu_half_tile_details_transform_u8_half.txt): dram read 4.12 GB + write 26.16 GB per launch = 1.027 x the algorithmic 29.49 GB at 131,072 samples, scaled linearly to this launch",
        "l2_note": "the kernel is bound by L2 throughput, not HBM: 37.1 GB of record reads + 26.2 GB of result writes cross L2<->SM per launch at 131,072 samples = 10.2 TB/s, against 9.2-9.3 TB/s for the same read:write mix with no arithmetic at all (profiles/micro/l2bw.cu, profiles/r1/micro_l2bw.txt)",
        "algorithmic_bytes_per_launch": ab["transform"], "ms_per_launch": tr_ms,
        "pack": {"kernel": "pack", "achieved": pack_gbs, "frac": pack_gbs / peak, "algorithmic_bytes_per_launch": ab["pack"], "ms_per_launch": pack_ms},
        "step": {"achieved": ab["total"] / (ms_per_step * 1e-3) / 1e9, "frac": ab["total"] / (ms_per_step * 1e-3) / 1e9 / peak,
                 "algorithmic_bytes": ab["total"]},
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        # about 10 s of CPU work, as slices of 2,048 samples one after the other so that the host
        # never holds more than one slice's equality array (1.6 GB at the UKB-scale plan)
        ns_total = args.cpu_samples or 8192
        ns = min(ns_total, 2048)
        n_slices = max(1, min(ns_total // ns, -(-n // ns)))  # never more slices than the shard has samples for
        arrays, hap_label = _oracle_plan_arrays(w, H)
        wall = sum(_cpu_slice((w, arrays, 10_000 * i, ns, hap_label))[0] for i in range(n_slices))
        v = n_slices * ns * H * 2 / wall
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{n_slices} x {ns} samples x all {p} variants x all {H} haplotypes, oracle/transform_oracle.py (numpy restatement of the reference), {wall:.1f} s inside the transform; host has {os.cpu_count()} cores"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {
            "workload": w["name"], "samples_per_gpu": n, "variants": p, "haplotypes": H,
            **({"samples_per_gpu_requested": n_requested, "note": "shard shrunk: the requested one did not fit in HBM"} if n != n_requested else {}),
            "haplotype_alleles": plan.n_refs, "distinct_keys": plan.n_keys, "layout": w["layout"], "transform_form": method,
            "haplotype_order": "position-sorted" if w["sort"] else "generation order (unsorted)",
            "genotypes": "blocky founders (16 per 64-variant block), device-generated from a counter hash",
            "l2": "inputs (%.1f GB) larger than L2, no flush needed" % (ab["gt"] / 1e9),
            "sharding": "samples, one shard per GPU, no collective on the path",
        },
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": (1 + (2 * -(-((n + 1023) // 1024) // 64) if two_phase else 1)) * K, "clocks": clocks,
        "verified_against_oracle": verified, "ones_fraction": ones_fraction,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    # libraries (NCCL's version banner, ...) must not pollute stdout: the contract is ONE JSON
    # line there.  Route fd 1 to stderr while running and print the line on the real stdout.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import builtins

    def emit(*a, **kw):
        kw.pop("flush", None)
        builtins.print(*a, file=real_stdout, flush=True, **kw)

    globals()["print"] = emit
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
