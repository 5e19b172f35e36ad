#!/usr/bin/env python
"""
Benchmark of the haplotype-transform hot path (BASELINE.json metric:
"Haplotypes.transform hap-calls/s at 1/2/4/8 B200; achieved HBM GB/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ukb|1000g|ancestry|stream]
                    [--scaling strong|weak] [--sorted-haps]
    python bench.py --impl reference ...       # the reference's CPU algorithm (oracle port)

A step = one pass of the hot path (pack kernel + transform kernel) over one synthetic cohort shard that is
already resident in HBM.  hap-calls = samples x haplotypes x 2.  With N > 1 (torchrun, one rank per GPU) the
workload's cohort is sharded over the ranks by samples (strong scaling, `dist.shard_bounds`; --scaling weak
gives every rank the whole cohort's worth of samples instead); there is no communication on the path.  The
optional all-gather of the bit-packed result (NCCL) is timed as its own leg.

The default (ukb) line also carries, under `other_workloads`, short runs of the other BASELINE configs
(1000g, ancestry, streamed PGEN input, and the ukb plan with position-sorted haplotypes).

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import hashlib
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
    # BASELINE.json configs[4] (a leading slice of it, see stream_workload)
    "stream": dict(n=200_000, p=1_000_000, H=100_000, layout="pgen_int32_chunks", n_pops=0, seed=5,
                   name="streamed GenotypesPLINK input: 200k samples x 1M variants x 100k haplotypes, 1,024-variant int32 chunks"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ukb", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: shard the workload's cohort over the ranks (strong, default) or give every rank a whole cohort (weak)")
    ap.add_argument("--samples", type=int, default=0, help="override the cohort's samples (per GPU with --scaling weak)")
    ap.add_argument("--haps", type=int, default=0, help="override the number of haplotypes")
    ap.add_argument("--e2e-samples", type=int, default=0,
                    help="samples of the host-buffer end-to-end legs over all ranks (default 65,536 PER RANK, capped by free host memory)")
    ap.add_argument("--cpu-samples", type=int, default=0,
                    help="samples of the CPU sample (default: 8192 for the cpu_baseline leg, in slices of 2048; 512 per process and step for --impl reference)")
    ap.add_argument("--cpu-procs", type=int, default=0, help="processes of --impl reference (0 = all cores, max 16)")
    ap.add_argument("--sorted-haps", action="store_true",
                    help="order the synthetic haplotypes by position, as .hap files are after `haptools index` "
                         "(default: generation order, SURVEY.md 8d)")
    ap.add_argument("--method", default="auto", choices=["auto", "direct", "staged", "local"],
                    help="transform form (engine.transform_u8)")
    ap.add_argument("--stream-chunks", type=int, default=12, help="1,024-variant chunks of the streamed workload's slice")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--no-allgather", action="store_true")
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
    Time = the slowest process's time inside the transform (inputs are prebuilt; the reference's planner loop,
    haplotypes.py:1375-1386, is NOT in it -- nor is this repo's planner in the GPU arm's engine-level number)."""
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
    name = args.workload if args.workload != "stream" else "ukb"
    w = dict(WORKLOADS[name])
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
    sample = (f"{procs} processes x {ns} samples x all {w['p']} variants x all {H} haplotypes per step (oracle port of the "
              f"numpy path, one slice per process; the reference itself is single-threaded); host has {os.cpu_count()} cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
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

    def window(self, t0, t1):
        """Clocks over the samples taken in [t0, t1] (the sampler keeps running)."""
        if not self.how:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        rows = [r for r in list(self.rows) if r[0] >= t0 and r[0] <= t1]
        reasons = sorted({x for r in rows for x in r[2]})
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None,
                "sm_max_mhz": getattr(self, "max_mhz", None), "reasons": reasons, "samples": len(rows),
                "source": self.how + ", sampled inside the timed region"}

    def stop(self):
        self._stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()


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


TRANSFORM_SOURCES = ("ht_transform_half.cu", "ht_transform.cuh", "ht_common.cuh")


def csrc_hash() -> str:
    """Identity of the sources of the kernel the roofline is reported for (the half-tile transform): the committed
    ncu capture in profiles/r2/traffic.json belongs to exactly these files."""
    h = hashlib.sha256()
    for name in TRANSFORM_SOURCES:
        h.update(name.encode())
        h.update((REPO / "haptools_b200" / "csrc" / name).read_bytes())
    return h.hexdigest()[:16]


def measured_traffic(kernel: str, alg_bytes: int):
    """dram read + write bytes per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/r2/traffic.json), scaled by the algorithmic bytes of this launch -- or None when the capture was made
    from other kernel sources than the ones running now."""
    f = REPO / "profiles" / "r2" / "traffic.json"
    if not f.exists():
        return None, "no capture committed (profiles/r2/traffic.json)"
    d = json.loads(f.read_text())
    e = d.get("kernels", {}).get(kernel)
    if e is None:
        return None, f"no capture of {kernel} in profiles/r2/traffic.json"
    if d.get("csrc_hash") != csrc_hash():
        return None, f"capture belongs to csrc {d.get('csrc_hash')}, running csrc {csrc_hash()}"
    ratio = (e["dram_read_bytes"] + e["dram_write_bytes"]) / e["algorithmic_bytes"]
    return int(alg_bytes * ratio), (f"{e['source']}: dram read {e['dram_read_bytes'] / 1e9:.2f} GB + write {e['dram_write_bytes'] / 1e9:.2f} GB "
                                     f"per launch = {ratio:.3f} x the algorithmic bytes at {e['samples']} samples, scaled linearly to this launch")


# ----------------------------------------------------------------------------------------
# GPU arm: one device-resident workload
# ----------------------------------------------------------------------------------------
class Ctx:
    pass


def verify_slices(w, H, plan, out, n, sample_offset, orc, synth, slices=((0, 32), (-32, 32))):
    """The device result of rows [s0, s0 + ns) against the oracle on regenerated inputs."""
    ok = True
    arrays, hap_label = _oracle_plan_arrays(w, H)
    key_col, key_allele, hap_off, hap_key = arrays
    cols = plan.var_col.astype(np.int64)
    for s0, ns in slices:
        s0 = max(0, n + s0) if s0 < 0 else s0
        ns = min(ns, n - s0)
        sub = synth.synth_gt(ns, w["p"], w["seed"] + 1, mode=1, sample_offset=sample_offset + s0, variants=cols)
        sub_anc = synth.synth_ancestry(ns, w["p"], w["seed"] + 2, w["n_pops"], 2000, sample_offset=sample_offset + s0, variants=cols) if w["n_pops"] else None
        want = orc.transform_from_csr(sub, np.searchsorted(cols, key_col), key_allele, hap_off, hap_key, sub_anc, hap_label)
        got = out[s0: s0 + ns, : 2 * H].cpu().numpy().reshape(ns, H, 2).view(np.bool_)
        ok = ok and bool(np.array_equal(got, want))
    return ok


def device_workload(c, w, H, n, sample_offset, steps, warmup, method_arg, keep=False):
    """Generate the shard in HBM, run warm-up + timed steps (CUDA events on the launching stream), verify two
    slices against the oracle.  Returns a dict (times are THIS rank's; the caller takes the max over ranks)."""
    torch, engine, synth, _cuda, orc = c.torch, c.engine, c.synth, c._cuda, c.orc
    dev, lib = c.dev, c._cuda.lib()
    p = w["p"]
    plan = synth.synth_plan(p, H, seed=w["seed"], n_pops=w["n_pops"], sort=w.get("sort", False))
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
    pitch = engine.out_pitch(H)
    _cuda.check(lib.ht_synth_gt(gt.data_ptr(), *gt.stride(), n, p, sample_offset, w["seed"] + 1, 1, 0, 0), "synth")
    if anc is not None:
        _cuda.check(lib.ht_synth_ancestry(anc.data_ptr(), *anc.stride(), n, p, sample_offset, w["seed"] + 2, w["n_pops"], 2000, 0), "synth_anc")
    torch.cuda.synchronize()
    anc_ptr = anc.data_ptr() if anc is not None else 0
    anc_strides = anc.stride() if anc is not None else (0, 0, 0)

    if method_arg == "auto":  # what engine.transform_u8(method="auto") runs
        method = "staged" if (engine.STAGED_AUTO and dplan.stage_lines > 0) else ("local" if dplan.use_local() else "direct")
    else:
        method = method_arg
    if method == "staged" and dplan.stage_lines == 0:
        method = "direct"
    workspace = engine.local_workspace(dplan, n, dev) if method == "local" else None

    def step():
        engine.pack(dplan, gt.data_ptr(), gt.stride(), n, planes, anc_ptr, anc_strides)
        engine.transform_u8(dplan, planes, n, out.data_ptr(), pitch, method=method, workspace=workspace)

    for _ in range(max(warmup, 3)):
        step()
    c.barrier()
    K = steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c.barrier()
    host_t0 = time.perf_counter()
    t_begin.record()
    for k in range(K):
        ev[k][0].record()
        engine.pack(dplan, gt.data_ptr(), gt.stride(), n, planes, anc_ptr, anc_strides)
        ev[k][1].record()
        engine.transform_u8(dplan, planes, n, out.data_ptr(), pitch, method=method, workspace=workspace)
        ev[k][2].record()
    t_end.record()
    c.barrier()
    host_t1 = time.perf_counter()
    total_ms = c.max_over_ranks(t_begin.elapsed_time(t_end))
    r = {
        "plan": plan, "dplan": dplan, "n": n, "n_requested": n_requested, "method": method, "sorted": bool(w.get("sort")),
        "ms_per_step": total_ms / K,
        "pack_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in ev])),
        "tr_ms": float(np.mean([e[1].elapsed_time(e[2]) for e in ev])),
        "window": (host_t0, host_t1),
        "launches_per_step": 2 if method != "local" else 1 + 2 * -(-((n + 1023) // 1024) // 64),
    }
    # launch-latency territory (the 1000G-scale config: two launches that take 0.14 ms together): the same step
    # captured in a CUDA graph, so that a replay costs one launch; the eager number above is kept next to it
    if r["ms_per_step"] < 1.0 and method != "local":
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                step()
            for _ in range(3):
                g.replay()
            c.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                g.replay()
            e1.record()
            c.barrier()
            r["graph_ms_per_step"] = c.max_over_ranks(e0.elapsed_time(e1)) / K
        except Exception as e:  # capture refused: keep the eager number
            r["graph_error"] = f"{type(e).__name__}: {e}"[:200]
    verified = verify_slices(w, H, plan, out, n, sample_offset, orc, synth)
    r["verified"] = c.all_ranks(verified)
    r["ones_fraction"] = float(out[: min(n, 4096), : 2 * H].float().mean().item())
    if keep:
        r["buffers"] = (gt, anc, planes, out)
    return r


def roofline_of(r, peak, peak_src, kernel_env):
    plan, n = r["plan"], r["n"]
    ab = plan.algorithmic_bytes(n)
    tr_gbs = ab["transform"] / (r["tr_ms"] * 1e-3) / 1e9
    pack_gbs = ab["pack"] / (r["pack_ms"] * 1e-3) / 1e9
    if r["method"] == "local":
        kernel = "and_local_bits_kernel + unpack_bits_u8_kernel"
    elif kernel_env[:1] == "t":
        kernel = "transform_u8_kernel"
    else:
        kernel = "transform_u8_half_kernel<true, 1>" if r["method"] == "staged" else "transform_u8_half_kernel<false, 1>"
    sorted_key = kernel + " (position-sorted haplotypes)"
    traffic, traffic_src = measured_traffic(sorted_key if r.get("sorted") else kernel, ab["transform"])
    if traffic is None and r.get("sorted"):
        traffic, traffic_src = measured_traffic(kernel, ab["transform"])
    step_ms = r["ms_per_step"]
    return {
        "bound": "hbm", "kernel": kernel, "achieved": tr_gbs, "peak": peak, "unit": "GB/s", "frac": tr_gbs / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "frac_of_nominal_8TBs": tr_gbs / 8000.0,
        "algorithmic_bytes_per_launch": ab["transform"], "ms_per_launch": r["tr_ms"],
        "pack": {"kernel": "pack", "achieved": pack_gbs, "frac": pack_gbs / peak, "algorithmic_bytes_per_launch": ab["pack"],
                 "ms_per_launch": r["pack_ms"]},
        "step": {"achieved": ab["total"] / (step_ms * 1e-3) / 1e9, "frac": ab["total"] / (step_ms * 1e-3) / 1e9 / peak,
                 "algorithmic_bytes": ab["total"]},
    }


def short_summary(r, H, peak, world=1):
    """One entry of `other_workloads`."""
    ab = r["plan"].algorithmic_bytes(r["n"])
    ms = r.get("graph_ms_per_step", r["ms_per_step"])
    d = {
        "value": world * r["n"] * H * 2 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "samples": r["n"],
        "haplotypes": H, "transform_form": r["method"], "pack_ms": r["pack_ms"], "transform_ms": r["tr_ms"],
        "pack_frac": ab["pack"] / (r["pack_ms"] * 1e-3) / 1e9 / peak, "transform_frac": ab["transform"] / (r["tr_ms"] * 1e-3) / 1e9 / peak,
        "step_frac": ab["total"] / (ms * 1e-3) / 1e9 / peak, "verified_against_oracle": r["verified"],
        "launches_per_step": r["launches_per_step"],
    }
    if "graph_ms_per_step" in r:
        d["eager_ms_per_step"] = r["ms_per_step"]
        d["note"] = "ms_per_step is a CUDA-graph replay of the two launches; eager launches next to it"
        d["launches_per_step"] = 1
    if r["n"] != r["n_requested"]:
        d["samples_requested"] = r["n_requested"]
    return d


def stream_workload(c, args):
    """BASELINE configs[4], bounded by chunk count: the leading `--stream-chunks` 1,024-variant
    chunks of the 200k-sample x 1M-variant config, read by a stand-in PgenReader with pgenlib's buffer contract
    (int32 (chunk, 2n), -9 = missing; haptools/data/genotypes.py:1373-1402) into page-locked ring buffers, H2D on
    a copy stream overlapped with ht_pack_i32, one transform at the end.  The haplotypes are those of the full
    config (100k over 1M variants) that lie inside the slice."""
    torch, engine, synth, stream, orc = c.torch, c.engine, c.synth, c.stream, c.orc
    from haptools_b200.plan import plan_from_refs

    w = WORKLOADS["stream"]
    n, chunk = (args.samples or w["n"]) if args.workload == "stream" else w["n"], 1024
    V = chunk * args.stream_chunks
    ref_col, ref_allele, hap_off, _ = synth.synth_haplotype_refs(w["p"], w["H"], w["seed"], sort=True)
    last = np.maximum.reduceat(ref_col, hap_off[:-1])  # position-sorted: the haplotypes inside the slice come first
    Hs = int(np.searchsorted(np.maximum.accumulate(last), V, side="left"))
    Hs = max(1, Hs)
    plan = plan_from_refs(ref_col[: hap_off[Hs]], ref_allele[: hap_off[Hs]], hap_off[: Hs + 1], V)
    rows = 64
    base = synth.synth_gt(n, rows, w["seed"] + 1, mode=1)  # (n, 64, 2): variant v of the slice = pattern row v % 64
    pool = np.ascontiguousarray(base.transpose(1, 0, 2)).reshape(rows, 2 * n).astype(np.int32)
    pool[3, 5::97] = -9  # missing calls, pgenlib's code
    tiled = engine.pinned_empty((chunk, 2 * n), np.int32)
    tiled[:] = pool[np.arange(chunk) % rows]

    filled = set()

    def fill(idx, out):
        # stands for pgenlib decoding into the caller's buffer.  Every chunk of the slice holds the same pattern
        # (1,024 is a multiple of the 64 pattern rows), so a ring buffer is written once and recycled afterwards:
        # the reader costs nothing and the measurement shows the copy / pack pipeline, not a host memcpy
        key = (out.ctypes.data, len(idx))
        if key not in filled:
            np.copyto(out, tiled[: len(idx)])
            filled.add(key)

    reader = stream.ArrayPgenReader(fill=fill, n_samples=n)
    idx = np.arange(V, dtype=np.uint32)
    best = None
    for rep in range(3):
        st = {}
        res = stream.transform_pgen_stream(reader, idx, n, plan, chunk_variants=chunk, stats=st, to_host=False)
        torch.cuda.synchronize()
        if best is None or st["packed_s"] < best["packed_s"]:
            best = st
    # parity of a slice: the oracle on the same pattern (variant v = pattern row v % 64, -9 -> 255)
    cols = plan.var_col.astype(np.int64)
    ns = 64
    pat = np.arange(len(cols)) % rows  # only referenced variants are read: the j-th of them receives pattern row j % 64
    sub = base[:ns, pat].copy()
    sub[pool[pat][:, : 2 * ns].reshape(len(cols), ns, 2).transpose(1, 0, 2) == -9] = 255
    want = orc.transform_from_csr(sub, np.searchsorted(cols, plan.key_col().astype(np.int64)), plan.key_allele,
                                  plan.hap_off.astype(np.int64), plan.hap_key.astype(np.int64))
    got = res[:ns].cpu().numpy().view(np.bool_)
    verified = bool(np.array_equal(got, want))
    wall = best["total_device_s"]
    h2d_gbs = best["h2d_bytes"] / best["h2d_s"] / 1e9
    pack_gbs = best["h2d_bytes"] / best["pack_s"] / 1e9
    full_chunks = w["p"] / chunk
    del res
    torch.cuda.empty_cache()
    return {
        "value": n * Hs * 2 / wall, "unit": UNIT, "ms": wall * 1e3, "samples": n, "variants_in_slice": V, "haplotypes_in_slice": Hs,
        "chunks": best["chunks"], "chunk_variants": chunk, "h2d_bytes": best["h2d_bytes"], "h2d_GBps": h2d_gbs,
        "pack_i32_GBps": pack_gbs, "reader_s": best["read_s"], "h2d_s": best["h2d_s"], "pack_s": best["pack_s"],
        "stream_wall_s": best["packed_s"],
        # wall = until the last chunk has been packed (the host loop itself only ENQUEUES: its end is not the end)
        "overlap_efficiency": max(best["read_s"], best["h2d_s"], best["pack_s"]) / best["packed_s"],
        "bound": "PCIe H2D of int32 alleles (4 bytes per allele call: 16x the bytes of the bit-planes they become)",
        "full_config_estimate_s": best["packed_s"] / best["chunks"] * full_chunks,
        "verified_against_oracle": verified,
        "workload": w["name"] + f"; bounded sample: its first {args.stream_chunks} chunks ({V} variants) and the {Hs} haplotypes inside them",
        # per chunk: narrow_i32_rows_kernel + pack_vm_kernel (one pack_generic_kernel with HT_I32_KERNEL=generic); + 1 transform
        "launches": best["chunks"] * (1 if os.environ.get("HT_I32_KERNEL", "")[:1] == "g" else 2) + 1,
    }


# ----------------------------------------------------------------------------------------
# end to end through the class API on host objects
# ----------------------------------------------------------------------------------------
def pcie_ceiling(c, nbytes=1 << 30):
    """What this box's host<->device links deliver with every rank copying at once: H2D alone, D2H alone and both
    together, page-locked buffers (the ceiling the end-to-end legs are judged against)."""
    torch = c.torch
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=c.dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=c.dev)
    s1, s2 = torch.cuda.Stream(c.dev), torch.cuda.Stream(c.dev)

    def run(h2d, d2h):
        c.barrier()
        t0 = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return c.max_over_ranks(time.perf_counter() - t0)

    out = {}
    for name, a in (("h2d", (True, False)), ("d2h", (False, True)), ("both", (True, True))):
        run(*a)
        t = min(run(*a) for _ in range(3))
        out[name + "_GBps_aggregate"] = c.world * nbytes * (2 if name == "both" else 1) / t / 1e9
    out["note"] = f"{c.world} rank(s) copying {nbytes >> 20} MiB each way at the same time, page-locked host memory, max time over ranks"
    return out


def e2e_legs(c, w, H, args, n_shard, sample_offset):
    """`Haplotypes.transform(gts)` -- the call a user of the reference makes -- on HOST objects: planner (cached
    after the first call), H2D of `gts.data`, kernels, the result back in a numpy array.  Once with a page-locked
    `gts.data` and once with an ordinary (pageable) one; next to them the engine-level call of round 1
    (prebuilt plan, page-locked input AND a reused page-locked result)."""
    torch, engine, synth, orc = c.torch, c.engine, c.synth, c.orc
    from haptools_b200 import integrate
    from haptools_b200.plan import plan_from_haplotypes

    p = w["p"]
    # every rank transforms its own batch at the same time (the e2e legs scale weakly: a caller with N GPUs brings N
    # batches); the batch is capped by what the host has free: page-locked + pageable input, the result twice
    per_sample = (2 * p * (2 if w["n_pops"] else 1)) * 2 + 2 * H * 3
    try:
        import psutil

        fit = int(psutil.virtual_memory().available * 0.45 / c.world / per_sample)
    except Exception:
        fit = 16_384
    want = args.e2e_samples // c.world if args.e2e_samples else 65_536
    ne = int(min(max(1024, min(want, fit) // 1024 * 1024), n_shard if c.world == 1 else max(n_shard, 1024)))
    t0 = time.perf_counter()
    haps, gts = synth.synth_objects(p, H, w["seed"], n_pops=w["n_pops"], sort=w.get("sort", False))
    objects_s = time.perf_counter() - t0

    def host_array(pinned=True):
        alloc = engine.pinned_empty if pinned else (lambda shape: np.empty(shape, dtype=np.uint8))
        if w["layout"] == "variant_major":  # the VCF reader's (p, n, 2) array seen as (n, p, 2)
            return alloc((p, ne, 2)).transpose(1, 0, 2)
        return alloc((ne, p, 2))

    def fill(dst, base):
        reps = min(ne, base.shape[0])
        for s0 in range(0, ne, reps):
            m = min(reps, ne - s0)
            dst[s0: s0 + m] = base[:m]

    reps = min(ne, 256)
    gt_pinned = host_array()
    fill(gt_pinned, synth.synth_gt(reps, p, w["seed"] + 1, mode=1, sample_offset=sample_offset))
    anc_pinned = None
    if w["n_pops"]:
        anc_pinned = host_array()
        fill(anc_pinned, synth.synth_ancestry(reps, p, w["seed"] + 2, w["n_pops"], 2000, sample_offset=sample_offset))
    gts.samples = tuple(map(str, range(ne)))
    gts.data = gt_pinned
    if w["n_pops"]:
        gts.ancestry = anc_pinned

    def timed(fn, reps=3):
        times = []
        for _ in range(reps):
            c.barrier()
            t0 = time.perf_counter()
            res = fn()
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
            del res
        return c.max_over_ranks(float(np.mean(times)))

    # cold: the first call of a process builds the plan (and uploads its tables)
    hap_list = [haps.data[k] for k in haps.data]
    t0 = time.perf_counter()
    plan = plan_from_haplotypes(hap_list, gts, ancestry=bool(w["n_pops"]))
    planner_ms = (time.perf_counter() - t0) * 1e3
    c.barrier()
    t0 = time.perf_counter()
    res = haps.transform(gts)
    torch.cuda.synchronize()
    cold_ms = c.max_over_ranks(time.perf_counter() - t0) * 1e3
    # what was computed: the first and last rows against the oracle on regenerated inputs (sample i = pattern i % reps)
    arrays, hap_label = _oracle_plan_arrays(w, H)
    key_col, key_allele, hap_off, hap_key = arrays
    cols = plan.var_col.astype(np.int64)
    m = min(32, reps)
    sub = synth.synth_gt(m, p, w["seed"] + 1, mode=1, sample_offset=sample_offset, variants=cols)
    sub_anc = synth.synth_ancestry(m, p, w["seed"] + 2, w["n_pops"], 2000, sample_offset=sample_offset, variants=cols) if w["n_pops"] else None
    want = orc.transform_from_csr(sub, np.searchsorted(cols, key_col), key_allele, hap_off, hap_key, sub_anc, hap_label)
    verified = bool(np.array_equal(res.data[:m], want))
    if ne >= reps + m:  # the pattern repeats every `reps` samples
        verified = verified and bool(np.array_equal(res.data[ne // reps * reps - reps:][:m], want))
    assert res.data.dtype == np.bool_ and res.data.shape == (ne, H, 2)
    d2h_bits = -(-2 * H // 8) * ne if res.data.nbytes >= engine.RESULT_BITS_MIN else res.data.nbytes
    del res
    warm_s = timed(lambda: haps.transform(gts))
    # engine-level (round 1's e2e): prebuilt plan, reused page-locked result
    dplan = engine.device_plan(plan, c.dev)
    out_pinned = engine.pinned_empty((ne, H, 2))
    engine.transform_host(gt_pinned, plan, ancestry=anc_pinned, out=out_pinned, dplan=dplan)
    engine_s = timed(lambda: engine.transform_host(gt_pinned, plan, ancestry=anc_pinned, out=out_pinned, dplan=dplan))
    del out_pinned
    # pageable source (what cyvcf2 / pgenlib / numpy hand the reference)
    gt_pageable = host_array(pinned=False)
    gt_pageable[:] = gt_pinned
    gts.data = gt_pageable
    if w["n_pops"]:
        anc_pageable = host_array(pinned=False)
        anc_pageable[:] = anc_pinned
        gts.ancestry = anc_pageable
    haps.transform(gts)
    pageable_s = timed(lambda: haps.transform(gts))
    calls = c.world * ne * H * 2
    h2d = engine.upload_bytes(gt_pinned, plan, anc_pinned) * c.world
    api = ("haptools_b200.data.Haplotypes.transform(gts) on host objects (the reference's API: planner -> H2D -> kernels -> "
           "numpy result allocated by the call); plan cache warm (identity-validated, integrate.py); all ranks concurrently, max time over ranks")
    e2e = {
        "value": calls / warm_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h_bits) * c.world,
        "ms": warm_s * 1e3, "samples_per_gpu": ne, "e2e_scaling": "weak (every rank its own batch of samples_per_gpu samples, all at once)",
        "api": api + "; gts.data page-locked",
        "planner_ms": planner_ms, "first_call_ms": cold_ms, "first_call_value": calls / (cold_ms * 1e-3),
        "api_overhead_ms": max(0.0, (warm_s - engine_s) * 1e3),
        "api_overhead_note": "class-API call minus the engine-level call: O(H) Python bookkeeping independent of the sample count "
                             "(haplotype list, cache validation, the result's `variants` array) plus first-touch of the result's fresh pages",
        "result_path": "bit-packed over PCIe + host widening (ht_host_submit_download_expand)" if d2h_bits != ne * H * 2 else "bytes over PCIe",
        "verified_against_oracle": c.all_ranks(verified), "synth_objects_s": objects_s, "host_copy_threads": c._cuda.lib().ht_host_workers(),
    }
    e2e_pageable = {"value": calls / pageable_s, "unit": UNIT, "ms": pageable_s * 1e3, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(d2h_bits) * c.world, "api": api + "; gts.data in ordinary (pageable) numpy memory",
                    "slowdown_vs_page_locked": pageable_s / warm_s}
    e2e_engine = {"value": calls / engine_s, "unit": UNIT, "ms": engine_s * 1e3,
                  "api": "haptools_b200.engine.transform_host with a prebuilt plan, page-locked input and a reused page-locked result "
                         "(round 1's e2e definition)"}
    return e2e, e2e_pageable, e2e_engine


def run_b200(args):
    import torch
    import torch.distributed as dist

    from haptools_b200 import _cuda, engine, stream, synth
    from haptools_b200 import dist as hdist
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

    c = Ctx()
    c.torch, c.engine, c.synth, c._cuda, c.orc, c.stream, c.dev, c.world, c.rank = torch, engine, synth, _cuda, orc, stream, dev, world, rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(flag):
        if world == 1:
            return bool(flag)
        t = torch.tensor([1.0 if flag else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() == 1.0)

    c.barrier, c.max_over_ranks, c.all_ranks = barrier, max_over_ranks, all_ranks

    peaks_file = REPO / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    if args.workload == "stream":
        s = stream_workload(c, args)
        if rank == 0:
            sampler.stop()
            line = {"metric": METRIC, "value": s["value"] * world, "unit": UNIT, "n_gpus": world, "steps": 3, "warmup": 0,
                    "ms_per_step": s["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32->bits",
                    "data": "synthetic", "config": {"workload": s["workload"]}, "stream": s, "gpu_launches": s["launches"],
                    "roofline": None, "cpu_baseline": None, "e2e": {"value": s["value"] * world, "unit": UNIT, "h2d_bytes_per_step": s["h2d_bytes"],
                                                                     "d2h_bytes_per_step": 0, "note": "the streamed workload IS an end-to-end path: host int32 chunks in, result resident on the device"}}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    w = dict(WORKLOADS[args.workload])
    w["sort"] = bool(args.sorted_haps)
    p, H = w["p"], args.haps or w["H"]
    cohort = args.samples or w["n"]
    if world > 1 and args.scaling == "strong":
        lo, hi = hdist.shard_bounds(cohort, world, rank)  # equal shards of whole 1,024-sample tiles
        n, sample_offset = hi - lo, lo
    else:
        n, sample_offset = cohort, rank * cohort
    n_local = torch.tensor([n], device=dev, dtype=torch.int64)

    # ---- the workload the metric is quoted on -------------------------------------------------------
    r = device_workload(c, w, H, n, sample_offset, args.steps, args.warmup, args.method, keep=True)
    n = r["n"]
    if world > 1:
        n_local[0] = n
        dist.all_reduce(n_local, op=dist.ReduceOp.SUM)
    n_total = int(n_local.item())
    ms_per_step = r["ms_per_step"]
    value = n_total * H * 2 / (ms_per_step * 1e-3)
    clocks = sampler.window(*r["window"]) if rank == 0 else None
    roofline = roofline_of(r, peak, peak_src, os.environ.get("HT_TRANSFORM_KERNEL", "")) if rank == 0 else None
    plan, dplan = r["plan"], r["dplan"]
    gpu_launches = r["launches_per_step"] * args.steps

    # ---- optional all-gather of the bit-packed result over NVLink (north_star) ----------
    allgather = None
    if world > 1 and not args.no_allgather:
        gt, anc, planes, out = r["buffers"]
        tiles = (n + engine.TILE - 1) // engine.TILE
        n_gather = n_total if args.scaling == "strong" else n * world  # samples of the gathered result
        per = hdist.shard_tiles(n_gather, world)
        try:
            del out
            r["buffers"] = (gt, anc, planes, None)
            torch.cuda.empty_cache()
            bits = torch.empty((tiles, H, 32, 2), dtype=torch.int32, device=dev)
            engine.transform_bits(dplan, planes, n, bits)
            gathered = None
            for _ in range(2):
                gathered = hdist.allgather_bits(bits, n_gather)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                gathered = hdist.allgather_bits(bits, n_gather)
            e1.record()
            barrier()
            ag_ms = max_over_ranks(e0.elapsed_time(e1)) / reps
            shard_bytes = per * H * 256
            # the gathered tiles of THIS rank's own shard must be what it sent
            same = bool(torch.equal(gathered[rank * per: rank * per + tiles], bits))
            allgather = {"ms": ag_ms, "bytes_per_rank_shard": shard_bytes, "gathered_bytes_per_gpu": shard_bytes * world,
                         "algbw_GBps": shard_bytes * world / (ag_ms * 1e-3) / 1e9,
                         "busbw_GBps": shard_bytes * (world - 1) / (ag_ms * 1e-3) / 1e9,
                         "nvlink5_per_direction_GBps": 900, "own_shard_intact": all_ranks(same),
                         "what": "dist.allgather_bits: NCCL all_gather_into_tensor of the bit-packed result (tiles, H, 32, 2), "
                                 "tile-major so the gather is a concatenation; CUDA events, max over ranks"}
            # ---- the same pair as ONE kernel: ht_transform_bits_gather stores every record of this rank's shard into
            # all ranks' gathered buffers over NVLink peer memory (dist.PeerGather); timed from the planes to the
            # complete gathered result on every rank, against transform_bits + NCCL all-gather for the same
            fused = None
            try:
                e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for _ in range(2):
                    engine.transform_bits(dplan, planes, n, bits)
                    gathered = hdist.allgather_bits(bits, n_gather)
                barrier()
                e2.record()
                for _ in range(reps):
                    engine.transform_bits(dplan, planes, n, bits)
                    gathered = hdist.allgather_bits(bits, n_gather)
                e3.record()
                barrier()
                pair_ms = max_over_ranks(e2.elapsed_time(e3)) / reps
                pg = hdist.PeerGather(n_gather, H)
                for _ in range(2):
                    fbits = pg.run(dplan, planes, n)
                barrier()
                e2.record()
                for _ in range(reps):
                    fbits = pg.run(dplan, planes, n)
                e3.record()
                barrier()
                fused_ms = max_over_ranks(e2.elapsed_time(e3)) / reps
                total_tiles = (n_gather + engine.TILE - 1) // engine.TILE
                same_f = bool(torch.equal(fbits[:total_tiles], gathered[:total_tiles]))
                barrier()
                fused = {"ms": fused_ms, "transform_bits_plus_nccl_allgather_ms": pair_ms, "speedup": pair_ms / fused_ms,
                         "busbw_GBps": shard_bytes * (world - 1) / (fused_ms * 1e-3) / 1e9,
                         "equals_nccl_result": all_ranks(same_f),
                         "what": "dist.PeerGather: ht_transform_bits_gather (AND + record stores into every rank's buffer over "
                                 "NVLink peer memory, CUDA IPC) + a stream-ordered one-element all-reduce as the closing barrier; "
                                 "CUDA events from the planes to the complete gathered result, max over ranks"}
                del fbits
                pg.close()
            except Exception as e:  # noqa: BLE001 -- peer memory may be unavailable on a box (no P2P / IPC);
                # PeerGather raises on ALL ranks together, so nobody is left waiting in a collective
                fused = {"skipped": f"{type(e).__name__}: {e}"[:300]}
            allgather["fused_kernel"] = fused
            del gathered, bits
        except torch.cuda.OutOfMemoryError:
            allgather = {"skipped": "not enough HBM next to the resident shard"}
        torch.cuda.empty_cache()
    r.pop("buffers", None)
    torch.cuda.empty_cache()

    # ---- short runs of the other BASELINE configs (inherit this run's clock record) --------------------
    others = None
    if args.workload == "ukb" and world == 1 and not args.no_others and not args.sorted_haps:
        others = {}
        t_others0 = time.perf_counter()
        ws = dict(WORKLOADS["ukb"], sort=True)
        rs = device_workload(c, ws, H, n, sample_offset, 5, 3, args.method)
        others["ukb_sorted"] = short_summary(rs, H, peak)
        others["ukb_sorted"]["note"] = ("the same plan with the haplotypes in position order -- what a .hap file is after `haptools index` "
                                        "(haptools/index.py:31-94); neighbouring haplotypes share plane records, which the half-tile kernel "
                                        "picks up as L1 hits")
        gpu_launches += rs["launches_per_step"] * 5
        del rs
        torch.cuda.empty_cache()
        rst = device_workload(c, ws, H, n, sample_offset, 3, 3, "staged")
        others["ukb_sorted"]["staged_form"] = {
            "transform_ms": rst["tr_ms"], "transform_frac": rst["plan"].algorithmic_bytes(rst["n"])["transform"] / (rst["tr_ms"] * 1e-3) / 1e9 / peak,
            "verified_against_oracle": rst["verified"], "stage_lines": rst["dplan"].stage_lines,
            "note": "ht_transform_u8_staged (key runs of the haplotype blocks in shared memory): bit-exact, 2.1 x instead of 3.1 x the plane "
                    "bytes from L2, not faster -- method='auto' does not pick it (engine.STAGED_AUTO)"}
        gpu_launches += rst["launches_per_step"] * 3
        del rst
        torch.cuda.empty_cache()
        for name in ("1000g", "ancestry"):
            wo = dict(WORKLOADS[name], sort=False)
            ro = device_workload(c, wo, wo["H"], wo["n"], 0, 10 if name == "1000g" else 5, 3, args.method)
            others[name] = short_summary(ro, wo["H"], peak)
            others[name]["workload"] = wo["name"]
            gpu_launches += ro["launches_per_step"] * (10 if name == "1000g" else 5)
            del ro
            torch.cuda.empty_cache()
        others["stream"] = stream_workload(c, args)
        gpu_launches += others["stream"]["launches"] * 3
        if rank == 0:
            others["clocks"] = sampler.window(t_others0, time.perf_counter())

    # ---- end to end through the class API -------------------------------------------------------------
    e2e = e2e_pageable = e2e_engine = None
    if not args.no_e2e:
        e2e, e2e_pageable, e2e_engine = e2e_legs(c, w, H, args, n, sample_offset)
        e2e["cpu_binding_rank0"] = numa
        e2e["pcie_ceiling"] = pcie_ceiling(c)
        moved = (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) / (e2e["ms"] * 1e-3) / 1e9
        e2e["host_link_GBps_aggregate"] = moved
        e2e["pcie_ceiling_gbs"] = e2e["pcie_ceiling"]["h2d_GBps_aggregate"]

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    sampler.stop()

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

    ab = plan.algorithmic_bytes(n)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling if world > 1 else "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {
            "workload": w["name"], "cohort_samples": n_total, "samples_per_gpu": n, "variants": p, "haplotypes": H,
            **({"samples_per_gpu_requested": r["n_requested"], "note": "shard shrunk: the requested one did not fit in HBM"} if n != r["n_requested"] else {}),
            "haplotype_alleles": plan.n_refs, "distinct_keys": plan.n_keys, "layout": w["layout"], "transform_form": r["method"],
            "haplotype_order": "position-sorted" if w["sort"] else "generation order (unsorted)",
            "genotypes": "blocky founders (16 per 64-variant block), device-generated from a counter hash",
            "l2": "inputs (%.1f GB) larger than L2, no flush needed" % (ab["gt"] / 1e9),
            "sharding": ("the cohort's samples in equal shards of whole 1,024-sample tiles, one shard per GPU, no collective on the path"
                         if world > 1 and args.scaling == "strong" else "one whole cohort per GPU, no collective on the path"),
        },
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_pageable": e2e_pageable, "e2e_engine": e2e_engine,
        "allgather": allgather, "other_workloads": others,
        "gpu_launches": gpu_launches, "clocks": clocks,
        "verified_against_oracle": r["verified"], "ones_fraction": r["ones_fraction"],
        **({"graph_ms_per_step": r["graph_ms_per_step"], "launches_per_step_in_graph": 1} if "graph_ms_per_step" in r else {}),
        "launches_per_step": r["launches_per_step"], "csrc_hash": csrc_hash(),
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
