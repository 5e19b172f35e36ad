#!/usr/bin/env python
"""Turn one kernel of an `ncu -i X.ncu-rep --page raw --csv` dump into an entry of profiles/r2/traffic.json
(the per-launch DRAM traffic bench.py reports as `roofline.traffic`, scaled by algorithmic bytes).

    python profiles/make_traffic.py raw.csv KEY SAMPLES ALGORITHMIC_BYTES SOURCE [PLANE_BYTES]

KEY is the name the entry gets (the kernel's template instance as bench.py names it); the csrc hash of the
transform sources (bench.csrc_hash) is refreshed, so run it for every transform kernel after a change to them.
"""
import csv, json, sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def main():
    raw, key, samples, alg, source = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    plane = int(sys.argv[6]) if len(sys.argv) > 6 else None
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}

    def get(name, as_bytes=False):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        return v * scale[units[i]] if as_bytes else v

    e = {
        "samples": samples, "algorithmic_bytes": alg,
        "dram_read_bytes": get("dram__bytes_read.sum", True), "dram_write_bytes": get("dram__bytes_write.sum", True),
        "l2_to_sm_read_bytes": get("lts__t_sectors_srcunit_tex_op_read.sum") * 32,
        "gpu_time_ms_under_ncu": get("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]],
        "registers": int(get("launch__registers_per_thread")), "warp_instructions": get("smsp__inst_executed.sum"),
        "dram_throughput_pct_of_ncu_peak": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "source": source,
    }
    if plane:
        e["plane_bytes"] = plane
        e["l2_to_sm_read_over_plane_bytes"] = e["l2_to_sm_read_bytes"] / plane
    e["traffic_over_algorithmic"] = (e["dram_read_bytes"] + e["dram_write_bytes"]) / alg
    f = REPO / "profiles" / "r2" / "traffic.json"
    d = json.loads(f.read_text()) if f.exists() else {"kernels": {}}
    d["kernels"][key] = e
    import bench

    d["csrc_hash"] = bench.csrc_hash()
    f.write_text(json.dumps(d, indent=1))
    print(json.dumps(e, indent=1))


if __name__ == "__main__":
    main()
