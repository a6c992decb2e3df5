#!/usr/bin/env python
"""profiles/roofline_counters.json from the `ncu --metrics ... --csv --log-file` passes of gpu_call_final.sh:
executed warp / thread instructions and DRAM bytes of one map (all per-frequency launches of one step), per
nominal ray-step, for bench.py's issue-bound roofline.

    python scripts/make_roofline_counters.py c4=gpurun_out/r2h_c4_counters.csv c5=gpurun_out/r2h_c5_counters.csv
"""
import collections
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def read(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    else:
        raise SystemExit(f"{path}: no ncu csv header")
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = collections.defaultdict(dict)
    for r in rows[start:]:
        if len(r) > iv and r[0].isdigit():
            try:
                per[int(r[0])][r[im]] = float(r[iv].replace(",", ""))
            except ValueError:
                pass
            per[int(r[0])]["kernel"] = r[ik]
    return [per[k] for k in sorted(per)]


def main():
    out = {}
    for arg in sys.argv[1:]:
        key, path = arg.split("=", 1)
        launches = read(path)
        w = bench.workload(key, 1)
        nominal = w["n_pix_x"] * w["n_pix_y"] * sum(p["n_steps"] for p in w["freq_params"])
        warp = sum(l["smsp__inst_executed.sum"] for l in launches)
        thr = sum(l["smsp__thread_inst_executed.sum"] for l in launches)
        dram = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in launches)
        t = [l["gpu__time_duration.sum"] for l in launches]
        assert len(launches) == w["n_freq"], (len(launches), w["n_freq"])
        out[key] = {
            "workload": f"{w['name']}/1gpu", "launches_per_map": len(launches),
            "warp_inst_per_map": warp, "warp_inst_per_nominal_ray_step": warp / nominal,
            "thread_inst_per_nominal_ray_step": thr / nominal, "dram_bytes_per_launch": int(dram),
            "dram_bytes_note": "sum over the map's per-frequency launches (one launch per frequency)",
            "per_launch_ms_under_ncu": [round(x / 1e6, 3) for x in t],
            "issue_active_pct_per_launch": [l.get("smsp__issue_active.avg.pct_of_peak_sustained_active") for l in launches],
            "warps_active_pct_per_launch": [l.get("sm__warps_active.avg.pct_of_peak_sustained_active") for l in launches],
            "registers": launches[0].get("launch__registers_per_thread"),
            "source": f"profiles/{Path(path).name} (ncu --metrics, bench.py --config {key} --steps 1 --warmup 3, the launches of the timed map)",
            "lib": bench.lib_fingerprint(),
        }
    json.dump(out, open(ROOT / "profiles" / "roofline_counters.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
