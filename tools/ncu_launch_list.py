"""ncu --csv (long format) launch log -> compact launch list + average DRAM traffic per launch of one step.
usage: python tools/ncu_launch_list.py gpurun_out/launches_raw.csv profiles/rNN_bench_launches.csv [launches_per_step]"""
import csv
import json
import re
import sys


def main(src, dst, per_step=224):
    rows = {}
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        d = rows.setdefault(k, {"kernel": r["Kernel Name"], "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name == "gpu__time_duration.sum":
            d["time_ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        elif name.startswith("dram__bytes"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["r" if "read" in name else "w"] = v * mult
    ids = sorted(rows)
    short = lambda n: re.sub(r"\(.*", "", re.sub(r"void |fp4b200::|<unnamed>::|\(anonymous namespace\)::", "", n))  # noqa: E731
    with open(dst, "w") as f:
        f.write("id,kernel,grid,time_ns,dram_read_bytes,dram_write_bytes\n")
        for i in ids:
            d = rows[i]
            f.write(f'{i},"{short(d["kernel"])}","{d["grid"]}",{d.get("time_ns", 0):.0f},{d.get("r", 0):.0f},{d.get("w", 0):.0f}\n')
    gemv = [rows[i] for i in ids if "gemv_stream_kernel" in rows[i]["kernel"]]
    step = gemv[-per_step:] if len(gemv) >= per_step else gemv
    tot_t = sum(rows[i].get("time_ns", 0) for i in ids)
    out = {"launches_in_log": len(ids), "gemv_stream_launches": len(gemv),
           "gemv_share_of_logged_time": sum(d.get("time_ns", 0) for d in gemv) / max(tot_t, 1),
           "avg_traffic_bytes_per_launch": sum(d.get("r", 0) + d.get("w", 0) for d in step) / max(len(step), 1),
           "avg_time_ns_per_launch_under_ncu": sum(d.get("time_ns", 0) for d in step) / max(len(step), 1),
           "launches_averaged": len(step)}
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 224)
