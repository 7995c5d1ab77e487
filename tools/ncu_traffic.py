"""DRAM traffic of one captured kernel launch: dram__bytes_read.sum + dram__bytes_write.sum from an
`ncu --set full` report -> profiles/<name>.json, which bench.py quotes as roofline.traffic.
usage: python tools/ncu_traffic.py report.ncu-rep profiles/r1_k2_traffic.json"""
import csv
import json
import subprocess
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    head, unit, val = rows[0], rows[1], rows[2]
    col = {c: i for i, c in enumerate(head)}

    def get(name):
        i = col[name]
        return float(val[i]) * SCALE.get(unit[i], 1.0)

    d = {"report": rep.split("/")[-1], "kernel": val[col["Kernel Name"]], "grid": val[col["launch__grid_size"]],
         "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
         "duration_ms_under_ncu": float(val[col["gpu__time_duration.sum"]]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[unit[col["gpu__time_duration.sum"]]]}
    d["traffic_bytes"] = d["dram_bytes_read"] + d["dram_bytes_write"]
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d))


if __name__ == "__main__":
    main()
