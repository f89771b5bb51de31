#!/usr/bin/env python
"""profiles/ncu_traffic.json <- measured DRAM traffic per unit of the dominant kernels, from `ncu --set full` captures of
this code (bench.py reads the file for `roofline.traffic`; nothing is hard-coded there).
  python tools/ncu_traffic.py <capture tag> <config> lin=<rep> <tasks per launch> qp=<rep> <QPs per launch>"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dram_bytes(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, r = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    def val(k):
        v = float(r[idx[k]].replace(",", ""))
        u = units[idx[k]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
    return val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), r[idx["Kernel Name"]], float(r[idx["gpu__time_duration.sum"]].replace(",", "")), units[idx["gpu__time_duration.sum"]]


def main():
    tag, cfg = sys.argv[1], sys.argv[2]
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    ent = data.setdefault(cfg, {})
    a = sys.argv[3:]
    for i in range(0, len(a), 2):
        kind, rep = a[i].split("=")
        units = float(a[i + 1])
        b, kname, dur, du = dram_bytes(rep)
        key = "dram_bytes_per_task" if kind == "lin" else "dram_bytes_per_qp"
        ent[kind] = {"kernel": kname.split("(")[0], key: round(b / units, 1), "units_per_launch": units, "dram_bytes_per_launch": b,
                     "launch_duration": "%g %s (under ncu)" % (dur, du), "capture": tag, "source": os.path.basename(rep)}
    json.dump(data, open(path, "w"), indent=1)
    print(json.dumps(ent))


if __name__ == "__main__":
    main()
