"""profiles/r02_traffic.json from `ncu --set full` captures of ONE pressure pass (raw CSV pages): DRAM bytes per pass =
dram__bytes_read.sum + dram__bytes_write.sum summed over the kernels of the pass (the streaming kernel and the frame launch
of the tiled kernel), keyed to the md5 of the kernel sources so that bench.py only quotes it for the build it was taken from.

  python tools/ncu_traffic.py --entry "cavity 8192x8192 T=4" --source profiles/r02_ncu_full_ppe_stream.md gpurun_out/x_stream_raw.csv gpurun_out/x_frame_raw.csv
"""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def dram_bytes(path):
    rows = list(csv.reader(open(path)))
    hdr, units, first = rows[0], rows[1], rows[2]
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(m)
        tot += float(first[i].replace(",", "")) * UNIT[units[i]]
    return tot, first[hdr.index("Kernel Name")].split("(")[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entry", required=True)
    ap.add_argument("--source", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_traffic.json"))
    ap.add_argument("csv", nargs="+")
    a = ap.parse_args()
    import bench
    t = {"_comment": "DRAM bytes per pressure pass from ncu --set full, summed over the kernels of the pass; see tools/ncu_traffic.py",
         "kernel_source_md5": bench.kernel_source_id(), "entries": {}}
    if os.path.exists(a.out):
        old = json.load(open(a.out))
        if old.get("kernel_source_md5") == t["kernel_source_md5"]:
            t["entries"] = old.get("entries", {})
    parts, tot = [], 0.0
    for p in a.csv:
        b, name = dram_bytes(p)
        parts.append({"kernel": name, "dram_bytes": b})
        tot += b
    t["entries"][a.entry] = {"dram_bytes_per_launch": tot, "kernels": parts, "source": a.source}
    json.dump(t, open(a.out, "w"), indent=1)
    print(json.dumps(t["entries"][a.entry]))


if __name__ == "__main__":
    main()
