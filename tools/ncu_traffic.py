"""Attribute an ncu launch list of ONE forecast step to the bench's kernel families.

    python tools/ncu_traffic.py gpurun_out/launches_dram.csv gpurun_out/step_tags.json profiles/traffic_by_family.json

The CSV comes from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
--profile-from-start off --csv python tools/profile_step.py` (tools/gpu_round2.sh); step_tags.json is written by the same
run.  Output: per family the launches, summed ncu time and the per-launch DRAM bytes (read + write) that bench.py reports as
`roofline.traffic` next to the algorithmic bytes / FLOPs."""
import csv
import io
import json
import sys


def read_launches(path):
    text = open(path).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        if r["Metric Name"].startswith("dram__bytes"):
            val *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        elif r["Metric Name"].startswith("gpu__time"):
            val *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(unit, 1e-6)
        d[r["Metric Name"]] = val
    return [launches[k] for k in sorted(launches)]


def main():
    csv_path, tags_path, out_path = sys.argv[1:4]
    launches = read_launches(csv_path)
    tags = json.load(open(tags_path))
    assert sum(n for _, n in tags) == len(launches), (sum(n for _, n in tags), len(launches))
    fam = {}
    i = 0
    for tag, n in tags:
        key = "embed" if tag.startswith("embed") else tag.split(".")[0]
        f = fam.setdefault(key, {"launches": 0, "kernels": 0, "ncu_ms": 0.0, "dram_read": 0.0, "dram_write": 0.0, "names": set()})
        f["launches"] += 1
        for L in launches[i:i + n]:
            f["kernels"] += 1
            f["ncu_ms"] += L.get("gpu__time_duration.sum", 0.0)
            f["dram_read"] += L.get("dram__bytes_read.sum", 0.0)
            f["dram_write"] += L.get("dram__bytes_write.sum", 0.0)
            f["names"].add(L["name"].split("(")[0].replace("void ", "")[:80])
        i += n
    total_ms = sum(f["ncu_ms"] for f in fam.values())
    out = {}
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ncu_ms"]):
        out[k] = {"launches": f["launches"], "ncu_ms": round(f["ncu_ms"], 4), "ncu_share": round(f["ncu_ms"] / total_ms, 4),
                  "bytes_per_launch": (f["dram_read"] + f["dram_write"]) / f["launches"],
                  "dram_read_per_launch": f["dram_read"] / f["launches"], "dram_write_per_launch": f["dram_write"] / f["launches"],
                  "kernels": sorted(f["names"]),
                  "note": f"ncu dram__bytes_read.sum + dram__bytes_write.sum averaged over the {f['launches']} launches of this "
                          f"family in one forecast step ({csv_path.split('/')[-1]})"}
    json.dump(out, open(out_path, "w"), indent=1)
    for k, v in out.items():
        print(f"{k:16s} {v['launches']:3d} launches {v['ncu_ms']:8.3f} ms ({100 * v['ncu_share']:5.1f} %)  "
              f"{v['bytes_per_launch'] / 1e6:9.1f} MB/launch")


if __name__ == "__main__":
    main()
