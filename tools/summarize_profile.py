"""Turn the ncu outputs of tools/gpu_profile_r1*.sh into the committed summaries under profiles/.

    python tools/summarize_profile.py launches gpurun_out/launches_c.csv <forwards> > profiles/<name>.md
    python tools/summarize_profile.py full gpurun_out/prof_r1c.ncu-rep profiles/<traffic>.json > profiles/<name>.md

`launches`: per-kernel launch counts / time / share of one forward from the `--metrics gpu__time_duration.sum` launch
list (the csv holds <forwards> forwards incl. warm-up and the e2e leg).  `full`: per kernel shape the mean duration, DRAM
bytes and pipe utilisations of the `--set full` capture; also writes the DRAM traffic json bench.py reads for
`roofline.traffic`.
"""
import collections
import csv
import io
import json
import subprocess
import sys


def short(name):
    name = name.replace("pangu::", "").replace("(anonymous namespace)::", "")
    cut = name.find("(")
    return (name[:cut] if cut > 0 else name)[:70]


def launches(path, forwards):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1000.0 if r[iu] in ("ns", "nsecond") else (v * 1000.0 if r[iu] in ("ms", "msecond") else v)  # -> us
        a = agg.setdefault(short(r[ik]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print("| kernel | launches / forward | us / forward | share |\n|---|---|---|---|")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %.1f | %.0f | %.1f %% |" % (k, c / forwards, us / forwards, 100.0 * us / tot))
    print("\nSum per forward: %.2f ms under ncu (%d launches = %g forwards)." % (tot / forwards / 1000.0, n, forwards))


METRICS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor"),
           ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts"),
           ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1"),
           ("dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram_r"),
           ("dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "dram_w"),
           ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue"),
           ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu"),
           ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3,
         "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}


def full(rep, traffic_json):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    ik = h.index("Kernel Name")
    groups = collections.OrderedDict()
    for r in rows[2:]:
        rec = {}
        for m, key in METRICS:
            if m not in h:
                continue
            j = h.index(m)
            try:
                rec[key] = float(r[j].replace(",", "")) * SCALE.get(units[j], 1.0)
            except ValueError:
                rec[key] = float("nan")
        groups.setdefault((short(r[ik]), int(rec.get("grid", 0)), round(rec["us"] / max(rec["us"], 1e-9))), []).append(rec)
    # split launches of one kernel name by similar duration (two shapes of the same template differ ~2x)
    shapes = []
    for (name, grid, _), recs in groups.items():
        recs.sort(key=lambda x: x["us"])
        cur = [recs[0]]
        for x in recs[1:]:
            if x["us"] > 1.35 * cur[0]["us"]:
                shapes.append((name, grid, cur))
                cur = []
            cur.append(x)
        shapes.append((name, grid, cur))
    print("| kernel (launches) | us | DRAM rd MB | DRAM wr MB | tensor pipe % | issue % | XU % | LTS % | L1TEX % | DRAM % | SM % | regs | grid x block |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    traffic = collections.OrderedDict()
    for name, grid, recs in shapes:
        mean = lambda k: sum(x.get(k, float("nan")) for x in recs) / len(recs)
        print("| `%s` (%d) | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %d x %d |" % (
            name, len(recs), mean("us"), mean("rd") / 1e6, mean("wr") / 1e6, mean("tensor"), mean("issue"), mean("xu"),
            mean("lts"), mean("l1"), mean("dram_r") + mean("dram_w"), mean("sm"), mean("regs"), grid, mean("block")))
        traffic.setdefault(name, []).append({"grid": grid, "us": mean("us"), "dram_read_bytes": mean("rd"), "dram_write_bytes": mean("wr")})
    if traffic_json:
        json.dump({"source": "ncu --set full, %s (tools/summarize_profile.py)" % rep, "kernels": traffic}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], float(sys.argv[3]))
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
