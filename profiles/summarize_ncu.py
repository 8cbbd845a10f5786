#!/usr/bin/env python
"""Turn ncu reports into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_summary.md
    python profiles/summarize_ncu.py --launches gpurun_out/launches.csv          > profiles/rNN_launches.md

Reads the report with `ncu -i ... --page raw --csv` / `--page source --csv` (no GPU needed).  Per kernel: duration,
tensor / XU / FMA / ALU pipe utilisation, issue-slot utilisation, DRAM bytes and throughput, L2 hit rate, registers,
warp-stall breakdown and the ten instructions with the most stall samples.
"""
import csv
import io
import subprocess
import sys
from collections import Counter, defaultdict

RAW_KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "L2 sector throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_summary(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"\n### {d['Kernel Name']}  (launch id {d.get('ID', '?')}, {rep.split('/')[-1]})\n")
        print("| metric | value |\n|---|---|")
        for k, label in RAW_KEYS:
            if k in d and d[k] != "":
                print(f"| {label} | {d[k]} {u.get(k, '')} |")
        stalls = {k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): float(d[k] or 0)
                  for k in hdr if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")}
        tot = sum(stalls.values()) or 1.0
        top = sorted(stalls.items(), key=lambda x: -x[1])[:6]
        print("| warp stalls (share of warp-cycles) | " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top) + " |")


def source_summary(rep, n=10):
    rows = ncu_csv(rep, "source", ("--print-source", "sass"))
    # one block per kernel: a "Kernel Name" line, a header line, then instructions
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = dict(name=r[1], rows=[])
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        if len(b["rows"]) < 2:
            continue
        hdr = b["rows"][0]
        idx = {h: i for i, h in enumerate(hdr)}
        if "# Samples" not in idx:
            continue
        stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        body = [r for r in b["rows"][1:] if len(r) == len(hdr)]
        total = sum(int(r[idx["# Samples"]]) for r in body) or 1
        mix = Counter()
        for r in body:
            toks = [t for t in r[idx["Source"]].split() if not t.startswith("@")]
            if toks:
                mix[toks[0].split(".")[0]] += int(r[idx["Instructions Executed"]])
        print(f"\n#### {b['name']}: instructions with the most stall samples ({total} samples)\n")
        print("| samples | executed | SASS | top stall reasons |\n|---|---|---|---|")
        for r in sorted(body, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
            s = sorted(((h[6:], int(r[idx[h]])) for h in stall if int(r[idx[h]]) > 0), key=lambda x: -x[1])[:2]
            print(f"| {100 * int(r[idx['# Samples']]) / total:.1f}% | {r[idx['Instructions Executed']]} | `{r[idx['Source']].strip()[:60]}` | "
                  + ", ".join(f"{k} {v}" for k, v in s) + " |")
        tot_i = sum(mix.values()) or 1
        print("\ninstruction mix: " + ", ".join(f"{k} {100 * v / tot_i:.1f}%" for k, v in mix.most_common(12)))


def launches(path, top=30):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    idx = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        name = r[idx["Kernel Name"]].split("(")[0]
        per[name][0] += 1
        per[name][1] += us
    total = sum(v[1] for v in per.values()) or 1.0
    print(f"launch list `{path.split('/')[-1]}`: {sum(v[0] for v in per.values())} launches, {total / 1e3:.2f} ms of kernel time "
          "(ncu: cold caches, serialised -- compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, us) in sorted(per.items(), key=lambda x: -x[1][1])[:top]:
        print(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f}% |")


def traffic(reps, points, out_path):
    """profiles/r02_traffic.json: dram__bytes_read + dram__bytes_write per point of every kernel in the reports, stamped with
    the hash of the kernel sources (bench.source_sha) -- bench.py prints roofline.traffic only while that hash matches."""
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import source_sha
    kernels = {}
    for rep in reps:
        rows = ncu_csv(rep, "raw")
        hdr = rows[0]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0].replace("rnb::", "").replace("_kernel", "")
            try:
                rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            except (KeyError, ValueError):
                continue
            units = dict(zip(hdr, rows[1]))

            def to_bytes(v, u):
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            tot = to_bytes(rd, units["dram__bytes_read.sum"]) + to_bytes(wr, units["dram__bytes_write.sum"])
            k = kernels.setdefault(name, dict(launches=0, dram_bytes=0.0))
            k["launches"] += 1
            k["dram_bytes"] += tot
    out = dict(source_sha=source_sha(), points=points, capture=", ".join(os.path.basename(r) for r in reps), kernels={})
    for name, k in kernels.items():
        out["kernels"][name] = dict(dram_bytes_per_launch=k["dram_bytes"] / k["launches"],
                                    dram_bytes_per_point=k["dram_bytes"] / k["launches"] / points, launches_captured=k["launches"])
    json.dump(out, open(out_path, "w"), indent=1)
    print("wrote", out_path, {k: round(v["dram_bytes_per_point"]) for k, v in out["kernels"].items()})


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--traffic":      # --traffic POINTS out.json rep [rep ...]
        traffic(args[3:], int(args[1]), args[2])
    elif args and args[0] == "--launches":
        for p in args[1:]:
            launches(p)
    else:
        for rep in args:
            raw_summary(rep)
            source_summary(rep)
