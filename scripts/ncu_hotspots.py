#!/usr/bin/env python
"""Summarise an ncu report (captured with --set full --import-source on) into the JSON kept under profiles/:
headline metrics of every launch plus, per launch, the SASS opcode histogram (share of executed warp
instructions and of stall samples) and the hottest straight-line regions.

    python scripts/ncu_hotspots.py gpurun_out/r2_attention_f16_kernel.ncu-rep > profiles/r2_attention_f16_hotspots.json
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
           "launch__block_size", "launch__registers_per_thread", "sm__cycles_elapsed.max",
           "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warp_latency_per_inst_issued.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True, check=True).stdout


def launches(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:100]}
        for m in METRICS:
            if m in hdr:
                d[m] = f"{r[hdr.index(m)]} {units[hdr.index(m)]}".strip()
        out.append(d)
    return out


def sass(rep):
    secs, cur = [], None
    for row in csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))):
        if row and row[0] == "Kernel Name":
            cur = {"hdr": None, "rows": [], "name": row[1] if len(row) > 1 else ""}
            secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    out = []
    for sec in secs:
        ix = {h: i for i, h in enumerate(sec["hdr"])}
        ie, isamp, isrc = ix["Instructions Executed"], ix["# Samples"], ix["Source"]
        tot_i = sum(int(r[ie]) for r in sec["rows"]) or 1
        tot_s = sum(int(r[isamp]) for r in sec["rows"]) or 1
        hi, hs = defaultdict(int), defaultdict(int)
        regions, reg = [], None
        for n, r in enumerate(sec["rows"]):
            t = r[isrc].strip().split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            e, s = int(r[ie]), int(r[isamp])
            hi[op] += e
            hs[op] += s
            if reg and reg["exec"] == e:
                reg["n"] += 1
                reg["samples"] += s
                reg["last"] = n
            else:
                reg = {"exec": e, "n": 1, "samples": s, "first": n, "last": n}
                regions.append(reg)
        out.append({
            "kernel": sec["name"][:100], "warp_instructions": tot_i, "stall_samples": tot_s,
            "opcodes_pct": {op: {"inst": round(v / tot_i * 100, 2), "samples": round(hs[op] / tot_s * 100, 2)}
                            for op, v in sorted(hi.items(), key=lambda kv: -kv[1])[:16]},
            "hot_regions": [dict(r, inst_share_pct=round(r["exec"] * r["n"] / tot_i * 100, 2))
                            for r in regions if r["exec"] * r["n"] / tot_i > 0.01],
        })
    return out


if __name__ == "__main__":
    rep = sys.argv[1]
    ls = launches(rep)
    try:  # reports captured without --import-source / the SourceCounters section have no source page
        ss = sass(rep)
    except Exception:  # noqa: BLE001
        ss = []
    # the source page may list a launch more than once (one section per view): drop consecutive repeats
    uniq = []
    for sec in ss:
        if uniq and all(uniq[-1][k] == sec[k] for k in ("kernel", "warp_instructions", "stall_samples")):
            continue
        uniq.append(sec)
    if len(uniq) == len(ls):
        for launch, sec in zip(ls, uniq):
            launch["sass"] = sec
    else:
        sys.stderr.write(f"source sections ({len(uniq)}) do not match launches ({len(ls)}): listed apart\n")
    doc = {"report": rep, "launches": ls}
    if len(uniq) != len(ls):
        doc["sass_sections"] = uniq
    json.dump(doc, sys.stdout, indent=1)
