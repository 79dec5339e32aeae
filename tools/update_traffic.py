#!/usr/bin/env python3
"""Fold `ncu --page raw --csv` captures into profiles/traffic.json (read by bench.py's roofline block).

    python tools/update_traffic.py <workload>/<kernel key> <raw csv> [note]   (repeatable: triples via --)
    python tools/update_traffic.py --round r02b     # the standard set of tools/final_sweep.sh captures

An entry keeps what the capture measured for ONE launch: DRAM bytes, duration under ncu, issue-slot
utilisation, lanes per instruction, the busiest pipe, cache hit rates, and `bound` = the unit the
counters show nearest its ceiling ("hbm", "tex", "lsu", "issue").  Nothing here is a bench value.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "profiles", "traffic.json")
HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def read_capture(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            d[h] = float(v.replace(",", "")) * UNIT.get(u, 1)
        except ValueError:
            d[h] = v
    return d


def entry(path, note=None):
    d = read_capture(os.path.join(ROOT, path))
    ms = d["gpu__time_duration.sum"]
    rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
    issue = d["smsp__issue_active.avg.pct_of_peak_sustained_active"]
    pipes = {k: d.get(f"sm__inst_executed_pipe_{k}.avg.pct_of_peak_sustained_active", 0.0)
             for k in ("alu", "fma", "lsu", "xu")}
    tex = d.get("l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", 0.0)
    lsu = d.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 0.0)
    dram_frac = (rd + wr) / (ms / 1e3) / 1e9 / HBM_PEAK
    ceilings = {"hbm": 100 * dram_frac, "tex": tex, "lsu": lsu, "issue": max(issue, pipes["alu"], pipes["fma"])}
    e = {
        "kernel": d.get("Kernel Name"),
        "dram_bytes_read": round(rd), "dram_bytes_write": round(wr),
        "kernel_ms_under_ncu": round(ms, 4), "capture_ms": round(ms, 4), "capture": path,
        "bound": max(ceilings, key=ceilings.get),
        "dram_frac_of_measured_peak": round(dram_frac, 4),
        "issue_active": round(issue, 1),
        "lanes_per_instruction": d["smsp__thread_inst_executed_per_inst_executed.ratio"],
        "pipe_alu": round(pipes["alu"], 1), "pipe_fma": round(pipes["fma"], 1),
        "tex_wavefronts": round(tex, 1), "lsu_wavefronts": round(lsu, 1),
        "l1_hit": round(d.get("l1tex__t_sector_hit_rate.pct", 0.0), 1),
        "l2_hit": round(d.get("lts__t_sector_hit_rate.pct", 0.0), 1),
        # achieved L2 rate: 32-byte sectors the L2 slices served per second of the capture
        "l2_gb_per_s": round(d.get("lts__t_sectors.sum", 0.0) * 32 / (ms / 1e3) / 1e9, 1),
        "l2_throughput_pct": round(d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", 0.0), 1),
        "warps_active": round(d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0), 1),
        "registers": int(d.get("launch__registers_per_thread", 0)),
        "warp_instructions": round(d.get("smsp__inst_executed.sum", 0)),
    }
    if note:
        e["note"] = note
    return e


# name in tools/final_sweep.sh -> key bench.py looks up (workload / kernel_name_for())
STANDARD = {
    "dda_cfg4_f120": ("cfg4/dda_skip_tex_kernel", "script frame 120 (inside the volume)"),
    "dda_cfg3_f120": ("cfg3/dda_skip_tex_kernel", "script frame 120 (inside the volume)"),
    "dda_cfg1": ("cfg1/dda_skip_tex_kernel", None),
    "dda_cfg2": ("cfg2/dda_skip_tex_kernel", None),
    "esvo_f120": ("cfg2/esvo_kernel", "script frame 120 (inside the volume)"),
    "esvo_cfg4e_f120": ("cfg4e/esvo_kernel", "script frame 120 (inside the volume)"),
    # (a later entry for the same key wins: the interior frame where both were captured)
    "dfr_f15": ("cfg2/svo_df_kernel", "script frame 15 (outside, distance 2)"),
    "df_f120": ("cfg2/svo_df_kernel", "script frame 120"),
    "rope_f15": ("cfg2/svo_rope_kernel", "script frame 15 (outside, distance 2)"),
    "rope_f120": ("cfg2/svo_rope_kernel", "script frame 120"),
    "rope_cfg3r_f120": ("cfg3r/svo_rope_kernel", "script frame 120"),
    "naive_f15": ("cfg2/svo_naive_kernel", "script frame 15"),
    "naive_f120": ("cfg2/svo_naive_kernel", "script frame 120"),
}


def main(argv):
    t = json.load(open(PATH))
    if argv and argv[0] == "--round":
        tag = argv[1]
        for name, (key, note) in STANDARD.items():
            path = f"profiles/{tag}_{name}_ncu_raw.csv"
            if os.path.exists(os.path.join(ROOT, path)):
                t[key] = entry(path, note)
                print(key, "<-", path, t[key]["bound"], t[key]["dram_frac_of_measured_peak"])
    else:
        while argv:
            key, path = argv[0], argv[1]
            note = argv[2] if len(argv) > 2 and argv[2] != "--" else None
            t[key] = entry(path, note)
            print(key, "<-", path, t[key]["bound"])
            argv = argv[(3 if note else 2):]
            if argv and argv[0] == "--":
                argv = argv[1:]
    json.dump(t, open(PATH, "w"), indent=2)
    open(PATH, "a").write("\n")


if __name__ == "__main__":
    main(sys.argv[1:])
