"""Turn an .ncu-rep (ncu --set full) into the per-kernel summary table kept under profiles/."""
import csv
import json
import subprocess
import sys


def num(v, u):
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1,
            "Kbyte/block": 1e3, "byte/block": 1, "second": 1, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}.get(u, 1)
    return float(v) * mult


def rows_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = {"gpu__time_duration.sum": "time", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
            "smsp__inst_executed.sum": "inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue",
            "sm__warps_active.avg.pct_of_peak_sustained_active": "warps",
            "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
            "smsp__thread_inst_executed_per_inst_executed.ratio": "thr",
            "launch__shared_mem_per_block_dynamic": "smem",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "bank"}
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in want:
                d[want[h]] = num(r[i], units[i]) if r[i] not in ("", "n/a") else 0.0
        out.append(d)
    return out


def main():
    print("| kernel | grid x block | regs | dyn smem KB | time us | DRAM read MB | DRAM write MB | DRAM GB/s | "
          "warp inst M | issue % | warps active % | threads/inst |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    traffic = {}
    for rep in sys.argv[1:]:
        for d in rows_of(rep):
            name = d["kernel"].split("(")[0].replace("void ", "").replace("smlvm::", "")
            t = d["time"]
            print("| %s | %dx%d | %d | %.1f | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %.1f |" % (
                name[:48], d["grid"], d["block"], d["regs"], d.get("smem", 0) / 1e3, t * 1e6, d["rd"] / 1e6,
                d["wr"] / 1e6, (d["rd"] + d["wr"]) / t / 1e9, d["inst"] / 1e6, d["issue"], d["warps"], d["thr"]))
            traffic[name] = d["rd"] + d["wr"]
    json.dump(traffic, sys.stderr)


if __name__ == "__main__":
    main()
