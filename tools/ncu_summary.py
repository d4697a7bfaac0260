"""Condense an .ncu-rep (captured with --set full --import-source on) into a committed text summary:
per kernel the headline metrics of the raw page and the hottest SASS instructions of the source page.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_ncu_summary.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main(rep):
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    names = []
    print(f"# ncu summary of {rep}\n")
    for r in raw[2:]:
        name = r[hdr.index("Kernel Name")]
        names.append(name)
        print(f"## {name[:110]}")
        print(f"   grid {r[hdr.index('Grid Size')]}  block {r[hdr.index('Block Size')]}")
        for k in KEYS:
            if k in hdr and r[hdr.index(k)] not in ("", "n/a"):
                print(f"   {k:86s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        print()
    for name in names:
        short = name.split("(")[0].split("::")[-1].split("<")[0]
        src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--kernel-name", "regex:" + short]))))
        if len(src) < 3:
            continue
        h = src[1]
        iS, iSrc, iEx = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
        stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        data = [r for r in src[2:] if len(r) > max(stall) and r[iS].isdigit() and r[0].startswith("0x")]
        data = data[: len(data) // 2] if len(data) > 1 and data[0][0] == data[len(data) // 2][0] else data
        tot = sum(int(r[iS]) for r in data) or 1
        print(f"## hottest SASS of {short} (warp-state samples {tot})")
        ops = Counter()
        for r in data:
            t = r[iSrc].split()
            if t:
                ops[t[1] if t[0].startswith("@") and len(t) > 1 else t[0]] += int(r[iS])
        print("   by opcode: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(10)))
        for r in sorted(data, key=lambda r: -int(r[iS]))[:12]:
            st = sorted(((h[i][6:], int(r[i])) for i in stall if r[i].isdigit() and int(r[i]) > 0), key=lambda kv: -kv[1])[:2]
            print(f"   {100 * int(r[iS]) / tot:5.1f}%  x{r[iEx]:>9s}  {r[iSrc].strip()[:72]:72s} {st}")
        mma = sum(1 for r in data if "UTCHMMA" in r[iSrc] or "UTCQMMA" in r[iSrc])
        tma = sum(1 for r in data if "UBLKCP" in r[iSrc] or "UTMALDG" in r[iSrc] or "UBLKRED" in r[iSrc])
        tm = sum(1 for r in data if "LDTM" in r[iSrc] or "STTM" in r[iSrc])
        print(f"   SASS mnemonics: UTCHMMA/UTCQMMA x{mma}, TMA (UTMALDG/UBLKCP/UBLKRED) x{tma}, LDTM/STTM x{tm}\n")


if __name__ == "__main__":
    main(sys.argv[1])
