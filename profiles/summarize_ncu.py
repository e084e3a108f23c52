"""Turn gpurun_out/*.ncu-rep and launch-list CSVs into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py full  gpurun_out/r2_prof_grad.ncu-rep  > profiles/r1_ppo_grad_full.txt
    python profiles/summarize_ncu.py list  gpurun_out/r1_launches.csv      > profiles/r1_launch_list.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|sm__throughput.avg.pct|"
    r"sm__warps_active.avg.pct|launch__(registers_per_thread|grid_size|block_size|occupancy_limit|shared_mem_per_block_dynamic)|"
    r"sm__inst_executed_pipe_(fma|lsu|xu|alu|fp64|tensor).*pct_of_peak_sustained_active$|sm__pipe_(fma|tensor|fp64|alu).*cycles_active.avg.pct_of_peak_sustained_active$|"
    r"smsp__issue_active.avg.pct|smsp__inst_executed.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|"
    r"l1tex__data_pipe_(lsu|tc)_wavefronts_mem_shared(_op_ld|_op_st)?.sum(.pct_of_peak_sustained_elapsed)?$|sm__cycles_elapsed.avg$|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_(ld|st).sum$|smsp__average_warps_issue_stalled_.*_per_issue_active.ratio$|"
    r"lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|smsp__cycles_active.avg$|sm__cycles_active.avg$")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print(f"# kernel: {name}")
        for h, u, v in zip(hdr, units, vals):
            if KEEP.search(h):
                print(f"{h:82s} {v:>18s} {u}")
        print()


def launch_list(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(row["Metric Unit"], 1e-3)
        agg[re.sub(r"\(.*", "", row["Kernel Name"])].append(v * scale)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':72s} {'launches':>8s} {'mean_us':>10s} {'total_ms':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:72]:72s} {len(v):8d} {sum(v) / len(v):10.1f} {sum(v) / 1e3:10.3f} {sum(v) / tot:7.3f}")


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2])
