"""Extract the judged metrics of every kernel in an .ncu-rep (ncu --set full) into small CSV summaries.
usage: python tools/ncu_full_summary.py file.ncu-rep out_prefix "header note"      (needs `ncu` on PATH; runs `ncu -i ... --page raw --csv`)"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(Kernel Name|dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum|l1tex__throughput\.avg\.pct|lts__throughput\.avg\.pct|"
    r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit)|"
    r"sm__cycles_active\.avg$|sm__inst_executed_pipe_tensor|sm__throughput\.avg\.pct|smsp__average_warps_issue_stalled_.*_per_issue_active|"
    r"smsp__issue_active\.avg\.pct|smsp__warps_active\.avg\.per_cycle_active|sm__warps_active\.avg\.pct_of_peak|smsp__inst_executed\.sum$|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|sm__pipe_fma_cycles_active\.avg\.pct)")


def main(rep, prefix, note):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    seen = {}
    for r in data:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]).replace("<unnamed>::", "").replace("void ", "")
        name = re.sub(r"[^A-Za-z0-9_]+", "_", name).strip("_")
        k = seen.get(name, 0)
        seen[name] = k + 1
        if k:  # first launch of each kernel only
            continue
        path = f"{prefix}_{name}_ncu_full_summary.csv"
        with open(path, "w") as f:
            f.write(f"# ncu --set full --clock-control none, first captured launch of {name}; {note}\nmetric,unit,value\n")
            for h, u, v in sorted(zip(hdr, units, r)):
                if KEEP.search(h):
                    f.write(f"{h},{u},{v}\n")
        print("wrote", path)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
