"""Extract the metrics DESIGN.md / bench.py cite from an `ncu --set full` report.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/<name>.md
Also writes <name>.json next to stdout target when --json PATH is given."""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC per SM"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy pipe (IMAD) cycles active % of elapsed"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe (heavy+lite avg) cycles active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe cycles active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("sm__sass_inst_executed_op_shared_ld.sum", "shared loads"), ("sm__sass_inst_executed_op_shared_st.sum", "shared stores"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
]


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    jpath = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary: {rep}\n")
    allj = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "")
        if sub and sub not in name:
            continue
        print(f"## `{name.split('(')[0]}`  (launch id {d.get('ID')})\n")
        print("| metric | value | unit |\n|---|---:|---|")
        j = {"kernel": name.split("(")[0]}
        for k, label in KEYS:
            if k in d and d[k] != "":
                print(f"| {label} (`{k}`) | {d[k]} | {u.get(k, '')} |")
                j[k] = d[k]; j[k + ".unit"] = u.get(k, "")
        print()
        allj.append(j)
    if jpath:
        json.dump(allj, open(jpath, "w"), indent=1)


if __name__ == "__main__":
    main()
