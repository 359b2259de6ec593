#!/usr/bin/env python3
"""tools/ncu_summary.py <report.ncu-rep> [--traffic OUT.json LANES COMMIT DATE] -- prints the per-kernel metrics we track
(CPU side, no GPU).  sm__pipe_fmaheavy_cycles_active is the pipe that executes IMAD.WIDE (the multiplier's instruction);
--traffic also writes the DRAM bytes per launch of the k_verify_* kernels (what bench.py's roofline.traffic scales per lane)."""
import csv, json, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum','launch__registers_per_thread','launch__block_size','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_selected_per_issue_active.ratio']
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:60])
    for k in keys:
        if k in hdr:
            i = hdr.index(k); print('  %-88s %s %s' % (k, r[i], units[i]))

if len(sys.argv) > 2 and sys.argv[2] == "--traffic":
    out, lanes, commit, date = sys.argv[3], int(sys.argv[4]), sys.argv[5], sys.argv[6]
    tr = {}
    ir, iw, ik = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('Kernel Name')
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = r[ik].split("(")[0]
        if name.startswith("k_verify_") and name not in tr:
            tr[name] = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    json.dump({"source": "ncu --set full, one full wave, " + rep, "lanes": lanes, "commit": commit, "date": date,
               "dram_bytes_per_launch": tr}, open(out, "w"), indent=1)
