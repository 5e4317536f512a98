"""Summarise one kernel of an .ncu-rep (ncu --set full) -- or of its `--page raw --csv` export -- into the text
form kept under profiles/.
    python tools/ncu_summary.py <report.ncu-rep | raw.csv> "<header line>" [launch index] > profiles/<name>.txt"""
import csv, subprocess, sys
if sys.argv[1].endswith(".csv"):
    out = open(sys.argv[1]).read()
else:
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
hdr, units, vals = rows[0], rows[1], rows[2 + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)]
d = dict(zip(hdr, zip(vals, units)))
print("# " + sys.argv[2])
print("# kernel: " + d["Kernel Name"][0])
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'sm__icc_request_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for k in KEYS:
    if k in d and d[k][0] not in ("", "n/a"):
        print("%s [%s] = %s" % (k, d[k][1], d[k][0]))
st = []
for k, v in d.items():
    if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued'):
        try:
            st.append((k.replace('smsp__pcsamp_warps_issue_stalled_', ''), float(v[0].replace(',', ''))))
        except ValueError:
            pass
tot = sum(v for _, v in st) or 1
print("# stall samples: " + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in sorted(st, key=lambda x: -x[1])[:10]))
