"""Summarise an .ncu-rep here (no GPU): key raw metrics + the top stall sites of the source page.
    python tools/ncu_top.py gpurun_out/prof.ncu-rep [n_top]"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
pat = re.compile(r"^(gpu__time_duration.sum|smsp__inst_executed.sum|sm__inst_issued.avg.pct_of_peak_sustained_active|"
                 r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active|"
                 r"sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed|sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active|"
                 r"l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed|l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|"
                 r"lts__throughput.avg.pct_of_peak_sustained_elapsed|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|"
                 r"dram__bytes_read.sum|dram__bytes_write.sum|launch__registers_per_thread|launch__shared_mem_per_block_dynamic|launch__grid_size|launch__block_size|"
                 r"gpc__cycles_elapsed.avg.per_second|l1tex__m_xbar2l1tex_read_bytes.sum|"
                 r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|smsp__warps_eligible.avg.per_cycle_active)$")
for h, u, v in zip(hdr, units, vals):
    if pat.match(h):
        print(f"{h:90s} {v:>16s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
for i, r in enumerate(data):
    r.append(i)
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("total samples", tot, "instructions", sum(int(r[idx["Instructions Executed"]]) for r in data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h[6:]: sum(int(r[idx[h]]) for r in data) for h in stalls}
print("stall totals:", sorted(agg.items(), key=lambda x: -x[1])[:10])
top = sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:ntop]
for r in sorted(top, key=lambda r: r[-1]):
    s = int(r[idx["# Samples"]])
    st = {h[6:]: int(r[idx[h]]) for h in stalls if int(r[idx[h]]) > 0}
    st = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{r[-1]:5d} {r[idx['Source']].strip()[:64]:64s} {s:6d} {100 * s / tot:5.1f}% x{r[idx['Instructions Executed']]:>9s} {st}")
