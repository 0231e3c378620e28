"""Summarise one kernel of an `ncu --set full --import-source on` report: headline metrics, warp-stall mix, opcode mix and
the share of PC samples per code region (runs of instructions with the same executed count = one loop body).
usage: python profiles/summarize_full.py gpurun_out/prof_xxx.ncu-rep > profiles/rNN_<kernel>_ncu_full.txt"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, units, d = r[0], r[1], dict(zip(r[0], r[2]))
u = dict(zip(h, units))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
for k in keys:
    if k in d:
        print("%-70s %s %s" % (k, d[k], u.get(k, "")))
print("\nwarp stall reasons (warps per issued instruction):")
st = sorted(((float(d[k]), k) for k in h if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k),
            reverse=True)
tot = sum(x for x, _ in st) or 1.0
for x, k in st:
    if x > 0.005:
        print("  %-24s %8.3f  %5.1f %%" % (k.split("issue_stalled_")[1].split("_per_issue")[0], x, 100 * x / tot))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 3:
    hh, data = rows[1], rows[2:]
    ia, isrc, isamp = hh.index("Instructions Executed"), hh.index("Source"), hh.index("# Samples")
    ti, ts = sum(int(x[ia]) for x in data), sum(int(x[isamp]) for x in data) or 1
    c, s = collections.Counter(), collections.Counter()
    for x in data:
        t = x[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(x[ia])
        s[op] += int(x[isamp])
    print("\nopcode mix (warp instructions executed %d, PC samples %d):" % (ti, ts))
    for op, n in c.most_common(18):
        print("  %-10s %10d %5.1f %%   samples %5.1f %%" % (op, n, 100 * n / ti, 100 * s[op] / ts))
    print("\ncode regions (consecutive SASS with one executed count; >= 1 % of the samples):")
    print("  first..last   len   executed/instr  samples   first instruction")
    a = 0
    for i in range(1, len(data) + 1):
        if i == len(data) or data[i][ia] != data[a][ia]:
            sm = sum(int(x[isamp]) for x in data[a:i])
            if sm >= 0.01 * ts:
                print("  %5d..%-5d %5d %12s %8.1f %%   %s" % (a, i - 1, i - a, data[a][ia], 100 * sm / ts, data[a][isrc].strip()[:48]))
            a = i
