"""Key metrics of every kernel in an .ncu-rep (raw page): duration, DRAM bytes, throughput %, issue, occupancy, registers."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = {"Kernel Name": "kernel", "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
        "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct"}
idx = {h: i for i, h in enumerate(hdr)}
print("| kernel | duration | DRAM read | DRAM write | DRAM GB/s | DRAM % | SM % | issue % | fp64 pipe % | occ % | regs | L1 hit % | L2 hit % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
def val(r, k):
    i = idx.get(k)
    return (r[i], units[i]) if i is not None else ("", "")
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def to_sec(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1)
for r in rows[2:]:
    name = val(r, "Kernel Name")[0].replace("smr::", "")
    name = name[:90]
    d = to_sec(*val(r, "gpu__time_duration.sum"))
    br, bw = to_bytes(*val(r, "dram__bytes_read.sum")), to_bytes(*val(r, "dram__bytes_write.sum"))
    f = lambda k: val(r, k)[0]
    print(f"| `{name}` | {d*1e6:.1f} us | {br/1e6:.1f} MB | {bw/1e6:.1f} MB | {(br+bw)/d/1e9:.0f} | {f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | "
          f"{f('sm__throughput.avg.pct_of_peak_sustained_elapsed')} | {f('smsp__issue_active.avg.pct_of_peak_sustained_active')} | "
          f"{f('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')} | {f('sm__warps_active.avg.pct_of_peak_sustained_active')} | "
          f"{f('launch__registers_per_thread')} | {f('l1tex__t_sector_hit_rate.pct')} | {f('lts__t_sector_hit_rate.pct')} |")
