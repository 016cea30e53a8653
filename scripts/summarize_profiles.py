#!/usr/bin/env python3
"""Turn the ncu captures under gpurun_out/final/ into the tracked summaries in profiles/.
usage: python scripts/summarize_profiles.py <tag>   (e.g. r1)"""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "final"); OUT = os.path.join(ROOT, "profiles")
KEYS = ['lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_srcunit_tex_lookup_hit.sum',
        'lts__t_sectors_srcunit_tex_lookup_miss.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_bytes.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum',
        'Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v.replace(",", "")) * m.get(unit, 1)
traffic = {}
for k in ("select", "expand", "expand_select", "resnet", "warmup", "trunk", "head", "play"):
    rep = os.path.join(SRC, f"prof_{k}.ncu-rep")
    if not os.path.exists(rep): continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, f"{tag}_ncu_{k}_summary.csv"), "w") as f:
        f.write(",".join(["metric"] + [f"launch{i}" for i in range(len(rows) - 2)]) + "\n")
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                f.write(",".join([f"{key} [{units[i]}]"] + [r[i].replace(",", ";") for r in rows[2:]]) + "\n")
    ir, iw, it = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('gpu__time_duration.sum')
    n = len(rows) - 2
    traffic[k] = {"dram_bytes_per_launch": sum(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in rows[2:]) / n,
                  "duration_us_under_ncu": sum(float(r[it]) for r in rows[2:]) / n, "launches_captured": n}
    lines = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    top = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), "25"], input=lines, capture_output=True, text=True).stdout
    open(os.path.join(OUT, f"{tag}_ncu_{k}_hot_lines.txt"), "w").write(top)
json.dump({"source": f"ncu --set full --clock-control none, profiles/{tag}_ncu_select_summary.csv", **traffic.get("select", {})},
          open(os.path.join(OUT, "select_traffic.json"), "w"), indent=1)
json.dump(traffic, open(os.path.join(OUT, f"{tag}_kernel_traffic.json"), "w"), indent=1)
# launch shares
rows = list(csv.reader(open(os.path.join(SRC, "launches.csv"))))
for i, r in enumerate(rows):
    if 'Kernel Name' in r: hdr = r; start = i + 1; break
kn, mv, mn = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start:]:
    if len(r) <= mv or r[mn] != 'gpu__time_duration.sum': continue
    name = r[kn].split('(')[0][:90]; agg[name][0] += 1; agg[name][1] += float(r[mv].replace(',', ''))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(OUT, f"{tag}_launch_shares.csv"), "w") as f:
    f.write("kernel,launches,total_ns,share_pct,avg_us\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f'"{k}",{v[0]},{v[1]:.0f},{100*v[1]/tot:.2f},{v[1]/v[0]/1000:.2f}\n')
print(open(os.path.join(OUT, f"{tag}_launch_shares.csv")).read())
print(json.dumps(traffic, indent=1))
