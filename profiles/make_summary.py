#!/usr/bin/env python
"""Regenerates profiles/rNN_ncu_summary.md from the ncu artefacts a gpurun call left in gpurun_out/.

    python profiles/make_summary.py r01 gpurun_out/launches_r1.csv gpurun_out/prof_rollout_r1.ncu-rep gpurun_out/prof_step_r1.ncu-rep

Inputs: the launch list (ncu --metrics gpu__time_duration.sum --csv) of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`
and the --set full reports of the rollout kernel and of the stand-alone step kernels (read with `ncu -i ... --page raw --csv`).
"""
import collections
import csv
import shutil
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__cycles_elapsed.avg.per_second"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    return [(dict(zip(h, r)), dict(zip(h, units))) for r in rows[2:]]


def section(title, rec, units, out):
    out.append(f"\n## ncu --set full: {title}\n")
    for k in METRICS:
        if k in rec:
            out.append(f"{k:<82} {rec[k]:>14} {units.get(k, '')}")
    stalls = {k: float(v) for k, v in rec.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
              or k.startswith("smsp__average_warp_latency_issue_stalled_")}
    if stalls:
        out.append("warp stall reasons (cycles per issue):")
        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]:
            out.append(f"   {k.split('stalled_')[1].split('.')[0].replace('_per_issue_active',''):<40} {v:.3f}")


def main():
    tag, launches, rollout_rep, step_rep = sys.argv[1:5]
    out = [f"# Round {tag} ncu evidence (B200, driver 580, CUDA 12.9, ncu --clock-control none)\n",
           "Commands (each after the same command exited 0 without ncu in the same gpurun call):",
           "  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra",
           "  ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o ... python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --no-aux",
           "  ncu --set full --clock-control none --import-source on -k regex:_step_ -c 2 -s 6 -o ... python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra\n"]
    # ---- launch list
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    iu = h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or not r[iv]:
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "s": 1e3}.get(r[iu], 1e-6)
        a = agg.setdefault(r[ik], [0.0, 0, []])
        a[0] += v
        a[1] += 1
        a[2].append(v)
    total = sum(a[0] for a in agg.values())
    out.append(f"## Launch list of `bench.py --steps 2 --warmup 3 --no-cpu-baseline` (full file: {tag}_launches_bench.csv)\n")
    out.append("   total_ms  count  share  kernel")
    for k, (ms, c, _) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:10]:
        out.append(f"{ms:11.3f} {c:6d} {100 * ms / total:6.2f}%  {k[:110]}")
    top = max(agg.items(), key=lambda kv: kv[1][0])
    full = [v for v in top[1][2] if v >= 0.6 * max(top[1][2])]      # launches over the whole env range (warm-up + timed steps)
    part = [v for v in top[1][2] if v < 0.6 * max(top[1][2])]       # env slices of the pipelined host-buffer (e2e) calls
    out.append(f"sum of all kernel time: {total:.1f} ms.  {top[0].split('(')[0][:60]}: {len(full)} launches over all envs (warm-up + timed: "
               f"{sum(full) / len(full):.2f} ms each; one timed bench step is exactly one launch of it)"
               + (f" + {len(part)} launches over env slices inside the e2e calls (csrc/host_pipe.cuh: {sum(part) / len(part):.2f} ms each, "
                  f"{sum(part):.1f} ms in total)." if part else "."))
    shutil.copy(launches, f"profiles/{tag}_launches_bench.csv")
    # ---- full reports
    for rec, units in raw(rollout_rep)[:1]:
        section(rec.get("Kernel Name", "rollout_kernel")[:120] + " (the bench launch: 2^20 envs x 200 steps)", rec, units, out)
        rd = float(rec["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units["dram__bytes_read.sum"]]
        wr = float(rec["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units["dram__bytes_write.sum"]]
        out.append(f"dram traffic of this launch: {rd + wr:.4e} B (read {rd:.3e} + write {wr:.3e})")
    for rec, units in raw(step_rep)[:2]:
        section(rec.get("Kernel Name", "step kernel")[:120] + " (2^25 envs, one step)", rec, units, out)
        rd = float(rec["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units["dram__bytes_read.sum"]]
        wr = float(rec["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units["dram__bytes_write.sum"]]
        out.append(f"dram traffic of this launch: {rd + wr:.4e} B (read {rd:.3e} + write {wr:.3e})")
    open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
