"""Summarise ncu captures into small tracked files under profiles/ (the .ncu-rep itself stays in gpurun_out/).

  python tools/ncu_summary.py full   gpurun_out/r1_coarse.ncu-rep profiles/r1_coarse_pair_ncu.json
  python tools/ncu_summary.py launch gpurun_out/r1_launches.csv   profiles/r1_launches.md
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__cluster_size", "launch__block_size", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "dram__bytes.sum.per_second", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        k = {"kernel": d.get("Kernel Name", ("?", ""))[0]}
        for key in KEYS:
            if key in d:
                v, u = d[key]
                try:
                    v = float(v)
                except ValueError:
                    pass
                k[key] = {"value": v, "unit": u}
        kernels.append(k)
    json.dump({"source": rep, "command": "ncu --set full --clock-control none --import-source on", "kernels": kernels},
              open(out, "w"), indent=1)
    print(open(out).read())


def launch(path, out):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.reader(lines[start:]))
    h = rows[0]
    ki, mi, vi, ii, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID"), h.index("Metric Unit")
    agg, order = {}, []
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        key = (r[ii], r[ki])
        if key not in agg:
            agg[key] = {}
            order.append(key)
        agg[key][r[mi]] = (r[vi].replace(",", ""), r[ui])
    per_kernel = {}
    with open(out, "w") as f:
        f.write("| id | kernel | time (us) | dram read (MB) | dram write (MB) |\n|---|---|---|---|---|\n")
        for key in order:
            m = agg[key]

            def val(name, scale):
                if name not in m:
                    return float("nan")
                v, u = m[name]
                v = float(v)
                mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                return v * mult

            t = val("gpu__time_duration.sum", 1)
            rd, wr = val("dram__bytes_read.sum", 1), val("dram__bytes_write.sum", 1)
            name = key[1].split("(")[0]
            f.write(f"| {key[0]} | {name} | {t:.1f} | {rd:.1f} | {wr:.1f} |\n")
            per_kernel.setdefault(name, [0.0, 0])
            per_kernel[name][0] += t
            per_kernel[name][1] += 1
        total = sum(v[0] for v in per_kernel.values())
        f.write("\nShare of device time over the captured launches (cold-cache, serialised under ncu):\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for name, (t, n) in sorted(per_kernel.items(), key=lambda kv: -kv[1][0]):
            f.write(f"| {name} | {n} | {t:.1f} | {100 * t / total:.1f}% |\n")
    print(open(out).read())


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2], sys.argv[3])


def raw_csv(raw, out, note):
    """`ncu -i rep --page raw --csv` output (already exported on the GPU box) -> the same small JSON as full()."""
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        k = {"kernel": d.get("Kernel Name", ("?", ""))[0]}
        for key in KEYS:
            if key in d:
                v, u = d[key]
                try:
                    v = float(v.replace(",", ""))
                except ValueError:
                    pass
                k[key] = {"value": v, "unit": u}
        kernels.append(k)
    json.dump({"source": raw, "note": note, "command": "ncu --set full --clock-control none --import-source on (tools/gpu_r2_ncu.sh)",
               "kernels": kernels}, open(out, "w"), indent=1)


def steps(path, title, f):
    """Launch list (gpu__time_duration.sum per launch) -> the kernels of the LAST step of the run and their share of it."""
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.reader(lines[start:]))
    h = rows[0]
    ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    seq = []
    for r in rows[1:]:
        if len(r) > vi and r[mi] == "gpu__time_duration.sum":
            mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
            seq.append((r[ki].split("(")[0].replace("void ", ""), float(r[vi].replace(",", "")) * mult))
    # a step starts at the query preparation kernel; take the last complete one before the verification launches
    starts = [i for i, (n, _) in enumerate(seq) if "quantise_queries" in n or "prepare_queries" in n or "prepare_bits_queries" in n or "query_range" in n]
    firsts = [i for j, i in enumerate(starts) if j == 0 or starts[j - 1] != i - 1]
    if len(firsts) < 2:
        return
    a, b = firsts[-2], firsts[-1]
    step = seq[a:b]
    total = sum(t for _, t in step)
    f.write(f"\n### {title}\n\n| kernel | time (us) | share of the step's device time |\n|---|---|---|\n")
    for n, t in step:
        f.write(f"| {n} | {t:.1f} | {100 * t / total:.1f}% |\n")
    f.write(f"| **sum** | **{total:.1f}** | |\n")
