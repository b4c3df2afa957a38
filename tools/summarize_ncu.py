"""Turn ncu artefacts from gpurun_out/ into small tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_x.csv profiles/r1_launches_x.md
  python tools/summarize_ncu.py rep gpurun_out/prof_x.ncu-rep profiles/r1_prof_x.txt
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] == "ns":
            v /= 1e3
        elif row["Metric Unit"] == "ms":
            v *= 1e3
        key = re.sub(r"\(.*", "", row["Kernel Name"])[:70] + " grid=" + row.get("Grid Size", "")
        agg[key][0] += 1
        agg[key][1] += v
        tot += v
    with open(dst, "w") as out:
        out.write(f"# ncu launch list summary of `{src}` (gpu__time_duration.sum, --clock-control none; cold-cache, "
                  "serialised: compare shares)\n\n| kernel | launches | total us | us/launch | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            out.write(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |\n")
        out.write(f"\ntotal {tot:.1f} us\n")


def rep(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as out:
        out.write(f"ncu --set full --clock-control none capture: {src}\n")
        for r in rows[2:]:
            out.write("\n" + r[hdr.index("Kernel Name")] + "\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    out.write(f"  {m:72s} {r[i]:>16s} {units[i]}\n")


def traffic(src, dst):
    """dram read + write bytes of the (first) captured launch -> the JSON bench.py reads for `roofline.traffic`."""
    import json
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]

    def val(name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    i = hdr.index("gpu__time_duration.sum")
    json.dump({"kernel": r[hdr.index("Kernel Name")], "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_launch": rd + wr, "gpu_time": f"{r[i]} {units[i]}",
               "source": f"profiles/{dst.split('/')[-1]} <- ncu --set full --clock-control none capture {src.split('/')[-1]}"},
              open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "rep": rep, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
