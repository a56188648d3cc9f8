"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/ncu_summarize.py launches gpurun_out/launches.csv profiles/r01_launch_list.txt
  python tools/ncu_summarize.py full gpurun_out/prof_dense2.ncu-rep profiles/r01_ncu_dense_tc2.txt [json_out]
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__cluster_size",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in r:
        name = re.sub(r"\(.*", "", row[ki]).replace("void ", "")
        v = float(row[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: first {n} launches of "
                f"`python bench.py --steps 1 --warmup 1 --no-cpu-baseline`\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'ms':>10} {'share':>7} {'count':>6} {'avg_us':>9}  kernel\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{t / 1e6:10.3f} {100 * t / tot:6.2f}% {c:6d} {t / c / 1e3:9.1f}  {k[:100]}\n")
        f.write(f"total {tot / 1e6:.3f} ms over {n} launches\n")
    print(open(dst).read())


def full(src, dst, json_out=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = {}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, {src}\n")
        f.write(f"# kernel: {rows[2][hdr.index('Kernel Name')]}\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals = [r[i] for r in rows[2:]]
                f.write(f"{k} [{units[i]}]: {vals}\n")
                out[k] = {"unit": units[i], "values": vals}
    print(open(dst).read())
    if json_out:
        def tobytes(k):
            u = out[k]["unit"].lower()
            scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
            return float(out[k]["values"][-1]) * scale
        j = {"kernel": rows[2][hdr.index("Kernel Name")], "source": src,
             "dram_bytes_per_launch": tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"),
             "tensor_pipe_active_pct_elapsed": float(out["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]["values"][-1]),
             "duration_us": float(out["gpu__time_duration.sum"]["values"][-1]),
             "launch": "fine-net 1024->1024 layer at bench.py --H 256 --W 256: M = 530432 rows (one default chunk of 4144 rays x 128 samples), N=1024, K=1024"}
        json.dump(j, open(json_out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
