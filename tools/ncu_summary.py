#!/usr/bin/env python
"""Turn ncu output into the small, tracked summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep  > profiles/rNN_full.md
    python tools/ncu_summary.py train    gpurun_out/train_launches.csv > profiles/rNN_train_launches.md

`launches` reads the CSV log of `ncu --metrics gpu__time_duration.sum --clock-control none --csv` (one row per launch)
and prints every launch of ONE forward pass plus the per-kernel share of the step.  `full` reads a `--set full`
report through `ncu -i ... --page raw --csv` (ncu runs here without a GPU) and prints the roofline-relevant counters
of every captured launch: duration, DRAM bytes read + written (the `traffic` of bench.py's roofline object), DRAM and
tensor-pipe utilisation, occupancy, registers.
"""
import csv
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = name.replace("air::", "").replace("tc::", "")
    return name.split("(")[0].replace("void ", "")


def launches(path, per_step=None):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    seq = [(short(r[ki]), float(r[vi].replace(",", "")) / 1e3, r[gi], r[bi]) for r in rows[1:]]
    # one forward pass = from one elbo_scalars_kernel (exclusive) to the next (inclusive)
    ends = [i for i, s in enumerate(seq) if s[0].startswith("elbo_scalars")]
    if len(ends) >= 2:
        step = seq[ends[-2] + 1: ends[-1] + 1]
    else:
        step = seq
    total = sum(s[1] for s in step)
    print(f"# ncu launch list: one forward pass ({len(step)} launches, {total:.1f} us serialised, cold-cache, "
          f"--clock-control none)\n")
    print("| # | kernel | grid | block | us | share |\n|---|---|---|---|---|---|")
    for i, (k, us, g, b) in enumerate(step):
        print(f"| {i} | {k} | {g} | {b} | {us:.2f} | {100 * us / total:.1f}% |")
    agg = OrderedDict()
    for k, us, _, _ in step:
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + us)
    print("\n| kernel | launches | us | share of step |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {t:.2f} | {100 * t / total:.1f}% |")


def train(path):
    """Launch list of ONE training step (tools/train_step_probe.py): from one rmsprop_centered_kernel (exclusive) to the next
    (inclusive), split at the forward's elbo_scalars_kernel."""
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    seq = [(short(r[ki]), float(r[vi].replace(",", "")) / 1e3, r[gi], r[bi]) for r in rows[1:]]
    ends = [i for i, s in enumerate(seq) if s[0].startswith("rmsprop_centered")]
    step = seq[ends[-2] + 1: ends[-1] + 1] if len(ends) >= 2 else seq
    cut = max(i for i, s in enumerate(step) if s[0].startswith("elbo_scalars")) + 1
    total = sum(s[1] for s in step)
    fwd, bwd = sum(s[1] for s in step[:cut]), sum(s[1] for s in step[cut:])
    print(f"# ncu launch list: one TRAINING step at B=4096 (tools/train_step_probe.py; AIR_PREC_TC_SPLIT handle in training "
          f"mode)\n\n{len(step)} launches, {total:.1f} us serialised (cold-cache, --clock-control none; ncu serialises the "
          f"streams, the real step overlaps the weight-gradient work with the critical path): forward {fwd:.1f} us, backward + "
          f"optimiser {bwd:.1f} us.\n")
    for title, part in (("forward (activations kept)", step[:cut]), ("backward + centered RMSProp", step[cut:])):
        agg = OrderedDict()
        for k, us, _, _ in part:
            n, t = agg.get(k, (0, 0.0))
            agg[k] = (n + 1, t + us)
        print(f"## {title}\n\n| kernel | launches | us | share of step |\n|---|---|---|---|")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"| {k} | {n} | {t:.1f} | {100 * t / total:.1f}% |")
        print()
    print("## every launch\n\n| # | kernel | grid | block | us |\n|---|---|---|---|---|")
    for i, (k, us, g, b) in enumerate(step):
        print(f"| {i} | {k} | {g} | {b} | {us:.2f} |")


WANT = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % (active)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    cols = [(hdr.index(m), m, label) for m, label in WANT if m in hdr]
    print(f"# ncu --set full summary of {path.split('/')[-1]} (per launch; --clock-control none)\n")
    print("| kernel | grid | " + " | ".join(f"{label} [{units[i]}]" if units[i] else label for i, _, label in cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for r in rows[2:]:
        print(f"| {short(r[ki])} | {r[gi]} | " + " | ".join(r[i] for i, _, _ in cols) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full, "train": train}[sys.argv[1]](sys.argv[2])
