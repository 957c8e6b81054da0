#!/usr/bin/env python3
"""Plot script -- successor of the reference's per-variant `plot.py`
(/root/reference/parallel_pivot/plot.py:7-91): reads benchmark_results_1M.json written by
sweep.py (same schema) and draws the same 2 x 2 figure (average runtime, variance, standard
deviation, fourth panel) against matrix size.  The reference's fourth panel plots register
counts scraped from its nohup compile log (plot.py:20-27); there is no per-configuration
compile here, so it shows the fraction of the HBM roofline (or cuBLAS speed-up when present).
matplotlib is imported lazily: `summarise()` works without it and prints the table.
"""
from __future__ import annotations

import argparse
import json

import numpy as np


def load(path):
    with open(path) as f:
        data = json.load(f)
    keys = sorted(data, key=lambda k: data[k]["matrix_size"])  # plot.py:30
    col = lambda name, default=None: [data[k].get(name, default) for k in keys]
    return {
        "matrix_sizes": col("matrix_size"),
        "runtime_avgs": col("runtime_avg"),
        "variances": col("variance"),
        "std_devs": col("std_dev"),
        "incorrect_inversions": [float(np.mean(v)) if v else 0.0 for v in col("incorrect_inversions", [])],
        "frac_hbm_roofline": col("frac_hbm_roofline"),
        "speedup_vs_cublas": col("speedup_vs_cublas"),
        "gbps": col("gbps"),
    }


def summarise(d):
    lines = ["N   avg_ms    std_ms   GB/s   roofline  xcuBLAS  incorrect"]
    for i, n in enumerate(d["matrix_sizes"]):
        f = lambda v, fmt: (fmt % v) if v is not None else "-"
        lines.append("%-3d %-9.4f %-8.4f %-6s %-9s %-8s %g" % (
            n, d["runtime_avgs"][i], d["std_devs"][i], f(d["gbps"][i], "%.0f"), f(d["frac_hbm_roofline"][i], "%.3f"),
            f(d["speedup_vs_cublas"][i], "%.1f"), d["incorrect_inversions"][i]))
    return "\n".join(lines)


def plot(d, out_png, title="Performance Analysis Metrics"):
    import matplotlib
    matplotlib.use("Agg")
    import matplotlib.pyplot as plt

    fig, ((ax1, ax2), (ax3, ax4)) = plt.subplots(2, 2, figsize=(15, 12), dpi=200)
    fig.suptitle(title, fontsize=16, y=0.95)
    x = d["matrix_sizes"]
    for ax, y, c, t, yl in ((ax1, d["runtime_avgs"], "b", "Average Runtime vs Matrix Size", "Average Runtime (milli seconds)"),
                            (ax2, d["variances"], "r", "Runtime Variance vs Matrix Size", "Variance"),
                            (ax3, d["std_devs"], "g", "Standard Deviation vs Matrix Size", "Standard Deviation")):
        ax.plot(x, y, c + "-o", linewidth=2, markersize=6)
        ax.set_title(t); ax.set_xlabel("Matrix Size"); ax.set_ylabel(yl); ax.grid(True, linestyle="--", alpha=0.7)
    y4 = d["speedup_vs_cublas"] if any(v is not None for v in d["speedup_vs_cublas"]) else d["frac_hbm_roofline"]
    ax4.plot(x, [v if v is not None else np.nan for v in y4], "m-o", linewidth=2, markersize=6)
    ax4.set_title("Speed-up over cuBLAS getrf+getri" if y4 is d["speedup_vs_cublas"] else "Fraction of HBM roofline")
    ax4.set_xlabel("Matrix Size"); ax4.grid(True, linestyle="--", alpha=0.7)
    fig.savefig(out_png)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("json", nargs="?", default="benchmark_results_1M.json")
    ap.add_argument("--png", default="")
    a = ap.parse_args(argv)
    d = load(a.json)
    print(summarise(d))
    if a.png:
        plot(d, a.png)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
