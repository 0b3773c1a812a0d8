"""Runs bench.py's kernel sweep alone (the `kernels` block of the bench line) and prints / stores the rows.

  python tools/kernel_sweep.py [--sizes-mb 64 1024] [--filter range_learning] [--out gpurun_out/kernels.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes-mb", type=int, nargs="*", default=[64, 1024])
    ap.add_argument("--filter", default="")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import bench
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    peak = bench.peak_hbm()[0]
    rows = bench.kernel_sweep(device, peak, sizes_mb=tuple(args.sizes_mb))
    rows = [r for r in rows if args.filter in r["kernel"]]
    for r in rows:
        print(r.get("kernel"), r.get("dtype"), r.get("mb"), r.get("us"), r.get("gbs"), r.get("frac"))
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"peak": peak, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
