"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/summarize_launches.py gpurun_out/launches.csv [--skip-first-half] > profiles/launches_rN_summary.txt

ncu serialises launches and runs them cold-cache, so absolute times are not bench values: the SHARE of the step per
kernel is what is compared with the CUDA-event shares `bench.py` reports.  `--skip-first-half` drops the first half of
the launches (the warm-up step of `bench.py --ncu`, which also contains cuDNN's algorithm trials)."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    body = [r for r in rows[hi + 1:] if len(r) > vi]
    if "--skip-first-half" in sys.argv:
        body = body[len(body) // 2:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in body:
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name)[:110]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("sdb::"))
    if "--categories" in sys.argv:
        # whose kernels: this library / cuBLAS(Lt) GEMMs / cuDNN convolutions / ATen (elementwise, reductions, copies)
        def category(k):
            if "sdb::" in k:
                return "this library"
            if "implicit_gemm" in k or "cudnn" in k or "xmma_fprop" in k or "xmma_dgrad" in k or "xmma_wgrad" in k or "conv" in k:
                return "cuDNN convolutions"
            if k.startswith("cutlass") or k.startswith("nvjet") or "gemm" in k or "cublas" in k or "splitKreduce" in k:
                return "cuBLAS"
            if "at::" in k:
                return "ATen"
            return "other"
        cat = collections.defaultdict(lambda: [0, 0.0])
        for k, v in agg.items():
            c = cat[category(k)]
            c[0] += v[0]
            c[1] += v[1]
        print(f"# categories of {path}")
        for k, v in sorted(cat.items(), key=lambda kv: -kv[1][1]):
            print(f"{v[1] / 1e3:7.3f} ms {100 * v[1] / tot:5.1f}% {v[0]:6d} launches  {k}")
        print(f"# {len(body)} launches, {tot / 1e3:.3f} ms of serialised kernel time")
        return
    print(f"# {path}: {len(body)} launches, {tot / 1e3:.3f} ms of serialised kernel time; sdb:: kernels {ours / 1e3:.3f} ms "
          f"({100 * ours / tot:.1f} %)")
    print(f"# {'total_us':>10} {'share':>7} {'count':>6} {'avg_us':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.1f} {100 * v[1] / tot:6.2f}% {v[0]:6d} {v[1] / v[0]:9.2f}  {k}")


if __name__ == "__main__":
    main()
