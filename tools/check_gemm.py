"""GPU check + timing of the tcgen05 TF32 GEMM (sdb_gemm_tf32) against torch / cuBLAS.

    python tools/check_gemm.py            # every case in its own subprocess (a trap in one does not hide the rest)
    python tools/check_gemm.py --case fwd 44446 256 256
Prints one JSON line per case: max error against an fp64 product of the TF32-truncated operands (what the tensor
core computes, up to fp32 accumulation order), error of cuBLAS-TF32 against the same, and both timings.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    ("fwd", 128, 128, 32), ("fwd", 128, 128, 256), ("fwd", 256, 256, 64), ("fwd", 1000, 384, 256), ("fwd", 44446, 256, 256),
    ("fwd", 44446, 2048, 256), ("fwd", 44446, 256, 2048), ("fwd", 2200, 256, 512), ("fwd", 77, 132, 36),
    ("dx", 128, 128, 128), ("dx", 1000, 256, 384), ("dx", 44446, 256, 256), ("dx", 44446, 256, 2048), ("dx", 44446, 2048, 256),
    ("dw", 128, 128, 128), ("dw", 256, 256, 1000), ("dw", 256, 256, 44446), ("dw", 2048, 256, 44446), ("dw", 256, 2048, 44446),
    ("dw", 384, 256, 44448), ("dw", 132, 36, 76),
]


def tf32_trunc(t):
    """round to nearest TF32, ties away from zero (cvt.rna.tf32.f32): what round_mode 3 feeds the tensor core"""
    import torch
    return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def run_case(kind, m, n, k):
    """Timing: `reps` back-to-back launches between one pair of CUDA events (so host launch latency is hidden behind
    the queue), each launch on its own copy of the streamed operands -- the copies together exceed the 126 MB L2."""
    import torch
    from semi_detr_b200.layers import gemm as G
    torch.manual_seed(0)
    dev = "cuda"
    torch.backends.cuda.matmul.allow_tf32 = True
    byt = 4.0 * (m * k + n * k + m * n)
    flops = 2.0 * m * n * k
    reps = 8
    ncopy = int(max(1, min(reps, (300e6 // byt) + 1)))

    def copies(t):
        return [t] + [t.clone() for _ in range(ncopy - 1)]
    if kind == "fwd":      # y (m,n) = x (m,k) W(n,k)^T + b, relu, row mask
        x = copies(torch.randn(m, k, device=dev))
        w = torch.randn(n, k, device=dev) * 0.1
        b = torch.randn(n, device=dev)
        mask = torch.rand(m, device=dev) < 0.1
        outs = [torch.empty(m, n, device=dev) for _ in range(ncopy)]
        ours = lambda i=0: G.gemm_tf32(x[i % ncopy], 0, w, 0, m, n, k, bias=b, row_mask=mask, relu=True, out=outs[i % ncopy])
        lib = lambda i=0: torch.relu(torch.nn.functional.linear(x[i % ncopy], w, b)).masked_fill(mask[:, None], 0.0)
        lib_plain = lambda i=0: torch.addmm(b, x[i % ncopy], w.t(), out=outs[i % ncopy])
        ref = torch.relu(tf32_trunc(x[0]).double() @ tf32_trunc(w).double().t() + b.double()).masked_fill(mask[:, None], 0.0)
    elif kind == "dx":     # dx (m,n) = dy (m,k) W (k,n)
        dy = copies(torch.randn(m, k, device=dev))
        w = torch.randn(k, n, device=dev) * 0.1
        outs = [torch.empty(m, n, device=dev) for _ in range(ncopy)]
        ours = lambda i=0: G.gemm_tf32(dy[i % ncopy], 0, w, 1, m, n, k, out=outs[i % ncopy])
        lib = lambda i=0: torch.mm(dy[i % ncopy], w, out=outs[i % ncopy])
        lib_plain = lib
        ref = tf32_trunc(dy[0]).double() @ tf32_trunc(w).double()
    else:                  # dW (m,n) = dy (k,m)^T x (k,n)
        dy = copies(torch.randn(k, m, device=dev))
        x = copies(torch.randn(k, n, device=dev))
        ours = lambda i=0: G.linear_grad_weight(dy[i % ncopy], x[i % ncopy])
        lib = lambda i=0: dy[i % ncopy].t() @ x[i % ncopy]
        lib_plain = lib
        ref = tf32_trunc(dy[0]).double().t() @ tf32_trunc(x[0]).double()
    y = ours().clone()
    torch.cuda.synchronize()
    yl = lib().clone()
    scale = ref.abs().max().item() + 1e-30
    err = (y.double() - ref).abs().max().item() / scale
    err_lib = (yl.double() - ref).abs().max().item() / scale

    def timeit(f, iters=7):
        for i in range(reps):
            f(i)
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(reps):
                f(i)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
        ts.sort()
        return ts[len(ts) // 2]
    t_ours = timeit(ours)
    t_lib = timeit(lib)
    t_plain = timeit(lib_plain)
    G_round = G.gemm_tf32.__defaults__
    G.gemm_tf32.__defaults__ = G_round[:-1] + (0,)
    t_trunc = timeit(ours)
    G.gemm_tf32.__defaults__ = G_round
    print(json.dumps({"case": f"{kind} m={m} n={n} k={k}", "rel_err": err, "rel_err_cublas_tf32": err_lib, "ok": bool(err < max(2e-5, 2 * err_lib)),
                      "us": round(t_ours, 2), "us_no_rounding": round(t_trunc, 2), "us_torch_same_epilogue": round(t_lib, 2),
                      "us_cublas_gemm_only": round(t_plain, 2), "tflops": round(flops / t_ours / 1e6, 1),
                      "gbs": round(byt / t_ours / 1e3, 1), "operand_copies": ncopy}), flush=True)


if __name__ == "__main__":
    if "--case" in sys.argv:
        i = sys.argv.index("--case")
        run_case(sys.argv[i + 1], int(sys.argv[i + 2]), int(sys.argv[i + 3]), int(sys.argv[i + 4]))
    elif "--kind" in sys.argv:     # every case of one kind, small to large, stopping at the first CUDA failure
        kind = sys.argv[sys.argv.index("--kind") + 1]
        for c in CASES:
            if c[0] != kind:
                continue
            try:
                run_case(*c)
            except Exception as e:  # noqa: BLE001 -- a trap poisons the context: report and stop
                print(json.dumps({"case": c, "error": str(e)[-300:]}), flush=True)
                break
    else:
        for kind in ("fwd", "dx", "dw"):
            try:
                r = subprocess.run([sys.executable, __file__, "--kind", kind], timeout=240, capture_output=True, text=True)
                print(r.stdout.strip(), flush=True)
                if r.returncode != 0:
                    print(json.dumps({"kind": kind, "rc": r.returncode, "stderr": r.stderr[-600:]}), flush=True)
            except subprocess.TimeoutExpired as e:
                print((e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), flush=True)
                print(json.dumps({"kind": kind, "timeout": True}), flush=True)
