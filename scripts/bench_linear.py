"""Microbenchmark of u3d_linear_tc against cuBLAS (+ the elementwise kernels it fuses) on the decoder's shapes.
python scripts/bench_linear.py [rows]   -> one line per variant: us per launch (CUDA events, L2-warm, 50 reps)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from uni3detr_b200 import ops

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 38400
reps = int(os.environ.get("REPS", "50"))
dev = "cuda"
g = torch.Generator().manual_seed(0)
bf = lambda *s: (torch.randn(*s, generator=g) * 0.5).to(torch.bfloat16).to(dev)


def timeit(fn):
    """us per call with the launches captured in one CUDA graph (no host time between kernels)."""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    gr.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / reps


def case(name, K, N, **kw):
    a = bf(rows, K)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    gamma, beta = torch.ones(N, device=dev), torch.zeros(N, device=dev)
    lin = ops.PackedLinear(W, b, ln=(gamma, beta, 1e-5) if N <= 256 and N % 32 == 0 else None)
    args = {}
    for k in ("mul", "res1", "res2", "add2"):
        if kw.get(k):
            args[k] = bf(rows, N)
    for k in ("relu", "ln", "relu_out", "out_f32"):
        if kw.get(k):
            args[k] = True
    t_ours = timeit(lambda: ops.linear_tc(a, lin, **args))
    dbg = ""
    if os.environ.get("LIN_DEBUG"):
        t1 = timeit(lambda: ops.linear_tc(a, lin, _debug=1 << 20, **args))
        t2 = timeit(lambda: ops.linear_tc(a, lin, _debug=1 << 21, **args))
        t3 = timeit(lambda: ops.linear_tc(a, lin, _debug=1 << 22, **args))
        dbg = f"  [no global epilogue traffic {t1:6.1f}, TMEM reads only {t3:6.1f}, no epilogue {t2:6.1f}]"
    Wb, bb = W.to(torch.bfloat16), b.to(torch.bfloat16)

    def torch_fn():
        y = torch._addmm_activation(bb, a, Wb.t()) if kw.get("relu") else F.linear(a, Wb, bb)
        if "mul" in args:
            y = y * args["mul"]
        for k in ("res1", "res2"):
            if k in args:
                y = y + args[k]
        if kw.get("ln"):
            y = F.layer_norm(y, (N,), gamma.to(torch.bfloat16), beta.to(torch.bfloat16), 1e-5)
        if kw.get("relu_out"):
            y = torch.relu(y)
        if "add2" in args:
            return y, y + args["add2"]
        return y
    t_torch = timeit(torch_fn)
    fl = 2.0 * rows * K * N
    print(f"{name:34s} K={K:3d} N={N:3d}  linear_tc {t_ours:7.1f} us ({fl / t_ours / 1e6:6.0f} TFLOP/s)   torch/cuBLAS(+eltwise) {t_torch:7.1f} us{dbg}")


case("plain", 256, 256)
case("relu", 256, 256, relu=True)
case("K=384 relu (ref_point_head.0)", 384, 256, relu=True)
case("N=512 (in_proj qk)", 256, 512)
case("N=512 relu (ffn.0)", 256, 512, relu=True)
case("K=512 +res +LN (ffn.1)", 512, 256, res1=True, ln=True)
case("+res +LN (out_proj)", 256, 256, res1=True, ln=True)
case("+2res +LN (output_proj)", 256, 256, res1=True, res2=True, ln=True)
case("LN relu (cls branch)", 256, 256, ln=True, relu_out=True)
case("mul + out2 (query_scale.2)", 256, 256, mul=True, add2=True)
case("narrow fp32 N=8", 256, 8, out_f32=True)
case("narrow fp32 N=1", 256, 1, out_f32=True)
