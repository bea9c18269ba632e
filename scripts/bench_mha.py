"""A/B of the self-attention core: u3d_mha_core (tcgen05, csrc/mha_tc.cu) vs torch SDPA (every backend torch offers)
on the decoder's shape: n_seq sequences x 8 heads x seq_len x 32, bf16. CUDA-graph timed (no host time)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel
from uni3detr_b200 import ops

reps = 20


def timeit(fn):
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


for n_seq, L in ((128, 300), (32, 300), (8, 900), (128, 900)):
    H, hd = 8, 32
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(n_seq * L, 3 * H * hd, generator=g) * 0.5).to(torch.bfloat16).cuda()
    q, k, v = qkv[:, :256], qkv[:, 256:512], qkv[:, 512:]
    t_ours = timeit(lambda: ops.mha_core(q, k, v, n_seq, L, H))
    out = ops.mha_core(q, k, v, n_seq, L, H)
    q4 = q.reshape(n_seq, L, H, hd).transpose(1, 2).contiguous()
    k4 = k.reshape(n_seq, L, H, hd).transpose(1, 2).contiguous()
    v4 = v.reshape(n_seq, L, H, hd).transpose(1, 2).contiguous()
    ref = F.scaled_dot_product_attention(q4.float(), k4.float(), v4.float()).transpose(1, 2).reshape(n_seq * L, H * hd)
    err = float((out.float() - ref).abs().max())
    line = f"n_seq={n_seq:4d} L={L:4d}: u3d_mha_core {t_ours:7.1f} us (max err vs fp32 SDPA {err:.3e})"
    flops = 4.0 * n_seq * H * L * L * hd
    line += f" = {flops / t_ours / 1e6:6.1f} TFLOP/s;  torch SDPA:"
    for name, be in (("flash", SDPBackend.FLASH_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION),
                     ("cudnn", SDPBackend.CUDNN_ATTENTION), ("math", SDPBackend.MATH)):
        try:
            with sdpa_kernel(be):
                t = timeit(lambda: F.scaled_dot_product_attention(q4, k4, v4))
            line += f" {name} {t:7.1f} us"
        except Exception as ex:   # noqa: BLE001
            line += f" {name} n/a"
    print(line, flush=True)
