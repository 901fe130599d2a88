"""Bring-up check of the raw C-ABI kernels on a B200 (run under gpurun): each case runs in its own process so a
trapped kernel cannot poison the next one.  usage: python tools/gpu_kernel_check.py [case ...]"""
import subprocess
import sys
import time

CASES = ["gemm_small", "gemm_tail", "gemm_epi", "gemm_big", "attn_small", "attn_tail", "attn_cross", "attn_big"]


def rel(a, b):
    import torch
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def time_cuda(fn, iters=5, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run_case(name):
    import torch
    import ctypes as C
    sys.path.insert(0, ".")
    from stableavatar_b200 import _lib as L
    torch.manual_seed(0)
    dev = "cuda"
    lib = L.lib()

    def gemm(a, w, bias=None, act=0, res=None, gate=None, rows_per_batch=0, round_y=1, out_dtype=torch.bfloat16):
        M, K = a.shape
        N = w.shape[0]
        out = torch.empty(M, N, device=dev, dtype=out_dtype)
        g = L.GemmArgs(a=a.data_ptr(), w=w.data_ptr(), out=out.data_ptr(), bias=L.ptr(bias), res=L.ptr(res),
                       gate=L.ptr(gate), lda=a.stride(0), ldw=w.stride(0), ldc=out.stride(0),
                       ldr=res.stride(0) if res is not None else 0, gate_ld=gate.stride(0) if gate is not None else 0,
                       M=M, N=N, K=K, bias_dtype=L.dt(bias) if bias is not None else 0, out_dtype=L.dt(out),
                       res_dtype=L.dt(res) if res is not None else 0, act=act,
                       res_mode=0 if res is None else (2 if gate is not None else 1), round_y=round_y,
                       rows_per_batch=rows_per_batch)
        L.check(lib.sa_gemm_bf16(C.byref(g), L.stream_ptr()), "sa_gemm_bf16")
        return out

    def attn(q, k, v, out=None, accumulate=0):
        B, Lq, H, D = q.shape
        Lk = k.shape[1]
        if out is None:
            out = torch.empty_like(q)
        g = L.AttnArgs(q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), out=out.data_ptr(), q_bs=q.stride(0),
                       q_ls=q.stride(1), k_bs=k.stride(0), k_ls=k.stride(1), v_bs=v.stride(0), v_ls=v.stride(1),
                       o_bs=out.stride(0), o_ls=out.stride(1), batch=B, heads=H, q_len=Lq, kv_len=Lk,
                       scale=D ** -0.5, accumulate=accumulate)
        L.check(lib.sa_flash_attn_d128(C.byref(g), L.stream_ptr()), "sa_flash_attn_d128")
        return out

    def sdpa(q, k, v):
        return torch.nn.functional.scaled_dot_product_attention(
            q.transpose(1, 2).float(), k.transpose(1, 2).float(), v.transpose(1, 2).float()).transpose(1, 2)

    if name.startswith("gemm"):
        if name == "gemm_small":
            M, N, K = 256, 512, 256
        elif name == "gemm_tail":
            M, N, K = 1000, 200, 136
        elif name == "gemm_epi":
            M, N, K = 777, 1536, 1536
        else:
            M, N, K = 98280, 1536, 1536
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        bias = torch.randn(N, device=dev).bfloat16()
        out = gemm(a, w, bias)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t() + bias.float()
        print(name, "plain rel", rel(out, ref), flush=True)
        if name == "gemm_epi":
            res = torch.randn(M, N, device=dev).bfloat16()
            gate = torch.randn(3, N, device=dev).bfloat16()
            rpb = (M + 2) // 3
            out = gemm(a, w, bias, act=1, res=res, gate=gate, rows_per_batch=rpb)
            y = torch.nn.functional.gelu((a.float() @ w.float().t() + bias.float()).bfloat16().float(), approximate="tanh")
            gidx = torch.arange(M, device=dev) // rpb
            ref = res.float() + (y.bfloat16().float() * gate.float()[gidx]).bfloat16().float()
            print(name, "gelu+gated rel", rel(out, ref), flush=True)
            out = gemm(a, w, None, act=2, res=res, out_dtype=torch.float32, round_y=0)
            y = a.float() @ w.float().t()
            ref = res.float() + y * torch.sigmoid(y)
            print(name, "silu+res f32 rel", rel(out, ref), flush=True)
        if name == "gemm_big":
            ms = time_cuda(lambda: gemm(a, w, bias))
            print(name, f"{ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
            w2 = (torch.randn(8960, K, device=dev) / K ** 0.5).bfloat16()
            b2 = torch.randn(8960, device=dev).bfloat16()
            ms = time_cuda(lambda: gemm(a, w2, b2, act=1))
            print(name, f"ffn1 {ms:.3f} ms  {2.0 * M * 8960 * K / ms / 1e9:.1f} TFLOP/s", flush=True)
            ms = time_cuda(lambda: a @ w.t())
            print(name, f"cublas {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    else:
        if name == "attn_small":
            B, Lq, Lk, H = 1, 256, 256, 1
        elif name == "attn_tail":
            B, Lq, Lk, H = 2, 300, 333, 3
        elif name == "attn_cross":
            B, Lq, Lk, H = 63, 1560, 15, 12
        else:
            B, Lq, Lk, H = 1, 32760, 32760, 12
        q = torch.randn(B, Lq, H, 128, device=dev).bfloat16()
        k = torch.randn(B, Lk, H, 128, device=dev).bfloat16()
        v = torch.randn(B, Lk, H, 128, device=dev).bfloat16()
        out = attn(q, k, v)
        torch.cuda.synchronize()
        if name == "attn_big":
            ref = torch.nn.functional.scaled_dot_product_attention(
                q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)).transpose(1, 2)
        else:
            ref = sdpa(q, k, v)
        print(name, "rel", rel(out, ref), flush=True)
        if name == "attn_tail":
            base = torch.randn_like(q)
            o2 = attn(q, k, v, out=base.clone(), accumulate=1)
            print(name, "accumulate rel", rel(o2, base.float() + ref.bfloat16().float()), flush=True)
        if name == "attn_big":
            ms = time_cuda(lambda: attn(q, k, v), iters=3, warm=1)
            fl = 4.0 * B * H * Lq * Lk * 128
            print(name, f"{ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
            qq, kk, vv = (t.transpose(1, 2) for t in (q, k, v))
            ms = time_cuda(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv), iters=3, warm=1)
            print(name, f"sdpa {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run_case(sys.argv[2])
        sys.exit(0)
    cases = sys.argv[1:] or CASES
    bad = 0
    for c in cases:
        t0 = time.time()
        r = subprocess.run(["timeout", "120", sys.executable, __file__, "--case", c], capture_output=True, text=True)
        print(f"=== {c}: exit {r.returncode} ({time.time() - t0:.1f}s)")
        print(r.stdout[-3000:])
        if r.returncode != 0:
            bad += 1
            print(r.stderr[-3000:])
    sys.exit(1 if bad else 0)
