"""CPU checks of the arithmetic behind the fp32 mode (csrc/fp32_kernels.cu, stableavatar_b200/fp32_mode.py): the
three-term bf16 split is exact to fp32 precision and the six cross products kept by the K-concatenation
[x0|x0|x1|x1|x0|x2] . [w0|w1|w0|w1|w2|w0] approximate the fp32 product to ~2^-24 when summed without loss."""
import torch


def split3(x):
    t0 = x.bfloat16().float()
    r1 = x - t0
    t1 = r1.bfloat16().float()
    t2 = (r1 - t1).bfloat16().float()
    return t0, t1, t2


def test_three_bf16_terms_reconstruct_fp32_exactly():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4096, generator=g) * torch.logspace(-6, 6, 4096)
    t0, t1, t2 = split3(x)
    assert torch.equal(t0 + t1 + t2, x)                      # 3 x 8 mantissa bits cover fp32's 24


def test_six_cross_products_are_fp32_accurate():
    g = torch.Generator().manual_seed(1)
    x, w = torch.randn(64, 512, generator=g) * 3, torch.randn(48, 512, generator=g)
    xs, ws = split3(x), split3(w)
    pattern_x, pattern_w = (0, 0, 1, 1, 0, 2), (0, 1, 0, 1, 2, 0)
    a = torch.cat([xs[i] for i in pattern_x], dim=1).double()                 # [M, 6K] as sa_f32_split3 lays it out
    b = torch.cat([ws[i] for i in pattern_w], dim=1).double()
    got = a @ b.t()                                                            # exact products, lossless accumulation
    want = x.double() @ w.double().t()
    rel = ((got - want).norm() / want.norm()).item()
    assert rel < 2e-7, rel                                                     # dropped terms: x1 w2, x2 w1, x2 w2 (<= 2^-24)
    plain = (x.bfloat16().double() @ w.bfloat16().double().t())
    assert ((plain - want).norm() / want.norm()).item() > 1e-3                 # what a single bf16 GEMM would give


def test_k_chunks_cover_every_column_once():
    from stableavatar_b200.fp32_mode import KC, _chunks
    for K in (1, 8, 144, 256, 257, 1536, 8960):
        ch = _chunks(K)
        assert ch[0][0] == 0 and ch[-1][1] == K and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
        assert all(0 < k1 - k0 <= KC for k0, k1 in ch)
