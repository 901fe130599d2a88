"""world_size-2/4 gloo tests (CPU) of the sequence-parallel routing: the head<->sequence all-to-all of
stableavatar_b200/sequence_parallel.py must hand every rank all tokens of its head group (K, V), the query tokens of
its query split (Q), and return attention outputs to the owners of the tokens — checked against a single-process
attention over the full sequence (the reference's single-GPU semantics)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _attn(q, k, v):
    s = torch.einsum("blhd,bmhd->bhlm", q, k) / q.shape[-1] ** 0.5
    return torch.einsum("bhlm,bmhd->blhd", torch.softmax(s, -1), v)


def _worker(rank, world, port, nh, q_out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stableavatar_b200 import sequence_parallel as sp
        B, L, d = 2, 8 * world, 16
        g = torch.Generator().manual_seed(0)
        q, k, v = (torch.randn(B, L, nh, d, generator=g) for _ in range(3))
        Ll = L // world
        sl = slice(rank * Ll, (rank + 1) * Ll)
        pl = sp.plan(nh, world, rank)
        Q, KV = sp.exchange_qkv(pl, q[:, sl], k[:, sl], v[:, sl])
        hs = slice(pl.g * pl.hp, (pl.g + 1) * pl.hp)
        # K/V: all tokens of my head group, token-major
        assert torch.equal(KV[:, :, 0], k[:, :, hs].permute(1, 0, 2, 3)) and torch.equal(KV[:, :, 1], v[:, :, hs].permute(1, 0, 2, 3))
        tok = torch.cat([torch.arange(s * Ll, (s + 1) * Ll) for s in pl.q_sources])
        assert torch.equal(Q, q[:, tok][:, :, hs].permute(1, 0, 2, 3))
        O = _attn(Q.transpose(0, 1), KV[:, :, 0].transpose(0, 1), KV[:, :, 1].transpose(0, 1)).transpose(0, 1).contiguous()
        out = sp.exchange_out(pl, O, B, Ll, nh, d)
        ref = _attn(q, k, v)[:, sl]
        q_out.put((rank, float((out - ref).abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nh", [(2, 12), (4, 12), (4, 6), (2, 3)])
def test_ulysses_exchange(world, nh):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nh, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert sorted(res) == list(range(world)) and max(res.values()) < 1e-5


def test_plan_head_groups_and_query_splits():
    from stableavatar_b200 import sequence_parallel as sp
    p = sp.plan(12, 8, 5)             # 12 heads on 8 GPUs: 4 head groups x 2 query splits
    assert (p.hg, p.qs, p.hp, p.g, p.s) == (4, 2, 3, 2, 1) and p.q_sources == [1, 3, 5, 7]
    p = sp.plan(12, 4, 3)             # pure Ulysses
    assert (p.hg, p.qs, p.hp, p.g, p.s) == (4, 1, 3, 3, 0) and p.q_sources == [0, 1, 2, 3]
    p = sp.plan(40, 8, 0)             # 14B: 40 heads
    assert (p.hg, p.qs, p.hp) == (8, 1, 5)


@pytest.mark.parametrize("P,nh", [(2, 12), (4, 12), (8, 12), (8, 40), (2, 40), (8, 4)])
def test_peer_scatter_index_math_gives_the_all_to_all_layout(P, nh):
    """csrc/sp_exchange.cu's destination arithmetic, restated in numpy on tagged 16-byte chunks and driven sample by
    sample like the pipelined exchange (b_first, b_count = b, 1): after every rank's scatters each rank holds K/V of its
    head group for ALL tokens ([B, source rank, Ll], i.e. batch-outermost with a uniform token stride), Q for its query
    split, and after the O scatter every rank holds [B, Ll, heads] with each head group written by its owner."""
    import math
    import numpy as np
    from stableavatar_b200 import sequence_parallel as sp
    B, Ll = 2, 3
    hg = math.gcd(nh, P)
    qs, hp = P // hg, nh // hg
    n_src = P // qs
    qkv = [(np.arange(B * Ll * 3 * nh * 16) + r * 10 ** 6).reshape(B, Ll, 3, nh, 16) for r in range(P)]
    kv_recv = [np.full(B * P * Ll * 2 * hp * 16, -1) for _ in range(P)]
    q_recv = [np.full(B * n_src * Ll * hp * 16, -1) for _ in range(P)]
    for rank in range(P):                                        # scatter_qkv_kernel, one launch per sample
        src = qkv[rank].reshape(B * Ll, 3 * nh * 16)
        for b_first in range(B):
            for e in range(3 * nh * 16):
                c, h, which = e & 15, (e >> 4) % nh, (e >> 4) // nh
                g, hl = h // hp, h % hp
                for bt in range(1 * Ll):
                    t, b = bt % Ll, b_first + bt // Ll
                    val = src[b_first * Ll + bt, e]
                    if which == 0:
                        q_recv[g * qs + rank % qs][((b * n_src + rank // qs) * Ll + t) * (hp * 16) + hl * 16 + c] = val
                    else:
                        off = ((b * P + rank) * Ll + t) * (2 * hp * 16) + ((which - 1) * hp + hl) * 16 + c
                        for s in range(qs):
                            kv_recv[g * qs + s][off] = val
    for r in range(P):
        g, pl = r // qs, sp.plan(nh, P, r)
        KV = kv_recv[r].reshape(B, P, Ll, 2, hp, 16)
        Q = q_recv[r].reshape(B, n_src, Ll, hp, 16)
        for s_r in range(P):
            assert (KV[:, s_r] == qkv[s_r][:, :, 1:3, g * hp:(g + 1) * hp]).all()             # [B, Ll, 2, hp, 16]
        for i, s_r in enumerate(pl.q_sources):
            assert (Q[:, i] == qkv[s_r][:, :, 0, g * hp:(g + 1) * hp]).all()
    o_loc = [(np.arange(B * n_src * Ll * hp * 16) + r * 10 ** 6).reshape(B, n_src, Ll, hp, 16) for r in range(P)]
    o_recv = [np.full(B * Ll * nh * 16, -1) for _ in range(P)]
    for rank in range(P):                                        # scatter_o_kernel, one launch per sample
        cols, g = hp * 16, rank // qs
        flat = o_loc[rank].reshape(-1)
        rows_b = n_src * Ll
        for b_first in range(B):
            for idx in range(1 * rows_b * cols):
                row, col = idx // cols, idx % cols
                bl, lt = row // rows_b, row % rows_b
                src, t = lt // Ll, lt % Ll
                b = b_first + bl
                o_recv[src * qs + rank % qs][(b * Ll + t) * (nh * 16) + g * cols + col] = flat[(b * rows_b + lt) * cols + col]
    for r in range(P):
        O = o_recv[r].reshape(B, Ll, nh, 16)
        for g in range(hg):
            owner = g * qs + r % qs
            assert (O[:, :, g * hp:(g + 1) * hp] == o_loc[owner][:, r // qs]).all()
