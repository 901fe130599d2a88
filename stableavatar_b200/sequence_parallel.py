"""Sequence parallelism for the DiT over one NVSwitch domain: one process per GPU, tokens sharded contiguously along L.

Replaces the reference's xfuser path (wan/dist/wan_xfuser.py:72-114 `usp_attn_forward`, `enable_multi_gpus_inference`
wan_fantasy_transformer3d_1B.py:918-923, token chunking :1018-1019, final gather :1150-1151). Everything except
self-attention is token-local; self-attention exchanges heads for sequence with an all-to-all (Ulysses):

  P | num_heads  : pure Ulysses — rank r attends all tokens for heads [r*hp, (r+1)*hp).
  otherwise      : hg = gcd(num_heads, P) head groups x qs = P/hg query splits (12 heads on 8 GPUs -> 4 x 2): rank
                   (g, s) receives K, V of head group g from every rank but Q only from ranks r' with r' % qs == s, so
                   FLOPs stay balanced and no softmax merge (ring / LSE) is needed.

Semantics are those of the SINGLE-GPU reference (SURVEY.md fact #9-iii): RoPE positions and audio-window groups are
computed from the global token index, unlike the reference's SP path which pairs local token groups with the wrong audio
windows. L must be a multiple of P (the model rounds seq_len up, 1B.py:980-981).
"""
from __future__ import annotations

import math
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

from . import ops


def plan(num_heads: int, world: int, rank: int) -> SimpleNamespace:
    hg = math.gcd(num_heads, world)
    qs = world // hg
    return SimpleNamespace(world=world, rank=rank, hg=hg, qs=qs, hp=num_heads // hg, g=rank // qs, s=rank % qs,
                           q_sources=[r for r in range(world) if r % qs == rank % qs])


class SegmentedGraph:
    """A function captured as a chain of CUDA graphs separated by eagerly issued communication calls.

    Capturing the NCCL all-to-alls themselves inside one graph hung a 2-GPU replay (round 1), so only the compute
    between them is captured: whenever the traced function reaches `comm_point(closure)` the current capture is closed,
    the closure (a NCCL call on static buffers) runs eagerly and is recorded, and a new capture begins. All segments
    share one memory pool, so tensors produced in one segment stay valid for the next ones and for every replay.
    With no communication (single GPU) this degenerates to one ordinary CUDA graph."""
    _active = None

    def __init__(self, fn, device=None, pool=None):
        self.segments, self.cur = [], None
        self.pool = pool if pool is not None else torch.cuda.graph_pool_handle()   # shareable between window shapes
        self.stream = torch.cuda.Stream(device=device)
        self.stream.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(self.stream):
            fn()                                         # warm-up: lazy operand preparation, attribute setup, NCCL init
        torch.cuda.current_stream(device).wait_stream(self.stream)
        torch.cuda.synchronize(device)
        SegmentedGraph._active = self
        try:
            with torch.cuda.stream(self.stream):
                self._begin()
                self.out = fn()
                self._end()
        finally:
            SegmentedGraph._active = None
        torch.cuda.current_stream(device).wait_stream(self.stream)

    def _begin(self):
        self.cur = torch.cuda.CUDAGraph()
        self.cur.capture_begin(pool=self.pool)

    def _end(self):
        self.cur.capture_end()
        self.segments.append(self.cur)
        self.cur = None

    def comm(self, closure):
        self._end()
        closure()
        self.segments.append(closure)
        self._begin()

    def replay(self):
        for seg in self.segments:
            if isinstance(seg, torch.cuda.CUDAGraph):
                seg.replay()
            else:
                seg()
        return self.out


def comm_point(closure):
    """Run a communication closure now; under SegmentedGraph capture it also splits the graph there."""
    sg = SegmentedGraph._active
    if sg is None:
        closure()
    else:
        sg.comm(closure)


def all_to_all_single(out, inp, out_splits, in_splits, group=None):
    """Ragged all-to-all on flat buffers (split sizes in elements, rank order). NCCL: one grouped collective over
    NVLink. gloo (CPU tests) has no all_to_all: isend/irecv on the slices."""
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=group)
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    io = [sum(in_splits[:r]) for r in range(world + 1)]
    oo = [sum(out_splits[:r]) for r in range(world + 1)]
    out[oo[rank]:oo[rank + 1]].copy_(inp[io[rank]:io[rank + 1]])
    reqs = []
    for r in range(world):
        if r == rank:
            continue
        peer = dist.get_global_rank(group, r) if group is not None else r
        if in_splits[r]:
            reqs.append(dist.isend(inp[io[r]:io[r + 1]].contiguous(), peer, group=group))
        if out_splits[r]:
            reqs.append(dist.irecv(out[oo[r]:oo[r + 1]], peer, group=group))
    for q in reqs:
        q.wait()


def exchange_qkv(pl, q, k, v, group=None, kv=None):
    """q, k, v: local [B, Ll, nh, d] views of one fused projection buffer. Returns (Q [Lq, B, hp, d],
    KV [L, B, 2, hp, d]) for this rank's head group: all tokens for K/V, the tokens of `pl.q_sources` (in rank order)
    for Q. Token-major so that the gathered sequence has a uniform token stride for the attention kernel's TMA
    descriptors; the receive buffers are used as they land (no unpack), the send side is two permuted copies."""
    B, Ll, nh, d = q.shape
    P, hp, hg, qs = pl.world, pl.hp, pl.hg, pl.qs
    n_kv, n_q = Ll * B * 2 * hp * d, Ll * B * hp * d
    # K|V for head group g -> every rank of that group: [hg, Ll, B, 2, hp, d], repeated for the qs query splits
    with ops.timed("sp_pack"):
        if kv is None:                                       # kv: optional [B, Ll, 2, nh, d] view holding k and v side by side
            kv = torch.stack([k, v], dim=2)
        kv = kv.reshape(B, Ll, 2, hg, hp, d).permute(3, 1, 0, 2, 4, 5)
        kv_send = (kv.repeat_interleave(qs, dim=0) if qs > 1 else kv.contiguous()).reshape(-1)
        kv_recv = torch.empty(P * n_kv, device=q.device, dtype=q.dtype)
        # Q for head group g -> only the rank (g, s = my query split): [hg, Ll, B, hp, d]
        q_send = q.view(B, Ll, hg, hp, d).permute(2, 1, 0, 3, 4).contiguous().reshape(-1)
    in_splits = [n_q if dst % qs == pl.s else 0 for dst in range(P)]
    out_splits = [n_q if src in pl.q_sources else 0 for src in range(P)]
    q_recv = torch.empty(len(pl.q_sources) * n_q, device=q.device, dtype=q.dtype)

    def exchange():
        with ops.timed("sp_a2a_qkv"):
            all_to_all_single(kv_recv, kv_send, [n_kv] * P, [n_kv] * P, group)
            all_to_all_single(q_recv, q_send, out_splits, in_splits, group)
    comm_point(exchange)
    return q_recv.view(len(pl.q_sources) * Ll, B, hp, d), kv_recv.view(P * Ll, B, 2, hp, d)


def exchange_out(pl, O, B, Ll, nh, d, group=None):
    """O: [Lq, B, hp, d] attention output of this rank (its head group, the tokens of pl.q_sources). Returns the local
    [B, Ll, nh, d] with every head group filled in by its owner."""
    P, hp, hg, qs = pl.world, pl.hp, pl.hg, pl.qs
    n = Ll * B * hp * d
    in_splits = [n if dst in pl.q_sources else 0 for dst in range(P)]       # O is already ordered by source rank
    out_splits = [n if src % qs == pl.s else 0 for src in range(P)]         # one block per head-group owner
    recv = torch.empty(hg * n, device=O.device, dtype=O.dtype)
    send = O.reshape(-1)
    def exchange():
        with ops.timed("sp_a2a_o"):
            all_to_all_single(recv, send, out_splits, in_splits, group)
    comm_point(exchange)
    with ops.timed("sp_unpack"):
        return recv.view(hg, Ll, B, hp, d).permute(2, 1, 0, 3, 4).reshape(B, Ll, nh, d)


# ---------------------------------------------------------------------------------------------- NVLink peer exchange
class PeerExchange:
    """The self-attention all-to-alls as direct peer stores (csrc/sp_exchange.cu) instead of NCCL: every rank maps the
    receive buffers and flag arrays of all ranks of the group through CUDA IPC (sa_ipc_export / sa_ipc_open, handles
    exchanged once with all_gather_object). Buffers are batch-outermost — kv_recv [B, P, Ll, 2, hp, d], q_recv
    [B, n_src, Ll, hp, d], attention output o_send [B, n_src * Ll, hp, d], o_recv [B, Ll, nh, d] — so one CFG sample is a
    contiguous slice and the exchange is pipelined per sample (`attention`):

        comm stream :  norm+RoPE+scatter(b0) | barrier | norm+RoPE+scatter(b1) | barrier | ...
        stream b    :                          wait    | attention(b) | scatter_o(b)
        main stream :  join of all sample streams | barrier | (output projection reads o_recv)

    so only the first sample's scatter, the last sample's O scatter and two barriers are exposed; the rest of the NVLink
    traffic runs under the attention of the neighbouring samples. Two independent flag sets keep the comm-stream barriers
    and the main-stream barrier apart. Single buffers suffice: a peer overwrites kv/q_recv only in the next block, after
    the main-stream barrier which every rank reaches after its attention kernels; o_recv only after the next block's comm
    barrier, which this rank reaches after its output projection. Everything is ordinary kernels + events, so the whole
    step is captured in one CUDA graph (the side streams fork from and rejoin the capturing stream)."""

    def __init__(self, pl, B, Ll, nh, d, device, group=None):
        self.pl, self.shape, self.group, self.device = pl, (B, Ll, nh, d), group, device
        P, hp, n_src = pl.world, pl.hp, len(pl.q_sources)
        bf = torch.bfloat16
        self.kv_recv = torch.empty(B, P * Ll, 2, hp, d, device=device, dtype=bf)
        self.q_recv = torch.empty(B, n_src * Ll, hp, d, device=device, dtype=bf)
        self.o_send = torch.empty(B, n_src * Ll, hp, d, device=device, dtype=bf)
        self.o_recv = torch.empty(B, Ll, nh, d, device=device, dtype=bf)
        self.sig = torch.zeros(2, 64, device=device, dtype=torch.int32)          # [main | comm] flag sets
        self.epoch = torch.zeros(2, device=device, dtype=torch.int32)
        self.comm = torch.cuda.Stream(device=device)
        self.sample_streams = [torch.cuda.Stream(device=device) for _ in range(B)]
        torch.cuda.synchronize(device)
        local = (self.kv_recv, self.q_recv, self.o_recv, self.sig)
        metas = [None] * P
        dist.all_gather_object(metas, [ops.ipc_export(t) for t in local], group=group)
        self._bases, ptrs = {}, [[], [], [], []]
        with torch.cuda.device(device):                      # peer access is enabled for the device current at open time
            for r in range(P):
                for j in range(4):
                    if r == pl.rank:
                        ptrs[j].append(local[j].data_ptr())
                        continue
                    handle, off = metas[r][j]
                    if (r, handle) not in self._bases:       # buffers of one peer may share a cudaMalloc segment
                        self._bases[(r, handle)] = ops.ipc_open(handle)
                    ptrs[j].append(self._bases[(r, handle)] + off)
        self.kv_ptrs, self.q_ptrs, self.o_ptrs, sig0 = ptrs
        self.sig_ptrs = [sig0, [q + 64 * 4 for q in sig0]]
        # fused O exchange: the attention epilogue stores the tokens of source rank src * qs + s straight into that rank's
        # o_recv [B, Ll, nh, d] at this rank's head columns
        self.o_dst = [self.o_ptrs[i * pl.qs + pl.s] + pl.g * hp * d * 2 for i in range(n_src)]
        dist.barrier(group=group)                               # every rank has mapped every buffer before the first store

    def close(self):
        """Unmap the peers' allocations (the local buffers are ordinary tensors). Call on every rank once no rank will
        store into another any more; also runs when the object is collected."""
        bases, self._bases = getattr(self, "_bases", {}), {}
        for base in bases.values():
            try:
                ops.ipc_close(base)
            except Exception:  # noqa: BLE001  (context already torn down at interpreter exit)
                pass

    def __del__(self):
        try:
            if not sys.is_finalizing():      # at interpreter exit the CUDA context unmaps everything itself
                self.close()
        except Exception:  # noqa: BLE001
            pass

    def barrier(self, which=0):
        ops.sp_barrier(self.sig_ptrs[which], self.epoch[which:which + 1], self.pl.world, self.pl.rank)

    def scatter_qkv(self, qkv, norm, b_first=0, b_count=0):
        """Local [B*Ll, 3*nh*d] projection rows -> the peers' q_recv / kv_recv. norm = (weight_q, weight_k, freqs, grid,
        tok_offset): RMSNorm + RoPE of q / k fused into the scatter (the normalised values only ever exist in the receive
        buffers); norm = None: rows are already normalised."""
        B, Ll, nh, d = self.shape
        pl = self.pl
        kw = dict(B=B, Ll=Ll, heads=nh, P=pl.world, rank=pl.rank, hg=pl.hg, b_first=b_first, b_count=b_count)
        if norm is None:
            ops.sp_scatter_qkv(qkv, self.kv_ptrs, self.q_ptrs, **kw)
        else:
            wq, wk, freqs, grid, tok0 = norm
            ops.sp_norm_rope_scatter(qkv, wq, wk, self.kv_ptrs, self.q_ptrs, freqs=freqs, grid=grid, tok_offset=tok0, **kw)

    def scatter_o(self, b_first=0, b_count=0):
        B, Ll, nh, d = self.shape
        pl = self.pl
        ops.sp_scatter_o(self.o_send, self.o_ptrs, B=B, Ll=Ll, heads=nh, P=pl.world, rank=pl.rank, hg=pl.hg,
                         b_first=b_first, b_count=b_count)

    def _attend(self, b0, nb, fused_o):
        """Attention of CFG samples [b0, b0 + nb) on the receive buffers, O either stored straight into the owners' o_recv
        by the kernel's epilogue (fused_o) or into o_send for a separate scatter."""
        B, Ll, nh, d = self.shape
        K, V = self.kv_recv[b0:b0 + nb, :, 0], self.kv_recv[b0:b0 + nb, :, 1]
        if fused_o:
            off = b0 * Ll * nh * d * 2
            ops.flash_attn_sp(self.q_recv[b0:b0 + nb], K, V, [p + off for p in self.o_dst], Ll, Ll * nh * d, nh * d)
        else:
            ops.flash_attn(self.q_recv[b0:b0 + nb], K, V, out=self.o_send[b0:b0 + nb])

    def attention(self, qkv, norm, pipelined=False, fused_o=True):
        """Exchange + attention + exchange back for one block; returns o_recv [B, Ll, nh, d]."""
        B = self.shape[0]
        if not pipelined or B == 1:
            with ops.timed("sp_a2a_qkv"):
                self.scatter_qkv(qkv, norm)
                self.barrier(0)
            with ops.timed("self_attn"):
                self._attend(0, B, fused_o)
            with ops.timed("sp_a2a_o"):
                if not fused_o:
                    self.scatter_o()
                self.barrier(0)
            return self.o_recv
        main = torch.cuda.current_stream(self.device)
        with ops.timed("sp_attn_region"):
            fork = torch.cuda.Event()
            fork.record(main)
            self.comm.wait_event(fork)
            ready = []
            with torch.cuda.stream(self.comm):
                for b in range(B):
                    self.scatter_qkv(qkv, norm, b, 1)
                    self.barrier(1)
                    ev = torch.cuda.Event()
                    ev.record(self.comm)
                    ready.append(ev)
            for b in range(B):
                sb = self.sample_streams[b]
                sb.wait_event(ready[b])
                with torch.cuda.stream(sb):
                    self._attend(b, 1, fused_o)
                    if not fused_o:
                        self.scatter_o(b, 1)
                    ev = torch.cuda.Event()
                    ev.record(sb)
                main.wait_event(ev)
            self.barrier(0)
        return self.o_recv


def _probe_peer(model, device):
    """Can this group use the peer-store exchange? Every rank must be a CUDA device of one NCCL group with working CUDA
    IPC + P2P towards every other rank; the ranks agree on the answer (all_reduce MIN) so nobody takes a different path."""
    ok = device.type == "cuda" and dist.get_backend(model.sp_group) == "nccl"
    if not ok:
        return False
    good = 1
    try:
        probe = torch.zeros(64, device=device, dtype=torch.int32)
        metas = [None] * model.sp_world_size
        dist.all_gather_object(metas, ops.ipc_export(probe), group=model.sp_group)
        with torch.cuda.device(device):
            for r, (handle, off) in enumerate(metas):
                if r != model.sp_world_rank:
                    ops.ipc_close(ops.ipc_open(handle))
    except Exception:  # noqa: BLE001  (no P2P / IPC, e.g. expandable_segments allocations or ranks on different nodes)
        good = 0
    flag = torch.tensor([good], device=device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=model.sp_group)
    return bool(flag.item())


def _peer_exchange(model, B, Ll, nh, d, device):
    """The model's PeerExchange for these shapes, or None for the NCCL all_to_all_single path. model.sp_exchange:
    "auto" (default: peer stores when the probe succeeds on every rank, else NCCL), "peer" (raise if unavailable), "nccl"."""
    mode = getattr(model, "sp_exchange", "auto")
    if mode == "nccl" or device.type != "cuda":
        return None
    ok = getattr(model, "_sp_peer_ok", None)
    if ok is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("sequence_parallel: the first sequence-parallel forward must run eagerly (it probes and "
                               "maps the peer buffers) before CUDA-graph capture")
        ok = model._sp_peer_ok = _probe_peer(model, device)
    if not ok:
        if mode == "peer":
            raise RuntimeError("sequence_parallel: sp_exchange='peer' but CUDA IPC / P2P is not available on every rank")
        return None
    px = getattr(model, "_sp_px", None)
    if px is None or px.shape != (B, Ll, nh, d):
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("sequence_parallel: the peer exchange buffers must be created by an eager (warm-up) "
                               "forward before CUDA-graph capture")
        px = model._sp_px = PeerExchange(model._sp, B, Ll, nh, d, device, model.sp_group)
    return px


# ---------------------------------------------------------------------------------------------- model hooks
def shard_tokens(model, h, st):
    """1B.py:1018-1019: keep this rank's contiguous chunk of the (padded) token sequence."""
    B, L, C = st["B"], st["L"], st["C"]
    P, r = model.sp_world_size, model.sp_world_rank
    Ll = L // P
    st = dict(st, Ll=Ll, tok0=r * Ll)
    return h.view(B, L, C)[:, r * Ll:(r + 1) * Ll].reshape(B * Ll, C).contiguous(), st


def self_attention(model, qkv, sa, st):
    """xf.py:72-114 with global RoPE positions: RMSNorm+RoPE on the local shard, all-to-all, attention over the full
    sequence for this rank's heads, all-to-all back. Returns [B, Ll, nh, 128]."""
    B, C, nh, Ll = st["B"], st["C"], st["nh"], st["Ll"]
    pl = model._sp
    px = _peer_exchange(model, B, Ll, nh, 128, qkv.device)
    if px is not None:
        norm = (sa.norm_q.weight, sa.norm_k.weight, st["freqs"], st["grid"], st["tok0"])
        if not getattr(model, "sp_fused_norm", True):        # test knob: norm in place, then the plain scatter
            ops.rmsnorm_rope_(qkv[:, :C], norm[0], qkv[:, C:2 * C], norm[1], freqs=norm[2], grid=norm[3],
                              rows_per_batch=Ll, tok_offset=norm[4])
            norm = None
        return px.attention(qkv, norm, pipelined=getattr(model, "sp_pipelined", False),
                            fused_o=getattr(model, "sp_fused_o", True))
    ops.rmsnorm_rope_(qkv[:, :C], sa.norm_q.weight, qkv[:, C:2 * C], sa.norm_k.weight, freqs=st["freqs"],
                      grid=st["grid"], rows_per_batch=Ll, tok_offset=st["tok0"])
    q5 = qkv.view(B, Ll, 3, nh, 128)
    Q, KV = exchange_qkv(pl, q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], model.sp_group, kv=q5[:, :, 1:3])
    with ops.timed("self_attn"):
        O = ops.flash_attn(Q.transpose(0, 1), KV[:, :, 0].transpose(0, 1), KV[:, :, 1].transpose(0, 1),
                           out=torch.empty_like(Q).transpose(0, 1))
    return exchange_out(pl, O.transpose(0, 1), B, Ll, nh, 128, model.sp_group)


def audio_attention(model, q, kvv, a, st, Ll):
    """Grouped audio cross-attention on a token shard: global token t belongs to audio window t // (L/G) (the
    single-GPU view(b*G, ...) of 1B.py:575-586); a shard straddles a few windows, one launch per window."""
    B, C, nh, G, L, tok0 = st["B"], st["C"], st["nh"], st["G"], st["L"], st["tok0"]
    if L % G != 0:
        raise RuntimeError(f"shape '[{B * G}, -1, {nh}, 128]' is invalid for input of size {B * L * C}")
    gs = L // G
    q4, a4 = q.view(B, Ll, nh, 128), a.view(B, Ll, nh, 128)
    kg = kvv.view(B, G, -1, 2, nh, 128)
    for g in range(tok0 // gs, (tok0 + Ll - 1) // gs + 1):
        lo, hi = max(g * gs, tok0) - tok0, min((g + 1) * gs, tok0 + Ll) - tok0
        ops.flash_attn(q4[:, lo:hi], kg[:, g, :, 0], kg[:, g, :, 1], out=a4[:, lo:hi], accumulate=True)


def gather_tokens(model, u):
    """Gather the 64-wide head output of every shard (instead of the 1536-wide hidden states, 1B.py:1150-1154)."""
    P = model.sp_world_size
    parts = [torch.empty_like(u) for _ in range(P)]
    src = u.contiguous()
    comm_point(lambda: dist.all_gather(parts, src, group=model.sp_group))
    return torch.cat(parts, dim=1)
