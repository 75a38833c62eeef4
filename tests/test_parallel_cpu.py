"""world_size-2 gloo tests (CPU) of the sequence-parallel host logic: global K mean, global V |max|, rank-major
gathers and scale re-ordering must reproduce the single-process quantities."""
import os, sys
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sageattention_b200 import parallel
    from oracle import sage_oracle as O
    torch.manual_seed(0)
    B, H, S, D = 2, 3, 512, 64
    k = (torch.randn(B, H, S, D) + 3.0 * torch.randn(B, H, 1, D)).half()
    v = torch.randn(B, H, S, D).half()
    Sl = S // world
    kl, vl = k[:, :, rank * Sl:(rank + 1) * Sl], v[:, :, rank * Sl:(rank + 1) * Sl]
    # global mean from local fp32 sums == k.mean up to fp16 rounding of an fp32 mean
    km = parallel.global_k_mean(kl.float().sum(2), S, torch.float16)
    ref = k.float().mean(2)
    ok1 = (km.float() - ref).abs().max().item() <= 2 ** -10 * ref.abs().max().item() + 1e-6
    # global |V| max is exact
    am = parallel.global_abs_max(vl.float().amax(2), vl.float().amin(2))
    ok2 = torch.equal(am, v.float().abs().amax(2))
    # quantise shards with the GLOBAL mean -> gathered tensors equal the single-process quantisation
    _, _, k8l, ksl = O.quant_per_thread_int8_triton(kl, kl, km.view(B, H, 1, D))
    _, _, k8, ks = O.quant_per_thread_int8_triton(k, k, km.view(B, H, 1, D))
    k_all = parallel.gather_rank_major(k8l.contiguous())          # [P*B,H,Sl,D]
    k_cat = torch.cat([k_all[r * B:(r + 1) * B] for r in range(world)], dim=2)
    ok3 = torch.equal(k_cat, k8)
    ok4 = torch.equal(parallel.gather_scales(ksl), ks)
    ret[rank] = (ok1, ok2, ok3, ok4)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sp_host_logic_gloo_world2(world):
    port = 29500 + (os.getpid() % 2000) + world
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        assert all(ret[r]), f"rank {r}: {ret[r]}"


def _ulysses_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sageattention_b200 import parallel
    from oracle import sage_oracle as O
    torch.manual_seed(0)
    B, H, Hk, S, D = 1, 4, 2, 256, 64
    q = torch.randn(B, H, S, D).half()
    k = (torch.randn(B, Hk, S, D) + 2.0 * torch.randn(B, Hk, 1, D)).half()
    v = torch.randn(B, Hk, S, D).half()
    Sl = S // world
    sl = slice(rank * Sl, (rank + 1) * Sl)
    oks = []
    for layout, causal in (("HND", False), ("NHD", True)):
        full = O.sageattn_qk_int8_pv_fp8_cuda(q, k, v, is_causal=causal)                      # single-process result, HND
        ql, kl, vl = q[:, :, sl], k[:, :, sl], v[:, :, sl]
        if layout == "NHD":
            ql, kl, vl = (t.transpose(1, 2).contiguous() for t in (ql, kl, vl))
        o = parallel.sageattn_ulysses(ql, kl, vl, tensor_layout=layout, is_causal=causal, attn_fn=O.sageattn_qk_int8_pv_fp8_cuda)
        if layout == "NHD":
            o = o.transpose(1, 2)
        oks.append(torch.equal(o, full[:, :, sl]))     # per-head statistics only: bit-identical to the unsharded call
    # the two all_to_all re-layouts are inverses and put sequence slices in rank order
    x = torch.rand(B, H, Sl, D).floor() + float(rank)             # every element == the owning rank
    y = parallel._seq_to_head_shard(x, world)
    oks.append(y.shape == (B, H // world, S, D) and bool((y[:, :, :Sl] == 0.0).all()) and bool((y[:, :, Sl:] == 1.0).all()))
    x2 = torch.randn(B, H, Sl, D)
    oks.append(torch.equal(parallel._head_to_seq_shard(parallel._seq_to_head_shard(x2, world), world), x2))
    ret[rank] = tuple(oks)
    dist.destroy_process_group()


def test_ulysses_head_parallel_gloo_world2():
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ulysses_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        assert all(ret[r]), f"rank {r}: {ret[r]}"


def _ring_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sageattention_b200 import parallel
    from oracle import sage_oracle as O
    torch.manual_seed(0)
    B, H, Hk, S, D = 1, 4, 2, 384, 64
    q = torch.randn(B, H, S, D).half()
    k = (torch.randn(B, Hk, S, D) + 2.0 * torch.randn(B, Hk, 1, D)).half()
    v = torch.randn(B, Hk, S, D).half()
    Sl = S // world
    sl = slice(rank * Sl, (rank + 1) * Sl)
    oks = []
    for layout, causal in (("HND", False), ("HND", True), ("NHD", True)):
        exact = O.sdpa_fp32(q, k, v, is_causal=causal)[:, :, sl]
        single = O.sageattn_qk_int8_pv_fp8_cuda(q, k, v, is_causal=causal)[:, :, sl].float()
        ql, kl, vl = q[:, :, sl], k[:, :, sl], v[:, :, sl]
        if layout == "NHD":
            ql, kl, vl = (t.transpose(1, 2).contiguous() for t in (ql, kl, vl))
        o = parallel.sageattn_ring(ql, kl, vl, tensor_layout=layout, is_causal=causal, attn_fn=O.sageattn_qk_int8_pv_fp8_cuda)
        if layout == "NHD":
            o = o.transpose(1, 2)
        err_ring, err_single = (o.float() - exact).abs().max().item(), (single - exact).abs().max().item()
        # as accurate as the unsharded quantised call (per-slice smoothing / scales), far from "wrong merge" territory
        oks.append(err_ring < max(2.0 * err_single, 2e-2))
    # merge identity: two halves of the keys merged == attention over all keys (fp32 reference)
    s_ = (q.float() @ k.float().repeat_interleave(2, 1).transpose(-1, -2)) * D ** -0.5
    vf = v.float().repeat_interleave(2, 1)
    h = S // 2
    oa, ob = torch.softmax(s_[..., :h], -1) @ vf[:, :, :h], torch.softmax(s_[..., h:], -1) @ vf[:, :, h:]
    om, lm = parallel.merge_attention_states(oa, torch.logsumexp(s_[..., :h], -1), ob, torch.logsumexp(s_[..., h:], -1))
    oks.append((om - torch.softmax(s_, -1) @ vf).abs().max().item() < 1e-5 and (lm - torch.logsumexp(s_, -1)).abs().max().item() < 1e-5)
    ret[rank] = tuple(oks)
    dist.destroy_process_group()


def test_ring_attention_gloo_world2():
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ring_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        assert all(ret[r]), f"rank {r}: {ret[r]}"


def test_fused_gather_pull_schedule():
    """Issue order of the peer copies of sageattn_sp(fused_gather=True): every (chunk, source) exactly once, chunk-major (the
    attention grid runs the heads in order), own shard first, and at every step all ranks pull from DIFFERENT peers."""
    sys.path.insert(0, ROOT)
    from sageattention_b200 import parallel
    for world in (2, 3, 4, 8):
        for n_chunks in (1, 2, 4):
            scheds = [parallel.pull_schedule(world, r, n_chunks) for r in range(world)]
            for r, s in enumerate(scheds):
                assert sorted(s) == sorted((c, x) for c in range(n_chunks) for x in range(world))
                assert [c for c, _ in s] == sorted(c for c, _ in s)
                assert all(s[c * world][1] == r for c in range(n_chunks))
            for i in range(n_chunks * world):
                assert len({scheds[r][i][1] for r in range(world)}) == world
