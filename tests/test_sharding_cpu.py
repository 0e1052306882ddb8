"""world_size-2 gloo tests (CPU) of the N>1 host logic: kv-head sharding reproduces the unsharded
result with no collective on the attention path, and job time is the max over ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import attn_oracle as O
from scratchpad_b200.sharding import head_shard, max_over_ranks, shard_decode_inputs


def test_head_shard_layout_matches_engine_tp():
    s = [head_shard(32, 8, 4, r) for r in range(4)]
    assert [x.hq_local for x in s] == [8] * 4 and [x.hkv_local for x in s] == [2] * 4
    assert [x.kv_head_start for x in s] == [0, 2, 4, 6] and [x.q_head_start for x in s] == [0, 8, 16, 24]
    # 70B at TP8: 64 q / 8 kv heads -> 8 q + 1 kv per rank
    s = head_shard(64, 8, 8, 5)
    assert (s.hq_local, s.hkv_local, s.kv_head_start, s.kv_replicas) == (8, 1, 5, 1)
    # tp > Hkv replicates kv heads (linear.py:716-722)
    s = [head_shard(32, 2, 8, r) for r in range(8)]
    assert [x.kv_head_start for x in s] == [0, 0, 0, 0, 1, 1, 1, 1] and all(x.hkv_local == 1 and x.kv_replicas == 4 for x in s)
    # each rank's q heads belong to the kv head it reads
    for x in s:
        g = 32 // 2
        assert x.q_head_start // g == x.kv_head_start
    with pytest.raises(ValueError):
        head_shard(32, 8, 3, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # TP replicas build identical state (tp_worker.py:100-105)
        hq, hkv, d, pool = 8, 2, 64, 200
        gen = torch.Generator().manual_seed(11)
        k = torch.randn((pool + 1, hkv, d), generator=gen).to(torch.bfloat16)
        v = torch.randn((pool + 1, hkv, d), generator=gen).to(torch.bfloat16)
        q = torch.randn((3, hq, d), generator=gen).to(torch.bfloat16)
        r2t = torch.randint(1, pool + 1, (4, 70), generator=gen, dtype=torch.int32)
        rpi, sl = torch.tensor([2, 0, 3]), torch.tensor([64, 5, 33])
        sh = head_shard(hq, hkv, world, rank)
        ql, kl, vl = shard_decode_inputs(q, k, v, sh)
        out_local = O.decode_ref(ql, kl, vl, r2t, rpi, sl, d ** -0.5)        # no collective on the path
        gathered = [torch.zeros_like(out_local) for _ in range(world)]
        dist.all_gather(gathered, out_local.contiguous())                    # test-only: reassemble heads
        full = O.decode_ref(q, k, v, r2t, rpi, sl, d ** -0.5)
        err = (torch.cat(gathered, dim=1) - full).abs().max().item()
        t = max_over_ranks(1.0 + rank, dist)                                  # slowest rank defines the job time
        ret[rank] = (err, t)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank in range(world):
        err, t = ret[rank]
        assert err < 1e-6, err
        assert t == 2.0


def test_paged_allocator_and_table_keep_the_page_invariant():
    """Host logic of page_size > 1 (no GPU): slot = page * page_size + position % page_size, pages never shared
    between requests except whole pages of a common prefix, page 0 never handed out."""
    import torch
    import scratchpad_b200 as sb
    from harness import synth

    ps = 16
    gen = torch.Generator().manual_seed(0)
    table = torch.zeros((7, 400), dtype=torch.int32)
    lens = [260, 33, 16, 1, 290, 48]
    rows = synth.build_paged_table(table, lens, 4096, ps, gen, shared_prefix=40)
    seen = {}
    for i, n in enumerate(lens):
        r = table[rows[i], :n]
        pos = torch.arange(n)
        assert torch.equal(r, r[pos - pos % ps] + (pos % ps).int()) and (r[pos - pos % ps] % ps == 0).all()
        assert r.min() >= ps
        for pg in set((r // ps).tolist()):
            seen.setdefault(pg, []).append(i)
    shared_pages = set((table[rows[0], :32] // ps).tolist())       # 40 // 16 = 2 whole pages
    assert all(len(v) == 1 or k in shared_pages for k, v in seen.items())
    assert all(sorted(seen[k]) == [0, 1, 4, 5] for k in shared_pages)  # every request of >= 32 tokens

    pool = sb.MHATokenToKVPool(256, ps, torch.bfloat16, 2, 64, 1, "cpu")
    assert pool.k_buffer[0].shape == (256 + ps, 2, 64)
    al = sb.PagedTokenToKVPoolAllocator(256, ps, torch.bfloat16, "cpu", pool)
    assert al.available_size() == 256 and al.page_size == ps
    pre, seq = torch.tensor([0, 20]), torch.tensor([35, 41])
    last = torch.tensor([-1, 5 * ps + 3])                            # position 19 lives at offset 3 of page 5
    loc = al.alloc_extend(pre, seq, last, 35 + 21)
    assert loc.numel() == 56 and al.available_size() == 256 - 4 * ps  # 3 pages + 1 fresh page behind the finished one
    pos = torch.cat([torch.arange(0, 35), torch.arange(20, 41)])
    assert torch.equal(loc % ps, pos % ps)
    assert loc[35:35 + 12].tolist() == list(range(5 * ps + 4, 6 * ps))  # the partially filled page is finished first
    state = al.backup_state()
    d = al.alloc_decode(torch.tensor([36, 49]), torch.tensor([int(loc[34]), 100]))
    assert d[0] == loc[34] + 1 and d[1] % ps == 0 and al.available_size() == 256 - 5 * ps
    al.restore_state(state)
    assert al.available_size() == 256 - 4 * ps
    al.free(loc[:32])                                                 # two whole pages come back
    assert al.available_size() == 256 - 2 * ps
    assert al.alloc(10 * ps + ps * 20) is None
    import pytest
    with pytest.raises(NotImplementedError):
        sb.MHATokenToKVPool(32, 8, torch.bfloat16, 2, 64, 1, "cpu")
