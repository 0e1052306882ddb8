"""-m gpu parity at BASELINE.json's FULL sizes: every row of the C2, C3B, C3 (shared prefix,
cascade), C3U and C5-shard decode batches against the library's independent one-warp-per-(request, head) fp32 kernel
on the same device buffers -- element-wise and in norm -- plus sampled rows against the CPU oracle; the C4 extend
chunk on sampled query rows against the CPU oracle.  Uses the very state builder and checkers bench.py uses, so what
the bench's `parity_check` reports is what is asserted here."""
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import bench


def _args(*argv):
    old = sys.argv
    sys.argv = ["bench.py", *argv]
    try:
        return bench.parse_args()
    finally:
        sys.argv = old


DECODE = [
    # workload, shard, extra flags
    ("C2", 1, ()),
    ("C2", 8, ()),                      # TP8 shard: one kv head, fused K|V rows
    ("C3B", 1, ()),
    ("C3B", 1, ("--page-size", "16")),  # paged pool, TMA tiles
    ("C3U", 1, ()),
    ("C3", 1, ()),                      # shared 4k prefix: cascade leg + split-KV suffixes + merge
    ("C3", 1, ("--cascade", "off")),
    ("C5", 8, ()),                      # 70B TP8 shard, 1k-32k contexts
]


@pytest.mark.parametrize("name,shard,extra", DECODE)
def test_decode_full_size(name, shard, extra):
    args = _args("--workload", name, "--shard", str(shard), "--layers", "1", *extra)
    torch.cuda.set_device(0)
    st = bench.build_state(args, 0, shard)
    w, fb, backend = st["w"], st["fb"], st["backend"]
    B, hq, hkv, d = w.batch, st["hq"], st["hkv"], st["d"]
    g = torch.Generator(device="cuda").manual_seed(5)
    qkv = torch.randn((B, (hq + 2 * hkv) * d), generator=g, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, : hq * d], qkv[:, hq * d: (hq + hkv) * d], qkv[:, (hq + hkv) * d:]
    backend.init_forward_metadata(fb)
    o = st["layers"][0](q, k, v, fb, save_kv_cache=True)
    torch.cuda.synchronize()
    res = bench.decode_parity(st, q.view(B, hq, d), {0: o})
    assert res["ok"], res
    if name == "C3" and "off" not in extra:
        assert int(backend.workspace[:64].view(torch.int32)[7].item()) == 4096, "the shared prefix was not detected"
    bench.free_cuda()


def test_extend_full_size_c4():
    args = _args("--workload", "C4")
    torch.cuda.set_device(0)
    r = bench.measure_extend(args, 0, 1, 0, None, 2, 3, want_e2e=False, want_parity=True)
    assert r["parity"]["ok"], r["parity"]
    bench.free_cuda()
