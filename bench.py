#!/usr/bin/env python
"""bench.py -- decode-attention tokens/s and achieved HBM GB/s of the cuda_b200 backend.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2]

A *step* is one layer-call of the hot path over one synthetic batch:
``set_kv_buffer`` (fused KV store) + split-KV paged decode + LSE merge, i.e. one
``AttentionBackend.forward_decode`` as the engine issues it per layer
(scratchpad/nn/attention/triton_backend.py:157-196 in the reference), with
``init_forward_metadata`` once per model step (every ``layers`` steps, as
model_runner.py:506-514 does).  Steps rotate over ``layers`` distinct per-layer
KV pools (>> L2), so every step streams its KV from HBM.

Default workload = BASELINE.json configs[1] (C2): Llama-3.2-1B decode, batch 256,
2k context, GQA 32q/8kv, D=64, bf16 KV, page size 1, random-permuted slot ids.
At N GPUs the kv heads are sharded N ways (the engine's tensor-parallel layout);
every rank serves the same batch with Hkv/N heads, no collective on the path.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2048)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--index-mode", default="random", choices=["random", "contiguous"])
    ap.add_argument("--layers", type=int, default=0, help="distinct per-layer KV pools rotated (0 = auto)")
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"],
                    help="device-resident leg: replay one CUDA graph per layer-call (engine default) or eager calls")
    ap.add_argument("--cascade", default="auto", choices=["auto", "on", "off"],
                    help="shared-prefix cascade decode (server flag b200_shared_prefix_cascade; the backend's own default "
                         "is on).  auto = what an operator would set for the deployment the workload stands for: on when "
                         "its requests share a prompt prefix (C3), off otherwise -- an idle cascade costs one early-exit "
                         "launch (~2 us) per layer-call")
    ap.add_argument("--force-path", default="auto", choices=["auto", "mma", "tc"],
                    help="extend workloads: kernel family (mma = legacy mma.sync kernel, tc = tcgen05/TMEM kernel)")
    ap.add_argument("--rope", default="off", choices=["off", "separate", "fused"],
                    help="decode only: include the model's rotary embedding in the layer-call -- 'separate' = the "
                         "reference's sequence rotary_emb(positions, q, k) then attention (llama.py:149-152), 'fused' = "
                         "one b200_rope_kv_store launch doing the rotation and the KV write (SURVEY 8f #1)")
    ap.add_argument("--page-size", type=int, default=1, choices=[-1, 1, 16, 32, 64, 128],
                    help="decode only: KV pool page size.  1 = the reference's token-granular pool (default, the headline "
                         "configuration); > 1 = paged pool, random pages, whole tiles fetched by TMA (SURVEY 8f #4)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--kv-layout", default="auto", choices=["auto", "split", "fused"],
                    help="split: the reference's separate k/v buffers; fused: K|V rows of a token adjacent "
                         "(pool option of this backend); auto: fused when a per-buffer token row is < 256 B")
    ap.add_argument("--shard", type=int, default=1,
                    help="development aid: run the per-GPU shard shapes of a (gpus*shard)-way head split on this GPU")
    return ap.parse_args()


# ----------------------------------------------------------------------------
# clocks (nvidia-smi equivalent through NVML), sampled DURING the timed region
# ----------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period_s: float = 0.01):
        self.samples, self.reasons, self.power = [], set(), []
        self.period = period_s
        self._stop = threading.Event()
        self._thr = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # honour CUDA_VISIBLE_DEVICES: map the torch index to the NVML index through the UUID
            uuid = None
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        self.h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        continue
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if r & bit and name != "gpu_idle":
                    self.reasons.add(name)
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample()
            self._stop.wait(self.period)

    def start(self):
        if self.ok:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": getattr(self, "err", "no samples")}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": max(self.power) if self.power else None}


# ----------------------------------------------------------------------------
# workload construction
# ----------------------------------------------------------------------------
def fill_pool(pool, seed: int):
    """bf16-valued N(0,1) KV in every layer (chunked); fp8 pools get fixed per-head scales."""
    dev = pool.device
    g = torch.Generator(device=dev).manual_seed(seed)
    fp8 = pool.dtype != torch.bfloat16
    h = pool.head_num
    for layer_id in range(pool.layer_num):
        if fp8:
            ks = (6.0 / 448.0) * (1.0 + 0.05 * torch.arange(h, dtype=torch.float32, device=dev))
            pool.k_scale[layer_id].copy_(ks)
            pool.v_scale[layer_id].copy_(ks.flip(0))
        for which, buf in (("k", pool.k_buffer[layer_id]), ("v", pool.v_buffer[layer_id])):
            n_tok = buf.shape[0]
            chunk = max(1, (1 << 24) // (buf.shape[1] * buf.shape[2]))
            for s in range(0, n_tok, chunk):
                n = min(chunk, n_tok - s)
                x = torch.randn((n, buf.shape[1], buf.shape[2]), generator=g, device=dev, dtype=torch.float32)
                if not fp8:
                    buf[s:s + n] = x.to(torch.bfloat16)
                else:
                    sc = (pool.k_scale if which == "k" else pool.v_scale)[layer_id].view(1, -1, 1)
                    y = (x.to(torch.bfloat16).float() / sc).clamp_(-448.0, 448.0).to(torch.bfloat16)
                    buf[s:s + n] = y.to(pool.dtype).view(torch.uint8)
            buf[0].zero_()


def pick_layers(per_layer_bytes: int, want: int, total_layers: int) -> int:
    if want > 0:
        return want
    budget = 40e9  # keep the rotated pools well inside 180 GB and the build fast
    return int(max(2, min(total_layers, 16, budget // max(per_layer_bytes, 1))))


def build_state(args, rank, world):
    from scratchpad_b200 import CudaB200AttnBackend
    from harness import synth

    w = synth.WORKLOADS[args.workload]
    assert w.hkv_total % world == 0 and w.hq_total % world == 0, "kv heads must divide over the GPUs"
    gen = torch.Generator().manual_seed(args.seed)  # same page table on every rank (TP replicas)
    seq_lens = synth.workload_seq_lens(w, gen)
    from scratchpad_b200.sharding import head_shard
    sh = head_shard(w.hq_total, w.hkv_total, world, rank % world)
    hkv_l, hq_l = sh.hkv_local, sh.hq_local
    shared = w.shared_prefix
    need = shared + sum(n - shared if shared and n >= shared else n for n in seq_lens)
    pool_tokens = need + 64
    ps = max(args.page_size, 1)
    if ps > 1:
        pool_tokens = (shared // ps + sum((n + ps - 1) // ps for n in seq_lens) + 4) * ps
    kv_elt = 2 if w.kv_dtype == torch.bfloat16 else 1
    per_layer = 2 * (pool_tokens + 1) * hkv_l * w.head_dim * kv_elt
    layers = pick_layers(per_layer, args.layers, w.layers_total)
    dev = f"cuda:{torch.cuda.current_device()}"
    row_bytes = hkv_l * w.head_dim * kv_elt
    fused = args.kv_layout == "fused" or (args.kv_layout == "auto" and row_bytes < 256)
    runner = synth.make_model_runner(w.hq_total, w.hkv_total, w.head_dim, tp_size=world, layer_num=layers,
                                     pool_tokens=pool_tokens, max_reqs=w.batch, context_len=max(seq_lens) + 1,
                                     kv_dtype=w.kv_dtype, device=dev, fused_kv=fused, page_size=ps)
    fill_pool(runner.token_to_kv_pool, 1000 + rank % max(world, 1))
    cascade = args.cascade == "on" or (args.cascade == "auto" and w.shared_prefix > 0)
    backend = CudaB200AttnBackend(runner, cascade=cascade, tma_contiguous_runs=args.page_size == -1)
    fb = synth.make_decode_batch(runner, seq_lens, gen, mode=args.index_mode, shared_prefix=shared,
                                 idx_dtype=torch.int64, backend=backend)
    lay = [synth.make_layer(runner, i) for i in range(layers)]
    algo_bytes = synth.decode_algorithmic_bytes(seq_lens, hq_l, hkv_l, w.head_dim, w.head_dim, kv_elt)
    uniq_tokens = need
    return dict(w=w, seq_lens=seq_lens, runner=runner, backend=backend, fb=fb, layers=lay, L=layers,
                hq=hq_l, hkv=hkv_l, d=w.head_dim, algo_bytes=algo_bytes, kv_elt=kv_elt, fused_kv=fused,
                row_bytes=row_bytes,
                unique_kv_bytes=uniq_tokens * hkv_l * 2 * w.head_dim * kv_elt, per_layer_bytes=per_layer)


# ----------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's torch-native port on host cores
# ----------------------------------------------------------------------------
def cpu_sample_state(w, seq_lens, hq, hkv, sample_reqs: int, seed: int):
    from oracle import attn_oracle as O
    from harness import synth

    gen = torch.Generator().manual_seed(seed)
    lens = list(seq_lens[:sample_reqs])
    shared = w.shared_prefix
    need = shared + sum(n - shared if shared and n >= shared else n for n in lens)
    pool = need + 16
    r2t = torch.zeros((len(lens) + 1, max(lens) + 4), dtype=torch.int32)
    rpi = synth.build_page_table(r2t, lens, pool, gen, "random", shared)
    d = w.head_dim
    k = torch.randn((pool + 1, hkv, d), generator=gen).to(torch.bfloat16)
    v = torch.randn((pool + 1, hkv, d), generator=gen).to(torch.bfloat16)
    ks = vs = None
    if w.kv_dtype != torch.bfloat16:
        ks = torch.full((hkv,), 6.0 / 448.0)
        vs = ks.clone()
        k = O.quantize_kv_ref(k, w.kv_dtype, ks)
        v = O.quantize_kv_ref(v, w.kv_dtype, vs)
    q = torch.randn((len(lens), hq, d), generator=gen).to(torch.bfloat16)
    k_new = torch.randn((len(lens), hkv, d), generator=gen).to(torch.bfloat16)
    v_new = torch.randn((len(lens), hkv, d), generator=gen).to(torch.bfloat16)
    sl = torch.tensor(lens, dtype=torch.int64)
    loc = torch.stack([r2t[rpi[i], n - 1] for i, n in enumerate(lens)]).long()
    return dict(q=q, k=k, v=v, r2t=r2t, rpi=rpi, sl=sl, ks=ks, vs=vs, k_new=k_new, v_new=v_new, loc=loc,
                scale=d ** -0.5, kv_dtype=w.kv_dtype, lens=lens)


def cpu_step(s):
    """One layer-call of the torch-native restatement: KV write, gather, SDPA (oracle/attn_oracle.py)."""
    from oracle import attn_oracle as O

    O.set_kv_buffer_ref(s["k"], s["v"], s["loc"], s["k_new"], s["v_new"], s["kv_dtype"], s["ks"], s["vs"])
    return O.decode_torch_native(s["q"], s["k"], s["v"], s["r2t"], s["rpi"], s["sl"], s["scale"],
                                 math_dtype=torch.float32, kv_dtype=s["kv_dtype"], k_scale=s["ks"], v_scale=s["vs"])


def time_cpu(s, steps: int, warmup: int, max_seconds: float = 120.0):
    for _ in range(warmup):
        cpu_step(s)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        cpu_step(s)
        done += 1
        if time.perf_counter() - t0 > max_seconds:
            break
    return (time.perf_counter() - t0) / done, done


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The fork ships no
    torch-native backend and its Triton/FlashInfer backends need a GPU, so this
    times the oracle port ("port") with every host thread, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from harness import synth

    w = synth.WORKLOADS[args.workload]
    gen = torch.Generator().manual_seed(args.seed)
    seq_lens = synth.workload_seq_lens(w, gen)
    # the CPU arm has no GPUs to shard over: it always computes every head of the workload (the job the N GPUs share),
    # with every host core (torchrun exports OMP_NUM_THREADS=1 to its workers, so the count is set explicitly)
    hq, hkv = w.hq_total, w.hkv_total
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, RuntimeError):
        torch.set_num_threads(os.cpu_count() or 1)
    sample = min(w.batch, 16)
    s = cpu_sample_state(w, seq_lens, hq, hkv, sample, args.seed)
    steps = min(args.steps, 64)
    warm = min(max(args.warmup, 1), 3)
    dt, done = time_cpu(s, steps, warm)
    cores = torch.get_num_threads()
    val = sample / dt
    desc = f"{sample} of {w.batch} requests of {w.name} (same shapes, ctx as in the workload), {done} steps"
    line = {
        "impl": "reference", "metric": "decode_attn_tokens_per_s", "value": val,
        "unit": "tokens/s per layer-call", "n_gpus": world, "steps": done, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if w.kv_dtype == torch.bfloat16 else "fp8_e4m3", "data": "synthetic",
        "config": workload_config(w, world, args, None),
        "cpu_baseline": {"value": val, "unit": "tokens/s per layer-call", "cores": cores, "kind": "port",
                         "sample": desc},
        "e2e": {"value": val, "unit": "tokens/s per layer-call", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(w, world, args, st):
    cfg = {
        "workload": f"{w.name}: {w.description}", "batch": w.batch, "hq_per_gpu": w.hq_total // world,
        "hkv_per_gpu": w.hkv_total // world, "head_dim": w.head_dim, "page_size": args.page_size,
        "kv_dtype": "bf16" if w.kv_dtype == torch.bfloat16 else "fp8_e4m3", "index": args.index_mode,
        "parallelism": f"kv-head sharded x{world} (tensor-parallel head layout), no collective on the path",
        "step": "one forward_decode layer-call: fused KV store + split-KV paged decode + LSE merge; "
                "init_forward_metadata once per `layers` steps",
    }
    if st is not None:
        cfg.update({
            "layers": st["L"], "sum_seq_len": int(sum(st["seq_lens"])), "mode": args.mode,
            "kv_layout": ("fused [P+1,2,Hkv,D] (K|V rows adjacent)" if st["fused_kv"] else
                          "split k_buffer/v_buffer [P+1,Hkv,D] (reference layout)"),
            "token_row_bytes_per_buffer": st["row_bytes"],
            "l2": f"inputs larger than L2: {st['L']} per-layer KV pools of {st['per_layer_bytes'] / 1e9:.2f} GB "
                  "rotated step to step",
        })
    return cfg



# ----------------------------------------------------------------------------
# extend / chunked-prefill workloads (C4): tensor-bound, metric TFLOP/s
# ----------------------------------------------------------------------------
def run_extend_bench(args, rank, world, local_rank, dist):
    from scratchpad_b200 import CudaB200AttnBackend
    from harness import synth

    w = synth.WORKLOADS[args.workload]
    tp = world * args.shard
    hq, hkv, d = w.hq_total // tp, w.hkv_total // tp, w.head_dim
    dev = torch.device(f"cuda:{local_rank}")
    pre, ext = [w.prefix_len] * w.batch, [w.extend_len] * w.batch
    need = sum(p + e for p, e in zip(pre, ext))
    L = args.layers or 2
    runner = synth.make_model_runner(w.hq_total, w.hkv_total, d, tp_size=tp, layer_num=L, pool_tokens=need + 64,
                                     max_reqs=w.batch, context_len=w.prefix_len + w.extend_len + 1,
                                     kv_dtype=w.kv_dtype, device=str(dev))
    fill_pool(runner.token_to_kv_pool, 1000 + rank % max(world, 1))
    backend = CudaB200AttnBackend(runner, force_path={"auto": -1, "mma": 0, "tc": 1}[args.force_path])
    gen = torch.Generator().manual_seed(args.seed)
    fb = synth.make_extend_batch(runner, pre, ext, gen, mode=args.index_mode, backend=backend)
    layers = [synth.make_layer(runner, i) for i in range(L)]
    T = sum(ext)
    gq = torch.Generator(device=dev).manual_seed(77 + rank)
    qkv = torch.randn((T, (hq + 2 * hkv) * d), generator=gq, device=dev, dtype=torch.float32).to(torch.bfloat16)
    q, k, v = qkv[:, : hq * d], qkv[:, hq * d: (hq + hkv) * d], qkv[:, (hq + hkv) * d:]
    flops = synth.extend_algorithmic_flops(pre, ext, hq, d)
    K, W = min(args.steps, 64), min(max(args.warmup, 3), 8)

    def step(i):
        if i % L == 0:
            backend.init_forward_metadata(fb)
        return layers[i % L](q, k, v, fb, save_kv_cache=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.sample()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        step(i)
    ev1.record()
    torch.cuda.synchronize()
    sampler.sample()
    sampler.stop()
    tt = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())
    # e2e: host q/k/v + metadata in, o out, every step
    h_qkv = qkv.cpu().pin_memory()
    h_o = torch.empty((T, hq * d), dtype=torch.bfloat16).pin_memory()
    d_qkv = torch.empty_like(qkv)
    dq, dk, dv_ = d_qkv[:, : hq * d], d_qkv[:, hq * d: (hq + hkv) * d], d_qkv[:, (hq + hkv) * d:]
    h_loc = fb.out_cache_loc.cpu().pin_memory()
    ke = min(K, 16)

    def e2e_step(i):
        d_qkv.copy_(h_qkv, non_blocking=True)
        fb.out_cache_loc.copy_(h_loc, non_blocking=True)
        backend.init_forward_metadata(fb)
        o = layers[i % L](dq, dk, dv_, fb, save_kv_cache=True)
        h_o.copy_(o, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(ke):
        e2e_step(i)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # K back-to-back launches of a multi-millisecond tensor kernel are the sustained regime of MEASURED_PEAKS.json
        # (cuBLAS bf16 back to back for 4 s); the burst figure (best single GEMM) is reported beside it
        burst = float(peaks.get("bf16_tflops", 1590.0))
        peak = float(peaks.get("bf16_tflops_sustained", burst))
        ach = flops / (t_max / K) / 1e12
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            tkey = f"{w.name}/{args.index_mode}/n{world}" + (f"/ps{args.page_size}" if args.page_size != 1 else "")
            traffic = tj.get(tkey, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": "extend_attn_tflops", "value": ach * world, "unit": "TFLOP/s (algorithmic, causal half counted exactly)",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_max / K * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if w.kv_dtype == torch.bfloat16 else "f16 (fp8-e4m3 pool widened on chip)", "data": "synthetic",
            "config": {"workload": f"{w.name}: {w.description}", "batch": w.batch, "prefix_len": w.prefix_len,
                       "extend_len": w.extend_len, "hq_per_gpu": hq, "hkv_per_gpu": hkv, "head_dim": d,
                       "index": args.index_mode, "layers": L, "kernel_family": args.force_path,
                       "step": "one forward_extend layer-call: fused KV store + causal extend attention over the "
                               "gathered prefix", "l2": "per-layer pools rotated; compute-bound"},
            "tokens_per_s": T * K / t_max,
            "roofline": {"bound": "tensor", "kernel": "extend_tc_kernel", "achieved": ach, "peak": peak,
                         "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                         "algorithmic_flops_per_launch": flops,
                         "peak_kind": "measured sustained cuBLAS bf16 (MEASURED_PEAKS.json bf16_tflops_sustained): the timed "
                                      "region is K back-to-back launches of this kernel",
                         "peak_burst": burst, "frac_of_burst": ach / burst, "frac_of_nominal_2250": ach / 2250.0,
                         "note": "step time used (KV store + extend kernel); the store is <1% of the step"},
            "cpu_baseline": None,
            "e2e": {"value": flops * ke / float(te.item()) / 1e12 * world, "unit": "TFLOP/s",
                    "h2d_bytes_per_step": h_qkv.numel() * 2 + h_loc.numel() * 8, "d2h_bytes_per_step": h_o.numel() * 2,
                    "steps": ke},
            "gpu_launches": K * 2, "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()

# ----------------------------------------------------------------------------
# b200 arm
# ----------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (WORLD_SIZE is 1)", file=sys.stderr)
        sys.exit(2)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the cuda_b200 path has no CPU fallback", file=sys.stderr)
        sys.exit(1)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    from harness import synth as _synth
    if _synth.WORKLOADS[args.workload].extend_len:
        run_extend_bench(args, rank, world, local_rank, dist)
        return
    st = build_state(args, rank, world * args.shard)
    w, backend, fb, layers, L = st["w"], st["backend"], st["fb"], st["layers"], st["L"]
    B, hq, hkv, d = w.batch, st["hq"], st["hkv"], st["d"]
    dev = torch.device(f"cuda:{local_rank}")
    K, W = args.steps, max(args.warmup, 3)

    gq = torch.Generator(device=dev).manual_seed(77 + rank)
    # q/k/v as column slices of the fused qkv GEMM output (llama.py:149-152)
    from harness import ForwardMode

    # Static step inputs, like the buffers of a graph runner: the activations (q/k/v as column slices of the fused qkv
    # GEMM output, llama.py:149-152) and the batch metadata the forward pass reads (seq_lens, req_pool_indices,
    # out_cache_loc) are views of ONE device buffer, so the e2e leg uploads a step's inputs with a single copy.
    qkv_bytes = B * (hq + 2 * hkv) * d * 2
    meta_off = (qkv_bytes + 255) // 256 * 256
    step_in = torch.zeros(meta_off + 3 * B * 8, dtype=torch.uint8, device=dev)
    qkv = step_in[:qkv_bytes].view(torch.bfloat16).view(B, (hq + 2 * hkv) * d)
    qkv.copy_(torch.randn((B, (hq + 2 * hkv) * d), generator=gq, device=dev, dtype=torch.float32))
    q, k, v = qkv[:, : hq * d], qkv[:, hq * d: (hq + hkv) * d], qkv[:, (hq + hkv) * d:]
    meta_dev = step_in[meta_off:].view(torch.int64).view(3, B)
    meta_dev.copy_(torch.stack([fb.seq_lens.to(torch.int64), fb.req_pool_indices.to(torch.int64),
                                fb.out_cache_loc.to(torch.int64)]))
    fb.seq_lens, fb.req_pool_indices, fb.out_cache_loc = meta_dev[0], meta_dev[1], meta_dev[2]

    launches = {"n": 0}
    # the cascade leg is launched (and exits at once when nothing is shared) wherever it is supported
    cascade_launched = (bool(backend.cascade) and (d == 128 or (d == 64 and w.kv_dtype == torch.bfloat16))
                        and B * (hq // hkv) >= 64 and B <= 1024 and 256 % (hq // hkv) == 0)
    # kv_store (or rope_kv_store), [extend_tc<shared>], decode_mma, merge; + the separate rotary kernel
    KERNELS_PER_LAYER_CALL = 3 + int(cascade_launched) + int(args.rope == "separate")
    KERNELS_PER_PLAN = 1 + 2 * int(cascade_launched)    # [group_kernel, lcp_kernel], plan_kernel

    def plan():
        backend.init_forward_metadata(fb)
        launches["n"] += KERNELS_PER_PLAN

    rotary_emb = positions = None
    if args.rope != "off":
        from scratchpad_b200.rotary_embedding import get_rope
        rotary_emb = get_rope(d, d, max(int(fb.seq_lens.max().item()) + 1, 8192), 500000.0, True,
                              {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
                               "original_max_position_embeddings": 8192}, dtype=torch.bfloat16).to(dev)
        positions = (fb.seq_lens - 1).clone()

    def run_layer(i):
        """One layer-call as the model's attention block issues it (llama.py:149-152)."""
        if args.rope == "fused":
            return layers[i](q, k, v, fb, save_kv_cache=True, rope=(rotary_emb, positions))
        if args.rope == "separate":
            rotary_emb(positions, q, k)
        return layers[i](q, k, v, fb, save_kv_cache=True)

    def layer_call(i):
        launches["n"] += KERNELS_PER_LAYER_CALL
        return run_layer(i % L)

    # --- device-resident leg -------------------------------------------------------------------
    graphs = None
    if args.mode == "graph":
        backend.init_cuda_graph_state(B)
        backend.init_forward_metadata_capture_cuda_graph(B, B, fb.req_pool_indices, fb.seq_lens, None,
                                                         ForwardMode.DECODE, None)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(min(L, 2)):
                run_layer(i)
        torch.cuda.current_stream().wait_stream(side)
        graphs, outs = [], []
        for i in range(L):  # one graph per layer-call so that any K can be timed exactly
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                outs.append(run_layer(i))
            graphs.append(g)

        def replan():
            backend.init_forward_metadata_replay_cuda_graph(B, fb.req_pool_indices, fb.seq_lens, fb.seq_lens_sum,
                                                            None, ForwardMode.DECODE, None, seq_lens_cpu=None)
            launches["n"] += KERNELS_PER_PLAN

        def step(i):
            if i % L == 0:
                replan()
            graphs[i % L].replay()
            launches["n"] += KERNELS_PER_LAYER_CALL
    else:
        def step(i):
            if i % L == 0:
                plan()
            layer_call(i)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.sample()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches["n"] = 0
    ev0.record()
    for i in range(K):
        step(i)
    ev1.record()
    torch.cuda.synchronize()
    sampler.sample()
    sampler.stop()
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    timed_launches = launches["n"]
    barrier()
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())

    # --- roofline leg: the split-KV kernel alone, CUDA events on its stream, same rotation --------
    plan()
    nk = min(K, 256)
    evs = []
    for i in range(nk + 8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); b.record()  # materialise the handles
        evs.append((a, b))
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(evs):
        backend.profile_events = (a.cuda_event, b.cuda_event)
        run_layer(i % L)
    backend.profile_events = None
    torch.cuda.synchronize()
    kt = sorted(a.elapsed_time(b) * 1e-3 for a, b in evs[8:])
    k_avg = sum(kt) / len(kt)
    k_med = kt[len(kt) // 2]

    # --- e2e leg: host buffers, H2D + forward + D2H every step ------------------------------------
    e2e = None
    if not args.no_e2e:
        h_in = step_in.cpu().pin_memory()
        h_o = torch.empty((B, hq * d), dtype=torch.bfloat16).pin_memory()
        h2d = h_in.numel()
        d2h = h_o.numel() * 2

        if graphs is not None:
            # the engine's decode path (CudaGraphRunner.replay, cuda_graph_runner.py:443-510): copy the step's inputs
            # into the static buffers the graph was captured on, refresh the metadata in place, replay, read o back
            def e2e_step(i):
                step_in.copy_(h_in, non_blocking=True)
                backend.init_forward_metadata_replay_cuda_graph(B, fb.req_pool_indices, fb.seq_lens, fb.seq_lens_sum,
                                                                None, ForwardMode.DECODE, None, seq_lens_cpu=None)
                graphs[i % L].replay()
                h_o.copy_(outs[i % L], non_blocking=True)
                torch.cuda.current_stream().synchronize()  # the caller consumes the result
        else:
            def e2e_step(i):
                step_in.copy_(h_in, non_blocking=True)
                backend.init_forward_metadata(fb)  # metadata depends on the step's seq_lens: once per call here
                o = run_layer(i % L)
                h_o.copy_(o, non_blocking=True)
                torch.cuda.current_stream().synchronize()  # the caller consumes the result

        ke = min(K, 512)
        for i in range(min(W, 8)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(ke):
            e2e_step(i)
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        tte = torch.tensor([te], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tte, op=dist.ReduceOp.MAX)
        e2e = {"value": B * ke / float(tte.item()), "unit": "tokens/s per layer-call",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": ke,
               "ms_per_step": float(tte.item()) / ke * 1e3,
               "note": "pinned host q/k/v + batch metadata -> device, metadata hook + forward_decode through the "
                       "AttentionBackend surface (" + ("CUDA-graph replay on static buffers, as the engine's decode "
                       "path does" if graphs is not None else "eager launches") + "), o -> pinned host, stream sync "
                       "every step; KV pool and req_to_token stay device-resident (engine state)"}

    # --- the untouched NCCL all-reduce after o_proj, timed only -----------------------------------
    allreduce_us = None
    if dist is not None:
        hidden = w.hq_total * w.head_dim
        x = torch.zeros((B, hidden), dtype=torch.bfloat16, device=dev)
        for _ in range(5):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            dist.all_reduce(x)
        b.record()
        torch.cuda.synchronize()
        ta = torch.tensor([a.elapsed_time(b) / 50 * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        allreduce_us = float(ta.item())

    # --- cpu baseline (rank 0, N=1 only) ------------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = min(B, 16)
        s = cpu_sample_state(w, st["seq_lens"], hq, hkv, sample, args.seed)
        dt1, _ = time_cpu(s, 1, 1)
        n = int(max(3, min(200, args.cpu_seconds / max(dt1, 1e-6))))
        dt, done = time_cpu(s, n, 1, max_seconds=3 * args.cpu_seconds)
        cpu_baseline = {"value": sample / dt, "unit": "tokens/s per layer-call", "cores": torch.get_num_threads(),
                        "kind": "port",
                        "sample": f"oracle torch-native port (gather + fp32 SDPA), {sample} of {B} requests of "
                                  f"{w.name}, {done} layer-calls, {dt * 1e3:.2f} ms each"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs, copy benchmark)" if "hbm_gbs" in peaks else "fallback 6650"
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            tkey = f"{w.name}/{args.index_mode}/n{world}" + (f"/ps{args.page_size}" if args.page_size != 1 else "")
            traffic = tj.get(tkey, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        achieved = st["algo_bytes"] / k_avg / 1e9
        shared_len = int(backend.workspace[:64].view(torch.int32)[7].item())
        cfg_out = workload_config(w, world * args.shard, args, st)
        cfg_out["rope"] = args.rope
        cfg_out["tile_fetch"] = {0: "LDGSTS row gather", 1: "TMA, one box per K/V tile of consecutive slots (row gather "
                                 "for the rest)"}[backend.lib.b200_debug_last_tma_mode()]
        cfg_out["cascade"] = {"enabled": bool(backend.cascade), "launched": cascade_launched,
                              "shared_prefix_tokens_detected": shared_len,
                              "row_tiles": int(backend.workspace[:64].view(torch.int32)[8].item())}
        roof_kernel = "decode_paged_kernel" if backend.lib.b200_debug_last_tma_mode() else "decode_mma_kernel"
        roof_note = None
        if shared_len > 0:
            roof_kernel = "extend_tc_kernel<shared> + decode_mma_kernel"
            roof_note = ("cascade active: the batch-shared prefix is read once per kv head by the dense tcgen05 leg, so "
                         "GB/s on the logical (per-request) bytes exceeds the HBM peak by design; flops of the dense "
                         f"leg per launch = {4 * d * hq * B * shared_len}")
        line = {
            "metric": "decode_attn_tokens_per_s", "value": B * K / t_max, "unit": "tokens/s per layer-call",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_max / K * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if w.kv_dtype == torch.bfloat16 else "fp8_e4m3", "data": "synthetic",
            "config": cfg_out,
            "hbm_gbs_step": st["algo_bytes"] * world / (t_max / K) / 1e9,
            "hbm_frac_of_8tbs_per_gpu": st["algo_bytes"] / (t_max / K) / 8e12,
            "tokens_per_s_model": B * K / t_max / w.layers_total,
            "roofline": {"bound": "hbm", "kernel": roof_kernel, "note": roof_note, "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": st["algo_bytes"], "kernel_us_avg": k_avg * 1e6,
                         "kernel_us_median": k_med * 1e6, "launches_timed": len(kt), "peak_kind": peak_kind,
                         "frac_of_8tbs": achieved / 8000.0},
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": timed_launches,
            "clocks": sampler.summary(),
        }
        if allreduce_us is not None:
            line["allreduce_after_o_proj_us"] = allreduce_us
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
