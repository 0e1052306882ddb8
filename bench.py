#!/usr/bin/env python
"""bench.py -- decode-attention tokens/s and achieved HBM GB/s of the cuda_b200 backend.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C3B]

A *step* is one layer-call of the hot path over one synthetic batch:
``set_kv_buffer`` (fused KV store) + split-KV paged decode + LSE merge, i.e. one
``AttentionBackend.forward_decode`` as the engine issues it per layer
(scratchpad/nn/attention/triton_backend.py:157-196 in the reference), with
``init_forward_metadata`` once per model step (every ``layers`` steps, as
model_runner.py:506-514 does).  Steps rotate over ``layers`` distinct per-layer
KV pools (>> L2), so every step streams its KV from HBM.  The device-resident leg
replays CUDA graphs that hold one model step's worth of layer-calls each, the way
the engine's graph runner captures a whole forward pass
(cuda_graph_runner.py:394-423).

Headline workload = the shape BASELINE.json's ``north_star`` sets its target on:
Llama-3.1-8B decode, batch 256, 4k context, bf16 KV, page size 1, random-permuted
slot ids (C3B).  The other BASELINE configs ride along in ``suite`` (C2, the
shared-prefix C3 and its fp8 control C3U at every N; the 70B TP8 shard C5 and the
extend chunk C4 at N=1).  At N GPUs the kv heads are sharded N ways (the engine's
tensor-parallel layout); every rank serves the same batch with Hkv/N heads, no
collective on the path.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import copy
import gc
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

HEADLINE = "C3B"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--index-mode", default="random", choices=["random", "contiguous"])
    ap.add_argument("--layers", type=int, default=0, help="distinct per-layer KV pools rotated (0 = auto)")
    ap.add_argument("--mode", default="graph", choices=["graph", "graph1", "eager"],
                    help="device-resident leg: graph = CUDA graphs of one model step (`layers` layer-calls) each, as the "
                         "engine captures a whole forward pass; graph1 = one graph per layer-call; eager = plain launches")
    ap.add_argument("--cascade", default="auto", choices=["auto", "on", "off"],
                    help="shared-prefix cascade decode (server flag b200_shared_prefix_cascade; the backend's own default "
                         "is on).  auto = what an operator would set for the deployment the workload stands for: on when "
                         "its requests share a prompt prefix (C3), off otherwise -- an idle cascade costs one early-exit "
                         "launch (~2 us) per layer-call")
    ap.add_argument("--force-path", default="auto", choices=["auto", "mma", "tc", "tc2"],
                    help="extend workloads: kernel family (mma = legacy mma.sync kernel, tc = tcgen05/TMEM kernel, one CTA "
                         "per tile, tc2 = the opt-in CTA-pair tcgen05 kernel)")
    ap.add_argument("--rope", default="off", choices=["off", "separate", "fused"],
                    help="decode only: include the model's rotary embedding in the layer-call -- 'separate' = the "
                         "reference's sequence rotary_emb(positions, q, k) then attention (llama.py:149-152), 'fused' = "
                         "one b200_rope_kv_store launch doing the rotation and the KV write (SURVEY 8f #1)")
    ap.add_argument("--page-size", type=int, default=1, choices=[-1, 0, 1, 16, 32, 64, 128],
                    help="decode only: KV pool page size.  1 = the reference's token-granular pool (default, the headline "
                         "configuration) with the backend's default tile fetch; 0 = the same pool, row gather forced "
                         "(tma_contiguous_runs off); -1 = the same pool, TMA for consecutive-slot tiles forced on; "
                         "> 1 = paged pool, random pages, whole tiles fetched by TMA (SURVEY 8f #4)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--kv-layout", default="auto", choices=["auto", "split", "fused"],
                    help="split: the reference's separate k/v buffers; fused: K|V rows of a token adjacent "
                         "(pool option of this backend); auto: fused when a per-buffer token row is < 256 B")
    ap.add_argument("--shard", type=int, default=1,
                    help="development aid: run the per-GPU shard shapes of a (gpus*shard)-way head split on this GPU")
    ap.add_argument("--suite", default="auto", choices=["auto", "on", "off"],
                    help="append the other BASELINE configs as compact records (auto: on for the default headline run)")
    ap.add_argument("--comparators", default="auto", choices=["auto", "on", "off"],
                    help="time the reference's own GPU path (unmodified FlashInferAttnBackend + in-tree Triton kernels from "
                         "baseline/_ref) on the same tensors, N=1 only (auto: with the suite)")
    ap.add_argument("--no-parity", action="store_true", help="skip the check of the timed output")
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
K_SCALE, V_SCALE = 6.0 / 448.0, 5.0 / 448.0  # fp8 pools: per-tensor scales, the reference's semantics


def fill_pool(pool, seed: int):
    """bf16-valued N(0,1) KV in every layer (chunked); fp8 pools get fixed per-tensor scales (equal per-head
    entries, so the reference's FlashInfer path -- one float per tensor -- computes the same thing)."""
    dev = pool.device
    g = torch.Generator(device=dev).manual_seed(seed)
    fp8 = pool.dtype != torch.bfloat16
    for layer_id in range(pool.layer_num):
        if fp8:
            pool.k_scale[layer_id].fill_(K_SCALE)
            pool.v_scale[layer_id].fill_(V_SCALE)
        for which, buf in (("k", pool.k_buffer[layer_id]), ("v", pool.v_buffer[layer_id])):
            n_tok = buf.shape[0]
            chunk = max(1, (1 << 24) // (buf.shape[1] * buf.shape[2]))
            for s in range(0, n_tok, chunk):
                n = min(chunk, n_tok - s)
                x = torch.randn((n, buf.shape[1], buf.shape[2]), generator=g, device=dev, dtype=torch.float32)
                if not fp8:
                    buf[s:s + n] = x.to(torch.bfloat16)
                else:
                    sc = K_SCALE if which == "k" else V_SCALE
                    y = (x.to(torch.bfloat16).float() / sc).clamp_(-448.0, 448.0).to(torch.bfloat16)
                    buf[s:s + n] = y.to(pool.dtype).view(torch.uint8)
            buf[0].zero_()


def pick_layers(per_layer_bytes: int, want: int, total_layers: int) -> int:
    if want > 0:
        return want
    budget = 40e9  # keep the rotated pools well inside 180 GB and the build fast
    return int(max(2, min(total_layers, 16, budget // max(per_layer_bytes, 1))))


def static_info(args, w, tp):
    """Everything about a decode workload that does not need a GPU: the reference arm prints the same `config`."""
    from harness import synth
    from scratchpad_b200.sharding import head_shard

    gen = torch.Generator().manual_seed(args.seed)  # same page table on every rank (TP replicas)
    seq_lens = synth.workload_seq_lens(w, gen)
    sh = head_shard(w.hq_total, w.hkv_total, tp, 0)
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
    row_bytes = hkv_l * w.head_dim * kv_elt
    fused = args.kv_layout == "fused" or (args.kv_layout == "auto" and row_bytes < 256)
    return dict(gen=gen, seq_lens=seq_lens, hq=hq_l, hkv=hkv_l, need=need, pool_tokens=pool_tokens, ps=ps,
                kv_elt=kv_elt, per_layer_bytes=per_layer, L=layers, row_bytes=row_bytes, fused_kv=fused)


def workload_config(w, tp, args, info):
    """Declarative description of the workload (identical in the b200 and the reference arm)."""
    return {
        "workload": f"{w.name}: {w.description}", "batch": w.batch, "hq_per_gpu": info["hq"],
        "hkv_per_gpu": info["hkv"], "head_dim": w.head_dim, "page_size": max(args.page_size, 1),
        "kv_dtype": "bf16" if w.kv_dtype == torch.bfloat16 else "fp8_e4m3", "index": args.index_mode,
        "parallelism": f"kv-head sharded x{tp} (tensor-parallel head layout), no collective on the path",
        "step": "one forward_decode layer-call: fused KV store + split-KV paged decode + LSE merge; "
                "init_forward_metadata once per `layers` steps",
        "layers": info["L"], "sum_seq_len": int(sum(info["seq_lens"])),
        "kv_layout": ("fused [P+1,2,Hkv,D] (K|V rows adjacent)" if info["fused_kv"] else
                      "split k_buffer/v_buffer [P+1,Hkv,D] (reference layout)"),
        "token_row_bytes_per_buffer": info["row_bytes"],
        "l2": f"inputs larger than L2: {info['L']} per-layer KV pools of {info['per_layer_bytes'] / 1e9:.2f} GB "
              "rotated step to step",
        "mode": args.mode, "rope": args.rope,
    }


def build_state(args, rank, tp):
    from scratchpad_b200 import CudaB200AttnBackend
    from harness import synth

    w = synth.WORKLOADS[args.workload]
    assert w.hkv_total % tp == 0 and w.hq_total % tp == 0, "kv heads must divide over the GPUs"
    info = static_info(args, w, tp)
    gen, seq_lens, layers = info["gen"], info["seq_lens"], info["L"]
    dev = f"cuda:{torch.cuda.current_device()}"
    runner = synth.make_model_runner(w.hq_total, w.hkv_total, w.head_dim, tp_size=tp, layer_num=layers,
                                     pool_tokens=info["pool_tokens"], max_reqs=w.batch, context_len=max(seq_lens) + 1,
                                     kv_dtype=w.kv_dtype, device=dev, fused_kv=info["fused_kv"], page_size=info["ps"])
    fill_pool(runner.token_to_kv_pool, 1000 + rank % max(tp, 1))
    cascade = args.cascade == "on" or (args.cascade == "auto" and w.shared_prefix > 0)
    tma_runs = {-1: True, 0: False}.get(args.page_size)  # None = the backend's default
    backend = CudaB200AttnBackend(runner, cascade=cascade, tma_contiguous_runs=tma_runs)
    fb = synth.make_decode_batch(runner, seq_lens, gen, mode=args.index_mode, shared_prefix=w.shared_prefix,
                                 idx_dtype=torch.int64, backend=backend)
    lay = [synth.make_layer(runner, i) for i in range(layers)]
    algo_bytes = synth.decode_algorithmic_bytes(seq_lens, info["hq"], info["hkv"], w.head_dim, w.head_dim, info["kv_elt"])
    return dict(info, w=w, runner=runner, backend=backend, fb=fb, layers=lay, d=w.head_dim, algo_bytes=algo_bytes)


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
        ks = torch.full((hkv,), K_SCALE)
        vs = torch.full((hkv,), V_SCALE)
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


def all_host_threads() -> int:
    # torchrun exports OMP_NUM_THREADS=1 to its workers, so the count is set explicitly
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_baseline_c1(seconds: float):
    """BASELINE.json configs[0] (C1: Llama-3.2-1B, batch 8, 1k context, bf16, page 1 -- the reference's own
    CPU-runnable case), the WHOLE config, at every host thread and at 1 thread (the engine pins
    torch.set_num_threads(1), model_runner.py:189)."""
    from harness import synth

    w = synth.WORKLOADS["C1"]
    gen = torch.Generator().manual_seed(1)
    lens = synth.workload_seq_lens(w, gen)
    s = cpu_sample_state(w, lens, w.hq_total, w.hkv_total, w.batch, 1)
    out = {"workload": f"{w.name}: {w.description}", "unit": "tokens/s per layer-call", "kind": "port"}
    for label, nthr in (("all_threads", all_host_threads()), ("one_thread", 1)):
        torch.set_num_threads(nthr)
        dt1, _ = time_cpu(s, 1, 1)
        n = int(max(3, min(400, seconds / max(dt1, 1e-6))))
        dt, done = time_cpu(s, n, 1, max_seconds=2 * seconds)
        out[label] = {"value": w.batch / dt, "cores": torch.get_num_threads(), "ms_per_step": dt * 1e3, "steps": done}
    all_host_threads()
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The fork ships no
    torch-native backend and its Triton/FlashInfer backends need a GPU, so this
    times the oracle port ("port") with every host thread, on a bounded sample of the same workload.
    (The reference's GPU path on the same box is the `comparators` record of the b200 arm.)"""
    if rank != 0:
        return
    from harness import synth

    w = synth.WORKLOADS[args.workload]
    tp = world * args.shard
    info = static_info(args, w, tp)
    # the CPU arm has no GPUs to shard over: it always computes every head of the workload (the job the N GPUs share)
    hq, hkv = w.hq_total, w.hkv_total
    cores = all_host_threads()
    sample = min(w.batch, 16)
    s = cpu_sample_state(w, info["seq_lens"], hq, hkv, sample, args.seed)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        cpu_step(s)
    dt, done = time_cpu(s, args.steps, 0)
    val = sample / dt
    desc = f"{sample} of {w.batch} requests of {w.name} per step (same shapes and context lengths), all heads, {done} steps"
    line = {
        "impl": "reference", "metric": "decode_attn_tokens_per_s", "value": val,
        "unit": "tokens/s per layer-call", "n_gpus": world, "steps": done, "warmup": warm,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if w.kv_dtype == torch.bfloat16 else "fp8_e4m3", "data": "synthetic",
        "config": workload_config(w, tp, args, info),
        "cpu_baseline": {"value": val, "unit": "tokens/s per layer-call", "cores": cores, "kind": "port",
                         "sample": desc},
        "e2e": {"value": val, "unit": "tokens/s per layer-call", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def traffic_for(key: str):
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return tj.get(key, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def free_cuda():
    gc.collect()
    torch.cuda.empty_cache()


# ----------------------------------------------------------------------------
# the reference's own GPU path on the same tensors (baseline/comparators.py): reported, never on the product path
# ----------------------------------------------------------------------------
def decode_comparators(st, q3, iters: int):
    """FlashInfer through the reference's unmodified FlashInferAttnBackend (flashinfer_backend.py:448-484,623-670) and
    the reference's in-tree Triton decode kernel (decode_attention.py:547-608), eager launches, attention only
    (save_kv_cache=False), rotating over the same per-layer pools as the b200 arm.  `ours_attn_us` is the b200
    backend measured the same way (eager forward_decode, save_kv_cache=False)."""
    out = {}
    try:
        from baseline import comparators as C, ref_loader
        if not ref_loader.available():
            return {"unavailable": "baseline/_ref is not installed (python baseline/install_ref.py)"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:200]}
    w, runner, backend, fb, L = st["w"], st["runner"], st["backend"], st["fb"], st["L"]
    pool = runner.token_to_kv_pool
    B, hq, d = w.batch, st["hq"], st["d"]
    fp8 = w.kv_dtype != torch.bfloat16
    q2 = q3.reshape(B, hq * d)
    backend.init_forward_metadata(fb)
    ours_o = [None]

    def ours(i):
        ours_o[0] = st["layers"][i % L](q2, None, None, fb, save_kv_cache=False)

    out["ours_attn_us"] = C.time_cuda(ours, iters) * 1e6
    ours_last = ours_o[0].float()
    last_layer = (iters - 1) % L
    sm_scale = float(st["layers"][0].scaling)
    try:
        rs = C.RefState(req_to_token=runner.req_to_token_pool.req_to_token, k_buffers=pool.k_buffer,
                        v_buffers=pool.v_buffer, kv_dtype=w.kv_dtype, hq=hq, hkv=st["hkv"], head_dim=d,
                        k_scale=K_SCALE if fp8 else None, v_scale=V_SCALE if fp8 else None, sm_scale=sm_scale)
        ref = C.RefFlashInfer(rs, context_len=runner.model_config.context_len)
        rfb = rs.decode_batch(fb.req_pool_indices, fb.seq_lens, fb.out_cache_loc, ref.backend)
        ref.init_forward_metadata(rfb)
        torch.cuda.synchronize()
        plan_t = []
        for _ in range(5):
            t0 = time.perf_counter()
            ref.init_forward_metadata(rfb)  # kv_indices flatten + host-side plan (flashinfer_backend.py:623-670)
            torch.cuda.synchronize()
            plan_t.append(time.perf_counter() - t0)
        out["flashinfer_plan_us"] = sorted(plan_t)[len(plan_t) // 2] * 1e6
        theirs_o = [None]

        def theirs(i):
            theirs_o[0] = ref.forward(i % L, q2, None, None, rfb, save_kv_cache=False)

        out["flashinfer_us"] = C.time_cuda(theirs, iters) * 1e6
        out["flashinfer_tensor_cores"] = ref.use_tensor_cores
        diff = (ours_last - theirs_o[0].float()).abs().max().item()
        out["flashinfer_max_abs_vs_ours"] = diff
        del ref, rs, rfb
    except Exception as e:  # noqa: BLE001
        out["flashinfer_error"] = repr(e)[:300]
    if not fp8 and not st["fused_kv"]:
        try:
            r2t = runner.req_to_token_pool.req_to_token
            total = int(fb.seq_lens.sum().item())
            logits = torch.empty((hq, total), dtype=torch.float32, device=q3.device)
            theirs_o = [None]

            def tri(i):
                theirs_o[0] = C.triton_decode(q3, pool.k_buffer[i % L], pool.v_buffer[i % L], r2t, fb.req_pool_indices,
                                              fb.seq_lens, sm_scale, 0.0, attn_logits=logits)

            out["triton_us"] = C.time_cuda(tri, iters) * 1e6
            out["triton_max_abs_vs_ours"] = (ours_last - theirs_o[0].reshape(B, -1).float()).abs().max().item()
            del logits
        except Exception as e:  # noqa: BLE001
            out["triton_error"] = repr(e)[:300]
    else:
        out["triton_note"] = ("the reference's Triton backend has no fp8 scale semantics (SURVEY section 0 fact 4); "
                              "its bf16 figure is on the bf16 twin of this shape")
    page = int(getattr(pool, "page_size", 1))
    if page >= 16 and not fp8 and not st["fused_kv"]:
        # SURVEY 8d stretch comparator: FlashInfer's trtllm-gen decode (NVIDIA's Blackwell-native FMHA cubins) on the
        # same paged pool bytes -- not a path of the reference engine (page size 1), the best library kernel there is
        try:
            tg = C.TrtllmGenDecode(runner.req_to_token_pool.req_to_token, fb.req_pool_indices, fb.seq_lens, page)
            theirs_o = [None]

            def trt(i):
                theirs_o[0] = tg(q3, pool.k_buffer[i % L], pool.v_buffer[i % L], sm_scale)

            out["trtllm_gen_us"] = C.time_cuda(trt, iters) * 1e6
            out["trtllm_gen_max_abs_vs_ours"] = (ours_last - theirs_o[0].reshape(B, -1).float()).abs().max().item()
            out["speedup_vs_trtllm_gen"] = out["trtllm_gen_us"] / out["ours_attn_us"]
            del tg
        except Exception as e:  # noqa: BLE001
            out["trtllm_gen_error"] = repr(e)[:300]
    if "flashinfer_us" in out:
        out["speedup_vs_flashinfer"] = out["flashinfer_us"] / out["ours_attn_us"]
    if "triton_us" in out:
        out["speedup_vs_triton"] = out["triton_us"] / out["ours_attn_us"]
    out["how"] = ("eager forward_decode(save_kv_cache=False) per layer-call, CUDA events, per-layer pools rotated; "
                  "flashinfer_plan_us = the reference's init_forward_metadata (once per model step, host-synchronous)")
    free_cuda()
    return out


# ----------------------------------------------------------------------------
# parity check of the timed output
# ----------------------------------------------------------------------------
def decode_parity(st, q3, outs_by_layer, n_oracle_rows: int = 8):
    """`outs_by_layer`: {layer index: o [B, Hq*Dv]} as the timed loop left them.  Every row is compared with the
    library's independent one-warp-per-(request, head) fp32 kernel on the same device buffers, and `n_oracle_rows`
    rows of one layer with the CPU oracle (oracle/attn_oracle.py) on a cut-out of the pool."""
    from harness import checks
    from oracle import attn_oracle as O
    from scratchpad_b200 import _lib

    w, runner, fb = st["w"], st["runner"], st["fb"]
    pool = runner.token_to_kv_pool
    B, hq, d = w.batch, st["hq"], st["d"]
    fp8 = w.kv_dtype != torch.bfloat16
    abs_tol, rel_tol = (5e-2, 5e-2) if fp8 else (2e-2, 1e-2)
    r2t = runner.req_to_token_pool.req_to_token
    sm_scale = float(st["layers"][0].scaling)
    res = {"tolerance": {"max_abs": abs_tol, "rel": rel_tol}, "layers_checked": sorted(outs_by_layer)}
    worst_abs = worst_rel = 0.0
    finite = elementwise = True
    for li, o in outs_by_layer.items():
        ks, vs = (pool.get_kv_scales(li) if fp8 else (None, None))
        naive = checks.naive_decode(q3, pool.k_buffer[li], pool.v_buffer[li], r2t, fb.req_pool_indices, fb.seq_lens,
                                    sm_scale, 0.0, _lib.KV_FP8_E4M3 if fp8 else _lib.KV_BF16, ks, vs)
        got = o.view(B, hq, d).float()
        ref = naive.float()
        finite = finite and bool(torch.isfinite(got).all().item())
        worst_abs = max(worst_abs, (got - ref).abs().max().item())
        worst_rel = max(worst_rel, ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item())
        # element-wise as well as in norm: |got - ref| <= abs_tol + rel_tol * |ref| for every output element
        elementwise = elementwise and bool(((got - ref).abs() <= abs_tol + rel_tol * ref.abs()).all().item())
    res["vs_naive_kernel"] = {"rows": B, "max_abs": worst_abs, "rel": worst_rel, "rel_kind": "Frobenius",
                              "elementwise_ok": elementwise}
    li = sorted(outs_by_layer)[-1]
    rows = sorted({int(x) for x in torch.linspace(0, B - 1, n_oracle_rows).round().tolist()})
    ks, vs = (pool.get_kv_scales(li) if fp8 else (None, None))
    cq, ck, cv, ct, crpi, csl = checks.sample_rows_to_cpu(rows, q3, pool.k_buffer[li], pool.v_buffer[li], r2t,
                                                          fb.req_pool_indices, fb.seq_lens)
    want = O.decode_ref(cq, ck, cv, ct, crpi, csl, sm_scale, 0.0, w.kv_dtype if fp8 else None,
                        None if ks is None else ks.cpu(), None if vs is None else vs.cpu())
    got = outs_by_layer[li].view(B, hq, d)[torch.tensor(rows, device=q3.device)].float().cpu()
    oa = (got - want).abs().max().item()
    orl = ((got - want).norm() / want.norm().clamp_min(1e-12)).item()
    res["vs_cpu_oracle"] = {"rows": rows, "layer": li, "max_abs": oa, "rel": orl}
    res["ok"] = bool(finite and elementwise and worst_abs <= abs_tol and worst_rel <= rel_tol and oa <= abs_tol
                     and orl <= rel_tol)
    return res


# ----------------------------------------------------------------------------
# extend / chunked-prefill workloads (C4): tensor-bound, metric TFLOP/s
# ----------------------------------------------------------------------------
def measure_extend(args, rank, world, local_rank, dist, K, W, want_e2e=True, want_parity=True, want_comparators=False):
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
    backend = CudaB200AttnBackend(runner, force_path={"auto": -1, "mma": 0, "tc": 1, "tc2": 2}[args.force_path])
    gen = torch.Generator().manual_seed(args.seed)
    fb = synth.make_extend_batch(runner, pre, ext, gen, mode=args.index_mode, backend=backend)
    layers = [synth.make_layer(runner, i) for i in range(L)]
    T = sum(ext)
    gq = torch.Generator(device=dev).manual_seed(77 + rank)
    qkv = torch.randn((T, (hq + 2 * hkv) * d), generator=gq, device=dev, dtype=torch.float32).to(torch.bfloat16)
    q, k, v = qkv[:, : hq * d], qkv[:, hq * d: (hq + hkv) * d], qkv[:, (hq + hkv) * d:]
    flops = synth.extend_algorithmic_flops(pre, ext, hq, d)
    K, W = min(K, 64), min(max(W, 3), 8)
    last = {}

    def step(i):
        if i % L == 0:
            backend.init_forward_metadata(fb)
        last["o"], last["layer"] = layers[i % L](q, k, v, fb, save_kv_cache=True), i % L

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
    torch.cuda._sleep(2_000_000)  # device time only: the first launches are queued before the device reaches ev0
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

    backend.lib.b200_debug_extend_tma_tiles(1)  # debug counter: K / V tiles fetched by one TMA box copy each
    step(0)
    tma_tiles = int(backend.lib.b200_debug_extend_tma_tiles(1))
    total_tiles = 2 * hkv * sum(((e + 256 // (hq // hkv) - 1) // (256 // (hq // hkv))) and
                                sum((p_ + min(e, (qb + 1) * (256 // (hq // hkv))) + 127) // 128
                                    for qb in range((e + 256 // (hq // hkv) - 1) // (256 // (hq // hkv))))
                                for p_, e in zip(pre, ext))
    parity = None
    if want_parity and rank == 0:
        # the timed output of the last step against the CPU oracle, on sampled query rows of every request (the
        # oracle is exact per query row: prefix rows + causal new rows), oracle/attn_oracle.py:extend_ref
        try:
            parity = extend_parity(runner, fb, last["o"], last["layer"], q, k, v, pre, ext, hq, hkv, d, w.kv_dtype,
                                   float(layers[0].scaling))
        except Exception as e:  # noqa: BLE001
            parity = {"ok": False, "error": repr(e)[:300]}

    comparators = None
    if want_comparators and rank == 0 and world == 1:
        comparators = extend_comparators(runner, fb, layers, q, k, v, hq, hkv, d, w.kv_dtype, L, backend,
                                         iters=min(K, 8))

    e2e = None
    if want_e2e:
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
        e2e = {"value": flops * ke / float(te.item()) / 1e12 * world, "unit": "TFLOP/s",
               "h2d_bytes_per_step": h_qkv.numel() * 2 + h_loc.numel() * 8, "d2h_bytes_per_step": h_o.numel() * 2,
               "steps": ke}
    peaks = load_peaks()
    # K back-to-back launches of a multi-millisecond tensor kernel are the sustained regime of MEASURED_PEAKS.json
    # (cuBLAS bf16 back to back for 4 s); the burst figure (best single GEMM) is reported beside it
    burst = float(peaks.get("bf16_tflops", 1590.0))
    peak = float(peaks.get("bf16_tflops_sustained", burst))
    ach = flops / (t_max / K) / 1e12
    return dict(
        w=w, K=K, W=W, t_max=t_max, T=T, flops=flops, ach=ach, peak=peak, burst=burst, e2e=e2e, parity=parity,
        comparators=comparators, clocks=sampler.summary(), hq=hq, hkv=hkv, d=d, L=L, tma_tiles=tma_tiles,
        total_tiles=total_tiles,
        traffic=traffic_for(f"{w.name}/{args.index_mode}/n{world}"))


def extend_parity(runner, fb, o, layer, q, k, v, pre, ext, hq, hkv, d, kv_dtype, sm_scale, rows_per_req: int = 6):
    """Sampled query rows of every request against the fp32 CPU oracle.  Row t of request i sees the request's prefix
    slots in the pool (dequantised for fp8 pools) and the new tokens 0..t from the call arguments: restated as a
    one-request, one-row extend problem -- cut-out fp32 pool [prefix rows | new tokens 0..t-1], new token t -- so the
    oracle's own extend_ref is used unchanged."""
    from oracle import attn_oracle as O

    fp8 = kv_dtype != torch.bfloat16
    abs_tol, rel_tol = (5e-2, 5e-2) if fp8 else (2e-2, 1e-2)
    pool = runner.token_to_kv_pool
    r2t = runner.req_to_token_pool.req_to_token
    kb, vb = pool.k_buffer[layer], pool.v_buffer[layer]
    ks, vs = (pool.get_kv_scales(layer) if fp8 else (None, None))
    T = q.shape[0]
    q3, k3, v3, o3 = q.reshape(T, hq, d), k.reshape(T, hkv, d), v.reshape(T, hkv, d), o.view(T, hq, d)
    worst_abs, num, den, n_rows = 0.0, 0.0, 0.0, 0
    start = 0
    zero = torch.zeros((1, hkv, d), dtype=torch.float32)
    one = torch.zeros(1, dtype=torch.int64)
    for i, (p, e) in enumerate(zip(pre, ext)):
        req = int(fb.req_pool_indices[i])
        slots = r2t[req, :p].long()
        kp = O.dequantize_kv_ref(kb[slots].cpu(), kv_dtype, None if ks is None else ks.cpu())
        vp = O.dequantize_kv_ref(vb[slots].cpu(), kv_dtype, None if vs is None else vs.cpu())
        kn, vn = k3[start:start + e].float().cpu(), v3[start:start + e].float().cpu()
        qn = q3[start:start + e].cpu()
        picks = sorted({0, e - 1, e // 2, min(e - 1, 127), min(e - 1, 128), (7 * e) // 8})[:rows_per_req]
        for t in picks:
            ck = torch.cat([zero, kp, kn[:t]])
            cv = torch.cat([zero, vp, vn[:t]])
            tab = torch.zeros((1, p + t + 4), dtype=torch.int32)
            tab[0, :p + t] = torch.arange(1, p + t + 1, dtype=torch.int32)
            want = O.extend_ref(qn[t:t + 1], kn[t:t + 1], vn[t:t + 1], ck, cv, tab, one, torch.tensor([p + t + 1]),
                                torch.tensor([1], dtype=torch.int32), torch.zeros(1, dtype=torch.int32), sm_scale, 0.0)[0]
            got = o3[start + t].float().cpu()
            worst_abs = max(worst_abs, (got - want).abs().max().item())
            num += float((got - want).pow(2).sum())
            den += float(want.pow(2).sum())
            n_rows += 1
        start += e
    rel = (num / max(den, 1e-30)) ** 0.5
    return {"tolerance": {"max_abs": abs_tol, "rel": rel_tol}, "vs_cpu_oracle": {"query_rows": n_rows, "layer": layer,
            "max_abs": worst_abs, "rel": rel}, "ok": bool(worst_abs <= abs_tol and rel <= rel_tol)}


def extend_comparators(runner, fb, layers, q, k, v, hq, hkv, d, kv_dtype, L, backend, iters: int):
    """The reference's FlashInfer extend (ragged causal + paged prefix + merge_state, flashinfer_backend.py:378-446) and
    its in-tree Triton extend kernel (extend_attention.py:229-327) on the same tensors, attention only."""
    out = {}
    try:
        from baseline import comparators as C, ref_loader
        if not ref_loader.available():
            return {"unavailable": "baseline/_ref is not installed (python baseline/install_ref.py)"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:200]}
    pool = runner.token_to_kv_pool
    fp8 = kv_dtype != torch.bfloat16
    T = q.shape[0]
    qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
    backend.init_forward_metadata(fb)
    ours_o = [None]

    def ours(i):
        ours_o[0] = layers[i % L](qc, kc, vc, fb, save_kv_cache=False)

    out["ours_attn_us"] = C.time_cuda(ours, iters, warmup=2) * 1e6
    ours_last = ours_o[0].float()
    sm_scale = float(layers[0].scaling)
    try:
        rs = C.RefState(req_to_token=runner.req_to_token_pool.req_to_token, k_buffers=pool.k_buffer,
                        v_buffers=pool.v_buffer, kv_dtype=kv_dtype, hq=hq, hkv=hkv, head_dim=d,
                        k_scale=1.0 if fp8 else None, v_scale=1.0 if fp8 else None, sm_scale=sm_scale)
        ref = C.RefFlashInfer(rs, context_len=runner.model_config.context_len)
        el = fb.extend_seq_lens.to(torch.int64)
        sl = fb.seq_lens.to(torch.int64)
        rfb = rs.extend_batch(fb.req_pool_indices, sl, fb.out_cache_loc, el, sl - el, ref.backend)
        t0 = time.perf_counter()
        ref.init_forward_metadata(rfb)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref.init_forward_metadata(rfb)
        torch.cuda.synchronize()
        out["flashinfer_plan_us"] = (time.perf_counter() - t0) * 1e6
        theirs_o = [None]

        def theirs(i):
            theirs_o[0] = ref.forward(i % L, qc, kc, vc, rfb, save_kv_cache=False)

        out["flashinfer_us"] = C.time_cuda(theirs, iters, warmup=2) * 1e6
        if not fp8:
            out["flashinfer_max_abs_vs_ours"] = (ours_last - theirs_o[0].float()).abs().max().item()
        del ref, rs, rfb
    except Exception as e:  # noqa: BLE001
        out["flashinfer_error"] = repr(e)[:300]
    if not fp8:
        try:
            r2t = runner.req_to_token_pool.req_to_token
            theirs_o = [None]
            q3, k3, v3 = qc.view(T, hq, d), kc.view(T, hkv, d), vc.view(T, hkv, d)

            def tri(i):
                theirs_o[0] = C.triton_extend(q3, k3, v3, pool.k_buffer[i % L], pool.v_buffer[i % L], r2t,
                                              fb.req_pool_indices, fb.seq_lens, fb.extend_seq_lens, fb.extend_start_loc,
                                              sm_scale)

            out["triton_us"] = C.time_cuda(tri, max(2, iters // 2), warmup=1) * 1e6
            out["triton_max_abs_vs_ours"] = (ours_last - theirs_o[0].reshape(T, -1).float()).abs().max().item()
        except Exception as e:  # noqa: BLE001
            out["triton_error"] = repr(e)[:300]
    if "flashinfer_us" in out:
        out["speedup_vs_flashinfer"] = out["flashinfer_us"] / out["ours_attn_us"]
    if "triton_us" in out:
        out["speedup_vs_triton"] = out["triton_us"] / out["ours_attn_us"]
    out["how"] = "eager forward_extend(save_kv_cache=False) per layer-call, CUDA events, contiguous q/k/v"
    free_cuda()
    return out


def extend_line(r, args, world):
    w = r["w"]
    K, t_max, ach, peak, burst = r["K"], r["t_max"], r["ach"], r["peak"], r["burst"]
    return {
        "metric": "extend_attn_tflops", "value": ach * world, "unit": "TFLOP/s (algorithmic, causal half counted exactly)",
        "n_gpus": world, "steps": K, "warmup": r["W"], "ms_per_step": t_max / K * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if w.kv_dtype == torch.bfloat16 else "f16 (fp8-e4m3 pool widened on chip)", "data": "synthetic",
        "config": {"workload": f"{w.name}: {w.description}", "batch": w.batch, "prefix_len": w.prefix_len,
                   "extend_len": w.extend_len, "hq_per_gpu": r["hq"], "hkv_per_gpu": r["hkv"], "head_dim": r["d"],
                   "index": args.index_mode, "layers": r["L"], "kernel_family": args.force_path,
                   "step": "one forward_extend layer-call: fused KV store + causal extend attention over the "
                           "gathered prefix", "l2": "per-layer pools rotated; compute-bound"},
        "tokens_per_s": r["T"] * K / t_max,
        "roofline": {"bound": "tensor", "kernel": "extend_tc_kernel", "achieved": ach, "peak": peak,
                     "unit": "TFLOP/s", "frac": ach / peak, "traffic": r["traffic"],
                     "algorithmic_flops_per_launch": r["flops"],
                     "peak_kind": "measured sustained cuBLAS bf16 (MEASURED_PEAKS.json bf16_tflops_sustained): the timed "
                                  "region is K back-to-back launches of this kernel",
                     "peak_burst": burst, "frac_of_burst": ach / burst, "frac_of_nominal_2250": ach / 2250.0,
                     "note": "step time used (KV store + extend kernel); the store is <1% of the step"},
        "runtime": {"kv_tiles_per_launch": r["total_tiles"], "kv_tiles_fetched_by_tma_per_launch": r["tma_tiles"],
                    "note": "tcgen05 kernel: a K / V tile of 128 consecutive rows (new tokens; prefix runs) is one TMA box copy"},
        "cpu_baseline": None, "e2e": r["e2e"], "parity_check": r["parity"], "comparators": r["comparators"],
        "gpu_launches": K * 2, "clocks": r["clocks"],
    }


# ----------------------------------------------------------------------------
# decode workloads
# ----------------------------------------------------------------------------
def measure_decode(args, rank, world, local_rank, dist, K, W, want_e2e=True, want_parity=True, want_comparators=False,
                   roofline_launches=256):
    """All legs of one decode workload on this rank; collective-free except the max-over-ranks of the timings."""
    from harness import ForwardMode

    tp = world * args.shard
    st = build_state(args, rank, tp)
    w, backend, fb, layers, L = st["w"], st["backend"], st["fb"], st["layers"], st["L"]
    B, hq, hkv, d = w.batch, st["hq"], st["hkv"], st["d"]
    dev = torch.device(f"cuda:{local_rank}")
    gq = torch.Generator(device=dev).manual_seed(77 + rank)

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
    cascade_launched = (bool(backend.cascade) and d in (64, 128)
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

    # --- device-resident leg -------------------------------------------------------------------
    last_outs = {}  # layer index -> o of the most recent layer-call on that layer (what the parity check reads)
    graphs = {}

    def capture(first: int, n: int):
        """One CUDA graph holding layer-calls first .. first+n-1 (n == L: a whole model step)."""
        key = (first, n)
        if key not in graphs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                outs = [run_layer((first + j) % L) for j in range(n)]
            graphs[key] = (g, outs)
        return graphs[key]

    if args.mode != "eager":
        backend.init_cuda_graph_state(B)
        backend.init_forward_metadata_capture_cuda_graph(B, B, fb.req_pool_indices, fb.seq_lens, None,
                                                         ForwardMode.DECODE, None)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(min(L, 2)):
                run_layer(i)
        torch.cuda.current_stream().wait_stream(side)
        per = L if args.mode == "graph" else 1

        def replan():
            backend.init_forward_metadata_replay_cuda_graph(B, fb.req_pool_indices, fb.seq_lens, fb.seq_lens_sum,
                                                            None, ForwardMode.DECODE, None, seq_lens_cpu=None)
            launches["n"] += KERNELS_PER_PLAN

        def segments(n):
            """n layer-calls as model steps: the metadata refresh, then graphs of <= `per` layer-calls."""
            i = 0
            while i < n:
                m = min(L - i % L, n - i)  # the rest of this model step
                yield i % L, m
                i += m

        for n in (W, K):
            for first, m in segments(n):
                j = 0
                while j < m:
                    capture((first + j) % L, min(per, m - j))
                    j += per

        # the first replay of a CUDA graph also uploads it (~100 us): every graph the timed loop will use is replayed
        # once here, outside the timed region (at K = 20, L = 16 the remainder graph would otherwise meet its first
        # replay inside it)
        replan()
        for g_, _ in list(graphs.values()):
            g_.replay()
        torch.cuda.synchronize()

        def run_steps(n):
            for first, m in segments(n):
                replan()
                j = 0
                while j < m:
                    c = min(per, m - j)
                    g, outs = graphs[((first + j) % L, c)]
                    g.replay()
                    for x, o in enumerate(outs):
                        last_outs[(first + j + x) % L] = o
                    j += per
                launches["n"] += KERNELS_PER_LAYER_CALL * m
    else:
        def run_steps(n):
            for i in range(n):
                if i % L == 0:
                    plan()
                last_outs[i % L] = run_layer(i % L)
                launches["n"] += KERNELS_PER_LAYER_CALL

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(W)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.sample()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches["n"] = 0
    last_outs.clear()
    # The timed region is device time: a ~1 ms delay kernel is queued in front of the first event so that the host
    # has enqueued the metadata refresh and the first graph before the device reaches it.  Without it the host-side
    # launch latency of the first model step (~150 us with eight ranks sharing the host cores) sits inside the region
    # and weighs 8 % of 20 steps of 92 us (N = 8) against 0.1 % of 20 steps of 646 us (N = 1); the host-inclusive
    # number is the e2e leg's job.
    torch.cuda._sleep(2_000_000)
    ev0.record()
    run_steps(K)
    ev1.record()
    torch.cuda.synchronize()
    sampler.sample()
    sampler.stop()
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    timed_launches = launches["n"]
    barrier()
    tt = torch.tensor([t_dev], dtype=torch.float64, device=dev)
    per_rank_us = None
    if dist is not None:
        allt = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allt, tt)  # every rank's own device time: the job's time is the slowest rank's
        per_rank_us = [float(x.item()) / K * 1e6 for x in allt]
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())

    # --- parity of the timed output (before anything else touches the static buffers) -----------------
    parity = None
    if want_parity and rank == 0 and args.rope == "off":
        try:
            keys = sorted(last_outs)
            pick = {keys[0]: last_outs[keys[0]], keys[-1]: last_outs[keys[-1]]}
            parity = decode_parity(st, q.view(B, hq, d), pick)
        except Exception as e:  # noqa: BLE001
            parity = {"ok": False, "error": repr(e)[:300]}

    # --- roofline leg: the split-KV kernel alone, CUDA events on its stream, same rotation --------
    plan()
    nk = max(8, min(K, roofline_launches))
    evs = []
    for i in range(nk + 8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); b.record()  # materialise the handles
        evs.append((a, b))
    torch.cuda.synchronize()
    backend.lib.b200_debug_tma_tiles(1)  # reset the debug counter of TMA-fetched KV tiles
    for i, (a, b) in enumerate(evs):
        backend.profile_events = (a.cuda_event, b.cuda_event)
        run_layer(i % L)
    backend.profile_events = None
    torch.cuda.synchronize()
    tma_tiles = int(backend.lib.b200_debug_tma_tiles(1)) / len(evs)
    kt = sorted(a.elapsed_time(b) * 1e-3 for a, b in evs[8:])
    k_avg = sum(kt) / len(kt)
    k_med = kt[len(kt) // 2]

    # --- e2e leg: host buffers, H2D + forward + D2H every step ------------------------------------
    e2e = None
    if want_e2e:
        h_in = step_in.cpu().pin_memory()
        h_o = torch.empty((B, hq * d), dtype=torch.bfloat16).pin_memory()
        h2d = h_in.numel()
        d2h = h_o.numel() * 2

        if args.mode != "eager":
            # the engine's decode path (CudaGraphRunner.replay, cuda_graph_runner.py:443-510): copy the step's inputs
            # into the static buffers the graph was captured on, refresh the metadata in place, replay, read o back.
            # A step is ONE layer-call, so these are single-layer graphs (the engine uploads once per model step; the
            # contract asks for the copies every step)
            Le = min(L, 4)
            for i in range(Le):
                capture(i, 1)

            def e2e_step(i):
                step_in.copy_(h_in, non_blocking=True)
                backend.init_forward_metadata_replay_cuda_graph(B, fb.req_pool_indices, fb.seq_lens, fb.seq_lens_sum,
                                                                None, ForwardMode.DECODE, None, seq_lens_cpu=None)
                g, outs = graphs[(i % Le, 1)]
                g.replay()
                h_o.copy_(outs[0], non_blocking=True)
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
                       "path does" if args.mode != "eager" else "eager launches") + "), o -> pinned host, stream sync "
                       "every step; KV pool and req_to_token stay device-resident (engine state)"}

    comparators = None
    if want_comparators and rank == 0 and world == 1 and args.shard == 1:
        graphs.clear()
        free_cuda()
        comparators = decode_comparators(st, q.view(B, hq, d), iters=max(8, min(K, 32)))

    hdr = backend.workspace[:64].view(torch.int32)
    shared_len, row_tiles = int(hdr[7].item()), int(hdr[8].item())
    tma_mode = backend.lib.b200_debug_last_tma_mode()
    res = dict(st=None, w=w, info={k_: st[k_] for k_ in ("seq_lens", "hq", "hkv", "L", "row_bytes", "fused_kv",
                                                           "per_layer_bytes", "kv_elt")},
               K=K, W=max(W, 0), t_max=t_max, k_avg=k_avg, k_med=k_med, n_kernel=len(kt), e2e=e2e, parity=parity,
               comparators=comparators, clocks=sampler.summary(), algo_bytes=st["algo_bytes"],
               timed_launches=timed_launches, shared_len=shared_len, row_tiles=row_tiles, tma_mode=tma_mode,
               per_rank_us=per_rank_us, tma_tiles=tma_tiles, total_tiles=int(hdr[0].item()) * (hkv // int(backend.config.heads_per_cta)),
               cascade_enabled=bool(backend.cascade), cascade_launched=cascade_launched, tp=tp)
    graphs.clear()
    last_outs.clear()
    del st, backend, fb, layers, step_in, qkv, q, k, v
    free_cuda()
    return res


def decode_roofline(r, args, world):
    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs, copy benchmark)" if "hbm_gbs" in peaks else "fallback 6650"
    w = r["w"]
    tkey = f"{w.name}/{args.index_mode}/n{r['tp']}" + (f"/ps{args.page_size}" if args.page_size not in (0, 1) else "")
    achieved = r["algo_bytes"] / r["k_avg"] / 1e9
    kernel = "decode_paged_kernel" if r["tma_mode"] else "decode_mma_kernel"
    note = None
    if r["shared_len"] > 0:
        kernel = "extend_tc_kernel<shared> + decode_mma_kernel"
        note = ("cascade active: the batch-shared prefix is read once per kv head by the dense tcgen05 leg, so "
                "GB/s on the logical (per-request) bytes exceeds the HBM peak by design; flops of the dense "
                f"leg per launch = {4 * w.head_dim * r['info']['hq'] * w.batch * r['shared_len']}")
    return {"bound": "hbm", "kernel": kernel, "note": note, "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_for(tkey),
            "algorithmic_bytes_per_launch": r["algo_bytes"], "kernel_us_avg": r["k_avg"] * 1e6,
            "kernel_us_median": r["k_med"] * 1e6, "launches_timed": r["n_kernel"], "peak_kind": peak_kind,
            "frac_of_8tbs": achieved / 8000.0}


def decode_runtime(r):
    return {"tile_fetch": {0: "LDGSTS row gather (TMA disabled or pages smaller than the tile)",
                           1: "TMA-capable kernel: one box per K/V tile whose slots are consecutive, row gather for "
                              "the rest; CTAs that start on scattered slots run the pure row-gather loop"}[r["tma_mode"]],
            "kv_tiles_per_launch": r["total_tiles"], "kv_tiles_fetched_by_tma_per_launch": r["tma_tiles"],
            "cascade": {"enabled": r["cascade_enabled"], "launched": r["cascade_launched"],
                        "shared_prefix_tokens_detected": r["shared_len"], "row_tiles": r["row_tiles"]}}


def decode_line(r, args, world):
    w, K, t_max = r["w"], r["K"], r["t_max"]
    B = w.batch
    info = dict(r["info"])
    cfg = workload_config(w, r["tp"], args, info)
    return {
        "metric": "decode_attn_tokens_per_s", "value": B * K / t_max, "unit": "tokens/s per layer-call",
        "n_gpus": world, "steps": K, "warmup": r["W"], "ms_per_step": t_max / K * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if w.kv_dtype == torch.bfloat16 else "fp8_e4m3", "data": "synthetic",
        "config": cfg, "runtime": decode_runtime(r),
        "hbm_gbs_step": r["algo_bytes"] * world / (t_max / K) / 1e9,
        "hbm_frac_of_8tbs_per_gpu": r["algo_bytes"] / (t_max / K) / 8e12,
        "tokens_per_s_model": B * K / t_max / w.layers_total,
        "roofline": decode_roofline(r, args, world),
        "e2e": r["e2e"], "parity_check": r["parity"], "comparators": r["comparators"],
        "gpu_launches": r["timed_launches"], "clocks": r["clocks"],
        **({"step_us_per_rank": r["per_rank_us"]} if r.get("per_rank_us") else {}),
    }


def compact(line):
    """A suite entry: the numbers of one full line without the prose."""
    out = {"workload": line["config"]["workload"].split(":")[0], "metric": line["metric"], "value": line["value"],
           "unit": line["unit"], "us_per_step": line["ms_per_step"] * 1e3, "steps": line["steps"],
           "dtype": line["dtype"]}
    rf = line.get("roofline") or {}
    out["roofline"] = {k: rf.get(k) for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "frac_of_8tbs",
                                              "frac_of_burst", "kernel_us_avg", "traffic") if k in rf}
    if "hbm_frac_of_8tbs_per_gpu" in line:
        out["step_frac_of_8tbs"] = line["hbm_frac_of_8tbs_per_gpu"]
    if line.get("runtime"):
        out["runtime"] = line["runtime"]
    for k in ("hq_per_gpu", "hkv_per_gpu", "kv_layout", "index", "page_size"):
        if k in line["config"]:
            out[k] = line["config"][k]
    if line.get("e2e"):
        out["e2e"] = {k: line["e2e"][k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")}
    pc = line.get("parity_check")
    if pc is not None:
        out["parity_check"] = pc
    if line.get("comparators"):
        out["comparators"] = line["comparators"]
    if "step_us_per_rank" in line:
        out["step_us_per_rank"] = line["step_us_per_rank"]
    out["clocks"] = {k: (line.get("clocks") or {}).get(k) for k in ("sm_mhz", "reasons")}
    return out


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

    from harness import synth
    default_run = (args.workload == HEADLINE and args.shard == 1 and args.page_size == 1 and args.rope == "off"
                   and args.index_mode == "random" and args.mode == "graph")
    want_suite = args.suite == "on" or (args.suite == "auto" and default_run)
    want_cmp = args.comparators == "on" or (args.comparators == "auto" and want_suite)
    K, W = args.steps, max(args.warmup, 3)
    dev = torch.device(f"cuda:{local_rank}")

    if synth.WORKLOADS[args.workload].extend_len:
        r = measure_extend(args, rank, world, local_rank, dist, K, W, want_e2e=not args.no_e2e,
                           want_parity=not args.no_parity, want_comparators=want_cmp)
        if rank == 0:
            print(json.dumps(extend_line(r, args, world)), flush=True)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    r = measure_decode(args, rank, world, local_rank, dist, K, W, want_e2e=not args.no_e2e,
                       want_parity=not args.no_parity, want_comparators=want_cmp)
    line = decode_line(r, args, world) if rank == 0 else None

    # --- the untouched NCCL all-reduce after o_proj, timed only -----------------------------------
    if dist is not None:
        w = r["w"]
        x = torch.zeros((w.batch, w.hq_total * w.head_dim), dtype=torch.bfloat16, device=dev)
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
        if rank == 0:
            line["allreduce_after_o_proj_us"] = float(ta.item())
        del x

    # --- suite: the other BASELINE configs, compact ------------------------------------------------
    if want_suite:
        suite = []
        # at every N (kv heads sharded like the headline): C2, and BASELINE configs[2] -- the shared-prefix decode C3 with
        # its unshared control C3U
        plan = [("C2", {}), ("C3U", {}), ("C3", {})]
        if world == 1:
            # C4 in both slot orders SURVEY 8d asks for: random-permuted (worst case) and contiguous runs (what an
            # unfragmented allocator hands out: prefix tiles then arrive by TMA like the new tokens do)
            plan += [("C5", {"shard": 8}), ("C2", {"shard": 8}), ("C3B", {"shard": 8}), ("C4", {}),
                     ("C4", {"index_mode": "contiguous"}),
                     # the headline shape over 64-token pages: whole tiles by TMA, and FlashInfer's trtllm-gen decode
                     # (which needs pages) as a comparator on the same bytes
                     ("C3B", {"page_size": 64})]
        for name, over in plan:
            a2 = copy.copy(args)
            a2.workload = name
            a2.layers = 0
            for k_, v_ in over.items():
                setattr(a2, k_, v_)
            wl = synth.WORKLOADS[name]
            try:
                if wl.extend_len:
                    rr = measure_extend(a2, rank, world, local_rank, dist, 16, 3, want_e2e=False,
                                        want_parity=not args.no_parity,
                                        want_comparators=want_cmp and a2.index_mode == "random")
                    entry = compact(extend_line(rr, a2, world)) if rank == 0 else None
                else:
                    if wl.hkv_total % (world * a2.shard):
                        continue
                    cmp_here = want_cmp and a2.shard == 1 and (name in ("C2", "C3U") or a2.page_size > 1)
                    rr = measure_decode(a2, rank, world, local_rank, dist, 64, 16, want_e2e=(a2.shard == 1),
                                        want_parity=not args.no_parity, want_comparators=cmp_here,
                                        roofline_launches=64)
                    entry = compact(decode_line(rr, a2, world)) if rank == 0 else None
                    if entry is not None and a2.shard > 1:
                        entry["workload"] += f" (per-GPU shard of a {a2.shard}-way kv-head split, run on one GPU)"
                del rr
            except Exception as e:  # noqa: BLE001  (a suite entry must never take the headline down)
                entry = {"workload": name, "error": repr(e)[:300]}
            free_cuda()
            if rank == 0 and entry is not None:
                suite.append(entry)
        if rank == 0:
            line["suite"] = suite

    # --- cpu baseline (rank 0, N=1 only) ------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        w = r["w"]
        B = w.batch
        all_host_threads()
        sample = min(B, 16)
        s = cpu_sample_state(w, r["info"]["seq_lens"], r["info"]["hq"], r["info"]["hkv"], sample, args.seed)
        dt1, _ = time_cpu(s, 1, 1)
        n = int(max(3, min(200, args.cpu_seconds / max(dt1, 1e-6))))
        dt, done = time_cpu(s, n, 1, max_seconds=3 * args.cpu_seconds)
        line["cpu_baseline"] = {"value": sample / dt, "unit": "tokens/s per layer-call", "cores": torch.get_num_threads(),
                                "kind": "port",
                                "sample": f"oracle torch-native port (gather + fp32 SDPA), {sample} of {B} requests of "
                                          f"{w.name}, {done} layer-calls, {dt * 1e3:.2f} ms each"}
        if want_suite:
            line["cpu_baseline_c1"] = cpu_baseline_c1(min(args.cpu_seconds, 4.0))
    elif rank == 0:
        line["cpu_baseline"] = None

    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
