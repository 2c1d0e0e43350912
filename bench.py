#!/usr/bin/env python
"""bench.py — headline benchmark of the plugin hook-chain hot path on B200.

Metric (BASELINE.json): tool-call payloads/sec on batched 16 KiB JSON payloads through the fused
regex_filter / deny_filter / harmful_content_detector scan (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W          # our arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                   # the reference's CPU path (CPython `re`) on host cores

A "step" is one pass of the hot path over one batch: `units` payloads of ~16 KiB each, packed as
unit 0xFF unit 0xFF ... (include/cfgpu.h).  Per-GPU work is fixed (weak scaling).  Rank 0 prints ONE
JSON line.  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAYLOAD_BYTES = 16384
UNITS_PER_GPU = 65536          # 65536 x 16 KiB = 1 GiB per step per GPU (8x the 126 MB L2; SURVEY §8d "64k batch")
DISTINCT = 256                 # distinct seeded payloads, tiled to the batch
MIX = (("A", 0.5), ("B", 0.25), ("C", 0.25))
METRIC = "tool-call payloads/sec (16 KiB JSON), fused regex/deny/harmful scan"

HARMFUL = None
DENY = ["innovative", "groundbreaking", "revolutionary"]          # plugins/config.yaml:171-174
SUBS = [("crap", 0, "crud"), ("crud", 0, "yikes")]                # plugins/config.yaml:149-153


def make_payloads(distinct: int = DISTINCT):
    from mcp_context_forge_b200 import synth

    out = []
    for i in range(distinct):
        r = (i * 0.61803398875) % 1.0
        acc = 0.0
        shape = "A"
        for s, w in MIX:
            acc += w
            if r < acc:
                shape = s
                break
        tgt = PAYLOAD_BYTES if shape != "B" else int(PAYLOAD_BYTES * 0.6)
        p = synth.payload(shape, tgt, seed=i, hit_rate=1e-4)
        b = p.encode("utf-8")
        if len(b) > PAYLOAD_BYTES + 512:   # keep units within ~3 % of 16 KiB
            p = b[: PAYLOAD_BYTES].decode("utf-8", "ignore")
        out.append(p)
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (CPython `re`, the reference's own matcher) on all host cores
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init():
    from oracle import hook_chain_ref as ref

    _W["cats"] = ref.harmful_compile(None)
    _W["rules"] = ref.regex_compile_rules([{"search": s, "replace": r} for s, _, r in SUBS])
    _W["ref"] = ref


def _cpu_chain(payloads):
    """The reference chain's data path for one payload (string unit): harmful scan (9 IGNORECASE
    searches), deny (3 substring tests), regex_filter (2 subs).  Returns a checksum."""
    ref = _W["ref"]
    n = 0
    for p in payloads:
        n += len(ref.harmful_scan_text(p, _W["cats"]))
        n += 1 if any(w in p for w in DENY) else 0
        n += len(ref.regex_apply_str(_W["rules"], p))
    return n


def cpu_run(payloads, total_units: int, cores: int) -> float:
    """payloads/s of the oracle chain over `total_units` payloads spread over `cores` processes."""
    import multiprocessing as mp

    per = max(1, total_units // cores)
    work = [[payloads[(c * per + i) % len(payloads)] for i in range(per)] for c in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        pool.map(_cpu_chain, [w[:2] for w in work])          # warm the workers
        t0 = time.perf_counter()
        pool.map(_cpu_chain, work)
        dt = time.perf_counter() - t0
    return per * cores / dt


def host_cores() -> int:
    """Usable host threads: affinity mask, capped by the cgroup CPU quota when one is set."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def side_stages(ctx, engine, payloads):
    """Informational device-resident timings of the other rows of the path on the same payload mix
    (not part of the headline metric): toon_encoder, the JSON structural index and request_logging_masking,
    4096 units each."""
    import ctypes

    import torch

    out = {}
    try:
        n = 4096
        units = [payloads[i % len(payloads)].encode("utf-8") for i in range(n)]
        stream, offs = engine.pack_units(units)
        batch = engine.Batch(ctx, len(stream), n)
        batch.upload(stream, offs)
        d_out = torch.empty(len(stream) + 16, dtype=torch.uint8, device="cuda")
        d_len = torch.empty(n, dtype=torch.int32, device="cuda")
        d_st = torch.empty(n, dtype=torch.int32, device="cuda")
        lib = ctx.lib
        for name in ("toon",):
            lib.cf_toon(ctx.h, batch.h, 0, d_out.data_ptr(), d_len.data_ptr(), d_st.data_ptr(), None)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.cf_toon(ctx.h, batch.h, 0, d_out.data_ptr(), d_len.data_ptr(), d_st.data_ptr(), None)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            out[name] = {"units": n, "ms": ms, "payloads_per_s": n / ms * 1e3, "gb_per_s": len(stream) / ms / 1e6, "converted": int((d_st == 0).sum())}
        toks = torch.empty((len(stream) + 64, 2), dtype=torch.int32, device="cuda")
        for _ in range(2):
            lib.cf_json_index(ctx.h, batch.h, 0, toks.data_ptr(), d_len.data_ptr(), None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.cf_json_index(ctx.h, batch.h, 0, toks.data_ptr(), d_len.data_ptr(), None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out["json_index_stage1"] = {"units": n, "ms": ms, "gb_per_s": len(stream) / ms / 1e6, "tokens": int((d_len & 0x7FFFFFFF).sum())}
        del toks
        engine.mask_host(batch, stream, offs, 10)          # warm-up: scratch buffers grow once
        t0 = time.perf_counter()
        st, _ = engine.mask_host(batch, stream, offs, 10)
        dt = time.perf_counter() - t0
        out["mask_e2e_host_buffers"] = {"units": n, "ms": dt * 1e3, "payloads_per_s": n / dt, "ok": int((st == 0).sum())}
    except Exception as exc:  # informational only
        out["error"] = str(exc)
    return out


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--units", type=int, default=UNITS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": f"configs[1]: batched 16 KiB payloads ({int(MIX[0][1]*100)}% tabular JSON / {int(MIX[1][1]*100)}% nested JSON / {int(MIX[2][1]*100)}% prose, hit rate 1e-4), "
                          "fused harmful(9 IGNORECASE regex)+deny(3 literals)+regex_filter(2 rules) scan",
              "payload_bytes": PAYLOAD_BYTES, "units_per_gpu": args.units, "batch_bytes_per_gpu": None,
              "patterns": 14, "l2_policy": "inputs_larger_than_l2 (1 GiB batch per GPU vs 126 MB L2)",
              "parallelism": f"shard{world}: independent payload shards per GPU, one NCCL all_gather of verdict bitmaps" if world > 1 else "single GPU"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        payloads = make_payloads(64)
        cores = host_cores()
        sample = max(cores * 8, min(args.units, 64 * cores))
        vals = []
        for i in range(args.warmup + args.steps):
            v = cpu_run(payloads, sample, cores)
            if i >= args.warmup:
                vals.append(v)
            if i == 0 and sample / v > 8.0:      # keep the whole run within a few minutes
                sample = max(cores * 4, int(sample * 4.0 / (sample / v)))
        val = sum(vals) / len(vals)
        config["batch_bytes_per_gpu"] = sample * PAYLOAD_BYTES
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "payloads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sample / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "payloads/s", "cores": cores, "kind": "port",
                                 "sample": f"{sample} payloads of 16 KiB per step over {cores} processes; oracle/hook_chain_ref.py = the reference plugins' loops on CPython re (the reference is pure Python; cpex/orjson absent here)"},
                "e2e": {"value": val, "unit": "payloads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    payloads = make_payloads()
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        sample = 96 * cores                               # ~3 ms per payload per core: ~10-40 s of CPU work in total
        v = cpu_run(payloads, sample, cores)
        cpu_base = {"value": v, "unit": "payloads/s", "cores": cores, "kind": "port",
                    "sample": f"{sample} payloads of 16 KiB (same mix) over {cores} processes, oracle/hook_chain_ref.py chain (CPython re)"}

    import numpy as np
    import torch

    from mcp_context_forge_b200 import engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = engine.Context.get(local_rank)
    prog = engine.Program()
    from mcp_context_forge_b200.plugins.harmful_content_detector import DEFAULT_LEXICONS   # product's copy of the reference defaults

    for pats in DEFAULT_LEXICONS.values():
        for pat in pats:
            prog.add_search(pat, re.I)
    for w in DENY:
        prog.add_literal(w)
    for s, f, r in SUBS:
        prog.add_sub(s, f, r)
    prog.compile(ctx)
    W = prog.words

    n = args.units
    units = [payloads[(i + rank * 7) % len(payloads)] for i in range(n)]
    stream, offs = engine.pack_units(units)
    nbytes = len(stream)
    config["batch_bytes_per_gpu"] = nbytes
    # pinned host copies (torch is plumbing: allocator / streams / NCCL)
    h_stream = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_stream.numpy()[:] = np.frombuffer(stream, dtype=np.uint8)
    h_offs = torch.from_numpy(offs.astype(np.int64)).pin_memory()
    h_bm = torch.empty(n * W, dtype=torch.int64, pin_memory=True)
    d_bm = torch.zeros(n * W, dtype=torch.int64, device="cuda")
    # N > 1: verdict bitmaps are all-gathered every step; two buffer pairs so that the gather of step k
    # (NCCL stream) overlaps the scan of step k+1 (compute stream)
    d_bms = [d_bm, torch.zeros(n * W, dtype=torch.int64, device="cuda")] if world > 1 else [d_bm]
    d_alls = [torch.zeros(world * n * W, dtype=torch.int64, device="cuda") for _ in range(2)] if world > 1 else None
    pending = []
    step_no = [0]
    batch = engine.Batch(ctx, nbytes, n)
    lib = ctx.lib
    cs = torch.cuda.current_stream().cuda_stream

    def upload():
        ctx.check(lib.cf_batch_upload(ctx.h, batch.h, h_stream.data_ptr(), nbytes, h_offs.data_ptr(), n, cs), "upload")

    def step_resident():
        k = step_no[0] & 1 if world > 1 else 0
        step_no[0] += 1
        if world > 1 and len(pending) >= 2:
            pending.pop(0).wait()                  # buffer pair k is free again
        ctx.check(lib.cf_scan(ctx.h, prog.h, batch.h, d_bms[k].data_ptr(), cs), "cf_scan")
        if world > 1:
            pending.append(dist.all_gather_into_tensor(d_alls[k], d_bms[k], async_op=True))

    def drain():
        while pending:
            pending.pop(0).wait()

    def step_e2e():
        ctx.check(lib.cf_scan_host(ctx.h, prog.h, batch.h, h_stream.data_ptr(), nbytes, h_offs.data_ptr(), n, h_bm.data_ptr()), "cf_scan_host")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, use_events=True):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        if world > 1:
            drain()                                # every step's gather is inside the timed region
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1) if use_events else wall
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    upload()
    for _ in range(args.warmup):
        step_resident()
    if world > 1:
        drain()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = ctx.kernel_launches
    ctx.check(lib.cf_profile_begin(ctx.h, args.steps), "profile_begin")
    ms_total = timed(step_resident, args.steps)
    import ctypes

    kms, kn = ctypes.c_double(), ctypes.c_uint32()
    ctx.check(lib.cf_profile_collect(ctx.h, ctypes.byref(kms), ctypes.byref(kn)), "profile_collect")
    ctx.check(lib.cf_profile_begin(ctx.h, 0), "profile_end")
    launches = ctx.kernel_launches - l0
    clocks = sampler.stop() if sampler else None
    cand, steps_dfa = ctx.scan_counters()

    # end to end through the C ABI with host (pinned) buffers: H2D + kernels + D2H each step
    for _ in range(2):
        step_e2e()
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = timed(step_e2e, e2e_steps, use_events=False)   # cf_scan_host is synchronous: wall clock

    # sanity: the e2e verdicts equal the resident ones, and flagged units are what the oracle says on a sample
    torch.cuda.synchronize()
    same = bool((h_bm.cuda() == d_bms[0]).all().item()) and (world == 1 or bool((d_alls[0][rank * n * W:(rank + 1) * n * W] == d_bms[0]).all().item()))
    stages = None
    if rank == 0 and world == 1:
        stages = side_stages(ctx, engine, payloads)
    if rank == 0:
        from oracle import hook_chain_ref as ref          # checker only: parity of a sample of this run's verdicts

        sample_units = units[:8]
        exp = ref.scan_bitmaps(sample_units, [(p, re.I) for pats in ref.DEFAULT_LEXICONS.values() for p in pats], DENY, [(s, f) for s, f, _ in SUBS])
        got = engine.bitmaps_to_ints(h_bm.numpy().view(np.uint64), 8, W)
        if got != exp or not same:
            raise SystemExit(f"bench.py: parity check failed (resident==e2e: {same})")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    ms_step = ms_total / args.steps
    value = world * n / (ms_step / 1e3)
    e2e_val = world * n / (ms_e2e / e2e_steps / 1e3)
    peak, peak_src = read_peaks()
    k_ms = kms.value / max(1, kn.value)
    alg_bytes = nbytes + 8 * W * n          # stream read once + verdict words
    achieved = alg_bytes / (k_ms / 1e3) / 1e9 if k_ms > 0 else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "scan_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("stream_bytes") == nbytes:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    line = {
        "metric": METRIC, "value": value, "unit": "payloads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": f"synthetic: {DISTINCT} distinct seeded payloads tiled to {n} units per GPU", "config": config,
        "e2e": {"value": e2e_val, "unit": "payloads/s", "h2d_bytes_per_step": world * (nbytes + 8 * (n + 1)), "d2h_bytes_per_step": world * 8 * W * n,
                "api": "cf_scan_host (C ABI, pinned host buffers, synchronous)", "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "kernel": "scan_kernel", "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
        "cpu_baseline": cpu_base,
        "clocks": clocks,
        "scan_counters": {"prefilter_candidates": cand, "dfa_steps": steps_dfa},
        "other_stages": stages,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
