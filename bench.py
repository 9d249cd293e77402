#!/usr/bin/env python3
"""bench.py — ops merged/sec of the Peritext op-log apply + flatten hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5] [--docs D] [--impl engine|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path (pt_batch_merge: apply every op of every log + flatten to spans + digest) over one
batch of synthetic logs.  Default workload = BASELINE.json configs[1] ("c2": 1K docs x 10K ops, insert/delete only,
2 replicas => 2000 logs, 2x10^7 op records, 320 MB of packed input — larger than the 126 MB L2, so consecutive steps
re-read their input from HBM).  `value` counts op records applied per second summed over docs AND replicas (each
replica really applies every op in the reference, src/micromerge.ts:513), inputs resident in HBM; `e2e` is the same
metric through the public API with pinned-host inputs (H2D) and results read back (D2H) inside the timed region.
N > 1: documents are sharded by doc id, one process per GPU, same per-GPU work ("weak"); the only collective is one
all-gather of the 32-byte per-log result headers (digests) per step for the convergence check.

--impl reference: the reference's sequential algorithm (C++ restatement in oracle/, Node.js is unavailable in this
image) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ops merged/sec across batch"
UNIT = "ops/s"


def sample_clocks(stop_evt, out, device_index):
    """Polls the B200_PROFILING.md clocks line (one nvidia-smi query per sample, back to back) while the bench runs."""
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    while not stop_evt.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(device_index)],
                               capture_output=True, text=True, timeout=5)
            if r.returncode == 0 and r.stdout.strip():
                parts = [x.strip() for x in r.stdout.strip().split(",")]
                if len(parts) >= 9:
                    out.append(parts)
        except Exception:
            pass
        stop_evt.wait(0.02)


def clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[1]) for s in samples if s[1].replace(".", "").isdigit())
    mx = max((float(s[2]) for s in samples if s[2].replace(".", "").isdigit()), default=None)
    reasons = set()
    for s in samples:
        for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[5:9]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(samples)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(config, n_docs):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json),
    valid for the same config and docs-per-GPU; None otherwise."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        e = json.load(open(p)).get(config)
        if e and int(e.get("docs_per_gpu", -1)) == int(n_docs):
            return float(e["dram_bytes_per_launch"]), e.get("source")
    except Exception:
        pass
    return None, None


def cpu_baseline(batch, budget_s=20.0):
    """Times the oracle (sequential reference algorithm) on a bounded sample of the same workload, all host cores."""
    from oracle.packed import replay_packed
    cores = os.cpu_count() or 1
    R = batch.meta["replicas"]
    n_logs = batch.n_logs
    # grow the sample until the run takes long enough to be meaningful but stays bounded
    sample = min(n_logs, max(R, cores * R))
    ops_s, info = 0.0, ""
    t_spent = 0.0
    while True:
        sub = batch.select(range(sample))
        _, dt = replay_packed(sub, threads=cores, flatten=True)
        t_spent += dt
        ops_s = sub.n_ops / dt
        info = f"first {sample} of {n_logs} logs ({sub.n_ops} op records), {dt:.2f} s, apply+flatten, one log per thread"
        if dt >= budget_s / 4 or sample >= n_logs or t_spent > budget_s:
            break
        grow = max(2.0, min(8.0, (budget_s / 2) / max(dt, 1e-3)))
        sample = min(n_logs, int(sample * grow) // R * R)
    return ops_s, cores, info


def run_reference(args, rank, world):
    """--impl reference arm: rank 0 only."""
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build()
    from peritext_b200 import workload
    cfg = workload.CONFIGS[args.config]
    n_docs = args.docs or cfg["n_docs"]
    cores = os.cpu_count() or 1
    # bounded sample: the reference is O(N^2) per document, so a step replays a slice of the workload
    sample_docs = min(n_docs, max(cores, 16) if args.config in ("c2", "c3") else max(cores * 64, 1024) if args.config == "c4" else 1)
    batch = workload.generate(args.config, n_docs=sample_docs, ops_per_doc=args.ops_per_doc)
    from oracle.packed import replay_packed
    for _ in range(max(0, min(args.warmup, 1))):
        replay_packed(batch, threads=cores)
    t = 0.0
    for _ in range(args.steps):
        _, dt = replay_packed(batch, threads=cores)
        t += dt
    ms = 1e3 * t / max(1, args.steps)
    value = batch.n_ops / (ms / 1e3)
    sample = (f"{sample_docs} of {n_docs} docs x {batch.meta['replicas']} replicas per step ({batch.n_ops} op records); C++ restatement "
              f"of the reference algorithm (Node.js unavailable in image), one log per thread")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfg['label']}", "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--docs", type=int, default=0, help="documents per GPU (default: the config's)")
    ap.add_argument("--ops-per-doc", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "engine" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; this engine has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as g
    g.build()
    from peritext_b200 import workload
    from peritext_b200.engine import BatchEngine
    from peritext_b200.packing import RESULT_DT

    cfg = workload.CONFIGS[args.config]
    n_docs = args.docs or cfg["n_docs"]
    t0 = time.time()
    batch = workload.generate(args.config, n_docs=n_docs, ops_per_doc=args.ops_per_doc, doc_first=rank * n_docs)
    gen_s = time.time() - t0
    R = batch.meta["replicas"]
    n_logs = batch.n_logs
    ops_per_step_local = batch.n_ops                      # op records applied (docs x replicas)
    in_bytes = batch.insdel.nbytes + batch.marks.nbytes + batch.desc.nbytes

    # pinned host copies of the inputs (the e2e leg uploads from these every step)
    def pinned(a):
        t = torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory() if a.nbytes else torch.zeros(16, dtype=torch.uint8).pin_memory()
        return t
    p_ins, p_mk = pinned(batch.insdel), pinned(batch.marks)
    from peritext_b200.packing import INSDEL_DT, MARK_DT, PackedBatch
    pbatch = PackedBatch(batch.desc,
                         p_ins.numpy()[: batch.insdel.nbytes].view(INSDEL_DT), p_mk.numpy()[: batch.marks.nbytes].view(MARK_DT),
                         batch.values, batch.link_attrs, batch.comment_ids, batch.other_attrs, batch.meta)

    # a non-default stream: the engine captures its launch sequence into a CUDA graph (not possible on the legacy stream)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = BatchEngine(local_rank, stream=stream.cuda_stream)
    eng.upload(pbatch)

    class _DevView:   # zero-copy torch view of the engine-owned per-log result headers
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    res_dev = torch.as_tensor(_DevView(eng.device_results_ptr(), n_logs * 32), device=dev) if n_logs else torch.zeros(0, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * n_logs * 32, dtype=torch.uint8, device=dev) if world > 1 else None

    def step():
        eng.merge()
        if world > 1:
            dist.all_gather_into_tensor(gathered, res_dev)   # the path's only exchange: digests for the convergence check

    stop_evt, samples = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop_evt, samples, local_rank), daemon=True)
    th.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches0 = eng.launch_count
    # keep the GPU busy for a moment so the clock sampler sees the loaded state as well
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        kernel_ms.append(None)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0

    # per-launch duration of the dominant kernel: one merge timed alone with the engine's own events
    lone = []
    for _ in range(min(5, args.steps)):
        eng.merge(); eng.sync(); lone.append(eng.last_merge_ms)

    results = eng.results()
    ok = bool((results["status"] == 0).all())
    dig = results["digest"].reshape(n_logs // R, R, 2) if n_logs else np.zeros((0, R, 2), np.uint64)
    converged = bool((dig == dig[:, :1, :]).all())
    if world > 1:
        allres = gathered.cpu().numpy().view(RESULT_DT).reshape(world, n_logs)
        ok = ok and bool((allres["status"] == 0).all())
        d2 = allres["digest"].reshape(world, n_logs // R, R, 2)
        converged = converged and bool((d2 == d2[:, :, :1, :]).all())

    # ---- e2e through the public API: pinned host -> device, merge, results (headers + text + spans) back to host -------
    e2e = None
    if not args.no_e2e:
        # the public batch API: PipelinedEngine cuts the batch into 4 runs of logs (own handle + stream each) so that the
        # upload of one overlaps the merge and the download of the others
        from peritext_b200.engine import PipelinedEngine, compress_runs
        pipe = PipelinedEngine(local_rank, chunks=4)
        # the host -> device leg uses the run-compressed wire form (typing runs / consecutive deletes as one record),
        # expanded on the device; compressed once here, like the packing itself, outside the timed region
        keep = []

        def pin(a):
            t = torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory() if a.nbytes else torch.zeros(16, dtype=torch.uint8).pin_memory()
            keep.append(t)
            return t.numpy()[: a.nbytes].view(a.dtype)
        pruns = compress_runs(pbatch, pin=pin)
        in_bytes_e2e = pruns.nbytes
        pbatch_e2e = pruns
        outs = None
        for _ in range(2):
            outs = pipe.run(pbatch_e2e)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            outs = pipe.run(pbatch_e2e)
        torch.cuda.synchronize()
        e_ms = 1e3 * (time.perf_counter() - t0) / e_steps
        d2h = sum(o.results.nbytes + o.text.nbytes + o.spans.nbytes + o.comment_pool.nbytes for o in outs)
        e_res = np.concatenate([o.results for o in outs])
        e2e_ok = bool((e_res["status"] == 0).all()) and e_res["digest"].tobytes() == results["digest"].tobytes()
        ok = ok and e2e_ok
        e2e = {"ms": e_ms, "h2d": in_bytes_e2e, "d2h": int(d2h)}
        pipe.close()

    stop_evt.set(); th.join(timeout=2)

    # reduce over ranks: time = max, work = sum
    t_ms = elapsed_ms
    if world > 1:
        t = torch.tensor([elapsed_ms, e2e["ms"] if e2e else 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ms = float(t[0]); e2e_ms = float(t[1])
    else:
        e2e_ms = e2e["ms"] if e2e else None
    total_ops_per_step = ops_per_step_local * world
    ms_per_step = t_ms / args.steps
    value = total_ops_per_step / (ms_per_step / 1e3)

    if rank == 0:
        alg_bytes = batch.algorithmic_bytes(results)
        lone_ms = sorted(lone)[len(lone) // 2]
        peak, peak_src = hbm_peak()
        achieved = alg_bytes / (lone_ms / 1e3) / 1e9
        traffic, traffic_src = ncu_traffic(args.config, n_docs)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfg['label']}", "docs_per_gpu": n_docs, "replicas": R,
                       "logs_per_gpu": n_logs, "op_records_per_step_per_gpu": ops_per_step_local,
                       "unique_ops_per_gpu": batch.meta["unique_ops"], "input_bytes_per_gpu": in_bytes,
                       "parallelism": f"doc-sharded x{world}", "l2": "input (%.0f MB) larger than the 126 MB L2; no flush needed" % (in_bytes / 1e6)
                       if in_bytes > 130e6 else "input smaller than L2 (steps may hit L2)",
                       "generator_s": round(gen_s, 2), "all_status_ok": ok, "replicas_converged": converged,
                       "docs_per_sec": (n_logs * world) / (ms_per_step / 1e3), "kernel_paths": eng.stats()},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "ptk::merge_logs_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "launch_ms": lone_ms},
            "clocks": clocks_summary(samples),
        }
        if e2e:
            line["e2e"] = {"value": total_ops_per_step / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(e2e["h2d"]),
                           "d2h_bytes_per_step": int(e2e["d2h"]), "ms_per_step": e2e_ms,
                           "api": "peritext_b200.engine.PipelinedEngine.run on the run-compressed wire form (4 chunks: pt_batch_upload_runs / merge / download_begin / download per chunk)",
                           "uncompressed_input_bytes": int(in_bytes)}
        if not args.no_cpu_baseline and world == 1:
            v, cores, info = cpu_baseline(batch)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": info}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if (ok and converged) else 3


if __name__ == "__main__":
    sys.exit(main())
