#!/usr/bin/env python
"""bench.py — sampled tokens/s of the collapsed-Gibbs LDA hot path (BASELINE.json metric).

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the CPU arm (Mallet-faithful oracle port)

A "step" is one Gibbs sweep over the whole corpus: per-sweep table build + the sampling kernel +
the AD-LDA count exchange (NCCL all-reduce of the int32 n_wk/n_k delta when N > 1). Default
workload = BASELINE.json config "synthetic PubMed-shaped corpus: 8.2M docs, V=141k, 738M tokens,
K=1000" (C4), split over the N GPUs by tokens (strong scaling, as that config states).
Prints ONE JSON line on rank 0.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALPHA_K = 0.1   # symmetric alpha_k (Mallet ctor alphaSum = 0.1*K), SURVEY.md §8(d)
BETA = 0.01
HBM_FALLBACK_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--mode", default="live", choices=["live", "deferred"])
    ap.add_argument("--docs", type=int, default=0, help="override document count (debug)")
    ap.add_argument("--topics", type=int, default=0, help="override K (C5 sweep)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-docs", type=int, default=20000, help="documents in the CPU baseline sample")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=1234)
    return ap.parse_args()


def measured_traffic(workload, K, default_K):
    """DRAM bytes per token of the sampling kernel from the committed ncu capture of this workload
    (profiles/r01_traffic_<workload>.json: dram__bytes_read.sum + dram__bytes_write.sum over one
    sweep's class launches). None when no capture exists for this exact workload / K."""
    if K != default_K:
        return None, None
    try:
        with open(os.path.join(ROOT, "profiles", f"r01_traffic_{workload}.json")) as f:
            d = json.load(f)
        return float(d["traffic_bytes_per_token"]), d["source"]
    except Exception:
        return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the Mallet-faithful oracle port on the host cores (cpu_baseline / --impl reference)
# ------------------------------------------------------------------------------------------------

def host_cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown CPU"


def run_cpu(workload, K, cpu_docs, warmup, steps=None, seconds=None):
    """Times AD-LDA sweeps of oracle/mallet_sparse_lda.c with T = all host threads on a bounded
    sample (cpu_docs documents of the workload's shape). Returns (tokens/s, ms_per_step, info)."""
    import bench_corpus as BC
    from oracle import oracle as O
    O.build()
    threads = os.cpu_count() or 1
    dp, tok, V, K0 = BC.cpu_sample(workload, cpu_docs)
    K = K or K0
    m = O.MalletModel(K, ALPHA_K * K, BETA, seed=1, threads=threads)
    m.add_instances(dp, tok, V)
    n = len(tok)
    for _ in range(warmup):
        m.estimate(1)
    times = []
    t_begin = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        m.estimate(1)
        times.append(time.perf_counter() - t0)
        if steps is not None and len(times) >= steps:
            break
        if steps is None and len(times) >= 2 and time.perf_counter() - t_begin >= seconds:
            break
    ll = m.model_log_likelihood() / n
    m.close()
    total = sum(times)
    info = {"cores": threads, "kind": "port",
            "sample": f"{cpu_docs} docs / {n} tokens of the {workload} shape (V={V}, K={K}), "
                      f"{warmup} warm-up + {len(times)} timed AD-LDA sweeps, T={threads} worker replicas "
                      f"on {threads} x {host_cpu_model()}",
            "ll_per_token": ll}
    return n * len(times) / total, 1e3 * total / len(times), info


def reference_arm(args):
    import bench_corpus as BC
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = BC.WORKLOADS[args.workload]
    K = args.topics or w["K"]
    value, ms, info = run_cpu(args.workload, K, args.cpu_docs, args.warmup, steps=args.steps)
    line = {
        "impl": "reference", "metric": "gibbs_sampled_tokens_per_s", "value": value, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32+f64",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "K": K, "alpha_k": ALPHA_K, "beta": BETA,
                   "note": "reference arm = CPU port of Mallet 2.0.7 SparseLDA/AD-LDA (oracle/), the "
                           "reference itself is Java + an un-vendored jar and no JVM exists here"},
        "cpu_baseline": {"value": value, "unit": "tokens/s", **{k: info[k] for k in ("cores", "kind", "sample")}},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "ll_per_token": info["ll_per_token"],
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------

class _DevBuf:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}


def b200_arm(args):
    import torch
    import torch.distributed as dist

    import bench_corpus as BC
    import ldagibbssampling_b200 as L
    from ldagibbssampling_b200.partition import partition_by_tokens

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a B200; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = BC.WORKLOADS[args.workload]
    D = args.docs or w["D"]
    V, K, k_true = w["V"], (args.topics or w["K"]), w["k_true"]

    # ---- corpus: lengths for all documents (identical on every rank), tokens for own shard -------
    t0 = time.perf_counter()
    lengths = BC.doc_lengths(D, w["mean_len"], w["seed"], dev)
    doc_ptr_g = np.zeros(D + 1, np.int64)
    doc_ptr_g[1:] = torch.cumsum(lengths, 0).cpu().numpy()
    N_global = int(doc_ptr_g[-1])
    shard = partition_by_tokens(doc_ptr_g, world)[rank]
    phi_cdf = BC.phi_flat_cdf(V, k_true, w["seed"], dev)
    words_dev = BC.generate_docs(shard.doc_begin, shard.doc_end, lengths, phi_cdf, V, k_true, w["seed"], dev)
    assert words_dev.numel() == shard.num_tokens
    del phi_cdf
    # host-resident copies in pinned memory: what a caller of the C ABI holds
    h_doc_ptr = torch.from_numpy(doc_ptr_g[shard.doc_begin:shard.doc_end + 1] - doc_ptr_g[shard.doc_begin]).pin_memory()
    h_words = torch.empty(shard.num_tokens, dtype=torch.int32, pin_memory=True)
    h_words.copy_(words_dev)
    h_z = torch.empty(shard.num_tokens, dtype=torch.int32, pin_memory=True)
    del words_dev, lengths
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    gen_s = time.perf_counter() - t0

    # ---- sampler on torch's current stream so torch CUDA events bracket its kernels --------------
    # (a dedicated non-default stream: the legacy default stream has handle 0, which the C ABI
    # reads as "create a private stream")
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    s = L.Sampler(K, V, ALPHA_K * K, BETA, seed=args.seed,
                  mode=L.MODE_LIVE if args.mode == "live" else L.MODE_DEFERRED, device=local_rank,
                  rank=rank, world_size=world, global_token_offset=shard.token_begin,
                  global_doc_offset=shard.doc_begin, stream=stream.cuda_stream)
    s.load_corpus_raw(shard.num_docs, h_doc_ptr.data_ptr(), h_words.data_ptr(), shard.num_tokens)
    s.init_assignments(None)
    ex = None
    if world > 1:
        ptr, n = s.exchange_buffer()
        ex = torch.as_tensor(_DevBuf(ptr, n), device=dev)

    def sync_counts():
        # every shard counted only its own documents: sum once so all n_wk / n_k replicas are global
        if ex is not None:
            s.counts_sync_begin()
            dist.all_reduce(ex, op=dist.ReduceOp.SUM)
            s.counts_sync_end()

    sync_counts()

    def one_sweep():
        s.sweep_begin()
        if ex is not None:
            dist.all_reduce(ex, op=dist.ReduceOp.SUM)
        s.sweep_end()

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_sweep()
    fence()
    s.reset_stats()
    launches0 = s.stats()["kernel_launches"]

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    ev0.record(stream)
    for _ in range(args.steps):
        one_sweep()
    ev1.record(stream)
    fence()
    ms_total = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    st = s.stats()

    t = torch.tensor([ms_total, st["cum_sample_ms"], st["cum_tables_ms"], st["cum_finish_ms"]],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, cum_sample_ms, cum_tables_ms, cum_finish_ms = [float(x) for x in t.tolist()]
    value = N_global * args.steps / (ms_total / 1e3)

    # ---- LL/token after the timed sweeps (doc parts summed over shards) --------------------------
    doc_part, word_part = s.loglik_parts()
    dp_t = torch.tensor([doc_part], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dp_t, op=dist.ReduceOp.SUM)
    ll_per_token = (float(dp_t.item()) + word_part) / N_global

    # ---- roofline of the dominant kernel (k_gibbs_sweep), this rank's launches -------------------
    tokens_rank = shard.num_tokens
    sampled = max(1, tokens_rank * st["cum_sweeps"])
    kd = st["cum_doc_topics"] / sampled
    f_moved = st["cum_tokens_moved"] / sampled
    f_prior = st["cum_prior_bucket"] / sampled
    mean_len = tokens_rank / max(1, shard.num_docs)
    nlev = 1
    n = K
    while n > 32:
        n = (n + 31) // 32
        nlev += 1
    # SURVEY.md §8(d) contract figure for the sampling kernel (the 28*V*K/N_g table term belongs to
    # the table/delta kernels, not to this launch): 36 + 4*Kd + 8*Kd/L bytes per token.
    a_alg = 36.0 + 4.0 * kd + 8.0 * kd / mean_len
    # bytes this implementation must move per token: word id 4 + z 2 + prior mass 4 +
    # 4*Kd gathers + doc row 8*Kd/L + moved tokens (z 2 + 2 n_wk RMW 16 + 2 n_k RMW 16) +
    # prior draws (P_w[o] 4 + one 128 B line per search level)
    a_impl = 10.0 + 4.0 * kd + 8.0 * kd / mean_len + f_moved * 34.0 + f_prior * (4.0 + 128.0 * nlev)
    kernel_ms = cum_sample_ms / max(1, st["cum_sweeps"])
    peak, peak_src = measured_peaks()
    achieved = tokens_rank * a_alg / (kernel_ms / 1e3) / 1e9
    bpt, traffic_src = measured_traffic(args.workload, K, w["K"])
    roofline = {"bound": "hbm", "kernel": "k_gibbs_sweep", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": (bpt * tokens_rank) if bpt is not None else None,  # DRAM bytes per sweep of this rank
                "traffic_source": traffic_src, "peak_source": peak_src,
                "dram_frac": (bpt * tokens_rank / (kernel_ms / 1e3) / 1e9 / peak) if bpt is not None else None,
                "alg_bytes_per_token": a_alg, "impl_bytes_per_token": a_impl,
                "achieved_impl_bytes": tokens_rank * a_impl / (kernel_ms / 1e3) / 1e9,
                "kernel_ms": kernel_ms, "tokens_per_launch": tokens_rank,
                "mean_doc_topics": kd, "moved_frac": f_moved, "prior_frac": f_prior,
                "kernel_share_of_step": cum_sample_ms / ms_total,
                "tables_ms": cum_tables_ms / max(1, st["cum_sweeps"]),
                "finish_ms": cum_finish_ms / max(1, st["cum_sweeps"])}
    launches = st["kernel_launches"] - launches0

    # ---- end to end through the C ABI with HOST buffers: corpus + topics in, one sweep, topics out
    s.assignments_raw(h_z.data_ptr())  # current chain state, host side
    fence()
    e2e_times = []
    for i in range(args.e2e_steps + 1 if args.e2e_steps > 0 else 0):
        fence()
        t0 = time.perf_counter()
        s.load_corpus_raw(shard.num_docs, h_doc_ptr.data_ptr(), h_words.data_ptr(), shard.num_tokens)
        s.init_assignments_raw(h_z.data_ptr())
        sync_counts()
        one_sweep()
        s.assignments_raw(h_z.data_ptr())
        fence()
        if i > 0:  # first repetition warms the path
            e2e_times.append(time.perf_counter() - t0)
    e2e_t = torch.tensor([float(np.mean(e2e_times)) if e2e_times else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    h2d = 8 * (shard.num_docs + 1) + 4 * shard.num_tokens + 4 * shard.num_tokens
    d2h = 4 * shard.num_tokens
    e2e = {"value": (N_global / e2e_s) if e2e_s > 0 else None, "unit": "tokens/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s,
           "call": "b200lda_load_corpus + b200lda_init_assignments(z) + sweep_begin/all-reduce/sweep_end "
                   "+ b200lda_get_assignments, pinned host buffers, per rank shard"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, _, info = run_cpu(args.workload, K, args.cpu_docs, warmup=1, seconds=args.cpu_seconds)
        cpu = {"value": v, "unit": "tokens/s", **{k: info[k] for k in ("cores", "kind", "sample")}}

    if rank == 0:
        line = {
            "metric": "gibbs_sampled_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32+f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "docs": D, "tokens": N_global, "V": V, "K": K,
                       "alpha_k": ALPHA_K, "beta": BETA, "mode": args.mode,
                       "parallelism": f"ad-lda docs/{world} + int32 all-reduce of n_wk/n_k delta per sweep",
                       "l2": "inputs exceed L2 (no flush needed)" if 4 * N_global / world > 256e6 else "inputs fit L2",
                       "corpus_gen_s": gen_s},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "ll_per_token": ll_per_token,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
