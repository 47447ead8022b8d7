#!/usr/bin/env python
"""bench.py — sampled tokens/s of the collapsed-Gibbs LDA hot path (BASELINE.json metric).

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the CPU arm (Mallet-faithful oracle port)

A "step" is one Gibbs sweep over the whole corpus: per-sweep table build + the sampling kernel +
the AD-LDA count exchange (the library's own in-place NCCL all-reduce of the n_wk replica when
N > 1: b200lda_comm_init + b200lda_group_sweep). Default
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
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="timed CPU work per thread count")
    ap.add_argument("--after-sweeps", type=int, default=50,
                    help="second timed block of --steps sweeps starting at this sweep of the chain (0: off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ll-marks", default="", help="comma-separated sweep counts: also report LL/token of a chain "
                                                   "started from oracle.init_z(seed 7) at those sweeps (1 GPU)")
    ap.add_argument("--seed", type=int, default=1234)
    return ap.parse_args()


def measured_traffic(workload, K, default_K):
    """DRAM bytes per token of the sampling kernel from the committed ncu capture of this workload
    (profiles/r02_traffic_<workload>.json: dram__bytes_read.sum + dram__bytes_write.sum over one
    sweep's class launches). None when no capture exists for this exact workload / K."""
    if K != default_K:
        return None, None
    for rnd in ("r02", "r01"):
        try:
            with open(os.path.join(ROOT, "profiles", f"{rnd}_traffic_{workload}.json")) as f:
                d = json.load(f)
            return float(d["traffic_bytes_per_token"]), d["source"]
        except Exception:
            continue
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


def mallet_probe():
    """BASELINE.md promises a CPU-Mallet row if a JVM and the pinned jar (reference pom.xml:107-111:
    cc.mallet:mallet:2.0.7) are found on the box: this is the probe. Neither exists in this image."""
    import shutil
    jar = os.path.join(ROOT, "baseline", "_ref", "mallet-2.0.7.jar")
    return {"java": shutil.which("java") is not None, "mallet_jar": os.path.exists(jar)}


def _time_port(dp, tok, V, K, threads, warmup, steps, seconds, z0):
    """One estimate() call per timed block, as the reference does (the worker replicas are built
    once per call): returns tokens/s, ms per sweep, the sampling / merge / set-up split, LL/token."""
    from oracle import oracle as O
    m = O.MalletModel(K, ALPHA_K * K, BETA, seed=1, threads=threads)
    m.add_instances(dp, tok, V, z_init=z0)
    n = len(tok)
    if warmup:
        m.estimate(warmup)
    if steps is None:  # size the block from one probe sweep
        t0 = time.perf_counter()
        m.estimate(1)
        probe = time.perf_counter() - t0
        steps = int(max(2, min(200, seconds / max(probe, 1e-3))))
        warmup += 1
    s0 = m.timers()
    t0 = time.perf_counter()
    m.estimate(steps)
    total = time.perf_counter() - t0
    s1 = m.timers()
    setup, sample, merge = (b - a for a, b in zip(s0, s1))
    ll = m.model_log_likelihood() / n
    m.close()
    return {"threads": threads, "tokens_per_s": n * steps / total, "ms_per_sweep": 1e3 * total / steps, "sweeps": steps,
            "warmup": warmup, "sample_frac": sample / total, "merge_frac": merge / total, "setup_frac": setup / total,
            "ll_per_token": ll, "ll_after_sweeps": warmup + steps}


def run_cpu(workload, K, cpu_docs, warmup, steps=None, seconds=None):
    """Times oracle/mallet_sparse_lda.c on a bounded sample (cpu_docs documents of the workload's
    shape) with 1 thread and with all host threads (Mallet's AD-LDA replicas), from the same initial
    topics. Returns (tokens/s of the faster one, its ms per sweep, info)."""
    import bench_corpus as BC
    from oracle import oracle as O
    O.build()
    threads = os.cpu_count() or 1
    dp, tok, V, K0 = BC.cpu_sample(workload, cpu_docs)
    K = K or K0
    z0 = O.init_z(len(tok), K, 7)
    runs = [_time_port(dp, tok, V, K, 1, warmup, steps, seconds, z0)]
    if threads > 1:
        runs.append(_time_port(dp, tok, V, K, threads, warmup, steps, seconds, z0))
    best = max(runs, key=lambda r: r["tokens_per_s"])
    n = len(tok)
    info = {"cores": best["threads"], "kind": "port",
            "sample": f"{cpu_docs} docs / {n} tokens of the {workload} shape (V={V}, K={K}) on {threads} x {host_cpu_model()}; "
                      + "; ".join(f"T={r['threads']}: {r['tokens_per_s']:.3g} tok/s, {r['warmup']} warm-up + {r['sweeps']} timed sweeps "
                                  f"in one estimate() call, sampling {100 * r['sample_frac']:.0f} % / merge {100 * r['merge_frac']:.0f} % "
                                  f"/ replica set-up {100 * r['setup_frac']:.0f} % of the time" for r in runs)
                      + f"; value = the faster (T={best['threads']})",
            "runs": runs, "ll_per_token": best["ll_per_token"], **mallet_probe(),
            "corpus": (dp, tok, V, K, z0)}
    return best["tokens_per_s"], best["ms_per_sweep"], info


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
        "cpu_baseline": {"value": value, "unit": "tokens/s",
                         **{k: info[k] for k in ("cores", "kind", "sample", "runs", "java", "mallet_jar")}},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "ll_per_token": info["ll_per_token"],
    }
    _emit(line)


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
    h_z = torch.empty(shard.num_tokens, dtype=torch.uint16, pin_memory=True)   # topics in the device's own width
    del words_dev, lengths
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    gen_s = time.perf_counter() - t0

    # ---- sampler on torch's current stream so torch CUDA events bracket its kernels --------------
    # (a dedicated non-default stream: the legacy default stream has handle 0, which the C ABI
    # reads as "create a private stream")
    from ldagibbssampling_b200 import _capi
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    s = L.Sampler(K, V, ALPHA_K * K, BETA, seed=args.seed,
                  mode=L.MODE_LIVE if args.mode == "live" else L.MODE_DEFERRED, device=local_rank,
                  rank=rank, world_size=world, global_token_offset=shard.token_begin,
                  global_doc_offset=shard.doc_begin, stream=stream.cuda_stream)
    if world > 1:
        # the exchange is the library's own: its NCCL communicator from a unique id that travels
        # through torch.distributed (plumbing); torch's NCCL is not on the data path
        box = [_capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        s.comm_init(box[0])
    s.load_corpus_raw(shard.num_docs, h_doc_ptr.data_ptr(), h_words.data_ptr(), shard.num_tokens)
    s.init_assignments(None)

    def sync_counts():
        # every shard counted only its own documents: sum once so all n_wk / n_k replicas are global
        if world > 1:
            _capi.group_sync_counts([s])

    sync_counts()

    def sweeps(n):
        _capi.group_sweep([s], n)   # n whole sweeps enqueued back to back (tables, sampling, exchange), one sync at the end

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_block(n):
        """n sweeps between two events on the sampler's stream, barrier + synchronize on both sides;
        returns (ms, stats of the block), max over ranks."""
        fence()
        s.reset_stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fence()
        ev0.record(stream)
        sweeps(n)
        ev1.record(stream)
        fence()
        st = s.stats()
        t = torch.tensor([ev0.elapsed_time(ev1), st["cum_sample_ms"], st["cum_tables_ms"], st["cum_finish_ms"]],
                         dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()], st

    sweeps(args.warmup)
    fence()
    launches0 = s.stats()["kernel_launches"]
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    (ms_total, cum_sample_ms, cum_tables_ms, cum_finish_ms), st = timed_block(args.steps)
    clk = clocks.stop() if rank == 0 else None
    launches = s.stats()["kernel_launches"] - launches0
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
                "finish_ms": cum_finish_ms / max(1, st["cum_sweeps"]),
                "table_refresh": st["table_refresh_last"], "hot_words": st["hot_words"],
                "prior_rows_rebuilt_last_sweep": st["rows_refreshed_last"]}

    # ---- count invariants on every rank, on the device (north star: sum n_wk = sum n_dk = N) ----
    def invariants():
        sum_nk, sum_nwk, bad_cols, sum_ndk = s.check_invariants()
        ok = sum_nk == N_global and sum_nwk == N_global and bad_cols == 0 and sum_ndk == shard.num_tokens
        nk_t = torch.from_numpy(s.nk().astype(np.int64)).to(dev)
        flags = torch.tensor([0 if ok else 1], dtype=torch.int64, device=dev)
        if world > 1:   # every replica must hold the same n_k
            lo, hi = nk_t.clone(), nk_t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            if not bool((lo == hi).all()):
                flags += 1
            dist.all_reduce(flags, op=dist.ReduceOp.SUM)
        return int(flags.item()) == 0

    invariants_ok = invariants()

    # ---- steady block: the same measurement further into the chain (a 1000-iteration estimate()
    # spends its time there, not in the first 25 sweeps from a uniform random init) ----------------
    steady = None
    done = args.warmup + args.steps
    if args.after_sweeps > done:
        sweeps(args.after_sweeps - done)
        (ms2, samp2, _, _), st2 = timed_block(args.steps)
        n2 = max(1, shard.num_tokens * st2["cum_sweeps"])
        steady = {"after_sweeps": args.after_sweeps, "steps": args.steps, "value": N_global * args.steps / (ms2 / 1e3),
                  "unit": "tokens/s", "ms_per_step": ms2 / args.steps, "kernel_ms": samp2 / max(1, st2["cum_sweeps"]),
                  "mean_doc_topics": st2["cum_doc_topics"] / n2, "moved_frac": st2["cum_tokens_moved"] / n2,
                  "prior_frac": st2["cum_prior_bucket"] / n2}
        dpt, wpt = s.loglik_parts()
        t = torch.tensor([dpt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        steady["ll_per_token"] = (float(t.item()) + wpt) / N_global
        invariants_ok = invariants() and invariants_ok

    # ---- end to end through the C ABI with HOST buffers: corpus + topics in, one sweep, topics out
    s.assignments_u16_raw(h_z.data_ptr())  # current chain state, host side
    fence()
    e2e_times = []
    for i in range(args.e2e_steps + 1 if args.e2e_steps > 0 else 0):
        fence()
        t0 = time.perf_counter()
        s.load_corpus_raw(shard.num_docs, h_doc_ptr.data_ptr(), h_words.data_ptr(), shard.num_tokens)
        s.init_assignments_u16_raw(h_z.data_ptr())
        sync_counts()
        sweeps(1)
        s.assignments_u16_raw(h_z.data_ptr())
        fence()
        if i > 0:  # first repetition warms the path
            e2e_times.append(time.perf_counter() - t0)
    e2e_t = torch.tensor([float(np.mean(e2e_times)) if e2e_times else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())
    h2d = 8 * (shard.num_docs + 1) + 4 * shard.num_tokens + 2 * shard.num_tokens
    d2h = 2 * shard.num_tokens
    e2e = {"value": (N_global / e2e_s) if e2e_s > 0 else None, "unit": "tokens/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s,
           "call": "b200lda_load_corpus + b200lda_init_assignments_u16(z) + (count sync) + b200lda_group_sweep(1) "
                   "+ b200lda_get_assignments_u16, pinned host buffers, per rank shard"}

    # ---- LL trajectory in the same run (BASELINE config 1: 500 sweeps against the committed Mallet-port curve)
    ll_trajectory = None
    if args.ll_marks and world == 1:
        from oracle import oracle as O
        O.build()
        marks = sorted(int(x) for x in args.ll_marks.split(","))
        tok_np, dp_np = h_words.numpy(), h_doc_ptr.numpy()
        g = L.Sampler(K, V, ALPHA_K * K, BETA, seed=7, mode=L.MODE_LIVE if args.mode == "live" else L.MODE_DEFERRED,
                      device=local_rank)
        g.load_corpus(dp_np, tok_np)
        g.init_assignments(O.init_z(len(tok_np), K, 7))
        done_t, curve = 0, []
        for mk in marks:
            g.sweep(mk - done_t)
            done_t = mk
            curve.append(g.loglik() / len(tok_np))
        g.close()
        ll_trajectory = {"init": "oracle.init_z(N, K, seed=7)", "sweeps": marks, "b200_ll_per_token": curve}
        gold = os.path.join(ROOT, "tests", "golden", f"{args.workload}_ll_trajectory.json")
        if os.path.exists(gold) and not args.docs and not args.topics:
            gj = json.load(open(gold))
            if gj.get("tokens") == len(tok_np):   # same corpus: the committed Mallet-port curve beside it
                ll_trajectory["mallet_port"] = {"sweeps": gj["sweeps"], "ll_per_token_by_seed": gj["mallet_ll_per_token"]}
            else:
                ll_trajectory["note"] = ("bench_corpus generates this workload on the GPU; the committed Mallet-port curve "
                                         "(tests/golden) is for the oracle generator's corpus of the same shape")

    cpu = None
    ll_same_corpus = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, _, info = run_cpu(args.workload, K, args.cpu_docs, warmup=1, seconds=args.cpu_seconds)
        cpu = {"value": v, "unit": "tokens/s", **{k: info[k] for k in ("cores", "kind", "sample", "runs", "java", "mallet_jar")}}
        # LL/token of both sides on the SAME corpus (the CPU sample), same initial topics, same sweep count
        cdp, ctok, cV, cK, cz0 = info["corpus"]
        t1 = info["runs"][0]
        g = L.Sampler(cK, cV, ALPHA_K * cK, BETA, seed=7, mode=L.MODE_LIVE if args.mode == "live" else L.MODE_DEFERRED,
                      device=local_rank)
        g.load_corpus(cdp, ctok)
        g.init_assignments(cz0)
        g.sweep(t1["ll_after_sweeps"])
        ll_same_corpus = {"corpus": f"the CPU baseline's {args.cpu_docs}-document sample", "sweeps": t1["ll_after_sweeps"],
                          "cpu_port_T1": t1["ll_per_token"], "b200": g.loglik() / len(ctok),
                          "note": "same initial topics, same sweep count; on a corpus this small the early chain is where "
                                  "the samplers differ most (tests/test_gpu_ll_parity.py follows it to sweep 200)"}
        if len(info["runs"]) > 1:   # Mallet's own AD-LDA with all host threads, for scale
            tn = info["runs"][-1]
            g.sweep(max(0, tn["ll_after_sweeps"] - t1["ll_after_sweeps"]))
            ll_same_corpus["cpu_port_all_threads"] = {"threads": tn["threads"], "sweeps": tn["ll_after_sweeps"],
                                                      "ll_per_token": tn["ll_per_token"],
                                                      "b200_at_that_sweep": g.loglik() / len(ctok)
                                                      if tn["ll_after_sweeps"] >= t1["ll_after_sweeps"] else None}
        g.close()

    if rank == 0:
        line = {
            "metric": "gibbs_sampled_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32+f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "docs": D, "tokens": N_global, "V": V, "K": K,
                       "alpha_k": ALPHA_K, "beta": BETA, "mode": args.mode,
                       "parallelism": f"ad-lda docs/{world} + in-place int32 all-reduce of the n_wk/n_k replica per sweep (library NCCL)",
                       "l2": "inputs exceed L2 (no flush needed)" if 4 * N_global / world > 256e6 else "inputs fit L2",
                       "corpus_gen_s": gen_s},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "ll_per_token": ll_per_token, "ll_same_corpus": ll_same_corpus,
            "invariants_ok": bool(invariants_ok), "steady": steady, "ll_trajectory": ll_trajectory,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def _emit(line):
    """The one JSON line of the contract, on the process's real stdout."""
    text = json.dumps(line) + "\n"
    if _RESULT_FD is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, text.encode())


def main():
    global _RESULT_FD
    args = parse_args()
    # stdout carries exactly one line. Native libraries write there too (NCCL prints its version
    # banner on communicator creation), so file descriptor 1 points at stderr while the arm runs and
    # the result goes to a duplicate of the original.
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
