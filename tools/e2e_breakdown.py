#!/usr/bin/env python
"""Wall-clock breakdown of the end-to-end step (host buffers in, host buffers out) per C-ABI call, per
rank (max over ranks): load_corpus, init_assignments_u16, count sync, sweep, get_assignments_u16.
  python tools/e2e_breakdown.py [--docs N]                                  # one GPU
  python -m torch.distributed.run --nproc-per-node 8 ... tools/e2e_breakdown.py
Prints one JSON line on rank 0 (committed as profiles/r02_e2e_breakdown_n{N}.json)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4"); ap.add_argument("--docs", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch, torch.distributed as dist, bench_corpus as BC, ldagibbssampling_b200 as L
    from ldagibbssampling_b200 import _capi
    from ldagibbssampling_b200.partition import partition_by_tokens
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = BC.WORKLOADS[a.workload]; D = a.docs or w["D"]; V = w["V"]; K = w["K"]
    lengths = BC.doc_lengths(D, w["mean_len"], w["seed"], dev)
    dpg = np.zeros(D + 1, np.int64); dpg[1:] = torch.cumsum(lengths, 0).cpu().numpy()
    sh = partition_by_tokens(dpg, world)[rank]
    phi = BC.phi_flat_cdf(V, w["k_true"], w["seed"], dev)
    words = BC.generate_docs(sh.doc_begin, sh.doc_end, lengths, phi, V, w["k_true"], w["seed"], dev)
    h_dp = torch.from_numpy(dpg[sh.doc_begin:sh.doc_end + 1] - dpg[sh.doc_begin]).pin_memory()
    h_w = torch.empty(sh.num_tokens, dtype=torch.int32, pin_memory=True); h_w.copy_(words)
    h_z = torch.empty(sh.num_tokens, dtype=torch.uint16, pin_memory=True)
    del phi, words, lengths; torch.cuda.empty_cache()
    s = L.Sampler(K, V, 0.1 * K, 0.01, seed=1, device=lr, rank=rank, world_size=world,
                  global_token_offset=sh.token_begin, global_doc_offset=sh.doc_begin)
    if world > 1:
        box = [_capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        s.comm_init(box[0])

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    s.load_corpus_raw(sh.num_docs, h_dp.data_ptr(), h_w.data_ptr(), sh.num_tokens); s.init_assignments(None)
    if world > 1:
        _capi.group_sync_counts([s])
    _capi.group_sweep([s], 2)
    s.assignments_u16_raw(h_z.data_ptr())
    rows = []
    for r in range(a.reps + 1):
        fence()
        t = [time.perf_counter()]
        s.load_corpus_raw(sh.num_docs, h_dp.data_ptr(), h_w.data_ptr(), sh.num_tokens); t.append(time.perf_counter())
        s.init_assignments_u16_raw(h_z.data_ptr()); t.append(time.perf_counter())
        if world > 1:
            _capi.group_sync_counts([s]); s.synchronize()
        t.append(time.perf_counter())
        _capi.group_sweep([s], 1); t.append(time.perf_counter())
        s.assignments_u16_raw(h_z.data_ptr()); t.append(time.perf_counter())
        fence(); t.append(time.perf_counter())
        if r > 0:
            rows.append(np.diff(t) * 1e3)
    d = torch.tensor(np.mean(rows, 0), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
    d = d.tolist()
    if rank == 0:
        N = int(dpg[-1])
        print(json.dumps({"workload": a.workload, "n_gpus": world, "docs": D, "tokens": N, "tokens_per_rank": sh.num_tokens,
                          "reps": a.reps, "ms_max_over_ranks": {"load_corpus": d[0], "init_assignments_u16": d[1],
                                                                "count_sync": d[2], "sweep": d[3], "get_assignments_u16": d[4],
                                                                "barrier": d[5], "total": float(sum(d))},
                          "h2d_bytes_per_rank": 8 * (sh.num_docs + 1) + 6 * sh.num_tokens, "d2h_bytes_per_rank": 2 * sh.num_tokens,
                          "e2e_tokens_per_s": N / sum(d) * 1e3}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
