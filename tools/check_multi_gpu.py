#!/usr/bin/env python
"""Multi-GPU parity check, one process per GPU (run under torch.distributed.run):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port 29511 tools/check_multi_gpu.py

DEFERRED mode through the real NCCL all-reduce must reproduce the oracle's single-shard chain bit
for bit for any N; LIVE mode must keep the count invariants on every replica."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import ldagibbssampling_b200 as L
    from ldagibbssampling_b200.partition import partition_by_tokens, shard_corpus
    from ldagibbssampling_b200.topic_model import _DevBuf
    from oracle import oracle as O

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    D, V, K, ALPHA, BETA, SEED = 4000, 1500, 64, 0.1, 0.01, 12
    dp, tok = O.gen_corpus(D, V, 70.0, 16, 41)
    sh = partition_by_tokens(dp, world)[rank]
    ldp, ltok = shard_corpus(dp, tok, sh)
    ok = True
    for mode, name in ((L.MODE_DEFERRED, "deferred"), (L.MODE_LIVE, "live")):
        stream = torch.cuda.Stream(dev)
        with torch.cuda.stream(stream):
            s = L.Sampler(K, V, ALPHA * K, BETA, seed=SEED, mode=mode, device=local, rank=rank, world_size=world,
                          global_token_offset=sh.token_begin, global_doc_offset=sh.doc_begin, stream=stream.cuda_stream)
            s.load_corpus(ldp, ltok)
            s.init_assignments(None)
            ex = None
            if world > 1:
                ptr, n = s.exchange_buffer()
                ex = torch.as_tensor(_DevBuf(ptr, n), device=dev)
                s.counts_sync_begin()
                dist.all_reduce(ex)
                s.counts_sync_end()
            for _ in range(4):
                s.sweep_begin()
                if ex is not None:
                    dist.all_reduce(ex)
                s.sweep_end()
            s.synchronize()
            z_local = torch.from_numpy(s.assignments()).to(dev)
            sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([z_local.numel()], device=dev))
            maxn = int(max(x.item() for x in sizes))
            pad = torch.zeros(maxn, dtype=torch.int32, device=dev)
            pad[:z_local.numel()] = z_local
            parts = [torch.zeros(maxn, dtype=torch.int32, device=dev) for _ in range(world)]
            dist.all_gather(parts, pad)
            z = np.concatenate([p[:int(n.item())].cpu().numpy() for p, n in zip(parts, sizes)])
            nwk, nk = O.count(dp, tok, z, V, K)
            good = np.array_equal(s.nwk(), nwk) and np.array_equal(s.nk(), nk)
            if mode == L.MODE_DEFERRED:
                want = O.spec_sweeps(dp, tok, O.init_z(len(tok), K, SEED), V, K, ALPHA, BETA, SEED, 1, 4)
                good = good and np.array_equal(z, want)
            doc, word = s.loglik_parts()
            t = torch.tensor([doc], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            ll = float(t.item()) + word
            want_ll = O.loglik(dp, tok, z, V, K, ALPHA, BETA)
            good = good and abs(ll - want_ll) <= 1e-9 * abs(want_ll)
            print(f"rank {rank}/{world} {name}: {'OK' if good else 'MISMATCH'}  LL/token={ll / len(tok):.5f}", flush=True)
            ok = ok and good
            s.close()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
