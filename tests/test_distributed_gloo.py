"""world_size-2 test of the N>1 path on CPU (gloo): token-balanced document sharding, global
Philox offsets, the start-up count sum and the per-sweep "delta -> all-reduce(sum) -> apply"
exchange. The per-shard sampling runs in the oracle (no GPU here); everything around it is the
host logic the GPU ranks use. Property checked: the sharded DEFERRED chain equals the single-shard
chain bit for bit, and the replicas' counts stay equal to a recount from z."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ALPHA, BETA, SEED, K, V = 0.1, 0.01, 17, 12, 150


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from ldagibbssampling_b200.partition import partition_by_tokens, shard_corpus
    from oracle import oracle as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    dp, tok = O.gen_corpus(240, V, 35.0, 6, 21)          # every rank sees the same corpus description
    shard = partition_by_tokens(dp, world)[rank]
    ldp, ltok = shard_corpus(dp, tok, shard)
    z = O.init_z(shard.num_tokens, K, SEED, global_off=shard.token_begin)

    # start-up: each shard counts its own documents, one all-reduce makes the replicas global
    nwk, nk = O.count(ldp, ltok, z, V, K)
    buf = torch.from_numpy(np.concatenate([nwk.reshape(-1), nk]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    nwk = buf[:V * K].numpy().reshape(V, K).copy()
    nk = buf[V * K:].numpy().copy()

    for sweep in (1, 2, 3):
        z, d_nwk, d_nk = O.spec_sweep_given_counts(ldp, ltok, z, nwk, nk, ALPHA, BETA, SEED, sweep,
                                                   global_off=shard.token_begin)
        ex = torch.from_numpy(np.concatenate([d_nwk.reshape(-1), d_nk]))   # the exchange buffer layout
        dist.all_reduce(ex, op=dist.ReduceOp.SUM)
        nwk = nwk + ex[:V * K].numpy().reshape(V, K)
        nk = nk + ex[V * K:].numpy()

    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), z=z, nwk=nwk, nk=nk, begin=shard.token_begin)
    dist.destroy_process_group()


def test_two_rank_deferred_chain_equals_single_shard(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert r0["begin"] == 0 and r1["begin"] == len(r0["z"])
    z = np.concatenate([r0["z"], r1["z"]])
    dp, tok = oracle.gen_corpus(240, V, 35.0, 6, 21)
    want = oracle.spec_sweeps(dp, tok, oracle.init_z(len(tok), K, SEED), V, K, ALPHA, BETA, SEED, 1, 3)
    assert np.array_equal(z, want)
    nwk, nk = oracle.count(dp, tok, z, V, K)
    for r in (r0, r1):
        assert np.array_equal(r["nwk"], nwk) and np.array_equal(r["nk"], nk)
