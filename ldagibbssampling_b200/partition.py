"""AD-LDA document partitioning (host logic, CPU-testable).

Mallet splits documents into contiguous ranges of D/T documents per worker thread
(setNumThreads(4), reference cmu_ron/TrainAndPredict.java:164, cmu/TrainAndPredict.java:262;
SURVEY.md Appendix A.3). Here a "thread" is a GPU and ranges are balanced by TOKEN count, since
a sweep's cost is proportional to tokens, not documents. Ranges stay contiguous so a shard is
just a slice of the flattened corpus and a global token index (the Philox counter) is
shard-local index + offset.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shard:
    rank: int
    world_size: int
    doc_begin: int
    doc_end: int
    token_begin: int
    token_end: int

    @property
    def num_docs(self) -> int:
        return self.doc_end - self.doc_begin

    @property
    def num_tokens(self) -> int:
        return self.token_end - self.token_begin


def partition_by_tokens(doc_ptr, world_size: int):
    """Contiguous document ranges whose token counts are as equal as contiguity allows.

    Shard r starts at the first document whose starting token offset is >= r * N / world_size.
    Returns a list of `world_size` Shards covering [0, D) exactly once, in order.
    """
    doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
    if doc_ptr.ndim != 1 or len(doc_ptr) < 1 or doc_ptr[0] != 0:
        raise ValueError("doc_ptr must be a CSR offset array starting at 0")
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    D = len(doc_ptr) - 1
    N = int(doc_ptr[-1])
    starts = [0]
    for r in range(1, world_size):
        target = (N * r) // world_size
        d = int(np.searchsorted(doc_ptr[:-1], target, side="left")) if D > 0 else 0
        starts.append(min(max(d, starts[-1]), D))
    starts.append(D)
    return [Shard(r, world_size, starts[r], starts[r + 1], int(doc_ptr[starts[r]]),
                  int(doc_ptr[starts[r + 1]])) for r in range(world_size)]


def shard_corpus(doc_ptr, tok_word, shard: Shard):
    """The shard's own CSR (doc_ptr rebased to 0) and token slice."""
    doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
    local_ptr = doc_ptr[shard.doc_begin:shard.doc_end + 1] - doc_ptr[shard.doc_begin]
    return np.ascontiguousarray(local_ptr), np.ascontiguousarray(tok_word[shard.token_begin:shard.token_end])
