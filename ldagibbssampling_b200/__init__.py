"""ldagibbssampling_b200 — B200-native (sm_100a) collapsed-Gibbs LDA sampler behind the entry
points the reference drives (Mallet's ParallelTopicModel, reference cmu_ron/TrainAndPredict.java:159-177).

The arithmetic lives in libb200lda.so (hand-written CUDA, C ABI in include/b200lda.h); this
package is the host-side mirror of the reference interface. No CPU fallback exists.
"""
from ._capi import (B200LDAError, MODE_DEFERRED, MODE_LIVE, Sampler, device_count, load_library)

__all__ = ["B200LDAError", "MODE_DEFERRED", "MODE_LIVE", "Sampler", "device_count", "load_library"]
