"""Host-side mirror of the Mallet entry points the reference drives, over the C ABI.

Same names, argument meaning and error behaviour as the used subset of
cc.mallet.topics.ParallelTopicModel / TopicInferencer, so the reference's
trainNewModel (cmu_ron/TrainAndPredict.java:159-171, cmu/TrainAndPredict.java:258-269) reads the
same with `ParallelTopicModel` imported from here:

    model = ParallelTopicModel(500, 100, 1)      # K, alphaSum (NOT alpha), beta
    model.addInstances(training)
    model.setOptimizeInterval(20); model.setNumThreads(4); model.setNumIterations(10000)
    model.estimate()
    inferencer = model.getInferencer()

All arithmetic happens in libb200lda.so on the GPU; there is no CPU fallback. setNumThreads(n)
means n AD-LDA shards = n GPUs (one context each; under torch.distributed one rank = one shard).
The per-sweep count exchange is the library's own (NCCL loaded by libb200lda.so): one communicator
set for the contexts of this process (b200lda_group_comm_init) or, one process per GPU, a
communicator per rank from a unique id broadcast through torch.distributed (b200lda_comm_init).
"""
from __future__ import annotations

import warnings
from typing import List, Optional

import numpy as np

from . import _capi
from .instances import Alphabet, FeatureSequence, Instance, InstanceList, LabelSequence, TopicAssignment
from .partition import partition_by_tokens, shard_corpus


class _DevBuf:
    """A device buffer of the library as something torch.as_tensor accepts (tests, tools)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 3}


class ParallelTopicModel:
    """Mirror of cc.mallet.topics.ParallelTopicModel (the subset the reference calls)."""

    DEFAULT_BETA = 0.01

    def __init__(self, numberOfTopics: int, alphaSum: Optional[float] = None, beta: float = DEFAULT_BETA):
        if numberOfTopics < 1:
            raise ValueError("numberOfTopics must be >= 1")
        self.numTopics = int(numberOfTopics)
        self.alphaSum = float(numberOfTopics if alphaSum is None else alphaSum)  # Mallet: (K) -> alphaSum = K
        self.alpha = np.full(self.numTopics, self.alphaSum / self.numTopics, np.float64)
        self.beta = float(beta)
        self.data: List[TopicAssignment] = []
        self.alphabet: Optional[Alphabet] = None
        self.numTypes = 0
        self.betaSum = 0.0
        # Mallet defaults
        self.numIterations = 1000
        self.burninPeriod = 200
        self.optimizeInterval = 50
        self.saveSampleInterval = 10
        self.showTopicsInterval = 50
        self.wordsPerTopic = 7
        self.numThreads = 1
        self.randomSeed = -1
        # B200 specifics
        self.mode = _capi.MODE_LIVE
        self.devices: Optional[List[int]] = None      # one entry per shard; default 0..numThreads-1
        self.distributed = False                      # True: this process is ONE shard of a torch.distributed job
        self.process_group = None
        self._samplers: List[_capi.Sampler] = []
        self._shards = []
        self._iterationsSoFar = 0
        self._dirty = True        # instances added since the device state was built
        self._z_host: Optional[np.ndarray] = None
        self._doc_ptr = np.zeros(1, np.int64)
        self._tok = np.zeros(0, np.int32)

    # -- configuration (same names as Mallet) -------------------------------------------------
    def setNumIterations(self, n: int):
        self.numIterations = int(n)

    def setBurninPeriod(self, n: int):
        self.burninPeriod = int(n)

    def setOptimizeInterval(self, n: int):
        self.optimizeInterval = int(n)

    def setNumThreads(self, n: int):
        """n AD-LDA shards (= n GPUs). The reference calls it AFTER addInstances
        (cmu_ron/TrainAndPredict.java:162-164): a change re-shards the documents at the next
        estimate(), keeping the chain state."""
        n = max(1, int(n))
        if n != self.numThreads:
            self._dirty = True
        self.numThreads = n

    def setRandomSeed(self, seed: int):
        self.randomSeed = int(seed)

    def setTopicDisplay(self, interval: int, n: int):
        self.showTopicsInterval = int(interval)
        self.wordsPerTopic = int(n)

    def setSamplingMode(self, mode):
        """B200 extension: 'live' (default) or 'deferred' (bit-reproducible, shard-count independent)."""
        self.mode = {"live": _capi.MODE_LIVE, "deferred": _capi.MODE_DEFERRED}.get(mode, mode)
        self._dirty = True

    def setDevices(self, devices):
        self.devices = list(devices)
        self._dirty = True

    def setSweepLog(self, path, logLikelihoodEvery: int = 0):
        """B200 extension (SURVEY.md §5 metrics): one JSON line per sweep to `path` - sweep number, device
        ms of the table build / sampling kernel / exchange, sampled tokens/s, moved and prior-bucket
        fractions, mean non-zero doc topics, table rebuilds, and LL/token every logLikelihoodEvery sweeps.
        Logging synchronises after every sweep."""
        self._sweep_log = path
        self._sweep_log_ll = int(logLikelihoodEvery)

    def setDistributed(self, flag: bool = True, group=None):
        self.distributed = bool(flag)
        self.process_group = group
        self._dirty = True

    def getAlphabet(self) -> Alphabet:
        return self.alphabet

    def getData(self):
        return self.data

    def getNumTopics(self) -> int:
        return self.numTopics

    # -- addInstances ---------------------------------------------------------------------------
    def addInstances(self, training: InstanceList):
        """Takes the documents (FeatureSequences over one alphabet), draws the initial topics and
        builds the counts. May be called again with more documents (the reference's updateModel,
        cmu_ron/TrainAndPredict.java:173-177): existing assignments are kept."""
        alphabet = training.getDataAlphabet()
        if self.alphabet is None:
            self.alphabet = alphabet
        elif alphabet is not self.alphabet:
            raise ValueError("instances must share the model's data alphabet")
        self._pull_assignments()          # keep the chain state of documents already in the model
        doc_ptr, tok = training.flatten()
        if len(tok) and int(np.diff(doc_ptr).max()) > 65535:
            raise ValueError("a document is longer than 65535 tokens")
        old_n = len(self._tok)
        self._doc_ptr = np.concatenate([self._doc_ptr, doc_ptr[1:] + self._doc_ptr[-1]])
        self._tok = np.concatenate([self._tok, tok])
        for inst in training:
            self.data.append(TopicAssignment(inst, LabelSequence(np.zeros(inst.getData().getLength(), np.int32))))
        self.numTypes = self.alphabet.size()
        self.betaSum = self.beta * self.numTypes
        self._new_from = old_n
        self._dirty = True
        self._build_device_state()

    def _seed(self) -> int:
        if self.randomSeed == -1:
            import time
            self.randomSeed = int(time.time_ns() & 0x7FFFFFFF)   # Mallet: clock-seeded unless setRandomSeed
        return self.randomSeed

    def _world(self):
        if self.distributed:
            import torch.distributed as dist
            return dist.get_world_size(self.process_group), dist.get_rank(self.process_group)
        return self.numThreads, None

    def _build_device_state(self):
        for s in self._samplers:
            s.close()
        self._samplers, self._shards = [], []
        world, my_rank = self._world()
        shards = partition_by_tokens(self._doc_ptr, world)
        seed = self._seed()
        ranks = [my_rank] if my_rank is not None else list(range(world))
        if my_rank is not None:
            import torch
            devices = {my_rank: torch.cuda.current_device()}
        else:
            if self.devices is not None:
                devs = self.devices
            else:
                # shard r on GPU r; more threads than GPUs (the reference's setNumThreads(4) on a
                # smaller box): the shards share the GPUs round-robin, as the Java shim does
                n_dev = max(1, _capi.device_count())
                devs = [r % n_dev for r in range(world)]
            if len(devs) != world:
                raise ValueError("setDevices needs one device per thread/shard")
            devices = dict(zip(ranks, devs))
        # initial topics of documents that are new to the model come from the device (Philox);
        # documents already sampled keep their topics
        for r in ranks:
            sh = shards[r]
            dp, tok = shard_corpus(self._doc_ptr, self._tok, sh)
            s = _capi.Sampler(self.numTopics, max(self.numTypes, 1), self.alphaSum, self.beta, seed=seed,
                              mode=self.mode, device=devices[r], rank=r, world_size=world,
                              global_token_offset=sh.token_begin, global_doc_offset=sh.doc_begin)
            s.set_alpha(self.alpha)
            s.load_corpus(dp, tok)
            restore = getattr(self, "_restore", None)
            if restore is not None and len(restore) == world:
                s.set_state(restore[r])      # alpha, beta, seed, sweep counter, z: the chain continues
                self._samplers.append(s)
                self._shards.append(sh)
                continue
            if self._z_host is None or self._new_from == 0:
                s.init_assignments(None)
            else:
                s.init_assignments(None)
                z = s.assignments()
                keep = min(sh.token_end, self._new_from) - sh.token_begin
                if keep > 0:
                    z[:keep] = self._z_host[sh.token_begin:sh.token_begin + keep]
                s.init_assignments(z)
            s.set_sweep_counter(self._iterationsSoFar)
            self._samplers.append(s)
            self._shards.append(sh)
        if world > 1:
            self._init_communicators(my_rank)
            _capi.group_sync_counts(self._samplers)   # Mallet's sumTypeTopicCounts at start-up
            for s in self._samplers:
                s.synchronize()
        self._dirty = False
        self._z_host = None
        self._restore = None
        self._push_assignments_to_data()

    def _init_communicators(self, my_rank):
        """NCCL communicators inside the library. One process per GPU: rank 0's unique id travels
        through torch.distributed. All shards in this process: one communicator set, when every
        shard has its own GPU (several contexts on one GPU - tests - fall back to peer copies)."""
        if my_rank is not None:
            import torch.distributed as dist
            box = [_capi.nccl_unique_id() if my_rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=self.process_group)
            self._samplers[0].comm_init(box[0])
        elif len({s.device for s in self._samplers}) == len(self._samplers):
            _capi.group_comm_init(self._samplers)

    # -- estimate ---------------------------------------------------------------------------------
    def estimate(self):
        """Runs numIterations sweeps (Mallet's estimate(); declared `throws IOException` there for
        its state files — nothing here writes files)."""
        if not self.data:
            raise RuntimeError("addInstances must be called before estimate")
        if self._dirty:
            self._pull_assignments()
            self._build_device_state()
        n = self.numIterations
        world, my_rank = self._world()
        optimizing = self.optimizeInterval != 0 and n > self.burninPeriod
        if getattr(self, "_sweep_log", None) and not optimizing:
            self._logged_sweeps(n)
        elif not optimizing:
            _capi.group_sweep(self._samplers, n)
        else:
            width = int(np.diff(self._doc_ptr).max()) + 1 if len(self._doc_ptr) > 1 else 1
            for s in self._samplers:
                s.hyper_begin(width)
            # Mallet's estimate(): iteration counts from 1 on every call; statistics are collected on
            # iterations > burninPeriod that are multiples of saveSampleInterval, alpha and beta are
            # re-estimated on those that are multiples of optimizeInterval.
            iteration = 0
            while iteration < n:
                # sweeps up to the next iteration that collects or optimises go down in one call
                nxt = n
                if iteration + 1 > self.burninPeriod:
                    nxt = iteration + 1
                    while nxt < n and nxt % self.saveSampleInterval and nxt % self.optimizeInterval:
                        nxt += 1
                else:
                    nxt = min(n, self.burninPeriod)
                _capi.group_sweep(self._samplers, nxt - iteration)
                iteration = nxt
                if iteration <= self.burninPeriod:
                    continue
                if iteration % self.saveSampleInterval == 0:
                    for s in self._samplers:
                        s.hyper_collect()
                if iteration % self.optimizeInterval == 0:
                    if world > 1:
                        _capi.group_allreduce(self._samplers, 1)
                    for s in self._samplers:
                        s.optimize_alpha()
                        s.optimize_beta()
                    self.alpha = self._samplers[0].alpha()
                    self.alphaSum = float(self.alpha.sum())
                    self.beta = self._samplers[0].beta()
                    self.betaSum = self.beta * self.numTypes
            for s in self._samplers:
                s.synchronize()
        self._iterationsSoFar += n
        self._push_assignments_to_data()

    def _logged_sweeps(self, n):
        import json
        tokens = int(self._doc_ptr[-1])
        with open(self._sweep_log, "a", encoding="utf-8") as out:
            for it in range(1, n + 1):
                _capi.group_sweep(self._samplers, 1)
                st = [s.stats() for s in self._samplers]
                ms = max(x["last_sweep_ms"] for x in st)
                local = max(1, sum(x["num_tokens"] for x in st))
                rec = {"sweep": self._iterationsSoFar + it, "ms": ms, "tokens_per_s": tokens / ms * 1e3 if ms > 0 else None,
                       "tables_ms": max(x["last_tables_ms"] for x in st), "sample_ms": max(x["last_sample_ms"] for x in st),
                       "exchange_ms": max(x["last_finish_ms"] for x in st),
                       "moved_frac": sum(x["tokens_moved_last"] for x in st) / local,
                       "prior_frac": sum(x["prior_bucket_last"] for x in st) / local,
                       "mean_doc_topics": sum(x["mean_doc_topics"] * x["num_tokens"] for x in st) / local,
                       "table_refresh": st[0]["table_refresh_last"], "prior_rows_rebuilt": sum(x["rows_refreshed_last"] for x in st),
                       "shards": len(st)}
                if self._sweep_log_ll and it % self._sweep_log_ll == 0 and not self.distributed:
                    rec["ll_per_token"] = self.modelLogLikelihood() / tokens
                out.write(json.dumps(rec) + "\n")

    def _pull_assignments(self):
        """Global z (document order) of the chain as it stands, before the device state is rebuilt."""
        if not self._samplers:
            return
        local = np.concatenate([s.assignments() for s in self._samplers])
        if self.distributed:
            local = _all_gather_ragged(local, self._samplers[0].device, self.process_group)
        self._z_host = local

    def _push_assignments_to_data(self):
        """Writes z back into every TopicAssignment.topicSequence so `data` consumers
        (cmu_ron/TrainAndPredict.java:135-143) keep working."""
        for s, sh in zip(self._samplers, self._shards):
            z = s.assignments()
            dp = self._doc_ptr
            for d in range(sh.doc_begin, sh.doc_end):
                self.data[d].topicSequence = LabelSequence(z[dp[d] - sh.token_begin:dp[d + 1] - sh.token_begin])

    # -- outputs ------------------------------------------------------------------------------------
    def getTopicProbabilities(self, topics) -> np.ndarray:
        """theta = (n_dk + alpha_k) / (L_d + alphaSum). Accepts a LabelSequence (as the reference
        passes, cmu_ron/TrainAndPredict.java:143) or a document index."""
        if isinstance(topics, (int, np.integer)):
            topics = self.data[int(topics)].topicSequence
        z = topics.getFeatures()
        counts = np.bincount(z, minlength=self.numTopics).astype(np.float64)
        return (counts + self.alpha) / (len(z) + self.alpha.sum())

    def getDocumentTopics(self) -> np.ndarray:
        """All local theta rows computed on the device (D x K)."""
        return np.concatenate([s.theta() for s in self._samplers], axis=0)

    def modelLogLikelihood(self) -> float:
        doc = 0.0
        word = 0.0
        for s in self._samplers:
            d, w = s.loglik_parts()
            doc += d
            word = w
        if self.distributed:
            import torch
            import torch.distributed as dist
            t = torch.tensor([doc], dtype=torch.float64, device=torch.device("cuda", self._samplers[0].device))
            dist.all_reduce(t, group=self.process_group)
            doc = float(t.item())
        return doc + word

    def getTypeTopicCounts(self):
        """(n_wk [V x K] int32, n_k [K] int32) — the dense form of Mallet's typeTopicCounts / tokensPerTopic."""
        s = self._samplers[0]
        return s.nwk(), s.nk()

    def getInferencer(self) -> "TopicInferencer":
        return TopicInferencer(self)

    def getTopWords(self, numWords: int):
        nwk, _ = self.getTypeTopicCounts()
        out = []
        for k in range(self.numTopics):
            col = nwk[:, k]
            # Mallet sorts each topic's types by count descending (ties: by type index, IDSorter)
            order = np.lexsort((np.arange(len(col)), -col))[:numWords]
            out.append([self.alphabet.lookupObject(int(w)) for w in order if col[w] > 0])
        return out

    def printTopWords(self, file, numWords: int, useNewLines: bool):
        """`topic \\t alpha_k \\t word word ...` per topic (format witnessed by reference
        data/Topics.java:13-17,40-49)."""
        words = self.getTopWords(numWords)
        with _open_for_write(file) as out:
            for k in range(self.numTopics):
                if useNewLines:
                    out.write(f"{k}\t{_fmt(self.alpha[k])}\n")
                    for w in words[k]:
                        out.write(f"{w}\n")
                else:
                    out.write(f"{k}\t{_fmt(self.alpha[k])}\t" + " ".join(str(w) + " " for w in words[k]).rstrip(" ") + " \n")

    def printDocumentTopics(self, file, threshold: float = 0.0, max: int = -1):
        """`doc name topic proportion ...` sorted by proportion descending (format witnessed by
        reference data/Docs.java:21-26,40-52)."""
        with _open_for_write(file) as out:
            out.write("#doc source topic proportion ...\n")
            for d, ta in enumerate(self.data):
                theta = self.getTopicProbabilities(ta.topicSequence)
                order = np.lexsort((np.arange(len(theta)), -theta))
                limit = len(order) if max < 0 else min(max, len(order))
                src = ta.instance.getSource()   # Mallet 2.0.7 prints the source, "null-source" when absent
                out.write(f"{d} {src if src is not None else 'null-source'}")
                for k in order[:limit]:
                    if theta[k] <= threshold:
                        break
                    out.write(f" {int(k)} {repr(float(theta[k]))}")
                out.write(" \n")

    # -- checkpoint (Mallet: the model is Serializable; the reference writes it with
    #    ObjectOutputStream and skips training when the file exists, cmu_ron/TrainAndPredict.java:179-226)
    def write(self, file):
        """The whole model as one .npz: corpus, alphabet, configuration and, per shard, the
        library's state blob (alpha, beta, Philox seed, sweep counter, z). read() continues the
        chain where it stopped - bit-identically in DEFERRED mode."""
        import json
        if self._dirty:
            self._pull_assignments()
            self._build_device_state()
        meta = {k: getattr(self, k) for k in ("numTopics", "alphaSum", "beta", "numIterations", "burninPeriod",
                                              "optimizeInterval", "saveSampleInterval", "showTopicsInterval",
                                              "wordsPerTopic", "numThreads", "randomSeed", "mode", "_iterationsSoFar")}
        meta["names"] = [ta.instance.getName() for ta in self.data]
        meta["sources"] = [ta.instance.getSource() for ta in self.data]
        meta["alphabet"] = [str(x) for x in self.alphabet.toArray()]
        blobs = {f"state_{i}": np.frombuffer(s.get_state(), np.uint8) for i, s in enumerate(self._samplers)}
        with (open(file, "wb") if not hasattr(file, "write") else _nullctx(file)) as out:
            np.savez(out, meta=np.frombuffer(json.dumps(meta).encode(), np.uint8), doc_ptr=self._doc_ptr, tok=self._tok,
                     alpha=self.alpha, **blobs)

    @classmethod
    def read(cls, file, devices=None) -> "ParallelTopicModel":
        import json
        with np.load(file) as f:
            meta = json.loads(f["meta"].tobytes().decode())
            doc_ptr, tok, alpha = f["doc_ptr"], f["tok"], f["alpha"]
            blobs = [f[k].tobytes() for k in sorted((k for k in f.files if k.startswith("state_")), key=lambda k: int(k[6:]))]
        m = cls(meta["numTopics"], meta["alphaSum"], meta["beta"])
        for k in ("numIterations", "burninPeriod", "optimizeInterval", "saveSampleInterval", "showTopicsInterval",
                  "wordsPerTopic", "numThreads", "randomSeed", "mode", "_iterationsSoFar"):
            setattr(m, k, meta[k])
        m.alpha = np.asarray(alpha, np.float64)
        m.devices = devices
        m.alphabet = Alphabet(meta["alphabet"])
        il = InstanceList.from_arrays(doc_ptr, tok, m.alphabet, names=meta["names"])
        for inst, src in zip(il, meta["sources"]):
            inst._source = src
            m.data.append(TopicAssignment(inst, LabelSequence(np.zeros(inst.getData().getLength(), np.int32))))
        m._doc_ptr, m._tok = np.asarray(doc_ptr, np.int64), np.asarray(tok, np.int32)
        m.numTypes = m.alphabet.size()
        m.betaSum = m.beta * m.numTypes
        m._new_from = 0
        m._restore = blobs
        m._build_device_state()
        return m

    def close(self):
        for s in self._samplers:
            s.close()
        self._samplers = []


class TopicInferencer:
    """Mirror of cc.mallet.topics.TopicInferencer as the reference uses it:
    getSampledDistribution(instance, 100, 10, 10) (cmu_ron/TrainAndPredict.java:144,
    cmu/TrainAndPredict.java:114). A snapshot of the trained n_wk / n_k / alpha / beta; held-out
    documents are sampled on the GPU against the frozen counts."""

    def __init__(self, model: ParallelTopicModel):
        self._model = model
        self.numTopics = model.numTopics
        self.randomSeed = 0

    def setRandomSeed(self, seed: int):
        self.randomSeed = int(seed)

    def getSampledDistribution(self, instance: Instance, numIterations: int, thinning: int, burnIn: int) -> np.ndarray:
        return self.getSampledDistributions([instance], numIterations, thinning, burnIn)[0]

    def getSampledDistributions(self, instances, numIterations: int, thinning: int, burnIn: int) -> np.ndarray:
        """Batched form: many held-out documents in one device pass (each row = one theta)."""
        m = self._model
        docs = []
        for inst in instances:
            f = inst.getData().getFeatures()
            docs.append(f[(f >= 0) & (f < m.numTypes)])   # unknown types are dropped, as in Mallet
        lens = np.array([len(d) for d in docs], np.int64)
        doc_ptr = np.zeros(len(docs) + 1, np.int64)
        np.cumsum(lens, out=doc_ptr[1:])
        tok = np.concatenate(docs).astype(np.int32) if len(docs) and doc_ptr[-1] > 0 else np.zeros(0, np.int32)
        return m._samplers[0].infer(doc_ptr, tok, numIterations, thinning, burnIn, self.randomSeed)


def _all_gather_ragged(local: np.ndarray, device: int, group) -> np.ndarray:
    """Concatenation over ranks (rank order = document order) of per-rank int32 arrays of different
    lengths: sizes first, then one padded all_gather on the rank's device."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(local)], dtype=torch.int64, device=dev), group=group)
    sizes = [int(x.item()) for x in sizes]
    pad = max(max(sizes), 1)
    mine = torch.zeros(pad, dtype=torch.int32, device=dev)
    mine[:len(local)] = torch.from_numpy(np.ascontiguousarray(local, np.int32)).to(dev)
    parts = [torch.empty(pad, dtype=torch.int32, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return np.concatenate([p[:n].cpu().numpy() for p, n in zip(parts, sizes)])


def _nullctx(f):
    import contextlib
    return contextlib.nullcontext(f)


def _fmt(x: float) -> str:
    return f"{x:.5f}"


def _open_for_write(file):
    if hasattr(file, "write"):
        import contextlib
        return contextlib.nullcontext(file)
    return open(file, "w", encoding="utf-8")
