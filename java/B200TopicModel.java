// B200TopicModel.java — the reference-side binding of libb200lda.so (include/b200lda.h).
//
// Drop-in for the subset of cc.mallet.topics.ParallelTopicModel that the reference calls
// (cmu_ron/TrainAndPredict.java:159-177,136-144,230-234 and cmu/TrainAndPredict.java:258-274,
// 109-114,436): in trainNewModel, `new ParallelTopicModel(500, 100, 1)` becomes
// `new B200TopicModel(500, 100, 1)` and nothing else changes.
//
// Binding: Panama FFM (java.lang.foreign, final in Java 22) — no C glue, so the .so the Python
// tests load is byte for byte what the JVM loads. NOT compiled in the build image (no JDK there);
// the symbols, argument orders and struct layout below are the ones tests/test_host_logic.py
// checks against the header.  javac --release 22 -cp mallet-2.0.7.jar B200TopicModel.java
package cmu_b200;

import java.io.File;
import java.io.IOException;
import java.io.PrintWriter;
import java.io.Serializable;
import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.Comparator;

import cc.mallet.topics.TopicAssignment;
import cc.mallet.types.Alphabet;
import cc.mallet.types.FeatureSequence;
import cc.mallet.types.Instance;
import cc.mallet.types.InstanceList;
import cc.mallet.types.LabelAlphabet;
import cc.mallet.types.LabelSequence;

import static java.lang.foreign.ValueLayout.*;

public class B200TopicModel implements Serializable, AutoCloseable {
  private static final long serialVersionUID = 1L;

  // ---- native handles ---------------------------------------------------------------------------
  private static final Linker LINKER = Linker.nativeLinker();
  private static final SymbolLookup LIB =
      SymbolLookup.libraryLookup(System.getProperty("b200lda.library", "libb200lda.so"), Arena.global());

  private static MethodHandle fn(String name, FunctionDescriptor fd) {
    return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
  }

  // b200lda_config: int32 x4, double x2, uint64, int32 x4, int64 x2, pointer  (80 bytes, no padding)
  private static final StructLayout CONFIG = MemoryLayout.structLayout(
      JAVA_INT.withName("struct_size"), JAVA_INT.withName("num_topics"), JAVA_INT.withName("num_types"),
      JAVA_INT.withName("mode"), JAVA_DOUBLE.withName("alpha_sum"), JAVA_DOUBLE.withName("beta"),
      JAVA_LONG.withName("seed"), JAVA_INT.withName("device"), JAVA_INT.withName("rank"),
      JAVA_INT.withName("world_size"), JAVA_INT.withName("table_refresh"), JAVA_LONG.withName("global_token_offset"),
      JAVA_LONG.withName("global_doc_offset"), ADDRESS.withName("stream"));

  private static final MethodHandle LAST_ERROR = fn("b200lda_last_error", FunctionDescriptor.of(ADDRESS));
  private static final MethodHandle CREATE = fn("b200lda_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle DESTROY = fn("b200lda_destroy", FunctionDescriptor.ofVoid(ADDRESS));
  private static final MethodHandle LOAD_CORPUS =
      fn("b200lda_load_corpus", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
  private static final MethodHandle INIT_ASSIGNMENTS =
      fn("b200lda_init_assignments", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle SWEEP = fn("b200lda_sweep", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
  private static final MethodHandle SWEEP_BEGIN = fn("b200lda_sweep_begin", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  private static final MethodHandle SWEEP_END = fn("b200lda_sweep_end", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  private static final MethodHandle SYNCHRONIZE = fn("b200lda_synchronize", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  private static final MethodHandle LOGLIK = fn("b200lda_loglik", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle GET_ASSIGNMENTS =
      fn("b200lda_get_assignments", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle GET_NWK = fn("b200lda_get_nwk", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle GET_NK = fn("b200lda_get_nk", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle SET_ALPHA = fn("b200lda_set_alpha", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle INFER = fn("b200lda_infer", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG,
      ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_LONG, ADDRESS));
  private static final MethodHandle HYPER_BEGIN = fn("b200lda_hyper_begin", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
  private static final MethodHandle HYPER_COLLECT = fn("b200lda_hyper_collect", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  private static final MethodHandle OPTIMIZE_ALPHA = fn("b200lda_optimize_alpha", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  private static final MethodHandle OPTIMIZE_BETA = fn("b200lda_optimize_beta", FunctionDescriptor.of(JAVA_INT, ADDRESS));
  private static final MethodHandle GET_ALPHA = fn("b200lda_get_alpha", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle GET_BETA = fn("b200lda_get_beta", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
  private static final MethodHandle SET_SWEEP_COUNTER =
      fn("b200lda_set_sweep_counter", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG));
  // setNumThreads(n) = n shards = n contexts driven by this thread through the library's group calls
  private static final MethodHandle GROUP_COMM_INIT =
      fn("b200lda_group_comm_init", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
  private static final MethodHandle GROUP_SYNC_COUNTS =
      fn("b200lda_group_sync_counts", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
  private static final MethodHandle GROUP_SWEEP =
      fn("b200lda_group_sweep", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
  private static final MethodHandle GROUP_ALLREDUCE =
      fn("b200lda_group_allreduce", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
  private static final MethodHandle LOGLIK_PARTS =
      fn("b200lda_loglik_parts", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
  private static final MethodHandle DEVICE_COUNT = fn("b200lda_device_count", FunctionDescriptor.of(JAVA_INT));

  private static void check(int rc) {
    if (rc == 0) return;
    String msg;
    try {
      msg = ((MemorySegment) LAST_ERROR.invokeExact()).reinterpret(512).getString(0);
    } catch (Throwable t) {
      msg = "?";
    }
    switch (rc) {
      case -1: case -6: throw new IllegalArgumentException("b200lda: " + msg);
      case -5: throw new IllegalStateException("b200lda: " + msg);
      case -3: throw new OutOfMemoryError("b200lda: " + msg);
      default: throw new RuntimeException("b200lda (" + rc + "): " + msg);   // ENODEV / ECUDA: no CPU fallback
    }
  }

  // ---- Mallet-visible state (same names as ParallelTopicModel) --------------------------------------
  public ArrayList<TopicAssignment> data = new ArrayList<>();   // read by cmu_ron/TrainAndPredict.java:135
  public int numTopics;
  public double alphaSum, beta, betaSum;
  public double[] alpha;
  public Alphabet alphabet;
  public LabelAlphabet topicAlphabet;
  public int numTypes;
  public int numIterations = 1000, burninPeriod = 200, optimizeInterval = 50, showTopicsInterval = 50,
      wordsPerTopic = 7, numThreads = 1, randomSeed = -1, saveSampleInterval = 10;

  private transient MemorySegment ctx = MemorySegment.NULL;   // shard 0 (the replica outputs are read from)
  private transient MemorySegment[] ctxs = {};                // one context per shard = per GPU
  private transient MemorySegment ctxArray = MemorySegment.NULL;  // the same handles as a C array for the group calls
  private transient long[] shardTok = {0}, shardDoc = {0};    // token / document offset of every shard (+ total)
  private transient Arena arena;
  private long[] docPtr = {0};
  private int[] tokens = {};
  private int[] topics = {};          // chain state, also what Java serialisation persists
  private long sweepsDone = 0;

  public B200TopicModel(int numberOfTopics) { this(numberOfTopics, numberOfTopics, 0.01); }

  /** Same argument meaning as Mallet: the second argument is alphaSum, not alpha. */
  public B200TopicModel(int numberOfTopics, double alphaSum, double beta) {
    this.numTopics = numberOfTopics;
    this.alphaSum = alphaSum;
    this.beta = beta;
    this.alpha = new double[numberOfTopics];
    Arrays.fill(alpha, alphaSum / numberOfTopics);
    this.topicAlphabet = new LabelAlphabet();
    for (int k = 0; k < numberOfTopics; k++) topicAlphabet.lookupIndex("topic" + k);
  }

  public void setNumIterations(int n) { numIterations = n; }
  public void setBurninPeriod(int n) { burninPeriod = n; }
  public void setOptimizeInterval(int n) { optimizeInterval = n; }   // optimizeAlpha / optimizeBeta every n sweeps
  /** = AD-LDA shards = GPUs (devices 0..n-1). The reference calls it after addInstances
   *  (cmu_ron/TrainAndPredict.java:164): estimate() re-shards and keeps the chain. */
  public void setNumThreads(int n) { numThreads = Math.max(1, n); }
  public void setRandomSeed(int seed) { randomSeed = seed; }
  public void setTopicDisplay(int interval, int n) { showTopicsInterval = interval; wordsPerTopic = n; }
  public Alphabet getAlphabet() { return alphabet; }
  public ArrayList<TopicAssignment> getData() { return data; }

  /** addInstances: flatten the FeatureSequences, upload, draw initial topics on the device. */
  public void addInstances(InstanceList training) {
    alphabet = training.getDataAlphabet();
    numTypes = alphabet.size();
    betaSum = beta * numTypes;
    int oldDocs = docPtr.length - 1, oldTokens = tokens.length;
    long add = 0;
    for (Instance inst : training) add += ((FeatureSequence) inst.getData()).getLength();
    docPtr = Arrays.copyOf(docPtr, oldDocs + training.size() + 1);
    tokens = Arrays.copyOf(tokens, (int) (oldTokens + add));
    int d = oldDocs, pos = oldTokens;
    for (Instance inst : training) {
      FeatureSequence fs = (FeatureSequence) inst.getData();
      if (fs.getLength() > 65535) throw new IllegalArgumentException("document longer than 65535 tokens");
      System.arraycopy(fs.getFeatures(), 0, tokens, pos, fs.getLength());
      pos += fs.getLength();
      docPtr[++d] = pos;
      data.add(new TopicAssignment(inst, new LabelSequence(topicAlphabet, new int[fs.getLength()])));
    }
    rebuildContexts(topics);   // documents already in the model keep their chain (updateModel, cmu_ron/TrainAndPredict.java:173-177)
    pullTopics();
  }

  /** Contiguous document ranges balanced by tokens, one context per range on device r. kept =
   *  topics of the documents already sampled (document order); new documents get the Philox draw. */
  private void rebuildContexts(int[] kept) {
    close();
    arena = Arena.ofShared();
    try {
      final int world = numThreads;
      final int docs = docPtr.length - 1;
      final long total = docPtr[docs];
      shardDoc = new long[world + 1];
      for (int r = 1; r < world; r++) {
        int d = Arrays.binarySearch(docPtr, 0, docs, total * r / world);
        if (d < 0) d = -d - 1;
        shardDoc[r] = Math.min(Math.max(d, shardDoc[r - 1]), docs);
      }
      shardDoc[world] = docs;
      shardTok = new long[world + 1];
      for (int r = 0; r <= world; r++) shardTok[r] = docPtr[(int) shardDoc[r]];
      if (randomSeed == -1) randomSeed = (int) (System.nanoTime() & 0x7fffffff);  // Mallet: clock seed
      final int devices = Math.max(1, (int) DEVICE_COUNT.invokeExact());
      ctxs = new MemorySegment[world];
      ctxArray = arena.allocate(ADDRESS, world);
      for (int r = 0; r < world; r++) {
        MemorySegment cfg = arena.allocate(CONFIG);
        cfg.set(JAVA_INT, 0, (int) CONFIG.byteSize());
        cfg.set(JAVA_INT, 4, numTopics);
        cfg.set(JAVA_INT, 8, Math.max(1, numTypes));
        cfg.set(JAVA_INT, 12, 0 /* B200LDA_MODE_LIVE */);
        cfg.set(JAVA_DOUBLE, 16, alphaSum);
        cfg.set(JAVA_DOUBLE, 24, beta);
        cfg.set(JAVA_LONG, 32, (long) randomSeed);
        cfg.set(JAVA_INT, 40, world <= devices ? r : r % devices);   // device
        cfg.set(JAVA_INT, 44, r);       // rank
        cfg.set(JAVA_INT, 48, world);   // world_size
        cfg.set(JAVA_INT, 52, 0);       // table_refresh: auto
        cfg.set(JAVA_LONG, 56, shardTok[r]);
        cfg.set(JAVA_LONG, 64, shardDoc[r]);
        cfg.set(ADDRESS, 72, MemorySegment.NULL);
        MemorySegment out = arena.allocate(ADDRESS);
        check((int) CREATE.invokeExact(cfg, out));
        ctxs[r] = out.get(ADDRESS, 0);
        ctxArray.setAtIndex(ADDRESS, r, ctxs[r]);
        MemorySegment a = arena.allocate(JAVA_DOUBLE, numTopics);
        MemorySegment.copy(alpha, 0, a, JAVA_DOUBLE, 0, numTopics);
        check((int) SET_ALPHA.invokeExact(ctxs[r], a));
        final int nd = (int) (shardDoc[r + 1] - shardDoc[r]);
        final int nt = (int) (shardTok[r + 1] - shardTok[r]);
        MemorySegment dp = arena.allocate(JAVA_LONG, nd + 1);
        for (int d = 0; d <= nd; d++) dp.setAtIndex(JAVA_LONG, d, docPtr[(int) shardDoc[r] + d] - shardTok[r]);
        MemorySegment tw = arena.allocate(JAVA_INT, Math.max(1, nt));
        MemorySegment.copy(tokens, (int) shardTok[r], tw, JAVA_INT, 0, nt);
        check((int) LOAD_CORPUS.invokeExact(ctxs[r], (long) nd, dp, tw));
        check((int) INIT_ASSIGNMENTS.invokeExact(ctxs[r], MemorySegment.NULL));
        final int keep = (int) (Math.min(shardTok[r + 1], (long) kept.length) - shardTok[r]);
        if (keep > 0) {
          MemorySegment z = arena.allocate(JAVA_INT, Math.max(1, nt));
          check((int) GET_ASSIGNMENTS.invokeExact(ctxs[r], z));
          MemorySegment.copy(kept, (int) shardTok[r], z, JAVA_INT, 0, keep);
          check((int) INIT_ASSIGNMENTS.invokeExact(ctxs[r], z));
        }
        check((int) SET_SWEEP_COUNTER.invokeExact(ctxs[r], sweepsDone));
      }
      ctx = ctxs[0];
      if (world > 1) {
        // one GPU per shard: NCCL communicators inside the library; fewer GPUs than shards: the
        // group calls fall back to peer copies
        if (world <= devices) check((int) GROUP_COMM_INIT.invokeExact(ctxArray, world));
        check((int) GROUP_SYNC_COUNTS.invokeExact(ctxArray, world));   // Mallet's sumTypeTopicCounts at start-up
      }
    } catch (RuntimeException | Error e) {
      throw e;
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
  }

  /** estimate(): numIterations sweeps on the GPU, then z is written back into every topicSequence. */
  public void estimate() throws IOException {
    try {
      if (ctxs.length != numThreads) rebuildContexts(topics);   // setNumThreads after addInstances
      final int world = ctxs.length;
      if (optimizeInterval == 0 || numIterations <= burninPeriod) {
        check((int) GROUP_SWEEP.invokeExact(ctxArray, world, numIterations));
      } else {
        // Mallet's schedule: statistics on iterations > burn-in that are multiples of
        // saveSampleInterval, optimizeAlpha + optimizeBeta on multiples of optimizeInterval
        int width = 1;
        for (int d = 0; d + 1 < docPtr.length; d++) width = Math.max(width, (int) (docPtr[d + 1] - docPtr[d]) + 1);
        for (MemorySegment c : ctxs) check((int) HYPER_BEGIN.invokeExact(c, width));
        for (int iteration = 1; iteration <= numIterations; iteration++) {
          check((int) GROUP_SWEEP.invokeExact(ctxArray, world, 1));
          if (iteration <= burninPeriod) continue;
          if (iteration % saveSampleInterval == 0)
            for (MemorySegment c : ctxs) check((int) HYPER_COLLECT.invokeExact(c));
          if (iteration % optimizeInterval == 0) {
            if (world > 1) check((int) GROUP_ALLREDUCE.invokeExact(ctxArray, world, 1 /* B200LDA_BUFFER_HYPER */));
            for (MemorySegment c : ctxs) {
              check((int) OPTIMIZE_ALPHA.invokeExact(c));
              check((int) OPTIMIZE_BETA.invokeExact(c));
            }
            try (Arena a = Arena.ofConfined()) {
              MemorySegment al = a.allocate(JAVA_DOUBLE, numTopics), be = a.allocate(JAVA_DOUBLE);
              check((int) GET_ALPHA.invokeExact(ctx, al));
              check((int) GET_BETA.invokeExact(ctx, be));
              alpha = al.toArray(JAVA_DOUBLE);
              alphaSum = Arrays.stream(alpha).sum();
              beta = be.get(JAVA_DOUBLE, 0);
              betaSum = beta * numTypes;
            }
          }
        }
      }
    } catch (RuntimeException | Error e) {
      throw e;
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
    sweepsDone += numIterations;
    pullTopics();
  }

  private void pullTopics() {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment z = a.allocate(JAVA_INT, Math.max(1, tokens.length));
      for (int r = 0; r < ctxs.length; r++)
        if (shardTok[r + 1] > shardTok[r])
          check((int) GET_ASSIGNMENTS.invokeExact(ctxs[r], z.asSlice(4L * shardTok[r])));
      topics = z.asSlice(0, 4L * tokens.length).toArray(JAVA_INT);
    } catch (RuntimeException | Error e) {
      throw e;
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
    for (int d = 0; d < data.size(); d++) {
      int[] dst = data.get(d).topicSequence.getFeatures();
      System.arraycopy(topics, (int) docPtr[d], dst, 0, dst.length);
    }
  }

  /** theta_k = (n_dk + alpha_k) / (L_d + alphaSum)    (cmu_ron/TrainAndPredict.java:143) */
  public double[] getTopicProbabilities(LabelSequence topicSequence) {
    double[] dist = new double[numTopics];
    int[] z = topicSequence.getFeatures();
    for (int t : z) dist[t]++;
    for (int k = 0; k < numTopics; k++) dist[k] = (dist[k] + alpha[k]) / (z.length + alphaSum);
    return dist;
  }

  public double modelLogLikelihood() {
    try (Arena a = Arena.ofConfined()) {
      MemorySegment dp = a.allocate(JAVA_DOUBLE), wp = a.allocate(JAVA_DOUBLE);
      double doc = 0.0, word = 0.0;   // every shard's document part + the (replicated) word / topic part
      for (MemorySegment c : ctxs) {
        check((int) LOGLIK_PARTS.invokeExact(c, dp, wp));
        doc += dp.get(JAVA_DOUBLE, 0);
        word = wp.get(JAVA_DOUBLE, 0);
      }
      return doc + word;
    } catch (RuntimeException | Error e) {
      throw e;
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
  }

  public B200TopicInferencer getInferencer() { return new B200TopicInferencer(this); }

  /** TopicInferencer.getSampledDistribution(instance, numIterations, thinning, burnIn). */
  double[] infer(int[] words, int numIterations, int thinning, int burnIn, long seed) {
    int n = 0;
    int[] known = new int[words.length];
    for (int w : words) if (w >= 0 && w < numTypes) known[n++] = w;   // unknown types dropped, as Mallet does
    try (Arena a = Arena.ofConfined()) {
      MemorySegment dp = a.allocate(JAVA_LONG, 2);
      dp.setAtIndex(JAVA_LONG, 0, 0L);
      dp.setAtIndex(JAVA_LONG, 1, (long) n);
      MemorySegment tw = a.allocate(JAVA_INT, Math.max(1, n));
      MemorySegment.copy(known, 0, tw, JAVA_INT, 0, n);
      MemorySegment th = a.allocate(JAVA_DOUBLE, numTopics);
      check((int) INFER.invokeExact(ctx, 1L, dp, tw, numIterations, thinning, burnIn, seed, th));
      return th.toArray(JAVA_DOUBLE);
    } catch (RuntimeException | Error e) {
      throw e;
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
  }

  /** `#doc source topic proportion ...` — the format data/Docs.java:40-52 parses. */
  public void printDocumentTopics(File f) throws IOException {
    try (PrintWriter out = new PrintWriter(f)) {
      out.print("#doc source topic proportion ...\n");
      for (int d = 0; d < data.size(); d++) {
        double[] th = getTopicProbabilities(data.get(d).topicSequence);
        Integer[] order = new Integer[numTopics];
        for (int k = 0; k < numTopics; k++) order[k] = k;
        Arrays.sort(order, Comparator.comparingDouble((Integer k) -> -th[k]).thenComparingInt(k -> k));
        Object src = data.get(d).instance.getSource();
        out.print(d + " " + (src != null ? src : "null-source") + " ");
        for (int k : order) out.print(k + " " + th[k] + " ");
        out.print(" \n");
      }
    }
  }

  /** `topic \t alpha_k \t word word ...` — the format data/Topics.java:40-49 parses. */
  public void printTopWords(File f, int numWords, boolean useNewLines) throws IOException {
    int[] nwk = new int[numTypes * numTopics];
    try (Arena a = Arena.ofConfined()) {
      MemorySegment m = a.allocate(JAVA_INT, Math.max(1, nwk.length));
      check((int) GET_NWK.invokeExact(ctx, m));
      MemorySegment.copy(m, JAVA_INT, 0, nwk, 0, nwk.length);
    } catch (RuntimeException | Error e) {
      throw e;
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
    try (PrintWriter out = new PrintWriter(f)) {
      for (int k = 0; k < numTopics; k++) {
        final int kk = k;
        Integer[] order = new Integer[numTypes];
        for (int w = 0; w < numTypes; w++) order[w] = w;
        Arrays.sort(order, Comparator.comparingInt((Integer w) -> -nwk[w * numTopics + kk]).thenComparingInt(w -> w));
        StringBuilder sb = new StringBuilder();
        sb.append(k).append('\t').append(String.format("%.5f", alpha[k])).append(useNewLines ? "\n" : "\t");
        for (int i = 0; i < Math.min(numWords, numTypes) && nwk[order[i] * numTopics + k] > 0; i++)
          sb.append(alphabet.lookupObject(order[i])).append(useNewLines ? "\n" : " ");
        out.print(sb + (useNewLines ? "" : "\n"));
      }
    }
  }

  @Override public void close() {
    if (ctxs != null)
      for (MemorySegment c : ctxs)
        if (c != null && !c.equals(MemorySegment.NULL)) {
          try { DESTROY.invokeExact(c); } catch (Throwable ignored) { }
        }
    ctxs = new MemorySegment[0];
    ctx = MemorySegment.NULL;
    ctxArray = MemorySegment.NULL;
    if (arena != null) { arena.close(); arena = null; }
  }

  // Java serialisation (cmu_ron/TrainAndPredict.java:179-200 writes the model with an
  // ObjectOutputStream): the chain state travels as plain arrays; the device context is rebuilt.
  private void readObject(java.io.ObjectInputStream in) throws IOException, ClassNotFoundException {
    in.defaultReadObject();
    ctx = MemorySegment.NULL;
    ctxs = new MemorySegment[0];
    ctxArray = MemorySegment.NULL;
    if (tokens.length > 0 || docPtr.length > 1) {
      try {
        rebuildContexts(topics);   // the chain continues from the serialised topics and sweep counter
      } catch (RuntimeException e) {
        throw new IOException(e);
      }
    }
  }

  /** Drop-in for cc.mallet.topics.TopicInferencer as used at cmu_ron/TrainAndPredict.java:144. */
  public static final class B200TopicInferencer implements Serializable {
    private static final long serialVersionUID = 1L;
    private final B200TopicModel model;
    private long seed = 0;
    B200TopicInferencer(B200TopicModel m) { model = m; }
    public void setRandomSeed(int s) { seed = s; }
    public double[] getSampledDistribution(Instance instance, int numIterations, int thinning, int burnIn) {
      FeatureSequence fs = (FeatureSequence) instance.getData();   // getFeatures() may be longer than getLength()
      return model.infer(Arrays.copyOf(fs.getFeatures(), fs.getLength()), numIterations, thinning, burnIn, seed++);
    }
  }
}
