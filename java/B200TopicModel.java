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

  // b200lda_config: int32 x4, double x2, uint64, int32 x4, int64 x2, pointer  (88 bytes)
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

  private transient MemorySegment ctx = MemorySegment.NULL;
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
  public void setNumThreads(int n) { numThreads = Math.max(1, n); }  // = AD-LDA shards = GPUs
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
    int[] kept = topics;
    rebuildContext();
    // documents already in the model keep their chain (updateModel, cmu_ron/TrainAndPredict.java:173-177)
    try (Arena a = Arena.ofConfined()) {
      MemorySegment z = a.allocate(JAVA_INT, Math.max(1, tokens.length));
      check((int) INIT_ASSIGNMENTS.invokeExact(ctx, MemorySegment.NULL));
      if (kept.length > 0) {
        check((int) GET_ASSIGNMENTS.invokeExact(ctx, z));
        MemorySegment.copy(kept, 0, z, JAVA_INT, 0, kept.length);
        check((int) INIT_ASSIGNMENTS.invokeExact(ctx, z));
      }
    } catch (RuntimeException | Error e) {
      throw e;
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
    pullTopics();
  }

  private void rebuildContext() {
    close();
    arena = Arena.ofShared();
    try {
      MemorySegment cfg = arena.allocate(CONFIG);
      cfg.set(JAVA_INT, 0, (int) CONFIG.byteSize());
      cfg.set(JAVA_INT, 4, numTopics);
      cfg.set(JAVA_INT, 8, Math.max(1, numTypes));
      cfg.set(JAVA_INT, 12, 0 /* B200LDA_MODE_LIVE */);
      cfg.set(JAVA_DOUBLE, 16, alphaSum);
      cfg.set(JAVA_DOUBLE, 24, beta);
      if (randomSeed == -1) randomSeed = (int) (System.nanoTime() & 0x7fffffff);  // Mallet: clock seed
      cfg.set(JAVA_LONG, 32, (long) randomSeed);
      cfg.set(JAVA_INT, 40, 0);   // device
      cfg.set(JAVA_INT, 44, 0);   // rank
      cfg.set(JAVA_INT, 48, 1);   // world_size (multi-GPU: one context per device + NCCL, INTEGRATION.md)
      cfg.set(JAVA_LONG, 56, 0L);
      cfg.set(JAVA_LONG, 64, 0L);
      cfg.set(ADDRESS, 72, MemorySegment.NULL);
      MemorySegment out = arena.allocate(ADDRESS);
      check((int) CREATE.invokeExact(cfg, out));
      ctx = out.get(ADDRESS, 0);
      MemorySegment a = arena.allocate(JAVA_DOUBLE, numTopics);
      MemorySegment.copy(alpha, 0, a, JAVA_DOUBLE, 0, numTopics);
      check((int) SET_ALPHA.invokeExact(ctx, a));
      MemorySegment dp = arena.allocate(JAVA_LONG, docPtr.length);
      MemorySegment.copy(docPtr, 0, dp, JAVA_LONG, 0, docPtr.length);
      MemorySegment tw = arena.allocate(JAVA_INT, Math.max(1, tokens.length));
      MemorySegment.copy(tokens, 0, tw, JAVA_INT, 0, tokens.length);
      check((int) LOAD_CORPUS.invokeExact(ctx, (long) (docPtr.length - 1), dp, tw));
      check((int) SET_SWEEP_COUNTER.invokeExact(ctx, sweepsDone));
    } catch (RuntimeException | Error e) {
      throw e;
    } catch (Throwable t) {
      throw new RuntimeException(t);
    }
  }

  /** estimate(): numIterations sweeps on the GPU, then z is written back into every topicSequence. */
  public void estimate() throws IOException {
    try {
      if (optimizeInterval == 0 || numIterations <= burninPeriod) {
        check((int) SWEEP.invokeExact(ctx, numIterations));
      } else {
        // Mallet's schedule: statistics on iterations > burn-in that are multiples of
        // saveSampleInterval, optimizeAlpha + optimizeBeta on multiples of optimizeInterval
        int width = 1;
        for (int d = 0; d + 1 < docPtr.length; d++) width = Math.max(width, (int) (docPtr[d + 1] - docPtr[d]) + 1);
        check((int) HYPER_BEGIN.invokeExact(ctx, width));
        for (int iteration = 1; iteration <= numIterations; iteration++) {
          check((int) SWEEP.invokeExact(ctx, 1));
          if (iteration <= burninPeriod) continue;
          if (iteration % saveSampleInterval == 0) check((int) HYPER_COLLECT.invokeExact(ctx));
          if (iteration % optimizeInterval == 0) {
            check((int) OPTIMIZE_ALPHA.invokeExact(ctx));
            check((int) OPTIMIZE_BETA.invokeExact(ctx));
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
      check((int) GET_ASSIGNMENTS.invokeExact(ctx, z));
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
      MemorySegment out = a.allocate(JAVA_DOUBLE);
      check((int) LOGLIK.invokeExact(ctx, out));
      return out.get(JAVA_DOUBLE, 0);
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
    if (ctx != null && !ctx.equals(MemorySegment.NULL)) {
      try { DESTROY.invokeExact(ctx); } catch (Throwable ignored) { }
      ctx = MemorySegment.NULL;
    }
    if (arena != null) { arena.close(); arena = null; }
  }

  // Java serialisation (cmu_ron/TrainAndPredict.java:179-200 writes the model with an
  // ObjectOutputStream): the chain state travels as plain arrays; the device context is rebuilt.
  private void readObject(java.io.ObjectInputStream in) throws IOException, ClassNotFoundException {
    in.defaultReadObject();
    ctx = MemorySegment.NULL;
    if (tokens.length > 0 || docPtr.length > 1) {
      int[] kept = topics;
      rebuildContext();
      try (Arena a = Arena.ofConfined()) {
        MemorySegment z = a.allocate(JAVA_INT, Math.max(1, kept.length));
        MemorySegment.copy(kept, 0, z, JAVA_INT, 0, kept.length);
        check((int) INIT_ASSIGNMENTS.invokeExact(ctx, z));
      } catch (RuntimeException | Error e) {
        throw e;
      } catch (Throwable t) {
        throw new IOException(t);
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
