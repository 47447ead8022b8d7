// b200lda_topic_model.hpp — C++ host-side mirror of the Mallet entry points the reference drives,
// header-only, over the C ABI of libb200lda.so (include/b200lda.h).
//
// The reference's host language is Java and this image has no JDK, so — besides the Java shim that
// ships as source (java/B200TopicModel.java) — this is the compiled-language host side: the same
// method names, argument meaning and error behaviour as the used subset of
// cc.mallet.topics.ParallelTopicModel / TopicInferencer, so that the reference's
//     ParallelTopicModel model = new ParallelTopicModel(500, 100, 1);          cmu_ron/TrainAndPredict.java:160
//     model.addInstances(training); model.setOptimizeInterval(20);             :162-163
//     model.setNumThreads(4); model.setNumIterations(10000); model.estimate(); :164-166
//     currentInferencer = model.getInferencer();                               :169
// reads the same in C++. Errors surface as exceptions (IllegalArgument -> std::invalid_argument,
// IllegalState -> std::logic_error, everything else -> std::runtime_error); nothing is swallowed and
// there is no CPU fallback.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <istream>
#include <numeric>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "b200lda.h"

namespace b200lda_host {

inline void check(int rc) {
  if (rc == B200LDA_OK) return;
  const std::string msg = std::string("b200lda: ") + b200lda_last_error();
  if (rc == B200LDA_EINVAL || rc == B200LDA_ERANGE) throw std::invalid_argument(msg);
  if (rc == B200LDA_ESTATE) throw std::logic_error(msg);
  throw std::runtime_error(msg);
}

// cc.mallet.types.Alphabet
class Alphabet {
 public:
  int lookupIndex(const std::string& entry, bool addIfNotPresent = true) {
    auto it = map_.find(entry);
    if (it != map_.end()) return it->second;
    if (!addIfNotPresent) return -1;
    const int idx = (int)entries_.size();
    map_.emplace(entry, idx);
    entries_.push_back(entry);
    return idx;
  }
  const std::string& lookupObject(int index) const { return entries_.at((size_t)index); }
  int size() const { return (int)entries_.size(); }

 private:
  std::unordered_map<std::string, int> map_;
  std::vector<std::string> entries_;
};

// cc.mallet.types.FeatureSequence / LabelSequence / Instance / InstanceList / TopicAssignment
struct FeatureSequence {
  std::vector<int32_t> features;
  int getLength() const { return (int)features.size(); }
  const std::vector<int32_t>& getFeatures() const { return features; }
};
struct LabelSequence {
  std::vector<int32_t> features;
  const std::vector<int32_t>& getFeatures() const { return features; }
};
struct Instance {
  FeatureSequence data;
  std::string target, name;
  const FeatureSequence& getData() const { return data; }
  const std::string& getTarget() const { return target; }
};
struct InstanceList {
  Alphabet* alphabet = nullptr;
  std::vector<Instance> instances;
  Alphabet* getDataAlphabet() const { return alphabet; }
};
struct TopicAssignment {
  const Instance* instance = nullptr;
  LabelSequence topicSequence;
};

// The reference's corpus reader (cmu_ron/InstanceImporter.java:24-75 + SFDCIterator.java:60-66; file
// written by ron/GenerateInverseDocs.java:43-57): one document per line, `target \t token \t token ...`,
// tokens = maximal runs of non-tab characters, lower-cased (ASCII), one growing alphabet.
inline InstanceList readInverseDocs(std::istream& in, Alphabet& alphabet) {
  InstanceList il;
  il.alphabet = &alphabet;
  std::string line;
  int index = 0;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    Instance inst;
    const size_t tab = line.find('\t');
    inst.target = line.substr(0, tab);
    inst.name = "example:" + std::to_string(index++);
    size_t pos = tab == std::string::npos ? line.size() : tab + 1;
    while (pos < line.size()) {
      size_t end = line.find('\t', pos);
      if (end == std::string::npos) end = line.size();
      if (end > pos) {
        std::string tok = line.substr(pos, end - pos);
        for (char& ch : tok)
          if (ch >= 'A' && ch <= 'Z') ch = (char)(ch - 'A' + 'a');
        inst.data.features.push_back(alphabet.lookupIndex(tok, true));
      }
      pos = end + 1;
    }
    il.instances.push_back(std::move(inst));
  }
  return il;
}

class ParallelTopicModel;

// cc.mallet.topics.TopicInferencer (getSampledDistribution, cmu_ron/TrainAndPredict.java:144)
class TopicInferencer {
 public:
  explicit TopicInferencer(ParallelTopicModel* m) : model_(m) {}
  void setRandomSeed(int seed) { seed_ = (uint64_t)seed; }
  std::vector<double> getSampledDistribution(const Instance& instance, int numIterations, int thinning, int burnIn);

 private:
  ParallelTopicModel* model_;
  uint64_t seed_ = 0;
};

class ParallelTopicModel {
 public:
  std::vector<TopicAssignment> data;  // public field, as in Mallet (cmu_ron/TrainAndPredict.java:135)
  int numTopics;
  double alphaSum, beta, betaSum = 0.0;
  std::vector<double> alpha;
  int numTypes = 0;
  int numIterations = 1000, burninPeriod = 200, optimizeInterval = 50, saveSampleInterval = 10,
      showTopicsInterval = 50, wordsPerTopic = 7, numThreads = 1, randomSeed = -1;
  int samplingMode = B200LDA_MODE_LIVE;  // B200 extension: B200LDA_MODE_DEFERRED is bit-reproducible

  /** Same argument meaning as Mallet: the second argument is alphaSum, not alpha. */
  ParallelTopicModel(int numberOfTopics, double alphaSum_, double beta_)
      : numTopics(numberOfTopics), alphaSum(alphaSum_), beta(beta_), alpha((size_t)numberOfTopics, alphaSum_ / numberOfTopics) {
    if (numberOfTopics < 1) throw std::invalid_argument("numberOfTopics must be >= 1");
  }
  explicit ParallelTopicModel(int numberOfTopics) : ParallelTopicModel(numberOfTopics, numberOfTopics, 0.01) {}
  ParallelTopicModel(const ParallelTopicModel&) = delete;
  ParallelTopicModel& operator=(const ParallelTopicModel&) = delete;
  ~ParallelTopicModel() { close(); }

  void setNumIterations(int n) { numIterations = n; }
  void setBurninPeriod(int n) { burninPeriod = n; }
  void setOptimizeInterval(int n) { optimizeInterval = n; }
  /** = AD-LDA shards = GPUs (device r mod #GPUs unless setDevices). The reference calls it AFTER
   *  addInstances (cmu_ron/TrainAndPredict.java:162-164): estimate() re-shards, keeping the chain. */
  void setNumThreads(int n) { numThreads = std::max(1, n); }
  void setRandomSeed(int seed) { randomSeed = seed; }
  void setTopicDisplay(int interval, int n) { showTopicsInterval = interval; wordsPerTopic = n; }
  void setDevices(const std::vector<int>& devices) { devices_ = devices; }
  const Alphabet* getAlphabet() const { return alphabet_; }

  /** addInstances: flatten the FeatureSequences, upload, draw the initial topics on the device.
   *  Called again (updateModel, cmu_ron/TrainAndPredict.java:173-177) it keeps the chain of the
   *  documents already in the model. */
  void addInstances(const InstanceList& training) {
    if (alphabet_ && training.alphabet != alphabet_) throw std::invalid_argument("instances must share the model's alphabet");
    alphabet_ = training.alphabet;
    std::vector<int32_t> kept = pullTopics();
    for (const Instance& inst : training.instances) {
      if (inst.data.getLength() > 65535) throw std::invalid_argument("document longer than 65535 tokens");
      tokens_.insert(tokens_.end(), inst.data.features.begin(), inst.data.features.end());
      docPtr_.push_back((int64_t)tokens_.size());
      TopicAssignment ta;
      ta.instance = &inst;
      ta.topicSequence.features.assign((size_t)inst.data.getLength(), 0);
      data.push_back(std::move(ta));
    }
    numTypes = alphabet_ ? alphabet_->size() : 0;
    betaSum = beta * numTypes;
    rebuild(kept);
  }

  /** estimate(): numIterations sweeps; Mallet's hyper-parameter schedule when optimizeInterval != 0. */
  void estimate() {
    if (ctx_.empty()) throw std::logic_error("addInstances must be called before estimate");
    if ((int)ctx_.size() != numThreads) rebuild(pullTopics());  // setNumThreads after addInstances
    const bool optimizing = optimizeInterval != 0 && numIterations > burninPeriod;
    const int n = (int)ctx_.size();
    if (!optimizing) {
      check(b200lda_group_sweep(ctx_.data(), n, numIterations));  // the library drives every shard and the exchange
    } else {
      int width = 1;
      for (size_t d = 0; d + 1 < docPtr_.size(); ++d) width = std::max(width, (int)(docPtr_[d + 1] - docPtr_[d]) + 1);
      for (auto* c : ctx_) check(b200lda_hyper_begin(c, width));
      for (int iteration = 1; iteration <= numIterations; ++iteration) {
        check(b200lda_group_sweep(ctx_.data(), n, 1));
        if (iteration <= burninPeriod) continue;
        if (iteration % saveSampleInterval == 0)
          for (auto* c : ctx_) check(b200lda_hyper_collect(c));
        if (iteration % optimizeInterval == 0) {
          if (n > 1) check(b200lda_group_allreduce(ctx_.data(), n, B200LDA_BUFFER_HYPER));
          for (auto* c : ctx_) {
            check(b200lda_optimize_alpha(c));
            check(b200lda_optimize_beta(c));
          }
          check(b200lda_get_alpha(ctx_[0], alpha.data()));
          check(b200lda_get_beta(ctx_[0], &beta));
          alphaSum = std::accumulate(alpha.begin(), alpha.end(), 0.0);
          betaSum = beta * numTypes;
        }
      }
      for (auto* c : ctx_) check(b200lda_synchronize(c));
    }
    sweepsDone_ += numIterations;
    pushTopicsToData(pullTopics());
  }

  /** Checkpoint (the reference serialises the model and skips training when the file exists,
   *  cmu_ron/TrainAndPredict.java:179-200, 215-226): configuration, corpus and one library state
   *  blob per shard (alpha, beta, Philox seed, sweep counter, z). read() needs the same
   *  InstanceList again (the reference keeps its pipe for that) and continues the chain. */
  void write(const std::string& file) {
    std::ofstream out(file, std::ios::binary);
    auto put = [&](const void* p, size_t n) { out.write(static_cast<const char*>(p), (std::streamsize)n); };
    const int32_t head[8] = {0x4c324d42, 1, numTopics, numTypes, numThreads, samplingMode, randomSeed, (int32_t)ctx_.size()};
    put(head, sizeof(head));
    put(&alphaSum, sizeof(alphaSum));
    put(&beta, sizeof(beta));
    put(&sweepsDone_, sizeof(sweepsDone_));
    for (auto* c : ctx_) {
      int64_t bytes = 0;
      check(b200lda_state_size(c, &bytes));
      std::vector<char> blob((size_t)bytes);
      check(b200lda_get_state(c, blob.data(), bytes));
      put(&bytes, sizeof(bytes));
      put(blob.data(), blob.size());
    }
    if (!out) throw std::runtime_error("cannot write " + file);
  }

  void read(const std::string& file, const InstanceList& training) {
    std::ifstream in(file, std::ios::binary);
    auto get = [&](void* p, size_t n) {
      in.read(static_cast<char*>(p), (std::streamsize)n);
      if (!in) throw std::runtime_error("truncated model file " + file);
    };
    int32_t head[8];
    get(head, sizeof(head));
    if (head[0] != 0x4c324d42 || head[1] != 1) throw std::runtime_error("not a b200lda model file: " + file);
    if (head[2] != numTopics) throw std::invalid_argument("model file has another number of topics");
    numThreads = head[4];
    samplingMode = head[5];
    randomSeed = head[6];
    get(&alphaSum, sizeof(alphaSum));
    get(&beta, sizeof(beta));
    get(&sweepsDone_, sizeof(sweepsDone_));
    std::vector<std::vector<char>> blobs((size_t)head[7]);
    for (auto& b : blobs) {
      int64_t bytes = 0;
      get(&bytes, sizeof(bytes));
      b.resize((size_t)bytes);
      get(b.data(), b.size());
    }
    restore_ = std::move(blobs);
    addInstances(training);  // shards the corpus as before and installs the blobs instead of fresh topics
    check(b200lda_get_alpha(ctx_[0], alpha.data()));
    alphaSum = std::accumulate(alpha.begin(), alpha.end(), 0.0);
  }

  /** theta_k = (n_dk + alpha_k) / (L_d + alphaSum)   (cmu_ron/TrainAndPredict.java:143) */
  std::vector<double> getTopicProbabilities(const LabelSequence& topics) const {
    std::vector<double> dist((size_t)numTopics, 0.0);
    for (int32_t t : topics.features) dist[(size_t)t] += 1.0;
    for (int k = 0; k < numTopics; ++k) dist[(size_t)k] = (dist[(size_t)k] + alpha[(size_t)k]) / ((double)topics.features.size() + alphaSum);
    return dist;
  }

  double modelLogLikelihood() {
    double doc = 0.0, word = 0.0;
    for (auto* c : ctx_) {
      double d = 0.0, w = 0.0;
      check(b200lda_loglik_parts(c, &d, &w));
      doc += d;
      word = w;
    }
    return doc + word;
  }

  TopicInferencer getInferencer() { return TopicInferencer(this); }

  /** `#doc source topic proportion ...` — the format reference data/Docs.java:40-52 parses. */
  void printDocumentTopics(const std::string& file) const {
    std::ofstream out(file);
    out << "#doc source topic proportion ...\n";
    out.precision(17);
    for (size_t d = 0; d < data.size(); ++d) {
      const std::vector<double> th = getTopicProbabilities(data[d].topicSequence);
      std::vector<int> order((size_t)numTopics);
      std::iota(order.begin(), order.end(), 0);
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return th[(size_t)a] > th[(size_t)b]; });
      out << d << " null-source";
      for (int k : order) out << ' ' << k << ' ' << th[(size_t)k];
      out << " \n";
    }
  }

  /** `topic \t alpha_k \t word word ...` — the format reference data/Topics.java:40-49 parses. */
  void printTopWords(const std::string& file, int numWords, bool useNewLines) {
    std::vector<int32_t> nwk((size_t)numTypes * (size_t)numTopics);
    check(b200lda_get_nwk(ctx_[0], nwk.data()));
    std::ofstream out(file);
    char buf[64];
    for (int k = 0; k < numTopics; ++k) {
      std::vector<int> order((size_t)numTypes);
      std::iota(order.begin(), order.end(), 0);
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        return nwk[(size_t)a * numTopics + k] > nwk[(size_t)b * numTopics + k];
      });
      std::snprintf(buf, sizeof(buf), "%.5f", alpha[(size_t)k]);
      out << k << '\t' << buf << (useNewLines ? "\n" : "\t");
      for (int i = 0; i < std::min(numWords, numTypes) && nwk[(size_t)order[(size_t)i] * numTopics + k] > 0; ++i)
        out << alphabet_->lookupObject(order[(size_t)i]) << (useNewLines ? "\n" : " ");
      if (!useNewLines) out << "\n";
    }
  }

  /** All current topic assignments, document order (the chain state a checkpoint needs). */
  std::vector<int32_t> getAssignments() { return pullTopics(); }

  void close() {
    for (auto* c : ctx_) b200lda_destroy(c);
    ctx_.clear();
  }

  // held-out inference against the trained counts (used by TopicInferencer)
  std::vector<double> infer(const std::vector<int32_t>& words, int numIterations, int thinning, int burnIn, uint64_t seed) {
    std::vector<int32_t> known;
    for (int32_t w : words)
      if (w >= 0 && w < numTypes) known.push_back(w);  // unknown types are dropped, as Mallet does
    const int64_t dp[2] = {0, (int64_t)known.size()};
    std::vector<double> theta((size_t)numTopics);
    check(b200lda_infer(ctx_.at(0), 1, dp, known.data(), numIterations, thinning, burnIn, seed, theta.data()));
    return theta;
  }

 private:
  const Alphabet* alphabet_ = nullptr;
  std::vector<int64_t> docPtr_{0};
  std::vector<int32_t> tokens_;
  std::vector<b200lda_ctx*> ctx_;
  std::vector<int64_t> shardTok_;  // token offset of every shard (+ total)
  std::vector<int64_t> shardDoc_;
  std::vector<int> devices_;
  int64_t sweepsDone_ = 0;
  std::vector<std::vector<char>> restore_;  // read(): one state blob per shard, consumed by rebuild()

  std::vector<int32_t> pullTopics() {
    std::vector<int32_t> z(tokens_.size());
    for (size_t r = 0; r < ctx_.size(); ++r)
      if (shardTok_[r + 1] > shardTok_[r]) check(b200lda_get_assignments(ctx_[r], z.data() + shardTok_[r]));
    return z;
  }

  void pushTopicsToData(const std::vector<int32_t>& z) {
    for (size_t d = 0; d < data.size(); ++d)
      data[d].topicSequence.features.assign(z.begin() + docPtr_[d], z.begin() + docPtr_[d + 1]);
  }

  // contiguous document ranges balanced by tokens (Mallet: D / numThreads documents per thread)
  void partition(int world) {
    const int64_t D = (int64_t)docPtr_.size() - 1, N = docPtr_.back();
    shardDoc_.assign(1, 0);
    for (int r = 1; r < world; ++r) {
      const int64_t target = N * r / world;
      int64_t d = std::lower_bound(docPtr_.begin(), docPtr_.end() - 1, target) - docPtr_.begin();
      d = std::min(std::max(d, shardDoc_.back()), D);
      shardDoc_.push_back(d);
    }
    shardDoc_.push_back(D);
    shardTok_.clear();
    for (int64_t d : shardDoc_) shardTok_.push_back(docPtr_[(size_t)d]);
  }

  void rebuild(const std::vector<int32_t>& kept) {
    close();
    const int world = numThreads;
    if (randomSeed == -1) randomSeed = 12345;  // Mallet seeds from the clock; a fixed default keeps runs repeatable
    partition(world);
    // default placement: shard r on GPU r; with more threads than GPUs (the reference's
    // setNumThreads(4) on a smaller box) the shards share the GPUs round-robin
    const int ndev = std::max(1, (int)b200lda_device_count());
    auto deviceOf = [&](int r) { return devices_.empty() ? r % ndev : devices_.at((size_t)r); };
    for (int r = 0; r < world; ++r) {
      b200lda_config cfg{};
      cfg.struct_size = (int32_t)sizeof(cfg);
      cfg.num_topics = numTopics;
      cfg.num_types = std::max(1, numTypes);
      cfg.mode = samplingMode;
      cfg.alpha_sum = alphaSum;
      cfg.beta = beta;
      cfg.seed = (uint64_t)randomSeed;
      cfg.device = deviceOf(r);
      cfg.rank = r;
      cfg.world_size = world;
      cfg.global_token_offset = shardTok_[(size_t)r];
      cfg.global_doc_offset = shardDoc_[(size_t)r];
      b200lda_ctx* c = nullptr;
      check(b200lda_create(&cfg, &c));
      ctx_.push_back(c);
      check(b200lda_set_alpha(c, alpha.data()));
      std::vector<int64_t> dp;
      for (int64_t d = shardDoc_[(size_t)r]; d <= shardDoc_[(size_t)r + 1]; ++d) dp.push_back(docPtr_[(size_t)d] - shardTok_[(size_t)r]);
      check(b200lda_load_corpus(c, (int64_t)dp.size() - 1, dp.data(), tokens_.data() + shardTok_[(size_t)r]));
      if (restore_.size() == (size_t)world) {
        check(b200lda_set_state(c, restore_[(size_t)r].data(), (int64_t)restore_[(size_t)r].size()));
        continue;
      }
      check(b200lda_init_assignments(c, nullptr));
      const int64_t n = shardTok_[(size_t)r + 1] - shardTok_[(size_t)r];
      const int64_t keep = std::min<int64_t>(shardTok_[(size_t)r + 1], (int64_t)kept.size()) - shardTok_[(size_t)r];
      if (keep > 0) {  // documents already sampled keep their topics; new ones keep the fresh draw
        std::vector<int32_t> z((size_t)n);
        check(b200lda_get_assignments(c, z.data()));
        std::copy(kept.begin() + shardTok_[(size_t)r], kept.begin() + shardTok_[(size_t)r] + keep, z.begin());
        check(b200lda_init_assignments(c, z.data()));
      }
      check(b200lda_set_sweep_counter(c, sweepsDone_));
    }
    restore_.clear();
    if (world > 1) {
      // one GPU per shard: NCCL communicators inside the library; several shards on one GPU
      // (tests): the group calls fall back to peer copies
      bool distinct = true;
      for (int a = 0; a < world; ++a)
        for (int b = a + 1; b < world; ++b)
          distinct = distinct && deviceOf(a) != deviceOf(b);
      if (distinct) check(b200lda_group_comm_init(ctx_.data(), world));
      // every shard counted its own documents only: sum once (sumTypeTopicCounts at start-up)
      check(b200lda_group_sync_counts(ctx_.data(), world));
      for (auto* c : ctx_) check(b200lda_synchronize(c));
    }
    pushTopicsToData(pullTopics());
  }
};

inline std::vector<double> TopicInferencer::getSampledDistribution(const Instance& instance, int numIterations,
                                                                   int thinning, int burnIn) {
  return model_->infer(instance.data.features, numIterations, thinning, burnIn, seed_++);
}

}  // namespace b200lda_host
