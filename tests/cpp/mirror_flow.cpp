// mirror_flow.cpp — the reference's trainNewModel / predict flow (cmu_ron/TrainAndPredict.java:159-171,
// 108-156) written against the C++ mirror, used by tests/test_cpp_mirror.py: reads a corpus in the
// reference's inverse_docs line format, trains in DEFERRED mode (bit-reproducible), prints the
// facts the Python test compares with the Python mirror / the oracle.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>

#include "b200lda_topic_model.hpp"

using namespace b200lda_host;

int main(int argc, char** argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: mirror_flow <inverse_docs.txt> <K> <seed> <threads> <iterations>\n");
    return 64;
  }
  const int K = std::atoi(argv[2]), seed = std::atoi(argv[3]), threads = std::atoi(argv[4]), iters = std::atoi(argv[5]);
  try {
    Alphabet alphabet;
    std::ifstream in(argv[1]);
    if (!in) throw std::runtime_error("cannot open corpus file");
    InstanceList training = readInverseDocs(in, alphabet);

    // the reference's order (cmu_ron/TrainAndPredict.java:160-166): construct, addInstances, THEN
    // setOptimizeInterval / setNumThreads / setNumIterations, estimate
    std::unique_ptr<ParallelTopicModel> first(new ParallelTopicModel(K, 0.1 * K, 0.01));
    first->samplingMode = B200LDA_MODE_DEFERRED;
    first->setRandomSeed(seed);
    first->addInstances(training);
    first->setOptimizeInterval(0);
    first->setNumThreads(threads);
    if (threads > 1) first->setDevices(std::vector<int>((size_t)threads, 0));
    std::unique_ptr<ParallelTopicModel> resumed;
    if (argc > 6) {
      // train half, save, throw the model away, load, train the rest (reference save / load,
      // cmu_ron/TrainAndPredict.java:179-200): the chain must be the uninterrupted one
      first->setNumIterations(iters / 2);
      first->estimate();
      const std::string path = std::string(argv[6]) + "/model.bin";
      first->write(path);
      first->close();
      first.reset();
      resumed.reset(new ParallelTopicModel(K, 0.1 * K, 0.01));
      if (threads > 1) resumed->setDevices(std::vector<int>((size_t)threads, 0));
      resumed->read(path, training);
      resumed->setOptimizeInterval(0);
      resumed->setNumIterations(iters - iters / 2);
      resumed->estimate();
    } else {
      first->setNumIterations(iters);
      first->estimate();
    }
    ParallelTopicModel& model = resumed ? *resumed : *first;
    TopicInferencer inferencer = model.getInferencer();

    uint64_t h = 1469598103934665603ull;  // FNV-1a over all topic assignments
    int64_t n = 0;
    for (const TopicAssignment& ta : model.data)
      for (int32_t z : ta.topicSequence.getFeatures()) {
        h = (h ^ (uint64_t)(uint32_t)z) * 1099511628211ull;
        ++n;
      }
    std::printf("docs %zu\ntypes %d\ntokens %" PRId64 "\nzhash %" PRIu64 "\n", model.data.size(), model.numTypes, n, h);
    std::printf("ll %.17g\n", model.modelLogLikelihood());
    const std::vector<double> th = model.getTopicProbabilities(model.data[0].topicSequence);
    std::printf("theta0");
    for (double v : th) std::printf(" %.17g", v);
    std::printf("\n");
    inferencer.setRandomSeed(5);
    const std::vector<double> inf = inferencer.getSampledDistribution(training.instances[1], 100, 10, 10);
    std::printf("infer1");
    for (double v : inf) std::printf(" %.17g", v);
    std::printf("\n");
    if (argc > 6) {
      model.printDocumentTopics(std::string(argv[6]) + "/doc_topics.txt");
      model.printTopWords(std::string(argv[6]) + "/topic_words.txt", 10, false);
    }
    return 0;
  } catch (const std::invalid_argument& e) {
    std::fprintf(stderr, "IllegalArgument: %s\n", e.what());
    return 3;
  } catch (const std::logic_error& e) {
    std::fprintf(stderr, "IllegalState: %s\n", e.what());
    return 4;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "RuntimeException: %s\n", e.what());
    return 2;
  }
}
