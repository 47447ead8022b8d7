"""Host-side mirrors of the Mallet data types the reference hands to the sampler, and of the
reference's own corpus importer.

The reference builds a cc.mallet.types.InstanceList through a Pipe
(cmu_ron/InstanceImporter.java:23-75, cmu/BrokenByImporter.java:27-42) and passes it to
ParallelTopicModel.addInstances. Only what the hot path reads is mirrored: the data alphabet,
each instance's FeatureSequence (int word ids), its name/target, and the LabelSequence of topics
the model keeps per document (TopicAssignment). Names and argument meaning follow Mallet.
"""
from __future__ import annotations

import gzip
import io
import re
from typing import Iterable, List, Optional, Sequence

import numpy as np


class Alphabet:
    """cc.mallet.types.Alphabet: bijection entry <-> dense index in arrival order."""

    def __init__(self, entries: Optional[Iterable] = None):
        self._map = {}
        self._entries: List = []
        self._growth_stopped = False
        if entries is not None:
            for e in entries:
                self.lookupIndex(e, True)

    def lookupIndex(self, entry, addIfNotPresent: bool = True) -> int:
        """Index of entry, adding it when absent and allowed; -1 when absent and not added
        (Mallet's contract; the reference calls lookupIndex(f, false), cmu_ron/TrainAndPredict.java:119)."""
        idx = self._map.get(entry)
        if idx is not None:
            return idx
        if not addIfNotPresent or self._growth_stopped:
            return -1
        idx = len(self._entries)
        self._map[entry] = idx
        self._entries.append(entry)
        return idx

    def lookupObject(self, index: int):
        return self._entries[index]

    def contains(self, entry) -> bool:
        return entry in self._map

    def size(self) -> int:
        return len(self._entries)

    def toArray(self) -> list:
        return list(self._entries)

    def stopGrowth(self):
        self._growth_stopped = True

    def __len__(self):
        return len(self._entries)


class FeatureSequence:
    """cc.mallet.types.FeatureSequence: a document as a sequence of alphabet indices."""

    def __init__(self, alphabet: Alphabet, features: Optional[Sequence[int]] = None):
        self._alphabet = alphabet
        self._features: List[int] = [] if features is None else [int(f) for f in features]

    def add(self, item):
        """add(int index) or add(Object key) (the key is added to the alphabet), as in Mallet."""
        if isinstance(item, (int, np.integer)):
            self._features.append(int(item))
        else:
            self._features.append(self._alphabet.lookupIndex(item, True))

    def getFeatures(self) -> np.ndarray:
        return np.asarray(self._features, dtype=np.int32)

    def getLength(self) -> int:
        return len(self._features)

    def getIndexAtPosition(self, pos: int) -> int:
        return self._features[pos]

    def getAlphabet(self) -> Alphabet:
        return self._alphabet

    def __len__(self):
        return len(self._features)


class LabelSequence:
    """The per-document topic sequence Mallet stores in TopicAssignment.topicSequence."""

    def __init__(self, features):
        self._features = np.asarray(features, dtype=np.int32)

    def getFeatures(self) -> np.ndarray:
        return self._features

    def getLength(self) -> int:
        return len(self._features)


class Instance:
    """cc.mallet.types.Instance(data, target, name, source)."""

    def __init__(self, data, target=None, name=None, source=None):
        self._data, self._target, self._name, self._source = data, target, name, source

    def getData(self):
        return self._data

    def getTarget(self):
        return self._target

    def getName(self):
        return self._name

    def getSource(self):
        return self._source


class TopicAssignment:
    """cc.mallet.topics.TopicAssignment: public fields `instance` and `topicSequence`
    (read by the reference at cmu_ron/TrainAndPredict.java:135-143)."""

    def __init__(self, instance: Instance, topicSequence: LabelSequence):
        self.instance = instance
        self.topicSequence = topicSequence


class InstanceList(list):
    """cc.mallet.types.InstanceList: a list of Instances sharing one data alphabet."""

    def __init__(self, dataAlphabet: Optional[Alphabet] = None):
        super().__init__()
        self._alphabet = dataAlphabet if dataAlphabet is not None else Alphabet()

    def getDataAlphabet(self) -> Alphabet:
        return self._alphabet

    def getAlphabet(self) -> Alphabet:
        return self._alphabet

    def add(self, instance: Instance):
        self.append(instance)

    def flatten(self):
        """(doc_ptr int64[D+1], tok_word int32[N]) — the two arrays the C ABI takes."""
        lens = np.fromiter((inst.getData().getLength() for inst in self), dtype=np.int64, count=len(self))
        doc_ptr = np.zeros(len(self) + 1, np.int64)
        np.cumsum(lens, out=doc_ptr[1:])
        if len(self):
            tok = np.concatenate([inst.getData().getFeatures() for inst in self]).astype(np.int32, copy=False)
        else:
            tok = np.zeros(0, np.int32)
        return doc_ptr, np.ascontiguousarray(tok)

    @classmethod
    def from_arrays(cls, doc_ptr, tok_word, alphabet: Optional[Alphabet] = None, names=None):
        """Build from flattened arrays (synthetic corpora); alphabet entries default to the ids."""
        doc_ptr = np.asarray(doc_ptr, np.int64)
        tok_word = np.asarray(tok_word, np.int32)
        if alphabet is None:
            v = int(tok_word.max()) + 1 if len(tok_word) else 0
            alphabet = Alphabet(range(v))
        il = cls(alphabet)
        for d in range(len(doc_ptr) - 1):
            fs = FeatureSequence(alphabet, tok_word[doc_ptr[d]:doc_ptr[d + 1]])
            name = names[d] if names is not None else str(d)
            il.append(Instance(fs, name, name, None))
        return il


class InstanceImporter:
    """The reference's corpus reader (cmu_ron/InstanceImporter.java:23-75 + SFDCIterator.java:60-66):
    one document per line, `target \\t token \\t token ...`; tokens are the maximal runs of
    non-tab characters ([^\\t]+), lower-cased (TokenSequenceLowercase), looked up in one growing
    alphabet (TokenSequence2FeatureSequence). The line format is what
    ron/GenerateInverseDocs.java:43-57 writes (`inverse_docs.txt.gz`)."""

    _token = re.compile(r"[^\t]+")

    def __init__(self, alphabet: Optional[Alphabet] = None):
        self.alphabet = alphabet if alphabet is not None else Alphabet()

    def readFile(self, filename_or_reader) -> InstanceList:
        if isinstance(filename_or_reader, str):
            if filename_or_reader.endswith(".gz"):
                with gzip.open(filename_or_reader, "rt", encoding="utf-8") as f:
                    return self._read(f)
            with open(filename_or_reader, "r", encoding="utf-8") as f:
                return self._read(f)
        return self._read(filename_or_reader)

    def _read(self, reader: io.TextIOBase) -> InstanceList:
        il = InstanceList(self.alphabet)
        index = 0
        for line in reader:
            line = line.rstrip("\n").rstrip("\r")
            # SFDCIterator: fields = line.split("\t", 2); target = fields[0], data = the rest
            fields = line.split("\t", 1)
            target = fields[0]
            data = fields[1] if len(fields) > 1 else ""
            fs = FeatureSequence(self.alphabet)
            for m in self._token.finditer(data):
                fs.add(m.group(0).lower())
            il.append(Instance(fs, target, f"example:{index}", None))  # SFDCIterator's default URI
            index += 1
        return il
