"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol (no
compute calls — there is no GPU here), AD-LDA partitioning, the Mallet-type mirrors and the
reference's corpus reader."""
import ctypes
import gzip
import io
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "b200lda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200lda_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_symbol_the_header_declares():
    from ldagibbssampling_b200 import _capi
    lib = _capi.load_library()
    declared = _header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"libb200lda.so does not export {name}"
    bound = sorted(n for n, _, _ in _capi.SYMBOLS)
    assert bound == declared, "ctypes binding and include/b200lda.h disagree"
    assert lib.b200lda_abi_version() == 2
    assert lib.b200lda_last_error() is not None


def test_config_struct_matches_the_header_layout():
    from ldagibbssampling_b200 import _capi
    # int32 x4, double x2, uint64, int32 x4, int64 x2, pointer  (no implicit padding holes)
    assert ctypes.sizeof(_capi.Config) == 16 + 16 + 8 + 16 + 16 + 8
    assert _capi.Config.alpha_sum.offset == 16 and _capi.Config.seed.offset == 32
    assert _capi.Config.global_token_offset.offset == 56 and _capi.Config.stream.offset == 72


def _header_prototypes():
    """{name: (return kind, [argument kinds])} parsed from include/b200lda.h; kinds: ptr, i32, i64, f64, void."""
    text = open(os.path.join(ROOT, "include", "b200lda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)

    def kind(decl):
        decl = decl.strip()
        if "*" in decl:
            return "ptr"
        if re.search(r"\b(int64_t|uint64_t|size_t)\b", decl):
            return "i64"
        if re.search(r"\b(int32_t|uint32_t|int)\b", decl):
            return "i32"
        if re.search(r"\bdouble\b", decl):
            return "f64"
        if decl == "void":
            return "void"
        raise AssertionError("unclassified C type: " + decl)

    protos = {}
    for ret, name, args in re.findall(r"([A-Za-z_][\w\s\*]*?)\b(b200lda_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        args = [a for a in (x.strip() for x in args.split(",")) if a and a != "void"]
        protos[name] = (kind(ret), [kind(a) for a in args])
    return protos


def _ctypes_kind(t):
    import ctypes as C
    if t is None:
        return "void"
    if t in (C.c_int, C.c_int32, C.c_uint32):
        return "i32"
    if t in (C.c_int64, C.c_uint64, C.c_size_t, C.c_longlong, C.c_ulonglong):
        return "i64"
    if t is C.c_double:
        return "f64"
    return "ptr"  # c_void_p, c_char_p, POINTER(...)


def test_ctypes_signatures_match_the_header_prototypes():
    from ldagibbssampling_b200 import _capi
    protos = _header_prototypes()
    assert len(protos) >= 50
    for name, res, args in _capi.SYMBOLS:
        want = protos[name]
        got = (_ctypes_kind(res), [_ctypes_kind(a) for a in args])
        assert got == want, f"{name}: ctypes {got} vs header {want}"


def test_java_shim_descriptors_match_the_header_prototypes():
    """The Java shim cannot be compiled here (no JDK), so its Panama FunctionDescriptors and the
    byte offsets it writes b200lda_config at are checked against the header / the ctypes struct."""
    from ldagibbssampling_b200 import _capi
    src = open(os.path.join(ROOT, "java", "B200TopicModel.java")).read()
    protos = _header_prototypes()
    jkind = {"ADDRESS": "ptr", "JAVA_INT": "i32", "JAVA_LONG": "i64", "JAVA_DOUBLE": "f64"}
    found = re.findall(r'fn\(\s*"(b200lda_[a-z0-9_]+)"\s*,\s*FunctionDescriptor\.(of|ofVoid)\(([^)]*)\)\)', src)
    assert len(found) >= 25
    for name, form, body in found:
        kinds = [jkind[x.strip()] for x in body.split(",") if x.strip()]
        got = ("void", kinds) if form == "ofVoid" else (kinds[0], kinds[1:])
        assert name in protos, f"the shim binds {name}, which the header does not declare"
        assert got == protos[name], f"{name}: Java {got} vs header {protos[name]}"
    # cfg.set(LAYOUT, offset, value) calls in rebuildContexts, in field order
    sets = re.findall(r"cfg\.set\((JAVA_INT|JAVA_LONG|JAVA_DOUBLE|ADDRESS),\s*(\d+),", src)
    fields = [(n, getattr(_capi.Config, n).offset, _ctypes_kind(t)) for n, t in _capi.Config._fields_]
    assert len(sets) == len(fields)
    for (layout, off), (fname, foff, fkind) in zip(sets, fields):
        assert int(off) == foff and jkind[layout] == fkind, (fname, layout, off, foff, fkind)
    layout = re.search(r"StructLayout CONFIG = MemoryLayout\.structLayout\((.*?)\);", src, flags=re.S).group(1)
    names = re.findall(r'withName\("([a-z_0-9]+)"\)', layout)
    assert names == [n for n, _ in _capi.Config._fields_]


def test_no_cpu_fallback_create_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ldagibbssampling_b200 as L
    assert L.device_count() == 0
    with pytest.raises(L.B200LDAError) as e:
        L.Sampler(4, 10, 1.0, 0.1)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    # argument validation happens before any device work
    with pytest.raises(L.B200LDAError) as e:
        L.Sampler(0, 10, 1.0, 0.1)
    assert e.value.code == -1


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ldagibbssampling_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "lda_oracle.h" not in src, f


# ---- partitioning ---------------------------------------------------------------------------------

def test_partition_by_tokens_covers_and_balances():
    from ldagibbssampling_b200.partition import partition_by_tokens, shard_corpus
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 200, 5000)
    dp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok = rng.integers(0, 50, int(dp[-1])).astype(np.int32)
    for world in (1, 2, 3, 8):
        shards = partition_by_tokens(dp, world)
        assert len(shards) == world
        assert shards[0].doc_begin == 0 and shards[-1].doc_end == 5000
        for a, b in zip(shards, shards[1:]):
            assert a.doc_end == b.doc_begin and a.token_end == b.token_begin
        sizes = [s.num_tokens for s in shards]
        assert max(sizes) - min(sizes) <= 2 * lens.max()
        rebuilt = np.concatenate([shard_corpus(dp, tok, s)[1] for s in shards])
        assert np.array_equal(rebuilt, tok)
        for s in shards:
            ldp, ltok = shard_corpus(dp, tok, s)
            assert ldp[0] == 0 and ldp[-1] == len(ltok) == s.num_tokens


def test_partition_edge_cases():
    from ldagibbssampling_b200.partition import partition_by_tokens
    # more shards than documents, empty documents, empty corpus
    shards = partition_by_tokens(np.array([0, 5, 5, 9], np.int64), 6)
    assert sum(s.num_docs for s in shards) == 3 and sum(s.num_tokens for s in shards) == 9
    shards = partition_by_tokens(np.array([0], np.int64), 2)
    assert all(s.num_docs == 0 for s in shards)
    with pytest.raises(ValueError):
        partition_by_tokens(np.array([1, 2], np.int64), 2)
    with pytest.raises(ValueError):
        partition_by_tokens(np.array([0, 2], np.int64), 0)


# ---- Mallet type mirrors + the reference's importer ----------------------------------------------

def test_alphabet_and_feature_sequence_follow_mallet():
    from ldagibbssampling_b200.instances import Alphabet, FeatureSequence
    a = Alphabet()
    assert a.lookupIndex("x") == 0 and a.lookupIndex("y") == 1 and a.lookupIndex("x") == 0
    assert a.lookupIndex("zzz", False) == -1 and a.size() == 2
    assert a.lookupObject(1) == "y" and a.toArray() == ["x", "y"]
    fs = FeatureSequence(a)
    fs.add("y"); fs.add("new"); fs.add(0)
    assert fs.getFeatures().tolist() == [1, 2, 0] and fs.getLength() == 3 and a.size() == 3


def test_instance_importer_reads_the_reference_line_format(tmp_path):
    """`target \\t token \\t token ...`, tokens = [^\\t]+ lower-cased, one growing alphabet
    (reference cmu_ron/InstanceImporter.java:24-39, SFDCIterator.java:60-66,
    ron/GenerateInverseDocs.java:43-57)."""
    from ldagibbssampling_b200.instances import InstanceImporter
    text = "101\tsrc/A.java\tsrc/b.java\tsrc/A.java\n102\n103\tSRC/a.JAVA\twith space.txt\n"
    p = tmp_path / "inverse_docs.txt.gz"
    with gzip.open(p, "wt", encoding="utf-8") as f:
        f.write(text)
    il = InstanceImporter().readFile(str(p))
    assert [inst.getTarget() for inst in il] == ["101", "102", "103"]
    assert [inst.getName() for inst in il] == ["example:0", "example:1", "example:2"]
    al = il.getDataAlphabet()
    assert al.toArray() == ["src/a.java", "src/b.java", "with space.txt"]
    assert il[0].getData().getFeatures().tolist() == [0, 1, 0]
    assert il[1].getData().getLength() == 0
    assert il[2].getData().getFeatures().tolist() == [0, 2]
    doc_ptr, tok = il.flatten()
    assert doc_ptr.tolist() == [0, 3, 3, 5] and tok.tolist() == [0, 1, 0, 0, 2]
    # plain-text reader path
    il2 = InstanceImporter().readFile(io.StringIO(text))
    assert il2.flatten()[1].tolist() == tok.tolist()


def test_instance_list_from_arrays_round_trip():
    from ldagibbssampling_b200.instances import InstanceList
    dp = np.array([0, 2, 2, 6], np.int64)
    tok = np.array([3, 1, 0, 0, 2, 3], np.int32)
    il = InstanceList.from_arrays(dp, tok)
    dp2, tok2 = il.flatten()
    assert np.array_equal(dp, dp2) and np.array_equal(tok, tok2)
    assert il.getDataAlphabet().size() == 4


def test_bench_reference_arm_line_shape():
    """bench.py --impl reference runs the CPU port only and prints the contract's JSON line."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--cpu-docs", "300", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tokens/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
