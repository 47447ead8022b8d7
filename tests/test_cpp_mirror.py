"""The C++ host mirror (include/b200lda_topic_model.hpp) over the C ABI: compiles and links with g++
on CPU (and fails loudly without a GPU); on a B200 it must reproduce the Python mirror / oracle
bit for bit in DEFERRED mode, for one shard and for two."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ldagibbssampling_b200")
ALPHA, BETA = 0.1, 0.01


def _build(tmp_path):
    from ldagibbssampling_b200 import _capi
    _capi.load_library()  # makes sure libb200lda.so exists
    exe = str(tmp_path / "mirror_flow")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_flow.cpp"), "-o", exe,
                           "-L", LIBDIR, "-lb200lda", f"-Wl,-rpath,{LIBDIR}"])
    return exe


def _write_corpus(path, oracle, D=250, V=180):
    dp, tok = oracle.gen_corpus(D, V, 35.0, 6, 51)
    with open(path, "w") as f:
        for d in range(D):
            f.write(f"test{d}\t" + "\t".join(f"Src/File{w}.java" for w in tok[dp[d]:dp[d + 1]]) + "\n")
    return dp, tok


def test_cpp_mirror_compiles_links_and_fails_loudly_without_gpu(tmp_path, oracle):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    corpus = tmp_path / "inverse_docs.txt"
    _write_corpus(corpus, oracle, D=20)
    out = subprocess.run([exe, str(corpus), "8", "3", "1", "2"], capture_output=True, text=True)
    assert out.returncode == 2 and "no CPU fallback" in out.stderr   # ENODEV -> RuntimeException


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 2])
def test_cpp_mirror_reproduces_the_python_mirror_and_oracle(tmp_path, oracle, threads):
    from ldagibbssampling_b200.instances import InstanceImporter
    exe = _build(tmp_path)
    corpus = tmp_path / "inverse_docs.txt"
    _write_corpus(corpus, oracle)
    K, seed, iters = 12, 9, 6
    out = subprocess.run([exe, str(corpus), str(K), str(seed), str(threads), str(iters), str(tmp_path)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    facts = {ln.split(" ", 1)[0]: ln.split(" ", 1)[1] for ln in out.stdout.strip().splitlines()}
    # the same corpus through the Python importer (lower-casing, first-seen word ids) and the oracle
    il = InstanceImporter().readFile(str(corpus))
    dp, tok = il.flatten()
    V = il.getDataAlphabet().size()
    assert int(facts["docs"]) == len(il) and int(facts["types"]) == V and int(facts["tokens"]) == len(tok)
    want = oracle.spec_sweeps(dp, tok, oracle.init_z(len(tok), K, seed), V, K, ALPHA, BETA, seed, 1, iters)
    h = 1469598103934665603
    for z in want.tolist():
        h = ((h ^ z) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert int(facts["zhash"]) == h                       # whole chain, bit for bit, 1 or 2 shards
    ll = oracle.loglik(dp, tok, want, V, K, ALPHA, BETA)
    assert abs(float(facts["ll"]) - ll) <= 1e-9 * abs(ll)
    th = np.array([float(x) for x in facts["theta0"].split()])
    assert np.allclose(th, oracle.theta(want[dp[0]:dp[1]], K, ALPHA), rtol=1e-14)
    nwk, nk = oracle.count(dp, tok, want, V, K)
    inf = np.array([float(x) for x in facts["infer1"].split()])
    if threads == 1:
        assert np.allclose(inf, oracle.spec_infer(np.array([0, dp[2] - dp[1]]), tok[dp[1]:dp[2]], nwk, nk, ALPHA, BETA,
                                                 100, 10, 10, 5)[0], rtol=1e-14)
    assert abs(inf.sum() - 1) < 1e-12
    # the reference's own parsers' formats (data/Docs.java:40-52, data/Topics.java:40-49)
    doc_lines = open(tmp_path / "doc_topics.txt").read().splitlines()
    assert doc_lines[0].startswith("#doc") and len(doc_lines) == len(il) + 1
    ar = doc_lines[1].split()
    assert ar[1] == "null-source" and (len(ar) - 2) % 2 == 0
    top = open(tmp_path / "topic_words.txt").read().splitlines()
    assert len(top) == K and top[0].split("\t")[0] == "0" and float(top[0].split("\t")[1]) == pytest.approx(ALPHA)
