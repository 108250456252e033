"""Drop-in check through the reference's own driver: the stock ALGA binary vs the same sources with
src/GraphCreators/GraphCreatorPrefSuf.cpp, src/GraphCreators/GraphCreatorLI.cpp, src/IO/InputReader.cpp and
src/IO/ReadPreprocess.cpp swapped for the four files of shim/ (oracle/_ref/ALGA_gpu, built by `make -C oracle ref` in
the dev container).  Same FASTA in, --threads=1: the contigs must be identical."""
import os
import subprocess

import numpy as np
import pytest

from alga_b200 import synth

pytestmark = pytest.mark.gpu

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
STOCK, GPU = os.path.join(REF, "ALGA"), os.path.join(REF, "ALGA_gpu")


def _write_fasta(path, reads):
    nt = np.frombuffer(b"ACGT", np.uint8)
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">r%d\n" % i)
            f.write(nt[r].tobytes())
            f.write(b"\n")


def _contigs(path):
    seqs, cur = [], []
    for line in open(path):
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur))
            cur = []
        else:
            cur.append(line.strip())
    if cur:
        seqs.append("".join(cur))
    comp = str.maketrans("ACGT", "TGCA")
    return sorted(min(s, s.translate(comp)[::-1]) for s in seqs)


def _run(binary, cwd, args, env=None):
    r = subprocess.run([binary] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


@pytest.mark.skipif(not (os.path.isfile(STOCK) and os.path.isfile(GPU)), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("paired,error,extra", [(False, 0.0, []), (True, 0.0, []), (True, 0.01, ["--error_rate=0.02"])])
def test_contigs_identical_through_reference_driver(gpu, tmp_path, paired, error, extra):
    rng = np.random.default_rng(77)
    genome = synth.make_genome(60_000, rng)
    if paired:
        m1, m2 = synth.sample_paired_end(genome, 150, 40, rng, error)
        files = {"x_1.fasta": m1, "x_2.fasta": m2}
        args = ["--file1=x_1.fasta", "--file2=x_2.fasta"]
    else:
        files = {"x_1.fasta": synth.sample_single_end(genome, 100, 30, rng, error)}
        args = ["--file1=x_1.fasta"]
    outs = {}
    for name, binary in (("stock", STOCK), ("gpu", GPU)):
        d = tmp_path / name
        d.mkdir()
        for fn, reads in files.items():
            _write_fasta(d / fn, reads)
        log = _run(binary, d, args + ["--threads=1", "--output=contigs.fasta"] + extra)
        if name == "gpu":
            assert "alga_gpu:" in log, "the GPU shim did not run"
            assert "alga_gpu reader:" in log, "the GPU reader shim did not run"
            assert "alga_gpu preprocess:" in log, "the GPU preprocess shim did not run"
            if extra:
                assert "alga_gpu supplement:" in log, "the GPU supplement shim did not run"
        outs[name] = _contigs(d / "contigs.fasta")
    assert len(outs["stock"]) > 0
    assert outs["gpu"] == outs["stock"]


@pytest.mark.skipif(not (os.path.isfile(STOCK) and os.path.isfile(GPU)), reason="oracle/_ref binaries not built")
def test_contigs_identical_with_several_gpus(gpu, tmp_path):
    """ALGA_GPU_DEVICES=8: the reference's single-process driver on all GPUs of the box through alga_gpu_prefsuf_build_multi
    (equal-length reads; the second GraphCreatorPrefSuf call, on contigs of ragged lengths, runs on one GPU).  With a single
    GPU on the box the same call falls back to it."""
    rng = np.random.default_rng(78)
    genome = synth.make_genome(120_000, rng)
    m1, m2 = synth.sample_paired_end(genome, 150, 40, rng, 0.0)
    outs = {}
    for name, binary, env in (("stock", STOCK, None), ("gpu", GPU, {"ALGA_GPU_DEVICES": "8"})):
        d = tmp_path / name
        d.mkdir()
        _write_fasta(d / "x_1.fasta", m1)
        _write_fasta(d / "x_2.fasta", m2)
        log = _run(binary, d, ["--file1=x_1.fasta", "--file2=x_2.fasta", "--threads=1", "--output=contigs.fasta"], env)
        if name == "gpu":
            assert "alga_gpu:" in log, "the GPU shim did not run"
        outs[name] = _contigs(d / "contigs.fasta")
    assert len(outs["stock"]) > 0
    assert outs["gpu"] == outs["stock"]
