"""The C++ host layer (pairhmm/): wire format, manager conf, plugin ABI, no-CPU-fallback contract on the CPU side;
test bench, client/worker path and multi-threaded clients against oracle-minted golden folders on the GPU side."""
import os
import subprocess

import numpy as np
import pytest

from acc_genomics_b200 import batch as B
from acc_genomics_b200 import fixtures, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "pairhmm", "bin")
LIB = os.path.join(ROOT, "pairhmm", "lib")


@pytest.fixture(scope="module")
def host(built):
    if not all(os.path.exists(os.path.join(BIN, b)) for b in ("selftest", "pairhmm_host_tb", "pairhmm_test")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "pairhmm")], check=True, stdout=subprocess.DEVNULL)
    return BIN


def run(*cmd, ok=(0,)):
    p = subprocess.run(list(cmd), capture_output=True, text=True, timeout=600)
    assert p.returncode in ok, f"{cmd}: rc={p.returncode}\n{p.stdout}\n{p.stderr}"
    return p.stdout


def test_wire_format_matches_python(host, tmp_path):
    """C++ deserialize + serialize (all overloads) reproduces the bytes of the reference's format bit for bit."""
    b = synth.config(5, scale=0.0008)[0]                      # ragged read lengths
    rs, hs = B.serialize_reads(b), B.serialize_haps(b)
    (tmp_path / "r.bin").write_bytes(rs); (tmp_path / "h.bin").write_bytes(hs)
    out = run(os.path.join(host, "selftest"), "reserialize", str(tmp_path / "r.bin"), str(tmp_path / "h.bin"),
              str(tmp_path / "r.out"), str(tmp_path / "h.out"))
    assert out.startswith("ok")
    assert (tmp_path / "r.out").read_bytes() == rs and (tmp_path / "h.out").read_bytes() == hs
    # sizes of configuration 1 quoted in SURVEY.md section 8a: 65 156 B of reads, 12 932 B of haplotypes
    b1 = synth.config(1)[0]
    assert len(B.serialize_reads(b1)) == 65156 and len(B.serialize_haps(b1)) == 12932


def test_wire_format_empty_batch(host, tmp_path):
    (tmp_path / "r.bin").write_bytes(np.int32(0).tobytes()); (tmp_path / "h.bin").write_bytes(np.int32(0).tobytes())
    out = run(os.path.join(host, "selftest"), "reserialize", str(tmp_path / "r.bin"), str(tmp_path / "h.bin"),
              str(tmp_path / "r.out"), str(tmp_path / "h.out"))
    assert "ok 0 reads 0 haps" in out


def test_manager_conf_parser(host):
    out = run(os.path.join(host, "selftest"), "conf", os.path.join(ROOT, "pairhmm", "cuda.conf"))
    assert "acc id=PairHMM path=lib/libPairHMMTask.so" in out and "devices=all" in out and "slots_per_device=2" in out


def test_manager_conf_parser_reads_reference_style_conf(host, tmp_path):
    """The shape of the reference's own conf (pairhmm/xlnx.conf): several platforms, many params, verbose."""
    (tmp_path / "x.conf").write_text('verbose: 2\nplatform {\n  id: "cpu"\n}\nplatform {\n  id: "xlnx_opencl"\n  path: "/p/libxlnx.so"\n'
                                     '  cache_loc: "cpu"\n  acc {\n    id: "PairHMM"\n    path: "lib/xlnx/libPairHMMTask.so"\n'
                                     '    param {\n      key: "program_path"\n      value: "/x/pmm.xclbin"\n    }\n'
                                     '    param { key: "kernel_name[0]" value: "pmm_core_top0" }\n'
                                     '    param { key: "pmm_core_top0.num_pe" value: "80" }\n  }\n}\n')
    out = run(os.path.join(host, "selftest"), "conf", str(tmp_path / "x.conf"))
    assert "verbose=2" in out and "kernel_name[0]=pmm_core_top0" in out and "pmm_core_top0.num_pe=80" in out
    (tmp_path / "bad.conf").write_text('platform { id: "x" acc { id: "y" ')
    assert "error" in run(os.path.join(host, "selftest"), "conf", str(tmp_path / "bad.conf"), ok=(1,))


def test_plugin_abi(host):
    """libPairHMMTask.so exports create()/destroy() and the task takes the reference's three input blocks."""
    out = run(os.path.join(host, "selftest"), "plugin", os.path.join(LIB, "libPairHMMTask.so"))
    assert "ok inputs=3" in out
    syms = subprocess.run(["nm", "-D", "--defined-only", os.path.join(LIB, "libPairHMMTask.so")], capture_output=True, text=True).stdout
    assert " T create" in syms and " T destroy" in syms
    syms = subprocess.run(["nm", "-D", "--defined-only", os.path.join(LIB, "libpairhmm_host.so")], capture_output=True, text=True).stdout
    for name in ("compute_gpu", "compute_fpga", "cleanup"):
        assert any(name in ln for ln in syms.splitlines()), name


def test_no_cpu_fallback(host):
    """Without an accelerator the client must throw: the product has no CPU compute path."""
    out = run(os.path.join(host, "selftest"), "nofallback")
    assert "ok threw" in out and "no CPU fallback" in out


def test_fixture_format_roundtrip(tmp_path):
    b = synth.config(3, scale=0.004)[0]
    fixtures.write_input(str(tmp_path / "input0"), b)
    b2 = fixtures.read_input(str(tmp_path / "input0"))
    for f in ("read_off", "rs", "q", "i", "d", "c", "hap_off", "hap"):
        assert np.array_equal(getattr(b, f), getattr(b2, f)), f
    v = np.array([-1.5, -np.inf, -123.456e-7, 0.0])
    fixtures.write_output(str(tmp_path / "output0"), v)
    assert np.array_equal(fixtures.read_output(str(tmp_path / "output0")).view(np.int64), v.view(np.int64))


# ---- GPU ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def golden_folder(host, checker, tmp_path_factory):
    """input<i>/output<i> for slices of configs 1-4 and a few ragged regions of config 5, outputs by the oracle."""
    folder = tmp_path_factory.mktemp("fixtures")
    batches = [synth.config(1, scale=0.25)[0], synth.config(2, scale=0.03)[0], synth.config(3, scale=0.03)[0],
               synth.config(4, scale=0.03)[0]] + synth.config(5, scale=0.0012)
    outs = [checker.batch(b, threads=8)[1] for b in batches]
    fixtures.write_folder(str(folder), batches, outs)
    return str(folder), sum(b.num_pairs for b in batches)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["direct", "client"])
def test_host_tb_against_golden_folder(host, golden_folder, mode):
    folder, pairs = golden_folder
    cmd = [os.path.join(host, "pairhmm_host_tb")] + (["--client"] if mode == "client" else []) + ["-", folder]
    out = run(*cmd)
    assert "0 out of 7 failed test" in out
    assert f"bit-identical results: {pairs} of {pairs}" in out
    assert "recalc" in out                                     # config 3 exercises the double re-run


@pytest.mark.gpu
def test_host_tb_with_manager_conf(host, golden_folder):
    folder, pairs = golden_folder
    out = run(os.path.join(host, "pairhmm_host_tb"), "--client", os.path.join(ROOT, "pairhmm", "cuda.conf"), folder)
    assert "0 out of 7 failed test" in out and f"bit-identical results: {pairs} of {pairs}" in out


@pytest.mark.gpu
def test_client_threads_share_the_manager(host, golden_folder):
    """Four client threads, one manager: every thread gets bit-identical results and the env slots all saw tasks."""
    folder, pairs = golden_folder
    out = run(os.path.join(host, "selftest"), "threads", folder, "4")
    assert f"bit-identical {pairs} of {pairs}" in out
    used = [ln for ln in out.splitlines() if ln.startswith("env ") and not ln.endswith("tasks 0")]
    assert len(used) >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("tiles", ["1", "3", "7"])
def test_worker_pipeline_through_the_c_entry(host, checker, tiles, monkeypatch):
    """PairHMMClient + PairHMMWorker over libPairHMMTask.so behind pairhmm_worker_forward: one tile, three tiles with two
    tasks in flight, seven ragged tiles -- always the oracle's doubles, bit for bit."""
    from acc_genomics_b200 import hostlayer
    monkeypatch.setenv("PAIRHMM_WORKER_TILES", tiles)
    for b in (synth.config(3, seed=71, scale=0.1)[0], synth.config(5, seed=72, scale=0.0004)[0]):
        out, nrecal = hostlayer.worker_forward(b)
        _, want, fb = checker.batch(b, threads=8)
        assert np.array_equal(out.view(np.int64), np.asarray(want).ravel().view(np.int64))
        assert nrecal == int(fb.sum())


@pytest.mark.gpu
def test_worker_c_entry_from_several_threads(host, checker):
    """One PairHMMClient per calling thread, one manager: GATK's threading model."""
    import threading
    from acc_genomics_b200 import hostlayer
    batches = [synth.config(2, seed=80 + k, scale=0.06)[0] for k in range(4)]
    want = [np.asarray(checker.batch(b, threads=8)[1]).ravel() for b in batches]
    got = [None] * 4

    def work(k):
        for _ in range(3):
            got[k] = hostlayer.worker_forward(hostlayer.WorkerJob(batches[k]))[0].copy()
    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in th]; [t.join() for t in th]
    for k in range(4):
        assert np.array_equal(got[k].view(np.int64), want[k].view(np.int64)), k


# ---- the standalone entry: class FalconPairHMM (reference: pairhmm/xlnx/host/FalconPairHMM.h) ------------------
def test_falcon_entry_builds_and_refuses_to_run_without_a_gpu(host):
    syms = subprocess.run(["nm", "-DC", "--defined-only", os.path.join(LIB, "libfalcon_pairhmm.so")], capture_output=True, text=True).stdout
    for name in ("FalconPairHMM::computePairhmm(", "FalconPairHMM::computePairhmmFalcon(", "FalconPairHMM::get_kernel_time(",
                 "worthFPGA(", "countCell("):
        assert name in syms, name
    import torch
    if not torch.cuda.is_available():
        out = run(os.path.join(host, "pairhmm_test"), "-", "--syn", "1", ok=(2,))
        assert "no CPU fallback" in out


@pytest.mark.gpu
def test_falcon_entry_against_golden_folder(host, golden_folder):
    folder, pairs = golden_folder
    out = run(os.path.join(host, "pairhmm_test"), "-", "--real", folder)
    assert f"bit-identical results: {pairs} of {pairs}" in out and "0 results with significant error" in out


@pytest.mark.gpu
def test_falcon_entry_synthetic_cases_checked_by_the_oracle(host, checker, tmp_path):
    """The bench's own synthetic shapes (16(i+1) reads x (i+1) haplotypes, ragged lengths): invariants inside the bench,
    values against the oracle here."""
    out = run(os.path.join(host, "pairhmm_test"), "-", "--syn", "6", "--dump", str(tmp_path))
    assert "invariant failures: 0" in out
    for i in range(6):
        b = fixtures.read_input(str(tmp_path / f"input{i}"))
        assert b.num_read == 16 * (i + 1) and b.num_hap == i + 1
        want = checker.batch(b, threads=8)[1]
        got = fixtures.read_output(str(tmp_path / f"output{i}"))
        assert np.array_equal(got.view(np.int64), np.asarray(want).ravel().view(np.int64)), i
