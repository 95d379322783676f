"""Multi-GPU checks (skipped on a single-GPU box): launch tests/dist_gpu_check.py and bench.py
under torchrun with 2 ranks."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(n, script, *args, port=29517):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, *args]      # script may be "-m", then the module name
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_row_block_sharded_ice_two_gpus():
    r = _torchrun(2, os.path.join("tests", "dist_gpu_check.py"))
    assert r.returncode == 0 and "DIST_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_bench_two_gpus_chromosome_sharding():
    r = _torchrun(2, "bench.py", "--gpus", "2", "--pairs", "20000000", "--steps", "2", "--warmup", "3", port=29519)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["roofline"]["frac"] > 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_cli_under_torchrun_equals_single_process(tmp_path):
    """`python -m hichap_master_b200 matrix -N ...` under torchrun (genome-wide matrix row-block sharded with the
    in-loop NCCL allreduce, chromosomes LPT-sharded) writes the same store as one process."""
    import numpy as np
    from conftest import SMALL_GENOME
    from hichap_master_b200 import synth
    from hichap_master_b200.construction import MatrixStore
    genome = {c: l for c, l in SMALL_GENOME.items() if c != "M"}
    names = list(genome)
    gs = synth.write_genome_size(str(tmp_path / "genomeSize"), SMALL_GENOME)
    d = tmp_path / "rep0"; d.mkdir()
    c1, p1, c2, p2 = synth.genome_pairs(genome, names, 200_000, 61, trans_frac=0.2)
    with open(d / "R0_Valid.bed", "w") as fh:
        fh.writelines(synth.valid23_lines(names, c1, p1, c2, p2))
    args = ["matrix", "-N", "-b", str(d), "-gs", gs, "-wR", "100000", "-lR", "40000", "-w", str(tmp_path)]
    one = subprocess.run([sys.executable, "-m", "hichap_master_b200", *args, "-o", str(tmp_path / "one")], cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    two = _torchrun(2, "-m", "hichap_master_b200", *args, "-o", str(tmp_path / "two"), port=29523)
    assert two.returncode == 0, two.stdout[-2000:] + two.stderr[-2000:]
    a = MatrixStore.load(str(tmp_path / "one" / "Cooler" / "Merged_Multi.npz"))
    b = MatrixStore.load(str(tmp_path / "two" / "Cooler" / "Merged_Multi.npz"))
    keys = sorted(k for k in a if "|" in k and not k.startswith(("weight", "bins")))
    assert keys == sorted(k for k in b if "|" in k and not k.startswith(("weight", "bins")))
    for k in keys:
        assert all(np.array_equal(a[k][f], b[k][f]) for f in ("bin1", "bin2", "IF")), k
    for k in ("weight|100000", "weight|40000"):
        wa, wb = a[k], b[k]
        if k == "weight|40000":        # ranks hold different chromosomes: reorder by the stored chromosome list
            assert list(b["weight_chroms|40000"]) != [] and wb.size == wa.size
            sizes = {c: genome[c] // 40000 + 1 for c in genome if c != "Y"}
            order = ["1", "2", "10", "X"]
            off = dict(zip(order, np.concatenate([[0], np.cumsum([sizes[c] for c in order])])))
            pos, parts = 0, {}
            for c in b["weight_chroms|40000"]:
                parts[str(c)] = wb[pos:pos + sizes[str(c)]]; pos += sizes[str(c)]
            wb = np.concatenate([parts[c] for c in order])
        assert np.array_equal(np.isnan(wa), np.isnan(wb)), k
        ok = ~np.isnan(wa)
        assert np.max(np.abs(wa[ok] - wb[ok]) / np.abs(wa[ok])) < 1e-6, k
