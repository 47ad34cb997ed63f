"""CPU-side checks: the C-ABI library loads and exports every symbol include/frcnn_b200.h declares, the ctypes
signature table matches the header, the product never imports the oracle, and the per-image sharding helpers
work under a world_size-2 gloo group.  No compute call is made (no GPU here)."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "2d_object_detection_b200")


def _declared():
    text = open(os.path.join(ROOT, "include", "frcnn_b200.h")).read()
    return re.findall(r"FRCNN_API\s+[\w\s\*]+?\b(frcnn_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol(det):
    names = _declared()
    assert len(names) >= 25 and len(set(names)) == len(names)
    lib = ctypes.CDLL(os.path.join(PKG, "libfrcnn_b200.so"))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    lib.frcnn_version.restype = ctypes.c_int
    assert lib.frcnn_version() == 100


def test_ctypes_table_matches_header(det):
    sig = importlib.import_module("2d_object_detection_b200._lib").SIGNATURES
    assert set(sig) == set(_declared())
    text = open(os.path.join(ROOT, "include", "frcnn_b200.h")).read()
    for name, (_, args) in sig.items():
        decl = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, re.S).group(1)
        n_params = 0 if decl.strip() in ("void", "") else decl.count(",") + 1
        assert n_params == len(args), f"{name}: header has {n_params} parameters, ctypes table {len(args)}"


def test_library_is_sm100a_with_vector_reductions():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(PKG, "libfrcnn_b200.so")], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_ops_fail_loudly_without_gpu(det):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        det.decode([[0.0, 0.0, 0.0, 0.0]], [[0.0, 0.0, 1.0, 1.0]])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        det.ROIPooling(7, 2)(torch.zeros(1, 4, 4, 4), torch.zeros(1, 2, 4))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "frcnn_oracle" not in src


def test_feature_grid_and_workloads(syn):
    assert syn.feature_grid((600, 1987)) == (38, 125)
    assert syn.feature_grid((375, 1242)) == (24, 78)
    assert syn.feature_grid((800, 1333)) == (50, 84)
    assert syn.WORKLOADS["dense_b16"].num_anchors == 117600
    assert syn.WORKLOADS["kitti_b8"].num_anchors == 16848


def test_shard_slices_partition_the_batch(det):
    sh = importlib.import_module("2d_object_detection_b200.sharding")
    for gb, world in ((64, 8), (8, 1), (10, 4), (3, 8)):
        seen = []
        for r in range(world):
            s = sh.shard_slice(gb, r, world)
            seen += list(range(gb))[s]
        assert seen == list(range(gb))
    with pytest.raises(ValueError):
        sh.shard_slice(8, 8, 8)


_WORKER = r"""
import importlib, os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
sh = importlib.import_module("2d_object_detection_b200.sharding")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
s = sh.shard_slice(16, rank, 2)
assert (s.start, s.stop) == ((0, 8) if rank == 0 else (8, 16))
assert sh.max_over_ranks(10.0 + rank) == 11.0          # timing = max over ranks
assert sh.sum_over_ranks(float(s.stop - s.start)) == 16.0  # units processed = sum over ranks
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_sharding(tmp_path):
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    for r, p in enumerate(procs):
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err[-2000:]
        assert f"ok {r}" in out
