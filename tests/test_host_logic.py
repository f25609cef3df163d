"""CPU: host-side logic of the product and the C-ABI surface (no compute calls — there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import masks as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_selection_draws_replay_the_reference_sequence(nib):
    S = 50
    u = np.arange(S)
    for mode, fn in (("window", lambda r: om.draw_window(r, u)),
                     ("subset_keep", lambda r: om.draw_subset_keep(r, u)),
                     ("mnist", lambda r: om.draw_subset_mnist(r, u)),
                     ("cifar", lambda r: om.draw_subset_cifar(r, u))):
        got = nib.draw_selections(mode, S, 64, seed=99)
        rng = om.make_rng(99)
        want = [fn(rng) for _ in range(64)]
        assert got == want, mode


def test_selection_bits_match_oracle(nib):
    rng = np.random.RandomState(1)
    for S in (9, 50, 64, 65, 130):
        sels = [list(rng.choice(S, size=min(S, 7), replace=False)) for _ in range(33)] + [[]]
        np.testing.assert_array_equal(nib.selection_bits(sels, S), om.selection_bits(sels, S))
    with pytest.raises(ValueError):
        nib.selection_bits([[50]], 50)


def test_window_never_selects_segment_zero_and_subset_never_last(nib):
    S = 50
    w = nib.draw_selections("window", S, 500, seed=1)
    assert all(0 not in s and len(s) == 20 for s in w)       # firstIndex >= 1 (imagenet :227)
    k = nib.draw_selections("subset_keep", S, 500, seed=1)
    assert all(S - 1 not in s and len(set(s)) == 20 for s in k)  # range(u[0], u[-1]) excludes the last label


def test_shard_range_covers_everything_once(nib):
    for N in (0, 1, 7, 16384, 1000):
        for world in (1, 2, 3, 8):
            seen = []
            pers = set()
            for r in range(world):
                lo, hi, per = nib.shard_range(N, r, world)
                assert 0 <= lo <= hi <= N and hi - lo <= per
                seen += list(range(lo, hi))
                pers.add(per)
            assert seen == list(range(N)) and len(pers) == 1


def test_abi_exports_every_declared_symbol(nib):
    hdr = open(os.path.join(ROOT, "include", "nib.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(nib_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 30
    lib = ctypes.CDLL(nib._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"libnib.so does not export {name}"
    assert declared == set(nib._lib.EXPORTED_SYMBOLS), declared ^ set(nib._lib.EXPORTED_SYMBOLS)
    assert lib.nib_abi_version() == nib._lib.ABI_VERSION == int(re.search(r"#define NIB_ABI_VERSION (\d+)", hdr).group(1))


def test_compute_fails_loudly_without_a_gpu(nib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = nib._lib.load()
    rc = lib.nib_score(None, 1, 10, 0, None, None, None, None, None, None)
    assert rc == -3  # NIB_ENODEVICE: no CPU fallback
    assert b"no CPU fallback" in lib.nib_last_error() or b"fallback" in lib.nib_last_error()
    with pytest.raises(Exception):
        nib.MaskSynth(np.zeros((3, 8, 8), np.float32), np.zeros((8, 8), np.int64), S=2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "network_interpretation_imagenet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    for f in ("models/resnet.py", "models/densenet.py"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert not re.search(r"^\s*(from|import)\s+oracle\b", open(p).read(), flags=re.M), f


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from network_interpretation_imagenet_b200.engine import shard_range, gather_scores
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
N = 13
rank = dist.get_rank()
lo, hi, per = shard_range(N, rank, 2)
local = torch.zeros(per, 2)
idx = torch.arange(lo, hi, dtype=torch.float32)
local[: hi - lo, 0] = idx * 0.5
local[: hi - lo, 1] = idx
table = gather_scores(local, N)
assert table.shape == (N, 2)
assert torch.equal(table[:, 1], torch.arange(N, dtype=torch.float32)), table
assert torch.equal(table[:, 0], torch.arange(N, dtype=torch.float32) * 0.5)
dist.destroy_process_group()
print("ok", rank)
"""


def test_score_gather_world_size_2_gloo(tmp_path):
    """The N>1 data path: each rank scores a contiguous slice, one all-gather rebuilds the global table."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + (os.getpid() % 2000))
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o
