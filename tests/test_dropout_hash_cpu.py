"""The counter-based dropout mask of the CUDA kernels (mebt_b200/csrc/common.cuh: make_drop_key / drop_row_key / drop_pair)
checked on the CPU: the helpers are __host__ __device__, so a small nvcc-built host program evaluates them and this test
compares the result with an independent Python restatement of the documented algorithm (splitmix64 site key, lowbias32
row key, one 32-bit hash per element pair, 16-bit threshold) and checks the statistics the training path relies on."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
M32, M64 = (1 << 32) - 1, (1 << 64) - 1


def _mix32(x):
    x &= M32
    x ^= x >> 16
    x = (x * 0x7feb352d) & M32
    x ^= x >> 15
    x = (x * 0x846ca68b) & M32
    x ^= x >> 16
    return x


def _key(p, seed, site):
    s = (seed + (site + 1) * 0x9E3779B97F4A7C15) & M64
    s ^= s >> 30
    s = (s * 0xBF58476D1CE4E5B9) & M64
    s ^= s >> 27
    s = (s * 0x94D049BB133111EB) & M64
    s ^= s >> 31
    thr = min(max(int(np.rint(np.float32(p) * np.float32(65536.0))), 0), 65535)
    return s & M32, s >> 32, thr


def _mask(p, seed, site, rows, pairs):
    k0, k1, thr = _key(p, seed, site)
    out = np.zeros((rows, 2 * pairs), dtype=np.uint8)
    for r in range(rows):
        rk = _mix32(k0 ^ _mix32((r + k1) & M32))
        for c in range(pairs):
            h = _mix32((rk + c * 0x9E3779B9) & M32)
            out[r, 2 * c] = (h & 0xffff) >= thr
            out[r, 2 * c + 1] = (h >> 16) >= thr
    return out, (k0, k1, thr)


@pytest.fixture(scope="module")
def host_binary(tmp_path_factory):
    nvcc = shutil.which("nvcc")
    if nvcc is None:
        pytest.skip("nvcc not available")
    exe = tmp_path_factory.mktemp("drop") / "drop_hash_host"
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--expt-relaxed-constexpr", "-I",
                        str(REPO / "include"), "-o", str(exe),
                        str(REPO / "tests" / "host" / "drop_hash_host.cu")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _run(exe, p, seed, site, rows, pairs):
    out = subprocess.run([str(exe), repr(p), str(seed), str(site), str(rows), str(pairs)], capture_output=True, text=True,
                         check=True).stdout.splitlines()
    k0, k1, thr, inv = out[0].split()
    mask = np.array([[int(ch) for ch in line] for line in out[1:]], dtype=np.uint8)
    return mask, (int(k0), int(k1), int(thr)), float(inv)


@pytest.mark.parametrize("p,seed,site", [(0.1, 20261017, 0), (0.1, 20261017, 5), (0.5, 3, 1 << 20), (0.25, 2 ** 62 - 1, 94)])
def test_cuda_helpers_match_the_documented_algorithm(host_binary, p, seed, site):
    mask, key, inv = _run(host_binary, p, seed, site, 24, 40)
    ref, ref_key = _mask(p, seed, site, 24, 40)
    assert key == ref_key
    assert abs(inv - 65536.0 / (65536 - key[2])) < 1e-6
    assert (mask == ref).all()


def test_mask_statistics(host_binary):
    p = 0.1
    mask, key, _ = _run(host_binary, p, 7, 3, 512, 512)            # 512 x 1024 decisions
    keep = mask.astype(np.float64)
    n = keep.size
    assert abs((1 - keep.mean()) - key[2] / 65536.0) < 5 * (p * (1 - p) / n) ** 0.5
    assert np.abs((1 - keep.mean(0)) - p).max() < 6 * (p * (1 - p) / keep.shape[0]) ** 0.5      # per column
    assert np.abs((1 - keep.mean(1)) - p).max() < 6 * (p * (1 - p) / keep.shape[1]) ** 0.5      # per row
    d = 1 - keep
    for a, b in ((d[:, 0::2], d[:, 1::2]), (d[:-1], d[1:]), (d[:, :-2], d[:, 2:])):             # pair halves, rows, pairs
        c = np.corrcoef(a.ravel(), b.ravel())[0, 1]
        assert abs(c) < 6 / np.sqrt(a.size), c
    other, _, _ = _run(host_binary, p, 7, 4, 64, 512)                                           # another site
    assert abs(np.corrcoef(other.ravel().astype(float), mask[:64].ravel().astype(float))[0, 1]) < 6 / np.sqrt(other.size)
