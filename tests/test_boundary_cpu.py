"""CPU: the C-ABI library loads and exports every symbol the header declares; the
host-side argument checks mirror the reference binding; the kernel's index logic is
emulated on the CPU.  No compute call needs a GPU here."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    from fftconv_b200 import _capi

    header = open(os.path.join(ROOT, "include", "fftconv_cuda.h")).read()
    declared = sorted(set(re.findall(r"\b(fftconv_cuda_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 19
    lib = _capi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(declared) == sorted(_capi.SYMBOLS)
    assert lib.fftconv_cuda_version() == b"0.2.0"


def test_segment_size_rule_matches_reference_table(built):
    import fftconv_b200 as fc
    from oracle import fftconv_oracle as O

    for k in list(range(1, 1200)) + [4095, 4096, 4097, 8192, 8193, 20000]:
        assert fc.optimal_fft_size(k) == O.get_optimal_fft_size(k), k


def test_argument_errors_mirror_the_reference_binding(built):
    """py/main.cpp:21-27,38-44: RuntimeError for bad mode / kernel longer than signal."""
    import fftconv_b200 as fc

    a, k = np.ones(5), np.ones(3)
    with pytest.raises(RuntimeError, match="Unsupported convolution mode"):
        fc.convolve(a, k, mode="valid")
    with pytest.raises(RuntimeError, match="Kernel size must be smaller than input size"):
        fc.oaconvolve(k, a)
    with pytest.raises(RuntimeError, match="Number of dimensions must be one"):
        fc.convolve(np.ones((2, 2, 2)), k)
    with pytest.raises(TypeError):
        fc.convolve(a.astype(np.int32), k.astype(np.int32))


def test_no_cpu_fallback(built):
    """Without a CUDA device the product path must fail loudly, never compute."""
    import torch

    import fftconv_b200 as fc

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="(?i)cuda"):
        fc.oaconvolve(np.ones(100, np.float32), np.ones(5, np.float32))
    with pytest.raises(RuntimeError, match="(?i)cuda"):
        fc.hilbert(np.ones(64))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fftconv_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "oracle/" not in src.replace("tests/cpu", ""), f


EMU_CASES = [
    # batch n K mode chunks slots
    (4, 65536, 245, 0, 1, 7),    # BASELINE config 2 rows, work split mid-row (warm-up path)
    (5, 10000, 245, 1, 1, 3),    # Same mode, batch not a multiple of 4
    (1, 100000, 245, 0, 6, 5),   # one long row cut into chunks (config 5 shape)
    (3, 5000, 300, 1, 2, 4),
    (2, 1804, 245, 0, 1, 1),     # exactly one full segment
    (1, 300, 245, 1, 1, 1),      # shorter than one segment
    (6, 7000, 410, 0, 3, 2),     # largest K served by N = 2048
    (2, 4000, 221, 1, 1, 1),     # smallest K served by N = 2048
    (4, 20000, 245, 0, 1, 2, 1),  # base pointers 4 bytes off 16-byte alignment (staged path with shifts)
    (4, 20000, 246, 1, 1, 2, 3),  # step not a multiple of 4: shift changes every segment
    (8, 30000, 247, 0, 2, 3, 2),
    (4, 9020, 245, 0, 1, 1),      # n an exact multiple of step: last segment is full
    (16, 65536, 245, 0, 1, 19),   # slices of 8 / 9 blocks starting in mid-row (the 8-GPU shard of config 2, scaled down)
    (1, 200000, 245, 1, 1, 29),   # one row: four virtual rows, many slices per virtual row
    (2, 50000, 300, 0, 1, 11, 2), # two rows: two virtual rows each
    (3, 30000, 245, 1, 1, 64),    # more slices than blocks
    (5, 1000, 245, 0, 1, 8),      # rows shorter than one block, more slices than rows
]
EMU_WIDE_CASES = [
    (6, 20000, 411, 0, 1, 3),     # more than 410 taps: the WIDE instantiation (tails reach register slots < 8)
    (5, 30000, 500, 1, 2, 3, 1),
    (4, 30000, 768, 0, 1, 2, 3),
    (3, 30000, 1024, 1, 1, 2),    # tail as long as the step allows
]


@pytest.mark.parametrize("case", EMU_CASES + EMU_WIDE_CASES)
def test_n2048_kernel_logic_emulated_on_cpu(built, case):
    """Runs the kernel's own __host__ __device__ phase functions thread by thread on
    the CPU (tests/cpu/emulate_oa_n2048.cpp) against direct convolution in double."""
    exe = os.path.join(ROOT, "tests", "cpu", "emulate_oa_n2048")
    r = subprocess.run([exe, *map(str, case)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (case, r.stdout, r.stderr)
    assert float(r.stdout.split()[0]) < 1e-6


EMU_F64_CASES = [
    # batch n K mode chunks slots misalign(0..1 doubles) recomp
    (4, 32768, 245, 0, 1, 7, 0, 1),   # work split mid-row (warm-up path)
    (5, 10000, 165, 1, 1, 3, 0, 1),   # config 3's kernel length on the throughput segment size, odd batch
    (1, 60000, 245, 0, 4, 5, 0, 1),   # one long row cut into chunks
    (2, 1804, 245, 0, 1, 1, 0, 0),    # exactly one full segment, full twiddle table
    (3, 7000, 410, 0, 3, 2, 1, 1),    # largest K, base pointers 8 bytes off 16-byte alignment
    (2, 4000, 8, 1, 1, 1, 1, 1),      # smallest K the staged path serves
    (3, 9001, 246, 1, 1, 2, 1, 0),    # odd step: the staging shift alternates between segments
    (4, 30000, 640, 0, 1, 3, 1, 1),   # WIDE instantiation, longest float64 kernel it serves
]


@pytest.mark.parametrize("case", EMU_F64_CASES)
def test_n2048_float64_kernel_logic_emulated_on_cpu(built, case):
    """The float64 instantiation of the same engine (two streams per group), error bound 1e-13."""
    exe = os.path.join(ROOT, "tests", "cpu", "emulate_oa_n2048")
    r = subprocess.run([exe, *map(str, case), "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (case, r.stdout, r.stderr)
    assert float(r.stdout.split()[0]) < 1e-13


EMU_W1K_CASES = [
    # batch n K mode n_warps [misalign]
    (4, 65536, 245, 0, 7),        # config 2's rows: 85 blocks each, units straddle rows
    (5, 10000, 245, 1, 3),        # odd batch (last unit half empty), Same mode
    (1, 100000, 100, 0, 5, 1),    # one row, base pointers 4 bytes off 16-byte alignment
    (3, 5000, 300, 1, 2, 3),
    (7, 1805, 8, 0, 4, 2),        # shortest kernel served
    (2, 780, 245, 0, 1),          # exactly one block of new samples
    (1, 300, 245, 1, 1),          # shorter than one block
    (6, 7000, 400, 0, 3),         # longest kernel served
    (3, 4096, 246, 0, 2),         # K - 1 not a multiple of 4: the staging shift differs between blocks
    (2, 20000, 129, 1, 64),       # more warps than units
    (64, 8192, 245, 1, 9),        # short rows (A-lines): every row's last block is clipped, all staged in bulk
    (33, 8192, 245, 0, 9),
    (40, 2048, 65, 1, 5),
    (9, 4099, 245, 0, 4),         # rows whose ends are not 16-byte aligned
    (9, 4101, 245, 1, 4, 2),
]


@pytest.mark.parametrize("case", EMU_W1K_CASES)
def test_w1k_kernel_logic_emulated_on_cpu(built, case):
    """The warp-per-FFT engine (oa_w1k_*.cuh): its own __host__ __device__ phase, staging and store
    functions lane by lane on the CPU (tests/cpu/emulate_oa_w1k.cpp) against direct convolution in double,
    NaN-poisoned buffers and guard bands around the outputs."""
    exe = os.path.join(ROOT, "tests", "cpu", "emulate_oa_w1k")
    r = subprocess.run([exe, *map(str, case)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (case, r.stdout, r.stderr)
    assert float(r.stdout.split()[0]) < 1e-6


def test_hilbert_r2c_engine_emulated_on_cpu(built):
    """The real-input Hilbert engine (hilbert_r2c_core.cuh): its own __host__ __device__ pass functions for all
    128 threads of a line, barrier interval by barrier interval (warp after warp where the kernel only has
    __syncwarp), on NaN-poisoned planes, against a float64 Hilbert transform from the definition; also that the
    plane layout has no collisions and that the third pass's thread map owns every butterfly exactly once."""
    exe = os.path.join(ROOT, "tests", "cpu", "emulate_hilbert_r2c")
    r = subprocess.run([exe, "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout + r.stderr
    assert float(r.stdout.split()[0]) < 1e-6


def test_fast_fft_engine_emulated_on_cpu(built):
    """fast_fft.cuh (N = 512 .. 16384, float and double): the kernel's own pass functions
    run for all threads pass by pass on the CPU against direct circular convolution."""
    exe = os.path.join(ROOT, "tests", "cpu", "emulate_fast_fft")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout + r.stderr
