"""The arithmetic of the sm_100a sinc kernel, compiled for the HOST from the same source
(csrc/sinc_core.cuh via tests/sinc_host_emulation.cu: packed operations lane by lane, rcp.approx as an IEEE
division) and checked against the float64 oracle without a GPU.  Catches indexing / table / series mistakes in the
per-output tap loop before GPU time is spent; the GPU tests then check the real thing."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    so = str(tmp_path_factory.mktemp("emu") / "libsinc_emu.so")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-w", "-o", so,
                           os.path.join(HERE, "sinc_host_emulation.cu")], stderr=subprocess.DEVNULL)
    L = ctypes.CDLL(so)
    L.sinc_emulate.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int,
                               ctypes.c_void_p]

    def run(pos, x, nt):
        pos = np.ascontiguousarray(pos, np.float64)
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(len(pos), np.float32)
        L.sinc_emulate(pos.ctypes.data, len(pos), x.ctypes.data, len(x), nt, out.ctypes.data)
        return out
    return run


def _case(sr, dur, depth, base, hop=1024):
    n = int(sr * dur)
    rng = np.random.default_rng(1234)
    t = np.arange(n) / sr
    x = (0.25 * np.sin(2 * np.pi * 1000 * t) + 0.1 * np.sin(2 * np.pi * (sr / 4.3) * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    k = int(dur * sr / hop)
    times = np.linspace(0, dur, k)
    pos = oracle.speed_to_pos_c(times * sr, base + depth * np.sin(2 * np.pi * 0.5556 * times), n)
    return x, pos


@pytest.mark.parametrize("sr,dur,nt,depth,base", [(96000, 1.5, 128, 0.01, 1.0), (96000, 1.0, 50, 0.01, 1.0), (44100, 1.5, 128, 0.3, 1.0),
                                                  (44100, 1.0, 64, 0.02, 0.6), (44100, 1.0, 33, 0.05, 1.7), (96000, 0.5, 512, 0.01, 1.0),
                                                  (48000, 0.5, 1, 0.1, 1.0), (48000, 0.5, 7, 0.1, 0.9), (48000, 0.5, 17, 0.02, 1.0)])
def test_emulated_kernel_arithmetic_matches_the_float64_oracle(emu, sr, dur, nt, depth, base):
    x, pos = _case(sr, dur, depth, base)
    ref = oracle.sinc_c(pos, x, nt, nthreads=1).astype(np.float64)
    y = emu(pos, x, nt).astype(np.float64)
    ok = ~np.isnan(y)                                  # the emulation covers interior outputs (all 2*NT taps inside the signal)
    assert ok.sum() >= len(y) - 2 * nt - 64
    d = y[ok] - ref[ok]
    assert np.linalg.norm(d) / np.linalg.norm(ref[ok]) <= 2e-7
    assert np.max(np.abs(d)) / np.max(np.abs(ref[ok])) <= 5e-7
