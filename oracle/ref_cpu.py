"""ctypes loader for oracle/_build/libbn254ref.so (ORACLE / CPU BASELINE -- test infrastructure only)."""
from __future__ import annotations

import ctypes
import os
import subprocess
import time
from ctypes import POINTER, c_int, c_size_t, c_uint64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def lib(count=False):
    name = "libbn254ref_count.so" if count else "libbn254ref.so"
    if name not in _libs:
        path = os.path.join(HERE, "_build", name)
        src = os.path.join(HERE, "bn254_ref.cpp")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
        L = ctypes.CDLL(path)
        L.ref_groth16_verify_batch.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_size_t,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_int]
        L.ref_pairing_product_batch.argtypes = [c_void_p, c_void_p, c_int, c_size_t, c_void_p, c_void_p, c_void_p, c_int]
        L.ref_groth16_synth.argtypes = [c_uint64, c_int, c_int, c_size_t, c_size_t, c_void_p, POINTER(c_size_t), c_void_p,
                                        c_void_p, c_void_p, c_int]
        L.ref_plonk_verify_batch.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_int, c_void_p,
                                             c_size_t, c_void_p, c_void_p, c_int]
        L.ref_fp_mul_count.restype = c_uint64
        _libs[name] = L
    return _libs[name]


def _p(a):
    return None if a is None else a.ctypes.data


def groth16_verify_batch(vk, proofs, inputs, threads=1, debug=False, count=False):
    """Reference-shaped Groth16Verifier::verify over a batch.  Returns (seconds, status[, L, miller, gt])."""
    L = lib(count)
    vkb = np.frombuffer(bytes(vk), dtype=np.uint8)
    proofs = np.ascontiguousarray(proofs, dtype=np.uint8)
    inputs = np.ascontiguousarray(inputs, dtype=np.uint8)
    n = proofs.shape[0]
    status = np.full(n, 255, dtype=np.uint8)
    dl = np.zeros((n, 64), np.uint8) if debug else None
    dm = np.zeros((n, 384), np.uint8) if debug else None
    dg = np.zeros((n, 384), np.uint8) if debug else None
    t0 = time.perf_counter()
    L.ref_groth16_verify_batch(_p(vkb), vkb.size, _p(proofs), proofs.shape[1], None, _p(inputs), inputs.shape[1], n,
                               _p(status), _p(dl), _p(dm), _p(dg), threads)
    dt = time.perf_counter() - t0
    return (dt, status, dl, dm, dg) if debug else (dt, status)


def pairing_product_batch(g1, g2, k, threads=1):
    L = lib()
    g1 = np.ascontiguousarray(g1, dtype=np.uint8)
    g2 = np.ascontiguousarray(g2, dtype=np.uint8)
    n = g1.size // (64 * k)
    one = np.zeros(n, np.uint8)
    ml = np.zeros((n, 384), np.uint8)
    gt = np.zeros((n, 384), np.uint8)
    t0 = time.perf_counter()
    L.ref_pairing_product_batch(_p(g1), _p(g2), k, n, _p(one), _p(ml), _p(gt), threads)
    return time.perf_counter() - t0, one, ml, gt


def groth16_synth(seed, n, n_public=2, sign_mode=0, first_index=0, threads=0):
    L = lib()
    threads = threads or (os.cpu_count() or 1)
    vk = np.zeros(4096, np.uint8)
    vk_len = c_size_t(0)
    proofs = np.zeros((n, 256), np.uint8)
    inputs = np.zeros((n, n_public, 32), np.uint8)
    expected = np.zeros(n, np.uint8)
    rc = L.ref_groth16_synth(seed, n_public, sign_mode, first_index, n, _p(vk), ctypes.byref(vk_len), _p(proofs),
                             _p(inputs), _p(expected), threads)
    assert rc == 0
    return bytes(vk[:vk_len.value]), proofs, inputs, expected


def plonk_verify_batch(vk, proofs, inputs, rnd, threads=1, lens=None, want_gt=False):
    """Reference-shaped PlonkVerifier::verify over a batch.  Returns (seconds, status[, gt])."""
    L = lib()
    vkb = np.frombuffer(bytes(vk), dtype=np.uint8)
    proofs = np.ascontiguousarray(proofs, dtype=np.uint8)
    inputs = np.ascontiguousarray(inputs, dtype=np.uint8)
    rnd = np.ascontiguousarray(rnd, dtype=np.uint8)
    n = proofs.shape[0]
    status = np.full(n, 255, dtype=np.uint8)
    gt = np.zeros((n, 384), np.uint8) if want_gt else None
    lens = None if lens is None else np.ascontiguousarray(lens, dtype=np.uint32)
    t0 = time.perf_counter()
    L.ref_plonk_verify_batch(_p(vkb), vkb.size, _p(proofs), proofs.shape[1], _p(lens), _p(inputs), inputs.shape[1], _p(rnd),
                             n, _p(status), _p(gt), threads)
    dt = time.perf_counter() - t0
    return (dt, status, gt) if want_gt else (dt, status)
