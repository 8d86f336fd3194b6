"""Host-side mirror of the reference's public interface over the bn254v C ABI.

The reference (succinctlabs/snark-bn254-verifier) exposes two associated functions,

    Groth16Verifier::verify(proof, vk, public_inputs) -> Result<bool, Groth16Error>   verifier/src/lib.rs:44-49
    PlonkVerifier::verify  (proof, vk, public_inputs) -> Result<bool, PlonkError>     verifier/src/lib.rs:69-74

This module keeps those names, argument meanings and outcomes (``True``/``False`` for ``Ok(bool)``,
``Groth16Error``/``PlonkError`` for ``Err``, ``VerifierPanic`` where the reference's ``unwrap`` would
panic) and adds ``verify_batch``.  Every call goes through ``libbn254v.so`` (CUDA, sm_100a) via
ctypes; there is no CPU fallback and nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p

import numpy as np

from . import build as _build

R_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617

# ---- status codes (include/bn254v.h enum bn254v_status) ------------------------------------------
OK_TRUE, OK_FALSE = 0, 1
ERR_PREPARE_INPUTS, ERR_BSB22_MISMATCH, ERR_INVALID_WITNESS, ERR_INVERSE_NOT_FOUND = 2, 3, 4, 5
ERR_OPENING_POLY_MISMATCH, ERR_INVALID_NUMBER_OF_DIGESTS, ERR_PAIRING_CHECK_FAILED = 6, 7, 8
PANIC_FIELD_NOT_MEMBER, PANIC_NOT_ON_CURVE, PANIC_NOT_IN_SUBGROUP, PANIC_IDENTITY = 16, 17, 18, 19
PANIC_SHORT_BUFFER, PANIC_DIV_BY_ZERO, PANIC_INDEX_OUT_OF_RANGE, PANIC_VK_PARSE, STATUS_UNSET = 20, 21, 22, 23, 255
KIND_GROTH16, KIND_PLONK = 0, 1

E_BAD_ARG, E_NO_DEVICE, E_CUDA, E_VK_PARSE, E_UNSUPPORTED = -1, -2, -3, -4, -5

# every symbol include/bn254v.h (the verifier) and include/bn254v_bench.h (measurement / test support) declare;
# tests check that the library exports them all
EXPORTS = [
    "bn254v_init", "bn254v_shutdown", "bn254v_device_count", "bn254v_last_error", "bn254v_status_name",
    "bn254v_groth16_vk_load", "bn254v_plonk_vk_load", "bn254v_vk_free", "bn254v_vk_n_public",
    "bn254v_groth16_verify_batch", "bn254v_groth16_batch_all_valid", "bn254v_plonk_verify_batch",
    "bn254v_pairing_product_batch",
    "bn254v_vk_cache_get", "bn254v_vk_cache_size", "bn254v_vk_cache_clear", "bn254v_verify_many",
]
BENCH_EXPORTS = [
    "bn254v_groth16_batch_upload", "bn254v_groth16_batch_verify", "bn254v_plonk_batch_upload",
    "bn254v_plonk_batch_verify", "bn254v_pairing_batch_upload", "bn254v_pairing_batch_verify", "bn254v_batch_free",
    "bn254v_last_stage_ms", "bn254v_last_kernel_split",
    "bn254v_groth16_synth", "bn254v_pairing_synth", "bn254v_imad_peak", "bn254v_launch_count", "bn254v_agg_host_sums",
    "bn254v_chacha20_block", "bn254v_chacha20_expand",
]


class LibraryError(RuntimeError):
    """A library-level failure (bn254v_error): bad argument, no device, CUDA error, malformed VK."""

    def __init__(self, code, msg):
        super().__init__(f"bn254v error {code}: {msg}")
        self.code = code


class Groth16Error(Exception):
    """verifier/src/groth16/error.rs"""

    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind


class PlonkError(Exception):
    """verifier/src/plonk/error.rs"""

    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind


class VerifierPanic(Exception):
    """The reference would panic here (`unwrap` on a parser error, slice index, bn's affine conversion)."""

    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind


class _Debug(ctypes.Structure):
    _fields_ = [("g1_out", c_void_p), ("fr_out", c_void_p), ("miller_out", c_void_p), ("gt_out", c_void_p)]


class _Item(ctypes.Structure):  # bn254v_item
    _fields_ = [("kind", c_int), ("n_inputs", c_int), ("proof", c_void_p), ("proof_len", c_size_t),
                ("vk", c_void_p), ("vk_len", c_size_t), ("inputs_be", c_void_p)]


_lib = None


def library_path() -> str:
    return _build.LIB


def load_library():
    """dlopen libbn254v.so (building it first when stale and nvcc is present) and declare prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if _build.is_stale() and _build.find_nvcc() is not None:
        _build.build_library()
    if not os.path.exists(_build.LIB):
        raise LibraryError(E_NO_DEVICE, "libbn254v.so is missing and cannot be built; there is no CPU fallback")
    lib = ctypes.CDLL(_build.LIB)
    u8p = c_void_p
    lib.bn254v_init.argtypes = [POINTER(c_int), c_int]
    lib.bn254v_init.restype = c_int
    lib.bn254v_shutdown.restype = None
    lib.bn254v_device_count.restype = c_int
    lib.bn254v_last_error.restype = c_char_p
    lib.bn254v_status_name.argtypes = [c_int]
    lib.bn254v_status_name.restype = c_char_p
    lib.bn254v_groth16_vk_load.argtypes = [u8p, c_size_t, c_int, POINTER(c_void_p)]
    lib.bn254v_groth16_vk_load.restype = c_int
    lib.bn254v_plonk_vk_load.argtypes = [u8p, c_size_t, POINTER(c_void_p)]
    lib.bn254v_plonk_vk_load.restype = c_int
    lib.bn254v_vk_free.argtypes = [c_void_p]
    lib.bn254v_vk_free.restype = None
    lib.bn254v_vk_n_public.argtypes = [c_void_p]
    lib.bn254v_vk_n_public.restype = c_int
    lib.bn254v_groth16_verify_batch.argtypes = [c_void_p, u8p, c_size_t, c_void_p, u8p, c_int, c_size_t, u8p,
                                                POINTER(_Debug)]
    lib.bn254v_groth16_verify_batch.restype = c_int
    lib.bn254v_groth16_batch_all_valid.argtypes = [c_void_p, u8p, c_size_t, c_void_p, u8p, c_int, u8p, c_size_t, u8p, u8p]
    lib.bn254v_groth16_batch_all_valid.restype = c_int
    lib.bn254v_plonk_verify_batch.argtypes = [c_void_p, u8p, c_size_t, c_void_p, u8p, c_int, u8p, c_size_t, u8p,
                                              POINTER(_Debug)]
    lib.bn254v_plonk_verify_batch.restype = c_int
    lib.bn254v_pairing_product_batch.argtypes = [u8p, u8p, c_int, c_size_t, u8p, u8p, u8p]
    lib.bn254v_pairing_product_batch.restype = c_int
    lib.bn254v_groth16_batch_upload.argtypes = [c_void_p, u8p, c_size_t, u8p, c_int, c_size_t, POINTER(c_void_p)]
    lib.bn254v_groth16_batch_upload.restype = c_int
    lib.bn254v_groth16_batch_verify.argtypes = [c_void_p, c_void_p, u8p, POINTER(c_float)]
    lib.bn254v_groth16_batch_verify.restype = c_int
    lib.bn254v_batch_free.argtypes = [c_void_p]
    lib.bn254v_batch_free.restype = None
    lib.bn254v_groth16_synth.argtypes = [c_uint64, c_int, c_int, c_size_t, c_size_t, u8p, POINTER(c_size_t), u8p, u8p,
                                         u8p]
    lib.bn254v_groth16_synth.restype = c_int
    lib.bn254v_pairing_synth.argtypes = [c_uint64, c_int, c_size_t, c_size_t, u8p, u8p, u8p]
    lib.bn254v_pairing_synth.restype = c_int
    lib.bn254v_imad_peak.argtypes = [c_int, POINTER(c_double), POINTER(c_double), POINTER(c_float)]
    lib.bn254v_imad_peak.restype = c_int
    lib.bn254v_launch_count.restype = c_uint64
    lib.bn254v_agg_host_sums.argtypes = [u8p, u8p, c_int, c_size_t, u8p]
    lib.bn254v_agg_host_sums.restype = None
    lib.bn254v_chacha20_block.argtypes = [u8p, ctypes.c_uint32, u8p, u8p]
    lib.bn254v_chacha20_block.restype = None
    lib.bn254v_chacha20_expand.argtypes = [u8p, c_size_t, u8p]
    lib.bn254v_chacha20_expand.restype = None
    lib.bn254v_last_kernel_split.argtypes = [POINTER(c_float), POINTER(c_float)]
    lib.bn254v_last_kernel_split.restype = c_int
    lib.bn254v_last_stage_ms.argtypes = [POINTER(c_float), c_int]
    lib.bn254v_last_stage_ms.restype = c_int
    lib.bn254v_plonk_batch_upload.argtypes = [c_void_p, u8p, c_size_t, u8p, c_int, u8p, c_size_t, POINTER(c_void_p)]
    lib.bn254v_plonk_batch_upload.restype = c_int
    lib.bn254v_plonk_batch_verify.argtypes = [c_void_p, c_void_p, u8p, POINTER(c_float)]
    lib.bn254v_plonk_batch_verify.restype = c_int
    lib.bn254v_pairing_batch_upload.argtypes = [u8p, u8p, c_int, c_size_t, POINTER(c_void_p)]
    lib.bn254v_pairing_batch_upload.restype = c_int
    lib.bn254v_pairing_batch_verify.argtypes = [c_void_p, u8p, POINTER(c_float)]
    lib.bn254v_pairing_batch_verify.restype = c_int
    lib.bn254v_vk_cache_get.argtypes = [c_int, u8p, c_size_t, c_int, POINTER(c_void_p)]
    lib.bn254v_vk_cache_get.restype = c_int
    lib.bn254v_vk_cache_size.restype = c_size_t
    lib.bn254v_vk_cache_clear.restype = None
    lib.bn254v_verify_many.argtypes = [POINTER(_Item), c_size_t, c_int, u8p, u8p]
    lib.bn254v_verify_many.restype = c_int
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise LibraryError(rc, load_library().bn254v_last_error().decode(errors="replace"))


def init(devices=None):
    """bn254v_init: select the CUDA devices the batches are sharded over (None = all visible)."""
    lib = load_library()
    if devices is None:
        _check(lib.bn254v_init(None, 0))
    else:
        arr = (c_int * len(devices))(*devices)
        _check(lib.bn254v_init(arr, len(devices)))
    return lib.bn254v_device_count()


def shutdown():
    _vk_memo.clear()
    load_library().bn254v_shutdown()


def status_name(s) -> str:
    return load_library().bn254v_status_name(int(s)).decode()


def launch_count() -> int:
    return int(load_library().bn254v_launch_count())


def _ptr(a):
    """address of a numpy array / bytes-like / None"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if isinstance(a, int):
        return a
    raise TypeError(type(a))


def _as_u8(buf):
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def fr_to_be(values) -> np.ndarray:
    """ints / 32-byte strings -> (len, 32) uint8 big-endian, the layout of `inputs_be`.
    Like bn::Fr the caller's values are expected to be < r (Fr::from_slice); a larger 32-byte value is
    passed through and reported per proof as PANIC_FIELD_NOT_MEMBER."""
    out = np.zeros((len(values), 32), dtype=np.uint8)
    for i, v in enumerate(values):
        b = v if isinstance(v, (bytes, bytearray)) else int(v).to_bytes(32, "big")
        if len(b) != 32:
            raise ValueError("field element must be 32 bytes")
        out[i] = np.frombuffer(bytes(b), dtype=np.uint8)
    return out


class _VkHandle:
    """A handle owned by the library's VK cache (bn254v_vk_cache_get): never freed from here."""

    def __init__(self, ptr, kind):
        self.ptr, self.kind = ptr, kind
        self.n_public = load_library().bn254v_vk_n_public(ptr)


def _vk(kind, vk_bytes, sign_mode=0) -> _VkHandle:
    """VK handles come from the library's cache, keyed by sha256(vk) (SP1's *_vkey_hash), kind and sign_mode: the
    VK-constant device tables are built once (the reference re-parses the VK on every call)."""
    vk_bytes = bytes(vk_bytes)
    memo_key = (kind, sign_mode, vk_bytes)
    h = _vk_memo.get(memo_key)  # saves re-hashing a 34 KB PlonK VK on every call; the handles live in the C cache
    if h is not None:
        return h
    lib = load_library()
    out = c_void_p()
    buf = np.frombuffer(vk_bytes, dtype=np.uint8)
    rc = lib.bn254v_vk_cache_get(KIND_GROTH16 if kind == "groth16" else KIND_PLONK, _ptr(buf), len(vk_bytes), sign_mode,
                                 ctypes.byref(out))
    if rc == E_VK_PARSE:
        raise VerifierPanic("VK_PARSE")  # the reference unwraps the VK parser (verifier/src/lib.rs:46,71)
    _check(rc)
    h = _vk_memo[memo_key] = _VkHandle(out.value, kind)
    return h


_vk_memo: dict = {}


def vk_cache_size() -> int:
    return int(load_library().bn254v_vk_cache_size())


def vk_cache_clear():
    _vk_memo.clear()
    load_library().bn254v_vk_cache_clear()


def _pack_proofs(proofs):
    """list of byte strings (ragged allowed) or a 2-D uint8 array -> (array[n, stride], lens or None)"""
    if isinstance(proofs, np.ndarray):
        assert proofs.dtype == np.uint8 and proofs.ndim == 2
        return np.ascontiguousarray(proofs), None
    n = len(proofs)
    stride = max([len(p) for p in proofs] + [1])
    arr = np.zeros((n, stride), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint32)
    for i, p in enumerate(proofs):
        arr[i, :len(p)] = np.frombuffer(bytes(p), dtype=np.uint8)
        lens[i] = len(p)
    return arr, (None if all(l == stride for l in lens) else lens)


def _pack_inputs(public_inputs, n):
    """per-proof lists of Fr (ints / 32-byte strings) or an (n, k, 32) uint8 array -> (array, k)"""
    if isinstance(public_inputs, np.ndarray):
        assert public_inputs.dtype == np.uint8 and public_inputs.ndim == 3 and public_inputs.shape[2] == 32
        return np.ascontiguousarray(public_inputs), public_inputs.shape[1]
    ks = {len(x) for x in public_inputs}
    if len(ks) > 1:
        raise ValueError("all proofs of a batch must have the same number of public inputs")
    k = ks.pop() if ks else 0
    arr = np.zeros((n, k, 32), dtype=np.uint8)
    for i, xs in enumerate(public_inputs):
        arr[i] = fr_to_be(xs)
    return arr, k


class DebugOutputs:
    """Canonical big-endian intermediates for the parity tests (bn254v_debug)."""

    def __init__(self, n, n_g1, n_fr):
        self.g1 = np.zeros((n, n_g1, 64), dtype=np.uint8)
        self.fr = np.zeros((n, max(n_fr, 1), 32), dtype=np.uint8)
        self.miller = np.zeros((n, 384), dtype=np.uint8)
        self.gt = np.zeros((n, 384), dtype=np.uint8)
        self.c = _Debug(self.g1.ctypes.data, self.fr.ctypes.data if n_fr else None, self.miller.ctypes.data,
                        self.gt.ctypes.data)


class Groth16Verifier:
    """Groth16Verifier (verifier/src/lib.rs:41-49)."""

    sign_mode = 0  # 0: the reference's equation as written; 1: gnark's

    @classmethod
    def verify(cls, proof, vk, public_inputs) -> bool:
        """Ok(true)/Ok(false) -> bool; Err(PrepareInputsFailed) -> Groth16Error; unwrap panics -> VerifierPanic.
        A batch of one through the same CUDA path (no CPU fallback)."""
        st = cls.verify_batch([proof], vk, [list(public_inputs)])[0]
        if st == OK_TRUE:
            return True
        if st == OK_FALSE:
            return False
        if st == ERR_PREPARE_INPUTS:
            raise Groth16Error("PrepareInputsFailed")
        raise VerifierPanic(status_name(st))

    @classmethod
    def verify_batch(cls, proofs, vk, public_inputs, debug=False, out=None):
        """Status byte per proof (numpy uint8).  `proofs`: list of gnark raw proofs (>= 256 bytes used) or an
        (n, stride) uint8 array; `public_inputs`: per-proof Fr lists or an (n, k, 32) big-endian array.
        `out`: optional preallocated (n,) uint8 array (e.g. pinned memory) that receives the status bytes."""
        lib = load_library()
        h = _vk("groth16", vk, cls.sign_mode)
        arr, lens = _pack_proofs(proofs)
        n = arr.shape[0]
        inp, k = _pack_inputs(public_inputs, n)
        if out is None:
            status = np.full(n, STATUS_UNSET, dtype=np.uint8)
        else:
            assert out.dtype == np.uint8 and out.shape == (n,) and out.flags["C_CONTIGUOUS"]
            status = out
        dbg = DebugOutputs(n, 1, 0) if debug else None
        _check(lib.bn254v_groth16_verify_batch(h.ptr, _ptr(arr), arr.shape[1], _ptr(lens), _ptr(inp), k, n,
                                               _ptr(status), ctypes.byref(dbg.c) if dbg else None))
        return (status, dbg) if debug else status

    @classmethod
    def batch_all_valid(cls, proofs, vk, public_inputs, rnd16=None, want_status=False):
        """OPT-IN aggregate check (include/bn254v.h: bn254v_groth16_batch_all_valid; no counterpart in the reference):
        True iff every proof of the batch is valid, except with probability <= 2^-126, at about half the cost of
        verify_batch; False says nothing about which proof fails.  `rnd16`: (n, 16) uint8 scalars for reproducible tests
        only -- leave None so that the library draws them.  `want_status`: also return the per-record validation statuses
        (OK_TRUE = well-formed and included in the aggregate, not a verdict)."""
        lib = load_library()
        h = _vk("groth16", vk, cls.sign_mode)
        arr, lens = _pack_proofs(proofs)
        n = arr.shape[0]
        inp, k = _pack_inputs(public_inputs, n)
        if rnd16 is not None:
            rnd16 = np.ascontiguousarray(rnd16, dtype=np.uint8)
            assert rnd16.size == 16 * n
        status = np.full(n, STATUS_UNSET, dtype=np.uint8)
        verdict = np.zeros(1, dtype=np.uint8)
        _check(lib.bn254v_groth16_batch_all_valid(h.ptr, _ptr(arr), arr.shape[1], _ptr(lens), _ptr(inp), k, _ptr(rnd16), n,
                                                  _ptr(verdict), _ptr(status)))
        return (bool(verdict[0]), status) if want_status else bool(verdict[0])


class PlonkVerifier:
    """PlonkVerifier (verifier/src/lib.rs:66-74).  verify_plonk never returns Ok(false)
    (verifier/src/plonk/verify.rs:316): the outcome is True or a PlonkError."""

    _ERR = {ERR_BSB22_MISMATCH: "Bsb22CommitmentMismatch", ERR_INVALID_WITNESS: "InvalidWitness",
            ERR_INVERSE_NOT_FOUND: "InverseNotFound", ERR_OPENING_POLY_MISMATCH: "OpeningPolyMismatch",
            ERR_INVALID_NUMBER_OF_DIGESTS: "InvalidNumberOfDigests", ERR_PAIRING_CHECK_FAILED: "PairingCheckFailed"}

    @classmethod
    def verify(cls, proof, vk, public_inputs, rnd=None) -> bool:
        st = cls.verify_batch([proof], vk, [list(public_inputs)], rnd=None if rnd is None else [rnd])[0]
        if st == OK_TRUE:
            return True
        if st in cls._ERR:
            raise PlonkError(cls._ERR[st])
        raise VerifierPanic(status_name(st))

    @classmethod
    def verify_batch(cls, proofs, vk, public_inputs, rnd=None, debug=False, out=None):
        """`rnd`: None (production) -- the library draws the batch-opening scalars from the OS CSPRNG where the reference
        calls Fr::random(OsRng) (verifier/src/plonk/kzg.rs:149-154).  Explicit per-proof scalars are for tests and
        reproducible runs only: a scalar the prover can predict makes the z-shifted opening forgeable (bn254v.h)."""
        lib = load_library()
        h = _vk("plonk", vk)
        arr, lens = _pack_proofs(proofs)
        n = arr.shape[0]
        inp, k = _pack_inputs(public_inputs, n)
        if rnd is None:
            rnd_arr = None
        elif isinstance(rnd, np.ndarray):
            rnd_arr = np.ascontiguousarray(rnd)
        else:
            rnd_arr = fr_to_be([int(v) % (1 << 256) for v in rnd])
        if out is None:
            status = np.full(n, STATUS_UNSET, dtype=np.uint8)
        else:
            assert out.dtype == np.uint8 and out.shape == (n,) and out.flags["C_CONTIGUOUS"]
            status = out
        dbg = DebugOutputs(n, 4, 8) if debug else None
        _check(lib.bn254v_plonk_verify_batch(h.ptr, _ptr(arr), arr.shape[1], _ptr(lens), _ptr(inp), k,
                                             _ptr(rnd_arr), n, _ptr(status), ctypes.byref(dbg.c) if dbg else None))
        return (status, dbg) if debug else status


def pairing_product_batch(g1, g2, k, want_values=False):
    """bn::pairing_batch over n independent k-pair sets: g1 (n, k, 64), g2 (n, k, 128) uint8.
    Returns is_one (n,) and, when want_values, the canonical Fq12 Miller and GT values (n, 384)."""
    lib = load_library()
    g1 = np.ascontiguousarray(g1, dtype=np.uint8)
    g2 = np.ascontiguousarray(g2, dtype=np.uint8)
    n = g1.size // (64 * k)
    assert g2.size == n * 128 * k
    is_one = np.zeros(n, dtype=np.uint8)
    ml = np.zeros((n, 384), dtype=np.uint8) if want_values else None
    gt = np.zeros((n, 384), dtype=np.uint8) if want_values else None
    _check(lib.bn254v_pairing_product_batch(_ptr(g1), _ptr(g2), k, n, _ptr(is_one), _ptr(ml), _ptr(gt)))
    return (is_one, ml, gt) if want_values else is_one


def groth16_synth(seed, n, n_public=2, sign_mode=0, first_index=0):
    """Trapdoor-simulated Groth16 workload (BASELINE.json config 2), generated on the device.
    Returns (vk_bytes, proofs[n,256], inputs[n,n_public,32], expected_status[n])."""
    lib = load_library()
    vk = np.zeros(4096, dtype=np.uint8)
    vk_len = c_size_t(vk.size)
    proofs = np.zeros((n, 256), dtype=np.uint8)
    inputs = np.zeros((n, n_public, 32), dtype=np.uint8)
    expected = np.zeros(n, dtype=np.uint8)
    _check(lib.bn254v_groth16_synth(seed, n_public, sign_mode, first_index, n, _ptr(vk), ctypes.byref(vk_len),
                                    _ptr(proofs), _ptr(inputs), _ptr(expected)))
    return bytes(vk[:vk_len.value]), proofs, inputs, expected


def pairing_synth(seed, n, k=4, first_index=0):
    lib = load_library()
    g1 = np.zeros((n, k, 64), dtype=np.uint8)
    g2 = np.zeros((n, k, 128), dtype=np.uint8)
    expected = np.zeros(n, dtype=np.uint8)
    _check(lib.bn254v_pairing_synth(seed, k, first_index, n, _ptr(g1), _ptr(g2), _ptr(expected)))
    return g1, g2, expected


class _DeviceBatch:
    """A batch staged in HBM (bn254v_*_batch_upload, include/bn254v_bench.h) for kernel-only timing."""
    ptr = None

    def _run(self, fn, args, want_status):
        status = np.full(self.n, STATUS_UNSET, dtype=np.uint8) if want_status else None
        ms = c_float(0)
        _check(fn(*args, _ptr(status), ctypes.byref(ms)))
        return status, ms.value

    def free(self):
        if self.ptr:
            load_library().bn254v_batch_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Groth16DeviceBatch(_DeviceBatch):
    def __init__(self, vk_bytes, proofs, inputs, sign_mode=0):
        lib = load_library()
        self.vk = _vk("groth16", vk_bytes, sign_mode)
        proofs = np.ascontiguousarray(proofs, dtype=np.uint8)
        inputs = np.ascontiguousarray(inputs, dtype=np.uint8)
        self.n = proofs.shape[0]
        out = c_void_p()
        _check(lib.bn254v_groth16_batch_upload(self.vk.ptr, _ptr(proofs), proofs.shape[1], _ptr(inputs),
                                               inputs.shape[1], self.n, ctypes.byref(out)))
        self.ptr = out.value

    def verify(self, want_status=True):
        """Returns (status or None, kernel milliseconds measured with CUDA events on the launching stream)."""
        return self._run(load_library().bn254v_groth16_batch_verify, (self.vk.ptr, self.ptr), want_status)


class PlonkDeviceBatch(_DeviceBatch):
    def __init__(self, vk_bytes, proofs, inputs, rnd):
        lib = load_library()
        self.vk = _vk("plonk", vk_bytes)
        proofs = np.ascontiguousarray(proofs, dtype=np.uint8)
        inputs = np.ascontiguousarray(inputs, dtype=np.uint8)
        rnd = np.ascontiguousarray(rnd, dtype=np.uint8)
        self.n = proofs.shape[0]
        out = c_void_p()
        _check(lib.bn254v_plonk_batch_upload(self.vk.ptr, _ptr(proofs), proofs.shape[1], _ptr(inputs), inputs.shape[1],
                                             _ptr(rnd), self.n, ctypes.byref(out)))
        self.ptr = out.value

    def verify(self, want_status=True):
        return self._run(load_library().bn254v_plonk_batch_verify, (self.vk.ptr, self.ptr), want_status)


class PairingDeviceBatch(_DeviceBatch):
    def __init__(self, g1, g2, k):
        lib = load_library()
        g1 = np.ascontiguousarray(g1, dtype=np.uint8)
        g2 = np.ascontiguousarray(g2, dtype=np.uint8)
        self.n = g1.size // (64 * k)
        out = c_void_p()
        _check(lib.bn254v_pairing_batch_upload(_ptr(g1), _ptr(g2), k, self.n, ctypes.byref(out)))
        self.ptr = out.value

    def verify(self, want_status=True):
        return self._run(load_library().bn254v_pairing_batch_verify, (self.ptr,), want_status)


def last_stage_ms():
    """Device times of the stages of the last *DeviceBatch.verify() (bn254v_last_stage_ms)."""
    buf = (c_float * 8)()
    n = load_library().bn254v_last_stage_ms(buf, 8)
    return [buf[i] for i in range(n)]


def last_kernel_split():
    """(Miller-loop kernel ms, final-exponentiation kernel ms) of the last Groth16DeviceBatch.verify()."""
    a, b = c_float(0), c_float(0)
    _check(load_library().bn254v_last_kernel_split(ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def imad_peak(iters=4096):
    """Measured int32 multiply-add issue rate of device 0 (roofline denominator)."""
    lib = load_library()
    wide, lo, clk = c_double(0), c_double(0), c_float(0)
    _check(lib.bn254v_imad_peak(iters, ctypes.byref(wide), ctypes.byref(lo), ctypes.byref(clk)))
    return {"wide_mac_per_s": wide.value, "lo_mac_per_s": lo.value, "sm_clock_mhz": clk.value}


def verify_many(items, rnd=None, sign_mode=0):
    """Mixed batch over several verifying keys and both proof systems (bn254v_verify_many): `items` is a list of
    (kind, proof_bytes, vk_bytes, public_inputs) with kind in {"groth16", "plonk"}.  The library groups the items by
    (kind, sha256(vk), number of inputs) -- one device batch per group, VK handles cached inside the library -- and
    the status bytes come back in input order.  `rnd`: None (production: drawn inside the library), or per-item scalars
    for the PlonK items (tests only, see bn254v.h)."""
    lib = load_library()
    n = len(items)
    arr = (_Item * max(n, 1))()
    keep = []  # buffers referenced by the item array
    vk_bufs = {}
    for i, (kind, proof, vk, inputs) in enumerate(items):
        if kind not in ("groth16", "plonk"):
            raise ValueError("kind must be 'groth16' or 'plonk'")
        vk = bytes(vk)
        vb = vk_bufs.get(vk)
        if vb is None:
            vb = vk_bufs[vk] = np.frombuffer(vk, dtype=np.uint8)
        pb = np.frombuffer(bytes(proof), dtype=np.uint8) if len(proof) else np.zeros(1, np.uint8)
        ib = fr_to_be(list(inputs)) if len(inputs) else np.zeros((1, 32), np.uint8)
        keep += [pb, ib]
        arr[i] = _Item(KIND_GROTH16 if kind == "groth16" else KIND_PLONK, len(inputs), pb.ctypes.data, len(proof),
                       vb.ctypes.data, len(vk), ib.ctypes.data)
    rnd_arr = None if rnd is None else fr_to_be([int(v) % (1 << 256) for v in rnd])
    out = np.full(n, STATUS_UNSET, dtype=np.uint8)
    _check(lib.bn254v_verify_many(arr, n, sign_mode, _ptr(rnd_arr), _ptr(out)))
    return out


_ITEM_DTYPE = np.dtype([("kind", "<i4"), ("n_inputs", "<i4"), ("proof", "<u8"), ("proof_len", "<u8"), ("vk", "<u8"),
                        ("vk_len", "<u8"), ("inputs", "<u8")])
assert _ITEM_DTYPE.itemsize == ctypes.sizeof(_Item)


class MixedItems:
    """The bn254v_item array of a BASELINE.json configs[4]-shaped batch: n Groth16 records and n PlonK records,
    interleaved one to one (item 2i is Groth16 record i, item 2i + 1 is PlonK record i).  The items point into the
    caller's arrays (no copies; the library gathers each group on the host cores inside the call).  Built once and
    verified any number of times: a C or Rust caller holds such an array already, and filling 2n structured records
    from numpy costs about as much as verifying them."""

    def __init__(self, vk_g, proofs_g, inputs_g, vk_p, proofs_p, inputs_p, rnd_p=None):
        proofs_g, inputs_g = np.ascontiguousarray(proofs_g, np.uint8), np.ascontiguousarray(inputs_g, np.uint8)
        proofs_p, inputs_p = np.ascontiguousarray(proofs_p, np.uint8), np.ascontiguousarray(inputs_p, np.uint8)
        n = proofs_g.shape[0]
        assert proofs_p.shape[0] == n
        vg, vp = np.frombuffer(bytes(vk_g), np.uint8), np.frombuffer(bytes(vk_p), np.uint8)
        items = np.zeros(2 * n, dtype=_ITEM_DTYPE)
        idx = np.arange(n, dtype=np.uint64)
        g, p = items[0::2], items[1::2]
        g["kind"], g["n_inputs"], g["proof_len"] = KIND_GROTH16, inputs_g.shape[1], proofs_g.shape[1]
        g["proof"] = proofs_g.ctypes.data + idx * np.uint64(proofs_g.shape[1])
        g["inputs"] = inputs_g.ctypes.data + idx * np.uint64(32 * inputs_g.shape[1])
        g["vk"], g["vk_len"] = vg.ctypes.data, vg.size
        p["kind"], p["n_inputs"], p["proof_len"] = KIND_PLONK, inputs_p.shape[1], proofs_p.shape[1]
        p["proof"] = proofs_p.ctypes.data + idx * np.uint64(proofs_p.shape[1])
        p["inputs"] = inputs_p.ctypes.data + idx * np.uint64(32 * inputs_p.shape[1])
        p["vk"], p["vk_len"] = vp.ctypes.data, vp.size
        self.rnd = None
        if rnd_p is not None:
            self.rnd = np.zeros((2 * n, 32), np.uint8)
            self.rnd[1::2] = rnd_p
        self.n_items, self.items = 2 * n, items
        self._keep = (proofs_g, inputs_g, proofs_p, inputs_p, vg, vp)  # the items point into these

    def verify(self, sign_mode=0, out=None):
        """ONE bn254v_verify_many call over the items; returns the status bytes (written into `out` when given)."""
        lib = load_library()
        status = np.full(self.n_items, STATUS_UNSET, dtype=np.uint8) if out is None else out
        _check(lib.bn254v_verify_many(ctypes.cast(self.items.ctypes.data, POINTER(_Item)), self.n_items, sign_mode,
                                      _ptr(self.rnd), _ptr(status)))
        return status


def verify_mixed_arrays(vk_g, proofs_g, inputs_g, vk_p, proofs_p, inputs_p, rnd_p=None, sign_mode=0, out=None):
    """MixedItems(...).verify(...) in one step."""
    return MixedItems(vk_g, proofs_g, inputs_g, vk_p, proofs_p, inputs_p, rnd_p).verify(sign_mode, out)
