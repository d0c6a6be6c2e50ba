"""Batch engine: a thin, typed wrapper over the C ABI (include/vrfs_b200.h).  Array conventions are the
ABI's: scalars (n,32) uint8 little-endian, points (n,64) uint8 affine x||y little-endian."""
import ctypes as C

import numpy as np

from . import _lib

BANDERSNATCH, ED25519, P256 = 0, 1, 2
SUITE_NAMES = {0: "Bandersnatch_SHA-512_ELL2", 1: "Ed25519_SHA-512_TAI", 2: "secp256r1 (RFC 9381 suite 0x01)"}


def _u8(a, shape):
    a = np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray, memoryview)) else np.asarray(a, dtype=np.uint8)
    return np.ascontiguousarray(a).reshape(shape)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_var(items):
    """list of bytes -> (uint8 data, uint64 offsets[n+1]); `items` may already be such a pair, or None."""
    if items is None:
        return None, None
    if isinstance(items, tuple):
        return items
    off = np.zeros(len(items) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in items], dtype=np.uint64)
    data = np.frombuffer(b"".join(items), dtype=np.uint8).copy() if off[-1] else np.zeros(16, dtype=np.uint8)
    return data, off


class Engine:
    """One context on one GPU (`device` = CUDA ordinal).  Not thread-safe; one per process per GPU."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        st = self._lib.vrfs_ctx_create(int(device), C.byref(self._ctx))
        if st != _lib.OK:
            msg = self._lib.vrfs_last_error(self._ctx).decode() if self._ctx else "context allocation failed"
            if self._ctx:
                self._lib.vrfs_ctx_destroy(self._ctx)
                self._ctx = None
            raise _lib.VrfsError(st, msg)
        self.device = device

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.vrfs_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st):
        if st != _lib.OK:
            raise _lib.VrfsError(st, self._lib.vrfs_last_error(self._ctx).decode())

    def sync(self):
        self._check(self._lib.vrfs_ctx_sync(self._ctx))

    @property
    def launch_count(self):
        return int(self._lib.vrfs_ctx_launch_count(self._ctx))

    @property
    def stream(self):
        return self._lib.vrfs_ctx_stream(self._ctx)

    # ---- ietf
    def ietf_verify(self, suite, pk, inp, outp, c, s, ad=None):
        pk = _u8(pk, (-1, 64)); n = len(pk)
        inp = _u8(inp, (n, 64)); outp = _u8(outp, (n, 64)); c = _u8(c, (n, 32)); s = _u8(s, (n, 32))
        adb, off = pack_var(ad)
        ok = np.zeros(n, np.uint8)
        self._check(self._lib.vrfs_ietf_verify_batch(self._ctx, suite, C.c_size_t(n), _p(pk), _p(inp), _p(outp), _p(c), _p(s), _p(adb), _p(off), _p(ok)))
        return ok

    def ietf_verify_dev(self, suite, n, d_pk, d_inp, d_outp, d_c, d_s, d_ok, d_ad=None, d_off=None):
        """device pointers (ints); enqueues on the context stream, returns immediately"""
        v = lambda x: C.c_void_p(x) if x else None
        self._check(self._lib.vrfs_ietf_verify_batch_dev(self._ctx, suite, C.c_size_t(n), v(d_pk), v(d_inp), v(d_outp), v(d_c), v(d_s), v(d_ad), v(d_off), v(d_ok)))

    def ietf_verify_host_ptrs(self, suite, n, pk, inp, outp, c, s, ok, ad=None, off=None):
        """raw host pointers (ints), e.g. of pinned torch tensors; synchronous"""
        v = lambda x: C.c_void_p(x) if x else None
        self._check(self._lib.vrfs_ietf_verify_batch(self._ctx, suite, C.c_size_t(n), v(pk), v(inp), v(outp), v(c), v(s), v(ad), v(off), v(ok)))

    # ---- measurement
    def enable_kernel_timing(self, on=True):
        self._check(self._lib.vrfs_ctx_enable_kernel_timing(self._ctx, int(bool(on))))

    def kernel_timings(self):
        """[(kernel name, device ms)] of the most recent batch call (syncs the stream)"""
        names = (C.c_char_p * 64)(); ms = (C.c_float * 64)()
        n = self._lib.vrfs_ctx_kernel_timings(self._ctx, names, ms, 64)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def measure_mac32_peak(self, variant=0):
        macs = C.c_double(); mhz = C.c_double()
        self._check(self._lib.vrfs_measure_mac32_peak(self._ctx, int(variant), C.byref(macs), C.byref(mhz)))
        return macs.value, mhz.value
