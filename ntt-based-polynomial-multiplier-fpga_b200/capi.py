"""ctypes binding of ``libnttb200.so`` (include/nttb200.h, include/nttb200_legacy.h).

No compute happens in Python and there is no fallback: if the shared library is missing
or no CUDA device is usable the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")

TRANSFORMS = {
    "ntt_std2rev": 0, "mulntt_std2rev": 1, "intt_rev2std": 2, "inttmul_rev2std": 3,
    "intt_rev2std_scaled": 4, "inttmul_rev2std_scaled": 5, "intt_std2rev": 6, "ntt_rev2std": 7,
}
DATAFLOWS = {"ct_std2rev": 0, "gs_rev2std": 1, "ct_rev2std": 2, "gs_std2rev": 3}
TABLES = {
    "psi_powers": 0, "inv_psi_powers": 1, "scaled_inv_psi_powers": 2, "omega_powers": 3,
    "omega_powers_rev": 4, "inv_omega_powers": 5, "inv_omega_powers_rev": 6, "mixed_powers": 7,
    "mixed_powers_rev": 8, "inv_mixed_powers": 9, "inv_mixed_powers_rev": 10,
    "inv_psi_powers_rev": 11,
}
PLAN_CYCLIC = 1
PLAN_NO_PLANTARD = 2
PLAN_CHECK_RANGE = 4


class NttError(RuntimeError):
    pass


def lib_path() -> str:
    """The in-tree library; NTTB200_LIB points tuning experiments at another build of it."""
    return os.environ.get("NTTB200_LIB") or os.path.join(PKG_DIR, "libnttb200.so")


def build_library(jobs: int = 8) -> str:
    """Compile every CUDA translation unit for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-j", str(jobs), "-C", CSRC], check=True)
    return lib_path()


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise NttError(f"{path} not built: run __graft_entry__.build() (make -C {CSRC}); "
                       "there is no Python/CPU fallback")
    L = C.CDLL(path)
    vp, i32p, u32p, sz = C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t
    L.nttb200_last_error.restype = C.c_char_p
    L.nttb200_plan_create.argtypes = [C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.nttb200_plan_destroy.argtypes = [vp]
    L.nttb200_plan_destroy.restype = None
    for nm in ("nttb200_plan_n", "nttb200_plan_q", "nttb200_plan_psi"):
        getattr(L, nm).argtypes = [vp]
        getattr(L, nm).restype = C.c_uint32
    L.nttb200_plan_device.argtypes = [vp]
    L.nttb200_plan_describe.argtypes = [vp]
    L.nttb200_plan_describe.restype = C.c_char_p
    L.nttb200_polymul_batch.argtypes = [vp, i32p, i32p, i32p, sz]
    L.nttb200_polymul_batch_async.argtypes = [vp, i32p, i32p, i32p, sz, C.POINTER(C.c_ulonglong)]
    L.nttb200_polymul_wait.argtypes = [vp, C.c_ulonglong]
    L.nttb200_polymul_batch_dev.argtypes = [vp, i32p, i32p, i32p, sz, vp]
    L.nttb200_plan_wire_stats.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong),
                                          C.POINTER(C.c_ulonglong), C.POINTER(C.c_int)]
    L.nttb200_ntt_batch.argtypes = [vp, C.c_int, i32p, sz]
    L.nttb200_ntt_batch_dev.argtypes = [vp, C.c_int, i32p, sz, vp]
    L.nttb200_ntt_table_batch.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, u32p, i32p, sz]
    L.nttb200_mul_array_batch.argtypes = [vp, i32p, i32p, i32p, sz]
    L.nttb200_scalar_mul_array_batch.argtypes = [vp, i32p, C.c_int32, sz]
    L.nttb200_make_table.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
    L.nttb200_make_red_table.argtypes = [C.c_int, C.c_uint32, C.c_uint32, i32p]
    L.nttb200_find_psi.argtypes = [C.c_uint32, C.c_uint32]
    L.nttb200_find_psi.restype = C.c_uint32
    L.nttb200_find_omega.argtypes = [C.c_uint32, C.c_uint32]
    L.nttb200_find_omega.restype = C.c_uint32
    L.nttb200_is_prime.argtypes = [C.c_uint32]
    L.nttb200_host_alloc.argtypes = [sz]
    L.nttb200_host_alloc.restype = vp
    L.nttb200_host_free.argtypes = [vp]
    L.nttb200_host_free.restype = None
    L.nttb200_dev_alloc.argtypes = [sz]
    L.nttb200_dev_alloc.restype = vp
    L.nttb200_dev_free.argtypes = [vp]
    L.nttb200_dev_free.restype = None
    L.nttb200_memcpy_h2d.argtypes = [vp, vp, sz, vp]
    L.nttb200_memcpy_d2h.argtypes = [vp, vp, sz, vp]
    L.nttb200_stream_sync.argtypes = [vp]
    L.nttb200_measure_int_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.nttb200_set_device.argtypes = [C.c_int]
    for nm in ("ntt256_product1", "ntt256_product4", "ntt_red256_product1", "ntt_red256_product4"):
        getattr(L, nm).argtypes = [i32p, i32p, i32p]
        getattr(L, nm).restype = None
    for nm in ("ntt_ct_rev2std_v1", "ntt_ct_rev2std", "mulntt_ct_rev2std", "ntt_ct_std2rev",
               "mulntt_ct_std2rev", "ntt_gs_rev2std", "nttmul_gs_rev2std", "ntt_gs_std2rev",
               "nttmul_gs_std2rev", "mul_array16"):
        getattr(L, nm).argtypes = [i32p, C.c_uint32, vp]
        getattr(L, nm).restype = None
    L.mul_array.argtypes = [i32p, C.c_uint32, i32p, i32p]
    L.mul_array.restype = None
    L.scalar_mul_array.argtypes = [i32p, C.c_uint32, C.c_int32]
    L.scalar_mul_array.restype = None
    for nm in ("ntt_red_ct_rev2std", "mulntt_red_ct_rev2std", "ntt_red_ct_std2rev", "mulntt_red_ct_std2rev",
               "ntt_red_gs_rev2std", "nttmul_red_gs_rev2std", "ntt_red_gs_std2rev", "nttmul_red_gs_std2rev",
               "mul_reduce_array16", "shuffle_with_table"):
        getattr(L, nm).argtypes = [i32p, C.c_uint32, vp] if nm != "shuffle_with_table" else [i32p, vp, C.c_uint32]
        getattr(L, nm).restype = None
    for nm in ("normalize", "normalize_inv3", "shift_array", "reduce_array", "reduce_array_twice", "correct",
               "bitrev_shuffle"):
        getattr(L, nm).argtypes = [i32p, C.c_uint32]
        getattr(L, nm).restype = None
    L.mul_reduce_array.argtypes = [i32p, C.c_uint32, i32p, i32p]
    L.mul_reduce_array.restype = None
    L.scalar_mul_reduce_array.argtypes = [i32p, C.c_uint32, C.c_int32]
    L.scalar_mul_reduce_array.restype = None
    L.nttb200_red_ntt_table_batch.argtypes = [C.c_uint32, C.c_int, C.c_int, i32p, i32p, sz]
    L.nttb200_red_elementwise_batch.argtypes = [C.c_int, i32p, i32p, i32p, C.c_int32, sz]
    L.nttb200_bitrev_shuffle_batch.argtypes = [i32p, C.c_uint32, sz]
    L.nttb200_shuffle_with_table.argtypes = [i32p, sz, vp, C.c_uint32]
    L.nttb200_polymul_batch_u16.argtypes = [vp, vp, vp, vp, sz]
    L.nttb200_polymul_batch_u16_dev.argtypes = [vp, vp, vp, vp, sz, vp]
    L.nttb200_multi_create.argtypes = [C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    L.nttb200_multi_destroy.argtypes = [vp]
    L.nttb200_multi_destroy.restype = None
    L.nttb200_multi_gpus.argtypes = [vp]
    L.nttb200_multi_polymul_batch.argtypes = [vp, i32p, i32p, i32p, sz]
    L.nttb200_shard_bounds.argtypes = [sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz)]
    L.nttb200_shard_bounds.restype = None
    L.nttb200_legacy_set_clobber.argtypes = [C.c_int]
    L.nttb200_legacy_set_clobber.restype = None
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise NttError(f"nttb200 error {rc}: {lib().nttb200_last_error().decode()}")


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def device_count() -> int:
    n = lib().nttb200_device_count()
    return max(n, 0)


def set_device(dev: int) -> None:
    _check(lib().nttb200_set_device(dev))


def make_table(kind: str | int, n: int, q: int, psi: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint32)
    k = TABLES[kind] if isinstance(kind, str) else kind
    _check(lib().nttb200_make_table(k, n, q, psi, _ptr(out)))
    return out


def find_psi(n: int, q: int) -> int:
    return int(lib().nttb200_find_psi(n, q))


def find_omega(n: int, q: int) -> int:
    return int(lib().nttb200_find_omega(n, q))


def is_prime(q: int) -> bool:
    return bool(lib().nttb200_is_prime(q))


def last_launch_count() -> int:
    return int(lib().nttb200_last_launch_count())


def measure_int_peak(which: int) -> float:
    v = C.c_double(0)
    _check(lib().nttb200_measure_int_peak(which, C.byref(v)))
    return v.value


class _Pinned:
    """A pinned host buffer exposed as a numpy int32 array."""

    def __init__(self, shape) -> None:
        self.nbytes = int(np.prod(shape)) * 4
        self.ptr = lib().nttb200_host_alloc(self.nbytes)
        if not self.ptr:
            raise NttError(lib().nttb200_last_error().decode())
        buf = (C.c_int32 * (self.nbytes // 4)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.int32).reshape(shape)

    def free(self) -> None:
        if self.ptr:
            self.array = None
            lib().nttb200_host_free(self.ptr)
            self.ptr = None


def host_alloc(shape) -> _Pinned:
    return _Pinned(shape)


def shard_bounds(batch: int, world: int, rank: int) -> tuple[int, int]:
    lo, hi = C.c_size_t(0), C.c_size_t(0)
    lib().nttb200_shard_bounds(batch, world, rank, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


class MultiPlan:
    """One plan per GPU of the box; host batches are split into contiguous slices (no collective)."""

    def __init__(self, n: int, q: int, psi: int = 0, ngpus: int = 0) -> None:
        h = C.c_void_p()
        _check(lib().nttb200_multi_create(C.byref(h), n, q, psi, 0, ngpus))
        self._h, self.n, self.q = h, n, q
        self.gpus = int(lib().nttb200_multi_gpus(h))

    def polymul_host_ptr(self, c_ptr: int, a_ptr: int, b_ptr: int, batch: int) -> None:
        _check(lib().nttb200_multi_polymul_batch(self._h, c_ptr, a_ptr, b_ptr, batch))

    def polymul(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, self.n)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, self.n)
        c = np.empty_like(a)
        self.polymul_host_ptr(_ptr(c), _ptr(a), _ptr(b), a.shape[0])
        return c

    def close(self) -> None:
        if self._h:
            lib().nttb200_multi_destroy(self._h)
            self._h = None


class Plan:
    """One (n, q, psi) parameter set on the current CUDA device."""

    def __init__(self, n: int, q: int, psi: int = 0, cyclic: bool = False, no_plantard: bool = False,
                 check_range: bool = False) -> None:
        h = C.c_void_p()
        flags = ((PLAN_CYCLIC if cyclic else 0) | (PLAN_NO_PLANTARD if no_plantard else 0) |
                 (PLAN_CHECK_RANGE if check_range else 0))
        _check(lib().nttb200_plan_create(C.byref(h), n, q, psi, flags))
        self._h = h
        self.n, self.q = n, q
        self.psi = int(lib().nttb200_plan_psi(h))

    def close(self) -> None:
        if self._h:
            lib().nttb200_plan_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # best effort
        try:
            self.close()
        except Exception:
            pass

    def describe(self) -> str:
        return lib().nttb200_plan_describe(self._h).decode()

    # ---- host buffers (numpy) -------------------------------------------------
    def polymul(self, a: np.ndarray, b: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, self.n)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, self.n)
        if a.shape != b.shape:
            raise ValueError("a and b must have the same shape")
        c = np.empty_like(a) if out is None else out
        _check(lib().nttb200_polymul_batch(self._h, _ptr(c), _ptr(a), _ptr(b), a.shape[0]))
        return c

    def wire_stats(self) -> dict:
        """Polynomials of the last host-buffer call that crossed PCIe as 16-bit / 32-bit words."""
        c16, c32, r16, th = C.c_ulonglong(0), C.c_ulonglong(0), C.c_ulonglong(0), C.c_int(0)
        _check(lib().nttb200_plan_wire_stats(self._h, C.byref(c16), C.byref(c32), C.byref(r16), C.byref(th)))
        return {"rows16": c16.value, "rows32": c32.value, "result_rows16": r16.value, "host_threads": th.value}

    def polymul_u16(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """Packed 16-bit extension (q <= 12385): uint16 in, uint16 out."""
        a = np.ascontiguousarray(a, dtype=np.uint16).reshape(-1, self.n)
        b = np.ascontiguousarray(b, dtype=np.uint16).reshape(-1, self.n)
        c = np.empty_like(a)
        _check(lib().nttb200_polymul_batch_u16(self._h, _ptr(c), _ptr(a), _ptr(b), a.shape[0]))
        return c

    def polymul_u16_dev(self, c_ptr: int, a_ptr: int, b_ptr: int, batch: int, stream: int = 0) -> None:
        _check(lib().nttb200_polymul_batch_u16_dev(self._h, c_ptr, a_ptr, b_ptr, batch, stream))

    def transform(self, kind: str, a: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, self.n).copy()
        _check(lib().nttb200_ntt_batch(self._h, TRANSFORMS[kind], _ptr(x), x.shape[0]))
        return x

    def mul_array(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, self.n)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, self.n)
        c = np.empty_like(a)
        _check(lib().nttb200_mul_array_batch(self._h, _ptr(c), _ptr(a), _ptr(b), a.shape[0]))
        return c

    def scalar_mul_array(self, a: np.ndarray, s: int) -> np.ndarray:
        x = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, self.n).copy()
        _check(lib().nttb200_scalar_mul_array_batch(self._h, _ptr(x), s, x.shape[0]))
        return x

    # ---- raw pointers (device-resident or pinned) -------------------------------
    def polymul_dev(self, c_ptr: int, a_ptr: int, b_ptr: int, batch: int, stream: int = 0) -> None:
        _check(lib().nttb200_polymul_batch_dev(self._h, c_ptr, a_ptr, b_ptr, batch, stream))

    def transform_dev(self, kind: str, a_ptr: int, batch: int, stream: int = 0) -> None:
        _check(lib().nttb200_ntt_batch_dev(self._h, TRANSFORMS[kind], a_ptr, batch, stream))

    def polymul_host_ptr(self, c_ptr: int, a_ptr: int, b_ptr: int, batch: int) -> None:
        _check(lib().nttb200_polymul_batch(self._h, c_ptr, a_ptr, b_ptr, batch))

    def polymul_async_ptr(self, c_ptr: int, a_ptr: int, b_ptr: int, batch: int) -> int:
        """Queue a host-buffer product (nttb200_polymul_batch_async); returns its ticket.  The buffers
        must stay alive and untouched until wait(ticket) has returned."""
        t = C.c_ulonglong(0)
        _check(lib().nttb200_polymul_batch_async(self._h, c_ptr, a_ptr, b_ptr, batch, C.byref(t)))
        return int(t.value)

    def wait(self, ticket: int = 0) -> None:
        """Block until the product with this ticket is complete (0: every product queued so far)."""
        _check(lib().nttb200_polymul_wait(self._h, ticket))


def ntt_table_batch(n: int, q: int, dataflow: str, table: np.ndarray, a: np.ndarray,
                    skip_j0: bool = False) -> np.ndarray:
    """skip_j0: the un-merged reference entry points (ntt_ct_*, ntt_gs_*), which do the j = 0
    butterflies without a multiplication and never read p[t] (R/NTT/ntt.C:313-317)."""
    x = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, n).copy()
    p = np.ascontiguousarray(table, dtype=np.uint32)
    _check(lib().nttb200_ntt_table_batch(n, q, DATAFLOWS[dataflow], int(skip_j0), _ptr(p), _ptr(x), x.shape[0]))
    return x


class legacy:
    """The reference's own call surface (include/nttb200_legacy.h), one polynomial per call."""

    @staticmethod
    def product(name: str, a: np.ndarray, b: np.ndarray, clobber: bool = False):
        a = np.ascontiguousarray(a, dtype=np.int32).copy()
        b = np.ascontiguousarray(b, dtype=np.int32).copy()
        c = np.zeros(256, dtype=np.int32)
        lib().nttb200_legacy_set_clobber(int(clobber))
        getattr(lib(), name)(_ptr(c), _ptr(a), _ptr(b))
        lib().nttb200_legacy_set_clobber(0)
        return c, a, b

    @staticmethod
    def transform(name: str, a: np.ndarray, table16: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(a, dtype=np.int32).copy()
        p = np.ascontiguousarray(table16, dtype=np.uint16)
        getattr(lib(), name)(_ptr(x), x.shape[0], _ptr(p))
        return x

    @staticmethod
    def red_transform(name: str, a: np.ndarray, table16: np.ndarray) -> np.ndarray:
        """ntt_red_* / mulntt_red_* / nttmul_red_* with the caller's int16 table; unreduced output."""
        x = np.ascontiguousarray(a, dtype=np.int32).copy()
        p = np.ascontiguousarray(table16, dtype=np.int16)
        getattr(lib(), name)(_ptr(x), x.shape[0], _ptr(p))
        return x

    @staticmethod
    def red_helper(name: str, a: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(a, dtype=np.int32).copy()
        getattr(lib(), name)(_ptr(x), x.shape[0])
        return x

    @staticmethod
    def mul_array16(a, table16):
        x = np.ascontiguousarray(a, dtype=np.int32).copy()
        p = np.ascontiguousarray(table16, dtype=np.uint16)
        lib().mul_array16(_ptr(x), x.shape[0], _ptr(p))
        return x

    @staticmethod
    def mul_array(a, b):
        a = np.ascontiguousarray(a, dtype=np.int32)
        b = np.ascontiguousarray(b, dtype=np.int32)
        c = np.zeros_like(a)
        lib().mul_array(_ptr(c), a.shape[0], _ptr(a), _ptr(b))
        return c

    @staticmethod
    def scalar_mul_array(a, s: int):
        x = np.ascontiguousarray(a, dtype=np.int32).copy()
        lib().scalar_mul_array(_ptr(x), x.shape[0], s)
        return x
