"""ctypes front-end to the CPU checkers (TEST INFRASTRUCTURE ONLY).

Two shared objects, both built by ``oracle/Makefile``:

* ``_build/libntt_oracle*.so`` -- the parametrised restatement (``ntt_oracle.c``);
* ``_ref/libntt_ref*.so``      -- the UNMODIFIED reference C compiled from
  ``/root/reference`` (only buildable where that tree exists; the prebuilt object
  travels to the GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product (``libnttb200.so``)
never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# enum orc_table
PSI_POWERS, INV_PSI_POWERS, SCALED_INV_PSI_POWERS, OMEGA_POWERS, OMEGA_POWERS_REV, \
    INV_OMEGA_POWERS, INV_OMEGA_POWERS_REV, MIXED_POWERS, MIXED_POWERS_REV, \
    INV_MIXED_POWERS, INV_MIXED_POWERS_REV, INV_PSI_POWERS_REV = range(12)
RED_SCALED_INV_PSI_POWERS_VAR = 100
# enum orc_variant
PRODUCT_CT, PRODUCT_GS, PRODUCT_MERGED, PRODUCT_SCHOOLBOOK, PRODUCT_CYCLIC = 1, 4, 10, 20, 30
# ref_product variants
REF_CT, REF_GS, REF_RED_CT, REF_RED_GS, REF_MERGED, REF_RED_MERGED = 1, 4, 101, 104, 10, 110

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def _has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = line.split()
                    return all(x in fl for x in ("avx2", "bmi2", "fma"))
    except OSError:
        pass
    return False


def build(ref: bool = True) -> None:
    """(Re)build the checkers; the reference object only where /root/reference exists."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)


def _pick(dirname: str, stem: str) -> str | None:
    names = [f"{stem}_v3.so", f"{stem}.so"] if _has_avx2() else [f"{stem}.so"]
    for nm in names:
        p = os.path.join(HERE, dirname, nm)
        if os.path.exists(p):
            return p
    return None


class Oracle:
    """The parametrised restatement (``ntt_oracle.c``)."""

    def __init__(self) -> None:
        path = _pick("_build", "libntt_oracle")
        if path is None:
            build(ref=False)
            path = _pick("_build", "libntt_oracle")
        self.path = path
        L = self.lib = C.CDLL(path)
        L.orc_powmod.restype = C.c_uint32
        L.orc_powmod.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32]
        L.orc_invmod.restype = C.c_uint32
        L.orc_invmod.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_smallest_psi.restype = C.c_uint32
        L.orc_smallest_psi.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_smallest_omega.restype = C.c_uint32
        L.orc_smallest_omega.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_make_table.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, _u32p]
        L.orc_make_omega_table.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, _u32p]
        L.orc_red_make_table.argtypes = [C.c_int, C.c_uint32, C.c_uint32, _i32p]
        for nm in ("orc_ntt_ct_rev2std_v1", "orc_ntt_ct_rev2std", "orc_mulntt_ct_rev2std",
                   "orc_ntt_ct_std2rev", "orc_mulntt_ct_std2rev", "orc_ntt_gs_rev2std",
                   "orc_nttmul_gs_rev2std", "orc_ntt_gs_std2rev", "orc_nttmul_gs_std2rev",
                   "orc_mul_array_tab"):
            getattr(L, nm).argtypes = [_i32p, C.c_uint32, _u32p, C.c_uint32]
            getattr(L, nm).restype = None
        for nm in ("orc_red_ct_rev2std", "orc_red_mulntt_ct_rev2std", "orc_red_ct_std2rev",
                   "orc_red_mulntt_ct_std2rev", "orc_red_gs_rev2std", "orc_red_nttmul_gs_rev2std",
                   "orc_red_gs_std2rev", "orc_red_nttmul_gs_std2rev", "orc_red_mul_reduce_array_tab"):
            getattr(L, nm).argtypes = [_i32p, C.c_uint32, _i32p]
            getattr(L, nm).restype = None
        for nm in ("orc_red_shift_array", "orc_red_reduce_array", "orc_red_reduce_array_twice",
                   "orc_red_correct", "orc_red_normalize", "orc_red_normalize_inv3",
                   "orc_bitrev_shuffle"):
            getattr(L, nm).argtypes = [_i32p, C.c_uint32]
            getattr(L, nm).restype = None
        L.orc_mul_array.argtypes = [_i32p, C.c_uint32, _i32p, _i32p, C.c_uint32]
        L.orc_mul_array.restype = None
        L.orc_scalar_mul_array.argtypes = [_i32p, C.c_uint32, C.c_int32, C.c_uint32]
        L.orc_scalar_mul_array.restype = None
        L.orc_red_mul_reduce_array.argtypes = [_i32p, C.c_uint32, _i32p, _i32p]
        L.orc_red_scalar_mul_reduce_array.argtypes = [_i32p, C.c_uint32, C.c_int32]
        L.orc_red_product.argtypes = [C.c_uint32, C.c_uint32, C.c_int, _i32p, _i32p, _i32p]
        L.orc_plan_create.restype = C.c_void_p
        L.orc_plan_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_plan_destroy.argtypes = [C.c_void_p]
        L.orc_plan_psi.restype = C.c_uint32
        L.orc_plan_psi.argtypes = [C.c_void_p]
        L.orc_product.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, _i32p]
        L.orc_product_batch.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, _i32p, C.c_size_t]
        L.orc_bench_loop.restype = C.c_double
        L.orc_bench_loop.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p, C.c_size_t, C.c_double,
                                     C.POINTER(C.c_uint64)]
        L.orc_fill_random.argtypes = [_i32p, C.c_size_t, C.c_uint32, C.c_uint64]
        L.orc_fill_random.restype = None
        self._plans: dict[tuple[int, int, int], int] = {}

    # -- helpers -----------------------------------------------------------
    def plan(self, n: int, q: int, psi: int = 0) -> int:
        key = (n, q, psi)
        if key not in self._plans:
            h = self.lib.orc_plan_create(n, q, psi)
            if not h:
                raise ValueError(f"no NTT plan for n={n} q={q} psi={psi}")
            self._plans[key] = h
        return self._plans[key]

    def psi(self, n: int, q: int, psi: int = 0) -> int:
        return int(self.lib.orc_plan_psi(self.plan(n, q, psi)))

    def table(self, kind: int, n: int, q: int, psi: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint32)
        if self.lib.orc_make_table(kind, n, q, psi, out) != 0:
            raise ValueError("bad table kind")
        return out

    def omega_table(self, kind: int, n: int, q: int, omega: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint32)
        if self.lib.orc_make_omega_table(kind, n, q, omega, out) != 0:
            raise ValueError("bad table kind")
        return out

    def red_table(self, kind: int, n: int, psi: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.int32)
        if self.lib.orc_red_make_table(kind, n, psi, out) != 0:
            raise ValueError("bad table kind")
        return out

    def random(self, shape, q: int, seed: int) -> np.ndarray:
        out = np.empty(shape, dtype=np.int32)
        self.lib.orc_fill_random(out.reshape(-1), out.size, q, seed)
        return out

    def transform(self, name: str, a: np.ndarray, p: np.ndarray, q: int) -> np.ndarray:
        """Apply orc_<name> row by row; returns a new array."""
        a = np.ascontiguousarray(a, dtype=np.int32).copy()
        n = a.shape[-1]
        fn = getattr(self.lib, "orc_" + name)
        p = np.ascontiguousarray(p, dtype=np.uint32)
        for row in a.reshape(-1, n):
            fn(row, n, p, q)
        return a

    def red_transform(self, name: str, a: np.ndarray, p: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32).copy()
        n = a.shape[-1]
        fn = getattr(self.lib, "orc_red_" + name)
        p = np.ascontiguousarray(p, dtype=np.int32)
        for row in a.reshape(-1, n):
            fn(row, n, p)
        return a

    def product(self, n: int, q: int, a: np.ndarray, b: np.ndarray, variant: int = PRODUCT_MERGED,
                psi: int = 0) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, n)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, n)
        c = np.empty_like(a)
        rc = self.lib.orc_product_batch(self.plan(n, q, psi), variant, c, a, b, a.shape[0])
        if rc != 0:
            raise ValueError(f"orc_product_batch rc={rc}")
        return c

    def red_product(self, n: int, psi: int, a: np.ndarray, b: np.ndarray, variant: int) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, n).copy()
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, n).copy()
        c = np.empty_like(a)
        for i in range(a.shape[0]):
            if self.lib.orc_red_product(n, psi, variant, c[i], a[i], b[i]) != 0:
                raise ValueError("orc_red_product failed")
        return c

    def bench_loop(self, n: int, q: int, a: np.ndarray, b: np.ndarray, variant: int,
                   seconds: float, psi: int = 0) -> tuple[float, int]:
        calls = C.c_uint64(0)
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, n)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, n)
        rate = self.lib.orc_bench_loop(self.plan(n, q, psi), variant, a, b, a.shape[0], seconds,
                                       C.byref(calls))
        return float(rate), int(calls.value)


class Reference:
    """The unmodified reference C, compiled (n=256, q=12289 only)."""

    N, Q, PSI = 256, 12289, 1002

    def __init__(self, flavour: str | None = None) -> None:
        if flavour == "O0":
            path = os.path.join(HERE, "_ref", "libntt_ref_O0.so")
            path = path if os.path.exists(path) else None
        else:
            path = _pick("_ref", "libntt_ref")
        if path is None and os.path.isdir("/root/reference"):
            build(ref=True)
            path = _pick("_ref", "libntt_ref")
        if path is None:
            raise FileNotFoundError("oracle/_ref/libntt_ref*.so not built (needs /root/reference)")
        self.path = path
        L = self.lib = C.CDLL(path)
        L.ref_product.argtypes = [C.c_int, _i32p, _i32p, _i32p]
        L.ref_product_batch.argtypes = [C.c_int, _i32p, _i32p, _i32p, C.c_size_t]
        L.ref_transform.argtypes = [C.c_int, _i32p]
        L.ref_table.restype = C.c_void_p
        L.ref_table.argtypes = [C.c_int, C.c_int]
        L.ref_param.argtypes = [C.c_int]
        L.ref_red_helper.argtypes = [C.c_int, _i32p, C.c_uint32]
        L.ref_red_helper.restype = None
        L.ref_bitrev_shuffle.argtypes = [_i32p, C.c_uint32]
        L.ref_bitrev_shuffle.restype = None
        L.ref_bench_loop.restype = C.c_double
        L.ref_bench_loop.argtypes = [C.c_int, _i32p, _i32p, C.c_size_t, C.c_double,
                                     C.POINTER(C.c_uint64)]
        if hasattr(L, "ref_transform_tab"):
            L.ref_transform_tab.argtypes = [C.c_int, _i32p, C.c_uint32,
                                            np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")]
            L.ref_red_transform_tab.argtypes = [C.c_int, _i32p, C.c_uint32,
                                                np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")]

    # ids of ref_transform_tab / ref_red_transform_tab -> reference function name
    TAB_IDS = {0: "ntt_ct_rev2std_v1", 1: "ntt_ct_rev2std", 2: "mulntt_ct_rev2std", 3: "ntt_ct_std2rev",
               4: "mulntt_ct_std2rev", 5: "ntt_gs_rev2std", 6: "nttmul_gs_rev2std", 7: "ntt_gs_std2rev",
               8: "nttmul_gs_std2rev"}

    def transform_tab(self, tid: int, a: np.ndarray, p: np.ndarray, red: bool = False) -> np.ndarray:
        """The generic reference entry point `tid` (TAB_IDS) run with the caller's table."""
        a = np.ascontiguousarray(a, dtype=np.int32).copy()
        n = a.shape[-1]
        p = np.ascontiguousarray(p, dtype=np.int16 if red else np.uint16)
        fn = self.lib.ref_red_transform_tab if red else self.lib.ref_transform_tab
        for row in a.reshape(-1, n):
            if fn(tid, row, n, p) != 0:
                raise ValueError("bad transform id")
        return a

    def product_post_state(self, a: np.ndarray, b: np.ndarray, variant: int):
        """(c, a_after, b_after) of ONE reference product call (the operands are modified)."""
        a = np.ascontiguousarray(a, dtype=np.int32).copy()
        b = np.ascontiguousarray(b, dtype=np.int32).copy()
        c = np.zeros(256, dtype=np.int32)
        if self.lib.ref_product(variant, c, a, b) != 0:
            raise ValueError("bad variant")
        return c, a, b

    def product(self, a: np.ndarray, b: np.ndarray, variant: int) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, 256)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, 256)
        c = np.empty_like(a)
        if self.lib.ref_product_batch(variant, c, a, b, a.shape[0]) != 0:
            raise ValueError("bad variant")
        return c

    def transform(self, tid: int, a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.int32).copy()
        for row in a.reshape(-1, 256):
            if self.lib.ref_transform(tid, row) != 0:
                raise ValueError("bad transform id")
        return a

    def table(self, kind: int, red: bool = False) -> np.ndarray:
        ptr = self.lib.ref_table(kind, int(red))
        if not ptr:
            raise ValueError("bad table kind")
        ct = C.c_int16 if red else C.c_uint16
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(256,))
        return arr.astype(np.int64)

    def param(self, which: int) -> int:
        return int(self.lib.ref_param(which))

    def bench_loop(self, a: np.ndarray, b: np.ndarray, variant: int, seconds: float) -> tuple[float, int]:
        calls = C.c_uint64(0)
        a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1, 256)
        b = np.ascontiguousarray(b, dtype=np.int32).reshape(-1, 256)
        rate = self.lib.ref_bench_loop(variant, a, b, a.shape[0], seconds, C.byref(calls))
        return float(rate), int(calls.value)


def reference_available() -> bool:
    return _pick("_ref", "libntt_ref") is not None or os.path.isdir("/root/reference")
