"""Host side of the 16-bit wire format (csrc/hostwire.c): narrowing, widening and the worker
pool, exercised on the CPU through the symbols libnttb200.so exports.  Pure byte shuffling --
no modular arithmetic -- so numpy's astype is the whole checker."""
import ctypes as C
import threading

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L(nttb200):
    lib = nttb200.lib()
    lib.nttb200_wire_narrow.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.nttb200_wire_narrow.restype = C.c_uint32
    lib.nttb200_wire_widen.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.nttb200_wire_widen.restype = None
    lib.nttb200_wire_post_narrow.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.nttb200_wire_post_narrow.restype = C.c_uint64
    lib.nttb200_wire_post_widen.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.nttb200_wire_post_widen.restype = C.c_uint64
    lib.nttb200_wire_post_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    lib.nttb200_wire_post_copy.restype = C.c_uint64
    lib.nttb200_wire_wait.argtypes = [C.c_uint64]
    lib.nttb200_wire_wait.restype = None
    lib.nttb200_wire_done.argtypes = [C.c_uint64]
    lib.nttb200_wire_begin.restype = None
    lib.nttb200_wire_end.restype = None
    return lib


@pytest.mark.parametrize("words", [0, 1, 15, 16, 31, 32, 33, 255, 4096, 16384, 16385, 100003])
@pytest.mark.parametrize("offset", [0, 1, 3])
def test_narrow_and_widen_every_length_and_alignment(L, words, offset):
    rng = np.random.default_rng(words * 7 + offset)
    src = rng.integers(0, 65536, words + offset, dtype=np.int32)[offset:]
    dst = np.full(words + 8, 0xABCD, np.uint16)
    mask = L.nttb200_wire_narrow(dst.ctypes.data, src.ctypes.data, words)
    assert (dst[:words] == src.astype(np.uint16)).all() and (dst[words:] == 0xABCD).all()
    assert mask == (int(np.bitwise_or.reduce(src.view(np.uint32))) if words else 0)
    back = np.full(words + offset + 8, -7, np.int32)
    L.nttb200_wire_widen(back[offset:].ctypes.data, dst.ctypes.data, words)
    assert (back[offset:offset + words] == src).all()
    assert (back[:offset] == -7).all() and (back[offset + words:] == -7).all()


@pytest.mark.parametrize("bad", [65536, 1 << 20, -1, -(1 << 31)])
def test_narrow_reports_words_that_do_not_fit(L, bad):
    src = np.arange(5000, dtype=np.int32)
    src[4321] = bad
    dst = np.zeros(5000, np.uint16)
    assert L.nttb200_wire_narrow(dst.ctypes.data, src.ctypes.data, 5000) & 0xFFFF0000


def test_pool_many_jobs_in_flight(L):
    rng = np.random.default_rng(5)
    L.nttb200_wire_begin()
    try:
        jobs = []
        for k in range(40):
            words = int(rng.integers(1, 200000))
            src = rng.integers(0, 12289, words, dtype=np.int32)
            if k % 7 == 3:
                src[words // 2] = 70000
            dst = np.zeros(words, np.uint16)
            mask = np.zeros(1, np.uint32)
            jid = L.nttb200_wire_post_narrow(dst.ctypes.data, src.ctypes.data, words, mask.ctypes.data)
            jobs.append((jid, src, dst, mask, k % 7 == 3))
        for jid, src, dst, mask, bad in jobs:
            L.nttb200_wire_wait(jid)
            assert L.nttb200_wire_done(jid)
            ok = src < 65536                    # what a word that does not fit becomes is unspecified
            assert (dst[ok] == src[ok].astype(np.uint16)).all()
            assert bool(mask[0] & 0xFFFF0000) == bad
        wide = []
        for jid, src, dst, mask, bad in jobs:
            out = np.full(src.size, -1, np.int32)
            wide.append((L.nttb200_wire_post_widen(out.ctypes.data, dst.ctypes.data, src.size), out, dst))
        for jid, out, dst in wide:
            L.nttb200_wire_wait(jid)
            assert (out == dst.astype(np.int32)).all()
    finally:
        L.nttb200_wire_end()


def test_pool_serves_several_posting_threads_and_recycles_its_ring(L):
    """More jobs than ring slots (256), posted from 4 threads at once (one per GPU in
    nttb200_multi_polymul_batch)."""
    errors = []

    def run(seed):
        rng = np.random.default_rng(seed)
        L.nttb200_wire_begin()
        try:
            for _ in range(150):
                words = int(rng.integers(1, 40000))
                src = rng.integers(0, 65536, words, dtype=np.int32)
                dst = np.zeros(words, np.uint16)
                jid = L.nttb200_wire_post_narrow(dst.ctypes.data, src.ctypes.data, words, None)
                L.nttb200_wire_wait(jid)
                if not (dst == src.astype(np.uint16)).all():
                    errors.append(seed)
        finally:
            L.nttb200_wire_end()

    ts = [threading.Thread(target=run, args=(s,)) for s in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors


@pytest.mark.parametrize("streaming", [0, 1])
@pytest.mark.parametrize("words,offset", [(1, 0), (31, 1), (32, 0), (40000, 0), (100003, 3)])
def test_pool_copies_32_bit_words(L, words, offset, streaming):
    rng = np.random.default_rng(words)
    src = rng.integers(-2**31, 2**31 - 1, words + offset, dtype=np.int64).astype(np.int32)[offset:]
    dst = np.full(words + offset + 8, 7, np.int32)
    L.nttb200_wire_begin()
    try:
        L.nttb200_wire_wait(L.nttb200_wire_post_copy(dst[offset:].ctypes.data, src.ctypes.data, words, streaming))
    finally:
        L.nttb200_wire_end()
    assert (dst[offset:offset + words] == src).all()
    assert (dst[:offset] == 7).all() and (dst[offset + words:] == 7).all()


def test_pool_survives_fork(nttb200):
    """A forked child (multiprocessing, the bench's CPU legs) starts with an empty pool of its own.
    Run in a fresh interpreter: pytest's own threads and hooks stay out of the fork."""
    import subprocess
    import sys
    code = f"""
import ctypes as C, numpy as np, os, warnings
warnings.simplefilter("ignore")
L = C.CDLL({nttb200.lib_path()!r})
L.nttb200_wire_post_narrow.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
L.nttb200_wire_post_narrow.restype = C.c_uint64
L.nttb200_wire_wait.argtypes = [C.c_uint64]
src = np.arange(50000, dtype=np.int32)
def narrow():
    dst = np.zeros(50000, np.uint16)
    L.nttb200_wire_begin()
    L.nttb200_wire_wait(L.nttb200_wire_post_narrow(dst.ctypes.data, src.ctypes.data, 50000, None))
    L.nttb200_wire_end()
    return bool((dst == src.astype(np.uint16)).all())
assert narrow()
pid = os.fork()
if pid == 0:
    os._exit(0 if narrow() and narrow() else 1)
_, status = os.waitpid(pid, 0)
assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0, status
assert narrow()
print("ok")
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr
