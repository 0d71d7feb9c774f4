/*
 * hostwire.c -- host side of the 16-bit WIRE FORMAT of nttb200_polymul_batch().
 *
 * The reference's call surface carries coefficients as int32_t (R/NTT/ntt256.h:76-86) and
 * so does nttb200_polymul_batch().  For half-word moduli (q <= 12385: 12289, 7681, 3329) the
 * upper half of every word is zero, and the host-buffer path is bound by the PCIe link
 * (2 n words in, n words out per product), not by the kernel.  So the
 * words are narrowed to uint16_t on the host, by a small pool of worker threads, into pinned
 * staging buffers; the link carries half the bytes; the kernel reads and writes uint16_t rows
 * (polymul_plant_kernel<..., uint16_t>); and the result rows are widened back into the
 * caller's int32_t buffer.  This replaces what the reference's transfer program does with its
 * 32-bit FIFO words (COMM/linux_app/NTT_PCIECommunicationv2.c:166-224).
 *
 * Callers that hold their polynomials in pageable (malloc) memory -- the reference's own
 * convention -- and cannot use the 16-bit wire (moduli above 12385) get the same pool as a
 * parallel stager: 32-bit words are copied into / out of the pinned staging by the workers,
 * 5x faster than the driver's pageable cudaMemcpyAsync path on the GPU box.
 *
 * This file holds only byte shuffling: narrowing, widening, copying and the thread pool that
 * runs them.  No modular arithmetic happens on the CPU.  Narrowing also ORs all the input words
 * together: if any word does not fit 16 bits the caller falls back to the 32-bit wire for that
 * part of the batch, so results never depend on the wire format.
 *
 * Pool: process-wide, created on first use, NTTB200_HOST_THREADS workers (default: the CPUs
 * this process may run on, at most 32, minus the calling thread).  Jobs (one array to narrow
 * or widen) go into a small ring; workers and the waiting caller grab fixed-size blocks of the
 * oldest unfinished job with a compare-and-swap on (job id, block index).  Workers spin while
 * a batch call is in flight (nttb200_wire_begin/end) and sleep on a condition variable
 * otherwise.
 */
#define _GNU_SOURCE
#include "hostwire.h"

#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- narrowing / widening kernels -------------------------------------------------- */

static uint32_t narrow_scalar(uint16_t *dst, const int32_t *src, size_t n) {
  uint32_t acc = 0;
  for (size_t i = 0; i < n; i++) {
    acc |= (uint32_t)src[i];
    dst[i] = (uint16_t)src[i];
  }
  return acc;
}
static void widen_scalar(int32_t *dst, const uint16_t *src, size_t n) {
  for (size_t i = 0; i < n; i++) dst[i] = src[i];
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static uint32_t narrow_avx2(uint16_t *dst, const int32_t *src, size_t n) {
  __m256i acc = _mm256_setzero_si256();
  size_t i = 0;
  /* (prefetchnta on the source, to keep it out of the way of the staging buffers in the caches,
   * measured 22 % slower end to end on the GPU box: the hardware prefetcher does better) */
  for (; i + 32 <= n; i += 32) {
    const __m256i a0 = _mm256_loadu_si256((const __m256i *)(src + i));
    const __m256i a1 = _mm256_loadu_si256((const __m256i *)(src + i + 8));
    const __m256i a2 = _mm256_loadu_si256((const __m256i *)(src + i + 16));
    const __m256i a3 = _mm256_loadu_si256((const __m256i *)(src + i + 24));
    acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(a0, a1), _mm256_or_si256(a2, a3)));
    /* packus works per 128-bit lane: restore element order with a 64-bit permute */
    const __m256i p0 = _mm256_permute4x64_epi64(_mm256_packus_epi32(a0, a1), 0xD8);
    const __m256i p1 = _mm256_permute4x64_epi64(_mm256_packus_epi32(a2, a3), 0xD8);
    _mm256_storeu_si256((__m256i *)(dst + i), p0);
    _mm256_storeu_si256((__m256i *)(dst + i + 16), p1);
  }
  uint32_t lanes[8];
  _mm256_storeu_si256((__m256i *)lanes, acc);
  uint32_t r = 0;
  for (int k = 0; k < 8; k++) r |= lanes[k];
  return r | narrow_scalar(dst + i, src + i, n - i);
}
/* the result buffer is written once and not read again by this library: non-temporal stores
 * (no read-for-ownership of the caller's lines) when the destination is 32-byte aligned */
__attribute__((target("avx2"))) static void widen_avx2(int32_t *dst, const uint16_t *src, size_t n) {
  size_t i = 0;
  if (((uintptr_t)dst & 31u) == 0) {
    for (; i + 16 <= n; i += 16) {
      const __m256i lo = _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i *)(src + i)));
      const __m256i hi = _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i *)(src + i + 8)));
      _mm256_stream_si256((__m256i *)(dst + i), lo);
      _mm256_stream_si256((__m256i *)(dst + i + 8), hi);
    }
    _mm_sfence();
  } else {
    for (; i + 16 <= n; i += 16) {
      const __m256i lo = _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i *)(src + i)));
      const __m256i hi = _mm256_cvtepu16_epi32(_mm_loadu_si128((const __m128i *)(src + i + 8)));
      _mm256_storeu_si256((__m256i *)(dst + i), lo);
      _mm256_storeu_si256((__m256i *)(dst + i + 8), hi);
    }
  }
  widen_scalar(dst + i, src + i, n - i);
}
#endif

#if defined(__x86_64__)
/* 32-bit words into a buffer that this library does not read again (the caller's c): non-temporal */
__attribute__((target("avx2"))) static void copy_stream_avx2(int32_t *dst, const int32_t *src, size_t n) {
  size_t i = 0;
  if (((uintptr_t)dst & 31u) == 0) {
    for (; i + 32 <= n; i += 32) {
      const __m256i a0 = _mm256_loadu_si256((const __m256i *)(src + i));
      const __m256i a1 = _mm256_loadu_si256((const __m256i *)(src + i + 8));
      const __m256i a2 = _mm256_loadu_si256((const __m256i *)(src + i + 16));
      const __m256i a3 = _mm256_loadu_si256((const __m256i *)(src + i + 24));
      _mm256_stream_si256((__m256i *)(dst + i), a0);
      _mm256_stream_si256((__m256i *)(dst + i + 8), a1);
      _mm256_stream_si256((__m256i *)(dst + i + 16), a2);
      _mm256_stream_si256((__m256i *)(dst + i + 24), a3);
    }
    _mm_sfence();
  }
  memcpy(dst + i, src + i, (n - i) * sizeof(int32_t));
}
#endif

#if defined(__x86_64__)
/* Variants of the two kernels above, picked by NTTB200_WIRE_SIMD = avx2 | avx2nt | avx512 | avx512nt
 * (default avx2; "nt" = non-temporal stores into the pinned staging too, sparing the read-for-ownership
 * of lines that only the DMA engine reads; avx512 = vpmovdw / vpmovzxwd on 512-bit registers where
 * the host has AVX-512BW).  Measured on the GPU box's host: DESIGN.md section 4. */
__attribute__((target("avx2"))) static uint32_t narrow_avx2_nt(uint16_t *dst, const int32_t *src, size_t n) {
  __m256i acc = _mm256_setzero_si256();
  size_t i = 0;
  if (((uintptr_t)dst & 31u) == 0) {
    for (; i + 32 <= n; i += 32) {
      const __m256i a0 = _mm256_loadu_si256((const __m256i *)(src + i));
      const __m256i a1 = _mm256_loadu_si256((const __m256i *)(src + i + 8));
      const __m256i a2 = _mm256_loadu_si256((const __m256i *)(src + i + 16));
      const __m256i a3 = _mm256_loadu_si256((const __m256i *)(src + i + 24));
      acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(a0, a1), _mm256_or_si256(a2, a3)));
      _mm256_stream_si256((__m256i *)(dst + i), _mm256_permute4x64_epi64(_mm256_packus_epi32(a0, a1), 0xD8));
      _mm256_stream_si256((__m256i *)(dst + i + 16), _mm256_permute4x64_epi64(_mm256_packus_epi32(a2, a3), 0xD8));
    }
    _mm_sfence();
  }
  uint32_t lanes[8];
  _mm256_storeu_si256((__m256i *)lanes, acc);
  uint32_t r = 0;
  for (int k = 0; k < 8; k++) r |= lanes[k];
  return r | narrow_avx2(dst + i, src + i, n - i);
}
__attribute__((target("avx512f,avx512bw"))) static uint32_t narrow_avx512(uint16_t *dst, const int32_t *src, size_t n, int nt) {
  __m512i acc = _mm512_setzero_si512();
  size_t i = 0;
  nt = nt && (((uintptr_t)dst & 63u) == 0);
  for (; i + 32 <= n; i += 32) {
    const __m512i a0 = _mm512_loadu_si512((const void *)(src + i));
    const __m512i a1 = _mm512_loadu_si512((const void *)(src + i + 16));
    acc = _mm512_or_si512(acc, _mm512_or_si512(a0, a1));
    const __m512i p = _mm512_inserti64x4(_mm512_castsi256_si512(_mm512_cvtepi32_epi16(a0)), _mm512_cvtepi32_epi16(a1), 1);
    if (nt) _mm512_stream_si512((__m512i *)(dst + i), p);
    else _mm512_storeu_si512((void *)(dst + i), p);
  }
  if (nt) _mm_sfence();
  const uint32_t r = (uint32_t)_mm512_reduce_or_epi32(acc);
  return r | narrow_scalar(dst + i, src + i, n - i);
}
__attribute__((target("avx512f,avx512bw"))) static void widen_avx512(int32_t *dst, const uint16_t *src, size_t n) {
  size_t i = 0;
  if (((uintptr_t)dst & 63u) == 0) {
    for (; i + 32 <= n; i += 32) {
      const __m512i v = _mm512_loadu_si512((const void *)(src + i));
      _mm512_stream_si512((__m512i *)(dst + i), _mm512_cvtepu16_epi32(_mm512_castsi512_si256(v)));
      _mm512_stream_si512((__m512i *)(dst + i + 16), _mm512_cvtepu16_epi32(_mm512_extracti64x4_epi64(v, 1)));
    }
    _mm_sfence();
  }
  widen_scalar(dst + i, src + i, n - i);
}
/* 0 avx2, 1 avx2 + nt, 2 avx512, 3 avx512 + nt */
static int g_simd = -1;
static int wire_simd(void) {
  if (g_simd < 0) {
    const char *e = getenv("NTTB200_WIRE_SIMD");
    int v = 0;
    if (e && !strcmp(e, "avx2nt")) v = 1;
    else if (e && !strcmp(e, "avx512")) v = 2;
    else if (e && !strcmp(e, "avx512nt")) v = 3;
    __builtin_cpu_init();
    if (v >= 2 && !(__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw"))) v -= 2;
    g_simd = v;
  }
  return g_simd;
}
#endif

static int g_have_avx2 = -1;
static int have_avx2(void) {
  if (g_have_avx2 < 0) {
#if defined(__x86_64__)
    __builtin_cpu_init();
    g_have_avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
#else
    g_have_avx2 = 0;
#endif
  }
  return g_have_avx2;
}

uint32_t nttb200_wire_narrow(uint16_t *dst, const int32_t *src, size_t n) {
#if defined(__x86_64__)
  if (have_avx2()) {
    const int v = wire_simd();
    if (v >= 2) return narrow_avx512(dst, src, n, v == 3);
    return v == 1 ? narrow_avx2_nt(dst, src, n) : narrow_avx2(dst, src, n);
  }
#endif
  return narrow_scalar(dst, src, n);
}
void nttb200_wire_widen(int32_t *dst, const uint16_t *src, size_t n) {
#if defined(__x86_64__)
  if (have_avx2()) {
    if (wire_simd() >= 2) widen_avx512(dst, src, n);
    else widen_avx2(dst, src, n);
    return;
  }
#endif
  widen_scalar(dst, src, n);
}

void nttb200_wire_copy(int32_t *dst, const int32_t *src, size_t n, int streaming) {
#if defined(__x86_64__)
  if (streaming && have_avx2()) { copy_stream_avx2(dst, src, n); return; }
#endif
  (void)streaming;
  memcpy(dst, src, n * sizeof(int32_t));
}

/* ---- the pool ---------------------------------------------------------------------- */

enum { JOB_NARROW = 1, JOB_WIDEN = 2, JOB_COPY = 3, JOB_COPY_STREAM = 4 };
#define RING 256                      /* jobs in flight (a batch call keeps < 4 per slot)      */
#define BLOCK_WORDS ((size_t)16384)   /* words per grabbed block: 64 KiB of int32              */

typedef struct {
  int kind;
  void *dst;
  const void *src;
  size_t words;
  uint32_t nblocks;
  _Atomic uint64_t cursor;            /* (job id << 32) | next block: a stale worker's CAS fails */
  _Atomic uint32_t finished;          /* blocks completed                                        */
  _Atomic uint32_t *mask_out;         /* narrowing: the caller's OR of every word that has high bits */
} job_t;

static struct {
  job_t jobs[RING];
  _Atomic uint64_t posted;            /* jobs 0 .. posted-1 exist                                */
  _Atomic uint64_t retired;           /* jobs 0 .. retired-1 are complete and their slot is free */
  _Atomic int active;                 /* batch calls in flight: workers spin while > 0           */
  _Atomic int stop;
  pthread_mutex_t mu;                 /* posting, retiring, sleeping                             */
  pthread_cond_t cv;
  int nworkers;
  int started;
  pthread_t th[64];
} G = {.mu = PTHREAD_MUTEX_INITIALIZER, .cv = PTHREAD_COND_INITIALIZER};

static void run_block(job_t *J, int kind, void *dst, const void *src, size_t words, uint32_t blk,
                      _Atomic uint32_t *mask_out) {
  const size_t lo = (size_t)blk * BLOCK_WORDS;
  const size_t cnt = words - lo < BLOCK_WORDS ? words - lo : BLOCK_WORDS;
  if (kind == JOB_NARROW) {
    const uint32_t m = nttb200_wire_narrow((uint16_t *)dst + lo, (const int32_t *)src + lo, cnt);
    if ((m & 0xffff0000u) && mask_out) atomic_fetch_or_explicit(mask_out, m, memory_order_relaxed);
  } else if (kind == JOB_WIDEN) {
    nttb200_wire_widen((int32_t *)dst + lo, (const uint16_t *)src + lo, cnt);
  } else {
    nttb200_wire_copy((int32_t *)dst + lo, (const int32_t *)src + lo, cnt, kind == JOB_COPY_STREAM);
  }
  atomic_fetch_add_explicit(&J->finished, 1, memory_order_release);
}

/* take one block of job `id` if it has any left; 1 = did a block */
static int help_job(uint64_t id) {
  job_t *J = &G.jobs[id % RING];
  uint64_t cur = atomic_load_explicit(&J->cursor, memory_order_acquire);
  for (;;) {
    if ((cur >> 32) != (id & 0xffffffffu)) return 0;          /* slot belongs to another job   */
    /* fields are stable while the cursor carries this id (written before the cursor) */
    const uint32_t blk = (uint32_t)cur, nblocks = J->nblocks;
    if (blk >= nblocks) return 0;
    const int kind = J->kind;
    void *dst = J->dst;
    const void *src = J->src;
    const size_t words = J->words;
    _Atomic uint32_t *mask_out = J->mask_out;
    if (atomic_compare_exchange_weak_explicit(&J->cursor, &cur, cur + 1, memory_order_acq_rel,
                                              memory_order_acquire)) {
      run_block(J, kind, dst, src, words, blk, mask_out);
      return 1;
    }
  }
}
static int help_any(void) {
  const uint64_t lo = atomic_load_explicit(&G.retired, memory_order_acquire);
  const uint64_t hi = atomic_load_explicit(&G.posted, memory_order_acquire);
  for (uint64_t id = lo; id < hi; id++)
    if (help_job(id)) return 1;
  return 0;
}

static void *worker(void *arg) {
  (void)arg;
  unsigned idle = 0;
  for (;;) {
    if (atomic_load_explicit(&G.active, memory_order_acquire) == 0) {
      /* calls tend to come back to back: stay hot for a moment (~0.2 ms) before going to sleep,
       * so that the next call does not pay for waking the pool up */
      int again = 0;
#if defined(__x86_64__)
      const unsigned long long t0 = __builtin_ia32_rdtsc();
      while (!again && __builtin_ia32_rdtsc() - t0 < 600000ull) {           /* ~0.2-0.3 ms of TSC */
        _mm_pause();
        again = atomic_load_explicit(&G.active, memory_order_acquire) != 0 ||
                atomic_load_explicit(&G.stop, memory_order_relaxed);
      }
#endif
      if (again) continue;
      pthread_mutex_lock(&G.mu);
      while (atomic_load(&G.active) == 0 && !atomic_load(&G.stop)) pthread_cond_wait(&G.cv, &G.mu);
      pthread_mutex_unlock(&G.mu);
      idle = 0;
    }
    if (atomic_load_explicit(&G.stop, memory_order_relaxed)) return NULL;
    if (help_any()) { idle = 0; continue; }
#if defined(__x86_64__)
    _mm_pause();
#endif
    if (++idle > 20000) { sched_yield(); idle = 0; }          /* oversubscribed hosts          */
  }
}

/* a forked child has none of the workers: start over with an empty pool of its own */
static void pool_after_fork_in_child(void) {
  pthread_mutex_init(&G.mu, NULL);
  pthread_cond_init(&G.cv, NULL);
  G.started = 0;
  G.nworkers = 0;
  atomic_store(&G.active, 0);
  atomic_store(&G.posted, 0);
  atomic_store(&G.retired, 0);
}

static int pool_start(void) {
  static int atfork_set = 0;
  pthread_mutex_lock(&G.mu);
  if (!atfork_set) {
    pthread_atfork(NULL, NULL, pool_after_fork_in_child);
    atfork_set = 1;
  }
  if (!G.started) {
    int want = -1;
    const char *e = getenv("NTTB200_HOST_THREADS");
    if (e) want = atoi(e) - 1;                                  /* the caller is one of them    */
    if (want < 0) {
      cpu_set_t set;
      int cpus = (sched_getaffinity(0, sizeof set, &set) == 0) ? CPU_COUNT(&set) : (int)sysconf(_SC_NPROCESSORS_ONLN);
      if (cpus > 32) cpus = 32;
      want = cpus - 1;
    }
    if (want > 63) want = 63;
    if (want < 0) want = 0;
    for (int i = 0; i < RING; i++) atomic_store(&G.jobs[i].cursor, ~0ull);
    int made = 0;
    for (int i = 0; i < want; i++)
      if (pthread_create(&G.th[made], NULL, worker, NULL) == 0) made++;
    G.nworkers = made;
    G.started = 1;
  }
  pthread_mutex_unlock(&G.mu);
  return G.nworkers;
}

int nttb200_wire_threads(void) { return pool_start() + 1; }

void nttb200_wire_begin(void) {
  pool_start();
  pthread_mutex_lock(&G.mu);
  atomic_fetch_add(&G.active, 1);
  pthread_cond_broadcast(&G.cv);
  pthread_mutex_unlock(&G.mu);
}
void nttb200_wire_end(void) { atomic_fetch_sub(&G.active, 1); }

static uint64_t post(int kind, void *dst, const void *src, size_t words, _Atomic uint32_t *mask_out) {
  for (;;) {
    pthread_mutex_lock(&G.mu);
    uint64_t r = atomic_load(&G.retired);
    const uint64_t p = atomic_load(&G.posted);
    while (r < p) {                                            /* retire finished jobs in order */
      job_t *O = &G.jobs[r % RING];
      if (atomic_load_explicit(&O->finished, memory_order_acquire) < O->nblocks) break;
      r++;
    }
    atomic_store_explicit(&G.retired, r, memory_order_release);
    if (p - r < RING) {
      job_t *J = &G.jobs[p % RING];
      /* Invalidate the cursor BEFORE the fields change: a worker that still holds the previous
       * job's cursor value (same slot, RING jobs ago) may read the new fields, but its
       * compare-and-swap then meets a cursor that is neither its stale value nor (yet) the new
       * job's, fails, and re-reads.  Without this it could win the CAS on the stale value while the
       * fields were half rewritten. */
      atomic_store_explicit(&J->cursor, ~0ull, memory_order_release);
      J->kind = kind;
      J->dst = dst;
      J->src = src;
      J->words = words;
      J->nblocks = (uint32_t)((words + BLOCK_WORDS - 1) / BLOCK_WORDS);
      J->mask_out = mask_out;
      atomic_store(&J->finished, 0);
      atomic_store_explicit(&J->cursor, (p & 0xffffffffu) << 32, memory_order_release);
      atomic_store_explicit(&G.posted, p + 1, memory_order_release);
      pthread_mutex_unlock(&G.mu);
      return p;
    }
    pthread_mutex_unlock(&G.mu);
    help_any();                                                /* ring full: work it off        */
  }
}

uint64_t nttb200_wire_post_narrow(uint16_t *dst, const int32_t *src, size_t words, uint32_t *mask_out) {
  return post(JOB_NARROW, dst, src, words, (_Atomic uint32_t *)mask_out);
}
uint64_t nttb200_wire_post_copy(int32_t *dst, const int32_t *src, size_t words, int streaming) {
  return post(streaming ? JOB_COPY_STREAM : JOB_COPY, dst, src, words, NULL);
}
uint64_t nttb200_wire_post_widen(int32_t *dst, const uint16_t *src, size_t words) {
  return post(JOB_WIDEN, dst, src, words, NULL);
}

/* A job's ring slot is recycled only after the job has been retired, and `retired` is
 * re-read on every turn, so a slot that changed hands is noticed one turn later at worst. */
int nttb200_wire_done(uint64_t id) {
  if (id < atomic_load_explicit(&G.retired, memory_order_acquire)) return 1;
  job_t *J = &G.jobs[id % RING];
  const uint64_t cur = atomic_load_explicit(&J->cursor, memory_order_acquire);
  if ((cur >> 32) != (id & 0xffffffffu)) return id < atomic_load_explicit(&G.retired, memory_order_acquire);
  return atomic_load_explicit(&J->finished, memory_order_acquire) >= J->nblocks;
}

void nttb200_wire_wait(uint64_t id) {
  while (!nttb200_wire_done(id)) {
    if (!help_job(id) && !help_any()) {
#if defined(__x86_64__)
      _mm_pause();
#endif
    }
  }
}

void nttb200_wire_help(void) {
  if (!help_any()) {
#if defined(__x86_64__)
    _mm_pause();
#endif
  }
}
