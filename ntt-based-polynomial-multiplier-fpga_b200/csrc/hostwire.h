/*
 * hostwire.h -- internal: 16-bit wire format of the host-buffer product path (hostwire.c).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* single-threaded kernels: narrow returns the OR of all source words */
uint32_t nttb200_wire_narrow(uint16_t *dst, const int32_t *src, size_t words);
void nttb200_wire_widen(int32_t *dst, const uint16_t *src, size_t words);
void nttb200_wire_copy(int32_t *dst, const int32_t *src, size_t words, int streaming);

/* worker pool.  begin/end bracket a batch call (workers spin in between, sleep otherwise);
 * post_* queue one array and return a job id; wait() helps until the job is complete.
 * *mask_out (may be NULL) receives, by atomic OR, every narrowed word that has bits above 15. */
int nttb200_wire_threads(void);
void nttb200_wire_begin(void);
void nttb200_wire_end(void);
uint64_t nttb200_wire_post_narrow(uint16_t *dst, const int32_t *src, size_t words, uint32_t *mask_out);
uint64_t nttb200_wire_post_widen(int32_t *dst, const uint16_t *src, size_t words);
uint64_t nttb200_wire_post_copy(int32_t *dst, const int32_t *src, size_t words, int streaming);
int nttb200_wire_done(uint64_t id);
void nttb200_wire_wait(uint64_t id);
void nttb200_wire_help(void);          /* do one block of any queued job, or pause */

#ifdef __cplusplus
}
#endif
