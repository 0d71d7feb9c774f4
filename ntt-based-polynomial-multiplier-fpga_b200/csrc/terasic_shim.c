/*
 * terasic_shim.c -- builds `terasic_pcie_qsys.so`: the 12-symbol dlopen ABI of the reference's
 * FPGA transfer path, served by the GPU library instead of the DE2i-150 board.
 *
 * Replaces (C/ = Multiplier_NTT_Based/Software_Hardware_Comunnicator/linux_app/):
 *   the vendor blob C/terasic_pcie_qsys.so, loaded by PCIE_Load(): dlopen("./terasic_pcie_qsys.so")
 *   + dlsym of PCIE_Open/Close/Read32/Write32/Read16/Write16/Read8/Write8/DmaWrite/DmaRead/
 *   DmaFifoWrite/DmaFifoRead (C/PCIE.c:59-103, typedefs C/TERASIC_PCIE.h:166-178), so that the
 *   reference host program C/NTT_PCIECommunicationv2.c runs unchanged: copy this library next
 *   to it under that name.
 *
 * Device model = the protocol NTT_HARDWARE_EXE speaks (C/NTT_PCIECommunicationv2.c:13-25, 43-51,
 * 166-224), which mirrors the Verilog testbench (Hardware_Multiplier/NTT_PolyMul_test.v:82-160):
 *   BAR0 + 0x00 CONTROL  bit0 = start (rising edge latches the command), bits 3:1 = mode
 *   BAR0 + 0x20 STATUS   bit0 = busy, bit1 = done_all
 *   FIFO 0x40 (in)       mode 0: W || W_INV || q || n_inv   (q is taken from the stream; the
 *                                 FPGA-order twiddles are not needed by the GPU kernels)
 *                        mode 1: polynomial A, mode 2: polynomial B   (n = bytes / 4 words)
 *   mode 3 = GO          H2D(A,B) -> fused NTT product kernel -> D2H(C), asynchronously on a CUDA
 *                        stream from pinned buffers; STATUS is answered from the stream state
 *   FIFO 0x80 (out)      polynomial C, natural order, canonical [0, q)
 * The ring is Z_q[x]/(x^n + 1) (what NTT_Software computes and the reference README states);
 * NTTB200_SHIM_CYCLIC=1 selects x^n - 1, the product the committed PolyMult.v datapath forms
 * (Hardware_Multiplier/PolyMult.v:279-283).  Plain reads/writes elsewhere hit a small register
 * file / 1 MiB local memory so that the board bring-up loopbacks (C/app.c:95-132) also pass.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nttb200.h"

typedef int BOOL;
typedef unsigned int DWORD;
typedef unsigned short WORD;
typedef unsigned char BYTE;
typedef int PCIE_HANDLE;
typedef int PCIE_BAR;

#define ADDR_CONTROL 0x00u
#define ADDR_STATUS 0x20u
#define FIFO_IN_ID 0x40u
#define FIFO_OUT_ID 0x80u
#define STATUS_BUSY 1u
#define STATUS_DONE 2u
#define LOCAL_MEM_BYTES (1u << 20)
#define REGFILE_BYTES 4096u
#define MAX_N (1u << 17)

static struct shim_state {
  int open;
  DWORD control, mode;
  int busy, done, running;
  uint32_t q, n_inv, n;
  nttb200_plan *plan;
  uint32_t plan_n, plan_q;
  int32_t *h_a, *h_b, *h_c;            /* pinned, MAX_N words each */
  int32_t *d_a, *d_b, *d_c;
  void *stream;
  size_t have_a, have_b;
  BYTE *local_mem;
  BYTE regs[REGFILE_BYTES];
  size_t out_pos;
} S;

static int shim_debug(void) {
  static int v = -1;
  if (v < 0) v = getenv("NTTB200_SHIM_DEBUG") != NULL;
  return v;
}
#define DBG(...) do { if (shim_debug()) fprintf(stderr, "[terasic-shim] " __VA_ARGS__); } while (0)

static void poll_stream(void) {
  if (S.running && nttb200_stream_query(S.stream) != 0) {   /* finished (or failed) */
    S.running = 0;
    S.busy = 0;
    S.done = 1;
  }
}

static BOOL launch_product(void) {
  if (!S.q || !S.n || S.have_a != S.n || S.have_b != S.n) {
    DBG("GO without a complete parameter/A/B load (q=%u n=%u a=%zu b=%zu)\n", S.q, S.n, S.have_a, S.have_b);
    return 0;
  }
  if (!S.plan || S.plan_n != S.n || S.plan_q != S.q) {
    if (S.plan) nttb200_plan_destroy(S.plan);
    S.plan = NULL;
    const char *cyc = getenv("NTTB200_SHIM_CYCLIC");
    if (nttb200_plan_create(&S.plan, S.n, S.q, 0, (cyc && *cyc == '1') ? NTTB200_PLAN_CYCLIC : 0) != 0) {
      fprintf(stderr, "terasic shim: %s\n", nttb200_last_error());
      return 0;
    }
    S.plan_n = S.n;
    S.plan_q = S.q;
  }
  const size_t bytes = (size_t)S.n * sizeof(int32_t);
  if (nttb200_memcpy_h2d(S.d_a, S.h_a, bytes, S.stream) || nttb200_memcpy_h2d(S.d_b, S.h_b, bytes, S.stream) ||
      nttb200_polymul_batch_dev(S.plan, S.d_c, S.d_a, S.d_b, 1, S.stream) ||
      nttb200_memcpy_d2h(S.h_c, S.d_c, bytes, S.stream)) {
    fprintf(stderr, "terasic shim: %s\n", nttb200_last_error());
    return 0;
  }
  S.running = 1;
  S.busy = 1;
  S.done = 0;
  S.out_pos = 0;
  return 1;
}

PCIE_HANDLE PCIE_Open(WORD vid, WORD did, WORD card) {
  (void)vid; (void)did; (void)card;
  if (S.open) return 1;
  if (nttb200_device_count() < 1) {
    fprintf(stderr, "terasic shim: no CUDA device (%s)\n", nttb200_last_error());
    return 0;                                            /* 0 = failure, as the vendor library */
  }
  memset(&S, 0, sizeof S);
  const size_t bytes = (size_t)MAX_N * sizeof(int32_t);
  S.h_a = (int32_t *)nttb200_host_alloc(bytes);
  S.h_b = (int32_t *)nttb200_host_alloc(bytes);
  S.h_c = (int32_t *)nttb200_host_alloc(bytes);
  S.d_a = (int32_t *)nttb200_dev_alloc(bytes);
  S.d_b = (int32_t *)nttb200_dev_alloc(bytes);
  S.d_c = (int32_t *)nttb200_dev_alloc(bytes);
  S.stream = nttb200_stream_create();
  S.local_mem = (BYTE *)calloc(LOCAL_MEM_BYTES, 1);
  if (!S.h_a || !S.h_b || !S.h_c || !S.d_a || !S.d_b || !S.d_c || !S.stream || !S.local_mem) {
    fprintf(stderr, "terasic shim: allocation failed (%s)\n", nttb200_last_error());
    return 0;
  }
  S.open = 1;
  return 1;
}

void PCIE_Close(PCIE_HANDLE h) {
  (void)h;
  if (!S.open) return;
  if (S.stream) nttb200_stream_sync(S.stream);
  if (S.plan) nttb200_plan_destroy(S.plan);
  nttb200_host_free(S.h_a); nttb200_host_free(S.h_b); nttb200_host_free(S.h_c);
  nttb200_dev_free(S.d_a); nttb200_dev_free(S.d_b); nttb200_dev_free(S.d_c);
  nttb200_stream_destroy(S.stream);
  free(S.local_mem);
  memset(&S, 0, sizeof S);
}

static void control_write(DWORD v) {
  const int rising = (v & 1u) && !(S.control & 1u);
  S.control = v;
  if (!rising) return;
  S.mode = (v >> 1) & 7u;
  DBG("command: mode %u\n", S.mode);
  switch (S.mode) {
    case 0: S.busy = 1; S.done = 0; break;               /* expects the parameter stream */
    case 1: S.busy = 1; S.done = 0; S.have_a = 0; break;
    case 2: S.busy = 1; S.done = 0; S.have_b = 0; break;
    case 3: if (!launch_product()) { S.busy = 0; S.done = 0; } break;
    default: break;
  }
}

BOOL PCIE_Write32(PCIE_HANDLE h, PCIE_BAR bar, DWORD addr, DWORD data) {
  (void)h;
  if (!S.open) return 0;
  if (bar == 0 && addr == ADDR_CONTROL) { control_write(data); return 1; }
  if (addr > REGFILE_BYTES - 4) return 0;
  memcpy(S.regs + addr, &data, 4);
  return 1;
}
BOOL PCIE_Read32(PCIE_HANDLE h, PCIE_BAR bar, DWORD addr, DWORD *data) {
  (void)h;
  if (!S.open || !data) return 0;
  if (bar == 0 && addr == ADDR_STATUS) {
    poll_stream();
    *data = (S.busy ? STATUS_BUSY : 0u) | (S.done ? STATUS_DONE : 0u);
    return 1;
  }
  if (bar == 0 && addr == ADDR_CONTROL) { *data = S.control; return 1; }
  if (addr > REGFILE_BYTES - 4) return 0;
  memcpy(data, S.regs + addr, 4);
  return 1;
}
BOOL PCIE_Write16(PCIE_HANDLE h, PCIE_BAR bar, DWORD addr, WORD data) {
  (void)h; (void)bar;
  if (!S.open || addr > REGFILE_BYTES - 2) return 0;
  memcpy(S.regs + addr, &data, 2);
  return 1;
}
BOOL PCIE_Read16(PCIE_HANDLE h, PCIE_BAR bar, DWORD addr, WORD *data) {
  (void)h; (void)bar;
  if (!S.open || !data || addr > REGFILE_BYTES - 2) return 0;
  memcpy(data, S.regs + addr, 2);
  return 1;
}
BOOL PCIE_Write8(PCIE_HANDLE h, PCIE_BAR bar, DWORD addr, BYTE data) {
  (void)h; (void)bar;
  if (!S.open || addr > REGFILE_BYTES - 1) return 0;
  S.regs[addr] = data;
  return 1;
}
BOOL PCIE_Read8(PCIE_HANDLE h, PCIE_BAR bar, DWORD addr, BYTE *data) {
  (void)h; (void)bar;
  if (!S.open || !data || addr > REGFILE_BYTES - 1) return 0;
  *data = S.regs[addr];
  return 1;
}

BOOL PCIE_DmaWrite(PCIE_HANDLE h, DWORD local, void *buf, DWORD bytes) {
  (void)h;
  if (!S.open || !buf || (uint64_t)local + bytes > LOCAL_MEM_BYTES) return 0;
  memcpy(S.local_mem + local, buf, bytes);
  return 1;
}
BOOL PCIE_DmaRead(PCIE_HANDLE h, DWORD local, void *buf, DWORD bytes) {
  (void)h;
  if (!S.open || !buf || (uint64_t)local + bytes > LOCAL_MEM_BYTES) return 0;
  memcpy(buf, S.local_mem + local, bytes);
  return 1;
}

BOOL PCIE_DmaFifoWrite(PCIE_HANDLE h, DWORD fifo, void *buf, DWORD bytes) {
  (void)h;
  if (!S.open || !buf || fifo != FIFO_IN_ID || (bytes & 3u)) return 0;
  const uint32_t *w = (const uint32_t *)buf;
  const size_t words = bytes / 4;
  switch (S.mode) {
    case 0:                                              /* W || W_INV || q || n_inv */
      if (words < 2) return 0;
      S.q = w[words - 2];
      S.n_inv = w[words - 1];
      S.busy = 0;
      DBG("parameters: %zu words, q=%u n_inv=%u\n", words, S.q, S.n_inv);
      return 1;
    case 1:
    case 2: {
      size_t *have = S.mode == 1 ? &S.have_a : &S.have_b;
      int32_t *dst = S.mode == 1 ? S.h_a : S.h_b;
      if (*have + words > MAX_N) return 0;
      memcpy(dst + *have, w, bytes);
      *have += words;
      /* the ring size is what the host streams: a power of two completes the load */
      if (*have >= 8 && (*have & (*have - 1)) == 0) {
        S.n = (uint32_t)*have;
        S.busy = 0;
      }
      DBG("poly %c: %zu words\n", S.mode == 1 ? 'A' : 'B', *have);
      return 1;
    }
    default:
      return 0;
  }
}

BOOL PCIE_DmaFifoRead(PCIE_HANDLE h, DWORD fifo, void *buf, DWORD bytes) {
  (void)h;
  if (!S.open || !buf || fifo != FIFO_OUT_ID || (bytes & 3u)) return 0;
  if (S.running) {                                       /* the DMA would block until data is there */
    nttb200_stream_sync(S.stream);
    poll_stream();
  }
  if (!S.done) return 0;
  const size_t words = bytes / 4;
  if (S.out_pos + words > S.n) return 0;
  memcpy(buf, S.h_c + S.out_pos, bytes);
  S.out_pos += words;
  return 1;
}
