/*
 * terasic_client.c -- test driver that talks to `terasic_pcie_qsys.so` the way the reference's
 * host program does (protocol of Software_Hardware_Comunnicator/linux_app/
 * NTT_PCIECommunicationv2.c:43-51,166-224 and loader of linux_app/PCIE.c:59-103), restated
 * for arbitrary (n, q) and inputs read from a file.
 *
 *   terasic_client <plugin.so> <n> <q> <in.txt> <out.txt>
 * in.txt: 2n decimals (A then B); out.txt: n decimals (C).  Exit code 0 = protocol completed.
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef int BOOL;
typedef unsigned int DWORD;
typedef unsigned short WORD;
typedef int PCIE_HANDLE;

typedef PCIE_HANDLE (*open_t)(WORD, WORD, WORD);
typedef void (*close_t)(PCIE_HANDLE);
typedef BOOL (*rd32_t)(PCIE_HANDLE, int, DWORD, DWORD *);
typedef BOOL (*wr32_t)(PCIE_HANDLE, int, DWORD, DWORD);
typedef BOOL (*fifo_t)(PCIE_HANDLE, DWORD, void *, DWORD);

static const char *SYMS[12] = {"PCIE_Open", "PCIE_Close", "PCIE_Read32", "PCIE_Write32", "PCIE_Read16",
                               "PCIE_Write16", "PCIE_Read8", "PCIE_Write8", "PCIE_DmaWrite", "PCIE_DmaRead",
                               "PCIE_DmaFifoWrite", "PCIE_DmaFifoRead"};

static wr32_t wr32;
static rd32_t rd32;

static void command(PCIE_HANDLE h, int mode) {
  wr32(h, 0, 0x00, (DWORD)(mode << 1) | 1u);     /* start high with the mode */
  wr32(h, 0, 0x00, (DWORD)(mode << 1));          /* start low */
}
static int wait_mask(PCIE_HANDLE h, DWORD mask, DWORD want) {
  for (long i = 0; i < 200000000L; i++) {
    DWORD s = 0;
    if (!rd32(h, 0, 0x20, &s)) return 0;
    if ((s & mask) == want) return 1;
  }
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 6) return 64;
  void *lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  void *fn[12];
  for (int i = 0; i < 12; i++) {
    fn[i] = dlsym(lib, SYMS[i]);
    if (!fn[i]) { fprintf(stderr, "missing symbol %s\n", SYMS[i]); return 3; }
  }
  open_t p_open = (open_t)fn[0];
  close_t p_close = (close_t)fn[1];
  rd32 = (rd32_t)fn[2];
  wr32 = (wr32_t)fn[3];
  fifo_t fifo_write = (fifo_t)fn[10], fifo_read = (fifo_t)fn[11];

  const unsigned n = (unsigned)atoi(argv[2]);
  const uint32_t q = (uint32_t)strtoul(argv[3], NULL, 10);
  uint32_t *a = calloc(n, 4), *b = calloc(n, 4), *c = calloc(n, 4);
  FILE *f = fopen(argv[4], "r");
  if (!f) return 4;
  for (unsigned i = 0; i < n; i++) if (fscanf(f, "%u", &a[i]) != 1) return 5;
  for (unsigned i = 0; i < n; i++) if (fscanf(f, "%u", &b[i]) != 1) return 5;
  fclose(f);

  PCIE_HANDLE h = p_open(0, 0, 0);
  if (!h) { fprintf(stderr, "PCIE_Open failed\n"); return 6; }

  /* board bring-up loopbacks of linux_app/app.c:95-132, 150-265: a register keeps what was
   * written, 128 KiB written by DMA to local memory read back identical */
  {
    typedef BOOL (*dma_t)(PCIE_HANDLE, DWORD, void *, DWORD);
    dma_t dma_write = (dma_t)fn[8], dma_read = (dma_t)fn[9];
    DWORD v = 0;
    if (!wr32(h, 0, 0x100, 0xDEADBEEFu) || !rd32(h, 0, 0x100, &v) || v != 0xDEADBEEFu) return 20;
    const unsigned len = 128 * 1024;
    unsigned char *w = malloc(len), *r = calloc(len, 1);
    for (unsigned i = 0; i < len; i++) w[i] = (unsigned char)(i * 131u + 7u);
    if (!dma_write(h, 0x0, w, len) || !dma_read(h, 0x0, r, len)) return 21;
    for (unsigned i = 0; i < len; i++) if (w[i] != r[i]) return 22;
    free(w); free(r);
  }

  /* mode 0: W || W_INV || q || n_inv in one FIFO write (the twiddle words are whatever the
   * host generated; a GPU back end only needs q) */
  const unsigned wcount = 272;
  uint32_t *params = calloc(2 * wcount + 2, 4);
  params[2 * wcount] = q;
  params[2 * wcount + 1] = 0;
  command(h, 0);
  if (!fifo_write(h, 0x40, params, (2 * wcount + 2) * 4)) return 7;
  if (!wait_mask(h, 1u, 0u)) return 8;
  command(h, 1);
  if (!fifo_write(h, 0x40, a, n * 4)) return 9;
  if (!wait_mask(h, 1u, 0u)) return 10;
  command(h, 2);
  if (!fifo_write(h, 0x40, b, n * 4)) return 11;
  if (!wait_mask(h, 1u, 0u)) return 12;
  command(h, 3);                                   /* GO */
  if (!wait_mask(h, 2u, 2u)) return 13;            /* done_all */
  if (!fifo_read(h, 0x80, c, n * 4)) return 14;

  f = fopen(argv[5], "w");
  if (!f) return 15;
  for (unsigned i = 0; i < n; i++) fprintf(f, "%u\n", c[i]);
  fclose(f);
  p_close(h);
  dlclose(lib);
  return 0;
}
