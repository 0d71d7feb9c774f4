/*
 * ref_wrappers_client.c -- a program written against the REFERENCE's own n = 256 headers
 * (NTT/ntt256.h, NTT-RED/ntt_red256.h: `static inline` wrappers that bind (256, table) to the
 * generic transforms), compiled with -I into the unmodified reference tree and linked with
 * -lnttb200 ONLY: every function and every table symbol it uses must come from the GPU library.
 * Built by oracle/Makefile into oracle/_ref/ref_wrappers_gpu where the reference tree exists.
 *
 * usage: ref_wrappers_gpu <id> <in.txt> <out.txt>     (256 integers each; ids as in
 *        oracle/ref_shim.c:ref_transform -- 0..12 plain, 100..111 Longa-Naehrig;
 *        200..203 = the four products on the pair in.txt holds (512 integers))
 */
#include <stdio.h>
#include <stdlib.h>

#include "NTT/ntt256.h"
#include "NTT-RED/ntt_red256.h"

static int run(int id, int32_t *a) {
  switch (id) {
    case 0:  ntt256_ct_rev2std(a); return 0;
    case 1:  ntt256_gs_rev2std(a); return 0;
    case 2:  ntt256_ct_std2rev(a); return 0;
    case 3:  ntt256_gs_std2rev(a); return 0;
    case 4:  intt256_ct_rev2std(a); return 0;
    case 5:  intt256_gs_rev2std(a); return 0;
    case 6:  intt256_ct_std2rev(a); return 0;
    case 7:  intt256_gs_std2rev(a); return 0;
    case 8:  mulntt256_ct_rev2std(a); return 0;
    case 9:  mulntt256_ct_std2rev(a); return 0;
    case 10: inttmul256_gs_rev2std(a); return 0;
    case 11: inttmul256_gs_std2rev(a); return 0;
    case 12: ntt_ct_rev2std_v1(a, 256, ntt256_psi_powers); return 0;
    case 100: ntt_red256_ct_rev2std(a); return 0;
    case 101: ntt_red256_gs_rev2std(a); return 0;
    case 102: ntt_red256_ct_std2rev(a); return 0;
    case 103: ntt_red256_gs_std2rev(a); return 0;
    case 104: intt_red256_ct_rev2std(a); return 0;
    case 105: intt_red256_gs_rev2std(a); return 0;
    case 106: intt_red256_ct_std2rev(a); return 0;
    case 107: intt_red256_gs_std2rev(a); return 0;
    case 108: mulntt_red256_ct_rev2std(a); return 0;
    case 109: mulntt_red256_ct_std2rev(a); return 0;
    case 110: inttmul_red256_gs_rev2std(a); return 0;
    case 111: inttmul_red256_gs_std2rev(a); return 0;
    default: return -1;
  }
}

int main(int argc, char **argv) {
  if (argc != 4) { fprintf(stderr, "usage: %s id in out\n", argv[0]); return 2; }
  const int id = atoi(argv[1]);
  int32_t a[256], b[256], c[256];
  FILE *f = fopen(argv[2], "r");
  if (!f) return 3;
  for (int i = 0; i < 256; i++) if (fscanf(f, "%d", &a[i]) != 1) return 3;
  if (id >= 200) for (int i = 0; i < 256; i++) if (fscanf(f, "%d", &b[i]) != 1) return 3;
  fclose(f);
  const int32_t *res = a;
  if (id >= 200) {
    switch (id) {
      case 200: ntt256_product1(c, a, b); break;
      case 201: ntt256_product4(c, a, b); break;
      case 202: ntt_red256_product1(c, a, b); break;
      case 203: ntt_red256_product4(c, a, b); break;
      default: return 4;
    }
    res = c;
  } else if (run(id, a)) return 4;
  f = fopen(argv[3], "w");
  if (!f) return 5;
  for (int i = 0; i < 256; i++) fprintf(f, "%d\n", res[i]);
  fclose(f);
  return 0;
}
