#define _GNU_SOURCE
#include "hostwire.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now(void){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+1e-9*t.tv_nsec;}
int main(int argc,char**argv){
  size_t words=(size_t)1<<24; /* 2^16 polys * 256 */
  int32_t *a=aligned_alloc(64,words*4),*b=aligned_alloc(64,words*4),*c=aligned_alloc(64,words*4);
  uint16_t *a16=aligned_alloc(64,words*2),*b16=aligned_alloc(64,words*2);
  for(size_t i=0;i<words;i++){a[i]=rand()%12289;b[i]=rand()%12289;}
  memset(c,0,words*4);memset(a16,0,words*2);memset(b16,0,words*2);
  printf("threads %d\n",nttb200_wire_threads());
  size_t chunk=(size_t)1<<20;
  for(int rep=0;rep<5;rep++){
    uint32_t mask=0;
    nttb200_wire_begin();
    double t0=now();
    for(size_t o=0;o<words;o+=chunk){
      uint64_t j1=nttb200_wire_post_narrow(a16+o,a+o,chunk,&mask);
      uint64_t j2=nttb200_wire_post_narrow(b16+o,b+o,chunk,&mask);
      nttb200_wire_wait(j1);nttb200_wire_wait(j2);
    }
    double t1=now();
    for(size_t o=0;o<words;o+=chunk){
      uint64_t j=nttb200_wire_post_widen(c+o,a16+o,chunk);
      nttb200_wire_wait(j);
    }
    double t2=now();
    nttb200_wire_end();
    printf("narrow a,b: %.3f ms (%.1f GB/s read)  widen c: %.3f ms (%.1f GB/s written) mask %x -> %.1f M polymul/s host-side ceiling\n",
      (t1-t0)*1e3, 2*words*4/(t1-t0)/1e9,(t2-t1)*1e3, words*4/(t2-t1)/1e9,mask, 65536/(t2-t0)/1e6);
  }
  for(size_t i=0;i<words;i++) if(c[i]!=a[i]||b16[i]!=b[i]){printf("MISMATCH %zu\n",i);return 1;}
  a[12345]=70000; uint32_t mask=0; nttb200_wire_begin(); nttb200_wire_wait(nttb200_wire_post_narrow(a16,a,words,&mask)); nttb200_wire_end();
  printf("mask after out-of-range word: %x\n",mask);
  return 0;}
