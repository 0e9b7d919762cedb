/* oracle_sanitize.c — the CPU oracle under AddressSanitizer + UBSan with extreme settings (NaN, inf,
 * arbitrary float bit patterns), every format mapping, and LUTs holding nan / inf / huge entries with
 * non-identity domains, in all three interpolation modes: the checker itself must be free of
 * undefined behaviour (float -> integer conversions above all) for its verdicts to mean anything. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../oracle/vf_oracle.h"
static uint64_t st = 12345;
static uint64_t nx(void){ uint64_t z=(st+=0x9E3779B97F4A7C15ull); z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
static float weird(void){ static const float v[]={0,-0.0f,1,-1,0.5f,360,-360,720,1e30f,-1e30f,INFINITY,-INFINITY,NAN,1e-40f,179.99999f,3.4e38f,-3.4e38f}; uint64_t r=nx(); if(r%3==0){ uint32_t b=(uint32_t)nx(); float f; memcpy(&f,&b,4); return f;} return v[r%(sizeof v/sizeof *v)]; }
int main(void){
  enum{W=64,H=16};
  uint8_t *a=malloc(W*H*8),*b=malloc(W*H*8);
  for(int it=0;it<3000;it++){
    for(int i=0;i<W*H*8;i++) a[i]=(uint8_t)nx();
    orc_hsvfilter_params p={weird(),weird(),weird(),weird(),weird()};
    for(int f=0;f<=9;f++){ memcpy(b,a,W*H*4); if(orc_hsvfilter_frame(b,(size_t)W*(f>=8?3:4),W,H,f,&p)) return 2; }
    orc_hsvdetector_params d={weird(),weird(),weird(),weird(),weird(),weird()};
    int ins[]={1,2,4,6,8,9}, outs[]={0,3,5,7};
    for(int i=0;i<6;i++)for(int o=0;o<4;o++) if(orc_hsvdetector_frame(a,(size_t)W*(ins[i]>=8?3:4),ins[i],b,W*4,outs[o],W,H,&d)) return 3;
  }
  /* LUTs with wild values and domains */
  for(int it=0;it<400;it++){
    char txt[8192]; int n=0; int size=2+(int)(nx()%3); int is3=nx()&1;
    n+=snprintf(txt+n,sizeof txt-n,"LUT_%dD_SIZE %d\nDOMAIN_MIN %g %g %g\nDOMAIN_MAX %g %g %g\n",is3?3:1,size,(double)(nx()%3)-1.0,0.0,-0.5,1.0+(double)(nx()%3),1.0,2.0);
    int cnt=is3?size*size*size:size;
    for(int i=0;i<cnt;i++){ const char*w[]={"nan","inf","-inf","1e38","-1e38","0.5","2","-1"}; n+=snprintf(txt+n,sizeof txt-n,"%s %s %s\n",w[nx()%8],w[nx()%8],w[nx()%8]); }
    orc_cube c; char err[256];
    if(orc_cube_parse(txt,(size_t)n,&c,err,sizeof err)) continue;
    for(int i=0;i<W*H*8;i++) a[i]=(uint8_t)nx();
    for(int mode=0;mode<3;mode++){
      if(orc_colorlut_frame_ex(&c,a,W*4,b,W*4,W,H,0,mode)) return 4;
      if(orc_colorlut_frame_ex(&c,a,W*8,b,W*8,W,H,10,mode)) return 5;
      if(orc_colorlut_frame_ex(&c,a,W*8,b,W*8,W,H,11,mode)) return 6;
    }
    orc_cube_free(&c);
  }
  free(a); free(b); puts("oracle_sanitize: ok"); return 0; }
