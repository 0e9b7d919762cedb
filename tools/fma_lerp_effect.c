// fma_lerp_effect.c — RGBA64 through a 3D LUT: the reference's operation order against lerps contracted to FMAs
// (DESIGN.md §18 item 5): how many 16-bit output codes change?  CPU only.
//   gcc -O2 -ffp-contract=off -o fma_lerp_effect tools/fma_lerp_effect.c -lm && ./fma_lerp_effect 17 33 65
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
static float six(double v){char b[32];snprintf(b,sizeof b,"%.6f",v<0?0:(v>1?1:v));return strtof(b,0);}
static float *lut;static int N;
static inline float at(int x,int y,int z,int c){return lut[(((size_t)z*N+y)*N+x)*3+c];}
static inline float l_ref(float a,float b,float t){return a+(b-a)*t;}
static inline float l_fma(float a,float b,float t){return fmaf(b-a,t,a);}
static uint64_t s=0x5EED0000u;static uint64_t sm(){uint64_t z=(s+=0x9E3779B97F4A7C15ull);z=(z^(z>>30))*0xBF58476D1CE4E5B9ull;z=(z^(z>>27))*0x94D049BB133111EBull;return z^(z>>31);}
int main(int argc,char**argv){
  for(int ni=1;ni<argc;ni++){N=atoi(argv[ni]);lut=malloc((size_t)N*N*N*3*4);
  for(int z=0;z<N;z++)for(int y=0;y<N;y++)for(int x=0;x<N;x++){double r=(double)x/(N-1),g=(double)y/(N-1),b=(double)z/(N-1);float*e=&lut[(((size_t)z*N+y)*N+x)*3];
    e[0]=six(pow(r,0.8)*0.9+0.1*g);e[1]=six(0.5-0.45*cos(M_PI*g)+0.05*b);e[2]=six(pow(b,1.2)*0.85+0.15*r);}
  size_t n=8000000,diffpx=0,diffch=0;int maxd=0;
  for(size_t i=0;i<n;i++){uint64_t r=sm();uint32_t c[3]={r&0xFFFF,(r>>16)&0xFFFF,(r>>32)&0xFFFF};int i0[3],i1[3];float t[3];
    for(int k=0;k<3;k++){float v=(float)c[k]/65535.0f;float p=v*((float)N-1.0f);int f=(int)floorf(p);if(f>N-1)f=N-1;i0[k]=f;i1[k]=f+1<N?f+1:N-1;t[k]=p-(float)f;}
    int bad=0;
    for(int ch=0;ch<3;ch++){
      float o[2];
      for(int m=0;m<2;m++){float(*L)(float,float,float)=m?l_fma:l_ref;
        float c00=L(at(i0[0],i0[1],i0[2],ch),at(i1[0],i0[1],i0[2],ch),t[0]);
        float c10=L(at(i0[0],i1[1],i0[2],ch),at(i1[0],i1[1],i0[2],ch),t[0]);
        float c01=L(at(i0[0],i0[1],i1[2],ch),at(i1[0],i0[1],i1[2],ch),t[0]);
        float c11=L(at(i0[0],i1[1],i1[2],ch),at(i1[0],i1[1],i1[2],ch),t[0]);
        o[m]=L(L(c00,c10,t[1]),L(c01,c11,t[1]),t[2]);}
      int a=(int)roundf(fminf(fmaxf(o[0],0),1)*65535.0f),b=(int)roundf(fminf(fmaxf(o[1],0),1)*65535.0f);
      if(a!=b){bad=1;diffch++;int d=abs(a-b);if(d>maxd)maxd=d;}
    }
    diffpx+=bad;}
  printf("N=%d: %zu random RGBA64 pixels: %.3f %% of pixels (%.3f %% of channel values) differ, max difference %d code of 65535\n",N,n,100.0*diffpx/n,100.0*diffch/(3.0*n),maxd);
  free(lut);}
}
