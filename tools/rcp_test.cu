#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp0(double x){ double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ __forceinline__ double rsq0(double x){ double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ double frcp(double x, int it){ double r=rcp0(x); for(int i=0;i<it;++i){ double e=fma(-x,r,1.0); r=fma(r,e,r);} return r; }
__device__ double frsq(double x, int it){ double y=rsq0(x); for(int i=0;i<it;++i){ double t=x*y; double e=fma(-t,y,1.0); y=fma(0.5*y,e,y);} return y; }
__global__ void k(double* out, int n){
  int i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n) return;
  // pseudo-random positive doubles across magnitudes
  unsigned long long s=0x9E3779B97F4A7C15ull*(i+1); s^=s>>29; s*=0xBF58476D1CE4E5B9ull; s^=s>>32;
  double m=1.0+(double)(s&0xFFFFFFFFFFFFFull)/4503599627370496.0; int e=(int)((s>>52)&63)-32; double x=ldexp(m,e);
  if (s&(1ull<<60)) x=-x;
  double ax=fabs(x);
  double ex=1.0/x;
  for(int it=0;it<4;++it){ double r=frcp(x,it); out[(size_t)it*n+i]=fabs(r-ex)/fabs(ex); }
  double es=1.0/sqrt(ax);
  for(int it=0;it<4;++it){ double r=frsq(ax,it); out[(size_t)(4+it)*n+i]=fabs(r-es)/es; }
}
int main(){ int n=1<<22; double* d; cudaMalloc(&d,8ull*8*n); k<<<n/256,256>>>(d,n); double* h=(double*)malloc(8ull*8*n); cudaMemcpy(h,d,8ull*8*n,cudaMemcpyDeviceToHost);
 for(int c=0;c<8;++c){ double mx=0; for(int i=0;i<n;++i) if(h[(size_t)c*n+i]>mx) mx=h[(size_t)c*n+i]; printf("%s iters=%d max rel err %.3e\n", c<4?"rcp":"rsqrt", c%4, mx);} return 0; }
