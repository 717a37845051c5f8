// accuracy probe for the branch-free reciprocal / rsqrt used by the optics kernels
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ __forceinline__ double rcp0(double x){ double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ __forceinline__ double rsq0(double x){ double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__device__ double rcp_newton2(double x){ double r=rcp0(x); double e=fma(-x,r,1.0); r=fma(r,e,r); e=fma(-x,r,1.0); r=fma(r,e,r); return r; }
__device__ double rcp_cubic(double x){ double r=rcp0(x); double e=fma(-x,r,1.0); double t=fma(e,e,e); return fma(r,t,r); }
__device__ double rsq_newton2(double x){ double y=rsq0(x); for(int i=0;i<2;++i){ double t=x*y; double e=fma(-t,y,1.0); y=fma(0.5*y,e,y);} return y; }
__device__ double rsq_cubic(double x){ double y=rsq0(x); double t=x*y; double e=fma(-t,y,1.0); double p=fma(0.375,e,0.5); p*=e; return fma(y,p,y); }
__device__ double sqrt_cubic(double x){ double y=rsq_cubic(x); double s=x*y; double r=fma(-s,s,x); return fma(r,0.5*y,s); }
__global__ void k(double* out, int n){
  int i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=n) return;
  unsigned long long s=0x9E3779B97F4A7C15ull*(i+1); s^=s>>29; s*=0xBF58476D1CE4E5B9ull; s^=s>>32;
  double m=1.0+(double)(s&0xFFFFFFFFFFFFFull)/4503599627370496.0; int e=(int)((s>>52)&63)-32; double x=ldexp(m,e);
  double ax=x; if (s&(1ull<<60)) x=-x;
  double ex=1.0/x, es=1.0/sqrt(ax), eq=sqrt(ax);
  out[0*(size_t)n+i]=fabs(rcp_newton2(x)-ex)/fabs(ex);
  out[1*(size_t)n+i]=fabs(rcp_cubic(x)-ex)/fabs(ex);
  out[2*(size_t)n+i]=fabs(rsq_newton2(ax)-es)/es;
  out[3*(size_t)n+i]=fabs(rsq_cubic(ax)-es)/es;
  out[4*(size_t)n+i]=fabs(sqrt_cubic(ax)-eq)/eq;
}
int main(){ int n=1<<22; const int C=5; double* d; cudaMalloc(&d,8ull*C*n); k<<<n/256,256>>>(d,n); double* h=(double*)malloc(8ull*C*n); cudaMemcpy(h,d,8ull*C*n,cudaMemcpyDeviceToHost);
 const char* names[C]={"rcp newton2","rcp cubic","rsqrt newton2","rsqrt cubic","sqrt via rsqrt cubic"};
 for(int c=0;c<C;++c){ double mx=0; for(int i=0;i<n;++i) if(h[(size_t)c*n+i]>mx) mx=h[(size_t)c*n+i]; printf("%-22s max rel err %.3e\n", names[c], mx);} return 0; }
