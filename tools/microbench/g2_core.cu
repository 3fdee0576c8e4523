// Micro-benchmark: the register-resident 2-qubit gate core of rs_pass.cu in isolation (no HBM, no
// shared memory): 16 complex128 amplitudes per thread, gate matrices in __constant__ memory reaching
// the DFMAs as uniform-register operands, one switch dispatch per op.  Tells how close the inner
// loop alone gets to the FP64 peak (18.4 TFMA/s measured by fp64_pipes.cu).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
struct Op { uint32_t code, moff; };
__constant__ Op c_ops[256];
__constant__ double2 c_mat[2048];

__device__ __forceinline__ double2 cmul(double2 a, double2 b){ double2 r; r.x=a.x*b.x-a.y*b.y; r.y=a.x*b.y+a.y*b.x; return r;}
__device__ __forceinline__ void cfma(double2&acc,double2 a,double2 b){acc.x=fma(a.x,b.x,acc.x);acc.x=fma(-a.y,b.y,acc.x);acc.y=fma(a.x,b.y,acc.y);acc.y=fma(a.y,b.x,acc.y);}

template<int A, int R1,int R2>
__device__ __forceinline__ void g2(double2 (&a)[A], const double2* M){
#pragma unroll
  for(int g=0;g<A/4;++g){
    const int t0=((g>>R1)<<(R1+1))|(g&((1<<R1)-1));
    const int i00=((t0>>R2)<<(R2+1))|(t0&((1<<R2)-1));
    const int i01=i00|(1<<R1), i10=i00|(1<<R2), i11=i01|(1<<R2);
    double2 x0=a[i00],x1=a[i01],x2=a[i10],x3=a[i11];
    double2 o[4];
#pragma unroll
    for(int r=0;r<4;++r){
      o[r]=cmul(M[4*r],x0); cfma(o[r],M[4*r+1],x1); cfma(o[r],M[4*r+2],x2); cfma(o[r],M[4*r+3],x3);
    }
    a[i00]=o[0];a[i01]=o[1];a[i10]=o[2];a[i11]=o[3];
  }
}
template <int A, int THREADS, int CTAS, bool CONSTMAT>
__global__ void __launch_bounds__(THREADS, CTAS) k(double2* st, const double2* gmat, int nops, int iters){
  double2 a[A];
#pragma unroll
  for(int r=0;r<A;++r) a[r]=st[threadIdx.x+THREADS*r+(size_t)A*THREADS*blockIdx.x];
  for (int it = 0; it < iters; ++it)
  for(int o=0;o<nops;++o){
    const Op op=c_ops[o];
    const double2* M = CONSTMAT ? c_mat+op.moff : gmat + op.moff;
    switch(op.code){
      case 0: g2<A,0,1>(a,M);break;
      case 1: g2<A,0,2>(a,M);break;
      case 2: g2<A,0,3>(a,M);break;
      case 3: g2<A,1,2>(a,M);break;
      case 4: g2<A,1,3>(a,M);break;
      case 5: g2<A,2,3>(a,M);break;
    }
  }
#pragma unroll
  for(int r=0;r<A;++r) st[threadIdx.x+THREADS*r+(size_t)A*THREADS*blockIdx.x]=a[r];
}

template <int A, int THREADS, int CTAS, bool CONSTMAT>
void run(const char* name, const double2* gmat) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int nops = 64, iters = 200;
  double2* st; cudaMalloc(&st, sizeof(double2) * A * THREADS * sms * CTAS);
  cudaMemset(st, 0, sizeof(double2) * A * THREADS * sms * CTAS);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<A,THREADS,CTAS,CONSTMAT><<<sms*CTAS, THREADS>>>(st, gmat, nops, 2);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<A,THREADS,CTAS,CONSTMAT><<<sms*CTAS, THREADS>>>(st, gmat, nops, iters);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)sms*CTAS*THREADS*A*16.0*nops*iters;  // 16 DFMA-equivalents per amplitude per 2q gate
  printf("%-44s %8.3f ms  %.2f TFMA/s (%s)\n", name, ms, fma/ms/1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(st);
}

int main() {
  Op ops[256]; double2* mat = (double2*)malloc(sizeof(double2)*2048);
  srand(1);
  for (int i = 0; i < 256; ++i) { ops[i].code = rand() % 6; ops[i].moff = (rand() % 100) * 16; }
  for (int i = 0; i < 2048; ++i) { mat[i].x = 0.3 * cos(i*0.7); mat[i].y = 0.3 * sin(i*1.3); }
  cudaMemcpyToSymbol(c_ops, ops, sizeof(ops)); cudaMemcpyToSymbol(c_mat, mat, sizeof(double2)*2048);
  double2* gmat; cudaMalloc(&gmat, sizeof(double2)*2048); cudaMemcpy(gmat, mat, sizeof(double2)*2048, cudaMemcpyHostToDevice);
  run<16,256,2,true>("16 amps x 256 thr x 2 CTA, const matrices", gmat);
  run<16,256,2,false>("16 amps x 256 thr x 2 CTA, global matrices", gmat);
  run<16,256,1,true>("16 amps x 256 thr x 1 CTA, const matrices", gmat);
  run<32,128,2,true>("32 amps x 128 thr x 2 CTA, const matrices", gmat);
  run<32,128,3,true>("32 amps x 128 thr x 3 CTA, const matrices", gmat);
  run<8,256,4,true>("8 amps x 256 thr x 4 CTA, const matrices", gmat);
  return 0;
}
