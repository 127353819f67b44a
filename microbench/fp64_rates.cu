// Microbenchmark: FP64 pipe facts on B200 that decide the decimator kernel design.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__constant__ double cc[1024];

template<int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[ILP];
  #pragma unroll
  for (int i=0;i<ILP;i++) acc[i]=threadIdx.x*1e-3+i;
  for (int it=0; it<iters; it++) {
    #pragma unroll
    for (int i=0;i<ILP;i++) acc[i]=fma(acc[i],a,b);
  }
  double s=0; 
  #pragma unroll
  for (int i=0;i<ILP;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// DFMA with coefficient operands from constant bank cycling through NC doubles
template<int NC>
__global__ void k_dfma_const(double* out, int iters, double x0) {
  double acc[16];
  #pragma unroll
  for (int i=0;i<16;i++) acc[i]=threadIdx.x*1e-3+i;
  double x = x0;
  for (int it=0; it<iters; it++) {
    #pragma unroll
    for (int c=0;c<NC;c+=16) {
      #pragma unroll
      for (int i=0;i<16;i++) acc[i]=fma(cc[c+i],x,acc[i]);
    }
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<16;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// DFMA with coefficient operands from shared memory (broadcast LDS.128), NC doubles
template<int NC>
__global__ void k_dfma_smem(double* out, int iters, double x0) {
  __shared__ double2 sc[NC/2];
  for (int i=threadIdx.x;i<NC/2;i+=blockDim.x) sc[i]=make_double2(1.0+i*1e-9,1.0-i*1e-9);
  __syncthreads();
  double acc[16];
  #pragma unroll
  for (int i=0;i<16;i++) acc[i]=threadIdx.x*1e-3+i;
  double x = x0;
  for (int it=0; it<iters; it++) {
    #pragma unroll
    for (int c=0;c<NC/2;c+=8) {
      #pragma unroll
      for (int i=0;i<8;i++) { double2 v=sc[c+i]; acc[2*i]=fma(v.x,x,acc[2*i]); acc[2*i+1]=fma(v.y,x,acc[2*i+1]); }
    }
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<16;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// mixed: per 8 DFMA, NI int->double conversions (I2F.F64) feeding an accumulator
template<int NI>
__global__ void k_mix_i2f(double* out, int iters, double a, double b, int seed) {
  double acc[8];
  #pragma unroll
  for (int i=0;i<8;i++) acc[i]=threadIdx.x*1e-3+i;
  int v = seed + threadIdx.x;
  double cs = 0;
  for (int it=0; it<iters; it++) {
    #pragma unroll
    for (int i=0;i<8;i++) acc[i]=fma(acc[i],a,b);
    #pragma unroll
    for (int i=0;i<NI;i++) { v = v*1664525+1013904223; cs = __longlong_as_double(__double_as_longlong(cs) ^ __double_as_longlong((double)(v>>16))); }
  }
  double s=cs;
  #pragma unroll
  for (int i=0;i<8;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// mixed: per 8 DFMA, NI integer ops (LOP/IADD chain)
template<int NI>
__global__ void k_mix_int(double* out, int iters, double a, double b, int seed) {
  double acc[8];
  #pragma unroll
  for (int i=0;i<8;i++) acc[i]=threadIdx.x*1e-3+i;
  unsigned v[4]; for (int i=0;i<4;i++) v[i]= seed + threadIdx.x*i;
  for (int it=0; it<iters; it++) {
    #pragma unroll
    for (int i=0;i<8;i++) acc[i]=fma(acc[i],a,b);
    #pragma unroll
    for (int i=0;i<NI;i++) { v[i&3] = (v[i&3] ^ (v[(i+1)&3]>>3)) + 0x9e3779b9u; }
  }
  double s=v[0]+v[1]+v[2]+v[3];
  #pragma unroll
  for (int i=0;i<8;i++) s+=acc[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// mixed: per 8 DFMA, NF FFMA ops
template<int NF>
__global__ void k_mix_ffma(double* out, int iters, double a, double b, float fa) {
  double acc[8]; float f[8];
  #pragma unroll
  for (int i=0;i<8;i++) { acc[i]=threadIdx.x*1e-3+i; f[i]=threadIdx.x*1e-3f+i; }
  for (int it=0; it<iters; it++) {
    #pragma unroll
    for (int i=0;i<8;i++) acc[i]=fma(acc[i],a,b);
    #pragma unroll
    for (int i=0;i<NF;i++) f[i&7]=fmaf(f[i&7],fa,0.5f);
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<8;i++) s+=acc[i]+f[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// DMMA m8n8k4: ACC independent accumulators per warp
template<int ACC>
__global__ void k_dmma884(double* out, int iters, double a0, double b0) {
  double c[ACC][2];
  #pragma unroll
  for (int i=0;i<ACC;i++){c[i][0]=threadIdx.x*1e-3;c[i][1]=i;}
  double a=a0+threadIdx.x*1e-9, b=b0;
  for (int it=0; it<iters; it++) {
    #pragma unroll
    for (int i=0;i<ACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<ACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// DMMA m16n8k16 (sm_90+): A 8 regs, B 4 regs, C 4 regs
template<int ACC>
__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
  double c[ACC][4];
  #pragma unroll
  for (int i=0;i<ACC;i++){c[i][0]=threadIdx.x*1e-3;c[i][1]=i;c[i][2]=1;c[i][3]=2;}
  double a[8], b[4];
  #pragma unroll
  for (int i=0;i<8;i++) a[i]=a0+threadIdx.x*1e-9+i*1e-10;
  #pragma unroll
  for (int i=0;i<4;i++) b[i]=b0+i*1e-10;
  for (int it=0; it<iters; it++) {
    #pragma unroll
    for (int i=0;i<ACC;i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
        : "d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),
          "d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
  }
  double s=0;
  #pragma unroll
  for (int i=0;i<ACC;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

__global__ void k_sincos(double* out, int iters, double w) {
  double s=0; double x = threadIdx.x*0.37;
  for (int it=0; it<iters; it++) { double sn, cs; sincos(x, &sn, &cs); s+=sn*cs; x+=w; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void k_atan2(double* out, int iters, double w) {
  double s=0; double x = threadIdx.x*0.37+0.1;
  for (int it=0; it<iters; it++) { s+=atan2(x, s+1.0); x+=w; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F> float timeit(F f, int reps=3) {
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for (int r=0;r<reps;r++){ CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms; }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int nsm=p.multiProcessorCount;
  double h[1024]; for(int i=0;i<1024;i++) h[i]=1.0+i*1e-9;
  CK(cudaMemcpyToSymbol(cc,h,sizeof(h)));
  double* out; CK(cudaMalloc(&out, sizeof(double)*nsm*16*1024));
  const int iters=20000;
  for (int wpsm : {4,8,16,32}) {
    int threads=256, blocks=nsm*wpsm*32/threads;
    double n=(double)blocks*threads*iters;
    float ms;
    ms=timeit([&]{k_dfma<8><<<blocks,threads>>>(out,iters,0.999,1e-3);});
    printf("warps/SM %2d  DFMA ilp8      : %8.3f ms  %7.2f TFMA/s (%.2f TFLOP/s)\n", wpsm, ms, n*8/ms/1e9, 2*n*8/ms/1e9);
    ms=timeit([&]{k_dmma884<8><<<blocks,threads>>>(out,iters/8,0.999,1e-3);});
    printf("warps/SM %2d  DMMA m8n8k4    : %8.3f ms  %7.2f TFMA/s\n", wpsm, ms, (double)blocks*threads/32*(iters/8)*8*256/ms/1e9);
    ms=timeit([&]{k_dmma16816<4><<<blocks,threads>>>(out,iters/32,0.999,1e-3);});
    printf("warps/SM %2d  DMMA m16n8k16  : %8.3f ms  %7.2f TFMA/s\n", wpsm, ms, (double)blocks*threads/32*(iters/32)*4*2048/ms/1e9);
  }
  {
    int wpsm=16, threads=256, blocks=nsm*wpsm*32/threads; double n=(double)blocks*threads*iters; float ms;
    ms=timeit([&]{k_dfma<8><<<blocks,threads>>>(out,iters,0.999,1e-3);});
    printf("base DFMA8                    : %8.3f ms\n", ms);
    ms=timeit([&]{k_mix_i2f<1><<<blocks,threads>>>(out,iters,0.999,1e-3,7);}); printf("8 DFMA + 1 I2F.F64 (+imad,xor): %8.3f ms\n", ms);
    ms=timeit([&]{k_mix_i2f<2><<<blocks,threads>>>(out,iters,0.999,1e-3,7);}); printf("8 DFMA + 2 I2F.F64            : %8.3f ms\n", ms);
    ms=timeit([&]{k_mix_i2f<4><<<blocks,threads>>>(out,iters,0.999,1e-3,7);}); printf("8 DFMA + 4 I2F.F64            : %8.3f ms\n", ms);
    ms=timeit([&]{k_mix_int<4><<<blocks,threads>>>(out,iters,0.999,1e-3,7);}); printf("8 DFMA + 4 intops(x2 instr)   : %8.3f ms\n", ms);
    ms=timeit([&]{k_mix_int<8><<<blocks,threads>>>(out,iters,0.999,1e-3,7);}); printf("8 DFMA + 8 intops(x2 instr)   : %8.3f ms\n", ms);
    ms=timeit([&]{k_mix_int<16><<<blocks,threads>>>(out,iters,0.999,1e-3,7);}); printf("8 DFMA + 16 intops(x2 instr)  : %8.3f ms\n", ms);
    ms=timeit([&]{k_mix_ffma<8><<<blocks,threads>>>(out,iters,0.999,1e-3,0.99f);}); printf("8 DFMA + 8 FFMA               : %8.3f ms\n", ms);
    ms=timeit([&]{k_mix_ffma<16><<<blocks,threads>>>(out,iters,0.999,1e-3,0.99f);}); printf("8 DFMA + 16 FFMA              : %8.3f ms\n", ms);
    // const/smem operand: per iter NC fmas per thread
    int it2=iters/16;
    ms=timeit([&]{k_dfma_const<64><<<blocks,threads>>>(out,it2*8,0.999);});  printf("DFMA const-operand NC=64  (512B) : %8.3f ms %7.2f TFMA/s\n", ms, (double)blocks*threads*it2*8*64/ms/1e9);
    ms=timeit([&]{k_dfma_const<256><<<blocks,threads>>>(out,it2*2,0.999);}); printf("DFMA const-operand NC=256 (2KB)  : %8.3f ms %7.2f TFMA/s\n", ms, (double)blocks*threads*it2*2*256/ms/1e9);
    ms=timeit([&]{k_dfma_const<512><<<blocks,threads>>>(out,it2,0.999);});   printf("DFMA const-operand NC=512 (4KB)  : %8.3f ms %7.2f TFMA/s\n", ms, (double)blocks*threads*it2*512/ms/1e9);
    ms=timeit([&]{k_dfma_const<1024><<<blocks,threads>>>(out,it2/2,0.999);});printf("DFMA const-operand NC=1024 (8KB) : %8.3f ms %7.2f TFMA/s\n", ms, (double)blocks*threads*(it2/2)*1024/ms/1e9);
    ms=timeit([&]{k_dfma_smem<512><<<blocks,threads>>>(out,it2,0.999);});    printf("DFMA smem-bcast LDS.128 NC=512   : %8.3f ms %7.2f TFMA/s\n", ms, (double)blocks*threads*it2*512/ms/1e9);
    ms=timeit([&]{k_sincos<<<blocks,threads>>>(out,2000,0.01);}); printf("sincos: %8.3f ms  %7.2f G/s\n", ms, (double)blocks*threads*2000/ms/1e6);
    ms=timeit([&]{k_atan2<<<blocks,threads>>>(out,2000,0.01);}); printf("atan2 : %8.3f ms  %7.2f G/s\n", ms, (double)blocks*threads*2000/ms/1e6);
    (void)n;
  }
  printf("done\n");
  return 0;
}
