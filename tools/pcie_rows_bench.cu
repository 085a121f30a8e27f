// tools/pcie_rows_bench.cu — microbenchmark behind the line-granular PCIe fetch of collect_kernel: SM reads of scattered
// 324-byte rows from pinned host memory with (A) 4-byte lane = class loads, (B) 4-byte loads over whole 128-byte lines,
// (C) one 16-byte load per lane over whole lines, (F) one bulk asynchronous copy (cp.async.bulk, the TMA engine) per row
// span into shared memory -- whole 128-byte lines or the row's own 16-byte-aligned bytes -- against a DMA of the same bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -include algorithm -o pcie_rows_bench tools/pcie_rows_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__);exit(1);} }while(0)
constexpr int C=81;
__device__ __forceinline__ float ldnc(const float* p){float r; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];":"=f"(r):"l"(p)); return r;}
// A: lane = class, 3 loads per row, 4 rows in flight per warp
__global__ void kA(const float* host, const int* rows, int nrows, float* out){
  int warp=(blockIdx.x*blockDim.x+threadIdx.x)>>5, lane=threadIdx.x&31, nw=(gridDim.x*blockDim.x)>>5; float acc=0;
  for(int r0=warp*4;r0<nrows;r0+=nw*4){ float v[4][3];
    #pragma unroll
    for(int k=0;k<4;++k){ int r=r0+k<nrows?rows[r0+k]:rows[0]; const float* p=host+(size_t)r*C;
      #pragma unroll
      for(int s=0;s<3;++s){int c=s*32+lane; v[k][s]= c<C? ldnc(p+c):0.f;} }
    #pragma unroll
    for(int k=0;k<4;++k) acc+=v[k][0]+v[k][1]+v[k][2]; }
  if(acc==123.456f) out[0]=acc;
}
// B: whole aligned 128-byte lines covering the row (4 lines), 4 rows in flight
__global__ void kB(const float* host, const int* rows, int nrows, float* out){
  int warp=(blockIdx.x*blockDim.x+threadIdx.x)>>5, lane=threadIdx.x&31, nw=(gridDim.x*blockDim.x)>>5; float acc=0;
  for(int r0=warp*4;r0<nrows;r0+=nw*4){ float v[4][4];
    #pragma unroll
    for(int k=0;k<4;++k){ int r=r0+k<nrows?rows[r0+k]:rows[0]; size_t f=(size_t)r*C; size_t l0=f&~(size_t)31; int nl=(int)((f+C-1-l0)/32)+1;
      #pragma unroll
      for(int s=0;s<4;++s) v[k][s]= s<nl? ldnc(host+l0+s*32+lane):0.f; }
    #pragma unroll
    for(int k=0;k<4;++k) acc+=v[k][0]+v[k][1]+v[k][2]+v[k][3]; }
  if(acc==123.456f) out[0]=acc;
}
// C: like B but 16-byte loads per lane: 8 lanes cover a 128-byte line, one instruction covers 4 lines = 512 B
__global__ void kC(const float* host, const int* rows, int nrows, float* out){
  int warp=(blockIdx.x*blockDim.x+threadIdx.x)>>5, lane=threadIdx.x&31, nw=(gridDim.x*blockDim.x)>>5; float acc=0;
  for(int r0=warp*8;r0<nrows;r0+=nw*8){ float4 v[8];
    #pragma unroll
    for(int k=0;k<8;++k){ int r=r0+k<nrows?rows[r0+k]:rows[0]; size_t f=(size_t)r*C; size_t l0=f&~(size_t)31; int nl=(int)((f+C-1-l0)/32)+1;
      v[k]= (lane>>3)<nl ? __ldg(reinterpret_cast<const float4*>(host+l0)+lane): make_float4(0,0,0,0); }
    #pragma unroll
    for(int k=0;k<8;++k) acc+=v[k].x+v[k].y+v[k].z+v[k].w; }
  if(acc==123.456f) out[0]=acc;
}
// D: one 16-byte item (a loc vector) per thread, scattered
__global__ void kD(const float4* host, const int* rows, int nrows, float* out){
  int t=blockIdx.x*blockDim.x+threadIdx.x, nt=gridDim.x*blockDim.x; float acc=0;
  for(int r=t;r<nrows;r+=nt){ float4 v=__ldg(host+rows[r]); acc+=v.x+v.y+v.z+v.w; }
  if(acc==123.456f) out[0]=acc;
}
// E: the whole 128-byte line around the item, 8 lanes x 16 bytes per item (4 items per warp instruction)
__global__ void kE(const float4* host, const int* rows, int nrows, float* out){
  int t=blockIdx.x*blockDim.x+threadIdx.x, nt=gridDim.x*blockDim.x; float acc=0; int sub=threadIdx.x&7;
  for(int r=t>>3;r<nrows;r+=nt>>3){ size_t i=(size_t)rows[r]; float4 v=__ldg(host+(i&~(size_t)7)+sub); acc+=v.x+v.y+v.z+v.w; }
  if(acc==123.456f) out[0]=acc;
}
// F: one bulk copy per row into shared memory, kInFlight rows per CTA behind one mbarrier; `lines`: the aligned 128-byte
//    lines that cover the row (512 B), else the 16-byte-aligned span (336 - 352 B)
constexpr int kInFlight=32;
__global__ void kF(const float* host, const int* rows, int nrows, float* out, int lines){
  __shared__ __align__(128) unsigned char buf[kInFlight][512];
  __shared__ __align__(8) unsigned long long bar;
  const int tid=threadIdx.x; float acc=0;
  if(tid==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"((unsigned)__cvta_generic_to_shared(&bar))); asm volatile("fence.mbarrier_init.release.cluster;":::"memory"); }
  __syncthreads();
  unsigned parity=0;
  for(int r0=blockIdx.x*kInFlight;r0<nrows;r0+=gridDim.x*kInFlight){
    if(tid==0){
      unsigned tot=0; unsigned nb[kInFlight]; const char* src[kInFlight]; int cnt=min(kInFlight,nrows-r0);
      for(int k=0;k<cnt;++k){ size_t f=(size_t)rows[r0+k]*C*4; size_t a0= lines? (f&~(size_t)127) : (f&~(size_t)15); size_t a1= lines? ((f+C*4+127)&~(size_t)127) : ((f+C*4+15)&~(size_t)15);
        src[k]=reinterpret_cast<const char*>(host)+a0; nb[k]=(unsigned)(a1-a0); tot+=nb[k]; }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"((unsigned)__cvta_generic_to_shared(&bar)),"r"(tot):"memory");
      for(int k=0;k<cnt;++k)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"((unsigned)__cvta_generic_to_shared(buf[k])),"l"(src[k]),"r"(nb[k]),"r"((unsigned)__cvta_generic_to_shared(&bar)):"memory");
    }
    unsigned ok=0; while(!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }":"=r"(ok):"r"((unsigned)__cvta_generic_to_shared(&bar)),"r"(parity):"memory");
    parity^=1;
    for(int i=tid;i<kInFlight*128;i+=blockDim.x) acc+=reinterpret_cast<const float*>(buf)[i];
    __syncthreads();
  }
  if(acc==123.456f) out[0]=acc;
}
int main(){
  const size_t total=(size_t)32*16320; float* h; CK(cudaHostAlloc(&h,total*C*4,cudaHostAllocDefault));
  for(size_t i=0;i<total*C;i+=1024) h[i]=1.f;
  int nrows=22560*4; std::vector<int> rows(nrows); srand(1); for(int i=0;i<nrows;++i) rows[i]=(int)(((size_t)rand()*7919+i*23)%total);
  // sorted-ish like real passing anchors: ascending order
  std::sort(rows.begin(),rows.end());
  int* drows; CK(cudaMalloc(&drows,nrows*4)); CK(cudaMemcpy(drows,rows.data(),nrows*4,cudaMemcpyHostToDevice));
  float* out; CK(cudaMalloc(&out,4)); float* dev; CK(cudaMalloc(&dev, 64<<20));
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  for(int grid: {148,296,592,1184}) for(int rep=0;rep<2;++rep){
    float ms;
    cudaEventRecord(a); kA<<<grid,256>>>(h,drows,nrows,out); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
    printf("grid %4d A lane=class   : %.3f ms  %.1f GB/s useful\n",grid,ms,nrows*324.0/ms/1e6);
    cudaEventRecord(a); kB<<<grid,256>>>(h,drows,nrows,out); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
    printf("grid %4d B full lines   : %.3f ms  %.1f GB/s useful\n",grid,ms,nrows*324.0/ms/1e6);
    cudaEventRecord(a); kC<<<grid,256>>>(h,drows,nrows,out); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
    printf("grid %4d C float4 lines : %.3f ms  %.1f GB/s useful\n",grid,ms,nrows*324.0/ms/1e6);
  }
  for(int lines: {1,0}) for(int grid: {148,296,592,1184}) for(int thr: {32,128}){ float ms;
    kF<<<grid,thr>>>(h,drows,nrows,out,lines);
    cudaEventRecord(a); kF<<<grid,thr>>>(h,drows,nrows,out,lines); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
    printf("grid %4d x%3d F bulk copy %s : %.3f ms  %.1f GB/s useful\n",grid,thr,lines?"128-B lines":"16-B span  ",ms,nrows*324.0/ms/1e6);
  }
  { // loc vectors: items of 16 bytes among total anchors
    for(int rep=0;rep<3;++rep){ float ms;
      cudaEventRecord(a); kD<<<592,256>>>((const float4*)h,drows,nrows,out); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
      printf("D 16-byte items        : %.3f ms  %.1f M items/s\n",ms,nrows/ms/1e3);
      cudaEventRecord(a); kE<<<592,256>>>((const float4*)h,drows,nrows,out); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b);
      printf("E 128-byte lines/item  : %.3f ms  %.1f M items/s\n",ms,nrows/ms/1e3); }
  }
  // DMA reference
  for(int rep=0;rep<2;++rep){ float ms; cudaEventRecord(a); cudaMemcpyAsync(dev,h,(size_t)nrows*324,cudaMemcpyHostToDevice); cudaEventRecord(b); CK(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms,a,b); printf("DMA same bytes contiguous: %.3f ms %.1f GB/s\n",ms,nrows*324.0/ms/1e6);}
  return 0;
}
