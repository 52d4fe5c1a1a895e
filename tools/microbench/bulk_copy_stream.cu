#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void*p){return (uint32_t)__cvta_generic_to_shared(p);}
__device__ __forceinline__ void mbar_init(uint32_t b,uint32_t c){asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"::"r"(b),"r"(c):"memory");}
__device__ __forceinline__ void expect(uint32_t b,uint32_t n){asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(b),"r"(n):"memory");}
__device__ __forceinline__ bool tryw(uint32_t b,uint32_t p){uint32_t d;asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(d):"r"(b),"r"(p):"memory");return d;}
__device__ __forceinline__ void bulk(uint32_t dst,const void*src,uint32_t n,uint32_t bar){asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(dst),"l"(src),"r"(n),"r"(bar):"memory");}
// NW producer warps per CTA, each with its own ring of STAGES x SB bytes and its own slice of the CTA's region;
// one 16 KB (or SB) bulk copy per stage issued by lane 0 of the warp.
__global__ void k(const uint8_t*src,size_t per_cta,int nw,int stages,int sb,unsigned long long*sink){
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar[64];
  const int warp=threadIdx.x>>5,lane=threadIdx.x&31;
  if(threadIdx.x==0){for(int s=0;s<nw*stages;++s)mbar_init(su32(&bar[s]),1);asm volatile("fence.mbarrier_init.release.cluster;":::"memory");}
  __syncthreads();
  if(warp<nw&&lane==0){
    const size_t per_w=per_cta/nw;const uint8_t*base=src+(size_t)blockIdx.x*per_cta+(size_t)warp*per_w;
    const int nst=(int)(per_w/sb);uint64_t*b=bar+warp*stages;const uint32_t ring=su32(sm)+warp*stages*sb;
    auto issue=[&](int i){const int s=i%stages;expect(su32(&b[s]),sb);bulk(ring+s*sb,base+(size_t)i*sb,sb,su32(&b[s]));};
    for(int i=0;i<stages&&i<nst;++i)issue(i);
    for(int i=0;i<nst;++i){const int s=i%stages;const uint32_t ph=(i/stages)&1;while(!tryw(su32(&b[s]),ph)){} if(i+stages<nst)issue(i+stages);}
  }
  if(threadIdx.x==0)sink[blockIdx.x]=sm[0];
}
int main(){const size_t total=32ull<<30;uint8_t*src;cudaMalloc(&src,total);cudaMemset(src,1,total);unsigned long long*sink;cudaMalloc(&sink,8*148*8);
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,200*1024);
  cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);
  struct C{int cps,nw,st,sb;};C cs[]={{1,1,3,16384},{1,1,6,16384},{1,1,12,16384},{1,1,3,32768},{1,1,6,32768},{1,2,3,32768},{1,4,3,16384},{1,4,6,8192},{2,1,3,32768},{2,1,6,16384},{2,2,3,16384},{4,1,3,16384},{1,1,3,65536},{1,3,1,65536}};
  for(auto c:cs){const int ctas=148*c.cps;const size_t per=(total/ctas)/(c.nw*c.sb)*(c.nw*c.sb);const size_t smem=(size_t)c.nw*c.st*c.sb;
    float best=1e9;for(int rep=0;rep<2;++rep){cudaEventRecord(e0);k<<<ctas,32*c.nw,smem>>>(src,per,c.nw,c.st,c.sb,sink);cudaEventRecord(e1);cudaEventSynchronize(e1);float ms;cudaEventElapsedTime(&ms,e0,e1);if(ms<best)best=ms;}
    printf("ctas/SM %d  warps %d  stages %2d x %5d B (in flight/SM %3zu KB): %.2f TB/s  %s\n",c.cps,c.nw,c.st,c.sb,smem*c.cps/1024,per*ctas/best/1e9,cudaGetErrorString(cudaGetLastError()));}
  return 0;}
