#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void*p){return (uint32_t)__cvta_generic_to_shared(p);}
__device__ __forceinline__ void mbar_init(uint32_t b,uint32_t c){asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"::"r"(b),"r"(c):"memory");}
__device__ __forceinline__ void expect(uint32_t b,uint32_t n){asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(b),"r"(n):"memory");}
__device__ __forceinline__ bool tryw(uint32_t b,uint32_t p){uint32_t d;asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(d):"r"(b),"r"(p):"memory");return d;}
__device__ __forceinline__ void bulk(uint32_t dst,const void*src,uint32_t n,uint32_t bar){asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(dst),"l"(src),"r"(n),"r"(bar):"memory");}
// Every CTA runs `nst` stages of 3-deep ring; per stage it pulls `pa` bytes from its PRIVATE region (HBM stream) and
// `pb` bytes from a region SHARED by `share` consecutive CTAs (the dW kernel's dout tiles: L2 resident, read by
// every m-tile CTA of a k-split).
__global__ void k(const uint8_t*priv,const uint8_t*shared_,int nst,int pa,int pb,int share,size_t priv_per,size_t shr_per,unsigned long long*sink){
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar[3];
  if(threadIdx.x==0){for(int s=0;s<3;++s)mbar_init(su32(&bar[s]),1);asm volatile("fence.mbarrier_init.release.cluster;":::"memory");}
  __syncthreads();
  if(threadIdx.x==0){
    const uint8_t*a=priv+(size_t)blockIdx.x*priv_per;const uint8_t*b=shared_+(size_t)(blockIdx.x/share)*shr_per;
    const int sb=pa+pb;
    auto issue=[&](int i){const int s=i%3;const uint32_t br=su32(&bar[s]);expect(br,sb);
      if(pa)bulk(su32(sm)+s*sb,a+(size_t)i*pa,pa,br); if(pb)bulk(su32(sm)+s*sb+pa,b+(size_t)i*pb,pb,br);};
    for(int i=0;i<3&&i<nst;++i)issue(i);
    for(int i=0;i<nst;++i){const int s=i%3;const uint32_t ph=(i/3)&1;while(!tryw(su32(&bar[s]),ph)){} if(i+3<nst)issue(i+3);}
  }
  if(threadIdx.x==0)sink[blockIdx.x]=sm[0];
}
int main(){const int ctas=148;const int nst=2048;uint8_t*priv,*shr;const size_t pp=(size_t)nst*32768;cudaMalloc(&priv,pp*ctas);cudaMalloc(&shr,pp*13);cudaMemset(priv,1,pp*ctas);cudaMemset(shr,1,pp*13);
  unsigned long long*sink;cudaMalloc(&sink,8*ctas);cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,200*1024);
  cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);
  struct C{int pa,pb,share;const char*what;};C cs[]={{32768,0,1,"private only (HBM)"},{0,32768,12,"shared only, 12 CTAs per region"},{0,32768,148,"shared only, all CTAs one region"},{0,32768,2,"shared only, 2 CTAs per region"},{32768,32768,12,"private + shared(12)"},{32768,32768,1,"private + private"},{32768,16384,12,"private 32K + shared(12) 16K"}};
  for(auto c:cs){float best=1e9;for(int rep=0;rep<3;++rep){cudaEventRecord(e0);k<<<ctas,32,3*(c.pa+c.pb)>>>(priv,shr,nst,c.pa,c.pb,c.share,pp,pp,sink);cudaEventRecord(e1);cudaEventSynchronize(e1);float ms;cudaEventElapsedTime(&ms,e0,e1);if(ms<best)best=ms;}
    printf("%-36s: %.3f ms, SM fill %.2f TB/s  %s\n",c.what,best,(double)(c.pa+c.pb)*nst*ctas/best/1e9,cudaGetErrorString(cudaGetLastError()));}
  return 0;}
