#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void*p){return (uint32_t)__cvta_generic_to_shared(p);}
__device__ __forceinline__ void mbar_init(uint32_t b,uint32_t c){asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"::"r"(b),"r"(c):"memory");}
__device__ __forceinline__ void expect(uint32_t b,uint32_t n){asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(b),"r"(n):"memory");}
__device__ __forceinline__ bool tryw(uint32_t b,uint32_t p){uint32_t d;asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}":"=r"(d):"r"(b),"r"(p):"memory");return d;}
__device__ __forceinline__ void bulk(uint32_t dst,const void*src,uint32_t n,uint32_t bar){asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(dst),"l"(src),"r"(n),"r"(bar):"memory");}
// The dW kernel's address pattern: CTA (mt, z) of a grid (MT, Z) reads, for unit u of its k-split, 32 pieces of 1 KB:
//   G + (u>>1)*KBLK*16K + (4*mt + j)*16K + part*8K + kc*2K + (u&1)*1K.   mode 1: the same bytes re-laid out so that each
// CTA's units are contiguous (64 KB per row tile, row tiles back to back).
__global__ void k(const uint8_t*G,int MT,int kblk,int units_per,int mode,unsigned long long*sink){
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar[3];
  const int lane=threadIdx.x;
  if(lane==0){for(int s=0;s<3;++s)mbar_init(su32(&bar[s]),1);asm volatile("fence.mbarrier_init.release.cluster;":::"memory");}
  __syncthreads();
  const int mt=blockIdx.x%MT,z=blockIdx.x/MT;const int u0=z*units_per;
  const int j=lane>>3,part=(lane>>2)&1,kc=lane&3;
  auto issue=[&](int i){const int s=i%3;const uint32_t br=su32(&bar[s]);if(lane==0)expect(br,32768);__syncwarp();const int u=u0+i;
    const uint8_t*src;
    if(mode==0)src=G+(size_t)(u>>1)*kblk*16384+(size_t)(4*mt+j)*16384+part*8192+kc*2048+(u&1)*1024;
    else src=G+((size_t)mt*(gridDim.x/MT)*units_per+(size_t)u)*32768+(size_t)lane*1024;
    bulk(su32(sm)+s*32768+lane*1024,src,1024,br);};
  for(int i=0;i<3&&i<units_per;++i)issue(i);
  for(int i=0;i<units_per;++i){const int s=i%3;const uint32_t ph=(i/3)&1;while(!tryw(su32(&bar[s]),ph)){} __syncwarp(); if(i+3<units_per)issue(i+3);}
  if(lane==0)sink[blockIdx.x]=sm[0];
}
int main(){const int MT=24,kblk=96,Z=19,units_per=126;  // ck=3072: 24 m tiles, 96 k-blocks; n = 19*126*64 = 153216 rows
  const size_t bytes=(size_t)(Z*units_per/2+1)*kblk*16384;uint8_t*G;cudaMalloc(&G,bytes);cudaMemset(G,1,bytes);unsigned long long*sink;cudaMalloc(&sink,8*MT*Z);
  cudaFuncSetAttribute(k,cudaFuncAttributeMaxDynamicSharedMemorySize,120*1024);
  cudaEvent_t e0,e1;cudaEventCreate(&e0);cudaEventCreate(&e1);
  for(int cfg=0;cfg<4;++cfg){const int mode=cfg&1;const size_t smem_req=(cfg<2)?3*32768:120*1024;float best=1e9;for(int rep=0;rep<3;++rep){cudaEventRecord(e0);k<<<MT*Z,32,smem_req>>>(G,MT,kblk,units_per,mode,sink);cudaEventRecord(e1);cudaEventSynchronize(e1);float ms;cudaEventElapsedTime(&ms,e0,e1);if(ms<best)best=ms;}
    printf("%s mode %d (%s): %.3f ms  %.2f TB/s  %s   [buffer %.2f GB]\n",(cfg<2)?"2 CTAs/SM":"1 CTA/SM ",mode,mode?"contiguous per CTA":"dW tile pattern",best,(double)MT*Z*units_per*32768/best/1e9,cudaGetErrorString(cudaGetLastError()),bytes/1e9);}
  return 0;}
