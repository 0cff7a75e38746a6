#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__device__ __forceinline__ unsigned s32(const void* p){ return (unsigned)__cvta_generic_to_shared(p);} 
template<bool PARAM>
__global__ void k(const __grid_constant__ CUtensorMap pm, const CUtensorMap* gm, float* out, int cx, int cy, int bw, int bh){
  extern __shared__ __align__(128) float st[];
  __shared__ __align__(8) unsigned long long bar;
  if(threadIdx.x==0){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;":::"memory"); }
  __syncthreads();
  if(threadIdx.x==0){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(&bar)),"r"(bw*bh*4):"memory");
    const CUtensorMap* m = PARAM ? &pm : gm;
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(s32(st)),"l"(m),"r"(cx),"r"(cy),"r"(s32(&bar)):"memory");
  }
  unsigned done=0; int spins=0;
  while(!done){ asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}":"=r"(done):"r"(s32(&bar)),"r"(0):"memory"); if(++spins>100000000) { if(threadIdx.x==0) printf("timeout\n"); return;} }
  for(int i=threadIdx.x;i<bw*bh;i+=blockDim.x) out[i]=st[i];
}
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc,char**argv){
  int W=48,H=46, bw=132,bh=33;
  if(argc>2){W=atoi(argv[1]);H=atoi(argv[2]);}
  if(argc>4){bw=atoi(argv[3]);bh=atoi(argv[4]);}
  std::vector<float> h(W*H); for(int i=0;i<W*H;i++)h[i]=i;
  float*d; cudaMalloc(&d,W*H*4); cudaMemcpy(d,h.data(),W*H*4,cudaMemcpyHostToDevice);
  void*fn=nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&fn,cudaEnableDefault,&q);
  CUtensorMap m; cuuint64_t gd[2]={(cuuint64_t)W,(cuuint64_t)H}; cuuint64_t gs[1]={(cuuint64_t)W*4}; cuuint32_t box[2]={(cuuint32_t)bw,(cuuint32_t)bh}; cuuint32_t es[2]={1,1};
  CUresult r=((PFN)fn)(&m,CU_TENSOR_MAP_DATA_TYPE_FLOAT32,2,d,gd,gs,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_L2_256B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode W=%d H=%d box=%dx%d -> %d\n",W,H,bw,bh,(int)r);
  CUtensorMap* gm; cudaMalloc(&gm,sizeof(m)); cudaMemcpy(gm,&m,sizeof(m),cudaMemcpyHostToDevice);
  float*o; cudaMalloc(&o,bw*bh*4); std::vector<float> ho(bw*bh);
  for(int variant=0;variant<2;variant++){
    cudaMemset(o,0,bw*bh*4);
    if(variant==0) k<true><<<1,128,bw*bh*4>>>(m,gm,o,-3,-2,bw,bh); else k<false><<<1,128,bw*bh*4>>>(m,gm,o,-3,-2,bw,bh);
    cudaError_t e=cudaDeviceSynchronize(); printf("variant %s: %s\n",variant==0?"param":"global",cudaGetErrorString(e));
    if(e!=cudaSuccess) return 1;
    cudaMemcpy(ho.data(),o,bw*bh*4,cudaMemcpyDeviceToHost);
    // expect out[y][x] = h[(y-2)*W + (x-3)] inside, 0 outside
    int bad=0; for(int y=0;y<bh;y++)for(int x=0;x<bw;x++){int gx=x-3,gy=y-2; float ex=(gx>=0&&gx<W&&gy>=0&&gy<H)?h[gy*W+gx]:0.f; if(ho[y*bw+x]!=ex)bad++;}
    printf("  mismatches %d\n",bad);
  }
  return 0;
}
