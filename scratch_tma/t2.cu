#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int BW=64, BH=16;
__global__ void k(const __grid_constant__ CUtensorMap pm, float* out, int cx, int cy){
  __shared__ alignas(128) float st[BH*BW];
  #pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if(threadIdx.x==0){ init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if(threadIdx.x==0){
    cde::cp_async_bulk_tensor_2d_global_to_shared(&st, &pm, cx, cy, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(st));
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for(int i=threadIdx.x;i<BW*BH;i+=blockDim.x) out[i]=st[i];
}
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc,char**argv){ int cx=argc>1?atoi(argv[1]):8, cy=argc>2?atoi(argv[2]):4; int l2=argc>3?atoi(argv[3]):0;
  int W=256,H=64;
  std::vector<float> h(W*H); for(int i=0;i<W*H;i++)h[i]=i;
  float*d; cudaMalloc(&d,W*H*4); cudaMemcpy(d,h.data(),W*H*4,cudaMemcpyHostToDevice);
  void*fn=nullptr; cudaDriverEntryPointQueryResult q; cudaError_t ee=cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&fn,cudaEnableDefault,&q);
  printf("entry %d %d %p\n",(int)ee,(int)q,fn);
  CUtensorMap m; cuuint64_t gd[2]={(cuuint64_t)W,(cuuint64_t)H}; cuuint64_t gs[1]={(cuuint64_t)W*4}; cuuint32_t box[2]={BW,BH}; cuuint32_t es[2]={1,1};
  CUresult r=((PFN)fn)(&m,CU_TENSOR_MAP_DATA_TYPE_FLOAT32,2,d,gd,gs,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,(CUtensorMapL2promotion)l2,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d\n",(int)r);
  float*o; cudaMalloc(&o,BW*BH*4); std::vector<float> ho(BW*BH);
  k<<<1,128>>>(m,o,cx,cy);
  cudaError_t e=cudaDeviceSynchronize(); printf("libcu++ variant: %s\n",cudaGetErrorString(e));
  if(e==cudaSuccess){ cudaMemcpy(ho.data(),o,BW*BH*4,cudaMemcpyDeviceToHost); int bad=0; for(int y=0;y<BH;y++)for(int x=0;x<BW;x++) {int gx=x+cx,gy=y+cy; float ex=(gx>=0&&gx<W&&gy>=0&&gy<H)?h[gy*W+gx]:0.f; if(ho[y*BW+x]!=ex)bad++;} printf(" mismatches %d\n",bad);} 
  int drv=0; cudaDriverGetVersion(&drv); int rt=0; cudaRuntimeGetVersion(&rt); printf("driver %d runtime %d\n",drv,rt);
  cudaDeviceProp p; cudaGetDeviceProperties(&p,0); printf("%s cc %d.%d\n",p.name,p.major,p.minor);
  return 0;
}
