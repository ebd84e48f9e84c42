// Dev probe: minimal 3-D u8 TMA load, descriptor in param space vs global memory.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BW, int BH, int RANK>
__device__ void body(const CUtensorMap* map, int x, int y, int z, unsigned* out) {
  __shared__ __align__(128) unsigned char raw[BW * BH];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(BW * BH) : "memory");
    if (RANK == 3)
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(su32(raw)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(su32(&bar)) : "memory");
    else
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(su32(raw)), "l"(map), "r"(x), "r"(y), "r"(su32(&bar)) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(su32(&bar)), "r"(0) : "memory");
  unsigned s = 0;
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) s += raw[i];
  atomicAdd(out, s);
}
__global__ void k_param(const __grid_constant__ CUtensorMap map, int x, int y, int z, unsigned* out) { body<80, 68, 3>(&map, x, y, z, out); }
__global__ void k_param2d(const __grid_constant__ CUtensorMap map, int x, int y, int z, unsigned* out) { body<80, 68, 2>(&map, x, y, z, out); }
__global__ void k_p80x64(const __grid_constant__ CUtensorMap map, int x, int y, int z, unsigned* out) { body<80, 64, 2>(&map, x, y, z, out); }
__global__ void k_p64x68(const __grid_constant__ CUtensorMap map, int x, int y, int z, unsigned* out) { body<64, 68, 2>(&map, x, y, z, out); }
__global__ void k_p128x68(const __grid_constant__ CUtensorMap map, int x, int y, int z, unsigned* out) { body<128, 68, 2>(&map, x, y, z, out); }
__global__ void k_param64(const __grid_constant__ CUtensorMap map, int x, int y, int z, unsigned* out) { body<64, 64, 2>(&map, x, y, z, out); }
__global__ void k_global(const CUtensorMap* map, int x, int y, int z, unsigned* out) { body<80, 68, 3>(map, x, y, z, out); }
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 256, H = 192, F = 2;
  unsigned char* h = (unsigned char*)malloc(W * H * F);
  for (int i = 0; i < W * H * F; i++) h[i] = (unsigned char)(i * 7 + (i >> 8));
  unsigned char* d; cudaMalloc(&d, W * H * F); cudaMemcpy(d, h, W * H * F, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point: err %d q %d fn %p\n", (int)e, (int)q, fn);
  auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  CUtensorMap map; memset(&map, 0, sizeof(map));
  cuuint64_t dims[3] = {W, H, F}; cuuint64_t strides[2] = {W, (cuuint64_t)W * H};
  cuuint32_t box[3] = {80, 68, 1}; cuuint32_t es[3] = {1, 1, 1};
  int rank = 3;
  if (variant >= 1 && variant != 3) rank = 2;
  if (variant == 2 || variant == 5 || variant >= 9) { box[0] = 64; box[1] = 64; }
  if (variant == 6) { box[0] = 80; box[1] = 64; }
  if (variant == 7) { box[0] = 64; box[1] = 68; }
  if (variant == 8) { box[0] = 128; box[1] = 68; }
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  unsigned* out; cudaMalloc(&out, 4);
  int x = 62, y = -2, z = 1;
  if (variant >= 1 && variant != 3) z = 0;
  if (variant == 2 || variant == 4 || variant >= 6) { x = 64; y = 64; }
  if (variant == 9) { x = 62; y = 64; }
  if (variant == 10) { x = 64; y = -2; }
  if (variant == 11) { x = 48; y = 64; }
  if (variant == 12) { x = 200; y = 150; }
  if (variant == 13) { x = -16; y = 64; }
  const int BWv = (int)box[0], BHv = (int)box[1];
  unsigned expect = 0;
  for (int r2 = 0; r2 < BHv; r2++) for (int c = 0; c < BWv; c++) { int gx = x + c, gy = y + r2; if (gx >= 0 && gx < W && gy >= 0 && gy < H) expect += h[(size_t)z * W * H + gy * W + gx]; }
  cudaMemset(out, 0, 4);
  if (variant == 3) {
    CUtensorMap* dm; cudaMalloc(&dm, sizeof(map)); cudaMemcpy(dm, &map, sizeof(map), cudaMemcpyHostToDevice);
    k_global<<<1, 128>>>(dm, x, y, z, out);
  } else if (variant == 1) k_param2d<<<1, 128>>>(map, x, y, z, out);
  else if (variant == 2 || variant == 5 || variant >= 9) k_param64<<<1, 128>>>(map, x, y, z, out);
  else if (variant == 4) k_param2d<<<1, 128>>>(map, x, y, z, out);
  else if (variant == 6) k_p80x64<<<1, 128>>>(map, x, y, z, out);
  else if (variant == 7) k_p64x68<<<1, 128>>>(map, x, y, z, out);
  else if (variant == 8) k_p128x68<<<1, 128>>>(map, x, y, z, out);
  else k_param<<<1, 128>>>(map, x, y, z, out);
  e = cudaDeviceSynchronize(); unsigned v = 0; cudaMemcpy(&v, out, 4, cudaMemcpyDeviceToHost);
  printf("variant %d: err %d (%s) sum %u expect %u\n", variant, (int)e, cudaGetErrorString(e), v, expect);
  return e != cudaSuccess;
  CUtensorMap* dm; cudaMalloc(&dm, sizeof(map)); cudaMemcpy(dm, &map, sizeof(map), cudaMemcpyHostToDevice);
  cudaMemset(out, 0, 4);
  k_global<<<1, 128>>>(dm, x, y, z, out);
  e = cudaDeviceSynchronize(); cudaMemcpy(&v, out, 4, cudaMemcpyDeviceToHost);
  printf("global-memory descriptor: err %d (%s) sum %u expect %u\n", (int)e, cudaGetErrorString(e), v, expect);
  return 0;
}
