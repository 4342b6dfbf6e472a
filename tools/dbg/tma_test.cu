#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
template <int BW, int BH>
__global__ void kp(const __grid_constant__ CUtensorMap mparam, int x0, int y0, int f, uint8_t* out) {
  const CUtensorMap* mp = &mparam;
  __shared__ __align__(128) uint8_t raw[BH][BW];
  __shared__ __align__(8) unsigned long long mbar;
  if (threadIdx.x == 0) {
    unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BW * BH) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(&raw[0][0])), "l"((unsigned long long)mp), "r"(x0), "r"(y0), "r"(f), "r"(bar) : "memory");
  }
  __syncthreads();
  unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar), done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = raw[i / BW][i % BW];
}
template <int BW, int BH>
__global__ void k(const CUtensorMap* mp, int x0, int y0, int f, uint8_t* out) {
  __shared__ __align__(128) uint8_t raw[BH][BW];
  __shared__ __align__(8) unsigned long long mbar;
  if (threadIdx.x == 0) {
    unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BW * BH) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(&raw[0][0])), "l"((unsigned long long)mp), "r"(x0), "r"(y0), "r"(f), "r"(bar) : "memory");
  }
  __syncthreads();
  unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar), done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = raw[i / BW][i % BW];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int BW, int BH> int run(Enc enc, int w, int h, int pitch, int B, int x0, int y0, bool asParam = false) {
  std::vector<uint8_t> img((size_t)pitch * h * B);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 7 + i / pitch);
  uint8_t *d, *o; cudaMalloc(&d, img.size()); cudaMalloc(&o, BW * BH); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
  CUtensorMap m; cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B}; cuuint64_t str[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
  cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("box %dx%d w=%d pitch=%d: encode failed %d\n", BW, BH, w, pitch, (int)r); return 1; }
  CUtensorMap* dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
  if (asParam) kp<BW, BH><<<1, 128>>>(m, x0, y0, B - 1, o); else k<BW, BH><<<1, 128>>>(dm, x0, y0, B - 1, o);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<uint8_t> out(BW * BH); cudaMemcpy(out.data(), o, BW * BH, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < BH; ++r2) for (int c = 0; c < BW; ++c) {
    int gx = x0 + c, gy = y0 + r2; uint8_t exp = (gx < 0 || gx >= w || gy < 0 || gy >= h) ? 0 : img[(size_t)(B - 1) * pitch * h + (size_t)gy * pitch + gx];
    bad += exp != out[r2 * BW + c];
  }
  printf("box %dx%d w=%d h=%d pitch=%d x0=%d y0=%d param=%d: %s mismatches=%d\n", BW, BH, w, h, pitch, x0, y0, (int)asParam, cudaGetErrorString(e), bad);
  if (e != cudaSuccess) { cudaDeviceReset(); cudaFree(0); }
  return 0;
}
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q; cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); Enc enc = (Enc)p;
  run<128, 22>(enc, 640, 480, 640, 2, 128, 16, true);
  run<128, 22>(enc, 640, 480, 640, 2, 125, 13, true);
  run<128, 22>(enc, 640, 480, 640, 2, 128, 16, false);
  if (run<128, 22>(enc, 640, 480, 640, 2, 125, 13)) return 1;
  if (run<144, 22>(enc, 640, 480, 640, 2, 125, 13)) return 1;
  if (run<144, 22>(enc, 640, 480, 640, 2, -3, -3)) return 1;
  if (run<144, 22>(enc, 444, 333, 448, 2, 381, 317)) return 1;
  if (run<144, 22>(enc, 533, 400, 544, 2, 509, -3)) return 1;
  return 0;
}
