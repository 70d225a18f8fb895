// Probe: which 3-D TMA tile shapes over a {pair, parity, row} view of a pitched FP64 plane work on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma3d_probe tools/tma3d_probe.cu -lcuda ; ./tools/tma3d_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap map, double* out, int cx, int cy, int cz, int bytes, int n) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(s32(smem)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(s32(&bar)), "r"(cx), "r"(cy), "r"(cz) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(s32(smem)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(s32(&bar)), "r"(cx), "r"(cz) : "memory");
  }
  __syncthreads();
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
  } while (!ok);
  const double* t = reinterpret_cast<const double*>(smem);
  for (int q = threadIdx.x; q < n; q += blockDim.x) out[q] = t[q];
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int pitch = 8224, rows = 64, SH = 40;
  std::vector<double> h(size_t(pitch) * rows);
  for (size_t q = 0; q < h.size(); ++q) h[q] = double(q);
  double *d, *o;
  cudaMalloc(&d, h.size() * 8);
  cudaMalloc(&o, 128 * SH * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  CUtensorMap map;
  CUresult r;
  cuInit(0);
  const cuuint32_t es[3] = {1, 1, 1};
  int bytes = 128 * SH * 8;
  if (variant == 0) {  // 2-D reference
    const cuuint64_t gd[2] = {cuuint64_t(pitch), cuuint64_t(rows)}, gs[1] = {cuuint64_t(pitch) * 8};
    const cuuint32_t bx[2] = {128, cuuint32_t(SH)};
    r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (variant == 4) {  // 3-D with a trivial third dimension
    const cuuint64_t gd[3] = {cuuint64_t(pitch), cuuint64_t(rows), 1}, gs[2] = {cuuint64_t(pitch) * 8, cuuint64_t(pitch) * rows * 8};
    const cuuint32_t bx[3] = {128, cuuint32_t(SH), 1};
    r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (variant == 5 || variant == 6) {  // split view, other element types
    const cuuint64_t gd[3] = {cuuint64_t(pitch / 2), 2, cuuint64_t(rows)}, gs[2] = {cuuint64_t(pitch / 2) * 8, cuuint64_t(pitch) * 8};
    const cuuint32_t bx[3] = {64, 2, cuuint32_t(SH)};
    r = cuTensorMapEncodeTiled(&map, variant == 5 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_INT64, 3, d, gd, gs, bx, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (variant == 7) {  // split view as 32-bit words: {2 words * pairs, parity, row}
    const cuuint64_t gd[3] = {cuuint64_t(pitch), 2, cuuint64_t(rows)}, gs[2] = {cuuint64_t(pitch / 2) * 8, cuuint64_t(pitch) * 8};
    const cuuint32_t bx[3] = {128, 2, cuuint32_t(SH)};
    r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, gd, gs, bx, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t gd[3] = {cuuint64_t(pitch / 2), 2, cuuint64_t(rows)}, gs[2] = {cuuint64_t(pitch / 2) * 8, cuuint64_t(pitch) * 8};
    const cuuint32_t bx[3] = {64, cuuint32_t(variant == 2 ? 1 : 2), cuuint32_t(SH)};
    if (variant == 2) bytes /= 2;
    r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               variant == 3 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  printf("variant %d encode -> %d\n", variant, int(r));
  cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  const int n = bytes / 8;
  if (variant == 0) probe<2><<<1, 128, 44 * 1024>>>(map, o, 10, 0, 3, bytes, n);
  else if (variant == 4) probe<3><<<1, 128, 44 * 1024>>>(map, o, 10, 3, 0, bytes, n);
  else if (variant == 7) probe<3><<<1, 128, 44 * 1024>>>(map, o, 10, 0, 3, bytes, n);
  else probe<3><<<1, 128, 44 * 1024>>>(map, o, 5, 0, 3, bytes, n);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run -> %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<double> res(n);
    cudaMemcpy(res.data(), o, n * 8, cudaMemcpyDeviceToHost);
    printf("first row: %.0f %.0f ... [64] %.0f [65] %.0f ; second row [128] %.0f\n", res[0], res[1], res[64 % n], res[65 % n], res[128 % n]);
  }
  return 0;
}
