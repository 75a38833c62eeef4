// HBM streaming ceilings for the access mixes of the quantisation front-end (268 MB of 16-bit input, configs[1] size):
//   read   : 2 B/elt read, nothing written                 (channel statistics)
//   r2w1   : 2 B/elt read, 1 B/elt written                (INT8 / FP8 quantisers)
//   copy   : 2 B/elt read, 2 B/elt written                 (what MEASURED_PEAKS.json calls hbm_gbs)
// Each in two shapes: `flat` (grid-stride, 16 B per thread and iteration, U loads in flight) and `tile` (one CTA per 32 KB
// tile, all loads issued, __syncthreads, all stores — the phase structure of the quantisers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/rw_mix tools/microbench/rw_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE, int U>   // MODE 0 read, 1 r2w1, 2 copy
__global__ void __launch_bounds__(256) flat_kernel(const uint4* __restrict__ in, void* __restrict__ out, size_t n16, uint32_t* sink) {
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldcs(in + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (MODE == 0) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
      if (MODE == 1) reinterpret_cast<uint2*>(out)[i + u * stride] = make_uint2(v[u].x ^ v[u].y, v[u].z ^ v[u].w);
      if (MODE == 2) reinterpret_cast<uint4*>(out)[i + u * stride] = v[u];
    }
  }
  if (MODE == 0 && acc == 0x12345678u) *sink = acc;
}

template <int MODE>
__global__ void __launch_bounds__(256, 3) tile_kernel(const uint4* __restrict__ in, void* __restrict__ out, uint32_t* sink) {
  // 128 rows x 256 B: thread (tr, tc) takes 16 B of rows tr, tr + 16, ... (the quantisers' mapping)
  const size_t base = size_t(blockIdx.x) * 2048;   // uint4 units per tile
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  uint4 v[8];
#pragma unroll
  for (int ps = 0; ps < 8; ++ps) v[ps] = in[base + (ps * 16 + tr) * 16 + tc];
  __shared__ uint32_t s[256];
  s[threadIdx.x] = v[0].x ^ v[7].w;
  __syncthreads();
  const uint32_t k = s[(threadIdx.x + 17) & 255];
  __syncthreads();
  uint32_t acc = 0;
#pragma unroll
  for (int ps = 0; ps < 8; ++ps) {
    const size_t o = base + (ps * 16 + tr) * 16 + tc;
    if (MODE == 0) acc ^= v[ps].x ^ v[ps].y ^ v[ps].z ^ v[ps].w ^ k;
    if (MODE == 1) reinterpret_cast<uint2*>(out)[o] = make_uint2(v[ps].x ^ v[ps].y ^ k, v[ps].z ^ v[ps].w);
    if (MODE == 2) reinterpret_cast<uint4*>(out)[o] = make_uint4(v[ps].x ^ k, v[ps].y, v[ps].z, v[ps].w);
  }
  if (MODE == 0 && acc == 0x12345678u) *sink = acc;
}

template <typename F>
static float time_ms(F f) {
  for (int i = 0; i < 3; ++i) f();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  const int n = 20;
  for (int i = 0; i < n; ++i) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / n;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const size_t elts = size_t(4) * 32 * 8192 * 128, bytes = elts * 2, n16 = bytes / 16;
  // three independent input sets so that nothing is re-read from the 126 MB L2
  uint4* in[3]; void* out; uint32_t* sink;
  for (auto& q : in) { cudaMalloc(&q, bytes); cudaMemset(q, 1, bytes); }
  cudaMalloc(&out, bytes); cudaMalloc(&sink, 4);
  printf("%s, %d SMs; %zu MB of 16-bit input per pass\n", p.name, p.multiProcessorCount, bytes >> 20);
  const char* names[3] = {"read (2 B/elt in)", "r2w1 (2 in + 1 out)", "copy (2 in + 2 out)"};
  const double traffic[3] = {2.0, 3.0, 4.0};
  int rot = 0;
  auto report = [&](const char* shape, int mode, float ms) {
    printf("  %-6s %-22s %7.1f us  %5.2f TB/s\n", shape, names[mode], ms * 1e3, traffic[mode] * elts / (ms * 1e-3) / 1e12);
  };
  const int G = p.multiProcessorCount * 8;
#define FLAT(M, U) report("flat" #U, M, time_ms([&] { flat_kernel<M, U><<<G, 256>>>(in[rot++ % 3], out, n16, sink); }))
  FLAT(0, 1); FLAT(0, 4); FLAT(0, 8);
  FLAT(1, 1); FLAT(1, 4); FLAT(1, 8);
  FLAT(2, 1); FLAT(2, 4); FLAT(2, 8);
#define TILE(M) report("tile", M, time_ms([&] { tile_kernel<M><<<unsigned(n16 / 2048), 256>>>(in[rot++ % 3], out, sink); }))
  TILE(0); TILE(1); TILE(2);
  cudaDeviceSynchronize();
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
