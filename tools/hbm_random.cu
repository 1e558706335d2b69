// tools/hbm_random.cu -- what the B200's HBM delivers for RANDOM 32-byte sectors.
//
// The recombination table of kd_advance_kernel is 8 MB per lane (8 GB per 1024 lanes): every entry access is
// one 32-byte sector at a random address that L2 does not hold.  The copy bandwidth in MEASURED_PEAKS.json is
// for streams; this measures the same memory with the table's access pattern: independent random 32-byte
// reads, writes (full-sector stores) and a 1:1 mix over a buffer far larger than L2, with as much memory
// parallelism as the chip takes (8 accesses in flight per thread, 2048 threads per SM).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hbm_random tools/hbm_random.cu && /tmp/hbm_random [GB]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                        \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

struct alignas(32) Sector { unsigned long long w[4]; };

__device__ __forceinline__ Sector ld32(const Sector *p) {
  Sector s;
  asm volatile("ld.global.cg.v4.b64 {%0, %1, %2, %3}, [%4];"
               : "=l"(s.w[0]), "=l"(s.w[1]), "=l"(s.w[2]), "=l"(s.w[3]) : "l"(p));
  return s;
}
__device__ __forceinline__ void st32(Sector *p, unsigned long long v) {
  asm volatile("st.global.v4.b64 [%0], {%1, %1, %1, %1};" ::"l"(p), "l"(v) : "memory");
}

// mode 0: reads, 1: writes, 2: read one sector + write another per step
template <int MODE>
__global__ void __launch_bounds__(256) k_random(Sector *buf, uint64_t n_sectors, int iters, uint64_t seed,
                                                unsigned long long *sink) {
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  unsigned long long acc = 0;
  constexpr int U = 8;
  for (int it = 0; it < iters; ++it) {
    Sector s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t a = mix(seed + tid * 0x9E3779B97F4A7C15ull + (uint64_t)(it * U + u)) % n_sectors;
      if (MODE == 0 || MODE == 2) s[u] = ld32(buf + a);
      if (MODE == 1) st32(buf + a, a);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (MODE == 0 || MODE == 2) acc += s[u].w[0] + s[u].w[3];
      if (MODE == 2) {
        const uint64_t b = mix(~seed + tid * 0xD1B54A32D192ED03ull + (uint64_t)(it * U + u)) % n_sectors;
        st32(buf + b, acc);
      }
    }
  }
  if (acc == 0x1234567u) *sink = acc;
}

template <int MODE>
double run(Sector *buf, uint64_t n_sectors, int blocks, int iters, unsigned long long *sink) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_random<MODE><<<blocks, 256>>>(buf, n_sectors, iters / 4, 1, sink);  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_random<MODE><<<blocks, 256>>>(buf, n_sectors, iters, 7, sink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double accesses = (double)blocks * 256 * iters * 8 * (MODE == 2 ? 2 : 1);
  return accesses * 32.0 / (ms * 1e-3) / 1e9;
}

int main(int argc, char **argv) {
  const double gb = argc > 1 ? atof(argv[1]) : 8.0;
  const uint64_t n_sectors = (uint64_t)(gb * 1e9 / 32);
  Sector *buf;
  unsigned long long *sink;
  CK(cudaMalloc(&buf, n_sectors * 32));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(buf, 1, n_sectors * 32));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  // streaming reference on the same buffer: device-to-device copy of half onto the other half
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t half = n_sectors / 2 * 32;
    CK(cudaMemcpy((char *)buf + half, buf, half, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 4; ++i) CK(cudaMemcpyAsync((char *)buf + half, buf, half, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%s, %d SMs, buffer %.1f GB\n", p.name, p.multiProcessorCount, gb);
    printf("streaming copy (read + write bytes)      %8.0f GB/s\n", 4 * 2.0 * half / (ms * 1e-3) / 1e9);
  }
  for (int per_sm = 4; per_sm <= 8; per_sm += 4) {
    const int blocks = p.multiProcessorCount * per_sm;
    printf("%d x 256 threads per SM, 8 independent accesses per thread and step:\n", per_sm);
    printf("  random 32-byte reads                   %8.0f GB/s\n", run<0>(buf, n_sectors, blocks, 256, sink));
    printf("  random 32-byte full-sector writes      %8.0f GB/s\n", run<1>(buf, n_sectors, blocks, 256, sink));
    printf("  random read + random write, 1:1        %8.0f GB/s\n", run<2>(buf, n_sectors, blocks, 256, sink));
  }
  return 0;
}
