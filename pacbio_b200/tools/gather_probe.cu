// Probe: how many bytes does ONE random small load cost on this GPU, by load flavour?
// (ncu on the library's random-gather microkernel showed 4 sectors L1->L2 and 117 B of DRAM traffic per
// random 16-byte __ldg.)  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo gather_probe.cu -o gather_probe
// Run under: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum ./gather_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template<int kFlavour, typename T> __device__ __forceinline__ T load(const T* p);
#define DEF(F, T, ASM, C, ...) template<> __device__ __forceinline__ T load<F, T>(const T* p) { T v; asm volatile(ASM : __VA_ARGS__ : "l"(p)); return v; }
DEF(0, uint4, "ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];", , "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w))
DEF(1, uint4, "ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];", , "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w))
DEF(2, uint4, "ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];", , "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w))
DEF(3, uint4, "ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];", , "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w))
DEF(4, uint4, "ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];", , "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w))
DEF(5, uint4, "ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];", , "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w))
DEF(6, uint4, "ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];", , "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w))

template<int kFlavour, int kBytes>
__global__ void __launch_bounds__(256) gather(const uint4* __restrict__ table, uint64_t nelem, uint32_t rounds, uint32_t* sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for(uint32_t r = 0; r < rounds; ++r) {
    uint4 v[8];
#pragma unroll
    for(int j = 0; j < 8; ++j) {
      const uint64_t h = ((uint64_t)mix32((tid * 8u + j) ^ mix32(r + 0x9e3779b9u)) << 32) | mix32((r + 1) * 0x9e3779b9u ^ (tid + j * 0x85ebca6bu));
      const uint4* p = table + (uint64_t)(((unsigned __int128)h * nelem) >> 64);
      if(kBytes == 16) v[j] = load<kFlavour, uint4>(p);
      else {   // 4-byte and 8-byte loads: the nc flavour only
        if(kBytes == 8) { const uint2 w = __ldg(reinterpret_cast<const uint2*>(p)); v[j] = make_uint4(w.x, w.y, 0, 0); }
        else            { v[j] = make_uint4(__ldg(reinterpret_cast<const uint32_t*>(p)), 0, 0, 0); }
      }
    }
#pragma unroll
    for(int j = 0; j < 8; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if(acc == 0x12345678u) sink[0] = acc;
}

template<int F, int B> void run(const char* name, const uint4* t, uint64_t nelem, uint32_t* sink) {
  const unsigned grid = 148 * 8; const uint32_t rounds = 64;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  gather<F, B><<<grid, 256>>>(t, nelem, rounds, sink);
  cudaEventRecord(a);
  gather<F, B><<<grid, 256>>>(t, nelem, rounds, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double loads = (double)grid * 256 * 8 * rounds;
  printf("%-28s %8.3f ms  %7.2f G loads/s  err=%s\n", name, ms, loads / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const uint64_t bytes = argc > 1 ? strtoull(argv[1], 0, 0) : (1ull << 30);
  if(argc > 2) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[2])); printf("set L2 fetch granularity %s: %s\n", argv[2], cudaGetErrorString(e)); }
  size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit: %zu; table %llu bytes\n", g, (unsigned long long)bytes);
  uint4* t; uint32_t* sink; cudaMalloc(&t, bytes); cudaMalloc(&sink, 64); cudaMemset(t, 0x5a, bytes);
  const uint64_t n = bytes / 16;
  run<0, 16>("nc v4 (=__ldg)", t, n, sink);
  run<1, 16>("cg v4", t, n, sink);
  run<2, 16>("ca v4", t, n, sink);
  run<3, 16>("cv v4", t, n, sink);
  run<4, 16>("nc L1::no_allocate v4", t, n, sink);
  run<5, 16>("nc L2::64B v4", t, n, sink);
  run<6, 16>("cs v4", t, n, sink);
  run<0, 8>("nc 8 bytes", t, n, sink);
  run<0, 4>("nc 4 bytes", t, n, sink);
  return 0;
}
