// Scratch micro-benchmark: what HBM bandwidth does a persistent 148 x 512-thread launch (the cooperative kernel's shape,
// 16 warps per SM) reach as a function of the loads each thread keeps in flight, and what does a TMA bulk-copy ring reach
// with the same occupancy?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench stream_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ double2 ldg2(const double2* p) {
  double2 v;
  asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg2(double2* p, double2 v) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// NA arrays of n2 double2 each; every thread keeps NA * UN 16-byte loads in flight per round
template <int NA, int UN, int TPB, int CPS>
__global__ void __launch_bounds__(TPB, CPS) rd(const double2* base, size_t n2, double* out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g + (UN - 1) * stride < n2; g += UN * stride) {
    double2 v[NA][UN];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int u = 0; u < UN; ++u) v[a][u] = ldg2(base + (size_t)a * n2 + g + u * stride);
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int u = 0; u < UN; ++u) acc += v[a][u].x * v[a][u].y;
  }
  if (acc == 1.2345) out[0] = acc;
}


// L2 re-read test: every thread loads its element and the same element of the rows above / below (row = 128 double2 = 2 KB):
// DRAM traffic 1x, L2 -> L1 traffic NB x (the neighbour rows are fetched by other CTAs at about the same time => L2 hits)
template <int NA, int NB, int OFF>
__global__ void __launch_bounds__(512, 1) rd_nbr(const double2* base, size_t n2, double* out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x + (size_t)OFF * NB; g + (size_t)OFF * NB < n2; g += stride) {
    double2 v[NA][NB];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int b = 0; b < NB; ++b) v[a][b] = ldg2(base + (size_t)a * n2 + g + (size_t)(b - NB / 2) * OFF);
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc += v[a][b].x * v[a][b].y;
  }
  if (acc == 1.2345) out[0] = acc;
}

// D-like: read NR arrays, write NW arrays, FL dependent DFMA per element pair in between
template <int NR, int NW, int UN, int FL, int TPB, int CPS>
__global__ void __launch_bounds__(TPB, CPS) rw(const double2* base, double2* obase, size_t n2) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g + (UN - 1) * stride < n2; g += UN * stride) {
    double2 v[NR][UN];
#pragma unroll
    for (int a = 0; a < NR; ++a)
#pragma unroll
      for (int u = 0; u < UN; ++u) v[a][u] = ldg2(base + (size_t)a * n2 + g + u * stride);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      double2 s = make_double2(0.0, 0.0);
#pragma unroll
      for (int a = 0; a < NR; ++a) { s.x += v[a][u].x; s.y += v[a][u].y; }
#pragma unroll
      for (int f = 0; f < FL; ++f) { s.x = s.x * 1.0000001 + 0.5; s.y = s.y * 0.9999999 + 0.25; }
#pragma unroll
      for (int a = 0; a < NW; ++a) stg2(obase + (size_t)a * n2 + g + u * stride, make_double2(s.x + a, s.y - a));
    }
  }
}

// ---- TMA bulk ring: warp 0 lane 0 streams CHUNK-byte pieces of NA arrays into an S-stage shared-memory ring, all other
// warps consume (sum) them.  full[s]: tx-count barrier; empty[s]: one arrival per consumer warp.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(512, 1) tma_ring(const char* base, size_t bytes, double* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + STAGES;
  unsigned char* ring = sm + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, ncons = (blockDim.x >> 5) - 1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncons); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t nchunks = bytes / CHUNK;
  double acc = 0.0;
  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], CHUNK);
        bulk_g2s(ring + (size_t)s * CHUNK, base + c * CHUNK, CHUNK, &full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    int s = 0; uint32_t ph = 0;
    const int ct = (warp - 1) * 32 + lane, nct = ncons * 32;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      mbar_wait(&full[s], ph);
      const double2* p = reinterpret_cast<const double2*>(ring + (size_t)s * CHUNK);
      for (int i = ct; i < CHUNK / 16; i += nct) { const double2 v = p[i]; acc += v.x * v.y; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }
  if (acc == 1.2345) out[0] = acc;
}

template <typename F> static float time_it(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t n = (size_t)64 * 256 * 256;       // doubles per array (cfg3 tsp65)
  const size_t n2 = n / 2;
  const int NARR = 14;
  double2 *in, *outb; double* o;
  CK(cudaMalloc(&in, NARR * n * 8)); CK(cudaMalloc(&outb, 6 * n * 8)); CK(cudaMalloc(&o, 64));
  CK(cudaMemset(in, 0, NARR * n * 8));
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d, array %.1f MB\n", sms, n * 8 / 1e6);
#define RD(NA, UN, TPB, CPS) { float ms = time_it([&] { rd<NA, UN, TPB, CPS><<<sms * CPS, TPB>>>(in, n2, o); }); \
    printf("read  %2d arrays, %2d loads in flight/thread, %4d thr x %d CTA/SM: %7.1f us  %6.0f GB/s\n", NA, NA * UN, TPB, CPS, ms * 1e3, NA * n * 8 / ms / 1e6); }
  RD(6, 1, 512, 1) RD(6, 2, 512, 1) RD(6, 4, 512, 1) RD(6, 8, 512, 1)
  RD(12, 1, 512, 1) RD(12, 2, 512, 1)
  RD(6, 1, 512, 2) RD(6, 2, 512, 2) RD(6, 4, 512, 2)
  RD(6, 1, 1024, 2) RD(6, 2, 1024, 2)
  RD(1, 8, 512, 1) RD(1, 16, 512, 1) RD(1, 8, 1024, 2)
#define RN(NA, NB, OFF) { float ms = time_it([&] { rd_nbr<NA, NB, OFF><<<sms, 512>>>(in, n2, o); }); \
    printf("nbr   %2d arrays x %d copies, offset %6d: %7.1f us  unique %6.0f GB/s  loads %6.0f GB/s\n", NA, NB, OFF, ms * 1e3, NA * n * 8 / ms / 1e6, NA * NB * n * 8 / ms / 1e6); }
  RN(6, 1, 128) RN(6, 3, 128) RN(6, 2, 18944) RN(6, 3, 18944) RN(4, 3, 18944) RN(2, 5, 18944) RN(2, 3, 18944) RN(6, 3, 3200) RN(2, 5, 3200)
#define RW(NR, NW, UN, FL, TPB, CPS) { float ms = time_it([&] { rw<NR, NW, UN, FL, TPB, CPS><<<sms * CPS, TPB>>>(in, outb, n2); }); \
    printf("r/w   %2d in %d out, x%d, %3d dfma: %4d thr x %d CTA/SM: %7.1f us  %6.0f GB/s\n", NR, NW, UN, FL, TPB, CPS, ms * 1e3, (NR + NW) * n * 8 / ms / 1e6); }
  { cudaFuncSetAttribute(rw<7, 5, 1, 100, 512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    float ms = time_it([&] { rw<7, 5, 1, 100, 512, 1><<<sms, 512, 160 * 1024>>>(in, outb, n2); });
    printf("r/w 7 in 5 out x1 100 dfma with 160 KB dynamic smem (L1 = 92 KB): %7.1f us  %6.0f GB/s\n", ms * 1e3, 12 * n * 8 / ms / 1e6);
    cudaFuncSetAttribute(rd<6, 2, 512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    ms = time_it([&] { rd<6, 2, 512, 1><<<sms, 512, 160 * 1024>>>(in, n2, o); });
    printf("read 6 arrays x2 with 160 KB dynamic smem: %7.1f us  %6.0f GB/s\n", ms * 1e3, 6 * n * 8 / ms / 1e6);
    cudaFuncSetAttribute(rd_nbr<6, 3, 18944>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    ms = time_it([&] { rd_nbr<6, 3, 18944><<<sms, 512, 160 * 1024>>>(in, n2, o); });
    printf("nbr 6 x 3 with 160 KB dynamic smem: %7.1f us  %6.0f GB/s\n", ms * 1e3, 6 * n * 8 / ms / 1e6); }
  RW(7, 5, 1, 0, 512, 1) RW(7, 5, 2, 0, 512, 1) RW(7, 5, 1, 100, 512, 1) RW(7, 5, 2, 100, 512, 1) RW(7, 5, 1, 0, 512, 2) RW(7, 5, 2, 0, 512, 2)
  RW(7, 5, 1, 100, 512, 2) RW(7, 5, 1, 0, 1024, 2) RW(7, 5, 1, 100, 1024, 2) RW(2, 2, 4, 0, 512, 1) RW(2, 2, 8, 0, 512, 1)
#define TR(ST, CH) { cudaFuncSetAttribute(tma_ring<ST, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + ST * CH); \
    float ms = time_it([&] { tma_ring<ST, CH><<<sms, 512, 1024 + ST * CH>>>((const char*)in, 6 * n * 8, o); }); \
    CK(cudaGetLastError()); \
    printf("tma ring %d stages x %5d B (%3d KB in flight/SM): %7.1f us  %6.0f GB/s\n", ST, CH, ST * CH / 1024, ms * 1e3, 6 * n * 8 / ms / 1e6); }
  TR(2, 16384) TR(4, 16384) TR(8, 16384) TR(4, 32768) TR(6, 32768) TR(8, 8192) TR(16, 8192) TR(16, 4096) TR(32, 2048)
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
