// Microbenchmarks added in round 2 (sm_100a): packed fp32x2 issue rates (FFMA2 / FADD2) and tensor memory used as
// thread-private scratch (tcgen05.st / tcgen05.ld round trips).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

// 8 independent 64-bit chains per thread; OP 0/1 count 2 lane-flops per instruction and lane
template <int OP>
__global__ void __launch_bounds__(256) k2(uint32_t* out, uint32_t seed) {
  unsigned long long v[ILP];
  uint32_t a[ILP];
  float f[ILP];
  const uint32_t t = threadIdx.x + seed;
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    v[i] = ((unsigned long long)__float_as_uint(1.0f + t * 1e-3f + i) << 32) | __float_as_uint(0.5f + i);
    a[i] = t * (i + 3) + 1u;
    f[i] = 1.0f + i + t * 1e-3f;
  }
  const unsigned long long fb = ((unsigned long long)__float_as_uint(1.000001f) << 32) | __float_as_uint(0.999999f);
  const uint32_t c = blockIdx.x + 7u;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      const int j = (i + 1) % ILP;
      if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(fb), "l"(v[j]));
      else if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(v[j]));
      else if (OP == 2) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(fb), "l"(v[j])); a[i] = a[i] + a[j] + c; }
      else if (OP == 3) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(fb), "l"(v[j])); asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(a[j]), "r"(c)); }
      else if (OP == 4) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(v[j])); a[i] = __byte_perm(a[i], a[j], 0x4140); }
      else if (OP == 5) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(fb), "l"(v[j])); f[i] = fmaf(f[i], 1.000001f, f[j]); }
      else if (OP == 6) { asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(fb)); }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += a[i] + (uint32_t)v[i] + (uint32_t)(v[i] >> 32) + __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// dependent-chain latency of one instruction class (1 warp per SM, ILP 1)
template <int OP>
__global__ void klat(uint32_t* out) {
  unsigned long long v = threadIdx.x * 3ull + 1, w = 0x3f8000013f800001ull;
  float f = 1.0f + threadIdx.x;
  uint32_t a = threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 16
  for (int it = 0; it < 4096; ++it) {
    if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(v) : "l"(w));
    else if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
    else if (OP == 2) f = fmaf(f, 1.000001f, f);
    else if (OP == 3) asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a) : "r"(a), "r"(a));
  }
  const long long t1 = clock64();
  out[threadIdx.x] = (uint32_t)v + __float_as_uint(f) + a;
  if (threadIdx.x == 0) out[32] = (uint32_t)(t1 - t0);
}

// tensor memory as thread-private scratch (CTA of 4 warps owns all 512 columns; warp w -> lanes 32w..32w+31)
__global__ void __launch_bounds__(128) ktm(uint32_t* out, int iters, int mode) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t ta = tbase + ((uint32_t)(warp * 32) << 16);
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 17u + i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int blk = 0; blk < 4; ++blk) {  // 4 x 16 columns per iteration
      const uint32_t addr = ta + (uint32_t)(((it & 7) * 64 + blk * 16) & 511);
      if (mode != 1)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
      if (mode == 2) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (mode != 0) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] += 1u;
      }
    }
  }
  if (mode == 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) out[gridDim.x * blockDim.x + blockIdx.x] = (uint32_t)(t1 - t0);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <int OP>
void run2(const char* name, int flops_per_iter, int instr_per_iter, uint32_t* d_out, int sms, double mhz) {
  int grid = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k2<OP><<<grid, 256>>>(d_out, 1);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k2<OP><<<grid, 256>>>(d_out, rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double n = (double)grid * 256 * ITERS * ILP;
  printf("{\"op\": \"%s\", \"ms\": %.4f, \"lane_ops_per_clk_per_sm_at_%.0fMHz\": %.1f, \"warp_instr_per_clk_per_smsp\": %.3f}\n",
         name, best, mhz, n * flops_per_iter / (best * 1e-3) / (mhz * 1e6) / sms,
         n / 32 * instr_per_iter / (best * 1e-3) / (mhz * 1e6) / sms / 4);
}

template <int OP>
void runlat(const char* name, uint32_t* d) {
  klat<OP><<<1, 32>>>(d);
  cudaDeviceSynchronize();
  klat<OP><<<1, 32>>>(d);
  cudaDeviceSynchronize();
  uint32_t cyc = 0;
  cudaMemcpy(&cyc, d + 32, 4, cudaMemcpyDeviceToHost);
  printf("{\"op\": \"%s\", \"dependent_issue_cycles\": %.2f}\n", name, cyc / 4096.0);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double mhz = khz / 1000.0;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f}\n", p.name, sms, mhz);
  uint32_t* d; cudaMalloc(&d, (size_t)sms * 8 * 256 * 4 + 4096);
  run2<0>("FFMA2", 2, 1, d, sms, mhz);
  run2<1>("FADD2", 2, 1, d, sms, mhz);
  run2<6>("FMUL2", 2, 1, d, sms, mhz);
  run2<2>("FFMA2+IADD3", 3, 2, d, sms, mhz);
  run2<3>("FFMA2+IDP2A", 3, 2, d, sms, mhz);
  run2<4>("FADD2+PRMT", 3, 2, d, sms, mhz);
  run2<5>("FFMA2+FFMA", 3, 2, d, sms, mhz);
  runlat<0>("FFMA2 latency", d);
  runlat<1>("FADD2 latency", d);
  runlat<2>("FFMA latency", d);
  runlat<3>("IDP2A latency", d);
  const char* names[3] = {"tcgen05.st.x16 stream", "tcgen05.ld.x16+wait (dependent)", "tcgen05.st+wait+ld+wait round trip"};
  for (int mode = 0; mode < 3; ++mode) {
    const int iters = 2048;
    ktm<<<sms, 128>>>(d, iters, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"op\": \"%s\", \"error\": \"%s\"}\n", names[mode], cudaGetErrorString(e)); break; }
    ktm<<<sms, 128>>>(d, iters, mode);
    cudaDeviceSynchronize();
    uint32_t cyc = 0;
    cudaMemcpy(&cyc, d + (size_t)sms * 128, 4, cudaMemcpyDeviceToHost);
    printf("{\"op\": \"%s\", \"cycles_per_x16_op_per_warp\": %.1f, \"words_per_clk_per_sm\": %.1f}\n", names[mode],
           (double)cyc / (iters * 4.0), 128.0 * 16 * iters * 4 * (mode == 2 ? 2 : 1) / (double)cyc);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"status\": \"%s\"}\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
