// Instruction-issue microbenchmark for sm_100a: measures lane-ops/clk/SM for the
// instruction classes the stereo kernels are built from. Used for (1) design decisions
// and (2) the measured ALU-issue roofline denominator (FFMA+IADD3 mix).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
  uint32_t a[ILP];
  float f[ILP];
  uint32_t t = threadIdx.x + seed;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = t * (i + 3) + 12345u; f[i] = (float)(t + i) * 1.0001f; }
  uint32_t b = t ^ 0x5a5a5a5au, c = t * 7u + 1u;
  float fb = 1.000001f, fc = 0.5f;
  __shared__ uint32_t sm[1024];
  sm[threadIdx.x] = t; sm[threadIdx.x + 256] = t; sm[threadIdx.x + 512] = t; sm[threadIdx.x + 768] = t;
  __syncthreads();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      const int j = (i + 1) % ILP;
      if (OP == 0) f[i] = fmaf(f[i], fb, f[j]);                      // FFMA
      else if (OP == 1) f[i] = f[i] + f[j];                           // FADD
      else if (OP == 2) a[i] = a[i] + a[j] + c;                       // IADD3
      else if (OP == 3) a[i] = a[i] * b + a[j];                       // IMAD
      else if (OP == 4) a[i] = __vabsdiffu4(a[i], a[j]);              // VABSDIFF4
      else if (OP == 5) a[i] = __dp4a(a[j], b, a[i]);                 // IDP.4A
      else if (OP == 6) { asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(a[j]), "r"(c)); } // IDP.2A
      else if (OP == 7) a[i] = __byte_perm(a[i], a[j], 0x4140);       // PRMT
      else if (OP == 8) { f[i] = (float)(a[i] & 0xff); a[i] = __float_as_uint(f[i]) + a[j]; }  // I2F.U8 + IADD
      else if (OP == 9) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1); // SHFL
      else if (OP == 10) a[i] = sm[(a[i] + threadIdx.x) & 1023];      // LDS.32 (dependent)
      else if (OP == 11) a[i] = min(a[i], a[j]) ;                     // IMNMX
      else if (OP == 12) f[i] = fminf(f[i], f[j]);                    // FMNMX
      else if (OP == 13) a[i] = (a[i] & a[j]) ^ b;                    // LOP3
      else if (OP == 14) { f[i] = fmaf(f[i], fb, f[j]); a[i] = a[i] + a[j] + c; } // FFMA + IADD3
      else if (OP == 15) { f[i] = fmaf(f[i], fb, f[j]); f[i] = f[i] + f[j]; a[i] = a[i] + a[j] + c; } // 2fp:1int
      else if (OP == 16) { f[i] = (float)(int)a[i]; a[i] = __float_as_uint(f[i]) + a[j]; }     // I2F.S32 + IADD
      else if (OP == 17) { a[i] = a[i] * b + a[j]; a[j] = a[j] + a[i] + c; }                     // IMAD + IADD3
      else if (OP == 18) { a[i] = __byte_perm(a[i], a[j], 0x4140); a[j] = a[j] + a[i] + c; }     // PRMT + IADD3
      else if (OP == 19) { asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(a[j]), "r"(c)); a[j] = a[j] + a[i] + b; } // IDP2A + IADD3
      else if (OP == 20) { asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(a[j]), "r"(c)); a[j] = a[j] * b + a[i]; } // IDP2A + IMAD
      else if (OP == 21) { f[i] = fmaf(f[i], fb, f[j]); a[i] = a[i] * b + a[j]; }                // FFMA + IMAD
      else if (OP == 22) { f[i] = fmaf(f[i], fb, f[j]); a[i] = __vabsdiffu4(a[i], a[j]); }       // FFMA + VABSDIFF4
      else if (OP == 23) { f[i] = fmaf(f[i], fb, f[j]); f[j] = f[j] + f[i]; a[i] = __byte_perm(a[i], a[j], 0x4140); a[j] = a[j] * b + a[i]; } // 2fp+alu+imad
      else if (OP == 24) { uint4 v = *reinterpret_cast<const uint4*>(&sm[((a[i] & 7) * 4 + i * 32) & 1020]); a[i] += v.x + v.y + v.z + v.w; } // LDS.128 bcast-ish
      else if (OP == 25) { uint4 v = *reinterpret_cast<const uint4*>(&sm[(threadIdx.x * 4 + (a[i] & 3) * 4 + i * 128) & 1020]); a[i] += v.x ^ v.y ^ v.z ^ v.w; } // LDS.128 distinct
      else if (OP == 26) { sm[(threadIdx.x + i * 32) & 1023] = a[i]; a[i] += c; }                 // STS.32
      else if (OP == 27) { f[i] = __uint_as_float(__byte_perm(a[i], 0x4b000000u, 0x7440)) - 8388608.0f; a[i] = __float_as_uint(f[i]) ^ a[j]; } // PRMT+FADD byte->float
      else if (OP == 28) { a[i] = __vimin3_s32(a[i], a[j], (int)c); }                              // VIMNMX3
      else if (OP == 30) { a[i] = __reduce_min_sync(0xffffffffu, (int)(a[i] + c)) + a[j]; }          // REDUX.MIN + IADD
      else if (OP == 31) { a[i] = __reduce_min_sync(0xffffffffu, (int)(a[i] + c)) + threadIdx.x; }     // REDUX.MIN (independent chains)
      else if (OP == 32) { int bb = (int)a[i]; a[i] = (uint32_t)(bb ^ ((bb >> 31) & 0x7fffffff)) + a[j]; } // sortable + IADD
      else if (OP == 33) { a[i] = a[i] ^ a[j]; }                                                        // LOP3 (xor)
      else if (OP == 34) { int bb = (int)a[i]; int k = bb ^ ((bb >> 31) & 0x7fffffff); a[i] = (uint32_t)((k & ~31) | (int)(threadIdx.x & 31)) + a[j]; } // key prep + IADD
      else if (OP == 35) { int bb = (int)a[i]; int sg = (int)__byte_perm((uint32_t)bb, 0u, 0xBBBB); int k = bb ^ (sg & 0x7fffffff); a[i] = (uint32_t)((k & ~31) | (int)(threadIdx.x & 31)) + a[j]; } // key prep, PRMT sign
      else if (OP == 36) { a[i] = (a[i] >> 3) + a[j]; }                                                 // SHF + IADD
      else if (OP == 37) { a[i] = (a[i] > a[j]) ? c : b; }                                              // ISETP + SEL
      else if (OP == 29) { f[i] = (float)((a[i] >> 8) & 0xff); a[i] = __float_as_uint(f[i]) + a[j]; } // I2F.U8.B1
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += a[i] + __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
void run(const char* name, int ops_per_iter, uint32_t* d_out, int sms, double mhz) {
  int grid = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<grid, 256>>>(d_out, 1);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<OP><<<grid, 256>>>(d_out, rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double laneops = (double)grid * 256 * ITERS * ILP * ops_per_iter;
  double tops = laneops / (best * 1e-3) / 1e12;
  printf("{\"op\": \"%s\", \"ms\": %.4f, \"Tlaneops_s\": %.3f, \"per_clk_per_sm_at_%.0fMHz\": %.1f}\n",
         name, best, tops, mhz, laneops / (best * 1e-3) / (mhz * 1e6) / sms);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double mhz = khz / 1000.0;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f}\n", p.name, sms, mhz);
  uint32_t* d; cudaMalloc(&d, (size_t)sms * 8 * 256 * 4);
  run<0>("FFMA", 1, d, sms, mhz);
  run<1>("FADD", 1, d, sms, mhz);
  run<2>("IADD3", 1, d, sms, mhz);
  run<3>("IMAD", 1, d, sms, mhz);
  run<4>("VABSDIFF4", 1, d, sms, mhz);
  run<5>("IDP4A", 1, d, sms, mhz);
  run<6>("IDP2A", 1, d, sms, mhz);
  run<7>("PRMT", 1, d, sms, mhz);
  run<8>("I2F.U8+IADD", 2, d, sms, mhz);
  run<9>("SHFL", 1, d, sms, mhz);
  run<10>("LDS32dep", 1, d, sms, mhz);
  run<11>("IMNMX", 1, d, sms, mhz);
  run<12>("FMNMX", 1, d, sms, mhz);
  run<13>("LOP3", 1, d, sms, mhz);
  run<14>("FFMA+IADD3", 2, d, sms, mhz);
  run<15>("FFMA+FADD+IADD3", 3, d, sms, mhz);
  run<16>("I2F.S32+IADD", 2, d, sms, mhz);
  run<17>("IMAD+IADD3", 2, d, sms, mhz);
  run<18>("PRMT+IADD3", 2, d, sms, mhz);
  run<19>("IDP2A+IADD3", 2, d, sms, mhz);
  run<20>("IDP2A+IMAD", 2, d, sms, mhz);
  run<21>("FFMA+IMAD", 2, d, sms, mhz);
  run<22>("FFMA+VABSDIFF4", 2, d, sms, mhz);
  run<23>("FFMA+FADD+PRMT+IMAD", 4, d, sms, mhz);
  run<24>("LDS128bcast+3IADD", 1, d, sms, mhz);
  run<25>("LDS128distinct+3LOP", 1, d, sms, mhz);
  run<26>("STS32+IADD", 1, d, sms, mhz);
  run<27>("PRMT+FADD(b2f)+LOP", 3, d, sms, mhz);
  run<28>("VIMNMX3", 1, d, sms, mhz);
  run<29>("I2F.U8.B1+IADD(+shift?)", 2, d, sms, mhz);
  run<30>("REDUX.MIN+IADD", 1, d, sms, mhz);
  run<31>("REDUX.MIN+IADD(lane)", 1, d, sms, mhz);
  run<32>("sortable(SHF+LOP3)+IADD", 1, d, sms, mhz);
  run<33>("LOP3xor", 1, d, sms, mhz);
  run<34>("keyprep(SHF+LOP3+LOP3)+IADD", 1, d, sms, mhz);
  run<35>("keyprep(PRMT+LOP3+LOP3)+IADD", 1, d, sms, mhz);
  run<36>("SHF+IADD", 1, d, sms, mhz);
  run<37>("ISETP+SEL", 1, d, sms, mhz);
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"status\": \"%s\"}\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
