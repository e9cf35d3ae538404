// Developer micro-benchmark: reciprocal throughput (cycles per warp instruction per SM sub-partition) of the
// instruction kinds the stepper leans on, alone and in the mixes its phases produce.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define CH 8   // independent chains per thread

template <int KIND>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, long long* cyc, uint32_t seed, float fseed) {
  uint32_t a[CH];
  float f[CH];
  unsigned long long d[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { a[i] = seed + threadIdx.x * 31 + i; f[i] = fseed + i * 0.01f + threadIdx.x * 1e-4f; d[i] = (unsigned long long)a[i] << 20 | i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (KIND == 0) {   // IMAD.WIDE.U32 imm
        unsigned long long p = (unsigned long long)a[i] * 0xD2511F53u; a[i] = (uint32_t)(p >> 32) ^ (uint32_t)p;   // + LOP3
      } else if (KIND == 1) {   // LOP3 only (2 per chain)
        a[i] = (a[i] ^ seed ^ (a[i] >> 3)); a[i] = (a[i] & 0x7fffffffu) | (seed << 1);
      } else if (KIND == 2) {   // IMAD lo
        a[i] = a[i] * 0xD2511F53u + seed;
      } else if (KIND == 3) {   // FFMA reg
        f[i] = fmaf(f[i], fseed, f[(i + 1) % CH]);
      } else if (KIND == 4) {   // FFMA2
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(d[(i + 1) % CH]), "l"(d[(i + 2) % CH]));
      } else if (KIND == 5) {   // MUFU.EX2
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      } else if (KIND == 6) {   // MUFU.RCP
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      } else if (KIND == 7) {   // MUFU.SIN
        asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      } else if (KIND == 8) {   // MUFU.LG2
        asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      } else if (KIND == 9) {   // MUFU.SQRT
        asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      } else if (KIND == 10) {  // IMAD.WIDE + MUFU alternating, independent
        unsigned long long p = (unsigned long long)a[i] * 0xD2511F53u; a[i] = (uint32_t)(p >> 32) ^ (uint32_t)p;
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      } else if (KIND == 11) {  // IMAD.WIDE alone (xor folded into next multiply through addend form)
        unsigned long long p = (unsigned long long)a[i] * 0xD2511F53u + d[i]; d[i] = p; a[i] = (uint32_t)(p >> 32);
      } else if (KIND == 12) {  // IMAD.HI only
        a[i] = __umulhi(a[i], 0xD2511F53u) + seed;
      } else if (KIND == 13) {  // FMUL
        f[i] = f[i] * fseed;
      } else if (KIND == 14) {  // FFMA imm
        f[i] = fmaf(f[i], 1.0001f, 0.5f);
      } else if (KIND == 15) {  // IADD3
        a[i] = a[i] + seed + a[(i + 1) % CH];
      }
    }
  }
  long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) r ^= a[i] ^ __float_as_uint(f[i]) ^ (uint32_t)d[i] ^ (uint32_t)(d[i] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int ops_per_chain_iter, int threads) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  k<KIND><<<148, threads>>>(out, cyc, 12345u, 0.7f);
  k<KIND><<<148, threads>>>(out, cyc, 12345u, 0.7f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double warp_instr_per_smsp = (double)ITERS * CH * ops_per_chain_iter * (threads / 32 / 4.0);
  printf("%-34s threads %4d: %7.3f cycles per warp-instr per SMSP (%d instr per chain-iter)\n", name, threads, c / warp_instr_per_smsp, ops_per_chain_iter);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 512, 1024}) {
    run<0>("IMAD.WIDE.U32 + LOP3", 2, threads);
    run<11>("IMAD.WIDE.U32 (with addend)", 1, threads);
    run<12>("IMAD.HI.U32 + IADD", 2, threads);
    run<2>("IMAD (lo)", 1, threads);
    run<1>("LOP3 x2", 2, threads);
    run<15>("IADD3", 1, threads);
    run<3>("FFMA reg", 1, threads);
    run<14>("FFMA imm", 1, threads);
    run<13>("FMUL", 1, threads);
    run<4>("FFMA2", 1, threads);
    run<5>("MUFU.EX2", 1, threads);
    run<6>("MUFU.RCP", 1, threads);
    run<7>("MUFU.SIN", 1, threads);
    run<8>("MUFU.LG2", 1, threads);
    run<9>("MUFU.SQRT", 1, threads);
    run<10>("IMAD.WIDE + LOP3 + MUFU.EX2", 3, threads);
  }
  return 0;
}
