// pipes.cu -- instruction-throughput microbenchmark for sm_100a (tuning aid, not product code).
// Each test runs ITERS x 16 independent instructions per thread, 8 warps per SM sub-partition;
// prints warp-instructions per clock per sub-partition (SMSP).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048

template <int T>
__device__ __forceinline__ void body(uint32_t (&r)[16], uint32_t a, uint32_t b, uint32_t sbase) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (T == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
    if (T == 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
    if (T == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
    if (T == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a));
    if (T == 4) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
    if (T == 5) asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(r[i]) : "r"(a));          // IMAD.IADD?
    if (T == 6) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a));
    if (T == 7) { unsigned long long t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(r[i]), "r"(a)); r[i] = (uint32_t)t ^ (uint32_t)(t >> 32); }
    if (T == 8) asm volatile("shr.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a));
    if (T == 9) asm volatile("max.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(a));
    if (T == 10) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r[i]) : "r"(sbase + ((r[i] & 0x3fc)) ) : "memory");
    if (T == 11) asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(sbase + 4 * (threadIdx.x & 31) + 128 * i), "r"(r[i]) : "memory");
    if (T == 12) asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(r[i]) : "r"(sbase + 4 * (threadIdx.x & 31) + 128 * i), "r"(a) : "memory");
    if (T == 13) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(sbase + 4 * (threadIdx.x & 31) + 128 * i), "r"(r[i]) : "memory");
    if (T == 14) { if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b)); else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b)); }
    if (T == 15) { if (i % 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b)); else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b)); }
    if (T == 16) asm volatile("shl.b32 %0, %0, 3;" : "+r"(r[i]));
    if (T == 17) { asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %0, %2, p; }" : "+r"(r[i]) : "r"(a), "r"(b)); }
    if (T == 18) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r[i]), "=r"(r[(i + 1) & 15]) : "r"(sbase + ((r[i] & 0x3f8))) : "memory");
    if (T == 19) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+r"(r[i]));
    if (T == 20) asm volatile("st.shared.u32 [%0], %1;" :: "r"(sbase + 4 * (threadIdx.x & 31) + 128 * i), "r"(r[i]) : "memory");
    if (T == 21) { if (i & 1) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r[i]) : "r"(sbase + ((r[i] & 0x3fc)) ) : "memory"); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b)); }
    if (T == 22) asm volatile("lop3.b32 %0, %0, 0xff00ff, %1, 0xea;" : "+r"(r[i]) : "r"(b));   // immediate form
    if (T == 23) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
    if (T == 24) asm volatile("popc.b32 %0, %0;" : "+r"(r[i]));
    if (T == 25) asm volatile("bfe.u32 %0, %0, 8, 8;" : "+r"(r[i]));
    if (T == 26) asm volatile("add.u32 %0, %0, 128;" : "+r"(r[i]));
    if (T == 27) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(r[(i + 1) & 15]), "r"(r[(i + 2) & 15]));
  }
}

template <int T>
__global__ void __launch_bounds__(1024) k(uint32_t a, uint32_t b, uint32_t* out, long long* cyc) {
  __shared__ uint32_t sm[2048 + 1024];
  for (int i = threadIdx.x; i < 3072; i += blockDim.x) sm[i] = i * 4 & 0x3fc;
  __syncthreads();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm) + ((T >= 11 && T <= 13) || T == 20 ? 128u * 16 * (threadIdx.x >> 5) % 8192 : 0);
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 17 + i * a;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) body<T>(r, a, b, sbase);
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int T>
void run(const char* name, uint32_t* out, long long* cyc) {
  k<T><<<148, 1024>>>(3, 5, out, cyc);
  k<T><<<148, 1024>>>(3, 5, out, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += (double)h[i];
  avg /= 148;
  const double winst = 32.0 * ITERS * 16;  // warp instructions per SM (32 warps)
  printf("%-28s %.3f warp-inst/clk/SMSP  (err %s)\n", name, winst / avg / 4.0, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>("LOP3 (3-reg)", out, cyc);
  run<22>("LOP3 (imm)", out, cyc);
  run<1>("SHF.L.W (3-reg)", out, cyc);
  run<27>("SHF.R.W (3 distinct regs)", out, cyc);
  run<8>("SHR var", out, cyc);
  run<16>("SHL imm", out, cyc);
  run<2>("IMAD (3-reg)", out, cyc);
  run<5>("IMAD x*1+c", out, cyc);
  run<3>("ADD", out, cyc);
  run<26>("ADD imm", out, cyc);
  run<4>("PRMT", out, cyc);
  run<6>("MUL.HI", out, cyc);
  run<7>("MUL.WIDE", out, cyc);
  run<9>("MAX", out, cyc);
  run<17>("SETP+SELP", out, cyc);
  run<24>("POPC", out, cyc);
  run<25>("BFE", out, cyc);
  run<23>("VABSDIFF4", out, cyc);
  run<14>("LOP3 + IMAD 1:1", out, cyc);
  run<15>("LOP3 + IMAD 2:1", out, cyc);
  run<10>("LDS.32 random 1KB", out, cyc);
  run<18>("LDS.64 random 1KB", out, cyc);
  run<21>("LDS.32 + LOP3 1:1", out, cyc);
  run<20>("STS.32 conflict-free", out, cyc);
  run<11>("RED.OR shared conflict-free", out, cyc);
  run<13>("RED.ADD shared conflict-free", out, cyc);
  run<12>("ATOM.EXCH shared", out, cyc);
  run<19>("SHFL.UP", out, cyc);
  return 0;
}
