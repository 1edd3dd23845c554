// Pipe-throughput micro-benchmarks for B200 (sm_100a): how many cycles does an SMSP need per
// warp-instruction of MUFU.EX2 / FFMA / FMNMX, alone and mixed?  One CTA per SM, W warps per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  float s = 0.f, m = -1e30f;
  unsigned long long acc2 = 0ull;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = ex2(a[i]);                                  // MUFU only
      if (MODE == 1) a[i] = fmaf(a[i], 1.0001f, 0.5f);                  // FFMA only
      if (MODE == 2) m = fmaxf(m, a[i]), a[i] = m * 0.5f;               // dependent; ignore
      if (MODE == 3) { float e = ex2(fmaf(a[i], 0.999f, -0.1f)); s += e; a[i] = e; }  // softmax inner loop
      if (MODE == 4) { asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s)); }  // FMNMX only
      if (MODE == 6) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a[i]), "f"(s)); a[i] = __uint_as_float(u | 0x3f000000u); }
      if (MODE == 7) { float e = ex2(fmaf(a[i], 0.999f, -0.1f)); s += e; if (i & 1) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(e), "f"(a[i-1])); m += __uint_as_float(u); } a[i] = e; }
      if (MODE == 8) { unsigned u = __float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u); }
      if (MODE == 9) { unsigned u = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u); }
      if (MODE == 10) { // pair: 2 ffma, pack f16x2, ex2.f16x2, unpack+2 fadd
        if (i & 1) { float x0 = fmaf(a[i-1], 0.999f, -0.1f), x1 = fmaf(a[i], 0.999f, -0.1f); unsigned u;
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(x1), "f"(x0));
          asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u));
          float p0, p1; asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=f"(p0), "=f"(p1) : "r"(u));
          s += p0; m += p1; a[i-1] = p0; a[i] = p1; } }
      if (MODE == 11) { float e = ex2(fmaf(a[i], 0.999f, -0.1f)); s += e; if (i & 1) { unsigned u0 = __float_as_uint(a[i-1]) + 0x8000u, u1 = __float_as_uint(e) + 0x8000u; unsigned u = __byte_perm(u0, u1, 0x7632); m += __uint_as_float(u); } a[i] = e; }
      if (MODE == 12) { if (i & 1) { // f32x2 FFMA + 2 EX2 + f32x2 FADD + integer pack
          unsigned long long xx, cc = 0x3f7fbe773f7fbe77ull, dd = 0xbdcccccdbdcccccdull, ss;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(a[i-1]), "f"(a[i]));
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xx) : "l"(cc), "l"(dd));
          float x0, x1; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(xx));
          float e0 = ex2(x0), e1 = ex2(x1);
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ss) : "f"(e0), "f"(e1));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc2) : "l"(ss));
          unsigned u = __byte_perm(__float_as_uint(e0) + 0x8000u, __float_as_uint(e1) + 0x8000u, 0x7632); m += __uint_as_float(u);
          a[i-1] = e0; a[i] = e1; } }
      if (MODE == 13 || MODE == 14 || MODE == 15) { if (i & 1) { // the kernel's packed loop: FFMA2 + 2 EX2 + FADD2 + F2FP (13), without F2FP (14), PRMT pack (15)
          unsigned long long xx, cc = 0x3f7fbe773f7fbe77ull, dd = 0xbdcccccdbdcccccdull, ss;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(a[i-1]), "f"(a[i]));
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xx) : "l"(cc), "l"(dd));
          float x0, x1; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(xx));
          float e0 = ex2(x0), e1 = ex2(x1);
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ss) : "f"(e0), "f"(e1));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc2) : "l"(ss));
          if (MODE == 13) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(e1), "f"(e0)); m += __uint_as_float(u); }
          if (MODE == 15) { unsigned u = __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632); m += __uint_as_float(u); }
          a[i-1] = e0; a[i] = e1; } }
      if (MODE == 5) { float e = ex2(fmaf(a[i], 0.999f, -0.1f)); s += e; a[i] = e; asm volatile("max.f32 %0, %0, %1;" : "+f"(m) : "f"(e)); }
    }
  }
  long long t1 = clock64();
  float r = s + m + __uint_as_float((unsigned)acc2) + __uint_as_float((unsigned)(acc2 >> 32));
  for (int i = 0; i < 16; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  int iters = 4096;
  k<MODE><<<148, warps * 32>>>(out, iters, cyc);
  k<MODE><<<148, warps * 32>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double per_smsp_warp_instr = (double)c / ((double)iters * 16 * (warps / 4.0));
  printf("%-28s warps/SM %2d: %.2f cycles per (loop element) per SMSP-warp\n", name, warps, per_smsp_warp_instr);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 12, 16}) {
    run<0>("MUFU.EX2", w);
    run<1>("FFMA", w);
    run<4>("FMNMX", w);
    run<3>("FFMA+EX2+FADD", w);
    run<5>("FFMA+EX2+FADD+FMNMX", w);
    run<6>("F2FP.BF16x2", w);
    run<8>("EX2.bf16x2", w);
    run<9>("EX2.f16x2", w);
    run<10>("pair f16x2 pipeline (per elem)", w);
    run<7>("FFMA+EX2+FADD+F2FP/2", w);
    run<11>("FFMA+EX2+FADD+intpack", w);
    run<12>("FFMA2+EX2+FADD2+intpack", w);
    run<13>("FFMA2+EX2+FADD2+F2FP/2", w);
    run<14>("FFMA2+EX2+FADD2", w);
    run<15>("FFMA2+EX2+FADD2+PRMT(trunc)", w);
  }
  return 0;
}
