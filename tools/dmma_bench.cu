// dmma_bench.cu -- FP64 tensor-core (DMMA, mma.sync m8n8k4) vs FMA pipe on the Riccati block products of an 18-wide stage
// (config #5, AWE: nx = 9, nz = nx + nu + ns + nsc = 18).  BASELINE.json's north_star asks for this comparison before any
// tensor-core use on the solve path: "FP64 DMMA tensor cores are used only for the Riccati block products when nx >= 16 (AWE
// model), and only where ncu shows they beat the FMA pipe".
//
// Product measured: the stage block of the backward recursion  F = Q + [A B]' P [A B]  (tunempc_b200/csrc/tmpc_qp.cuh,
// tm_qp_factor), nz x nz output, inner dimension nx, done as  T = P [A B] (nx x nz)  then  F = Q + [A B]' T  -- for B
// independent instances, one warp per instance (the layout of the warp-per-instance QP kernel), operands in shared memory.
//   fma   : lanes own output entries, plain DFMA loops
//   dmma  : 8 x 8 output tiles, k in steps of 4, operands zero-padded to multiples of 8 / 4 (18 -> 24, 9 -> 12)
// Both kernels write F to global memory; the host checks them against each other.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/dmma_bench.cu -o tools/dmma_bench
//   tools/dmma_bench [instances = 131072] [repeats = 20]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define NXB 9
#define NZB 18
#define NXP 12     /* nx padded to a multiple of 4 (k) and 8-row tiles need 16: see below */
#define NZP 24     /* nz padded to a multiple of 8 */
#define WARPS 4

// per-instance input: AB (nx x nz), P (nx x nx, symmetric), Q (nz x nz, symmetric): 162 + 81 + 324 doubles
#define IN_DBL (NXB * NZB + NXB * NXB + NZB * NZB)

__global__ void __launch_bounds__(WARPS * 32) k_fma(const double* __restrict__ in, double* __restrict__ out, int n, int inner) {
  __shared__ double sAB[WARPS][NXB * NZB], sP[WARPS][NXB * NXB], sT[WARPS][NXB * NZB];
  const int w = threadIdx.x / 32, lane = threadIdx.x & 31;
  const long inst = (long)blockIdx.x * WARPS + w;
  if (inst >= n) return;
  const double* src = in + inst * IN_DBL;
  for (int e = lane; e < NXB * NZB; e += 32) sAB[w][e] = src[e];
  for (int e = lane; e < NXB * NXB; e += 32) sP[w][e] = src[NXB * NZB + e];
  __syncwarp();
  double acc[(NZB * NZB + 31) / 32];
  for (int rep = 0; rep < inner; ++rep) {
    for (int e = lane; e < NXB * NZB; e += 32) {          // T = P [A B]
      const int i = e / NZB, c = e % NZB;
      double v = 0.0;
#pragma unroll
      for (int l = 0; l < NXB; ++l) v = fma(sP[w][i * NXB + l], sAB[w][l * NZB + c], v);
      sT[w][e] = v;
    }
    __syncwarp();
    int q = 0;
    for (int e = lane; e < NZB * NZB; e += 32, ++q) {     // F = Q + [A B]' T
      const int a = e / NZB, c = e % NZB;
      double v = rep == 0 ? src[NXB * NZB + NXB * NXB + e] : acc[q];
#pragma unroll
      for (int i = 0; i < NXB; ++i) v = fma(sAB[w][i * NZB + a], sT[w][i * NZB + c], v);
      acc[q] = v;
    }
    __syncwarp();
  }
  int q = 0;
  for (int e = lane; e < NZB * NZB; e += 32, ++q) out[inst * NZB * NZB + e] = acc[q];
}

// D (8x8) += A (8x4, row) * B (4x8, col):  lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4) .. +1]
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(WARPS * 32) k_dmma(const double* __restrict__ in, double* __restrict__ out, int n, int inner) {
  // padded operands: AB as [NXP][NZP] (rows 9..11 and columns 18..23 zero), P as [16][NXP] (8-row tiles: 16 rows), T as [16][NZP]
  __shared__ double sAB[WARPS][NXP * NZP], sP[WARPS][16 * NXP], sT[WARPS][16 * NZP];
  const int w = threadIdx.x / 32, lane = threadIdx.x & 31;
  const long inst = (long)blockIdx.x * WARPS + w;
  if (inst >= n) return;
  const double* src = in + inst * IN_DBL;
  for (int e = lane; e < NXP * NZP; e += 32) { const int i = e / NZP, c = e % NZP; sAB[w][e] = (i < NXB && c < NZB) ? src[i * NZB + c] : 0.0; }
  for (int e = lane; e < 16 * NXP; e += 32) { const int i = e / NXP, c = e % NXP; sP[w][e] = (i < NXB && c < NXB) ? src[NXB * NZB + i * NXB + c] : 0.0; }
  for (int e = lane; e < 16 * NZP; e += 32) sT[w][e] = 0.0;
  __syncwarp();
  const int r = lane / 4, kq = lane % 4;
  double f[9][2];                                           // 3 x 3 output tiles of F, two entries per lane each
  for (int rep = 0; rep < inner; ++rep) {
    // T (16 x NZP, rows >= 9 stay zero) = P (16 x NXP) * AB (NXP x NZP): 2 x 3 tiles, 3 k-steps
#pragma unroll
    for (int ti = 0; ti < 2; ++ti)
#pragma unroll
      for (int tj = 0; tj < 3; ++tj) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int kk = 0; kk < NXP / 4; ++kk)
          dmma(d0, d1, sP[w][(ti * 8 + r) * NXP + kk * 4 + kq], sAB[w][(kk * 4 + kq) * NZP + tj * 8 + r]);
        sT[w][(ti * 8 + r) * NZP + tj * 8 + 2 * kq] = d0;
        sT[w][(ti * 8 + r) * NZP + tj * 8 + 2 * kq + 1] = d1;
      }
    __syncwarp();
    // F (NZP x NZP) = Q + AB' (NZP x NXP) * T (NXP x NZP): 3 x 3 tiles, 3 k-steps;  A operand = AB' read transposed
#pragma unroll
    for (int ti = 0; ti < 3; ++ti)
#pragma unroll
      for (int tj = 0; tj < 3; ++tj) {
        double d0, d1;
        if (rep == 0) {
          const int a = ti * 8 + r, c = tj * 8 + 2 * kq;
          d0 = (a < NZB && c < NZB) ? src[NXB * NZB + NXB * NXB + a * NZB + c] : 0.0;
          d1 = (a < NZB && c + 1 < NZB) ? src[NXB * NZB + NXB * NXB + a * NZB + c + 1] : 0.0;
        } else { d0 = f[ti * 3 + tj][0]; d1 = f[ti * 3 + tj][1]; }
#pragma unroll
        for (int kk = 0; kk < NXP / 4; ++kk)
          dmma(d0, d1, sAB[w][(kk * 4 + kq) * NZP + ti * 8 + r], sT[w][(kk * 4 + kq) * NZP + tj * 8 + r]);
        f[ti * 3 + tj][0] = d0; f[ti * 3 + tj][1] = d1;
      }
    __syncwarp();
  }
#pragma unroll
  for (int ti = 0; ti < 3; ++ti)
#pragma unroll
    for (int tj = 0; tj < 3; ++tj) {
      const int a = ti * 8 + r, c = tj * 8 + 2 * kq;
      if (a < NZB && c < NZB) out[inst * NZB * NZB + a * NZB + c] = f[ti * 3 + tj][0];
      if (a < NZB && c + 1 < NZB) out[inst * NZB * NZB + a * NZB + c + 1] = f[ti * 3 + tj][1];
    }
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 131072, reps = argc > 2 ? atoi(argv[2]) : 20;
  std::vector<double> h((size_t)n * IN_DBL);
  unsigned s = 12345u;
  for (auto& v : h) { s = s * 1664525u + 1013904223u; v = ((s >> 8) & 0xffff) / 65536.0 - 0.5; }
  double *din, *o1, *o2;
  cudaMalloc(&din, h.size() * sizeof(double));
  cudaMalloc(&o1, (size_t)n * NZB * NZB * sizeof(double));
  cudaMalloc(&o2, (size_t)n * NZB * NZB * sizeof(double));
  cudaMemcpy(din, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice);
  const int grid = (n + WARPS - 1) / WARPS;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double useful = 2.0 * NXB * NXB * NZB + 2.0 * NZB * NZB * NXB;         // flops of one (T, F) product pair
  const double padded = 2.0 * 16 * NXP * NZP + 2.0 * NZP * NZP * NXP;           // flops the DMMA tiles execute
  printf("{\"instances\": %d, \"useful_flop_per_product\": %.0f, \"dmma_executed_flop_per_product\": %.0f", n, useful, padded);
  for (int inner : {1, 16}) {
    for (int which = 0; which < 2; ++which) {
      float best = 1e30f;
      for (int rp = 0; rp < reps + 3; ++rp) {
        cudaEventRecord(e0);
        if (which == 0) k_fma<<<grid, WARPS * 32>>>(din, o1, n, inner); else k_dmma<<<grid, WARPS * 32>>>(din, o2, n, inner);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rp >= 3 && ms < best) best = ms;
      }
      printf(", \"%s_inner%d_ms\": %.4f, \"%s_inner%d_useful_tflops\": %.3f", which ? "dmma" : "fma", inner, best, which ? "dmma" : "fma", inner,
             useful * inner * n / (best * 1e-3) / 1e12);
    }
    std::vector<double> a((size_t)n * NZB * NZB), b(a.size());
    cudaMemcpy(a.data(), o1, a.size() * sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), o2, b.size() * sizeof(double), cudaMemcpyDeviceToHost);
    double err = 0.0, mx = 0.0;
    for (size_t i = 0; i < a.size(); ++i) { err = fmax(err, fabs(a[i] - b[i])); mx = fmax(mx, fabs(a[i])); }
    printf(", \"maxdiff_inner%d\": %.3e, \"maxabs_inner%d\": %.3e", inner, err, inner, mx);
  }
  cudaError_t ce = cudaDeviceSynchronize();
  printf(", \"cuda\": \"%s\"}\n", cudaGetErrorString(ce));
  return ce != cudaSuccess;
}
