// Upper bound of the Legendre inner loop on sm_100a: 8 warps per CTA, each a 32x32 warp tile fed from shared memory
// (4 A + 4 B LDS.64 and 16 DMMA.8x8x4 per k-step of 4), no global traffic.  Variants: CTAs per SM, a __syncthreads per
// chunk of 4 k-steps, and plain FP64 work mixed in (the parity fold), to see what each costs the tensor pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 dmma_loop.cu -o dmma_loop
#include <cstdio>
#include <cuda_runtime.h>

#define LD 20
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int SYNC, int FOLD>
__global__ void __launch_bounds__(256, 2) k(double *out, int chunks) {
  extern __shared__ double sm[];
  double(*A)[64][LD] = reinterpret_cast<double(*)[64][LD]>(sm);                  // [2][64][LD]
  double(*F)[64][LD] = reinterpret_cast<double(*)[64][LD]>(sm + 2 * 64 * LD);    // [2][64][LD]
  double *R = sm + 4 * 64 * LD;                                                  // raw 2 x 32 x 16 complex
  for (int i = threadIdx.x; i < 4 * 64 * LD + 2048; i += 256) sm[i] = 1e-3 * (i % 17);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int par = warp >> 2, wr = (warp >> 1) & 1, wc = warp & 1, fr = lane >> 2, fk = lane & 3;
  double acc[4][4][2];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int c = 0; c < chunks; ++c) {
    if (SYNC) __syncthreads();
    if (FOLD) {   // the per-chunk fold of the forward kernel: 2 x (2 LDS.128, 4 add, 4 mul, 4 STS.64) per thread
      const int kk = threadIdx.x & 15;
      const double wk = R[kk];
      for (int j = 0; j < 2; ++j) {
        const int kzl = (threadIdx.x >> 4) + 16 * j;
        double2 t = reinterpret_cast<double2 *>(R)[kzl * 16 + kk], b = reinterpret_cast<double2 *>(R)[512 + kzl * 16 + kk];
        F[0][2 * kzl][kk + (c & 1)] = (t.x + b.x) * wk;
        F[0][2 * kzl + 1][kk + (c & 1)] = (t.y + b.y) * wk;
        F[1][2 * kzl][kk + (c & 1)] = (t.x - b.x) * wk;
        F[1][2 * kzl + 1][kk + (c & 1)] = (t.y - b.y) * wk;
      }
      if (SYNC) __syncthreads();
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int kq = ks * 4 + fk;
      double af[4], bf[4];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) af[mt] = A[par][(2 * mt + wr) * 8 + fr][kq];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bf[nt] = F[par][wc * 32 + nt * 8 + fr][kq];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
  }
  double s = 0.0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
  if (s == 123.456) out[0] = s;
}

template <int SYNC, int FOLD>
void run(int sms, double *d, int ctas_per_sm, const char *name) {
  const size_t smem = (4 * 64 * LD + 2048) * sizeof(double);
  cudaFuncSetAttribute(k<SYNC, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int chunks = 4096;
  k<SYNC, FOLD><<<sms * ctas_per_sm, 256, smem>>>(d, chunks);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    k<SYNC, FOLD><<<sms * ctas_per_sm, 256, smem>>>(d, chunks);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const double fl = (double)sms * ctas_per_sm * 8 * chunks * 4 * 16 * 512.0;
  printf("%-28s CTAs/SM %d : %6.2f TFLOP/s\n", name, ctas_per_sm, fl / (best * 1e-3) / 1e12);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *d;
  cudaMalloc(&d, 8);
  for (int c = 1; c <= 2; ++c) {
    run<0, 0>(sms, d, c, "LDS + DMMA");
    run<1, 0>(sms, d, c, "LDS + DMMA + sync/chunk");
    run<0, 1>(sms, d, c, "LDS + DMMA + fold");
    run<1, 1>(sms, d, c, "LDS + DMMA + fold + 2 sync");
  }
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
