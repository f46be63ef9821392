// FP64 tensor-instruction micro-benchmark for sm_100a: throughput of mma.sync f64 shapes (m8n8k4, m16n8k4, m16n8k8,
// m16n8k16) against warps per SM and independent accumulator chains per warp.  Decides the fragment shape and the
// occupancy the Legendre kernels need.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 dmma_shapes.cu -o dmma_shapes
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int CHAINS>
__global__ void k(double *out, int iters) {
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int i = 0; i < 4; ++i) b[i] = 1.0 - 1e-9 * (threadIdx.x + i);
  double c[CHAINS][4];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
      if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      if (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      if (SHAPE == 3)
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, "
            "{%0,%1,%2,%3};\n"
            : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
            : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
              "d"(b[2]), "d"(b[3]));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

template <int SHAPE, int CHAINS>
void run(int sms, double *d) {
  const double flops_per[4] = {512.0, 1024.0, 2048.0, 4096.0};
  const char *names[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int wps_list[] = {4, 8, 12, 16, 24, 32};   // warps per SM
  for (int wps : wps_list) {
    const int threads = wps * 32 > 1024 ? 1024 : wps * 32;
    const int ctas_per_sm = (wps * 32) / threads;
    const int iters = 4096 * 8 / (SHAPE == 0 ? 1 : SHAPE == 1 ? 2 : SHAPE == 2 ? 4 : 8);
    k<SHAPE, CHAINS><<<sms * ctas_per_sm, threads>>>(d, iters);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      cudaEventRecord(e0);
      k<SHAPE, CHAINS><<<sms * ctas_per_sm, threads>>>(d, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    double fl = (double)sms * wps * iters * CHAINS * flops_per[SHAPE];
    printf("%-9s chains %2d warps/SM %2d : %7.2f TFLOP/s (%.3f ms)\n", names[SHAPE], CHAINS, wps, fl / (best * 1e-3) / 1e12,
           best);
  }
}

__global__ void kd(double *out, int iters) {
  double c[16];
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

void run_dfma(int sms, double *d) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int wps : {8, 16, 32}) {
    const int iters = 4096;
    kd<<<sms * wps / 8, 256>>>(d, iters);
    cudaEventRecord(e0);
    kd<<<sms * wps / 8, 256>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA      chains 16 warps/SM %2d : %7.2f TFLOP/s\n", wps, (double)sms * wps * 32 * iters * 16 * 2 / (ms * 1e-3) / 1e12);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *d;
  cudaMalloc(&d, 8);
  run<0, 16>(sms, d);
  run<0, 8>(sms, d);
  run<0, 4>(sms, d);
  run<3, 4>(sms, d);   // ptxas splits the larger shapes into DMMA.8x8x4 (cuobjdump): same pipe, same rate
  run_dfma(sms, d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
