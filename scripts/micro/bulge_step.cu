// Microbenchmark (round-2 groundwork, not product code): latency of ONE bulge-chasing step of the band -> tridiagonal
// stage in a single CTA with the window in shared memory.  DESIGN.md section 5 estimates the two-stage reduction at
// n = 4096 with "~3 n dependent steps x 2-3 us"; this measures the per-step figure on the B200:
//     step = Householder vector from a length-b column (norm: warp shuffles + one barrier), then H W H on the
//            (2b x 2b) symmetric window it touches (left: b x 2b, right: 2b x b), everything in shared memory.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulge_step bulge_step.cu ; run: ./bulge_step
#include <cstdio>
#include <cuda_runtime.h>

template <int B, int T>
__global__ void __launch_bounds__(T) bulge_step_kernel(double* out, long long* cycles, int steps) {
  constexpr int W = 2 * B, LD = W + 1;
  extern __shared__ double win[];  // [W * LD], win[c * LD + r]
  __shared__ double v[B], tmp[W];
  __shared__ double s_red[T / 32], s_tau, s_beta;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < W * LD; i += T) {
    const int c = i / LD, r = i - c * LD;
    win[i] = r < W ? sin(0.37 * (r + 1) * (c + 2)) + (r == c ? 4.0 : 0.0) : 0.0;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    // ---- reflector from column 0, rows 1 .. B (refreshed each step so that the tail never vanishes) ----
    double x = 0.0;
    if (tid < B) {
      x = win[1 + tid] + 1e-3 * (double)((s * 7 + tid * 13) % 11);
      win[1 + tid] = x;
    }
    double p = (tid >= 1 && tid < B) ? x * x : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    if (lane == 0) s_red[warp] = p;
    __syncthreads();
    if (tid == 0) {
      double xn2 = 0.0;
      for (int w = 0; w < (B + 31) / 32; ++w) xn2 += s_red[w];
      const double alpha = win[1];
      const double beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
      s_tau = (beta - alpha) / beta;
      s_beta = 1.0 / (alpha - beta);
    }
    __syncthreads();
    if (tid < B) v[tid] = tid == 0 ? 1.0 : x * s_beta;
    __syncthreads();
    const double tau = s_tau;
    // ---- left: rows 1..B of every column:  col -= tau v (v^T col) ; one warp per column group ----
    for (int c = warp; c < W; c += T / 32) {
      double d = 0.0;
      for (int r = lane; r < B; r += 32) d = fma(v[r], win[c * LD + 1 + r], d);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      d *= tau;
      for (int r = lane; r < B; r += 32) win[c * LD + 1 + r] = fma(-d, v[r], win[c * LD + 1 + r]);
    }
    __syncthreads();
    // ---- right: columns 1..B of every row:  row -= tau (row v) v^T ; one thread per row ----
    if (tid < W) {
      double d = 0.0;
      for (int c = 0; c < B; ++c) d = fma(win[(1 + c) * LD + tid], v[c], d);
      tmp[tid] = tau * d;
    }
    __syncthreads();
    for (int i = tid; i < W * B; i += T) {
      const int c = i / W, r = i - c * W;
      win[(1 + c) * LD + r] = fma(-tmp[r], v[c], win[(1 + c) * LD + r]);
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) { cycles[0] = t1 - t0; out[0] = win[3 * LD + 2]; }
}

template <int B, int T>
void run(int steps) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const size_t smem = (size_t)(2 * B) * (2 * B + 1) * sizeof(double);
  cudaFuncSetAttribute(bulge_step_kernel<B, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bulge_step_kernel<B, T><<<1, T, smem>>>(out, cyc, 10);
  bulge_step_kernel<B, T><<<1, T, smem>>>(out, cyc, steps);
  long long h = 0; double o = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&o, out, 8, cudaMemcpyDeviceToHost);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("b=%d T=%d: %lld cycles/step = %.2f us at %d MHz (check %.6f) -> n=4096: 3n steps = %.1f ms on the critical path\n",
         B, T, h / steps, (double)h / steps / (khz / 1e3), khz / 1000, o, 3.0 * 4096 * (double)h / steps / (khz * 1e3) * 1e3);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  run<32, 128>(2000); run<32, 256>(2000); run<64, 256>(2000); run<64, 512>(2000);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
