// GPU diagnostic: accuracy of rcp.approx.ftz.f64 (the seed of shape64_step's reciprocal) and of
// one / two Newton steps on it, over the range of total masses the SALP body takes (4.5 .. 5.6 kg).
#include <cstdio>
#include <cmath>
__global__ void probe(double lo, double hi, int n, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = lo + (hi - lo) * i / n;
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
  double r1 = r0 * (2.0 - x * r0);
  double r2 = r1 * (2.0 - x * r1);
  double ex = 1.0 / x;
  out[3 * i] = fabs(r0 - ex) / ex;
  out[3 * i + 1] = fabs(r1 - ex) / ex;
  out[3 * i + 2] = fabs(r2 - ex) / ex;
}
int main() {
  const int n = 1 << 20;
  double* d;
  cudaMalloc(&d, sizeof(double) * 3 * n);
  probe<<<n / 256, 256>>>(4.5, 5.6, n, d);
  double* h = new double[3 * n];
  cudaMemcpy(h, d, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost);
  double m[3] = {0, 0, 0};
  for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) m[k] = fmax(m[k], h[3 * i + k]);
  printf("max rel err: seed %.3e  newton1 %.3e  newton2 %.3e\n", m[0], m[1], m[2]);
  return 0;
}
