// Integer-pipe microbenchmarks: the measured denominators for the MSM / NTT rooflines (SURVEY.md §8d asks for an
// IMAD / IMAD.WIDE issue-rate measurement committed next to the results; MEASURED_PEAKS.json has no integer peak).
//   which = 0: IMAD.WIDE.U32 (32x32+64 -> 64) multiply-accumulates per second, 8 independent chains per thread
//   which = 1: fp Montgomery multiplications per second with this library's fe_mul, 2 independent chains per thread
//   which = 2: fr Montgomery multiplications per second
#include "common.cuh"
#include "field.cuh"

namespace b200zk {

static constexpr int MB_ITERS = 2048;

__global__ void __launch_bounds__(256) mb_imad_kernel(unsigned long long* out, unsigned seed) {
  // 4 independent 256-bit accumulators, each fed by the same 4-wide IMAD.WIDE.U32.X carry chain fe_mul is built from
  unsigned a0 = seed + threadIdx.x, a1 = seed * 3 + blockIdx.x, a2 = a0 ^ 0x9e3779b9u, a3 = a1 ^ 0x7f4a7c15u;
  unsigned acc[4][8];
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[k][j] = seed + k * 8 + j;
  unsigned b = seed | 1u;
  for (int it = 0; it < MB_ITERS / 2; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int k = 0; k < 4; k++) madw4(acc[k], a0, a1, a2, a3, b + u);
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= acc[k][j];
  if (s == 0x12345678u && a0 == 0x9abcdef0u) out[0] = s;  // practically never: keeps the chains alive
}

template <class P>
__global__ void __launch_bounds__(256) mb_mul_kernel(uint4* out, unsigned seed) {
  Fe<P> x = fe_one<P>(), y = fe_one<P>(), m = fe_one<P>();
  x.l[0] ^= seed + threadIdx.x;
  y.l[1] ^= seed + blockIdx.x;
  m.l[2] ^= seed;
  for (int it = 0; it < MB_ITERS; it++) {
    x = fe_mul(x, m);
    y = fe_mul(y, m);
  }
  Fe<P> s = fe_add(x, y);
  if (s.l[0] == 0x12345678u && s.l[7] == 0x9abcdef0u) fe_store(out, s);
}

int microbench_run(b200zk_ctx* ctx, int which, double* out_ops_per_s) {
  if (!out_ops_per_s || which < 0 || which > 2) return B200ZK_ERR_BAD_ARG;
  void* sink = nullptr;
  B200ZK_CUDA(ctx, cudaMalloc(&sink, 64));
  cudaEvent_t e0, e1;
  B200ZK_CUDA(ctx, cudaEventCreate(&e0));
  B200ZK_CUDA(ctx, cudaEventCreate(&e1));
  const unsigned blocks = (unsigned)ctx->sm_count * 16, threads = 256;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    B200ZK_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (which == 0) mb_imad_kernel<<<blocks, threads, 0, ctx->stream>>>((unsigned long long*)sink, 17u + rep);
    else if (which == 1) mb_mul_kernel<FpParams><<<blocks, threads, 0, ctx->stream>>>((uint4*)sink, 17u + rep);
    else mb_mul_kernel<FrParams><<<blocks, threads, 0, ctx->stream>>>((uint4*)sink, 17u + rep);
    B200ZK_LAUNCH_CHECK(ctx, "microbench kernel");
    B200ZK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    B200ZK_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    B200ZK_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    double ops = (double)blocks * threads * (which == 0 ? (MB_ITERS / 2) * 64.0 : MB_ITERS * 2.0);
    double rate = ops / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;  // rep 0 is the warm-up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  *out_ops_per_s = best;
  return B200ZK_OK;
}

}  // namespace b200zk
