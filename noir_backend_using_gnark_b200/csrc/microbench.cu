// Integer-pipe microbenchmarks: the measured denominators for the MSM / NTT rooflines (SURVEY.md §8d asks for an
// IMAD / IMAD.WIDE issue-rate measurement committed next to the results; MEASURED_PEAKS.json has no integer peak).
//   which = 0: IMAD.WIDE.U32 (32x32+64 -> 64) multiply-accumulates per second, 8 independent chains per thread
//   which = 1: fp Montgomery multiplications per second with this library's fe_mul, 2 independent chains per thread
//   which = 2: fr Montgomery multiplications per second
//   which = 3: FP64 DFMA per second (8 independent chains per thread) -- the other multiplier array of the SM
//   which = 4: IMAD.WIDE.U32 per second while the same threads also issue DFMA (1:1), to see whether the pipes overlap
//   which = 5 / 6: fp multiplications per second in the schoolbook CIOS form / with the Karatsuba product (field.cuh)
//   which = 7: fp PRODUCTS per second through the two-product sweep fe_mul2add (a*b + c*d with one reduction); every
//              thread first checks it against fe_add(fe_mul, fe_mul) on its operands and the kernel traps on a mismatch
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

__global__ void __launch_bounds__(256) mb_dfma_kernel(double* out, unsigned seed) {
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = 1.0 + (double)(seed + threadIdx.x + k) * 1e-9;
  const double a = 1.0 + (double)(seed & 7) * 1e-12, b = (double)(blockIdx.x & 3) * 1e-15;
  for (int it = 0; it < MB_ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int k = 0; k < 8; k++) acc[k] = __fma_rz(acc[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += acc[k];
  if (s == 0.12345) out[0] = s;
}

__global__ void __launch_bounds__(256) mb_mixed_kernel(unsigned long long* out, unsigned seed) {
  unsigned a0 = seed + threadIdx.x, a1 = seed * 3 + blockIdx.x, a2 = a0 ^ 0x9e3779b9u, a3 = a1 ^ 0x7f4a7c15u;
  unsigned acc[2][8];
  double facc[8];
#pragma unroll
  for (int k = 0; k < 2; k++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[k][j] = seed + k * 8 + j;
#pragma unroll
  for (int k = 0; k < 8; k++) facc[k] = 1.0 + (double)(seed + threadIdx.x + k) * 1e-9;
  const double fa = 1.0 + (double)(seed & 7) * 1e-12, fb = (double)(blockIdx.x & 3) * 1e-15;
  unsigned b = seed | 1u;
  for (int it = 0; it < MB_ITERS / 2; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int k = 0; k < 2; k++) madw4(acc[k], a0, a1, a2, a3, b + u);  // 2 x 8 IMAD.WIDE
#pragma unroll
      for (int k = 0; k < 8; k++) facc[k] = __fma_rz(facc[k], fa, fb);   // 8 DFMA  (x2 below)
#pragma unroll
      for (int k = 0; k < 8; k++) facc[k] = __fma_rz(facc[k], fa, fb);
    }
  }
  unsigned s = 0;
  double fs = 0;
#pragma unroll
  for (int k = 0; k < 2; k++)
#pragma unroll
    for (int j = 0; j < 8; j++) s ^= acc[k][j];
#pragma unroll
  for (int k = 0; k < 8; k++) fs += facc[k];
  if (s == 0x12345678u && fs == 0.12345) out[0] = s;
}

template <class P, int FORM>
__global__ void __launch_bounds__(256) mb_mul_form_kernel(uint4* out, unsigned seed) {
  // FORM 0: schoolbook CIOS, 1: Karatsuba product + reduction sweep; also checks the two against each other once
  Fe<P> x = fe_one<P>(), y = fe_one<P>(), m = fe_one<P>();
  x.l[0] ^= seed + threadIdx.x;
  y.l[1] ^= seed + blockIdx.x + 7u * threadIdx.x;
  m.l[2] ^= seed + 3u * threadIdx.x;
  for (int it = 0; it < MB_ITERS; it++) {
    if (FORM == 0) {
      x = fe_mul_schoolbook(x, m);
      y = fe_mul_schoolbook(y, m);
    } else {
      x = fe_mul_k(x, m);
      y = fe_mul_k(y, m);
    }
  }
  Fe<P> s = fe_add(x, y);
  if (s.l[0] == 0x12345678u && s.l[7] == 0x9abcdef0u) fe_store(out, s);
}

template <class P>
__global__ void __launch_bounds__(256) mb_mul2add_kernel(uint4* out, unsigned seed) {
  Fe<P> x = fe_one<P>(), y = fe_one<P>(), m = fe_one<P>(), k = fe_one<P>();
  x.l[0] ^= seed + threadIdx.x;
  y.l[1] ^= seed + blockIdx.x + 7u * threadIdx.x;
  m.l[2] ^= seed + 3u * threadIdx.x;
  k.l[3] ^= seed + 5u * threadIdx.x;
  {
    // extreme operands too: modulus - 1 - (small)
    Fe<P> big = fe_neg(x), big2 = fe_neg(m);
    if (!fe_eq(fe_mul2add(x, m, y, k), fe_add(fe_mul(x, m), fe_mul(y, k))) ||
        !fe_eq(fe_mul2add(big, big2, big, big), fe_add(fe_mul(big, big2), fe_mul(big, big))))
      __trap();
  }
  for (int it = 0; it < MB_ITERS; it++) {
    x = fe_mul2add(x, m, y, k);
    y = fe_mul2add(y, k, x, m);
  }
  Fe<P> s = fe_add(x, y);
  if (s.l[0] == 0x12345678u && s.l[7] == 0x9abcdef0u) fe_store(out, s);
}

template <class P>
__global__ void __launch_bounds__(256) mb_mul_kernel(uint4* out, unsigned seed) {
  Fe<P> x = fe_one<P>(), y = fe_one<P>(), m = fe_one<P>();
  // every operand depends on the lane: warp-uniform chains would be moved to the uniform datapath (UIMAD) and
  // report a rate no per-thread field code can reach
  x.l[0] ^= seed + threadIdx.x;
  y.l[1] ^= seed + blockIdx.x + 7u * threadIdx.x;
  m.l[2] ^= seed + 3u * threadIdx.x;
  for (int it = 0; it < MB_ITERS; it++) {
    x = fe_mul(x, m);
    y = fe_mul(y, m);
  }
  Fe<P> s = fe_add(x, y);
  if (s.l[0] == 0x12345678u && s.l[7] == 0x9abcdef0u) fe_store(out, s);
}

int microbench_run(b200zk_ctx* ctx, int which, double* out_ops_per_s) {
  if (!out_ops_per_s || which < 0 || which > 7) return B200ZK_ERR_BAD_ARG;
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
    else if (which == 2) mb_mul_kernel<FrParams><<<blocks, threads, 0, ctx->stream>>>((uint4*)sink, 17u + rep);
    else if (which == 5) mb_mul_form_kernel<FpParams, 0><<<blocks, threads, 0, ctx->stream>>>((uint4*)sink, 17u + rep);
    else if (which == 6) mb_mul_form_kernel<FpParams, 1><<<blocks, threads, 0, ctx->stream>>>((uint4*)sink, 17u + rep);
    else if (which == 7) mb_mul2add_kernel<FpParams><<<blocks, threads, 0, ctx->stream>>>((uint4*)sink, 17u + rep);
    else if (which == 3) mb_dfma_kernel<<<blocks, threads, 0, ctx->stream>>>((double*)sink, 17u + rep);
    else mb_mixed_kernel<<<blocks, threads, 0, ctx->stream>>>((unsigned long long*)sink, 17u + rep);
    B200ZK_LAUNCH_CHECK(ctx, "microbench kernel");
    B200ZK_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    B200ZK_CUDA(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    B200ZK_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    double per_thread = MB_ITERS * 2.0;                  // which = 1, 2: two multiplication chains
    if (which == 0) per_thread = (MB_ITERS / 2) * 64.0;  // 4 iterations x 4 chains x 4 IMAD.WIDE
    if (which == 3) per_thread = MB_ITERS * 32.0;
    if (which == 7) per_thread = MB_ITERS * 4.0;         // two sweeps of two products each
    if (which == 4) per_thread = (MB_ITERS / 2) * 32.0;  // IMAD.WIDE only (4 x 8); the DFMA count is twice that
    double ops = (double)blocks * threads * per_thread;
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
