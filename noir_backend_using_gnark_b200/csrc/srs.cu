// KZG SRS generation on the device: G1[i] = alpha^i * G  (kzg.NewSRS(size, alpha) of gnark-crypto v0.9.1,
// called at /root/reference/gnark_backend_ffi/backend/common.go:137 and main.go:176).  Fixed-base method:
// a table d * 2^(8w) * G (w < 32, d < 256) is built once per call; every output point is 32 mixed additions
// plus one inversion.  Used by the reference flow on an SRS-cache miss and by the benchmarks to synthesise bases.
#include "common.cuh"
#include "g1.cuh"

namespace b200zk {

// pw[k] = alpha^(2^k)
__global__ void srs_seed_kernel(const uint4* alpha, Fr* pw) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Fr a = fe_load<FrParams>(alpha);
  for (int k = 0; k < 32; k++) {
    pw[k] = a;
    a = fe_sqr(a);
  }
}

// table[w*256 + d] = d * 2^(8w) * G, affine
__global__ void srs_table_kernel(void* table) {
  const unsigned w = blockIdx.x, d = threadIdx.x;
  __shared__ G1XYZZ base_sh;
  if (d == 0) {
    G1XYZZ b;
    b.x = fe_one<FpParams>();               // G = (1, 2)
    b.y = fe_dbl(fe_one<FpParams>());
    b.zz = fe_one<FpParams>();
    b.zzz = fe_one<FpParams>();
    for (unsigned k = 0; k < 8 * w; k++) g1_double(b);
    base_sh = b;
  }
  __syncthreads();
  G1XYZZ base = base_sh;
  G1XYZZ acc = g1_xyzz_inf();
  for (int bit = 7; bit >= 0; bit--) {
    g1_double(acc);
    if ((d >> bit) & 1) g1_add(acc, base);
  }
  G1Affine r = g1_to_affine(acc);
  g1_store_affine(table, w * 256 + d, r);
}

__global__ void __launch_bounds__(128) srs_points_kernel(const Fr* __restrict__ pw, const void* __restrict__ table,
                                                         size_t first, size_t n, void* __restrict__ out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  // s = alpha^(first + i)
  Fr s = fe_one<FrParams>();
  size_t e = first + i;
  for (int k = 0; e != 0; k++, e >>= 1) {
    if (e & 1) s = fe_mul(s, pw[k]);
  }
  s = fe_from_mont(s);
  G1XYZZ acc = g1_xyzz_inf();
  for (unsigned w = 0; w < 32; w++) {
    unsigned d = (s.l[w >> 2] >> (8 * (w & 3))) & 0xffu;
    if (d) {
      G1Affine t = g1_load_affine(table, w * 256 + d);
      g1_add_mixed(acc, t);
    }
  }
  G1Affine r = g1_to_affine(acc);
  g1_store_affine(out, i, r);
}

int srs_generate_run(b200zk_ctx* ctx, const void* alpha_dev, size_t first, size_t n, void* out_dev) {
  if (first + n >= ((size_t)1 << 32)) return B200ZK_ERR_UNSUPPORTED;
  void* scratch = nullptr;
  const size_t table_bytes = 32 * 256 * 64;
  B200ZK_CUDA(ctx, cudaMalloc(&scratch, table_bytes + 32 * sizeof(Fr)));
  Fr* pw = (Fr*)((char*)scratch + table_bytes);
  srs_seed_kernel<<<1, 32, 0, ctx->stream>>>((const uint4*)alpha_dev, pw);
  B200ZK_LAUNCH_CHECK(ctx, "srs_seed_kernel");
  srs_table_kernel<<<32, 256, 0, ctx->stream>>>(scratch);
  B200ZK_LAUNCH_CHECK(ctx, "srs_table_kernel");
  if (n) {
    srs_points_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(pw, scratch, first, n, out_dev);
    B200ZK_LAUNCH_CHECK(ctx, "srs_points_kernel");
  }
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  B200ZK_CUDA(ctx, cudaFree(scratch));
  return B200ZK_OK;
}

}  // namespace b200zk
