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

// ---------------------------------------------------------------------------------------------------
// gnark compressed G1 encoding on the device: G1Affine.Bytes() / SetBytes() of gnark-crypto v0.9.1
// (32-byte big-endian X, two flag bits in the most significant byte: 10 = smallest y, 11 = largest y,
// 01 = infinity).  This is the format of the SRS file the reference re-reads and decompresses on EVERY FFI call
// (/root/reference/gnark_backend_ffi/backend/common.go:86-105: one fp square root per point, >= 10^6 points).
// ---------------------------------------------------------------------------------------------------
namespace b200zk {

__device__ __forceinline__ Fp fp_from_be(const uint8_t* b, bool mask_flags) {
  Fp v;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t* q = b + 28 - 4 * i;
    uint32_t w = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    v.l[i] = w;
  }
  if (mask_flags) v.l[7] &= 0x3fffffffu;
  return v;
}

__device__ __forceinline__ void fp_to_be(uint8_t* b, const Fp& v) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint8_t* q = b + 28 - 4 * i;
    q[0] = (uint8_t)(v.l[i] >> 24); q[1] = (uint8_t)(v.l[i] >> 16); q[2] = (uint8_t)(v.l[i] >> 8); q[3] = (uint8_t)v.l[i];
  }
}

// regular-form a > (p-1)/2 ?
__device__ __forceinline__ bool fp_lexicographically_largest(const Fp& a_regular) {
  // (p-1)/2 = 0x183227397098d014dc2822db40c0ac2ecbc0b548b438e5469e10460b6c3e7ea3
  const uint32_t half[8] = {0x6c3e7ea3u, 0x9e10460bu, 0xb438e546u, 0xcbc0b548u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
  for (int i = 7; i >= 0; i--) {
    if (a_regular.l[i] != half[i]) return a_regular.l[i] > half[i];
  }
  return false;
}

// flags[i]: 0 ok, 1 = not on the curve / malformed
__global__ void __launch_bounds__(128) srs_decompress_kernel(const uint8_t* __restrict__ in, size_t n, void* __restrict__ out,
                                                             unsigned* __restrict__ bad) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* b = in + 32 * i;
  const unsigned flag = b[0] >> 6;
  G1Affine p;
  if (flag == 1) {  // infinity
    p.x = fe_zero<FpParams>();
    p.y = fe_zero<FpParams>();
    g1_store_affine(out, i, p);
    return;
  }
  if (flag == 0) {  // the uncompressed marker cannot appear in a 32-byte slot
    atomicAdd(bad, 1u);
    return;
  }
  Fp x = fe_to_mont(fp_from_be(b, true));
  // y = (x^3 + 3)^((p+1)/4)   (p = 3 mod 4)
  Fp three = fe_add(fe_add(fe_one<FpParams>(), fe_one<FpParams>()), fe_one<FpParams>());
  Fp rhs = fe_add(fe_mul(fe_sqr(x), x), three);
  // (p+1)/4 = 0xc19139cb84c680a6e14116da060561765e05aa45a1c72a34f082305b61f3f52
  const uint32_t e[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u, 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};
  Fp y = fe_one<FpParams>();
  for (int k = 7; k >= 0; k--) {
    for (int bit = 31; bit >= 0; bit--) {
      y = fe_sqr(y);
      if ((e[k] >> bit) & 1u) y = fe_mul(y, rhs);
    }
  }
  if (!fe_eq(fe_sqr(y), rhs)) {
    atomicAdd(bad, 1u);
    return;
  }
  const bool largest = fp_lexicographically_largest(fe_from_mont(y));
  if (largest != (flag == 3)) y = fe_neg(y);
  p.x = x;
  p.y = y;
  g1_store_affine(out, i, p);
}

__global__ void __launch_bounds__(128) srs_compress_kernel(const void* __restrict__ in, size_t n, uint8_t* __restrict__ out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p = g1_load_affine(in, i);
  uint8_t* b = out + 32 * i;
  if (g1_is_inf(p)) {
    for (int k = 0; k < 32; k++) b[k] = 0;
    b[0] = 0x40;
    return;
  }
  fp_to_be(b, fe_from_mont(p.x));
  b[0] |= fp_lexicographically_largest(fe_from_mont(p.y)) ? 0xC0 : 0x80;
}

int srs_decompress_run(b200zk_ctx* ctx, const void* compressed_host, size_t n, void* out_dev, unsigned* bad_count) {
  void* in_dev = nullptr;
  unsigned* bad = nullptr;
  B200ZK_CUDA(ctx, cudaMalloc(&in_dev, n * 32 + 16));
  bad = (unsigned*)((char*)in_dev + n * 32);
  cudaError_t e = cudaMemcpyAsync(in_dev, compressed_host, n * 32, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(bad, 0, 4, ctx->stream);
  if (e == cudaSuccess) {
    srs_decompress_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const uint8_t*)in_dev, n, out_dev, bad);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(bad_count, bad, 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(in_dev);
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "srs_decompress");
  return B200ZK_OK;
}

int srs_compress_run(b200zk_ctx* ctx, const void* points_dev, size_t n, void* out_host) {
  void* out_dev = nullptr;
  B200ZK_CUDA(ctx, cudaMalloc(&out_dev, n * 32));
  srs_compress_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(points_dev, n, (uint8_t*)out_dev);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, out_dev, n * 32, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(out_dev);
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "srs_compress");
  return B200ZK_OK;
}

}  // namespace b200zk
