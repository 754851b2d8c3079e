// C ABI of libb200zk (declared in include/b200zk.h).  No exceptions, no aborts: every failure is an error code.
#include <new>
#include "common.cuh"

using namespace b200zk;

namespace {

int enter(b200zk_ctx* ctx) {
  if (!ctx) return B200ZK_ERR_BAD_ARG;
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaSetDevice");
  return B200ZK_OK;
}

void free_buf(DeviceBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

}  // namespace

extern "C" {

int b200zk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int b200zk_init(int device, b200zk_ctx** out) {
  if (!out) return B200ZK_ERR_BAD_ARG;
  *out = nullptr;
  int n = b200zk_device_count();
  if (n <= 0 || device < 0 || device >= n) return B200ZK_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) {
    cudaGetLastError();
    return B200ZK_ERR_NO_DEVICE;
  }
  b200zk_ctx* ctx = new (std::nothrow) b200zk_ctx();
  if (!ctx) return B200ZK_ERR_OOM;
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return B200ZK_ERR_CUDA;
  }
  if (cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return B200ZK_ERR_CUDA;
  }
  for (auto& ev : ctx->ev_chunk)
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      delete ctx;
      return B200ZK_ERR_CUDA;
    }
  ctx->ws[0].stream = ctx->stream;
  for (int l = 0; l < MSM_LANES; l++) {
    if ((l > 0 && cudaStreamCreateWithFlags(&ctx->ws[l].stream, cudaStreamNonBlocking) != cudaSuccess) ||
        cudaEventCreateWithFlags(&ctx->ws[l].done, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      delete ctx;
      return B200ZK_ERR_CUDA;
    }
  }
  // tuning / test default of b200zk_msm_set_pair_rounds (lets a whole test run exercise one setting)
  if (const char* pr = getenv("B200ZK_MSM_PAIR_ROUNDS")) {
    const int r = atoi(pr);
    if (r >= -1 && r <= 6) ctx->msm_pair_rounds = r;
  }
  if (const char* sm = getenv("B200ZK_MSM_SCATTER_PASSES")) {
    const int v = atoi(sm);
    if (v >= 0 && v <= 256) ctx->msm_scatter_passes = v;
  }
  if (const char* pk = getenv("B200ZK_MSM_PAIR_KMAX")) {
    const int k = atoi(pk);
    if (k >= 0 && k <= 4096) ctx->msm_pair_kmax = k;
  }
  *out = ctx;
  return B200ZK_OK;
}

void b200zk_destroy(b200zk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ntt_free_domains(ctx);
  free_buf(ctx->stage);
  for (int l = 0; l < MSM_LANES; l++) {
    MsmWorkspace& w = ctx->ws[l];
    DeviceBuf* bufs[] = {&w.msm_digits, &w.msm_sorted, &w.msm_counts, &w.msm_starts, &w.msm_cursor, &w.msm_buckets,
                         &w.msm_tmp, &w.msm_small, &w.msm_scan_tmp, &w.msm_big, &w.msm_pairs, &w.msm_keys};
    for (auto bp : bufs) free_buf(*bp);
    if (l > 0 && w.stream) cudaStreamDestroy(w.stream);
    if (w.done) cudaEventDestroy(w.done);
  }
  cudaStreamDestroy(ctx->stream);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  for (auto ev : ctx->ev_chunk)
    if (ev) cudaEventDestroy(ev);
  delete ctx;
}

const char* b200zk_strerror(int code) {
  switch (code) {
    case B200ZK_OK: return "ok";
    case B200ZK_ERR_NO_DEVICE: return "no usable CUDA device (libb200zk has no CPU fallback)";
    case B200ZK_ERR_CUDA: return "CUDA runtime error";
    case B200ZK_ERR_BAD_ARG: return "bad argument";
    case B200ZK_ERR_OOM: return "out of device memory";
    case B200ZK_ERR_UNSUPPORTED: return "unsupported size";
    case B200ZK_ERR_UNSATISFIED: return "constraint system not satisfied by the solution";
    default: return "unknown error";
  }
}

const char* b200zk_last_cuda_error(const b200zk_ctx* ctx) { return ctx ? ctx->cuda_err : ""; }
void* b200zk_stream(b200zk_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t b200zk_launch_count(const b200zk_ctx* ctx) { return ctx ? ctx->launches : 0; }

int b200zk_sync(b200zk_ctx* ctx) {
  B200ZK_TRY(enter(ctx));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200ZK_OK;
}

// ---- NTT ---------------------------------------------------------------------------------------------------
int b200zk_ntt_dev(b200zk_ctx* ctx, void* a_dev, unsigned log2n, int inverse, int decimation, int coset) {
  B200ZK_TRY(enter(ctx));
  if (!a_dev || log2n > B200ZK_MAX_LOG2N || (decimation != B200ZK_DIF && decimation != B200ZK_DIT))
    return B200ZK_ERR_BAD_ARG;
  return ntt_run(ctx, a_dev, log2n, inverse != 0, decimation, coset != 0);
}

int b200zk_ntt(b200zk_ctx* ctx, void* a_host, unsigned log2n, int inverse, int decimation, int coset) {
  B200ZK_TRY(enter(ctx));
  if (!a_host || log2n > B200ZK_MAX_LOG2N || (decimation != B200ZK_DIF && decimation != B200ZK_DIT))
    return B200ZK_ERR_BAD_ARG;
  const size_t bytes = ((size_t)1 << log2n) * 32;
  B200ZK_TRY(ensure(ctx, ctx->stage, bytes));
  B200ZK_CUDA(ctx, cudaMemcpyAsync(ctx->stage.p, a_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  B200ZK_TRY(ntt_run(ctx, ctx->stage.p, log2n, inverse != 0, decimation, coset != 0));
  B200ZK_CUDA(ctx, cudaMemcpyAsync(a_host, ctx->stage.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200ZK_OK;
}

int b200zk_ntt_dist_half_dev(b200zk_ctx* ctx, const void* src_dev, void* dst_dev, unsigned log2n, unsigned log2g,
                             unsigned rank, unsigned log2c, int half, int inverse, int decimation, int coset) {
  B200ZK_TRY(enter(ctx));
  if (!src_dev || !dst_dev || (half != 0 && half != 1) || (decimation != B200ZK_DIF && decimation != B200ZK_DIT))
    return B200ZK_ERR_BAD_ARG;
  return ntt_dist_run(ctx, src_dev, dst_dev, log2n, log2g, rank, log2c, half, inverse != 0, decimation, coset != 0);
}

int b200zk_ntt_dist_half0_p2p_dev(b200zk_ctx* ctx, void* src_dev, void* const* peer_bufs, unsigned log2n, unsigned log2g,
                                  unsigned rank, unsigned log2c, int inverse, int decimation, int coset) {
  B200ZK_TRY(enter(ctx));
  if (!src_dev || !peer_bufs || (decimation != B200ZK_DIF && decimation != B200ZK_DIT)) return B200ZK_ERR_BAD_ARG;
  // the local destination is unused by the scattering pass; pass src so in-place passes before it work on src
  return ntt_dist_run(ctx, src_dev, src_dev, log2n, log2g, rank, log2c, 0, inverse != 0, decimation, coset != 0, peer_bufs);
}

int b200zk_dev_alloc(b200zk_ctx* ctx, size_t bytes, void** out) {
  B200ZK_TRY(enter(ctx));
  if (!out || !bytes) return B200ZK_ERR_BAD_ARG;
  B200ZK_CUDA(ctx, cudaMalloc(out, bytes));
  return B200ZK_OK;
}

int b200zk_host_alloc(b200zk_ctx* ctx, size_t bytes, void** out) {
  B200ZK_TRY(enter(ctx));
  if (!out || !bytes) return B200ZK_ERR_BAD_ARG;
  B200ZK_CUDA(ctx, cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return B200ZK_OK;
}

int b200zk_host_free(b200zk_ctx* ctx, void* p) {
  B200ZK_TRY(enter(ctx));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  B200ZK_CUDA(ctx, cudaFreeHost(p));
  return B200ZK_OK;
}

int b200zk_dev_free(b200zk_ctx* ctx, void* p) {
  B200ZK_TRY(enter(ctx));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  B200ZK_CUDA(ctx, cudaFree(p));
  return B200ZK_OK;
}

int b200zk_ipc_export(b200zk_ctx* ctx, void* dev_ptr, void* handle_out_64) {
  B200ZK_TRY(enter(ctx));
  if (!dev_ptr || !handle_out_64) return B200ZK_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  B200ZK_CUDA(ctx, cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle_out_64, &h, 64);
  return B200ZK_OK;
}

int b200zk_ipc_import(b200zk_ctx* ctx, const void* handle_64, void** out) {
  B200ZK_TRY(enter(ctx));
  if (!handle_64 || !out) return B200ZK_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, 64);
  B200ZK_CUDA(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return B200ZK_OK;
}

int b200zk_ipc_close(b200zk_ctx* ctx, void* imported) {
  B200ZK_TRY(enter(ctx));
  B200ZK_CUDA(ctx, cudaIpcCloseMemHandle(imported));
  return B200ZK_OK;
}

int b200zk_bit_reverse_dev(b200zk_ctx* ctx, void* a_dev, unsigned log2n) {
  B200ZK_TRY(enter(ctx));
  if (!a_dev || log2n > B200ZK_MAX_LOG2N) return B200ZK_ERR_BAD_ARG;
  return bit_reverse_run(ctx, a_dev, log2n);
}

int b200zk_bit_reverse(b200zk_ctx* ctx, void* a_host, unsigned log2n) {
  B200ZK_TRY(enter(ctx));
  if (!a_host || log2n > B200ZK_MAX_LOG2N) return B200ZK_ERR_BAD_ARG;
  const size_t bytes = ((size_t)1 << log2n) * 32;
  B200ZK_TRY(ensure(ctx, ctx->stage, bytes));
  B200ZK_CUDA(ctx, cudaMemcpyAsync(ctx->stage.p, a_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  B200ZK_TRY(bit_reverse_run(ctx, ctx->stage.p, log2n));
  B200ZK_CUDA(ctx, cudaMemcpyAsync(a_host, ctx->stage.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200ZK_OK;
}

// ---- MSM ---------------------------------------------------------------------------------------------------
int b200zk_bases_upload(b200zk_ctx* ctx, const void* g1_affine_host, size_t n, b200zk_bases** out) {
  B200ZK_TRY(enter(ctx));
  if (!out || (n && !g1_affine_host)) return B200ZK_ERR_BAD_ARG;
  *out = nullptr;
  b200zk_bases* b = new (std::nothrow) b200zk_bases();
  if (!b) return B200ZK_ERR_OOM;
  void* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, n ? n * 64 : 64);
  if (e != cudaSuccess) {
    delete b;
    return set_cuda_error(ctx, e, "cudaMalloc(bases)");
  }
  if (n) {
    e = cudaMemcpyAsync(dev, g1_affine_host, n * 64, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      cudaFree(dev);
      delete b;
      return set_cuda_error(ctx, e, "cudaMemcpy(bases)");
    }
  }
  b->dev = dev;
  b->n = n;
  b->owned = true;
  *out = b;
  return B200ZK_OK;
}

int b200zk_bases_wrap_dev(b200zk_ctx* ctx, const void* g1_affine_dev, size_t n, b200zk_bases** out) {
  B200ZK_TRY(enter(ctx));
  if (!out || (n && !g1_affine_dev)) return B200ZK_ERR_BAD_ARG;
  b200zk_bases* b = new (std::nothrow) b200zk_bases();
  if (!b) return B200ZK_ERR_OOM;
  b->dev = g1_affine_dev;
  b->n = n;
  b->owned = false;
  *out = b;
  return B200ZK_OK;
}

void b200zk_bases_free(b200zk_ctx* ctx, b200zk_bases* bases) {
  if (!bases) return;
  if (bases->owned && bases->dev) {
    if (ctx) {
      cudaSetDevice(ctx->device);
      cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(const_cast<void*>(bases->dev));
  }
  if (bases->table) {
    if (ctx) {
      cudaSetDevice(ctx->device);
      cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(bases->table);
    if (bases->small_table) cudaFree(bases->small_table);
  }
  delete bases;
}

int b200zk_bases_precompute(b200zk_ctx* ctx, b200zk_bases* bases, int c) {
  B200ZK_TRY(enter(ctx));
  if (!bases || (c != 0 && (c < 10 || c > 23))) return B200ZK_ERR_BAD_ARG;
  return msm_precompute_run(ctx, bases, c);
}

size_t b200zk_bases_len(const b200zk_bases* bases) { return bases ? bases->n : 0; }

int b200zk_msm_g1_dev(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first_base, const void* scalars_dev,
                      size_t n, void* out_dev, int out_kind) {
  B200ZK_TRY(enter(ctx));
  if (out_kind != 0 && out_kind != 1) return B200ZK_ERR_BAD_ARG;
  return msm_run(ctx, bases, first_base, scalars_dev, n, out_dev, out_kind);
}

// host scalars -> result left on the device (stage + 0: affine, or extended-Jacobian partial), on the context stream
static int msm_host_scalars(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first_base, const void* scalars_host, size_t n,
                            int out_kind, char** result_dev) {
  const size_t head = 256;  // the result (affine or extended-Jacobian)
  const size_t bytes = n * 32 + head;
  B200ZK_TRY(ensure(ctx, ctx->stage, bytes));
  char* stage = (char*)ctx->stage.p;
  char* sc = stage + head;
  *result_dev = stage;
  // Large inputs are split by point range so that the PCIe copy of chunk i+1 runs under the MSM of chunk i (the sum of
  // the partial MSMs is the MSM: the canonical affine result does not depend on the split).  Only the copy of the FIRST
  // chunk is exposed, so the automatic split is uneven — 1/8, 3/8, 1/2 of the points from 2^23 points (a short first
  // copy, and every later copy still shorter than the MSM it hides under), 1/4, 3/4 from 2^20 points (the shards of a
  // multi-GPU MSM: one more front end costs less than three quarters of the copy).
  unsigned chunks = ctx->msm_host_chunks > 0 ? (unsigned)ctx->msm_host_chunks
                                             : (n >= ((size_t)1 << 23) ? 3u : (n >= ((size_t)1 << 20) ? 2u : 1u));
  if (chunks > 8) chunks = 8;
  if (n < 4096 * (size_t)chunks) chunks = 1;
  if (chunks == 1) {
    if (n) B200ZK_CUDA(ctx, cudaMemcpyAsync(sc, scalars_host, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    return msm_run(ctx, bases, first_base, sc, n, stage, out_kind);
  }
  size_t bound[9];
  if (ctx->msm_host_chunks == 0 && chunks == 3) {
    bound[0] = 0; bound[1] = n / 8; bound[2] = n / 2; bound[3] = n;
  } else if (ctx->msm_host_chunks == 0) {
    bound[0] = 0; bound[1] = n / 4; bound[2] = n;
  } else {
    const size_t per = (n + chunks - 1) / chunks;
    for (unsigned i = 0; i <= chunks; i++) bound[i] = (size_t)i * per < n ? (size_t)i * per : n;
  }
  // the copy stream must not overtake earlier work on the compute stream that still reads the staging buffer
  B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
  auto copy_chunk = [&](unsigned i) -> int {
    const size_t first = bound[i], cnt = bound[i + 1] - first;
    if (cnt)
      B200ZK_CUDA(ctx, cudaMemcpyAsync(sc + first * 32, (const char*)scalars_host + first * 32, cnt * 32, cudaMemcpyHostToDevice,
                                       ctx->side));
    B200ZK_CUDA(ctx, cudaEventRecord(ctx->ev_chunk[i], ctx->side));
    return B200ZK_OK;
  };
  B200ZK_TRY(copy_chunk(0));
  // every chunk is scattered and accumulated into the SAME buckets (shape of the whole MSM); the bucket reduction and
  // the normalisation run once, after the last chunk
  for (unsigned i = 0; i < chunks; i++) {
    if (i + 1 < chunks) B200ZK_TRY(copy_chunk(i + 1));
    const size_t first = bound[i], cnt = bound[i + 1] - first;
    B200ZK_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[i], 0));
    const int part = i == 0 ? MSM_PART_FIRST : (i + 1 == chunks ? MSM_PART_LAST : MSM_PART_MORE);
    B200ZK_TRY(msm_run(ctx, bases, first_base + first, sc + first * 32, cnt, stage, out_kind, 0, part, n));
  }
  return B200ZK_OK;
}

int b200zk_msm_g1(b200zk_ctx* ctx, const b200zk_bases* bases, const void* scalars_host, size_t n,
                  void* out_affine_host) {
  B200ZK_TRY(enter(ctx));
  if (!bases || !out_affine_host || (n && !scalars_host) || n > bases->n) return B200ZK_ERR_BAD_ARG;
  char* result = nullptr;
  B200ZK_TRY(msm_host_scalars(ctx, bases, 0, scalars_host, n, 0, &result));
  B200ZK_CUDA(ctx, cudaMemcpyAsync(out_affine_host, result, 64, cudaMemcpyDeviceToHost, ctx->stream));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200ZK_OK;
}

int b200zk_msm_g1_shard(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first_base, const void* scalars_host, size_t n,
                        void* out_partial_dev) {
  B200ZK_TRY(enter(ctx));
  if (!bases || !out_partial_dev || (n && !scalars_host) || first_base > bases->n || n > bases->n - first_base)
    return B200ZK_ERR_BAD_ARG;
  char* result = nullptr;
  B200ZK_TRY(msm_host_scalars(ctx, bases, first_base, scalars_host, n, 1, &result));
  B200ZK_CUDA(ctx, cudaMemcpyAsync(out_partial_dev, result, 128, cudaMemcpyDeviceToDevice, ctx->stream));
  return B200ZK_OK;
}

int b200zk_plonk_set_commit_lanes(b200zk_ctx* ctx, int lanes) {
  if (!ctx || (lanes != 1 && lanes != MSM_LANES)) return B200ZK_ERR_BAD_ARG;
  ctx->msm_single_lane = lanes == 1;
  return B200ZK_OK;
}

int b200zk_msm_set_reduce_chunk(b200zk_ctx* ctx, int chunk_log) {
  if (!ctx || (chunk_log != 0 && chunk_log != 3 && chunk_log != 4 && chunk_log != 5)) return B200ZK_ERR_BAD_ARG;
  ctx->msm_chunk_log = chunk_log;
  return B200ZK_OK;
}

int b200zk_msm_set_scatter_passes(b200zk_ctx* ctx, int passes) {
  if (!ctx || passes < 0 || passes > 256) return B200ZK_ERR_BAD_ARG;
  ctx->msm_scatter_passes = passes;
  return B200ZK_OK;
}

int b200zk_msm_set_pair_rounds(b200zk_ctx* ctx, int rounds) {
  if (!ctx || rounds < -1 || rounds > 6) return B200ZK_ERR_BAD_ARG;
  ctx->msm_pair_rounds = rounds;
  return B200ZK_OK;
}

int b200zk_msm_set_small_path(b200zk_ctx* ctx, int on) {
  if (!ctx) return B200ZK_ERR_BAD_ARG;
  ctx->msm_no_tiny = on ? 0 : 1;
  return B200ZK_OK;
}

int b200zk_msm_set_host_chunks(b200zk_ctx* ctx, int chunks) {
  if (!ctx || chunks < 0 || chunks > 8) return B200ZK_ERR_BAD_ARG;
  ctx->msm_host_chunks = chunks;
  return B200ZK_OK;
}

int b200zk_g1_sum_dev(b200zk_ctx* ctx, const void* partials_dev, size_t count, void* out_affine_dev) {
  B200ZK_TRY(enter(ctx));
  return g1_sum_run(ctx, partials_dev, count, out_affine_dev, 0);
}

int b200zk_srs_generate(b200zk_ctx* ctx, const void* alpha_host, size_t first, size_t n, b200zk_bases** out) {
  B200ZK_TRY(enter(ctx));
  if (!alpha_host || !out) return B200ZK_ERR_BAD_ARG;
  *out = nullptr;
  b200zk_bases* b = new (std::nothrow) b200zk_bases();
  if (!b) return B200ZK_ERR_OOM;
  void* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, (n ? n : 1) * 64 + 32);
  if (e != cudaSuccess) {
    delete b;
    return set_cuda_error(ctx, e, "cudaMalloc(srs)");
  }
  char* alpha_dev = (char*)dev + (n ? n : 1) * 64;
  e = cudaMemcpyAsync(alpha_dev, alpha_host, 32, cudaMemcpyHostToDevice, ctx->stream);
  int rc = e == cudaSuccess ? srs_generate_run(ctx, alpha_dev, first, n, dev) : set_cuda_error(ctx, e, "cudaMemcpy(alpha)");
  if (rc != B200ZK_OK) {
    cudaFree(dev);
    delete b;
    return rc;
  }
  b->dev = dev;
  b->n = n;
  b->owned = true;
  *out = b;
  return B200ZK_OK;
}

int b200zk_bases_upload_compressed(b200zk_ctx* ctx, const void* compressed_host, size_t n, b200zk_bases** out) {
  B200ZK_TRY(enter(ctx));
  if (!out || (n && !compressed_host)) return B200ZK_ERR_BAD_ARG;
  *out = nullptr;
  b200zk_bases* b = new (std::nothrow) b200zk_bases();
  if (!b) return B200ZK_ERR_OOM;
  void* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, (n ? n : 1) * 64);
  if (e != cudaSuccess) {
    delete b;
    return set_cuda_error(ctx, e, "cudaMalloc(bases)");
  }
  unsigned bad = 0;
  int rc = n ? srs_decompress_run(ctx, compressed_host, n, dev, &bad) : B200ZK_OK;
  if (rc == B200ZK_OK && bad != 0) rc = B200ZK_ERR_BAD_ARG;  // a point is not on the curve / malformed encoding
  if (rc != B200ZK_OK) {
    cudaFree(dev);
    delete b;
    return rc;
  }
  b->dev = dev;
  b->n = n;
  b->owned = true;
  *out = b;
  return B200ZK_OK;
}

int b200zk_bases_download_compressed(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first, size_t n, void* out_host) {
  B200ZK_TRY(enter(ctx));
  if (!bases || !out_host || first > bases->n || n > bases->n - first) return B200ZK_ERR_BAD_ARG;
  if (n == 0) return B200ZK_OK;
  return srs_compress_run(ctx, (const char*)bases->dev + first * 64, n, out_host);
}

int b200zk_bases_download(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first, size_t n, void* out_host) {
  B200ZK_TRY(enter(ctx));
  if (!bases || !out_host || first > bases->n || n > bases->n - first) return B200ZK_ERR_BAD_ARG;
  B200ZK_CUDA(ctx, cudaMemcpyAsync(out_host, (const char*)bases->dev + first * 64, n * 64, cudaMemcpyDeviceToHost,
                                   ctx->stream));
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return B200ZK_OK;
}

int b200zk_profile_enable(b200zk_ctx* ctx, int on) {
  B200ZK_TRY(enter(ctx));
  ctx->profiling = on != 0;
  return B200ZK_OK;
}

int b200zk_profile_read(b200zk_ctx* ctx, double* ms_per_phase, uint64_t* count_per_phase, int nphases) {
  B200ZK_TRY(enter(ctx));
  if (!ms_per_phase || !count_per_phase || nphases < PH_COUNT) return B200ZK_ERR_BAD_ARG;
  B200ZK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < nphases; i++) {
    ms_per_phase[i] = 0;
    count_per_phase[i] = 0;
  }
  for (auto& r : ctx->records) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      ms_per_phase[r.phase] += ms;
      count_per_phase[r.phase]++;
    } else {
      cudaGetLastError();
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  ctx->records.clear();
  return B200ZK_OK;
}

int b200zk_microbench(b200zk_ctx* ctx, int which, double* out_ops_per_s) {
  B200ZK_TRY(enter(ctx));
  return microbench_run(ctx, which, out_ops_per_s);
}

int b200zk_ntt_set_radix2(b200zk_ctx* ctx, int on) {
  if (!ctx) return B200ZK_ERR_BAD_ARG;
  ctx->ntt_radix2 = on != 0;
  return B200ZK_OK;
}


int b200zk_msm_windows(const b200zk_ctx* ctx, const b200zk_bases* bases, size_t n) {
  if (!ctx || !bases || n == 0 || n > bases->n) return B200ZK_ERR_BAD_ARG;
  return (int)msm_window_count(ctx, bases, n);
}

int b200zk_msm_set_window(b200zk_ctx* ctx, int c) {
  if (!ctx || (c != 0 && (c < 6 || c > 16))) return B200ZK_ERR_BAD_ARG;
  ctx->forced_window = c;
  return B200ZK_OK;
}

}  // extern "C"
