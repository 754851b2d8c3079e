/* b200zk — C ABI of the sm_100a PLONK hot path (BN254 G1 MSM + fr NTT) for noir_backend_using_gnark.
 *
 * This is the inner drop-in boundary (SURVEY.md §8b "B-inner"): the functions a cgo shim inside
 * gnark_backend_ffi binds in place of gnark-crypto's CPU arithmetic.  The outer boundary — the four cgo exports
 * PlonkProveWithPK / PlonkVerifyWithVK / PlonkPreprocess / PlonkVerifyWithMeta at
 * /root/reference/gnark_backend_ffi/main.go:24,44,58,39 — stays byte-for-byte as the reference defines it.
 *
 * Conventions (identical to gnark-crypto v0.9.1 in-memory layouts, so Go slices can be passed without copies):
 *   fr.Element / fp.Element : 32 bytes = 4 x uint64 little-endian limbs, Montgomery form (R = 2^256), fully reduced
 *   G1Affine                : 64 bytes = X || Y, point at infinity = 64 zero bytes
 *   decimation              : 0 = DIF (natural in -> bit-reversed out), 1 = DIT (bit-reversed in -> natural out)
 *
 * Every function returns 0 on success or a negative b200zk_error; nothing aborts or throws across the boundary
 * (the Go shim turns a non-zero code into log.Fatal to match /root/reference/gnark_backend_ffi/backend/plonk/plonk.go:69).
 * A context is bound to ONE device and may be used from one thread at a time; every entry point re-selects the
 * context's device, so it can be called from any OS thread (Go moves goroutines between threads).
 * There is no CPU fallback: without a CUDA device b200zk_init fails with B200ZK_ERR_NO_DEVICE.
 */
#ifndef B200ZK_H
#define B200ZK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200zk_ctx b200zk_ctx;
typedef struct b200zk_bases b200zk_bases;

typedef enum b200zk_error {
  B200ZK_OK = 0,
  B200ZK_ERR_NO_DEVICE = -1,   /* no CUDA device / device index out of range */
  B200ZK_ERR_CUDA = -2,        /* a CUDA runtime call failed; see b200zk_last_cuda_error */
  B200ZK_ERR_BAD_ARG = -3,     /* null pointer, log2n > 28, n > bases, ... */
  B200ZK_ERR_OOM = -4,         /* device allocation failed */
  B200ZK_ERR_UNSUPPORTED = -5,
  B200ZK_ERR_UNSATISFIED = -6  /* b200zk_plonk_prove: the solution violates a constraint (what spr.Solve reports) */
} b200zk_error;

enum { B200ZK_DIF = 0, B200ZK_DIT = 1 };
enum { B200ZK_MAX_LOG2N = 28 }; /* 2-adicity of BN254 fr */

/* ---- lifecycle ------------------------------------------------------------------------------------------ */
int b200zk_device_count(void);
int b200zk_init(int device, b200zk_ctx** out);
void b200zk_destroy(b200zk_ctx* ctx);
const char* b200zk_strerror(int code);
const char* b200zk_last_cuda_error(const b200zk_ctx* ctx);
/* CUDA stream (cudaStream_t) all work of this context is enqueued on; *_dev calls are asynchronous on it. */
void* b200zk_stream(b200zk_ctx* ctx);
int b200zk_sync(b200zk_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t b200zk_launch_count(const b200zk_ctx* ctx);

/* ---- fr NTT ---------------------------------------------------------------------------------------------
 * b200zk_ntt == (*fft.Domain).FFT(a, decimation, coset) when inverse == 0 and
 *               (*fft.Domain).FFTInverse(a, decimation, coset) when inverse != 0,
 * for domain = fft.NewDomain(1 << log2n) of gnark-crypto v0.9.1 ecc/bn254/fr/fft (Generator = g^(2^(28-log2n)),
 * FrMultiplicativeGen = 5), reached in the reference through plonk.Prove / plonk.Setup
 * (/root/reference/gnark_backend_ffi/backend/plonk/plonk.go:67, :21).  In place; `a` holds 2^log2n fr.Element.
 * b200zk_ntt takes a HOST pointer (copies in and out); b200zk_ntt_dev takes a DEVICE pointer on the context's
 * device and only enqueues work on the context stream. */
int b200zk_ntt(b200zk_ctx* ctx, void* a_host, unsigned log2n, int inverse, int decimation, int coset);
int b200zk_ntt_dev(b200zk_ctx* ctx, void* a_dev, unsigned log2n, int inverse, int decimation, int coset);
/* One half of the multi-GPU four-step transform of 2^log2n points over g = 2^log2g ranks (NTTs >= 2^24; the
 * caller performs the all-to-all between the halves, see noir_backend_using_gnark_b200/dist_ntt.py).  The logical
 * vector is viewed as R x C (C = 2^log2c).  Shard layouts (2^(log2n-log2g) elements each):
 *   column-block  X[r][c_lo]         logical index r*C + rank*(C/g) + c_lo      (DIF input, DIT output)
 *   row-block     Y[r_lo][c]         logical index (rank*(R/g) + r_lo)*C + c    (DIF output, DIT input)
 *   exchange      Z[peer][r_lo][c_lo]  what all_to_all_single delivers / expects on the row-block side
 * DIF: half 0 in place on X (src == dst); all_to_all(Z <- X); half 1 reads Z (src), writes Y (dst).
 * DIT: half 0 works in place on Y (src) and writes Z (dst); all_to_all(X <- Z); half 1 in place on X.
 * Same inverse / coset semantics as b200zk_ntt (scalings are applied in the half that owns the first / last stage). */
int b200zk_ntt_dist_half_dev(b200zk_ctx* ctx, const void* src_dev, void* dst_dev, unsigned log2n, unsigned log2g,
                             unsigned rank, unsigned log2c, int half, int inverse, int decimation, int coset);
/* Fused variant of half 0: the last pass stores every element straight into the destination rank's exchange buffer
 * through peer-mapped pointers (NVLink stores, tile by tile, overlapping the butterflies) instead of writing locally and
 * calling an all-to-all.  peer_bufs[k] = rank k's exchange buffer of 2^(log2n-log2g) elements as seen from THIS process
 * (own buffer for k == rank, b200zk_ipc_import'ed pointers otherwise); g <= 8.  After it (and a cross-rank barrier)
 * every rank's buffer holds what all_to_all_single would have delivered: Z for DIF, X for DIT.  src is clobbered. */
int b200zk_ntt_dist_half0_p2p_dev(b200zk_ctx* ctx, void* src_dev, void* const* peer_bufs, unsigned log2n, unsigned log2g,
                                  unsigned rank, unsigned log2c, int inverse, int decimation, int coset);
/* device buffers that can be shared with the other ranks' processes (CUDA IPC) */
int b200zk_dev_alloc(b200zk_ctx* ctx, size_t bytes, void** out);
int b200zk_dev_free(b200zk_ctx* ctx, void* p);
int b200zk_ipc_export(b200zk_ctx* ctx, void* dev_ptr, void* handle_out_64);
int b200zk_ipc_import(b200zk_ctx* ctx, const void* handle_64, void** out);
int b200zk_ipc_close(b200zk_ctx* ctx, void* imported);
/* page-locked host buffers for callers without a CUDA runtime of their own (the Go / C++ host side): host arguments
 * of the *_host entry points living in such a buffer are copied by DMA at PCIe speed instead of being staged */
int b200zk_host_alloc(b200zk_ctx* ctx, size_t bytes, void** out);
int b200zk_host_free(b200zk_ctx* ctx, void* p);
/* tests / tuning: force the plain radix-2 pass kernel instead of the radix-4 register kernel (same results) */
int b200zk_ntt_set_radix2(b200zk_ctx* ctx, int on);
/* fft.BitReverse(a): in-place index bit-reversal permutation. */
int b200zk_bit_reverse(b200zk_ctx* ctx, void* a_host, unsigned log2n);
int b200zk_bit_reverse_dev(b200zk_ctx* ctx, void* a_dev, unsigned log2n);

/* ---- G1 MSM ---------------------------------------------------------------------------------------------
 * Bases are the KZG SRS G1 powers kzg.SRS.G1 (loaded at /root/reference/gnark_backend_ffi/backend/common.go:86-105,
 * attached by InitKZG at backend/plonk/plonk.go:62); they are static across calls, so they are uploaded once.
 * b200zk_msm_g1 == (*G1Affine).MultiExp(points[:n], scalars[:n], cfg) of gnark-crypto v0.9.1 ecc/bn254/multiexp.go,
 * as called by kzg.Commit: scalars are Montgomery-form fr.Element; result is the canonical affine point. */
int b200zk_bases_upload(b200zk_ctx* ctx, const void* g1_affine_host, size_t n, b200zk_bases** out);
/* wrap n points already resident on the context's device (no copy; caller keeps them alive) */
int b200zk_bases_wrap_dev(b200zk_ctx* ctx, const void* g1_affine_dev, size_t n, b200zk_bases** out);
/* kzg.NewSRS(size, alpha).G1[first .. first+n) computed on the device: bases[i] = alpha^(first+i) * G
 * (/root/reference/gnark_backend_ffi/backend/common.go:137, main.go:176).  alpha: one Montgomery-form fr.Element;
 * first > 0 produces the point-range shard of one GPU. */
int b200zk_srs_generate(b200zk_ctx* ctx, const void* alpha_host, size_t first, size_t n, b200zk_bases** out);
/* Bases from / to gnark's COMPRESSED G1 encoding (G1Affine.Bytes(): 32-byte big-endian X, flag bits 10 / 11 / 01 in the
 * top byte), the element format of kzg.SRS.WriteTo / ReadFrom — i.e. of the srs.hex cache the reference re-reads on
 * every FFI call (/root/reference/gnark_backend_ffi/backend/common.go:86-105, :107-125).  Decompression (one fp square
 * root per point) and compression run on the device; a malformed / off-curve point fails with B200ZK_ERR_BAD_ARG. */
int b200zk_bases_upload_compressed(b200zk_ctx* ctx, const void* compressed_host, size_t n, b200zk_bases** out);
int b200zk_bases_download_compressed(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first, size_t n, void* out_host);
/* copy bases[first .. first+n) back to the host (64 B each), e.g. to serialise a generated SRS */
int b200zk_bases_download(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first, size_t n, void* out_host);
/* One-time precomputation for static bases (the SRS does not change between commitments): stores the window
 * multiples 2^(c*j) * P_i, j < ceil(255/c), next to the bases (W times the memory).  MSMs over these bases then use a
 * single shared bucket set: no per-window reduction, no Horner tail, c ~ log2(n) - 2.
 * c = 0 chooses c from the number of bases; 10 <= c <= 23 otherwise.  Results are unchanged (canonical affine). */
int b200zk_bases_precompute(b200zk_ctx* ctx, b200zk_bases* bases, int c);
void b200zk_bases_free(b200zk_ctx* ctx, b200zk_bases* bases);
size_t b200zk_bases_len(const b200zk_bases* bases);

int b200zk_msm_g1(b200zk_ctx* ctx, const b200zk_bases* bases, const void* scalars_host, size_t n,
                  void* out_affine_host /* 64 B */);
/* One rank's share of a point-range-sharded MultiExp fed from the host: scalars_host[0..n) pair with
 * bases[first_base .. first_base+n); the extended-Jacobian partial (128 B) is left at out_partial_dev for the
 * cross-GPU combination (all-gather + b200zk_g1_sum_dev).  Same chunked copy / compute overlap as b200zk_msm_g1;
 * asynchronous on the context stream once the copies are enqueued. */
int b200zk_msm_g1_shard(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first_base, const void* scalars_host, size_t n,
                        void* out_partial_dev);
/* Device-resident variant.  first_base = index of the base paired with scalars[0] (point-range shard).
 * out_kind 0: canonical affine (64 B); 1: extended-Jacobian partial X,Y,ZZ,ZZZ (128 B) to be combined across
 * GPUs with b200zk_g1_sum_dev.  Asynchronous on the context stream. */
int b200zk_msm_g1_dev(b200zk_ctx* ctx, const b200zk_bases* bases, size_t first_base, const void* scalars_dev,
                      size_t n, void* out_dev, int out_kind);
/* out_affine_dev (64 B) = canonical affine of the sum of `count` extended-Jacobian partials (128 B each). */
int b200zk_g1_sum_dev(b200zk_ctx* ctx, const void* partials_dev, size_t count, void* out_affine_dev);
/* b200zk_msm_g1 (host scalars) splits large inputs by point range so that the host-to-device copy of one chunk runs
 * under the MSM of the previous one; 0 = choose from n (from 2^23 points: three chunks of 1/8, 3/8 and 1/2 of the points — only the first copy is exposed;
 * from 2^20 points: 1/4 and 3/4), 1 = never split, up to 8 equal chunks */
int b200zk_msm_set_host_chunks(b200zk_ctx* ctx, int chunks);
/* tests / tuning: buckets per running-sum chunk of the bucket reduction, as a power of two: 3 (short dependent chains,
 * chosen up to 2^17 buckets where the reduction is latency-bound), 4 (up to 2^19 buckets), 5 (fewer chunk results, chosen
 * above), 0 = automatic (measured: scripts/reduce_chunk_sweep.py) */
int b200zk_msm_set_reduce_chunk(b200zk_ctx* ctx, int chunk_log);
/* tests: with a window table, MSMs of up to 2^17 (point, window) terms skip the bucket pipeline (one thread per term,
 * two launches: the sizes of the reference's own test circuits); 0 forces the bucket pipeline there too, 1 = default */
int b200zk_msm_set_small_path(b200zk_ctx* ctx, int on);
/* The counting sort's scatter places the (point, window) entries in bucket-range passes so that the partially written
 * sectors of one pass stay in L2 until they are complete (random 4-byte stores into an array far larger than L2 otherwise
 * become 32-byte read-modify-writes in DRAM): 0 = number of passes from the bucket count (default; 7 at 2^21 buckets),
 * 1 = single pass, k <= 256 = forced (tests / tuning).  The result does not depend on it. */
int b200zk_msm_set_scatter_passes(b200zk_ctx* ctx, int passes);
/* Batched-affine pair rounds of the bucket accumulation: the counting sort pads every bucket's run to a multiple of
 * 2^rounds entries and `rounds` passes of out[o] = in[2o] + in[2o+1] (affine additions, one inversion per lane per
 * batch: 5M + 1S per addition instead of 8M + 2S) pre-sum the runs before the extended-Jacobian walk.
 * -1 = choose from the size (default), 0 = off, 1..6 = forced (tests / tuning).  The result does not depend on it. */
int b200zk_msm_set_pair_rounds(b200zk_ctx* ctx, int rounds);
/* force the Pippenger window size (0 = choose from n); for tests and tuning */
int b200zk_msm_set_window(b200zk_ctx* ctx, int c);
/* measurement: the number of windows (= bucket additions per point) an MSM of n points of these bases will use
 * (12 with the window table at 2^24 points, ceil(255/c) classic windows otherwise); < 0 = error code */
int b200zk_msm_windows(const b200zk_ctx* ctx, const b200zk_bases* bases, size_t n);


/* ---- device-resident PLONK prover ----------------------------------------------------------------------
 * The orchestration of gnark v0.8.0 plonk.Setup / plonk.Prove (called at
 * /root/reference/gnark_backend_ffi/backend/plonk/plonk.go:21 and :67) with all polynomials resident in HBM.
 * The host (Go shim / tests) keeps what the reference's glue computes on the CPU: the SparseR1CS rows and the wire
 * permutation (O(n), sequential), and solving the witness.
 *
 * b200zk_plonk_setup: n = 2^log2n rows (placeholders for the nb_public public inputs first, then the constraints,
 *   zero padding), big domain 2^log2n_big (gnark: 4n, or 8n when nbConstraints+nbPublic < 6).
 *   ql..qk: selector columns in LAGRANGE form (n Montgomery fr each; ql = -1 on placeholder rows, qk = constraint
 *   constants with zeros on placeholder rows).  permutation: gnark's pk.Permutation (3n int64).  lro: wire id of
 *   the L | R | O column at every row (3n uint32; padding rows and the R,O of placeholder rows use wire 0).
 *   Computes the canonical forms, S1..S3, the 8 verifying-key commitments and the Lagrange-coset forms.
 * b200zk_plonk_vk: the 8 commitments S[0],S[1],S[2],Ql,Qr,Qm,Qo,Qk as G1Affine (64 B each).
 * b200zk_plonk_prove: solution = value of every wire (nb_wires Montgomery fr; public wires first);
 *   blinding = the 9 fr.SetRandom draws in gnark's order L,L,R,R,O,O,Z,Z,Z (their limbs are the Montgomery form).
 *   proof_out (832 B) = LRO[0..2], Z, H[0..2], BatchedProof.H, ZShiftedOpening.H as G1Affine (64 B each), then
 *   BatchedProof.ClaimedValues[0..6] and ZShiftedOpening.ClaimedValue as fr.Element (32 B each).
 *   Like plonk.Prove (whose spr.Solve fails first), it refuses a solution that violates a constraint: every row is
 *   checked on the device and B200ZK_ERR_UNSATISFIED is returned; b200zk_plonk_unsatisfied_row then gives the first
 *   failing row (row - nb_public = index of the constraint), -1 after a successful prove.
 * Must be called with the context the key was set up on. */
typedef struct b200zk_plonk_pk b200zk_plonk_pk;
/* Commitment hook: when set, every kzg.Commit inside b200zk_plonk_prove calls fn(user, scalars_dev, n, out_dev)
 * instead of the local MSM.  fn must leave the canonical affine commitment (64 B) at out_dev, ordered on the context
 * stream (b200zk_stream).  For hosts that bring their own commitment engine; the built-in multi-GPU prover does not
 * use it (b200zk_plonk_join below) and refuses keys that have one. */
typedef int (*b200zk_commit_fn)(void* user, const void* scalars_dev, size_t n, void* out_affine_dev);
int b200zk_plonk_set_commit_hook(b200zk_plonk_pk* pk, b200zk_commit_fn fn, void* user);
int b200zk_plonk_setup(b200zk_ctx* ctx, const b200zk_bases* bases, unsigned log2n, unsigned log2n_big,
                       unsigned nb_public, unsigned nb_wires, const void* ql_l, const void* qr_l, const void* qm_l,
                       const void* qo_l, const void* qk_l, const int64_t* permutation, const uint32_t* lro,
                       b200zk_plonk_pk** out);
/* plonk.Setup(spr, srs) straight from the SparseR1CS the reference builds
 * (/root/reference/gnark_backend_ffi/backend/plonk/sparse_r1cs.go:98-106): one entry per constraint
 * qL*xa + qR*xb + qO*xc + qM*(xa*xb) + qC == 0 — coefficient columns (Montgomery fr, qm = coeff(M[0])*coeff(M[1]))
 * and the wire ids of L, R, O (public wires are 0..nb_public-1, secret wires follow).  Row layout, padding and gnark's
 * buildPermutation are done inside (host, O(n)). */
int b200zk_plonk_setup_r1cs(b200zk_ctx* ctx, const b200zk_bases* bases, unsigned nb_public, unsigned nb_secret,
                            size_t nb_constraints, const void* ql, const void* qr, const void* qm, const void* qo,
                            const void* qk, const uint32_t* wire_a, const uint32_t* wire_b, const uint32_t* wire_c,
                            b200zk_plonk_pk** out);
void b200zk_plonk_pk_free(b200zk_ctx* ctx, b200zk_plonk_pk* pk);
int b200zk_plonk_vk(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, void* out_8_points);
/* copy a key polynomial to the host (n fr): which = 0..8 -> Ql,Qr,Qm,Qo,CQk (canonical), LQk (Lagrange), S1,S2,S3 */
int b200zk_plonk_pk_poly(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, int which, void* out_host);
int b200zk_plonk_prove(b200zk_ctx* ctx, b200zk_plonk_pk* pk, const void* solution_host, const void* blinding_host,
                       void* proof_out);
long long b200zk_plonk_unsatisfied_row(const b200zk_plonk_pk* pk);
/* ---- the same prover over the 2, 4 or 8 GPUs of one box (one process / context / key per GPU, SPMD) -------
 * Every rank runs b200zk_plonk_setup on its own GPU (same circuit, full SRS), then the ranks map each other's key
 * arena (b200zk_plonk_arena + b200zk_ipc_export / b200zk_ipc_import for other processes; the raw pointers for contexts
 * of one process) and call b200zk_plonk_join with the world's arena pointers as seen from the calling process
 * (arena_ptrs[rank] = the own arena).  From then on b200zk_plonk_prove must be called by ALL ranks: rank 0 passes the
 * solution and the blinding (the other ranks may pass NULL: they fetch both from rank 0 over NVLink), every rank
 * returns the same 832-byte proof.  What is sharded (BASELINE.json north_star): every commitment's MSM by point range
 * (partials combined through peer loads), the five coset transforms of the 4n domain and the inverse one as four-step
 * NTTs whose transpose rides on the last butterfly pass as NVLink peer stores, and the quotient kernel by index range;
 * the n-sized stages are replicated.  There is no collective library on the data path: barriers are epoch counters in
 * the mapped arenas (a rank that does not arrive within 20 s fails the proof instead of hanging the device).
 * Circuits need log2n_big == log2n + 2 and at least 4 * world columns per rank (any circuit worth sharding).
 * b200zk_plonk_leave returns the key to single-GPU proving (call it on every rank before freeing a key). */
int b200zk_plonk_arena(b200zk_ctx* ctx, const b200zk_plonk_pk* pk, void** base_dev, size_t* bytes);
int b200zk_plonk_join(b200zk_ctx* ctx, b200zk_plonk_pk* pk, unsigned rank, unsigned world, void* const* arena_ptrs);
int b200zk_plonk_leave(b200zk_ctx* ctx, b200zk_plonk_pk* pk);
/* tests / tuning: the independent commitments of one prover round (L,R,O; H1,H2,H3; the 8 of Setup) run on 3 MSM lanes
 * of the context (own stream and workspace each) so that the latency-bound phases of one MSM hide under the bucket
 * accumulation of another; 1 = one after the other.  Results do not depend on it. */
int b200zk_plonk_set_commit_lanes(b200zk_ctx* ctx, int lanes);
/* The same prover fed with what PlonkProveWithPK receives (/root/reference/gnark_backend_ffi/main.go:24-37): the hex text of
 * the value vector, 64 characters per element (32 bytes big-endian, regular form), WITHOUT the 8-character count prefix.
 * DeserializeFelts (fr.SetBytes: reduce, Montgomery form) and BuildWitnesses (values -> publics then secrets) run on the
 * device: set_solution_map stores, once per key, which value feeds which wire (wire k <- values[src[k]], nb_wires
 * entries, every entry < nb_values); prove_hex uploads the text, decodes, gathers and proves.  A non-hex character
 * returns B200ZK_ERR_BAD_ARG (hex.DecodeString's error), before anything is committed. */
int b200zk_plonk_set_solution_map(b200zk_ctx* ctx, b200zk_plonk_pk* pk, const uint32_t* src_host, size_t nb_values);
int b200zk_plonk_prove_hex(b200zk_ctx* ctx, b200zk_plonk_pk* pk, const char* values_hex_host, size_t nb_values,
                           const void* blinding_host, void* proof_out);

/* ---- measurement ----------------------------------------------------------------------------------------
 * Integer-pipe microbenchmarks used as roofline denominators (synchronous).  which = 0: IMAD.WIDE.U32 multiply-
 * accumulates per second; 1: fp Montgomery multiplications per second; 2: fr Montgomery multiplications per second;
 * 3: FP64 DFMA per second; 4: IMAD.WIDE while DFMA issues beside it; 5 / 6: fp multiplications per second in the
 * schoolbook CIOS form / with the Karatsuba product (whichever the library was built with is also what 1 and 2 run);
 * 7: fp products per second through the two-product sweep (a*b + c*d under one reduction; self-checked in the kernel). */
int b200zk_microbench(b200zk_ctx* ctx, int which, double* out_ops_per_s);
/* Per-phase device timing with CUDA events on the context stream.  Phases: 0 msm digits+histogram, 1 msm scan,
 * 2 msm scatter, 3 msm bucket accumulation, 4 msm long-run path, 5 msm bucket reduction, 6 msm final, 7 ntt pass.
 * b200zk_profile_read synchronises, returns the summed milliseconds and launch-group counts since the last read
 * (arrays of at least 8 entries) and clears the records. */
int b200zk_profile_enable(b200zk_ctx* ctx, int on);
int b200zk_profile_read(b200zk_ctx* ctx, double* ms_per_phase, uint64_t* count_per_phase, int nphases);

#ifdef __cplusplus
}
#endif
#endif /* B200ZK_H */
