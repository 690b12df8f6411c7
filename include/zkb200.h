/* zkb200 -- C ABI of the B200-native Groth16/BN254 proving backend.
 *
 * This is the drop-in boundary for the reference's proving hot path.  Each entry point names the
 * reference interface it replaces (paths relative to the Zelana-Labs/zelana tree).  The reference's
 * `Groth16Prover: BatchProver` (core/src/sequencer/settlement/prover.rs:160-169, 252-447) would bind
 * these through a Rust `extern "C"` block; tests bind them through ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative zkb_status; nothing throws or aborts
 *     (a failed prove must surface as an error, pipeline.rs:398-425 retries it next tick);
 *   - a zkb_ctx owns one CUDA device + stream and is NOT thread-safe; distinct contexts may be
 *     used concurrently from distinct threads (one per GPU, or several per GPU);
 *   - field elements cross the ABI as 32-byte little-endian CANONICAL integers (what arkworks'
 *     `into_bigint().to_bytes_le()` yields, prover.rs:311-331);
 *   - G1 affine = x || y (64 B), G2 affine = x.c0 || x.c1 || y.c0 || y.c1 (128 B), no flag bits,
 *     infinity = all zero bytes (prover/src/bin/convert_vk.rs:163-191);
 *   - "host" pointers are ordinary CPU memory, "dev" pointers are CUDA device memory on ctx's device.
 *   - there is NO CPU fallback: without a CUDA device zkb_ctx_create fails with ZKB_ERR_NO_DEVICE.
 */
#ifndef ZKB200_H
#define ZKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ZKB_OK = 0,
  ZKB_ERR_NO_DEVICE = -1,
  ZKB_ERR_CUDA = -2,
  ZKB_ERR_INVALID_ARG = -3,
  ZKB_ERR_OOM = -4,
  ZKB_ERR_NOT_CANONICAL = -5, /* a field element >= modulus, or a point not on the curve */
  ZKB_ERR_SHAPE = -6          /* witness / key / matrix sizes disagree */
} zkb_status;

typedef struct zkb_ctx zkb_ctx;
typedef struct zkb_g1_bases zkb_g1_bases; /* device-resident G1 affine points, Montgomery form */
typedef struct zkb_g2_bases zkb_g2_bases;
typedef struct zkb_r1cs zkb_r1cs;         /* device-resident constraint matrices (CSR) */
typedef struct zkb_pk zkb_pk;             /* device-resident Groth16 proving key */

/* ---- library / context --------------------------------------------------------------------- */
const char* zkb_version(void);
int zkb_device_count(void);
int zkb_ctx_create(int device, zkb_ctx** out);
/* Run on an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); NULL = own stream. */
int zkb_ctx_set_stream(zkb_ctx* ctx, void* cuda_stream);
int zkb_ctx_synchronize(zkb_ctx* ctx);
void zkb_ctx_destroy(zkb_ctx* ctx);
const char* zkb_last_error(zkb_ctx* ctx);
/* page-locked host memory: assignments handed to zkb_prove_batch_begin from such a buffer are fetched by the GPU without a
 * staging copy, overlapped with compute (any host memory works; pageable memory is copied synchronously) */
int zkb_host_alloc_pinned(size_t bytes, void** out);
void zkb_host_free_pinned(void* p);
/* Memory-safety check without compute-sanitizer: with ZKB_GUARD=1 in the environment every scratch buffer a context allocates is
 * wrapped in two 4 KiB canaries; this verifies them all (ZKB_OK = intact; ZKB_ERR_INVALID_ARG if nothing was guarded). */
int zkb_debug_check_guards(zkb_ctx* ctx);
/* number of this library's kernels launched on ctx since creation (bench.py's gpu_launches) */
unsigned long long zkb_launch_count(zkb_ctx* ctx);
/* A prove's device part (~140 launches on five streams for the L2 circuit) is captured as a CUDA graph the second time a
 * (key, matrices) pair is proved on a context and replayed afterwards with one launch; on by default, off while profiling
 * (zkb_prof_enable).  zkb_launch_count keeps counting the kernels a replay stands for. */
int zkb_ctx_set_graphs(zkb_ctx* ctx, int on);
int zkb_graph_stats(zkb_ctx* ctx, unsigned long long* captures, unsigned long long* replays);
/* on: host calls that wait for the GPU sleep on a blocking event instead of spinning in cudaStreamSynchronize (default off;
 * zkb_l2_batch lanes switch it on: they outnumber the host cores and a spinning waiter takes the core an assignment needs). */
int zkb_ctx_set_blocking_sync(zkb_ctx* ctx, int on);
/* force the MSM window width c in [2, 23] (0 = automatic) for bases loaded AFTER this call: the width is fixed when a
 * bases handle builds its window tables 2^(c j) P.  Benchmarking/tests only. */
int zkb_ctx_set_msm_window(zkb_ctx* ctx, int c);

/* ---- per-phase device timing (replaces the reference's only timer, `proving_time_ms = start.elapsed()`,
 * core/src/sequencer/settlement/prover.rs:351,418, with CUDA-event spans on ctx's stream) -------------- */
int zkb_prof_phase_count(void);
const char* zkb_prof_phase_name(int phase);   /* "msm_g1_accumulate", "ntt", ... */
int zkb_prof_enable(zkb_ctx* ctx, int on);    /* off by default; when on, every phase is bracketed by two events */
int zkb_prof_reset(zkb_ctx* ctx);             /* synchronises the stream, zeroes the totals */
/* synchronises the stream; total_ms = device time spent in `phase` since the last reset, spans = how many times */
int zkb_prof_read(zkb_ctx* ctx, int phase, double* total_ms, unsigned long long* spans);

/* INT32 multiply-pipe peak of this GPU (the MSM roofline denominator): dependency-free
 * 32x32->64 multiply-accumulates per second.  variant 0 = mad.wide.u32, 1 = mad.lo.cc/madc.hi pairs. */
int zkb_bench_int32_peak(zkb_ctx* ctx, int variant, int iters, double* mul32_per_s, double* elapsed_ms);

/* ---- field arithmetic parity hooks (ark-ff Fp<MontBackend<_,4>>: SURVEY.md 8a row a9) -------- */
/* field: 0 = Fr, 1 = Fq.  op: 0 add, 1 sub, 2 mul, 3 inverse (b ignored; 0 -> 0), 4 neg (b ignored).
 * a, b, out: n x 32 B canonical LE, host memory. */
int zkb_field_op(zkb_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out);

/* ---- curve parity hooks (ark-ec short_weierstrass) ----------------------------------------- */
/* out[i] = scalars[i] * points[i]; group: 1 = G1 (64 B points), 2 = G2 (128 B points); host memory. */
int zkb_scalar_mul(zkb_ctx* ctx, int group, const uint8_t* points, const uint8_t* scalars, size_t n, uint8_t* out);
/* out = sum_i points[i] (plain point sum, exercises add/double/inverse edge cases); host memory. */
int zkb_point_sum(zkb_ctx* ctx, int group, const uint8_t* points, size_t n, uint8_t* out);

/* ---- bases (proving-key query vectors resident in HBM) ---------------------------------------
 * validate != 0 means what arkworks' Validate::Yes means, on every loading path (raw affine or compressed): coordinates
 * canonical, point on the curve and -- G2 only, G1 has cofactor 1 -- in the prime-order subgroup. */
int zkb_g1_bases_load(zkb_ctx* ctx, const uint8_t* affine_host, size_t n, int validate, zkb_g1_bases** out);
int zkb_g2_bases_load(zkb_ctx* ctx, const uint8_t* affine_host, size_t n, int validate, zkb_g2_bases** out);
/* bases[i] = k_i * G (G = (1,2)), k_i = 32 B canonical LE Fr scalars in DEVICE memory: synthetic keys with
 * known discrete logs, generated on the GPU. */
int zkb_g1_bases_generate(zkb_ctx* ctx, const void* k_dev, size_t n, zkb_g1_bases** out);
int zkb_g2_bases_generate(zkb_ctx* ctx, const void* k_dev, size_t n, zkb_g2_bases** out);
/* window width c and window count of the handle's resident tables: nwin x len points (64 / 128 B each) live in HBM */
int zkb_g1_bases_window(const zkb_g1_bases* b, int* c, int* nwin);
int zkb_g2_bases_window(const zkb_g2_bases* b, int* c, int* nwin);
size_t zkb_g1_bases_len(const zkb_g1_bases* b);
size_t zkb_g2_bases_len(const zkb_g2_bases* b);
/* copy bases [offset, offset+n) back as canonical affine bytes (host) */
int zkb_g1_bases_read(zkb_ctx* ctx, const zkb_g1_bases* b, size_t offset, size_t n, uint8_t* out_host);
int zkb_g2_bases_read(zkb_ctx* ctx, const zkb_g2_bases* b, size_t offset, size_t n, uint8_t* out_host);
void zkb_g1_bases_free(zkb_g1_bases* b);
void zkb_g2_bases_free(zkb_g2_bases* b);

/* ---- MSM: ark-ec VariableBaseMSM::msm_bigint -------------------- */
/* sum_{i<n} scalars[i] * bases[offset + i].  Host variant copies scalars H2D and the result D2H. */
int zkb_msm_g1(zkb_ctx* ctx, const zkb_g1_bases* bases, size_t offset, const uint8_t* scalars_host, size_t n,
               uint8_t out_affine_host[64]);
int zkb_msm_g2(zkb_ctx* ctx, const zkb_g2_bases* bases, size_t offset, const uint8_t* scalars_host, size_t n,
               uint8_t out_affine_host[128]);
/* Device variant: asynchronous on ctx's stream.  out_affine_dev: 64 / 128 B canonical (may be NULL);
 * out_partial_dev: 128 / 256 B opaque projective partial sum for zkb_msm_g{1,2}_combine (may be NULL). */
int zkb_msm_g1_dev(zkb_ctx* ctx, const zkb_g1_bases* bases, size_t offset, const void* scalars_dev, size_t n,
                   void* out_affine_dev, void* out_partial_dev);
int zkb_msm_g2_dev(zkb_ctx* ctx, const zkb_g2_bases* bases, size_t offset, const void* scalars_dev, size_t n,
                   void* out_affine_dev, void* out_partial_dev);
/* Host scalars in, projective partial sum out (device memory, asynchronous on ctx's stream): one rank's share of a
 * range-sharded MSM with the sliced upload pipeline of zkb_msm_g1; feed the partials of all ranks to zkb_msm_g{1,2}_combine. */
int zkb_msm_g1_partial(zkb_ctx* ctx, const zkb_g1_bases* bases, size_t offset, const uint8_t* scalars_host, size_t n,
                       void* out_partial_dev);
int zkb_msm_g2_partial(zkb_ctx* ctx, const zkb_g2_bases* bases, size_t offset, const uint8_t* scalars_host, size_t n,
                       void* out_partial_dev);
/* The whole multi-GPU MSM behind one call, no torch / NCCL needed (SURVEY.md 8b `zkb_msm_g1_multi`, 8e): one process, one
 * context per GPU; bases[i] (loaded on ctxs[i]) holds range i of the points, in order, and the n = sum_i len(bases[i]) host
 * scalars are split the same way.  Each GPU runs its partial MSM from its own host thread (sliced upload overlapped with the
 * accumulation); the n_gpus partial sums (128 / 256 B) are gathered through host memory and added on ctxs[0].
 * What a Rust `GpuGroth16Prover` holding one context per device calls (core/src/sequencer/settlement/prover.rs:160-169). */
int zkb_msm_g1_multi(zkb_ctx* const* ctxs, const zkb_g1_bases* const* bases, int n_gpus, const uint8_t* scalars_host, size_t n,
                     uint8_t out_affine_host[64]);
int zkb_msm_g2_multi(zkb_ctx* const* ctxs, const zkb_g2_bases* const* bases, int n_gpus, const uint8_t* scalars_host, size_t n,
                     uint8_t out_affine_host[128]);
/* Multi-GPU combine: k partial sums (gathered from k ranks, device memory) -> canonical affine (device). */
int zkb_msm_g1_combine(zkb_ctx* ctx, const void* partials_dev, int k, void* out_affine_dev);
int zkb_msm_g2_combine(zkb_ctx* ctx, const void* partials_dev, int k, void* out_affine_dev);
/* Parity hook for the batched MSM behind zkb_prove_batch: `batch` scalar vectors (vector p at scalars_dev + p * stride * 32 B)
 * against the same bases [offset, offset + n) -> batch canonical affine points in device memory.  group: 1 = G1, 2 = G2
 * (bases: the matching handle type). */
int zkb_debug_msm_batch(zkb_ctx* ctx, int group, const void* bases, size_t offset, const void* scalars_dev, size_t n, size_t stride,
                        int batch, void* out_affine_dev);
/* Parity hook for the comb-table MSM a batched prove uses for small keys: builds, on the bases handle, the table of every
 * multiple a signed c-bit digit can select (d * 2^(c w) * P_i, 2 <= c <= 16: windows * len * 2^(c-1) points of HBM) and runs
 * `batch` MSMs as plain sums of gathered points -> batch canonical affine points (device). */
int zkb_debug_msm_comb(zkb_ctx* ctx, int group, void* bases, size_t offset, const void* scalars_dev, size_t n, size_t stride,
                       int batch, int c, void* out_affine_dev);
/* Parity hook for the MSM front end (signed-digit extraction fused with the radix sort): the entries of `batch` scalar vectors
 * sorted by key = vector * 2^(c-1) + |digit| - 1, value = table index | sign << 31.  out_keys_dev / out_vals_dev: device
 * buffers of windows * n * batch u32 each; out_count_dev: one u32 = number of entries (zero digits produce none). */
int zkb_debug_msm_entries(zkb_ctx* ctx, const zkb_g1_bases* bases, size_t offset, const void* scalars_dev, size_t n, size_t stride,
                          int batch, void* out_keys_dev, void* out_vals_dev, void* out_count_dev);
#define ZKB_G1_PARTIAL_BYTES 128
#define ZKB_G2_PARTIAL_BYTES 256

/* ---- NTT: ark-poly Radix2EvaluationDomain ---------------------------- */
/* direction: 0 forward (fft_in_place), 1 inverse (ifft_in_place, 1/n folded in).
 * coset: 0 plain, 1 = domain.get_coset(Fr::GENERATOR = 5).  Natural order in and out; n = 2^log_n.
 * Data is n x 32 B; canonical stays canonical.  in == out is allowed. */
int zkb_ntt(zkb_ctx* ctx, const uint8_t* in_host, uint8_t* out_host, int log_n, int direction, int coset);
int zkb_ntt_dev(zkb_ctx* ctx, const void* in_dev, void* out_dev, int log_n, int direction, int coset);

/* ---- MiMC-7 over Fr and the depth-32 account Merkle tree of the forge stack ----------------------------
 * The hash: forge/circuits/zelana_lib/src/poseidon.nr:15-94 (c_i = (i+1)^3 + (i+1), 91 rounds of x -> (x + c_i)^7, sponge over
 * [arity, v_1 .. v_arity]) = core/src/sequencer/storage/account_tree.rs:48-125, where the sequencer evaluates it with BigUint
 * for every tree node it touches.  Field elements: 32 B little-endian canonical, like everywhere in this header (the reference
 * stores tree nodes big-endian: account_tree.rs:187-203 -- the host mirror zelana_b200/account_tree.py converts).
 * zkb_mimc_hash: out[i] = hash_arity(in[i * arity ..]) for n independent hashes, 1 <= arity <= 6 (hash_2 = a tree node,
 * hash_4 = an account leaf `compute_account_leaf`, account_tree.rs:109-125): one level of a batched tree update is one call.
 * zkb_mimc_merkle_roots: out[i] = root reached from leaves[i] along siblings[i * depth ..] with index bits[i * depth ..]
 * (one byte each, 1 = the running node is the RIGHT child) = AccountMerklePath::compute_root (account_tree.rs:222-237).
 * _dev variants: device buffers, asynchronous on ctx's stream. */
int zkb_mimc_hash(zkb_ctx* ctx, int arity, const uint8_t* in_host, size_t n, uint8_t* out_host);
int zkb_mimc_hash_dev(zkb_ctx* ctx, int arity, const void* in_dev, size_t n, void* out_dev);
int zkb_mimc_merkle_roots(zkb_ctx* ctx, const uint8_t* leaves, const uint8_t* siblings, const uint8_t* bits, size_t n, int depth,
                          uint8_t* out);
int zkb_mimc_merkle_roots_dev(zkb_ctx* ctx, const void* leaves_dev, const void* siblings_dev, const uint8_t* bits_dev, size_t n,
                              int depth, void* out_dev);

/* ---- R1CS matrices + witness map: LibsnarkReduction::witness_map_from_matrices (row a4) ----- */
typedef struct {
  const uint64_t* row_ptr; /* num_constraints + 1 offsets into col/coeff */
  const uint32_t* col;     /* variable index: instance j -> j (0 = constant ONE), witness w -> num_instance + w */
  const uint8_t* coeff;    /* nnz x 32 B canonical LE */
} zkb_csr;

typedef struct {
  uint64_t num_constraints;
  uint64_t num_instance; /* including the constant ONE */
  uint64_t num_witness;
  zkb_csr a, b, c;
} zkb_r1cs_desc;

int zkb_r1cs_load(zkb_ctx* ctx, const zkb_r1cs_desc* desc, zkb_r1cs** out);
void zkb_r1cs_free(zkb_r1cs* m);
/* log2 of the QAP domain: next_pow2(num_constraints + num_instance) */
int zkb_r1cs_log_domain(const zkb_r1cs* m);
/* num_instance + num_witness: the length (in 32-byte elements) every z_host passed with these matrices must have */
uint64_t zkb_r1cs_num_variables(const zkb_r1cs* m);
uint64_t zkb_r1cs_num_constraints(const zkb_r1cs* m);
/* z = full assignment [1, instance.., witness..], zkb_r1cs_num_variables(m) x 32 B canonical, host (the callee cannot check
 * the length of a bare pointer: size the buffer from that accessor).
 * h_out: domain_size x 32 B canonical coefficients of h(X), host. */
int zkb_witness_map(zkb_ctx* ctx, const zkb_r1cs* m, const uint8_t* z_host, uint8_t* h_out_host);

/* ---- proving key + prove: ark-groth16 create_proof_with_reduction_and_matrices (rows a3, a8) - */
typedef struct {
  const uint8_t* alpha_g1; /* 64 B */
  const uint8_t* beta_g1;  /* 64 B */
  const uint8_t* beta_g2;  /* 128 B */
  const uint8_t* delta_g1; /* 64 B */
  const uint8_t* delta_g2; /* 128 B */
  const uint8_t* a_query;    size_t a_len;    /* G1, num_instance + num_witness */
  const uint8_t* b_g1_query; size_t b_g1_len; /* G1, same */
  const uint8_t* b_g2_query; size_t b_g2_len; /* G2, same */
  const uint8_t* h_query;    size_t h_len;    /* G1, domain_size - 1 */
  const uint8_t* l_query;    size_t l_len;    /* G1, num_witness */
} zkb_pk_desc;

int zkb_pk_load(zkb_ctx* ctx, const zkb_pk_desc* desc, int validate, zkb_pk** out);
/* The key exactly as the reference stores it: `ProvingKey::<Bn254>::serialize_compressed` bytes (prover/src/bin/keygen.rs:100),
 * i.e. what Groth16Prover::from_bytes / from_files reads (prover.rs:263-286).  Points are decompressed (square roots) and
 * validated (on curve; G2 in the prime-order subgroup, as Validate::Yes does) on the GPU.  Errors: ZKB_ERR_SHAPE for a
 * truncated / over-long buffer, ZKB_ERR_NOT_CANONICAL for a bad point (zkb_last_error says which). */
int zkb_pk_load_compressed(zkb_ctx* ctx, const uint8_t* ark_bytes, size_t len, int validate, zkb_pk** out);
void zkb_pk_free(zkb_pk* pk);
/* BENCHMARK ONLY: a key of the given shape whose query points are [k_i] G for k_len >= max(num_vars + 2, h_len) + 4
 * canonical Fr scalars in device memory (no trusted setup; proofs do not verify, timing is that of a real key). */
int zkb_pk_synthetic(zkb_ctx* ctx, size_t num_vars, size_t num_witness, size_t h_len, const void* k_dev, size_t k_len,
                     zkb_pk** out);

/* Proof = (A in G1, B in G2, C in G1), canonical affine, A NOT negated (negation and flag bits are host-side
 * formatting: prover.rs:304-334 / ark-serialize).  r, s: 32 B canonical Fr, the prover's randomness
 * (StdRng::seed_from_u64(batch_id) -> Fr::rand twice, prover.rs:354).  z_host as in zkb_witness_map. */
int zkb_prove(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t r[32],
              const uint8_t s[32], uint8_t out_a[64], uint8_t out_b[128], uint8_t out_c[64]);

/* ---- a batch of proofs of ONE circuit with ONE key (BASELINE.json config 5; the forge coordinator's chunk-per-worker
 * dispatch, forge/crates/prover-coordinator/src/dispatcher.rs:290-330, as one set of kernels): K x
 * create_proof_with_reduction_and_matrices.  z_host: K assignments back to back, K x zkb_r1cs_num_variables(m) x 32 B;
 * rs_host: K x (r || s), 64 B each; out: K x 256 B, proof i = A (64) | B (128) | C (64) canonical affine, A not negated
 * (byte-identical to K calls of zkb_prove).  1 <= K <= 4096.  _begin queues the work and returns; the host buffers must stay
 * valid until _end, which waits and copies the results out.  One batch in flight per context. */
int zkb_prove_batch_begin(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t* rs_host,
                          size_t k);
/* flags = ZKB_BATCH_Z_MONTGOMERY: z_host holds Montgomery limbs (x * 2^256 mod r as 32 B little-endian, < r) instead of
 * canonical values -- the form ark-ff's Fp<MontBackend<_, 4>> keeps in memory, so a host synthesiser
 * hands its assignment over without 6 000 from-Montgomery products per L2-circuit proof; (r, s) stay canonical.  Same proofs. */
#define ZKB_BATCH_Z_MONTGOMERY 1u
int zkb_prove_batch_begin_ex(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t* rs_host,
                             size_t k, unsigned flags);
int zkb_prove_batch_end(zkb_ctx* ctx, size_t k, uint8_t* out);
int zkb_prove_batch(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t* rs_host, size_t k,
                    uint8_t* out);

/* ---- trusted setup: ark-groth16 generate_parameters_with_qap (Groth16::circuit_specific_setup, keygen.rs:87-91) -------- */
typedef struct {
  uint8_t alpha[32], beta[32], gamma[32], delta[32], tau[32]; /* canonical Fr, the toxic waste (caller's RNG) */
  uint8_t g1_generator[64];                                    /* the random generators arkworks draws (G1::rand, G2::rand) */
  uint8_t g2_generator[128];
} zkb_setup_params;

typedef struct { /* caller-allocated host buffers, raw canonical affine (infinity = zeros) */
  uint8_t *alpha_g1, *beta_g1, *delta_g1; /* 64 B each */
  uint8_t *beta_g2, *gamma_g2, *delta_g2; /* 128 B each */
  uint8_t* gamma_abc_g1;                  /* num_instance x 64 */
  uint8_t* a_query;                       /* (num_instance + num_witness) x 64 */
  uint8_t* b_g1_query;                    /* same */
  uint8_t* b_g2_query;                    /* same x 128 */
  uint8_t* h_query;                       /* (domain_size - 1) x 64 */
  uint8_t* l_query;                       /* num_witness x 64 */
} zkb_setup_out;

/* Errors: ZKB_ERR_INVALID_ARG if tau is a domain element or gamma / delta is zero; ZKB_ERR_NOT_CANONICAL for values >= r. */
int zkb_setup(zkb_ctx* ctx, const zkb_r1cs_desc* desc, const zkb_setup_params* params, const zkb_setup_out* out);

/* ---- one proof over several GPUs (MSMs sharded by contiguous range of the key's query vectors) ----------
 * Rank `shard` of `world` loads its range of the key (zkb_pk_load_shard), runs zkb_prove_partial (witness map replicated,
 * five MSMs over its range) and contributes ZKB_PROVE_PARTIAL_BYTES; the records are all-gathered (NCCL) and any rank
 * finishes the proof with zkb_prove_combine.  world = 1 reproduces zkb_prove. */
#define ZKB_PROVE_PARTIAL_BYTES 768 /* A, B1, L, H: 4 x 128 B ; B2: 256 B (opaque projective partial sums) */
int zkb_pk_load_shard(zkb_ctx* ctx, const zkb_pk_desc* desc, int validate, int shard, int world, zkb_pk** out);
int zkb_pk_synthetic_shard(zkb_ctx* ctx, size_t num_vars, size_t num_witness, size_t h_len, const void* k_dev, size_t k_len,
                           int shard, int world, zkb_pk** out); /* BENCHMARK ONLY, as zkb_pk_synthetic */
int zkb_prove_partial(zkb_ctx* ctx, const zkb_pk* pk_shard, const zkb_r1cs* m, const uint8_t* z_host, const uint8_t r[32],
                      const uint8_t s[32], void* out_partial_dev);
int zkb_prove_combine(zkb_ctx* ctx, const void* partials_dev, int world, const uint8_t r[32], const uint8_t s[32],
                      uint8_t out_a[64], uint8_t out_b[128], uint8_t out_c[64]);

/* The same behind one call for a single-process host (no torch / NCCL): ctxs[i] is on GPU i, pk_shards[i] =
 * zkb_pk_load_shard(ctxs[i], desc, validate, i, n_gpus), ms[i] = the matrices loaded on ctxs[i].  Every GPU proves its share from
 * its own host thread; the partial records are gathered through host memory; ctxs[0] finishes. */
int zkb_prove_multi(zkb_ctx* const* ctxs, const zkb_pk* const* pk_shards, const zkb_r1cs* const* ms, int n_gpus,
                    const uint8_t* z_host, const uint8_t r[32], const uint8_t s[32], uint8_t out_a[64], uint8_t out_b[128],
                    uint8_t out_c[64]);

/* ---- the L2 batch circuit on the host: prover/src/l2_circuit.rs (8f.3) ----------------------
 * What `Groth16Prover::prove` (core/src/sequencer/settlement/prover.rs:350-425) does around the arkworks call, natively:
 * build `L2BlockCircuit`, synthesise its constraints (l2_circuit.rs:179-505 over ark-r1cs-std / ark-crypto-primitives 0.5.0
 * gadgets), seed StdRng with the batch id, draw (r, s), prove on the GPU, format the 256-byte Solana proof.
 * Host-only functions (no zkb_ctx) report through zkb_l2_last_error() (thread-local). */
typedef struct {              /* = BatchPublicInputs, prover.rs:48-63; roots are `Fr::from_le_bytes_mod_order` inputs */
  uint8_t pre_state_root[32];
  uint8_t post_state_root[32];
  uint8_t pre_shielded_root[32];
  uint8_t post_shielded_root[32];
  uint8_t withdrawal_root[32];
  uint8_t batch_hash[32];
  uint64_t batch_id;
} zkb_l2_public_inputs;

typedef struct {              /* L2BlockCircuit's private witness, l2_circuit.rs:110-119 (prover.rs:357-404 fills it) */
  const uint8_t* account_pks;       /* n_accounts x 32: initial_accounts keys (any order; folded in BTreeMap order) */
  const uint64_t* account_balances; /* n_accounts */
  size_t n_accounts;
  const uint8_t* tx_senders;        /* n_txs x 32 */
  const uint8_t* tx_recipients;     /* n_txs x 32 */
  const uint64_t* tx_amounts;       /* n_txs */
  size_t n_txs;
  const uint8_t* commitments;       /* n_commitments x 32 (shielded_commitments) */
  size_t n_commitments;
  const uint8_t* wd_recipients;     /* n_withdrawals x 32 */
  const uint64_t* wd_amounts;       /* n_withdrawals */
  size_t n_withdrawals;
} zkb_l2_witness;

typedef struct zkb_l2_circuit zkb_l2_circuit; /* host: constraint matrices of ONE circuit shape (what a key is made for);
                                                 immutable once created: any number of threads may share one */

const char* zkb_l2_last_error(void);
/* Structure pass = `generate_constraints` under SynthesisMode::Setup + `cs.finalize(); cs.to_matrices()`
 * (OptimizationGoal::Constraints, as ark-groth16 sets).  Only the SHAPE of `shape` matters: the counts and which account
 * each transfer debits / credits (keygen uses L2BlockCircuit::dummy(), l2_circuit.rs:147-170).
 * ZKB_ERR_SHAPE: a transfer's sender is not an initial account (SynthesisError::AssignmentMissing, l2_circuit.rs:266). */
int zkb_l2_circuit_create(const zkb_l2_witness* shape, zkb_l2_circuit** out);
void zkb_l2_circuit_free(zkb_l2_circuit* c);
/* Host CSR views (valid until zkb_l2_circuit_free) for zkb_r1cs_load / zkb_setup. */
int zkb_l2_circuit_desc(const zkb_l2_circuit* c, zkb_r1cs_desc* out);
/* Assignment pass = `generate_constraints` in prove mode: z_out = [1, 7 public inputs, witness..] as
 * (num_instance + num_witness) x 32 B canonical.  Values only, no matrices.  ZKB_ERR_SHAPE if the witness does not have
 * the circuit's shape.  Like the reference (release build: `debug_assert!(cs.is_satisfied())`), an unsatisfying witness is
 * NOT an error here; use zkb_l2_circuit_is_satisfied. */
int zkb_l2_circuit_assign(const zkb_l2_circuit* c, const zkb_l2_public_inputs* inputs, const zkb_l2_witness* witness,
                          uint8_t* z_out);
/* `cs.is_satisfied()`: *satisfied = 1 iff (A z) o (B z) = C z; first_bad_row (may be NULL) = `cs.which_is_unsatisfied()`. */
int zkb_l2_circuit_is_satisfied(const zkb_l2_circuit* c, const uint8_t* z, int* satisfied, uint64_t* first_bad_row);
/* The six roots a satisfying batch carries, computed off-circuit with the native Poseidon sponge
 * (prover/src/main.rs.bak:93-154 `calculate_new_root_offchain`, same folds for withdrawals / batch hash / commitments). */
int zkb_l2_roots(const zkb_l2_witness* witness, uint64_t batch_id, const uint8_t pre_shielded_root[32],
                 zkb_l2_public_inputs* out);
/* Poseidon(get_poseidon_config()) of n <= 3 field elements (32 B LE, reduced mod r): fresh sponge, absorb, squeeze one. */
int zkb_l2_poseidon_hash(const uint8_t* elems, size_t n, uint8_t out[32]);
/* The same hash for n independent inputs on the GPU (the account / transfer / withdrawal LEAF hashes of
 * l2_circuit.rs:315-330,477-490 are independent; the folds over them are sequential chains and stay on the host).
 * in: n x arity x 32 B (reduced mod r like the host function), 0 <= arity <= 3; out: n x 32 B.  _dev: device buffers, async. */
int zkb_l2_poseidon_hash_batch(zkb_ctx* ctx, int arity, const uint8_t* in_host, size_t n, uint8_t* out_host);
int zkb_l2_poseidon_hash_batch_dev(zkb_ctx* ctx, int arity, const void* in_dev, size_t n, void* out_dev);
/* the same n hashes on `threads` host threads (the native sponge the witness walkers use): host baseline / convenience */
int zkb_l2_poseidon_hash_batch_host(int arity, const uint8_t* in, size_t n, int threads, uint8_t* out);
/* get_poseidon_config() as canonical bytes: ark_out = 64 rounds x 3 lanes x 32 B, mds_out = 3 x 3 x 32 B (row-major). */
int zkb_l2_poseidon_params(uint8_t* ark_out, uint8_t* mds_out);
/* `StdRng::seed_from_u64(batch_id)` then `Fr::rand` twice (prover.rs:354 + ark-groth16 prove): canonical r, s. */
int zkb_l2_prover_randomness(uint64_t batch_id, uint8_t r[32], uint8_t s[32]);
/* `impl BatchProver for Groth16Prover { fn prove }` (prover.rs:350-425): proof_out = -A || B || C, 256 B
 * (proof_to_solana_bytes, prover.rs:304-334).  m = zkb_r1cs_load(zkb_l2_circuit_desc(c)); pk = the key made for c. */
int zkb_l2_prove(zkb_ctx* ctx, const zkb_pk* pk, const zkb_r1cs* m, const zkb_l2_circuit* c,
                 const zkb_l2_public_inputs* inputs, const zkb_l2_witness* witness, uint8_t proof_out[256]);


/* ---- batches of independent L2 proofs on one GPU (BASELINE.json config 5; the forge coordinator's chunk-per-worker
 * parallelism, forge/crates/prover-coordinator/src/dispatcher.rs:290-330, inside one process) ------------------------------
 * A zkb_l2_batch owns a pool of `lanes` host threads and a few device contexts (env ZKB_L2_SLOTS, default 4).  A batch is cut into sub-batches (<= 256 proofs,
 * env ZKB_L2_SUBBATCH): the pool assigns a sub-batch's witnesses into pinned memory, zkb_prove_batch_begin proves it with one
 * set of batched kernels, and the next sub-batch is assigned meanwhile.  pk, m and c are shared, read-only.
 * A call of at most 8 proofs runs zkb_prove per proof instead, one context each (lower latency when there is nothing to batch).
 * proofs_out: n x 256 B.  status_out (may be NULL): per-proof zkb_status.  Returns the first error, ZKB_OK if all succeeded;
 * a failed proof does not invalidate the others. */
typedef struct zkb_l2_batch zkb_l2_batch;
int zkb_l2_batch_create(int device, int lanes /* 1..64 */, zkb_l2_batch** out);
void zkb_l2_batch_destroy(zkb_l2_batch* b);
int zkb_l2_batch_lanes(const zkb_l2_batch* b);
int zkb_l2_batch_prove(zkb_l2_batch* b, const zkb_pk* pk, const zkb_r1cs* m, const zkb_l2_circuit* c,
                       const zkb_l2_public_inputs* inputs /* n */, const zkb_l2_witness* witnesses /* n */, size_t n,
                       uint8_t* proofs_out, int* status_out);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
