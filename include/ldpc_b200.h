/* ldpc_b200.h -- C ABI of the B200 (sm_100a) batched LDPC decoding engine.
 *
 * This is the drop-in boundary for ONE hot path of GreatDrake/acg-alp-ldpc:
 * belief-propagation and QP-ADMM decoding of many independent AWGN frames.
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 * Each entry point names the reference interface it replaces (file:line in the
 * reference tree).  The reference-side bindings (the C++ Decoder adapters and
 * the ctypes stub) are in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns an int status: LDPC_OK (0) or a negative LDPC_E_*;
 *     ldpc_last_error() gives the message of the calling thread's last failure.
 *     There is NO CPU fallback: without a usable CUDA device every compute entry
 *     point fails with LDPC_E_CUDA.
 *   - a code handle is immutable after creation and may be used from many host
 *     threads at once (the reference calls Decoder::decode concurrently from up to
 *     200 pthreads, experiment.h:128-130); per-call workspaces are internal.
 *   - frames are rows: y is frames x n doubles (raw channel samples, NOT LLRs:
 *     the LLR scaling 2*y/sigma^2 happens on the device exactly as
 *     utils/channel.h:12-16 does it), bits is frames x n bytes (0/1).
 *   - snr is Es/N0 in dB (utils/channel.h:12).
 *   - "_device" variants take device pointers and a cudaStream_t (as void*) and
 *     are asynchronous; the plain variants take host pointers, copy in and out,
 *     and return when the results are in the caller's buffers.
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_B200_ABI_VERSION 1

enum {
    LDPC_OK = 0,
    LDPC_E_INVALID = -1,   /* bad argument */
    LDPC_E_CUDA = -2,      /* CUDA runtime / no device */
    LDPC_E_NOMEM = -3,
    LDPC_E_UNSUPPORTED = -4 /* code too large for the on-chip layout */
};

typedef struct ldpc_code ldpc_code_t;

/* Shape of a compiled code. */
typedef struct {
    int32_t m, n;        /* checks, variables */
    int32_t edges;       /* ones in H */
    int32_t max_row_deg, max_col_deg;
    int32_t admm_blocks; /* three-variable checks (+ degree-1/2 special blocks) */
    int32_t admm_n_var;  /* n + auxiliary variables */
    int32_t admm_rows;   /* inequality rows R */
    int32_t admm_nnz;
    int32_t admm_e_min;  /* min_i sum_j A_ji^2, used by the feasibility test qp_admm.h:108-114 */
    int32_t k;           /* generator rows attached with ldpc_code_set_generator, else 0 */
    int32_t device;
    /* shared-memory wavefronts replayed per QP-ADMM iteration because of bank conflicts, for the natural
       ordering of variables/blocks and for the ordering the graph compiler chose */
    int32_t admm_conflicts_natural, admm_conflicts_laid_out;
} ldpc_code_info_t;

/* Counter block of one Monte-Carlo point.  The first fields are the reference's
 * ExperimentResult / HammingDistanceTracker (experiment.h:25-68) widened to
 * 64 bit; bit_errors / sum_iters / frames_with_bits are extensions:
 * bit errors are counted over frames that return n bits (QP-ADMM: all frames;
 * BP: converged frames). */
enum {
    LDPC_CNT_TOTAL = 0,
    LDPC_CNT_CORRECT,
    LDPC_CNT_PSEUDO,
    LDPC_CNT_DECODER_FAIL,
    LDPC_CNT_BIT_ERRORS,
    LDPC_CNT_SUM_HAMMING,
    LDPC_CNT_SUM_HAMMING_OK,
    LDPC_CNT_SUM_HAMMING_WRONG,
    LDPC_CNT_SUM_ITERS,
    LDPC_CNT_FRAMES_WITH_BITS,
    LDPC_CNT_COUNT
};

enum { LDPC_ALGO_BP = 0, LDPC_ALGO_QPADMM = 1 };

/* Where the transmitted codewords of an experiment come from. */
enum {
    LDPC_CW_ZERO = 0,      /* all-zero codeword */
    LDPC_CW_TABLE = 1,     /* caller's table: frame f sends words[f % n_words] (experiment.h:86-93) */
    LDPC_CW_GENERATOR = 2  /* c = u*G on device, u from Philox stream 0 (utils/channel.h:29-36) */
};

/* Decoder configuration: the constructor arguments of BeliefPropagationDecoder
 * (algo/bp.h:210) and QPADMMDecoder (algo/qp_admm.h:182). */
typedef struct {
    int32_t algo;       /* LDPC_ALGO_* */
    int32_t max_iter;
    int32_t early_exit; /* 1 = reference behaviour (BP: syndrome exit bp.h:195-196).  0 = fixed-iteration
                           measurement mode for BP (QP-ADMM uses eps_stop = 0 for that, as the reference can) */
    int32_t reserved;
    double alpha, mu, eps_stop; /* QP-ADMM only */
} ldpc_algo_cfg_t;

int ldpc_abi_version(void);
const char *ldpc_last_error(void);
int ldpc_device_count(int *count);

/* ---- code handle: replaces the per-frame graph builds from_biadjacency_matrix
 * (algo/bp.h:136-153) and ConstructADMMProblem (algo/qp_admm.h:13-102) by one
 * compile + upload per H.  CSR: row_ptr[m+1], col_idx[edges], columns ascending
 * within a row.  `device` is the CUDA ordinal the tables are uploaded to. */
int ldpc_code_create(int32_t m, int32_t n, const int32_t *row_ptr, const int32_t *col_idx, int device,
                     ldpc_code_t **out);
/* Same from the reference's dense TMatrix layout (m x n bytes, row-major, nonzero = 1). */
int ldpc_code_create_dense(int32_t m, int32_t n, const uint8_t *H, int device, ldpc_code_t **out);
void ldpc_code_destroy(ldpc_code_t *code);
int ldpc_code_info(const ldpc_code_t *code, ldpc_code_info_t *info);
/* Attach G (k x n dense bytes, e.g. GetOrtogonal(H).first, utils/codeword.h:97-128)
 * for LDPC_CW_GENERATOR experiments.  Not thread-safe against running calls. */
int ldpc_code_set_generator(ldpc_code_t *code, int32_t k, const uint8_t *G);

/* ---- belief propagation: BeliefPropagationDecoder::decode (algo/bp.h:208-222),
 * batched.  ok[f] = the decoder's bool; bits of frames with ok == 0 are zero
 * (the reference returns an EMPTY codeword, bp.h:198).  iters[f] = iterations
 * run.  post_llr (frames x n, may be NULL) = VNode::estimate() (bp.h:85-90) of
 * the last iteration. */
int ldpc_bp_decode(const ldpc_code_t *code, const double *y, int64_t frames, double snr, int32_t max_iter,
                   int32_t early_exit, uint8_t *bits, uint8_t *ok, int32_t *iters, double *post_llr);
int ldpc_bp_decode_device(const ldpc_code_t *code, const double *d_y, int64_t frames, double snr,
                          int32_t max_iter, int32_t early_exit, uint8_t *d_bits, uint8_t *d_ok,
                          int32_t *d_iters, double *d_post_llr, void *stream);

/* ---- QP-ADMM: QPADMMDecoder::decode -> DecodeQPADMM (algo/qp_admm.h:104-194),
 * batched, fp64, the reference's floating-point operation order.  ok[f] = 0
 * only for the infeasible-parameter exit (min(e)*mu <= alpha, qp_admm.h:108-114;
 * bits all zero, iters 0).  v_out (frames x n, may be NULL) = relaxed solution
 * v[0..n) at exit. */
int ldpc_qpadmm_decode(const ldpc_code_t *code, const double *y, int64_t frames, double snr, double alpha,
                       double mu, int32_t max_iter, double eps_stop, uint8_t *bits, uint8_t *ok,
                       int32_t *iters, double *v_out);
int ldpc_qpadmm_decode_device(const ldpc_code_t *code, const double *d_y, int64_t frames, double snr,
                              double alpha, double mu, int32_t max_iter, double eps_stop, uint8_t *d_bits,
                              uint8_t *d_ok, int32_t *d_iters, double *d_v_out, void *stream);

/* ---- channel: transmit() (utils/channel.h:19-26) from a counter-based
 * Philox4x32-10 stream instead of mt19937 + std::normal_distribution.
 * Global frame index f uses counter (f, block, stream); the CPU oracle replays
 * the same y bit for bit.  codewords: NULL (all-zero) or frames x n bytes.
 * y: frames x n doubles (host). */
int ldpc_channel_generate(const ldpc_code_t *code, uint64_t seed, uint64_t frame_begin, int64_t frames,
                          double snr, const uint8_t *codewords, double *y);
/* Same with device pointers (d_codewords may be NULL), asynchronous on `stream`. */
int ldpc_channel_generate_device(const ldpc_code_t *code, uint64_t seed, uint64_t frame_begin, int64_t frames,
                                 double snr, const uint8_t *d_codewords, double *d_y, void *stream);
/* The codewords an LDPC_CW_GENERATOR experiment transmits (frames x n bytes, host). */
int ldpc_generator_codewords(const ldpc_code_t *code, uint64_t seed, uint64_t frame_begin, int64_t frames,
                             uint8_t *codewords);

/* ---- Monte-Carlo point: exp() + multithread_experiment() (experiment.h:80-139)
 * for global frames [frame_begin, frame_begin + frame_count): device-side
 * codeword selection, AWGN, decoding, verdict (experiment.h:109-118) and
 * counting; only the counter block comes back.  words/n_words are used by
 * LDPC_CW_TABLE (n_words x n bytes, host).  gpu_seconds (may be NULL) = device
 * time of the decode kernels (CUDA events). */
int ldpc_experiment_run(const ldpc_code_t *code, const ldpc_algo_cfg_t *cfg, double snr, uint64_t seed,
                        uint64_t frame_begin, uint64_t frame_count, int32_t codeword_source,
                        const uint8_t *words, uint64_t n_words, uint64_t counters[LDPC_CNT_COUNT],
                        double *gpu_seconds);

/* ---- (alpha, mu) grid search: the double loop of qpadmm_params.cpp:51-67, whose body (estimate_qpadmm,
 * :16-30) is multithread_experiment() with QPADMMDecoder(alpha, mu, max_iter, eps_stop).  ONE launch evaluates
 * all `points` parameter pairs on the SAME frames [frame_begin, frame_begin + frame_count) (as the reference
 * reuses its codewords and noise seeds for every pair): the work items (point, frame) share one queue, so the
 * heavy-tailed iteration counts of one pair never idle the GPU.  counters: points x LDPC_CNT_COUNT (host), in
 * the order of alpha[] / mu[].  Pairs with min(e) * mu <= alpha take the reference's {zeros, false} exit
 * (qp_admm.h:108-114) as in ldpc_experiment_run. */
int ldpc_qpadmm_grid_run(const ldpc_code_t *code, int32_t points, const double *alpha, const double *mu,
                         int32_t max_iter, double eps_stop, double snr, uint64_t seed, uint64_t frame_begin,
                         uint64_t frame_count, int32_t codeword_source, const uint8_t *words, uint64_t n_words,
                         uint64_t *counters, double *gpu_seconds);

/* ---- multi-GPU: frames are sharded by GLOBAL frame index, so the counters are identical for any
 * number of GPUs, and the path's only collective is one all-reduce (sum, uint64) of the counter block over NVLink,
 * done by NCCL (libnccl.so.2, loaded at first use; LDPC_E_UNSUPPORTED when it cannot be loaded).
 * It replaces merge_exp_results() after pthread_join (experiment.h:70-78, 125-139).
 *
 * One process, several GPUs (the C++ drivers): codes[g] is the handle of the same H on device g.  The shards
 * [frame_begin + frame_count * g / n, frame_begin + frame_count * (g + 1) / n) run concurrently, the counter blocks are
 * all-reduced on the devices (ncclCommInitAll communicator, cached per device list) and the sum is returned.
 * gpu_seconds (may be NULL) = the longest device time of a shard. */
int ldpc_experiment_run_multi(const ldpc_code_t *const *codes, int32_t n_devices, const ldpc_algo_cfg_t *cfg, double snr,
                              uint64_t seed, uint64_t frame_begin, uint64_t frame_count, int32_t codeword_source,
                              const uint8_t *words, uint64_t n_words, uint64_t counters[LDPC_CNT_COUNT],
                              double *gpu_seconds);

/* One process per GPU (torchrun / MPI style): rank 0 makes an id (128 bytes) and hands it to the other ranks by any
 * means (bench.py: torch.distributed broadcast of the bytes); every rank joins with ldpc_comm_init on its device, then
 * ldpc_allreduce_counters sums `count` 64-bit counters in place (host array) over all ranks. */
#define LDPC_COMM_ID_BYTES 128
typedef struct ldpc_comm ldpc_comm_t;
int ldpc_comm_unique_id(uint8_t id[LDPC_COMM_ID_BYTES]);
int ldpc_comm_init(int32_t rank, int32_t world, const uint8_t id[LDPC_COMM_ID_BYTES], int device, ldpc_comm_t **out);
int ldpc_allreduce_counters(ldpc_comm_t *comm, uint64_t *counters, int32_t count);
void ldpc_comm_destroy(ldpc_comm_t *comm);

/* ---- pinned host buffers for the host-pointer entry points (optional: any
 * host memory works, pinned memory makes the copies asynchronous). */
int ldpc_host_alloc(void **ptr, uint64_t bytes);
int ldpc_host_free(void *ptr);

/* ---- roofline support: measured issue rate of dependent-free fp64 FMAs on the
 * device of `device` (G FMA instructions/s per thread-op, i.e. lanes x clock),
 * used as the denominator of the FP64 roofline in bench.py. */
int ldpc_measure_fp64_peak(int device, double *gfma_per_s);
/* ... and the measured shared-memory load bandwidth (GB/s, conflict-free 16-byte loads on all SMs), the
 * denominator of the shared-memory roofline of the QP-ADMM kernel. */
int ldpc_measure_smem_peak(int device, double *gbytes_per_s);

/* ---- testing hook: the two fp64 kernels of the BP message update evaluated element-wise on the
 * device (host arrays of `count` doubles): out_exp[i] = exp(-a[i]) for a >= 0, out_log[i] =
 * log(ev[i] / od[i]) for ev >= od >= 0. */
int ldpc_debug_bpmath(int device, int32_t count, const double *a, const double *ev, const double *od,
                      double *out_exp, double *out_log);

/* ---- testing hook: the shared-memory layout of the likelihood-ratio BP kernel for `frames_per_cta` (2, 4, 8, 16)
 * frames per CTA.  out = {message slots per frame, 1 if even-degree check classes are padded to an odd stride,
 * variable-pass accesses of paired lane groups that share a half line, ... examined, the same two numbers for the
 * check pass}.  With 8 frames per CTA a shared half line is a replayed shared-memory wavefront. */
int ldpc_debug_bp_layout(const ldpc_code_t *code, int32_t frames_per_cta, int32_t out[6]);

/* ---- testing hook: which kernel served the last QP-ADMM launch of this process (decode or experiment mode):
 * 1 = check-centric (qpadmm_chk_kernel.cu), 2 = block-per-lane (qpadmm_kernel.cu), 0 = none yet.  The parity tests
 * assert it, so a silent fall-back to the slower kernel fails them. */
int ldpc_debug_last_qpadmm_kernel(void);
/* ... and the last BP launch: 1 = likelihood-ratio kernel (bp_lr_kernel.cu), 2 = log-domain kernel (bp_kernel.cu). */
int ldpc_debug_last_bp_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
