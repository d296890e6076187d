// QP-ADMM decoding kernel (sm_100a) -- DecodeQPADMM, algo/qp_admm.h:104-178.
//
// A persistent CTA keeps F frames in flight (slots.cuh); a lane is an (element,
// frame) pair and the whole iteration state stays on chip:
//   registers      yl[row], the dual of the previous iteration, 4 rows per block, KB blocks
//                  per lane.  The reference's z[] is not state: it is only read through
//                  w below, and yl' = max(0, yl - r) = max(0, -(r - yl)) exactly.
//   shared memory  per frame: w = yl + mu (z - b) of block rank R as two 16-byte chunks
//                  w01[R], w23[R] (the only form in which z, yl are read by the v-update,
//                  qp_admm.h:137); v and q + alpha/2 of variable rank R.  Per CTA:
//                  inv_coef and the incidence words.  Ranks come from admm_layout.cu, which
//                  orders variables and blocks so that the two static gathers below hit
//                  distinct shared-memory banks.
// A trip of the main loop is one iteration of every active slot:
//   v-phase  lane per variable: B = (q_i + alpha/2) + sum_j cf w_j with the rows j
//            ascending -- the reference's order, kept exactly (the iteration is
//            chaotic) -- then v = clip(B * inv_coef, 0, 1)
//   r-phase  lane per block: r = b - A v with the variables in ascending index order
//            (qp_admm.h:144-151), then t, z, yl, w and the partial stop sum
// Arithmetic is fp64 with every operation written as an _rn intrinsic so that nvcc
// cannot contract a multiply into a neighbouring add: v, the hard decisions and the
// iteration count are bit-identical to the reference.  The only licence taken is the
// summation order of the stop test sum2 (tree instead of row-ascending), which can
// move the exit by one iteration only if sum2 lands within an ulp-scale band around
// eps_stop.
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "admm_rows.cuh"
#include "slots.cuh"

namespace ldpc {

struct AdmmParams {
    KernelIO io;
    const AdmmBlock *blocks;
    const AdmmVarRec *vars;
    const uint32_t *inc;
    const uint16_t *var_id;     // rank -> variable index
    const uint16_t *var_rank;   // variable index (< n) -> rank
    int n_blocks, n_var, n_inc;
    int ne;                     // elements (ranks) per round = blockDim / F
    int stride_c;               // per-frame stride of w01 / w23 in 16-byte chunks
    int stride_v;               // per-frame stride of v / qa in doubles
    int max_iter;
    int infeasible;   // min(e) * mu <= alpha: DecodeQPADMM returns {zeros, false} (qp_admm.h:108-114)
    double alpha, mu, eps_stop;
};

template <int F, int KB>
__global__ void __launch_bounds__(512) qpadmm_kernel(const AdmmParams p) {
    extern __shared__ __align__(16) double smem[];
    const KernelIO &io = p.io;
    const int n = io.n, n_var = p.n_var, n_blocks = p.n_blocks;
    const int tid = threadIdx.x, nt = blockDim.x;
    // a warp holds 32/F consecutive elements of each of the F frames, frame-major: every quarter- and
    // half-warp then reads one frame at consecutive ranks, which is what the layout is optimised for
    constexpr int LPF = 32 / F;
    const int warp = tid >> 5, nwarps = nt >> 5, lane = tid & 31;
    const int f_lane = lane / LPF, elem = warp * LPF + lane % LPF, ne = p.ne;

    double2 *w01 = reinterpret_cast<double2 *>(smem);                 // F x stride_c chunks
    double2 *w23 = w01 + (size_t) F * p.stride_c;
    double *v = reinterpret_cast<double *>(w23 + (size_t) F * p.stride_c);   // F x stride_v
    double *qa = v + (size_t) F * p.stride_v;                          // F x stride_v: q_i + alpha/2
    double *inv = qa + (size_t) F * p.stride_v;                        // n_var: -1 / (mu e_i - alpha)
    double *red = inv + n_var;                                         // F x 32 partial stop sums
    SlotBlock<F> *S = reinterpret_cast<SlotBlock<F> *>(red + F * 32);
    uint32_t *inc_s = reinterpret_cast<uint32_t *>(S + 1);             // n_inc incidence words
    uint8_t *cw = reinterpret_cast<uint8_t *>(inc_s + p.n_inc);        // F x n (experiment mode)

    slots_init(S);
    // inv_coef, qp_admm.h:123-127 (A = (mu e - alpha)/2; inv = -1/(2A))
    for (int r = tid; r < n_var; r += nt) {
        const double A = __dmul_rn(__dadd_rn(__dmul_rn(p.mu, (double) p.vars[r].e), -p.alpha), 0.5);
        inv[r] = __ddiv_rn(-1.0, __dmul_rn(2.0, A));
    }
    for (int a = tid; a < p.n_inc; a += nt) inc_s[a] = p.inc[a];
    if (tid < F) v[(size_t) tid * p.stride_v + n_var] = 0.0;           // the absent-variable sentinel
    const double half_alpha = __dmul_rn(p.alpha, 0.5);

    double2 *w01_f = w01 + (size_t) f_lane * p.stride_c;
    double2 *w23_f = w23 + (size_t) f_lane * p.stride_c;
    double *v_f = v + (size_t) f_lane * p.stride_v;
    const double *qa_f = qa + (size_t) f_lane * p.stride_v;

    // this lane's blocks (static for the whole launch)
    AdmmBlock blk[KB];
    bool have[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const int rank = k * ne + elem;
        have[k] = rank < n_blocks;
        if (have[k]) blk[k] = p.blocks[rank];
        else { blk[k].var[0] = blk[k].var[1] = blk[k].var[2] = (uint16_t) n_var; blk[k].meta = 0x24; }
    }
    double t[KB][4];
#pragma unroll
    for (int k = 0; k < KB; ++k) t[k][0] = t[k][1] = t[k][2] = t[k][3] = 0.0;
    __syncthreads();

    for (;;) {
        slots_refill(io, S);
        if (S->alive == 0) break;
        unsigned newmask = 0, livemask = 0;
        int iter[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const int st = S->state[f];
            newmask |= (st == SLOT_NEW ? 1u : 0u) << f;
            livemask |= (st != SLOT_DEAD ? 1u : 0u) << f;
            iter[f] = S->iter[f];
        }
        if (newmask) {
            // q + alpha/2 and the v before the first update (qp_admm.h:116-119, visible only if max_iter == 0)
            slots_load<F>(io, S, newmask, nullptr, 0, cw, [&](int i, int f, double l) {
                const int r = p.var_rank[i];
                qa[(size_t) f * p.stride_v + r] = __dadd_rn(l, half_alpha);
                v[(size_t) f * p.stride_v + r] = (!p.infeasible && l > 0.0) ? 1.0 : 0.0;
            });
            if ((newmask >> f_lane) & 1u) {
                for (int r = elem; r < n_var; r += ne)
                    if (p.var_id[r] >= n) {                            // auxiliary variables: q = 0
                        qa[(size_t) f_lane * p.stride_v + r] = half_alpha;
                        v_f[r] = 0.0;
                    }
                // z = yl = 0 (t[][] holds yl), w = mu * (0 - b)
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    t[k][0] = t[k][1] = t[k][2] = t[k][3] = 0.0;
                    if (have[k]) {
                        w01_f[k * ne + elem] = make_double2(0.0, 0.0);
                        w23_f[k * ne + elem] = make_double2(0.0, (blk[k].meta >> 8) == 4 ? __dmul_rn(p.mu, -2.0) : 0.0);
                    }
                }
            }
            __syncthreads();
        }
        unsigned active = 0;
#pragma unroll
        for (int f = 0; f < F; ++f)
            if (((livemask >> f) & 1u) && !p.infeasible && iter[f] < p.max_iter) active |= 1u << f;
        const bool mine = (active >> f_lane) & 1u;

        // ---- v-phase, qp_admm.h:132-142
        if (mine) {
            for (int rank = elem; rank < n_var; rank += ne) {
                const AdmmVarRec rec = p.vars[rank];
                double B = qa_f[rank];
                const uint32_t *iw = inc_s + rec.inc_start;
                for (int a = 0; a < rec.inc_count; ++a) {
                    const uint32_t word = iw[a];
                    const int chunk = word & 0xffff;
                    const double2 a01 = w01_f[chunk], a23 = w23_f[chunk];
                    B = __dadd_rn(B, flip_by(a01.x, word));
                    B = __dadd_rn(B, flip_by(a01.y, word << 1));
                    B = __dadd_rn(B, flip_by(a23.x, word << 2));
                    B = __dadd_rn(B, a23.y);
                }
                double x = __dmul_rn(B, inv[rank]);
                x = (x < 0.0) ? 0.0 : x;
                x = (1.0 < x) ? 1.0 : x;
                v_f[rank] = x;
            }
        }
        __syncthreads();

        // ---- r-phase, qp_admm.h:144-159
        double part = 0.0;
        if (mine) {
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                if (!have[k]) continue;
                const int meta = blk[k].meta;
                const double v0 = v_f[blk[k].var[0]], v1 = v_f[blk[k].var[1]], v2 = v_f[blk[k].var[2]];
                double r0, r1, r2, r3, w0, w1, w2, w3;
                if ((meta >> 8) == 4) {              // a three-variable check: the common, branch-free case
                    switch (meta & 0x3f) {
                        case 0x24: residual_rows<0, 1, 2>(v0, v1, v2, 2.0, r0, r1, r2, r3); break;   // slots visited 0,1,2
                        case 0x21: residual_rows<1, 0, 2>(v0, v1, v2, 2.0, r0, r1, r2, r3); break;   // 1,0,2
                        default:   residual_rows<1, 2, 0>(v0, v1, v2, 2.0, r0, r1, r2, r3); break;   // 1,2,0 (0x09)
                    }
                    w0 = row_update<false>(r0, t[k][0], part, p.mu, 2.0);
                    w1 = row_update<false>(r1, t[k][1], part, p.mu, 2.0);
                    w2 = row_update<false>(r2, t[k][2], part, p.mu, 2.0);
                    w3 = row_update<true>(r3, t[k][3], part, p.mu, 2.0);
                } else {                             // degree-1 / degree-2 checks (slots 0,1,2 in order, b = 0):
                    const int rows = meta >> 8;      // the missing rows stay identically zero
                    residual_rows<0, 1, 2>(v0, v1, v2, 0.0, r0, r1, r2, r3);
                    w0 = row_update<false>(r0, t[k][0], part, p.mu, 0.0);
                    w1 = row_update<false>(rows < 2 ? 0.0 : r1, t[k][1], part, p.mu, 0.0);
                    w2 = row_update<false>(0.0, t[k][2], part, p.mu, 0.0);
                    w3 = row_update<false>(0.0, t[k][3], part, p.mu, 0.0);
                }
                w01_f[k * ne + elem] = make_double2(w0, w1);
                w23_f[k * ne + elem] = make_double2(w2, w3);
            }
        }
        // the lanes of one frame are LPF consecutive lanes of the warp
#pragma unroll
        for (int off = LPF / 2; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (lane % LPF == 0) red[f_lane * 32 + warp] = part;
        __syncthreads();

        // ---- per slot: stop test (qp_admm.h:161-163) / out of iterations
        unsigned finished = 0;
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (!((livemask >> f) & 1u)) continue;
            int iters = iter[f];
            bool done = true;
            if ((active >> f) & 1u) {
                double sum2 = 0.0;
                for (int w = 0; w < nwarps; ++w) sum2 += red[f * 32 + w];
                iters = iter[f] + 1;
                done = sum2 < p.eps_stop || iters >= p.max_iter;
            }
            if (!done) continue;
            finished |= 1u << f;
            // hard decisions, qp_admm.h:166-177
            const double *vf = v + (size_t) f * p.stride_v;
            int valid = 1;
            if (io.experiment) {
                int bad = 0;
                for (int c = tid; c < io.m; c += nt) {
                    int parity = 0;
                    for (int e = io.row_ptr[c]; e < io.row_ptr[c + 1]; ++e)
                        parity ^= vf[p.var_rank[io.col_idx[e]]] > 0.5 ? 1 : 0;
                    bad |= parity;
                }
                valid = !__syncthreads_or(bad);
            }
            slot_finish<F>(io, S, f, !p.infeasible, 1, valid, iters, cw,
                           [&](int i) { return vf[p.var_rank[i]] > 0.5 ? 1 : 0; },
                           [&](int i) { return vf[p.var_rank[i]]; });
        }
        if (finished) __syncthreads();      // finishing reads S->frame, which the refill overwrites
        if (tid == 0) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                if (!((livemask >> f) & 1u)) continue;
                if ((finished >> f) & 1u) S->state[f] = SLOT_EMPTY;
                else { S->iter[f] = iter[f] + 1; S->state[f] = SLOT_ACTIVE; }
            }
        }
    }
    slots_flush(io, S);
}

// ---------------------------------------------------------------- host side

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct AdmmShape {
    int F, kb, threads, ne, stride_c, stride_v;
    size_t smem;
};

static AdmmShape admm_shape(const ldpc_code *c, int F, int kb) {
    AdmmShape s;
    s.F = F; s.kb = kb;
    s.ne = round_up((c->n_blocks + kb - 1) / kb, 32 / F);   // whole warps: 32/F elements x F frames each
    s.threads = s.ne * F;
    s.stride_c = round_up(s.ne * kb, 8);
    s.stride_v = round_up(c->n_var + 1, 16);
    s.smem = (size_t) 2 * F * s.stride_c * 16 + (size_t) 2 * F * s.stride_v * 8 + (size_t) c->n_var * 8 +
             (size_t) F * 32 * 8 + sizeof(SlotBlock<4>) + (size_t) c->n_inc * 4 + (size_t) F * c->n + 32;
    return s;
}

using AdmmKernel = void (*)(const AdmmParams);

template <int F>
static AdmmKernel kernel_for(int kb) {
    switch (kb) {
        case 1: return qpadmm_kernel<F, 1>;
        case 2: return qpadmm_kernel<F, 2>;
        case 3: return qpadmm_kernel<F, 3>;
        case 4: return qpadmm_kernel<F, 4>;
        case 5: return qpadmm_kernel<F, 5>;
        case 6: return qpadmm_kernel<F, 6>;
        case 7: return qpadmm_kernel<F, 7>;
        default: return qpadmm_kernel<F, 8>;
    }
}

static AdmmKernel kernel_for(int F, int kb) {
    return F == 4 ? kernel_for<4>(kb) : (F == 2 ? kernel_for<2>(kb) : kernel_for<1>(kb));
}

// Frames per CTA (F) and blocks per lane (KB): the shape that keeps the most lanes busy per SM, i.e.
// resident warps (registers and shared memory both limit them) x the fraction of lanes that own a block.
static int choose_shape(const ldpc_code *c, int64_t frames, AdmmShape *out, int *per_sm_out, bool ignore_env = false) {
    int want_f = 0, want_kb = 0;
    if (!ignore_env) {
        if (const char *s = getenv("LDPC_ADMM_F")) want_f = atoi(s);
        if (const char *s = getenv("LDPC_ADMM_KB")) want_kb = atoi(s);
    }
    double best_score = -1;
    for (int F = 4; F >= 1; F >>= 1) {
        if ((want_f == 1 || want_f == 2 || want_f == 4) && F != want_f) continue;
        if (!want_f && F > 1 && frames < 2ll * 148 * F) continue;     // tiny batches: spread frames over the SMs
        if (!want_f && F == 4) continue;                              // layout is tuned for one frame per half-warp
        for (int kb = 1; kb <= 8; ++kb) {
            if (want_kb >= 1 && want_kb <= 8 && kb != want_kb) continue;
            const AdmmShape s = admm_shape(c, F, kb);
            if (s.threads > 512 || s.smem > 227 * 1024) continue;
            AdmmKernel fn = kernel_for(F, kb);
            LDPC_CUDA(allow_max_dynamic_smem(fn));
            int per_sm = 0;
            LDPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, s.threads, s.smem));
            if (per_sm < 1) continue;
            const double lanes_busy = (double) c->n_blocks * F / ((double) kb * s.threads);
            // measured on B200 (profiles/r01_admm_sweep.txt): throughput follows the number of frames resident
            // per SM (the dual state lives in registers, so 3..5 frames fit) times the lane utilisation;
            // more than 6 blocks per lane serialises too much work behind one dependency chain
            const double frames_per_sm = std::min(per_sm * F, 6);
            const double score = frames_per_sm * lanes_busy * (kb <= 6 ? 1.0 : 0.75);
            if (score > best_score + 1e-9) { best_score = score; *out = s; *per_sm_out = per_sm; }
        }
    }
    if (best_score < 0 && (want_f || want_kb))        // a forced shape that does not fit: fall back to the heuristic
        return choose_shape(c, frames, out, per_sm_out, true);   // (the caller's environment is left alone)
    if (best_score < 0)
        return fail(LDPC_E_UNSUPPORTED, "QP-ADMM state of this code exceeds one SM (shared memory / threads)");
    return LDPC_OK;
}

// Decoder::decode / exp() for QP-ADMM: the check-centric kernel where it applies (all checks of degree 0..12, a
// feasible (alpha, mu)), else the block-per-lane kernel below; LDPC_ADMM_KERNEL=block|check overrides (A/B runs).
// which kernel served the last QP-ADMM launch of this process (ldpc_debug_last_qpadmm_kernel): 1 check-centric, 2 block-per-lane
std::atomic<int> g_last_qpadmm_kernel{0};

int launch_qpadmm(const ldpc_code *c, const FrameIO &fio, int64_t frames, double var, double alpha, double mu,
                  int max_iter, double eps_stop, unsigned long long *queue, cudaStream_t stream) {
    bool chk = true;
    if (const char *k = getenv("LDPC_ADMM_KERNEL")) chk = k[0] != 'b';
    if (chk) {
        const int st = launch_qpadmm_chk(c, fio, frames, var, alpha, mu, max_iter, eps_stop, queue, stream);
        if (st != LDPC_E_UNSUPPORTED) {
            g_last_qpadmm_kernel = 1;
            return st;
        }
    }
    g_last_qpadmm_kernel = 2;
    return launch_qpadmm_blk(c, fio, frames, var, alpha, mu, max_iter, eps_stop, queue, stream);
}

int launch_qpadmm_blk(const ldpc_code *c, const FrameIO &fio, int64_t frames, double var, double alpha, double mu,
                      int max_iter, double eps_stop, unsigned long long *queue, cudaStream_t stream) {
    if (frames <= 0) return LDPC_OK;
    AdmmParams p;
    KernelIO &io = p.io;
    io.y = fio.y; io.bits = fio.bits; io.ok = fio.ok; io.iters = fio.iters; io.soft = fio.soft;
    io.experiment = fio.experiment; io.cw_source = fio.cw_source; io.seed = fio.seed;
    io.frame_begin = fio.frame_begin; io.words = fio.words; io.n_words = fio.n_words;
    io.counters = fio.counters; io.gen_cols = c->d.gen_cols; io.k = c->k; io.k_words = c->k_words;
    io.frames = frames; io.queue = queue; io.var = var; io.sigma = std::sqrt(var);
    io.n = c->n; io.m = c->m; io.row_ptr = c->d.row_ptr; io.col_idx = c->d.col_idx;
    p.blocks = c->d.blocks; p.vars = c->d.admm_var; p.inc = c->d.admm_inc; p.var_id = c->d.admm_var_id;
    p.var_rank = c->d.admm_var_rank;
    p.n_blocks = c->n_blocks; p.n_var = c->n_var; p.n_inc = c->n_inc; p.max_iter = max_iter;
    p.alpha = alpha; p.mu = mu; p.eps_stop = eps_stop;
    p.infeasible = (double) c->e_min * mu <= alpha;

    AdmmShape shape{};
    int per_sm = 0, sms = 0;
    int st = choose_shape(c, frames, &shape, &per_sm);
    if (st) return st;
    LDPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    p.ne = shape.ne; p.stride_c = shape.stride_c; p.stride_v = shape.stride_v;
    const long long want = (frames + shape.F - 1) / shape.F;
    const long long grid = std::min<long long>((long long) per_sm * sms, want);
    kernel_for(shape.F, shape.kb)<<<(unsigned) grid, shape.threads, shape.smem, stream>>>(p);
    LDPC_CUDA(cudaGetLastError());
    return LDPC_OK;
}

}  // namespace ldpc
