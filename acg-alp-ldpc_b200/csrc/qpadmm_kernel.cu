// QP-ADMM decoding kernel (sm_100a) -- DecodeQPADMM, algo/qp_admm.h:104-178.
//
// One persistent CTA decodes one frame at a time; the whole iteration state stays
// on chip:
//   registers      t[row] = r - yl of the previous iteration, 4 rows per block,
//                  KB blocks per thread.  z = max(t, 0) and yl = max(-t, 0)
//                  (qp_admm.h:156-157: yl - r == -(r - yl) exactly), so one double
//                  per inequality row replaces the reference's z[] and yl[].
//   shared memory  w[row]  = yl + mu (z - b)   (the only form in which z, yl are
//                  read by the v-update, qp_admm.h:137), v[var], q + alpha/2, inv_coef
// An iteration is two phases separated by barriers:
//   v-phase  thread per variable: B = (q_i + alpha/2) + sum_j cf (w_j), rows j
//            ascending -- the reference's order, kept exactly (the iteration is
//            chaotic) -- then v = clip(B * inv_coef, 0, 1)
//   r-phase  thread per block: r = b - A v with the variables in ascending
//            index order (qp_admm.h:144-151), t, z, yl, w, and the partial stop sum
// Arithmetic is fp64 with every operation written as an _rn intrinsic so that nvcc
// cannot contract a multiply into a neighbouring add: the results (v, hard
// decisions, iteration count) are bit-identical to the reference.  The only
// licence taken is the summation order of the stop test sum2 (tree instead of
// row-ascending), which can move the exit by one iteration only if sum2 lands
// within an ulp-scale band around eps_stop.
#include <algorithm>
#include <cstdlib>

#include "frame.cuh"

namespace ldpc {

struct AdmmParams {
    KernelIO io;
    const AdmmBlock *blocks;
    const uint16_t *blk_order;
    const uint32_t *var_ptr;
    const uint16_t *inc;
    const uint16_t *var_order;
    const uint8_t *var_e;
    int n_blocks, n_var;
    int max_iter;
    int infeasible;   // min(e) * mu <= alpha: DecodeQPADMM returns {zeros, false} (qp_admm.h:108-114)
    double alpha, mu, eps_stop;
};

__device__ __forceinline__ double flip_if(double x, bool neg) {
    return __hiloint2double(__double2hiint(x) ^ (neg ? 0x80000000 : 0), __double2loint(x));
}

// r[q] = b[q] - sum_k cf(q, S_k) * vs[k] with the block's variables visited in
// ascending index order; S_k = slot of the k-th visited variable; cf(q, s) = +1 if
// q == 3 or q == s, else -1.  The leading "0 -/+ v" of rows 0..2 is folded into a
// sign (differs from the reference only in the sign of an exact zero).
template <int S0, int S1, int S2>
__device__ __forceinline__ void residual_rows(const double vs[3], double b3, double r[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const bool p0 = (q == 3 || q == S0), p1 = (q == 3 || q == S1), p2 = (q == 3 || q == S2);
        double x = (q == 3) ? __dadd_rn(b3, -vs[0]) : (p0 ? -vs[0] : vs[0]);
        x = __dadd_rn(x, p1 ? -vs[1] : vs[1]);
        x = __dadd_rn(x, p2 ? -vs[2] : vs[2]);
        r[q] = x;
    }
}

template <int KB>
__global__ void __launch_bounds__(512) qpadmm_kernel(const AdmmParams p) {
    extern __shared__ __align__(16) double smem[];
    const KernelIO &io = p.io;
    const int n = io.n, n_var = p.n_var, n_blocks = p.n_blocks;
    const int tid = threadIdx.x, nt = blockDim.x;

    double *w = smem;                         // 4 * n_blocks (rows padded to 4 per block)
    double *v = w + 4 * n_blocks;             // n_var + 1 (last = 0, the absent-variable sentinel)
    double *qa = v + (n_var + 1);             // n_var: q_i + alpha/2
    double *inv = qa + n_var;                 // n_var: -1 / (mu e_i - alpha)
    double *red = inv + n_var;                // 32 warp partials of the stop sum
    FrameScratch *scratch = (FrameScratch *) (red + 32);
    uint8_t *hard = (uint8_t *) (scratch + 1);
    uint8_t *cw = hard + n;

    scratch_init(scratch);
    // inv_coef, qp_admm.h:123-127 (A = (mu e - alpha)/2; inv = -1/(2A))
    for (int i = tid; i < n_var; i += nt) {
        const double A = __dmul_rn(__dadd_rn(__dmul_rn(p.mu, (double) p.var_e[i]), -p.alpha), 0.5);
        inv[i] = __ddiv_rn(-1.0, __dmul_rn(2.0, A));
    }
    if (tid == 0) v[n_var] = 0.0;
    const double half_alpha = __dmul_rn(p.alpha, 0.5);

    // this thread's blocks (static for the whole launch)
    int blk_id[KB];
    AdmmBlock blk[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) {
        const int idx = tid + k * nt;
        blk_id[k] = idx < n_blocks ? (int) p.blk_order[idx] : -1;
        if (blk_id[k] >= 0) blk[k] = p.blocks[blk_id[k]];
        else { blk[k].var[0] = blk[k].var[1] = blk[k].var[2] = (uint16_t) n_var; blk[k].meta = 0x24; }
    }

    for (;;) {
        const long long f = next_frame(io, scratch);
        if (f < 0) break;
        load_frame(io, f, qa, cw, scratch);           // qa[0..n) = LLR for now
        // v before the first update (qp_admm.h:116-119); only visible when max_iter == 0
        for (int i = tid; i < n; i += nt) v[i] = p.infeasible ? 0.0 : (qa[i] > 0.0 ? 1.0 : 0.0);
        for (int i = tid; i < n_var; i += nt) qa[i] = __dadd_rn(i < n ? qa[i] : 0.0, half_alpha);

        // z = yl = 0  ->  t = 0, w = mu * (0 - b)
        double t[KB][4];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            t[k][0] = t[k][1] = t[k][2] = t[k][3] = 0.0;
            if (blk_id[k] >= 0) {
                const int rows = blk[k].meta >> 8;
                double *wb = w + 4 * blk_id[k];
                wb[0] = wb[1] = wb[2] = 0.0;
                wb[3] = rows == 4 ? __dmul_rn(p.mu, -2.0) : 0.0;
            }
        }
        __syncthreads();


        int iters = 0;
        const int max_iter = p.infeasible ? 0 : p.max_iter;
        for (int iter = 0; iter < max_iter; ++iter) {
            iters = iter + 1;
            // ---- v-phase, qp_admm.h:132-142
            for (int j = tid; j < n_var; j += nt) {
                const int i = p.var_order[j];
                double B = qa[i];
                const uint32_t end = p.var_ptr[i + 1];
                for (uint32_t a = p.var_ptr[i]; a < end; ++a) {
                    const uint32_t code = p.inc[a];
                    const int slot = code & 3;
                    const double2 *wb = reinterpret_cast<const double2 *>(w + 4 * (code >> 2));
                    const double2 w01 = wb[0], w23 = wb[1];
                    B = __dadd_rn(B, flip_if(w01.x, slot != 0));
                    B = __dadd_rn(B, flip_if(w01.y, slot != 1));
                    B = __dadd_rn(B, flip_if(w23.x, slot != 2));
                    B = __dadd_rn(B, w23.y);
                }
                double x = __dmul_rn(B, inv[i]);
                x = (x < 0.0) ? 0.0 : x;
                x = (1.0 < x) ? 1.0 : x;
                v[i] = x;
            }
            __syncthreads();
            // ---- r-phase, qp_admm.h:144-159
            double part = 0.0;
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                if (blk_id[k] < 0) continue;
                const int rows = blk[k].meta >> 8;
                const double vs[3] = {v[blk[k].var[0]], v[blk[k].var[1]], v[blk[k].var[2]]};
                const double b3 = rows == 4 ? 2.0 : 0.0;
                double r[4];
                switch (blk[k].meta & 0x3f) {
                    case 0x24: residual_rows<0, 1, 2>(vs, b3, r); break;   // slots visited 0,1,2
                    case 0x21: residual_rows<1, 0, 2>(vs, b3, r); break;   // 1,0,2
                    default:   residual_rows<1, 2, 0>(vs, b3, r); break;   // 1,2,0 (0x09)
                }
                double wn[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double rq = q < rows ? r[q] : 0.0;
                    const double told = t[k][q];
                    const double yl_old = (told < 0.0) ? -told : 0.0;
                    const double tn = __dadd_rn(rq, -yl_old);
                    const double z = (0.0 < tn) ? tn : 0.0;
                    const double yl = (tn < 0.0) ? -tn : 0.0;
                    const double d = __dadd_rn(z, -rq);
                    part = __fma_rn(d, d, part);
                    const double zb = (q == 3) ? __dadd_rn(z, -b3) : z;
                    wn[q] = __fma_rn(p.mu, zb, yl);
                    t[k][q] = tn;
                }
                double2 *wb = reinterpret_cast<double2 *>(w + 4 * blk_id[k]);
                wb[0] = make_double2(wn[0], wn[1]);
                wb[1] = make_double2(wn[2], wn[3]);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
            if ((tid & 31) == 0) red[tid >> 5] = part;
            __syncthreads();
            double sum2 = 0.0;
            for (int wi = 0; wi < (nt >> 5); ++wi) sum2 += red[wi];
            if (sum2 < p.eps_stop) break;             // qp_admm.h:161-163
        }

        // hard decisions, qp_admm.h:166-177
        for (int i = tid; i < n; i += nt) hard[i] = (v[i] <= 0.5) ? 0 : 1;
        int valid = 1;
        if (io.experiment) {
            __syncthreads();
            valid = syndrome_ok(io, hard);
        } else {
            __syncthreads();
        }
        finish_frame(io, f, hard, cw, v, !p.infeasible, 1, valid, iters, scratch);
    }
    scratch_flush(io, scratch);
}

// ---------------------------------------------------------------- host side

static size_t admm_smem_bytes(const ldpc_code *c) {
    size_t doubles = (size_t) 4 * c->n_blocks + (c->n_var + 1) + 2 * (size_t) c->n_var + 32;
    return doubles * sizeof(double) + sizeof(FrameScratch) + 2 * (size_t) c->n + 16;
}

// blocks per thread and CTA size: best lane utilisation with 96..512 threads
static void admm_shape(const ldpc_code *c, int *kb_out, int *threads_out) {
    int best_kb = 1, best_nt = 32;
    double best = -1;
    const char *force = getenv("LDPC_ADMM_KB");
    if (force && (atoi(force) < 1 || atoi(force) > 8 || (c->n_blocks + atoi(force) - 1) / atoi(force) > 512))
        force = nullptr;   // not a usable shape for this code: fall back to the heuristic
    for (int kb = 1; kb <= 8; ++kb) {
        if (force && atoi(force) != kb) continue;
        int nt = ((c->n_blocks + kb - 1) / kb + 31) / 32 * 32;
        if (nt > 512) continue;
        if (nt < 32) nt = 32;
        double eff = (double) c->n_blocks / ((double) kb * nt);
        if (nt < 96 || nt > 512) eff *= 0.8;
        if (eff > best + 1e-9) { best = eff; best_kb = kb; best_nt = nt; }
    }
    *kb_out = best_kb;
    *threads_out = best_nt;
}

template <int KB>
static int launch_kb(const AdmmParams &p, int threads, size_t smem, int64_t frames, int device,
                     cudaStream_t stream) {
    LDPC_CUDA(cudaFuncSetAttribute(qpadmm_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int per_sm = 0, sms = 0;
    LDPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, qpadmm_kernel<KB>, threads, smem));
    LDPC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (per_sm < 1) return fail(LDPC_E_UNSUPPORTED, "QP-ADMM state of this code does not fit on one SM");
    long long grid = std::min<long long>((long long) per_sm * sms, frames);
    qpadmm_kernel<KB><<<(unsigned) grid, threads, smem, stream>>>(p);
    LDPC_CUDA(cudaGetLastError());
    return LDPC_OK;
}

int launch_qpadmm(const ldpc_code *c, const FrameIO &fio, int64_t frames, double var, double alpha, double mu,
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
    p.blocks = c->d.blocks; p.blk_order = c->d.blk_order; p.var_ptr = c->d.var_ptr; p.inc = c->d.inc;
    p.var_order = c->d.var_order; p.var_e = c->d.var_e;
    p.n_blocks = c->n_blocks; p.n_var = c->n_var; p.max_iter = max_iter;
    p.alpha = alpha; p.mu = mu; p.eps_stop = eps_stop;
    p.infeasible = (double) c->e_min * mu <= alpha;

    size_t smem = admm_smem_bytes(c);
    if (smem > 227 * 1024) return fail(LDPC_E_UNSUPPORTED, "QP-ADMM state of this code exceeds 227 KB of shared memory");
    int kb, threads;
    admm_shape(c, &kb, &threads);
    switch (kb) {
        case 1: return launch_kb<1>(p, threads, smem, frames, c->device, stream);
        case 2: return launch_kb<2>(p, threads, smem, frames, c->device, stream);
        case 3: return launch_kb<3>(p, threads, smem, frames, c->device, stream);
        case 4: return launch_kb<4>(p, threads, smem, frames, c->device, stream);
        case 5: return launch_kb<5>(p, threads, smem, frames, c->device, stream);
        case 6: return launch_kb<6>(p, threads, smem, frames, c->device, stream);
        case 7: return launch_kb<7>(p, threads, smem, frames, c->device, stream);
        default: return launch_kb<8>(p, threads, smem, frames, c->device, stream);
    }
}

}  // namespace ldpc
