/* oracle/ldpc_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the reference's hot path
 * (GreatDrake/acg-alp-ldpc), used as the checker for the CUDA kernels.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libldpc_b200.so) never does.
 *
 * What is restated, and the reference lines each part follows:
 *   orc_llr_variance / orc_llr      utils/channel.h:12-16
 *   orc_bp_decode_*                 algo/bp.h:34 (phi), :49-57 (check message),
 *                                   :77-83 (variable message), :85-90 (estimate),
 *                                   :183-199 (schedule, syndrome exit, empty
 *                                   codeword on failure)
 *   orc_admm_build                  algo/qp_admm.h:13-102 (ConstructADMMProblem)
 *   orc_qpadmm_decode               algo/qp_admm.h:104-178 (DecodeQPADMM) -- the
 *                                   floating-point operation ORDER is kept exactly
 *   orc_syndrome_ok                 utils/codeword.h:90-95 (IsCodeword)
 *   orc_experiment                  experiment.h:33-46 (channel Hamming distance),
 *                                   :109-118 (verdict), :70-78 (counter merge)
 *
 * What is NOT in the reference and is specified here instead:
 *   orc_philox4x32_10, orc_channel_frame, orc_info_bits -- the counter-based
 *   AWGN / information-bit stream the GPU generates on device.  The Gaussian
 *   transform uses only IEEE-754 correctly rounded operations (+ - * / sqrt fma)
 *   in a fixed order, so this file replays the device's y bit for bit.
 *
 * Parity pinning: the reference has no golden vectors or tests.
 * This oracle is pinned against the reference itself, compiled unmodified into
 * oracle/_ref/libref_oracle.so (oracle/Makefile), in tests/test_oracle_vs_ref.py,
 * and against fixtures generated from that library (tests/golden/).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (never -ffast-math, never
 * -march=native: the QP-ADMM arithmetic must stay plain IEEE double).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ channel */

double orc_llr_variance(double snr) { return pow(10, -(snr / 10)) / 2; }

double orc_llr(double v, double snr) { return 2 * v / orc_llr_variance(snr); }

/* -------------------------------------------------------------- Tanner graph */

/* Dense 0/1 bytes (m x n, row-major) -> CSR.  row_ptr has m+1 entries,
 * col_idx has E entries, columns ascending inside a row.  Returns E. */
int orc_dense_to_csr(const uint8_t *H, int m, int n, int *row_ptr, int *col_idx) {
    int e = 0;
    for (int r = 0; r < m; ++r) {
        row_ptr[r] = e;
        for (int c = 0; c < n; ++c)
            if (H[(size_t) r * n + c]) {
                if (col_idx) col_idx[e] = c;
                ++e;
            }
    }
    row_ptr[m] = e;
    return e;
}

/* IsCodeword over the edge list: every check has even parity. */
int orc_syndrome_ok(int m, const int *row_ptr, const int *col_idx, const uint8_t *bits) {
    for (int r = 0; r < m; ++r) {
        int parity = 0;
        for (int e = row_ptr[r]; e < row_ptr[r + 1]; ++e) parity ^= bits[col_idx[e]] & 1;
        if (parity) return 0;
    }
    return 1;
}

/* var -> list of edge ids (edge ids follow CSR order, so rows ascend). */
static void build_csc(int m, int n, const int *row_ptr, const int *col_idx, int *col_ptr, int *edge_of) {
    int E = row_ptr[m];
    memset(col_ptr, 0, sizeof(int) * (n + 1));
    for (int e = 0; e < E; ++e) col_ptr[col_idx[e] + 1]++;
    for (int v = 0; v < n; ++v) col_ptr[v + 1] += col_ptr[v];
    int *fill = (int *) malloc(sizeof(int) * (n + 1));
    memcpy(fill, col_ptr, sizeof(int) * (n + 1));
    for (int e = 0; e < E; ++e) edge_of[fill[col_idx[e]]++] = e;
    free(fill);
}

/* ---------------------------------------------------- belief propagation (BP) */

/* Generated twice: long double (x87 fp80, what bp.h computes in) and double.
 * Flooding sum-product in the phi domain.
 *   ok      1 when the syndrome vanished within max_iter iterations
 *   bits    n hard decisions (zero-filled when ok == 0: the reference returns
 *           an EMPTY codeword then, bp.h:198)
 *   iters   iterations executed (1-based index of the converging iteration,
 *           max_iter on failure)
 *   post    posterior LLR (estimate(), bp.h:85-90) of the last iteration, as double
 * early_exit == 0 is the fixed-iteration measurement mode (not in the reference):
 * all max_iter iterations run, bits/ok describe the last one. */
#define ORC_DEFINE_BP(NAME, REAL, TANH, LOG, FABS)                                                     \
    static REAL NAME##_phi(REAL x) { return -LOG(TANH(x / 2)); }                                       \
    int NAME(int m, int n, const int *row_ptr, const int *col_idx, const double *y, double snr,        \
             int max_iter, int early_exit, uint8_t *bits, int *iters, double *post) {                  \
        int E = row_ptr[m];                                                                            \
        int *col_ptr = (int *) malloc(sizeof(int) * (n + 1));                                          \
        int *edge_of = (int *) malloc(sizeof(int) * (E > 0 ? E : 1));                                  \
        REAL *llr = (REAL *) malloc(sizeof(REAL) * n);                                                 \
        REAL *c2v = (REAL *) calloc(E > 0 ? E : 1, sizeof(REAL));                                      \
        REAL *mag = (REAL *) malloc(sizeof(REAL) * (E > 0 ? E : 1));                                   \
        REAL *sgn = (REAL *) malloc(sizeof(REAL) * (E > 0 ? E : 1));                                   \
        build_csc(m, n, row_ptr, col_idx, col_ptr, edge_of);                                           \
        for (int v = 0; v < n; ++v) llr[v] = (REAL) orc_llr(y[v], snr);                                \
        int ok = 0, it = 0;                                                                            \
        for (int pass = 0; pass <= max_iter; ++pass) {                                                 \
            if (pass > 0) { /* check -> variable, bp.h:49-57 */                                        \
                for (int c = 0; c < m; ++c)                                                            \
                    for (int e = row_ptr[c]; e < row_ptr[c + 1]; ++e) {                                \
                        REAL sum = 0, s = 1;                                                           \
                        for (int o = row_ptr[c]; o < row_ptr[c + 1]; ++o)                              \
                            if (o != e) {                                                              \
                                sum += mag[o];                                                         \
                                s *= sgn[o];                                                           \
                            }                                                                          \
                        c2v[e] = s * NAME##_phi(sum);                                                  \
                    }                                                                                  \
            }                                                                                          \
            /* variable -> check, bp.h:77-83 (pass 0 = the initial send, bp.h:184) */                  \
            for (int v = 0; v < n; ++v)                                                                \
                for (int k = col_ptr[v]; k < col_ptr[v + 1]; ++k) {                                    \
                    REAL sum = 0;                                                                      \
                    for (int o = col_ptr[v]; o < col_ptr[v + 1]; ++o)                                  \
                        if (o != k) sum += c2v[edge_of[o]];                                            \
                    REAL t = llr[v] + sum;                                                             \
                    mag[edge_of[k]] = NAME##_phi(FABS(t));                                             \
                    sgn[edge_of[k]] = (t <= 0) ? -1 : 1;                                               \
                }                                                                                      \
            if (pass == 0) continue;                                                                   \
            it = pass;                                                                                 \
            for (int v = 0; v < n; ++v) { /* estimate + decision, bp.h:85-90, :191-193 */              \
                REAL sum = 0;                                                                          \
                for (int k = col_ptr[v]; k < col_ptr[v + 1]; ++k) sum += c2v[edge_of[k]];              \
                REAL est = llr[v] + sum;                                                               \
                bits[v] = (est <= 0) ? 1 : 0;                                                          \
                if (post) post[v] = (double) est;                                                      \
            }                                                                                          \
            ok = orc_syndrome_ok(m, row_ptr, col_idx, bits);                                           \
            if (ok && early_exit) break;                                                               \
        }                                                                                              \
        if (!ok) memset(bits, 0, n);                                                                   \
        if (iters) *iters = it;                                                                        \
        free(col_ptr); free(edge_of); free(llr); free(c2v); free(mag); free(sgn);                      \
        return ok;                                                                                     \
    }

ORC_DEFINE_BP(orc_bp_decode_fp80, long double, tanhl, logl, fabsl)
ORC_DEFINE_BP(orc_bp_decode_fp64, double, tanh, log, fabs)

/* ------------------------------------------------------------------- QP-ADMM */

/* The penalised-QP ADMM problem of qp_admm.h:6-11, flattened.
 *   columns (what the reference calls A[i]):  entries col_ptr[i]..col_ptr[i+1],
 *       each (col_row, col_cf), rows ascending -- the v-update gathers in this order
 *   b[row], e[i] = sum of squared coefficients of column i
 * Sizes: n_var = n + n_aux, R rows, nnz entries.  Call with all output pointers
 * NULL to obtain the sizes, then again with buffers. */
int orc_admm_build(int m, int n, const int *row_ptr, const int *col_idx, int *n_var_out, int *R_out,
                   int *nnz_out, int *col_ptr, int *col_row, double *col_cf, double *b, double *e) {
    /* pass 1: sizes (qp_admm.h:14-22, :59-92) */
    int n_aux = 0, R = 0, nnz = 0;
    for (int c = 0; c < m; ++c) {
        int d = row_ptr[c + 1] - row_ptr[c];
        if (d > 3) n_aux += d - 3;
        if (d == 1) { R += 1; nnz += 1; }
        else if (d == 2) { R += 2; nnz += 4; }
        else if (d >= 3) { R += 4 * (d - 2); nnz += 12 * (d - 2); }
    }
    int n_var = n + n_aux;
    if (n_var_out) *n_var_out = n_var;
    if (R_out) *R_out = R;
    if (nnz_out) *nnz_out = nnz;
    if (!col_ptr) return 0;

    /* pass 2: emit rows in the reference's order as (row, var, cf) triples */
    int *t_row = (int *) malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
    int *t_var = (int *) malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
    double *t_cf = (double *) malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
    int k = 0, row = 0, next_aux = n;
#define EMIT(r, v, c) do { t_row[k] = (r); t_var[k] = (v); t_cf[k] = (c); ++k; } while (0)
    for (int c = 0; c < m; ++c) {
        const int *idx = col_idx + row_ptr[c];
        int d = row_ptr[c + 1] - row_ptr[c];
        if (d == 0) continue;                       /* qp_admm.h:67-69 */
        if (d == 1) {                               /* :70-74  x <= 0 */
            EMIT(row, idx[0], 1.0);
            b[row] = 0.0;
            row += 1;
            continue;
        }
        if (d == 2) {                               /* :75-83  x0 - x1 <= 0, x1 - x0 <= 0 */
            b[row] = 0.0; b[row + 1] = 0.0;
            EMIT(row, idx[0], 1.0); EMIT(row + 1, idx[0], -1.0);
            EMIT(row, idx[1], -1.0); EMIT(row + 1, idx[1], 1.0);
            row += 2;
            continue;
        }
        int last = idx[0];                          /* :84-91 chain of three-variable checks */
        for (int j = 1; j <= d - 2; ++j) {
            int mid = idx[j];
            int third = (j == d - 2) ? idx[d - 1] : next_aux++;
            int tri[3] = {last, mid, third};
            /* add_three, :34-57: rows (+--;0) (-+-;0) (--+;0) (+++;2) */
            b[row] = 0.0; b[row + 1] = 0.0; b[row + 2] = 0.0; b[row + 3] = 2.0;
            for (int p = 0; p < 3; ++p)
                for (int q = 0; q < 4; ++q) EMIT(row + q, tri[p], (q == 3 || q == p) ? 1.0 : -1.0);
            row += 4;
            last = third;
        }
    }
#undef EMIT
    /* columns: stable counting sort by variable keeps the emission (= row) order,
     * which is what A[var].emplace_back produces */
    memset(col_ptr, 0, sizeof(int) * (n_var + 1));
    for (int i = 0; i < nnz; ++i) col_ptr[t_var[i] + 1]++;
    for (int v = 0; v < n_var; ++v) col_ptr[v + 1] += col_ptr[v];
    int *fill = (int *) malloc(sizeof(int) * (n_var + 1));
    memcpy(fill, col_ptr, sizeof(int) * (n_var + 1));
    for (int i = 0; i < nnz; ++i) {
        int p = fill[t_var[i]]++;
        col_row[p] = t_row[i];
        col_cf[p] = t_cf[i];
    }
    for (int v = 0; v < n_var; ++v) {               /* :94-99 */
        e[v] = 0.0;
        for (int p = col_ptr[v]; p < col_ptr[v + 1]; ++p) e[v] += col_cf[p] * col_cf[p];
    }
    free(fill); free(t_row); free(t_var); free(t_cf);
    return 0;
}

/* DecodeQPADMM, qp_admm.h:104-178.  Returns the decoder's bool (0 only for the
 * infeasible-parameter exit, :108-114, where bits are all zero).
 *   iters  iterations executed (0 for the infeasible exit)
 *   v_out  n doubles: the relaxed solution v[0..n) at exit (NULL to skip) */
int orc_qpadmm_decode(int m, int n, const int *row_ptr, const int *col_idx, const double *y, double snr,
                      double alpha, double mu, int max_iter, double eps_stop, uint8_t *bits, int *iters,
                      double *v_out) {
    int n_var, R, nnz;
    orc_admm_build(m, n, row_ptr, col_idx, &n_var, &R, &nnz, NULL, NULL, NULL, NULL, NULL);
    int *col_ptr = (int *) malloc(sizeof(int) * (n_var + 1));
    int *col_row = (int *) malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
    double *col_cf = (double *) malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
    double *b = (double *) malloc(sizeof(double) * (R > 0 ? R : 1));
    double *e = (double *) malloc(sizeof(double) * n_var);
    orc_admm_build(m, n, row_ptr, col_idx, &n_var, &R, &nnz, col_ptr, col_row, col_cf, b, e);

    double *q = (double *) calloc(n_var, sizeof(double));
    for (int i = 0; i < n; ++i) q[i] = orc_llr(y[i], snr);

    int ok = 1, it = 0;
    double e_min = 1e9;
    for (int i = 0; i < n_var; ++i) e_min = (e[i] < e_min) ? e[i] : e_min;
    double *v = (double *) calloc(n_var, sizeof(double));
    if (e_min * mu <= alpha) {
        memset(bits, 0, n);
        ok = 0;
    } else {
        double *z = (double *) calloc(R > 0 ? R : 1, sizeof(double));
        double *yl = (double *) calloc(R > 0 ? R : 1, sizeof(double));
        double *r = (double *) malloc(sizeof(double) * (R > 0 ? R : 1));
        double *inv_coef = (double *) malloc(sizeof(double) * n_var);
        for (int i = 0; i < n_var; ++i) v[i] = q[i] > 0.0 ? 1.0 : 0.0;   /* :116-119, visible only if max_iter == 0 */
        for (int i = 0; i < n_var; ++i) {
            double A = (mu * e[i] - alpha) / 2;
            inv_coef[i] = -1.0 / (2 * A);
        }
        for (int iter = 0; iter < max_iter; ++iter) {
            it = iter + 1;
            for (int i = 0; i < n_var; ++i) {                 /* :132-142 */
                double B = q[i] + (alpha / 2);
                for (int p = col_ptr[i]; p < col_ptr[i + 1]; ++p) {
                    int j = col_row[p];
                    B += col_cf[p] * (yl[j] + mu * (z[j] - b[j]));
                }
                double x = B * inv_coef[i];
                x = (x < 0.0) ? 0.0 : x;                      /* std::max(v, 0.0) */
                x = (1.0 < x) ? 1.0 : x;                      /* std::min(v, 1.0) */
                v[i] = x;
            }
            for (int j = 0; j < R; ++j) r[j] = b[j];          /* :144-151 */
            for (int i = 0; i < n_var; ++i)
                for (int p = col_ptr[i]; p < col_ptr[i + 1]; ++p) r[col_row[p]] -= col_cf[p] * v[i];
            double sum2 = 0;                                   /* :154-159 */
            for (int j = 0; j < R; ++j) {
                double t = r[j] - yl[j];
                z[j] = (0.0 < t) ? t : 0.0;                   /* std::max(0.0, r - yl) */
                double u = yl[j] - r[j];
                yl[j] = (0.0 < u) ? u : 0.0;
                sum2 += (z[j] - r[j]) * (z[j] - r[j]);
            }
            if (sum2 < eps_stop) break;                        /* :161-163 */
        }
        for (int i = 0; i < n; ++i) bits[i] = (v[i] <= 0.5) ? 0 : 1;   /* :166-177 */
        free(z); free(yl); free(r); free(inv_coef);
    }
    if (v_out) memcpy(v_out, v, sizeof(double) * n);
    if (iters) *iters = it;
    free(v); free(q); free(col_ptr); free(col_row); free(col_cf); free(b); free(e);
    return ok;
}

/* -------------------------------------------- counter-based channel (Philox) */

/* Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11): ten rounds, multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key increments
 * 0x9E3779B9 / 0xBB67AE85. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t) 0xD2511F53u * c0;
        uint64_t p1 = (uint64_t) 0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t) p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 52 random bits -> (0,1), never 0 or 1:  (a + 1/2) * 2^-52 */
static double unit_open(uint32_t hi, uint32_t lo) {
    uint64_t a = ((uint64_t) hi << 20) | (lo >> 12);
    return ((double) a + 0.5) * 0x1p-52;
}

/* ln(u) for u in (0,1), from correctly rounded operations only. */
static double det_log(double u) {
    uint64_t bits;
    memcpy(&bits, &u, 8);
    int ex = (int) ((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double mant;
    memcpy(&mant, &bits, 8);                 /* [1, 2) */
    if (mant > 1.4142135623730951) {
        mant = mant * 0.5;
        ex += 1;
    }
    double f = mant - 1.0;
    double s = f / (2.0 + f);
    double s2 = s * s;
    double p = 0x1.642c8590b2164p-5;         /* 1/23 */
    p = fma(p, s2, 0x1.8618618618618p-5);    /* 1/21 */
    p = fma(p, s2, 0x1.af286bca1af28p-5);    /* 1/19 */
    p = fma(p, s2, 0x1.e1e1e1e1e1e1ep-5);    /* 1/17 */
    p = fma(p, s2, 0x1.1111111111111p-4);    /* 1/15 */
    p = fma(p, s2, 0x1.3b13b13b13b14p-4);    /* 1/13 */
    p = fma(p, s2, 0x1.745d1745d1746p-4);    /* 1/11 */
    p = fma(p, s2, 0x1.c71c71c71c71cp-4);    /* 1/9 */
    p = fma(p, s2, 0x1.2492492492492p-3);    /* 1/7 */
    p = fma(p, s2, 0x1.999999999999ap-3);    /* 1/5 */
    p = fma(p, s2, 0x1.5555555555555p-2);    /* 1/3 */
    double two_s = s + s;
    double lm = fma(two_s, p * s2, two_s);   /* ln(mant) = 2s (1 + s2 p) */
    return fma((double) ex, 0x1.62e42fefa39efp-1, lm);
}

/* sin and cos of 2*pi*u, u in (0,1), from correctly rounded operations only. */
static void det_sincos2pi(double u, double *sn, double *cs) {
    double t = u * 4.0;
    int quad = (int) (t + 0.5);              /* 0..4, truncation == floor here */
    double f = t - (double) quad;            /* [-1/2, 1/2], exact */
    double x = f * 0x1.921fb54442d18p+0;     /* * pi/2 */
    double x2 = x * x;
    double ps = -0x1.2f49b46814157p-57;      /* -1/19! */
    ps = fma(ps, x2, 0x1.952c77030ad4ap-49); /*  1/17! */
    ps = fma(ps, x2, -0x1.ae7f3e733b81fp-41);/* -1/15! */
    ps = fma(ps, x2, 0x1.6124613a86d09p-33); /*  1/13! */
    ps = fma(ps, x2, -0x1.ae64567f544e4p-26);/* -1/11! */
    ps = fma(ps, x2, 0x1.71de3a556c734p-19); /*  1/9!  */
    ps = fma(ps, x2, -0x1.a01a01a01a01ap-13);/* -1/7!  */
    ps = fma(ps, x2, 0x1.1111111111111p-7);  /*  1/5!  */
    ps = fma(ps, x2, -0x1.5555555555555p-3); /* -1/3!  */
    double s = fma(x * x2, ps, x);
    double pc = 0x1.6827863b97d97p-53;       /*  1/18! */
    pc = fma(pc, x2, -0x1.ae7f3e733b81fp-45);/* -1/16! */
    pc = fma(pc, x2, 0x1.93974a8c07c9dp-37); /*  1/14! */
    pc = fma(pc, x2, -0x1.1eed8eff8d898p-29);/* -1/12! */
    pc = fma(pc, x2, 0x1.27e4fb7789f5cp-22); /*  1/10! */
    pc = fma(pc, x2, -0x1.a01a01a01a01ap-16);/* -1/8!  */
    pc = fma(pc, x2, 0x1.6c16c16c16c17p-10); /*  1/6!  */
    pc = fma(pc, x2, -0x1.5555555555555p-5); /* -1/4!  */
    pc = fma(pc, x2, 0x1p-1);                /*  1/2!  */
    double c = fma(-x2, pc, 1.0);
    switch (quad & 3) {
        case 0: *sn = s;  *cs = c;  break;
        case 1: *sn = c;  *cs = -s; break;
        case 2: *sn = -s; *cs = -c; break;
        default: *sn = -c; *cs = s; break;
    }
}

/* Two standard normals from one Philox block (Box-Muller):
 * z0 = r cos(2 pi u2), z1 = r sin(2 pi u2), r = sqrt(-2 ln u1). */
void orc_gauss_pair(const uint32_t w[4], double *z0, double *z1) {
    double u1 = unit_open(w[0], w[1]);
    double u2 = unit_open(w[2], w[3]);
    double r = sqrt(-2.0 * det_log(u1));
    double sn, cs;
    det_sincos2pi(u2, &sn, &cs);
    *z0 = r * cs;
    *z1 = r * sn;
}

/* Counter layout: key = seed (lo, hi); counter = (frame lo,
 * frame hi, block index, stream).  Stream 0 = information bits, stream 1 = noise. */
enum { ORC_STREAM_INFO = 0, ORC_STREAM_NOISE = 1 };

/* y_i = (c_i ? -1 : +1) + sigma * w_i for one frame; sigma = sqrt(llr_variance). */
void orc_channel_frame(uint64_t seed, uint64_t frame, int n, const uint8_t *codeword, double sigma,
                       double *y) {
    uint32_t key[2] = {(uint32_t) seed, (uint32_t) (seed >> 32)};
    for (int blk = 0; 2 * blk < n; ++blk) {
        uint32_t ctr[4] = {(uint32_t) frame, (uint32_t) (frame >> 32), (uint32_t) blk, ORC_STREAM_NOISE};
        uint32_t w[4];
        double z[2];
        orc_philox4x32_10(ctr, key, w);
        orc_gauss_pair(w, &z[0], &z[1]);
        for (int h = 0; h < 2; ++h) {
            int i = 2 * blk + h;
            if (i >= n) break;
            double sym = (codeword && codeword[i]) ? -1.0 : 1.0;
            y[i] = fma(sigma, z[h], sym);
        }
    }
}

/* k information bits of one frame: bit i = bit (i % 32) of word (i / 32) % 4 of
 * block i / 128 on stream 0. */
void orc_info_bits(uint64_t seed, uint64_t frame, int k, uint8_t *u) {
    uint32_t key[2] = {(uint32_t) seed, (uint32_t) (seed >> 32)};
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < k; ++i) {
        if (i % 128 == 0) {
            uint32_t ctr[4] = {(uint32_t) frame, (uint32_t) (frame >> 32), (uint32_t) (i / 128), ORC_STREAM_INFO};
            orc_philox4x32_10(ctr, key, w);
        }
        u[i] = (w[(i / 32) % 4] >> (i % 32)) & 1u;
    }
}

/* c = u * G over GF(2), G dense k x n bytes (rows of GetOrtogonal's output). */
void orc_encode(const uint8_t *G, int k, int n, const uint8_t *u, uint8_t *c) {
    memset(c, 0, n);
    for (int i = 0; i < k; ++i)
        if (u[i])
            for (int j = 0; j < n; ++j) c[j] ^= G[(size_t) i * n + j] & 1;
}

/* ------------------------------------------------------------- experiment */

/* Counter block of one (decoder, SNR) point.  The first seven are the
 * reference's ExperimentResult/HammingDistanceTracker fields (experiment.h:25-68)
 * widened to 64 bit; the rest are extensions named in SURVEY.md 8a-a15 / 8e. */
enum {
    ORC_CNT_TOTAL = 0, ORC_CNT_CORRECT, ORC_CNT_PSEUDO, ORC_CNT_DECODER_FAIL, ORC_CNT_BIT_ERRORS,
    ORC_CNT_SUM_HAMMING, ORC_CNT_SUM_HAMMING_OK, ORC_CNT_SUM_HAMMING_WRONG, ORC_CNT_SUM_ITERS,
    ORC_CNT_FRAMES_WITH_BITS, ORC_CNT_COUNT
};

/* Classify one decoded frame and fold it into counters (experiment.h:109-120).
 * ok = decoder bool; bits = its n hard decisions (ignored when the decoder
 * returned an empty word, i.e. BP failure: pass has_bits = 0). */
void orc_account_frame(int m, int n, const int *row_ptr, const int *col_idx, const uint8_t *codeword,
                       const double *y, int ok, int has_bits, const uint8_t *bits, int iters,
                       uint64_t *cnt) {
    int correct = 0;
    if (ok && has_bits && orc_syndrome_ok(m, row_ptr, col_idx, bits)) {
        if (memcmp(bits, codeword, n) == 0) {
            cnt[ORC_CNT_CORRECT]++;
            correct = 1;
        } else
            cnt[ORC_CNT_PSEUDO]++;
    }
    if (!ok) cnt[ORC_CNT_DECODER_FAIL]++;
    if (has_bits) {
        cnt[ORC_CNT_FRAMES_WITH_BITS]++;
        for (int i = 0; i < n; ++i) cnt[ORC_CNT_BIT_ERRORS] += (bits[i] != codeword[i]);
    }
    cnt[ORC_CNT_TOTAL]++;
    int hamming = 0;
    for (int i = 0; i < n; ++i) {
        if (!codeword[i] && y[i] <= 0) hamming++;
        if (codeword[i] && y[i] > 0) hamming++;
    }
    cnt[ORC_CNT_SUM_HAMMING] += hamming;
    if (correct) cnt[ORC_CNT_SUM_HAMMING_OK] += hamming;
    else cnt[ORC_CNT_SUM_HAMMING_WRONG] += hamming;
    cnt[ORC_CNT_SUM_ITERS] += iters;
}

/* Monte-Carlo point over global frame indices [frame_begin, frame_begin+count).
 * algo: 0 = BP (fp80), 1 = QP-ADMM.
 * codeword source: G != NULL -> c = info_bits(frame) * G (k rows);
 *                  else words != NULL -> c = words[frame % n_words];
 *                  else all-zero. */
void orc_experiment(int algo, int m, int n, const int *row_ptr, const int *col_idx, double snr, int max_iter,
                    int early_exit, double alpha, double mu, double eps_stop, uint64_t seed,
                    uint64_t frame_begin, uint64_t count, const uint8_t *G, int k, const uint8_t *words,
                    uint64_t n_words, uint64_t *cnt) {
    double sigma = sqrt(orc_llr_variance(snr));
    uint8_t *c = (uint8_t *) calloc(n, 1);
    uint8_t *u = (uint8_t *) calloc(k > 0 ? k : 1, 1);
    uint8_t *bits = (uint8_t *) calloc(n, 1);
    double *y = (double *) malloc(sizeof(double) * n);
    memset(cnt, 0, sizeof(uint64_t) * ORC_CNT_COUNT);
    for (uint64_t f = frame_begin; f < frame_begin + count; ++f) {
        if (G) {
            orc_info_bits(seed, f, k, u);
            orc_encode(G, k, n, u, c);
        } else if (words)
            memcpy(c, words + (size_t) (f % n_words) * n, n);
        orc_channel_frame(seed, f, n, c, sigma, y);
        int iters = 0, ok, has_bits = 1;
        if (algo == 0) {
            ok = orc_bp_decode_fp80(m, n, row_ptr, col_idx, y, snr, max_iter, early_exit, bits, &iters, NULL);
            has_bits = ok;
        } else
            ok = orc_qpadmm_decode(m, n, row_ptr, col_idx, y, snr, alpha, mu, max_iter, eps_stop, bits, &iters,
                                   NULL);
        orc_account_frame(m, n, row_ptr, col_idx, c, y, ok, has_bits, bits, iters, cnt);
    }
    free(c); free(u); free(bits); free(y);
}
