// QP-ADMM problem compiler: ConstructADMMProblem (algo/qp_admm.h:13-102) once per
// H, expressed as BLOCKS (three-variable checks of the chain decomposition, or the
// degree-1 / degree-2 special cases) and laid out for the B200's shared memory.
//
// The kernel stores, per frame, the row terms w of block rank R at 16-byte chunk R of
// two arrays (w01, w23) and the variable values v of variable rank R at double R.
// Its two gathers are static: the v-update reads the chunks of the blocks a
// variable belongs to, the residual update reads the v of a block's variables.
// A quarter-warp (8 lanes x 16 B) is conflict-free iff its 8 chunk indices differ
// mod 8; a half-warp (16 lanes x 8 B) iff its 16 double indices differ mod 16.
// Lanes handle consecutive ranks, so the RANKS decide the bank conflicts.  They are
// chosen here by a deterministic annealing pass that swaps ranks inside a class
// (variables of equal degree, blocks of equal slot pattern -- so warps stay
// uniform) to minimise the number of replayed wavefronts.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <numeric>

#include "ldpc_internal.h"

namespace ldpc {

namespace {

struct RawBlock {
    int var[3];   // by slot: (last, middle, third) of add_three, qp_admm.h:34-57; -1 = absent
    int nvars, rows;
    int chain_pos;   // position of the block in its check's chain (0 = first)
};

struct Layout {
    int nv = 0, nb = 0;
    std::vector<std::vector<int>> var_blocks;   // per variable: blocks in row order (the gather order)
    std::vector<std::array<int, 3>> blk_vars;   // per block: variables by ASCENDING index (-1 = absent)
    std::vector<int> var_class, blk_class;
    std::vector<int> rank_v, rank_b;            // id -> rank
    std::vector<int> at_v, at_b;                // rank -> id

    // The search minimises the number of clashing PAIRS (a smooth surrogate); the figure of merit
    // reported to the caller is the number of replayed wavefronts (max multiplicity - 1 per access).
    // w gather of variable ranks [8g, 8g+8): one access per step of the incidence lists
    int cost_vgroup(int g, bool pairs = true) const {
        int cost = 0;
        for (int step = 0;; ++step) {
            int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, seen[8], any = 0, worst = 1, clash = 0;
            for (int i = 0; i < 8; ++i) seen[i] = -1;
            for (int r = 8 * g; r < std::min(8 * g + 8, nv); ++r) {
                const std::vector<int> &bl = var_blocks[at_v[r]];
                if (step < (int) bl.size()) {
                    any = 1;
                    const int bank = rank_b[bl[step]] & 7;
                    if (seen[bank] == bl[step]) continue;     // same chunk: broadcast, not a conflict
                    seen[bank] = bl[step];
                    clash += cnt[bank];
                    worst = std::max(worst, ++cnt[bank]);
                }
            }
            if (!any) break;
            cost += pairs ? clash : worst - 1;
        }
        return cost;
    }
    // v gather of block ranks [16h, 16h+16): one access per visited variable
    int cost_bgroup(int h, bool pairs = true) const {
        int cost = 0;
        for (int k = 0; k < 3; ++k) {
            int cnt[16] = {0}, worst = 1, clash = 0;
            int seen[16];
            for (int i = 0; i < 16; ++i) seen[i] = -1;
            for (int r = 16 * h; r < std::min(16 * h + 16, nb); ++r) {
                const int v = blk_vars[at_b[r]][k];
                if (v < 0) continue;
                const int bank = rank_v[v] & 15;
                if (seen[bank] == v) continue;    // same address: broadcast, not a conflict
                seen[bank] = v;
                clash += cnt[bank];
                worst = std::max(worst, ++cnt[bank]);
            }
            cost += pairs ? clash : worst - 1;
        }
        return cost;
    }
    long replayed_wavefronts() const {
        long c = 0;
        for (int g = 0; 8 * g < nv; ++g) c += cost_vgroup(g, false);
        for (int h = 0; 16 * h < nb; ++h) c += cost_bgroup(h, false);
        return c;
    }
};

// xorshift: deterministic across platforms (the layout must not depend on libstdc++)
struct Rng {
    uint64_t s;
    uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t) (s >> 11); }
    double unit() { return (next() & 0xffffff) / 16777216.0; }
};

void anneal(Layout &L, int moves) {
    Rng rng{0x9E3779B97F4A7C15ull};
    std::vector<int> touched;
    auto eval = [&](const std::vector<int> &vg, const std::vector<int> &bg) {
        int c = 0;
        for (int g : vg) c += L.cost_vgroup(g);
        for (int h : bg) c += L.cost_bgroup(h);
        return c;
    };
    auto uniq = [](std::vector<int> &x) { std::sort(x.begin(), x.end()); x.erase(std::unique(x.begin(), x.end()), x.end()); };
    for (int it = 0; it < moves; ++it) {
        const double temp = 0.6 * std::pow(0.02, (double) it / moves);
        std::vector<int> vg, bg;
        const bool swap_vars = (rng.next() & 1) || L.nb < 2;
        if (swap_vars) {
            if (L.nv < 2) continue;
            const int a = L.at_v[rng.next() % L.nv], b = L.at_v[rng.next() % L.nv];
            if (a == b || L.var_class[a] != L.var_class[b]) continue;
            vg = {L.rank_v[a] / 8, L.rank_v[b] / 8};
            for (int x : {a, b})
                for (int blk : L.var_blocks[x]) bg.push_back(L.rank_b[blk] / 16);
            uniq(vg); uniq(bg);
            const int before = eval(vg, bg);
            std::swap(L.rank_v[a], L.rank_v[b]);
            L.at_v[L.rank_v[a]] = a; L.at_v[L.rank_v[b]] = b;
            const int delta = eval(vg, bg) - before;
            if (delta > 0 && rng.unit() >= std::exp(-delta / temp)) {
                std::swap(L.rank_v[a], L.rank_v[b]);
                L.at_v[L.rank_v[a]] = a; L.at_v[L.rank_v[b]] = b;
            }
        } else {
            const int a = L.at_b[rng.next() % L.nb], b = L.at_b[rng.next() % L.nb];
            if (a == b || L.blk_class[a] != L.blk_class[b]) continue;
            bg = {L.rank_b[a] / 16, L.rank_b[b] / 16};
            for (int x : {a, b})
                for (int v : L.blk_vars[x])
                    if (v >= 0) vg.push_back(L.rank_v[v] / 8);
            uniq(vg); uniq(bg);
            const int before = eval(vg, bg);
            std::swap(L.rank_b[a], L.rank_b[b]);
            L.at_b[L.rank_b[a]] = a; L.at_b[L.rank_b[b]] = b;
            const int delta = eval(vg, bg) - before;
            if (delta > 0 && rng.unit() >= std::exp(-delta / temp)) {
                std::swap(L.rank_b[a], L.rank_b[b]);
                L.at_b[L.rank_b[a]] = a; L.at_b[L.rank_b[b]] = b;
            }
        }
    }
}

}  // namespace

int compile_admm(ldpc_code *c) {
    const int m = c->m, n = c->n;
    // ---- the chain decomposition of qp_admm.h:59-92
    std::vector<RawBlock> raw;
    int next_aux = n;
    for (int r = 0; r < m; ++r) {
        const int *idx = c->col_idx.data() + c->row_ptr[r];
        const int d = c->row_ptr[r + 1] - c->row_ptr[r];
        if (d == 0) continue;                                                    // :67-69
        if (d == 1) { raw.push_back({{idx[0], -1, -1}, 1, 1, 0}); continue; }       // :70-74
        if (d == 2) { raw.push_back({{idx[0], idx[1], -1}, 2, 2, 0}); continue; }   // :75-83
        int last = idx[0];                                                       // :84-91
        for (int j = 1; j <= d - 2; ++j) {
            const int third = (j == d - 2) ? idx[d - 1] : next_aux++;
            raw.push_back({{last, idx[j], third}, 3, 4, j - 1});
            last = third;
        }
    }
    const int nb = c->n_blocks = (int) raw.size();
    const int nv = c->n_var = next_aux;
    if (nv >= 65535 || nb >= 65535) return fail(LDPC_E_UNSUPPORTED, "code too large for the 16-bit QP-ADMM tables");

    Layout L;
    L.nv = nv; L.nb = nb;
    L.var_blocks.assign(nv, {});
    L.blk_vars.assign(nb, {-1, -1, -1});
    std::vector<std::vector<int>> var_slots(nv);     // slot of the variable in each of its blocks
    std::vector<int> e(nv, 0), blk_meta(nb, 0);
    c->n_rows = 0; c->nnz = 0;
    for (int b = 0; b < nb; ++b) {
        const RawBlock &rb = raw[b];
        int order[3] = {0, 1, 2};   // slots by ascending variable index; absent slots last
        std::sort(order, order + 3, [&](int x, int y) {
            const int vx = rb.var[x] < 0 ? 1 << 30 : rb.var[x], vy = rb.var[y] < 0 ? 1 << 30 : rb.var[y];
            return vx < vy || (vx == vy && x < y);
        });
        int meta = rb.rows << 8;
        for (int k = 0; k < 3; ++k) {
            L.blk_vars[b][k] = rb.var[order[k]];
            meta |= order[k] << (2 * k);
        }
        blk_meta[b] = meta;
        for (int slot = 0; slot < rb.nvars; ++slot) {
            L.var_blocks[rb.var[slot]].push_back(b);
            var_slots[rb.var[slot]].push_back(slot);
            e[rb.var[slot]] += rb.rows;               // one +-1 coefficient per row of the block (:94-99)
        }
        c->n_rows += rb.rows;
        c->nnz += rb.rows * rb.nvars;
    }
    // e_min as DecodeQPADMM computes it: over ALL variables, starting from 1e9 (qp_admm.h:108-111)
    c->e_min = 1000000000;
    for (int v = 0; v < nv; ++v) c->e_min = std::min(c->e_min, e[v]);
    c->n_inc = 0;
    for (int v = 0; v < nv; ++v) {
        if (L.var_blocks[v].size() > 255 || e[v] > 65535) return fail(LDPC_E_UNSUPPORTED, "column weight too large");
        c->n_inc += (int) L.var_blocks[v].size();
    }

    // ---- classes and initial ranks: variables by degree (descending), blocks by pattern
    L.var_class.resize(nv); L.blk_class.resize(nb);
    for (int v = 0; v < nv; ++v) L.var_class[v] = (int) L.var_blocks[v].size();
    for (int b = 0; b < nb; ++b) L.blk_class[b] = blk_meta[b] & 0xf3f;
    L.at_v.resize(nv); L.at_b.resize(nb);
    std::iota(L.at_v.begin(), L.at_v.end(), 0);
    std::iota(L.at_b.begin(), L.at_b.end(), 0);
    std::stable_sort(L.at_v.begin(), L.at_v.end(), [&](int a, int b) { return L.var_class[a] > L.var_class[b]; });
    // blocks of one class: all first blocks of the chains, then all second blocks, ... -- neighbouring checks of a
    // quasi-cyclic H touch neighbouring variables, so this keeps consecutive ranks on consecutive banks
    std::stable_sort(L.at_b.begin(), L.at_b.end(), [&](int a, int b) {
        if (L.blk_class[a] != L.blk_class[b]) return L.blk_class[a] < L.blk_class[b];
        return raw[a].chain_pos < raw[b].chain_pos;
    });
    L.rank_v.resize(nv); L.rank_b.resize(nb);
    for (int r = 0; r < nv; ++r) L.rank_v[L.at_v[r]] = r;
    for (int r = 0; r < nb; ++r) L.rank_b[L.at_b[r]] = r;
    c->admm_conflicts_before = L.replayed_wavefronts();
    int moves = 40 * (nv + nb);   // ~40 ms for the 160 x 280 codes; H changes per proposal in optimize_H.cpp
    // Codes the check-centric kernel decodes (qpadmm_chk_kernel.cu: every check of degree <= 12) use this kernel only for
    // the zero-iteration exit of infeasible (alpha, mu): the natural order will do
    bool chk_kernel = true;
    for (int r = 0; r < m && chk_kernel; ++r) {
        const int d = c->row_ptr[r + 1] - c->row_ptr[r];
        chk_kernel = d <= 12;
    }
    for (int v = 0; v < n && chk_kernel; ++v) chk_kernel = L.var_blocks[v].size() <= 15;
    if (chk_kernel) moves = 0;
    if (const char *s = getenv("LDPC_ADMM_LAYOUT_MOVES")) moves = atoi(s);
    if (moves > 0) {
        Layout start = L;
        anneal(L, moves);
        if (L.replayed_wavefronts() > start.replayed_wavefronts()) L = start;   // never worse than the natural order
    }
    c->admm_conflicts_after = L.replayed_wavefronts();

    // ---- device tables in rank order
    std::vector<AdmmVarRec> vrec(nv);
    std::vector<uint32_t> inc;
    std::vector<uint16_t> var_id(nv);
    for (int r = 0; r < nv; ++r) {
        const int v = L.at_v[r];
        var_id[r] = (uint16_t) v;
        vrec[r] = AdmmVarRec{(uint16_t) inc.size(), (uint8_t) L.var_blocks[v].size(), 0, (uint16_t) e[v], 0};
        for (size_t a = 0; a < L.var_blocks[v].size(); ++a) {
            const int slot = var_slots[v][a];
            // bits 31/30/29: flip the sign of row 0/1/2 (coefficient -1 unless the row index equals the slot)
            uint32_t word = (uint32_t) L.rank_b[L.var_blocks[v][a]];
            for (int q = 0; q < 3; ++q)
                if (q != slot) word |= 0x80000000u >> q;
            inc.push_back(word);
        }
    }
    if (inc.size() >= 65535) return fail(LDPC_E_UNSUPPORTED, "too many QP-ADMM incidences for the 16-bit tables");
    std::vector<AdmmBlock> blocks(nb);
    for (int r = 0; r < nb; ++r) {
        const int b = L.at_b[r];
        AdmmBlock ab;
        for (int k = 0; k < 3; ++k) ab.var[k] = (uint16_t) (L.blk_vars[b][k] < 0 ? nv : L.rank_v[L.blk_vars[b][k]]);
        ab.meta = (uint16_t) blk_meta[b];
        blocks[r] = ab;
    }
    int st;
    std::vector<uint16_t> var_rank(n);
    for (int i = 0; i < n; ++i) var_rank[i] = (uint16_t) L.rank_v[i];
    TableStager stage;
    stage.add(&c->d.blocks, blocks);
    stage.add(&c->d.admm_var, vrec);
    stage.add(&c->d.admm_inc, inc);
    stage.add(&c->d.admm_var_id, var_id);
    stage.add(&c->d.admm_var_rank, var_rank);
    if ((st = stage.commit(&c->d.blob_admm))) return st;
    return LDPC_OK;
}

}  // namespace ldpc
