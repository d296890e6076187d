// GF(2) vectors and matrices with the reference's public types and functions
// (utils/codeword.h:17-128 of GreatDrake/acg-alp-ldpc): TCodeword, TMatrix,
// operator^ / & / *, IsCodeword, GetOrtogonal.  Same results, different engine:
// everything is done on 64-bit packed rows instead of vector<bool> element loops.
#ifndef LDPC_B200_UTILS_CODEWORD_H
#define LDPC_B200_UTILS_CODEWORD_H

#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

using namespace std;

typedef vector<bool> TCodeword;
typedef vector<TCodeword> TMatrix;

namespace gf2 {

typedef vector<uint64_t> Packed;

inline Packed pack(const TCodeword &v) {
    Packed p((v.size() + 63) / 64, 0);
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i]) p[i >> 6] |= uint64_t(1) << (i & 63);
    return p;
}

inline TCodeword unpack(const Packed &p, size_t bits) {
    TCodeword v(bits);
    for (size_t i = 0; i < bits; ++i) v[i] = (p[i >> 6] >> (i & 63)) & 1;
    return v;
}

inline bool dot(const Packed &a, const Packed &b) {
    uint64_t acc = 0;
    for (size_t w = 0; w < a.size(); ++w) acc ^= a[w] & b[w];
    return __builtin_parityll(acc);
}

inline void xor_into(Packed &dst, const Packed &src) {
    for (size_t w = 0; w < dst.size(); ++w) dst[w] ^= src[w];
}

inline bool get(const Packed &p, size_t i) { return (p[i >> 6] >> (i & 63)) & 1; }

}  // namespace gf2

// "0101..." text form (utils/codeword.h:20-35): any character other than '0' reads as 1
inline istream &operator>>(istream &in, TCodeword &word) {
    string text;
    in >> text;
    word.assign(text.size(), false);
    for (size_t i = 0; i < text.size(); ++i) word[i] = text[i] != '0';
    return in;
}

inline ostream &operator<<(ostream &out, const TCodeword &word) {
    for (bool bit : word) out << bit;
    return out;
}

template <typename T>
ostream &operator<<(ostream &out, const vector<T> &items) {
    for (const T &item : items) out << item << "\n";
    return out;
}

inline TCodeword operator^(const TCodeword &a, const TCodeword &b) {
    assert(a.size() == b.size());
    gf2::Packed pa = gf2::pack(a);
    gf2::xor_into(pa, gf2::pack(b));
    return gf2::unpack(pa, a.size());
}

inline TCodeword operator&(const TCodeword &a, const TCodeword &b) {
    assert(a.size() == b.size());
    gf2::Packed pa = gf2::pack(a), pb = gf2::pack(b);
    for (size_t w = 0; w < pa.size(); ++w) pa[w] &= pb[w];
    return gf2::unpack(pa, a.size());
}

// matrix product over GF(2)
inline TMatrix operator*(const TMatrix &a, const TMatrix &b) {
    assert(!a.empty() && a[0].size() == b.size());
    const size_t inner = b.size(), cols = b.empty() ? 0 : b[0].size();
    vector<gf2::Packed> b_cols(cols, gf2::Packed((inner + 63) / 64, 0));
    for (size_t k = 0; k < inner; ++k)
        for (size_t j = 0; j < cols; ++j)
            if (b[k][j]) b_cols[j][k >> 6] |= uint64_t(1) << (k & 63);
    TMatrix c(a.size(), TCodeword(cols, false));
    for (size_t i = 0; i < a.size(); ++i) {
        const gf2::Packed row = gf2::pack(a[i]);
        for (size_t j = 0; j < cols; ++j) c[i][j] = gf2::dot(row, b_cols[j]);
    }
    return c;
}

// H * v: the syndrome
inline TCodeword operator*(const TMatrix &H, const TCodeword &v) {
    const gf2::Packed pv = gf2::pack(v);
    TCodeword s(H.size());
    for (size_t i = 0; i < H.size(); ++i) {
        assert(H[i].size() == v.size());
        s[i] = gf2::dot(gf2::pack(H[i]), pv);
    }
    return s;
}

// v * M: a combination of the rows of M
inline TCodeword operator*(const TCodeword &v, const TMatrix &M) {
    assert(v.size() == M.size());
    const size_t cols = M.empty() ? 0 : M[0].size();
    gf2::Packed acc((cols + 63) / 64, 0);
    for (size_t i = 0; i < M.size(); ++i)
        if (v[i]) gf2::xor_into(acc, gf2::pack(M[i]));
    return gf2::unpack(acc, cols);
}

// utils/codeword.h:90-95
inline bool IsCodeword(const vector<TCodeword> &H, const TCodeword &c) {
    const gf2::Packed pc = gf2::pack(c);
    for (const TCodeword &row : H)
        if (gf2::dot(gf2::pack(row), pc)) return false;
    return true;
}

// Generator of the null space of H (utils/codeword.h:97-128).  Row i is reduced
// on its FIRST set column (in row order, no row swaps -- that choice defines which
// basis comes out), every other row is cleared on that column, and one generator
// row is emitted per free column, in ascending column order.  {empty, false} when
// a row reduces to zero (H not of full row rank).
inline pair<TMatrix, bool> GetOrtogonal(TMatrix H) {
    const size_t m = H.size(), n = H.empty() ? 0 : H[0].size();
    vector<gf2::Packed> rows(m);
    for (size_t i = 0; i < m; ++i) rows[i] = gf2::pack(H[i]);
    vector<int> pivot(m, -1);
    vector<char> is_pivot(n, 0);
    for (size_t i = 0; i < m; ++i) {
        for (size_t w = 0; w < rows[i].size() && pivot[i] < 0; ++w)
            if (rows[i][w]) pivot[i] = int(w * 64 + __builtin_ctzll(rows[i][w]));
        if (pivot[i] < 0) return {TMatrix(), false};
        for (size_t k = 0; k < m; ++k)
            if (k != i && gf2::get(rows[k], pivot[i])) gf2::xor_into(rows[k], rows[i]);
        is_pivot[pivot[i]] = 1;
    }
    TMatrix G;
    G.reserve(n - m);
    for (size_t j = 0; j < n; ++j) {
        if (is_pivot[j]) continue;
        TCodeword g(n, false);
        g[j] = true;
        for (size_t i = 0; i < m; ++i)
            if (gf2::get(rows[i], j)) g[pivot[i]] = true;
        G.push_back(g);
    }
    return {G, true};
}

#endif
